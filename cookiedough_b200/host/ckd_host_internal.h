// ckd_host_internal.h -- what the compositor (ckd_demo.cpp) shares with the effect shims (ckd_host.cpp)
#pragma once

#include "../../include/ckd_host.h"

#include <stddef.h>
#include <stdint.h>
#include <vector>

namespace ckdhost
{
	bool Check(int rc, const char *what);          // false + SetLastError when rc != CKD_OK

	// registered (pre-decoded) image, as handed to CkdHost_RegisterImage; nullptr when the path is unknown
	struct ImageView { const void *pixels; int width, height, bpp; };
	bool FindImage(const char *path, ImageView &view);
	void ReleaseImage(const char *path);           // drop the host copy (after it went to the device)

	// host/ckd_image.cpp: decode `path` (PNG or JPEG, relative to the asset root) to BGRA (bpp 4) or L8 (bpp 1)
	bool DecodeImageFile(const char *path, int bpp, std::vector<uint8_t> &pixels, int &width, int &height);
	// ... and brought to the output resolution by the rules of SURVEY 8 f3 (output-sized art, FX-map sized maps, the ribbon
	// strip; the missing tunnelscape colour map synthesised from the landscape's)
	bool DecodeImageForResolution(const char *path, int bpp, int resX, int resY, std::vector<uint8_t> &pixels, int &width, int &height);

	// While composing, X_Draw(pDest, ...) renders into d_frame and leaves it on the device (pDest is ignored):
	// the compositor downloads the finished frame once.
	uint32_t *BeginCompose();                      // returns the device frame the effects will render into
	void EndCompose(uint32_t *pDest);              // device frame -> pDest (pipelined or synchronous, like X_Draw)
}

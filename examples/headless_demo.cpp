// examples/headless_demo.cpp -- the reference's main loop (code/main.cpp:263-374) without a window, on top of the drop-in
// host layer: what a maintainer of cookiedough gets after replacing the effect translation units by libckd_b200.so.
//
//   g++ -std=c++17 -O2 -Iinclude examples/headless_demo.cpp -Lcookiedough_b200 -lckd_b200 -Wl,-rpath,$PWD/cookiedough_b200 -o headless_demo
//   cd /path/to/cookiedough/target && headless_demo [seconds=222] [fps=60] [out.ckdf] [start=0]
//
// Run from the reference's target/ directory (or pass CKD_ASSET_ROOT / CKD_ROCKET): the art is decoded from the PNG / JPEG
// files by the library itself, the timeline comes from directors-cut.rocket (or the sync/ directory), every frame is
// composed on the GPU by Demo_Draw and lands in the page-locked pDest exactly like the reference's pDest (main.cpp:307),
// from where Display::Update would present it -- here it goes to a raw frame stream instead (tools/ckdf_to_png.py reads it).
//
// Only names of the reference appear in the loop (Image_Create, Shared_Create, Polar_Create, FxBlitter_Create, BoxBlur_Create,
// Demo_Create, Demo_Draw, Demo_Destroy ...); the three CkdHost_ calls replace what SDL/BASS provided: a device, the music
// position and a place to put the pixels.

#include "ckd_host.h"

#include <chrono>
#include <stdio.h>
#include <stdlib.h>

// main.h:37-38 -- compile-time constants in the reference; here CKD_RES_X / CKD_RES_Y may override them (3840 x 2160 renders
// the demo at 4K from the same 1280x720 art: the library resamples it by the rules of host/ckd_image.cpp)
static unsigned kResX = 1280, kResY = 720;

int main(int argc, char **argv)
{
	const double seconds = argc > 1 ? atof(argv[1]) : 222.0;
	const double fps = argc > 2 ? atof(argv[2]) : 60.0;
	const char *outPath = (argc > 3 && argv[3][0]) ? argv[3] : nullptr;
	const double startSeconds = argc > 4 ? atof(argv[4]) : 0.0;
	const char *rocket = getenv("CKD_ROCKET") ? getenv("CKD_ROCKET") : "directors-cut.rocket";
	if (getenv("CKD_RES_X") && getenv("CKD_RES_Y")) { kResX = unsigned(atoi(getenv("CKD_RES_X"))); kResY = unsigned(atoi(getenv("CKD_RES_Y"))); }

	if (!CkdHost_Create(kResX, kResY, 0))
	{
		fprintf(stderr, "%s\n", CkdHost_GetLastError().c_str());
		return 1;
	}
	CkdHost_SetRocketSource(rocket);
	if (getenv("CKD_ASSET_ROOT"))
		CkdHost_SetAssetRoot(getenv("CKD_ASSET_ROOT"));

	// main.cpp:263-279 -- the same calls, the same order
	bool ok = Image_Create();
	ok = ok && Shared_Create();
	ok = ok && Polar_Create();
	ok = ok && FxBlitter_Create();
	ok = ok && BoxBlur_Create();
	ok = ok && Demo_Create();
	if (!ok)
	{
		fprintf(stderr, "%s\n", CkdHost_GetLastError().c_str());
		return 1;
	}

	const unsigned numFrames = unsigned(seconds*fps);
	uint32_t *pDest = nullptr;
	if (outPath)
	{
		if (!CkdSink_Open(outPath, kResX, kResY, numFrames, 4, true, true))
		{
			fprintf(stderr, "%s\n", CkdHost_GetLastError().c_str());
			return 1;
		}
	}
	else
	{
		pDest = static_cast<uint32_t *>(aligned_alloc(64, size_t(kResX)*kResY*sizeof(uint32_t))); // main.cpp:307
		CkdHost_PinFrameBuffer(pDest);
	}

	const auto start = std::chrono::steady_clock::now();
	unsigned frame = 0;
	for (; frame < numFrames; ++frame)
	{
		const double audioTime = startSeconds + frame/fps; // Audio_Get_Pos_In_Sec(), main.cpp:336
		CkdHost_SetTime(audioTime);
		uint32_t *target = outPath ? CkdSink_Acquire() : pDest;
		if (!Demo_Draw(target, float(audioTime), float(100.0/fps)))
			break;                                    // demo:quit, demo.cpp:473-474
		if (outPath && !CkdSink_Commit(target, frame))
			break;
	}
	const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
	printf("%u frames of %ux%u in %.3f s: %.1f fps\n", frame, kResX, kResY, elapsed, frame/elapsed);

	if (outPath)
		CkdSink_Close();
	else
	{
		CkdHost_UnpinFrameBuffer(pDest);
		free(pDest);
	}

	// main.cpp:370-374
	Demo_Destroy();
	BoxBlur_Destroy();
	FxBlitter_Destroy();
	Polar_Destroy();
	Shared_Destroy();
	Image_Destroy();
	CkdHost_Destroy();
	return 0;
}

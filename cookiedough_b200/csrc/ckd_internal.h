// ckd_internal.h -- context layout and helpers shared by the translation units of libckd_b200.so
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>

#include "../../include/ckd.h"

constexpr int kCkdNumFxMaps = 4;         // fx-blitter.h:14
constexpr int kCkdNumRenderTargets = 4;  // shared-resources.h:11
constexpr int kCkdCosTabSize = 2048;     // sincos-lut.h:8

struct ckd_image_slot {
	void *d_pixels = nullptr;
	int width = 0, height = 0, bpp = 0;
	uint32_t firstPixel = 0; // host copy of the first 4 bytes (the voxel effects clear with s_pFogGradient[0])
	// block-linear twin for the voxel casters' bilinear footprints (ckd_footprint_texture), made on first use from d_pixels
	cudaArray_t gatherArray = nullptr;
	cudaTextureObject_t gatherTex = 0;
};

struct ckd_ctx {
	int device = 0;
	int resX = 0, resY = 0;     // kResX, kResY
	int fxX = 0, fxY = 0;       // kFxMapResX, kFxMapResY
	int numSMs = 148;
	cudaStream_t stream = nullptr;
	cudaStream_t ownedStream = nullptr;             // ckd_own_stream: a non-blocking stream that belongs to the context
	cudaEvent_t evJoin = nullptr;                  // ckd_join: marks what has been enqueued on this context's stream
	cudaEvent_t evStart = nullptr, evStop = nullptr;
	cudaStream_t copyStream = nullptr;             // read-back overlapped with rendering (ckd_download_overlapped)
	cudaEvent_t evRendered[2] = {}, evCopied[2] = {};
	bool copyPending[2] = { false, false };

	// banded read-back (ckd_arm_readback): the next effect draw whose last stages are raymarch + Fx_Blit_2x2 renders and
	// copies its frame in row bands
	static constexpr int kMaxBands = 8;
	void *rbHost = nullptr;                         // armed destination (page-locked host memory), cleared by the draw that uses it
	int rbBands = 0;
	bool rbIssued = false;                          // band copies are on the copy stream
	cudaEvent_t evBand[kMaxBands] = {};
	bool newBlurAttrSet = false;                   // opt-in shared-memory size set for new_blur_line_kernel on this device
	bool blurAttrSet[72] = {};                     // opt-in shared-memory size set for the staged blur variants on this device
	int scapeCols = 0;                             // columns per landscape CTA, picked on the first draw (PickScapeColumns)
	unsigned long long launches = 0;

	// device twins of the reference's global buffers (all carved out of one allocation, with guard rows)
	uint32_t *d_frame = nullptr;
	uint32_t *d_frame2 = nullptr;   // second frame buffer of the overlapped read-back (ckd_frame_slot(ctx, 1))
	uint32_t *d_fxMap[kCkdNumFxMaps] = {};
	uint32_t *d_renderTarget[kCkdNumRenderTargets] = {};
	uint32_t *d_scratch[2] = {};    // blur / effect scratch, output sized (+guard)
	uint32_t *d_spikeBlurMap = nullptr; // s_pSpikeBlurMap (shadertoy.cpp:99,185), FX-map sized
	uint8_t *d_ballHeightMix = nullptr; // s_heightMapMix (ball.cpp:21), 1024x1024
	uint32_t *d_ballBeamMix = nullptr;  // s_pBeamMapMix (ball.cpp:22), 1024x1024
	void *d_pool = nullptr;

	// tables
	float h_cosLUT[kCkdCosTabSize+1];
	float2 *d_cosLUT2 = nullptr;      // [i] = (LUT[i], LUT[i+1]-LUT[i]) so one 8-byte load feeds the lerp
	double h_fastCosTab[1025];        // g_fastCosTab (fast-cosine.cpp:9)
	double *d_fastCosTab = nullptr;
	uint32_t *d_rsqrtTab = nullptr;   // host RSQRTPS table
	uint32_t *h_rsqrtTab = nullptr;
	int rsqrtLog2Bin = 13;
	size_t rsqrtEntries = 0;
	int32_t *d_polarMap = nullptr, *d_polarInvMap = nullptr; // 2 ints per pixel
	int32_t *d_polarMap2x2 = nullptr, *d_polarInvMap2x2 = nullptr; // FX-map sized twins (s_pMap2x2/s_pInvMap2x2, polar.cpp:15-16), built on first use
	int *d_voxelTables = nullptr;     // per-frame projection / lighting tables for ball & twister
	float *d_rayParams = nullptr;     // per-ray host-computed parameters (ball fan deltas, twister origins)
	unsigned *d_tileCounters = nullptr; // two alternating work-queue counters of the raymarch kernels
	unsigned long long tileLaunches = 0;
	unsigned long long *d_checksumWork = nullptr; // ckd_frame_checksum scratch, allocated on first use
	bool frameIndependent = false;    // ckd_set_frame_independent: no pixel of a frame may depend on an earlier frame
	unsigned long long inputsGen = 1; // bumped by every setter of an image or table: lets ckd_clone_inputs skip what it has already copied
	const ckd_ctx *clonedFrom = nullptr;
	unsigned long long clonedGen = 0;

	ckd_image_slot images[CKD_IMG_COUNT];

	// optional per-launch CUDA-event profiler (ckd_profile_begin / ckd_profile_end)
	bool profiling = false;
	int profPending = -1;
	struct ProfEntry { const char *name; double algoBytes; cudaEvent_t start, stop; };
	std::vector<ProfEntry> profEntries;
	size_t profUsed = 0;
};

// brackets the next kernel launch with CUDA events when profiling is on; algoBytes = algorithmic bytes of that launch
void ckd_prof_begin(ckd_ctx *ctx, const char *name, double algoBytes);
void ckd_prof_end(ckd_ctx *ctx);

int ckd_footprint_texture(ckd_ctx *ctx, int slot, cudaTextureObject_t *pTex); // single-channel wrap-addressed texture over the image (8- or 32-bit texels) for tex2Dgather
void ckd_release_footprint_texture(ckd_image_slot &slot);
int ckd_ensure_polar_maps_2x2(ckd_ctx *ctx);
int ckd_ensure_copy_stream(ckd_ctx *ctx);         // copy stream + its events, created on first use
int ckd_polar_tail(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, int inverse, bool alpha, const uint32_t *d_softLightSrc); // Polar_Blit[A] (+ halo) as a frame's last stage: banded when a read-back is armed
int ckd_fx_blit_2x2_rows(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, int y0, int y1); // FX rows [y0, y1) -> output rows [2*y0, 2*y1) // builds and uploads the FX-map sized polar maps once

void ckd_set_error(const std::string &message);
int ckd_cuda_fail(cudaError_t err, const char *what, const char *file, int line);

#define CKD_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return ckd_cuda_fail(_e, #expr, __FILE__, __LINE__); } while (0)
#define CKD_CHECK_LAUNCH(ctx) do { (ctx)->launches++; ckd_prof_end(ctx); cudaError_t _e = cudaPeekAtLastError(); if (_e != cudaSuccess) return ckd_cuda_fail(_e, "kernel launch", __FILE__, __LINE__); } while (0)
#define CKD_REQUIRE(cond, msg) do { if (!(cond)) { ckd_set_error(std::string(__func__) + ": " + (msg)); return CKD_ERR_INVALID; } } while (0)
#define CKD_TRY(expr) do { int _r = (expr); if (_r != CKD_OK) return _r; } while (0)

static inline unsigned ckd_div_up(size_t a, size_t b) { return unsigned((a + b - 1)/b); }

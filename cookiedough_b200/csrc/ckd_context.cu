// ckd_context.cu -- context, device buffers, tables, copies.
//
// Replaces the reference's start-up allocations and tables: Shared_Create (shared-resources.cpp:14-37),
// FxBlitter_Create (fx-blitter.cpp:10-20), Polar_Create/CalculateMaps (polar.cpp:18-72), BoxBlur_Create
// (boxblur.cpp:23-28), CalculateCosLUT (sincos-lut.cpp:9-16).  Tables that the reference computes with the host's
// libm or CPU (cos LUT, polar UV maps, the RSQRTPS approximation) are computed the same way here, on the host,
// once, and uploaded.

#include "ckd_internal.h"
#include "ckd_hostmath.h"

#include <xmmintrin.h>
#include <vector>
#include <mutex>

static thread_local std::string t_lastError;
static std::string g_lastError;
static std::mutex g_errMutex;

void ckd_set_error(const std::string &message)
{
	t_lastError = message;
	std::lock_guard<std::mutex> lock(g_errMutex);
	g_lastError = message;
}

int ckd_cuda_fail(cudaError_t err, const char *what, const char *file, int line)
{
	ckd_set_error(std::string("CUDA error '") + cudaGetErrorString(err) + "' in " + what + " (" + file + ":" + std::to_string(line) + ")");
	return CKD_ERR_CUDA;
}

extern "C" const char *ckd_last_error(void)
{
	if (!t_lastError.empty())
		return t_lastError.c_str();
	return g_lastError.c_str();
}

extern "C" const char *ckd_version(void) { return "cookiedough_b200 0.1 (sm_100a)"; }

// ---------------------------------------------------------------------------------------------------------------
// host-side table builders
// ---------------------------------------------------------------------------------------------------------------

// CalculateCosLUT, sincos-lut.cpp:9-16 (host cosf, like the reference)
static void BuildCosLUT(float *lut)
{
	for (unsigned i = 0; i < kCkdCosTabSize; ++i)
		lut[i] = cosf(float(i)*(ckdh::k2PI/kCkdCosTabSize));
	lut[kCkdCosTabSize] = lut[0];
}

static inline uint32_t HostRsqrtBits(uint32_t bits)
{
	float x, r;
	memcpy(&x, &bits, 4);
	_mm_store_ss(&r, _mm_rsqrt_ss(_mm_set_ss(x)));
	uint32_t out;
	memcpy(&out, &r, 4);
	return out;
}

// Enumerates the host CPU's RSQRTSS over both exponent parities and all 2^23 mantissas, finds the widest aligned
// power-of-two mantissa bin on which the instruction is constant and returns the table (hard part 1).
static void ProbeHostRsqrt(std::vector<uint32_t> &table, int &log2Bin)
{
	std::vector<uint32_t> full(size_t(2) << 23);
	for (unsigned parity = 0; parity < 2; ++parity)
		for (uint32_t mant = 0; mant < (1u<<23); ++mant)
			full[(size_t(parity)<<23) + mant] = HostRsqrtBits(((126u+parity)<<23) | mant);

	log2Bin = 0;
	for (int bits = 23; bits >= 1; --bits)
	{
		const uint32_t bin = 1u<<bits;
		bool constant = true;
		for (size_t base = 0; base < full.size() && constant; base += bin)
		{
			const uint32_t v = full[base];
			for (uint32_t i = 1; i < bin; ++i)
				if (full[base+i] != v) { constant = false; break; }
		}
		if (constant) { log2Bin = bits; break; }
	}

	const size_t entries = size_t(2) << (23-log2Bin);
	table.resize(entries);
	for (size_t i = 0; i < entries; ++i)
		table[i] = full[i << log2Bin];
}

// InitializeFastCosine, fast-cosine.cpp:11-17 (host cos, double)
static void BuildFastCosTable(double *table)
{
	for (unsigned iPhase = 0; iPhase < 1024+1; ++iPhase)
		table[iPhase] = cos(double(iPhase)*(ckdh::k2PI/1024));
}

// CalculateMaps, polar.cpp:18-59 (host sqrtf/atan2f, like the reference)
static void BuildPolarMaps(int32_t *pDest, int32_t *pInvDest, unsigned srcResX, unsigned srcResY, unsigned destResX, unsigned destResY)
{
	const float halfResX = destResX/2.f;
	const float halfResY = destResY/2.f;

	size_t iPixel = 0;
	const float maxDist = sqrtf(halfResX*halfResX + halfResY*halfResY);
	for (float Y = -halfResY; Y < halfResY; Y += 1.f)
	{
		for (float X = -halfResX + ckdh::kEpsilon; X < halfResX; X += 1.f)
		{
			const float distance = sqrtf(X*X + Y*Y) / maxDist;
			float theta = atan2f(Y, X);
			theta += ckdh::kPI;
			theta /= ckdh::kPI*2.f;
			const float U    = distance*(srcResX-1.f);
			const float invU = (1.f-distance) * (srcResX-1.f);
			const float V    = theta * (srcResY-1.f);

			if (U >= srcResX-1.f)
				pDest[iPixel] = ((srcResX-2)<<8) | 0xff;
			else
				pDest[iPixel] = ckdh::ftofp24(U);

			if (invU >= srcResX-1.f)
				pInvDest[iPixel] = ((srcResX-2)<<8) | 0xff;
			else
				pInvDest[iPixel] = ckdh::ftofp24(invU);

			if (V >= srcResY-1.f)
				pInvDest[iPixel+1] = pDest[iPixel+1] = ((srcResY-2)<<8) | 0xff;
			else
				pInvDest[iPixel+1] = pDest[iPixel+1] = ckdh::ftofp24(V);

			iPixel += 2;
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------------

static size_t AlignUp(size_t v, size_t a) { return (v + a - 1)/a*a; }

extern "C" int ckd_set_cos_lut(ckd_ctx *ctx, const float *lut2049)
{
	CKD_REQUIRE(ctx && lut2049, "null argument");
	ctx->inputsGen++;
	memcpy(ctx->h_cosLUT, lut2049, sizeof(ctx->h_cosLUT));
	std::vector<float2> pairs(kCkdCosTabSize);
	for (int i = 0; i < kCkdCosTabSize; ++i)
		pairs[i] = make_float2(lut2049[i], lut2049[i+1] - lut2049[i]); // b-a of lerpf (Math.h:52-56), same float subtraction
	CKD_CUDA(cudaMemcpy(ctx->d_cosLUT2, pairs.data(), pairs.size()*sizeof(float2), cudaMemcpyHostToDevice));
	return CKD_OK;
}

extern "C" int ckd_set_fast_cos_table(ckd_ctx *ctx, const double *table1025)
{
	CKD_REQUIRE(ctx && table1025, "null argument");
	ctx->inputsGen++;
	memcpy(ctx->h_fastCosTab, table1025, sizeof(ctx->h_fastCosTab));
	CKD_CUDA(cudaMemcpy(ctx->d_fastCosTab, table1025, sizeof(ctx->h_fastCosTab), cudaMemcpyHostToDevice));
	return CKD_OK;
}

extern "C" int ckd_get_fast_cos_table(ckd_ctx *ctx, double *out_table1025)
{
	CKD_REQUIRE(ctx && out_table1025, "null argument");
	memcpy(out_table1025, ctx->h_fastCosTab, sizeof(ctx->h_fastCosTab));
	return CKD_OK;
}

extern "C" int ckd_set_frame_independent(ckd_ctx *ctx, int enabled)
{
	CKD_REQUIRE(ctx, "null context");
	ctx->frameIndependent = 0 != enabled;
	return CKD_OK;
}

extern "C" int ckd_set_rsqrt_table(ckd_ctx *ctx, const uint32_t *table, int log2_bin)
{
	CKD_REQUIRE(ctx && table, "null argument");
	CKD_REQUIRE(log2_bin >= 0 && log2_bin <= 23, "log2_bin out of range");
	ctx->inputsGen++;
	const size_t entries = size_t(2) << (23-log2_bin);
	if (ctx->d_rsqrtTab) { cudaFree(ctx->d_rsqrtTab); ctx->d_rsqrtTab = nullptr; }
	free(ctx->h_rsqrtTab);
	ctx->h_rsqrtTab = static_cast<uint32_t *>(malloc(entries*sizeof(uint32_t)));
	memcpy(ctx->h_rsqrtTab, table, entries*sizeof(uint32_t));
	CKD_CUDA(cudaMalloc(&ctx->d_rsqrtTab, entries*sizeof(uint32_t)));
	CKD_CUDA(cudaMemcpy(ctx->d_rsqrtTab, table, entries*sizeof(uint32_t), cudaMemcpyHostToDevice));
	ctx->rsqrtLog2Bin = log2_bin;
	ctx->rsqrtEntries = entries;
	return CKD_OK;
}

extern "C" int ckd_get_rsqrt_table(ckd_ctx *ctx, uint32_t *out_table, size_t max_entries, int *out_log2_bin, size_t *out_entries)
{
	CKD_REQUIRE(ctx, "null context");
	if (out_log2_bin) *out_log2_bin = ctx->rsqrtLog2Bin;
	if (out_entries) *out_entries = ctx->rsqrtEntries;
	if (out_table)
	{
		CKD_REQUIRE(max_entries >= ctx->rsqrtEntries, "table buffer too small");
		memcpy(out_table, ctx->h_rsqrtTab, ctx->rsqrtEntries*sizeof(uint32_t));
	}
	return CKD_OK;
}

extern "C" int ckd_set_polar_maps(ckd_ctx *ctx, const int32_t *map, const int32_t *inv_map)
{
	CKD_REQUIRE(ctx && map && inv_map, "null argument");
	ctx->inputsGen++;
	const size_t bytes = size_t(ctx->resX)*ctx->resY*2*sizeof(int32_t);
	CKD_CUDA(cudaMemcpy(ctx->d_polarMap, map, bytes, cudaMemcpyHostToDevice));
	CKD_CUDA(cudaMemcpy(ctx->d_polarInvMap, inv_map, bytes, cudaMemcpyHostToDevice));
	return CKD_OK;
}

extern "C" int ckd_get_polar_maps(ckd_ctx *ctx, int32_t *out_map, int32_t *out_inv_map)
{
	CKD_REQUIRE(ctx, "null context");
	const size_t bytes = size_t(ctx->resX)*ctx->resY*2*sizeof(int32_t);
	if (out_map) CKD_CUDA(cudaMemcpy(out_map, ctx->d_polarMap, bytes, cudaMemcpyDeviceToHost));
	if (out_inv_map) CKD_CUDA(cudaMemcpy(out_inv_map, ctx->d_polarInvMap, bytes, cudaMemcpyDeviceToHost));
	return CKD_OK;
}

// s_pMap2x2 / s_pInvMap2x2 (polar.cpp:15-16,69): only Polar_Blit_2x2 reads them and the demo never calls it, so they are
// built on first use instead of in ckd_create (2 x 16.7 MB at 3840x2160)
int ckd_ensure_polar_maps_2x2(ckd_ctx *ctx)
{
	if (ctx->d_polarMap2x2)
		return CKD_OK;
	const size_t fxPixels = size_t(ctx->fxX)*ctx->fxY;
	const size_t bytes = fxPixels*2*sizeof(int32_t);
	std::vector<int32_t> map(fxPixels*2), invMap(fxPixels*2);
	BuildPolarMaps(map.data(), invMap.data(), ctx->fxX, ctx->fxY, ctx->fxX, ctx->fxY); // polar.cpp:69
	int32_t *d_maps = nullptr;
	CKD_CUDA(cudaMalloc(&d_maps, 2*bytes));
	cudaError_t err = cudaMemcpy(d_maps, map.data(), bytes, cudaMemcpyHostToDevice);
	if (cudaSuccess == err) err = cudaMemcpy(d_maps + fxPixels*2, invMap.data(), bytes, cudaMemcpyHostToDevice);
	if (cudaSuccess != err) { cudaFree(d_maps); return ckd_cuda_fail(err, "cudaMemcpy(polar maps 2x2)", __FILE__, __LINE__); }
	ctx->d_polarMap2x2 = d_maps;
	ctx->d_polarInvMap2x2 = d_maps + fxPixels*2;
	return CKD_OK;
}

extern "C" int ckd_create(ckd_ctx **out_ctx, int res_x, int res_y, int device)
{
	CKD_REQUIRE(out_ctx, "null out_ctx");
	*out_ctx = nullptr;
	CKD_REQUIRE(res_x >= 64 && res_y >= 64 && res_x <= 16384 && res_y <= 16384, "resolution out of range");
	CKD_REQUIRE(0 == (res_x & 7) && 0 == (res_y & 7), "resolution must be a multiple of 8");

	int numDevices = 0;
	cudaError_t err = cudaGetDeviceCount(&numDevices);
	if (err != cudaSuccess || numDevices <= 0)
	{
		ckd_set_error(std::string("ckd_create: no usable CUDA device (") + cudaGetErrorString(err) + "); this library has no CPU fallback");
		return CKD_ERR_CUDA;
	}
	CKD_REQUIRE(device >= 0 && device < numDevices, "device index out of range");
	CKD_CUDA(cudaSetDevice(device));

	cudaDeviceProp prop;
	CKD_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10)
	{
		ckd_set_error(std::string("ckd_create: device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + "; this library ships sm_100a code only");
		return CKD_ERR_CUDA;
	}

	ckd_ctx *ctx = new ckd_ctx;
	ctx->device = device;
	ctx->resX = res_x;
	ctx->resY = res_y;
	ctx->fxX = res_x/2 + 4; // fx-blitter.h:16-17
	ctx->fxY = res_y/2 + 4;
	ctx->numSMs = prop.multiProcessorCount;

	const size_t outPixels = size_t(res_x)*res_y;
	const size_t fxPixels = size_t(ctx->fxX)*ctx->fxY;
	const size_t guard = size_t(res_x)*4; // guard rows: the reference reads/writes a little past several buffers
	const size_t outBytes = AlignUp((outPixels + guard)*4, 256);
	const size_t fxBytes = AlignUp((fxPixels + guard)*4, 256);
	const size_t mapBytes = AlignUp(outPixels*2*sizeof(int32_t), 256);
	const size_t ballMapPixels = 1024*1024;

	size_t total = 0;
	auto carve = [&](size_t bytes) { size_t off = total; total += AlignUp(bytes, 256); return off; };
	const size_t offFrame = carve(outBytes);
	const size_t offFrame2 = carve(outBytes); // ckd_frame_slot(ctx, 1): its own image, no effect or blur scratch aliases it
	size_t offRT[kCkdNumRenderTargets], offFx[kCkdNumFxMaps], offScratch[2];
	for (auto &o : offRT) o = carve(outBytes);
	for (auto &o : offFx) o = carve(fxBytes);
	for (auto &o : offScratch) o = carve(outBytes);
	const size_t offSpikeBlur = carve(fxBytes);
	const size_t offBallH = carve(ballMapPixels + 4096);
	const size_t offBallB = carve(ballMapPixels*4 + 4096);
	const size_t offMap = carve(mapBytes), offInvMap = carve(mapBytes);
	const size_t offCos = carve(kCkdCosTabSize*sizeof(float2));
	const size_t offFastCos = carve(1025*sizeof(double));
	const size_t offVox = carve(8192*sizeof(int));
	const size_t offRay = carve(size_t(res_y)*8*sizeof(float) + 4096);
	const size_t offCounters = carve(256);

	err = cudaMalloc(&ctx->d_pool, total);
	if (err != cudaSuccess) { delete ctx; return ckd_cuda_fail(err, "cudaMalloc(pool)", __FILE__, __LINE__); }
	err = cudaMemset(ctx->d_pool, 0, total);
	if (err != cudaSuccess) { cudaFree(ctx->d_pool); delete ctx; return ckd_cuda_fail(err, "cudaMemset(pool)", __FILE__, __LINE__); }

	uint8_t *base = static_cast<uint8_t *>(ctx->d_pool);
	ctx->d_frame = reinterpret_cast<uint32_t *>(base + offFrame);
	ctx->d_frame2 = reinterpret_cast<uint32_t *>(base + offFrame2);
	for (int i = 0; i < kCkdNumRenderTargets; ++i) ctx->d_renderTarget[i] = reinterpret_cast<uint32_t *>(base + offRT[i]);
	for (int i = 0; i < kCkdNumFxMaps; ++i) ctx->d_fxMap[i] = reinterpret_cast<uint32_t *>(base + offFx[i]);
	for (int i = 0; i < 2; ++i) ctx->d_scratch[i] = reinterpret_cast<uint32_t *>(base + offScratch[i]);
	ctx->d_spikeBlurMap = reinterpret_cast<uint32_t *>(base + offSpikeBlur);
	ctx->d_ballHeightMix = base + offBallH;
	ctx->d_ballBeamMix = reinterpret_cast<uint32_t *>(base + offBallB);
	ctx->d_polarMap = reinterpret_cast<int32_t *>(base + offMap);
	ctx->d_polarInvMap = reinterpret_cast<int32_t *>(base + offInvMap);
	ctx->d_cosLUT2 = reinterpret_cast<float2 *>(base + offCos);
	ctx->d_fastCosTab = reinterpret_cast<double *>(base + offFastCos);
	ctx->d_voxelTables = reinterpret_cast<int *>(base + offVox);
	ctx->d_rayParams = reinterpret_cast<float *>(base + offRay);
	ctx->d_tileCounters = reinterpret_cast<unsigned *>(base + offCounters);

	int rc = CKD_OK;
	do
	{
		if (cudaSuccess != (err = cudaEventCreate(&ctx->evStart)) || cudaSuccess != (err = cudaEventCreate(&ctx->evStop)))
		{ rc = ckd_cuda_fail(err, "cudaEventCreate", __FILE__, __LINE__); break; }

		float lut[kCkdCosTabSize+1];
		BuildCosLUT(lut);
		if (CKD_OK != (rc = ckd_set_cos_lut(ctx, lut))) break;

		double fastCos[1025];
		BuildFastCosTable(fastCos);
		if (CKD_OK != (rc = ckd_set_fast_cos_table(ctx, fastCos))) break;

		std::vector<uint32_t> rsqrtTab;
		int log2Bin = 0;
		ProbeHostRsqrt(rsqrtTab, log2Bin);
		if (CKD_OK != (rc = ckd_set_rsqrt_table(ctx, rsqrtTab.data(), log2Bin))) break;

		std::vector<int32_t> map(outPixels*2), invMap(outPixels*2);
		BuildPolarMaps(map.data(), invMap.data(), res_x, res_y, res_x, res_y); // kTargetRes == kRes (shared-resources.h:22-23)
		if (CKD_OK != (rc = ckd_set_polar_maps(ctx, map.data(), invMap.data()))) break;
	}
	while (false);

	if (rc != CKD_OK)
	{
		ckd_destroy(ctx);
		return rc;
	}

	*out_ctx = ctx;
	return CKD_OK;
}

extern "C" void ckd_destroy(ckd_ctx *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	for (auto &slot : ctx->images)
	{
		ckd_release_footprint_texture(slot);
		if (slot.d_pixels) cudaFree(slot.d_pixels);
	}
	if (ctx->d_rsqrtTab) cudaFree(ctx->d_rsqrtTab);
	if (ctx->d_polarMap2x2) cudaFree(ctx->d_polarMap2x2);
	if (ctx->d_checksumWork) cudaFree(ctx->d_checksumWork);
	free(ctx->h_rsqrtTab);
	for (auto &e : ctx->profEntries) { cudaEventDestroy(e.start); cudaEventDestroy(e.stop); }
	for (int i = 0; i < 2; ++i)
	{
		if (ctx->evRendered[i]) cudaEventDestroy(ctx->evRendered[i]);
		if (ctx->evCopied[i]) cudaEventDestroy(ctx->evCopied[i]);
	}
	for (auto &ev : ctx->evBand)
		if (ev) cudaEventDestroy(ev);
	if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
	if (ctx->ownedStream) cudaStreamDestroy(ctx->ownedStream);
	if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
	if (ctx->evStart) cudaEventDestroy(ctx->evStart);
	if (ctx->evStop) cudaEventDestroy(ctx->evStop);
	if (ctx->d_pool) cudaFree(ctx->d_pool);
	delete ctx;
}

extern "C" int ckd_set_stream(ckd_ctx *ctx, void *cuda_stream)
{
	CKD_REQUIRE(ctx, "null context");
	ctx->stream = static_cast<cudaStream_t>(cuda_stream);
	return CKD_OK;
}

extern "C" int ckd_own_stream(ckd_ctx *ctx)
{
	CKD_REQUIRE(ctx, "null context");
	if (nullptr == ctx->ownedStream)
		CKD_CUDA(cudaStreamCreateWithFlags(&ctx->ownedStream, cudaStreamNonBlocking));
	ctx->stream = ctx->ownedStream;
	return CKD_OK;
}

extern "C" int ckd_join(ckd_ctx *ctx, ckd_ctx *other)
{
	CKD_REQUIRE(ctx && other, "null context");
	if (nullptr == other->evJoin)
		CKD_CUDA(cudaEventCreateWithFlags(&other->evJoin, cudaEventDisableTiming));
	CKD_CUDA(cudaEventRecord(other->evJoin, other->stream));
	CKD_CUDA(cudaStreamWaitEvent(ctx->stream, other->evJoin, 0));
	return CKD_OK;
}

extern "C" int ckd_clone_inputs(ckd_ctx *dst, const ckd_ctx *src)
{
	CKD_REQUIRE(dst && src && dst != src, "two different contexts are needed");
	CKD_REQUIRE(dst->resX == src->resX && dst->resY == src->resY && dst->device == src->device, "the contexts must have the same resolution and device");
	dst->frameIndependent = src->frameIndependent;
	if (dst->clonedFrom == src && dst->clonedGen == src->inputsGen)
		return CKD_OK;                     // nothing was uploaded or set on the source since the last copy
	CKD_CUDA(cudaDeviceSynchronize()); // whatever still writes the source's inputs has to land first; this is a set-up call
	for (int i = 0; i < CKD_IMG_COUNT; ++i)
	{
		const ckd_image_slot &from = src->images[i];
		ckd_image_slot &to = dst->images[i];
		ckd_release_footprint_texture(to);
		if (to.d_pixels) { cudaFree(to.d_pixels); to.d_pixels = nullptr; }
		to = from;
		to.d_pixels = nullptr;
		to.gatherArray = nullptr; // the twin is made from this context's own copy on first use
		to.gatherTex = 0;
		if (nullptr == from.d_pixels)
			continue;
		const size_t bytes = size_t(from.width)*from.height*from.bpp + 256;
		CKD_CUDA(cudaMalloc(&to.d_pixels, bytes));
		CKD_CUDA(cudaMemcpy(to.d_pixels, from.d_pixels, bytes, cudaMemcpyDeviceToDevice));
	}
	CKD_TRY(ckd_set_cos_lut(dst, src->h_cosLUT));
	CKD_TRY(ckd_set_fast_cos_table(dst, src->h_fastCosTab));
	if (nullptr != src->h_rsqrtTab)
		CKD_TRY(ckd_set_rsqrt_table(dst, src->h_rsqrtTab, src->rsqrtLog2Bin));
	const size_t mapBytes = size_t(src->resX)*src->resY*2*sizeof(int32_t);
	CKD_CUDA(cudaMemcpy(dst->d_polarMap, src->d_polarMap, mapBytes, cudaMemcpyDeviceToDevice));
	CKD_CUDA(cudaMemcpy(dst->d_polarInvMap, src->d_polarInvMap, mapBytes, cudaMemcpyDeviceToDevice));
	dst->clonedFrom = src;
	dst->clonedGen = src->inputsGen;
	return CKD_OK;
}

extern "C" int ckd_sync(ckd_ctx *ctx)
{
	CKD_REQUIRE(ctx, "null context");
	CKD_CUDA(cudaStreamSynchronize(ctx->stream));
	return CKD_OK;
}

extern "C" int ckd_res_x(const ckd_ctx *ctx) { return ctx ? ctx->resX : 0; }
extern "C" int ckd_res_y(const ckd_ctx *ctx) { return ctx ? ctx->resY : 0; }
extern "C" int ckd_fxmap_res_x(const ckd_ctx *ctx) { return ctx ? ctx->fxX : 0; }
extern "C" int ckd_fxmap_res_y(const ckd_ctx *ctx) { return ctx ? ctx->fxY : 0; }
extern "C" uint32_t *ckd_frame(ckd_ctx *ctx) { return ctx ? ctx->d_frame : nullptr; }
extern "C" uint32_t *ckd_fxmap(ckd_ctx *ctx, int index) { return (ctx && index >= 0 && index < kCkdNumFxMaps) ? ctx->d_fxMap[index] : nullptr; }
extern "C" uint32_t *ckd_render_target(ckd_ctx *ctx, int index) { return (ctx && index >= 0 && index < kCkdNumRenderTargets) ? ctx->d_renderTarget[index] : nullptr; }
extern "C" unsigned long long ckd_launch_count(const ckd_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int ckd_malloc(ckd_ctx *ctx, void **out_d_ptr, size_t bytes)
{
	CKD_REQUIRE(ctx && out_d_ptr, "null argument");
	CKD_CUDA(cudaSetDevice(ctx->device));
	CKD_CUDA(cudaMalloc(out_d_ptr, bytes));
	return CKD_OK;
}

extern "C" int ckd_free(ckd_ctx *ctx, void *d_ptr)
{
	CKD_REQUIRE(ctx, "null context");
	CKD_CUDA(cudaFree(d_ptr));
	return CKD_OK;
}

extern "C" int ckd_malloc_host(void **out_h_ptr, size_t bytes)
{
	CKD_REQUIRE(out_h_ptr, "null argument");
	CKD_CUDA(cudaMallocHost(out_h_ptr, bytes));
	return CKD_OK;
}

extern "C" int ckd_free_host(void *h_ptr)
{
	CKD_CUDA(cudaFreeHost(h_ptr));
	return CKD_OK;
}

// page-lock a caller-owned host buffer in place (the reference allocates pDest with an aligned malloc, main.cpp:307): frame
// copies into pageable memory go through the driver's staging buffers at a fraction of the PCIe rate
extern "C" int ckd_pin_host(void *h_ptr, size_t bytes)
{
	CKD_REQUIRE(h_ptr && bytes, "null argument");
	CKD_CUDA(cudaHostRegister(h_ptr, bytes, cudaHostRegisterDefault));
	return CKD_OK;
}

extern "C" int ckd_unpin_host(void *h_ptr)
{
	CKD_REQUIRE(h_ptr, "null argument");
	CKD_CUDA(cudaHostUnregister(h_ptr));
	return CKD_OK;
}

extern "C" int ckd_upload(ckd_ctx *ctx, void *d_dst, const void *h_src, size_t bytes)
{
	CKD_REQUIRE(ctx && d_dst && h_src, "null argument");
	CKD_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
	return CKD_OK;
}

extern "C" int ckd_copy(ckd_ctx *ctx, void *d_dst, const void *d_src, size_t bytes)
{
	CKD_REQUIRE(ctx && d_dst && d_src, "null argument");
	CKD_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	ctx->launches++;
	return CKD_OK;
}

extern "C" int ckd_download(ckd_ctx *ctx, void *h_dst, const void *d_src, size_t bytes)
{
	CKD_REQUIRE(ctx && h_dst && d_src, "null argument");
	CKD_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	return CKD_OK;
}

extern "C" uint32_t *ckd_frame_slot(ckd_ctx *ctx, int slot)
{
	if (!ctx || slot < 0 || slot > 1) return nullptr;
	return slot ? ctx->d_frame2 : ctx->d_frame;
}

int ckd_ensure_copy_stream(ckd_ctx *ctx)
{
	if (ctx->copyStream)
		return CKD_OK;
	CKD_CUDA(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
	for (int i = 0; i < 2; ++i)
	{
		CKD_CUDA(cudaEventCreateWithFlags(&ctx->evRendered[i], cudaEventDisableTiming));
		CKD_CUDA(cudaEventCreateWithFlags(&ctx->evCopied[i], cudaEventDisableTiming));
	}
	for (auto &ev : ctx->evBand)
		CKD_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
	return CKD_OK;
}

// ---- banded read-back ------------------------------------------------------------------------------------------------
// A synchronous X_Draw at 4K spends three quarters of its time copying the finished frame over PCIe (33 MB, 0.6 ms) after
// 0.15-0.25 ms of rendering.  Armed with the caller's page-locked frame buffer, the next draw whose last stages are a
// raymarch kernel + Fx_Blit_2x2 issues them per band of FX-map rows and sends every finished band of output rows to the host
// on the copy stream while the next band renders (csrc/ckd_raymarch.cu, RaymarchAndBlit).  Draws that end differently
// ignore the arm; ckd_finish_readback tells the caller which of the two happened.
extern "C" int ckd_arm_readback(ckd_ctx *ctx, void *h_dest, int bands)
{
	CKD_REQUIRE(ctx && h_dest, "null argument");
	ctx->rbHost = nullptr;
	ctx->rbIssued = false;
	if (bands < 2)
		return CKD_OK;
	cudaPointerAttributes attr;
	if (cudaSuccess != cudaPointerGetAttributes(&attr, h_dest) || attr.type != cudaMemoryTypeHost)
	{
		cudaGetLastError(); // pageable memory: an asynchronous copy would be staged and serialise the bands -- leave it to the caller
		return CKD_OK;
	}
	CKD_TRY(ckd_ensure_copy_stream(ctx));
	ctx->rbHost = h_dest;
	ctx->rbBands = bands < ckd_ctx::kMaxBands ? bands : ckd_ctx::kMaxBands;
	return CKD_OK;
}

extern "C" int ckd_finish_readback(ckd_ctx *ctx, int *out_done)
{
	CKD_REQUIRE(ctx && out_done, "null argument");
	*out_done = 0;
	ctx->rbHost = nullptr;
	if (!ctx->rbIssued)
		return CKD_OK;
	ctx->rbIssued = false;
	CKD_CUDA(cudaStreamSynchronize(ctx->copyStream));
	*out_done = 1;
	return CKD_OK;
}

extern "C" int ckd_download_overlapped(ckd_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, int slot)
{
	CKD_REQUIRE(ctx && h_dst && d_src, "null argument");
	CKD_REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
	CKD_TRY(ckd_ensure_copy_stream(ctx));
	CKD_CUDA(cudaEventRecord(ctx->evRendered[slot], ctx->stream));
	CKD_CUDA(cudaStreamWaitEvent(ctx->copyStream, ctx->evRendered[slot], 0));
	CKD_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->copyStream));
	CKD_CUDA(cudaEventRecord(ctx->evCopied[slot], ctx->copyStream));
	// the next kernels that overwrite d_src must not start before the copy has read it
	ctx->copyPending[slot] = true;
	return CKD_OK;
}

extern "C" int ckd_wait_download(ckd_ctx *ctx, int slot)
{
	CKD_REQUIRE(ctx, "null context");
	CKD_REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
	if (ctx->copyPending[slot])
	{
		CKD_CUDA(cudaEventSynchronize(ctx->evCopied[slot]));
		ctx->copyPending[slot] = false;
	}
	return CKD_OK;
}

extern "C" int ckd_timer_start(ckd_ctx *ctx)
{
	CKD_REQUIRE(ctx, "null context");
	CKD_CUDA(cudaEventRecord(ctx->evStart, ctx->stream));
	return CKD_OK;
}

extern "C" int ckd_timer_stop_ms(ckd_ctx *ctx, float *out_ms)
{
	CKD_REQUIRE(ctx && out_ms, "null argument");
	CKD_CUDA(cudaEventRecord(ctx->evStop, ctx->stream));
	CKD_CUDA(cudaEventSynchronize(ctx->evStop));
	CKD_CUDA(cudaEventElapsedTime(out_ms, ctx->evStart, ctx->evStop));
	return CKD_OK;
}

void ckd_prof_begin(ckd_ctx *ctx, const char *name, double algoBytes)
{
	if (!ctx->profiling)
		return;
	if (ctx->profUsed == ctx->profEntries.size())
	{
		ckd_ctx::ProfEntry e = { name, algoBytes, nullptr, nullptr };
		if (cudaSuccess != cudaEventCreate(&e.start) || cudaSuccess != cudaEventCreate(&e.stop))
			return;
		ctx->profEntries.push_back(e);
	}
	ckd_ctx::ProfEntry &e = ctx->profEntries[ctx->profUsed];
	e.name = name;
	e.algoBytes = algoBytes;
	cudaEventRecord(e.start, ctx->stream);
	ctx->profPending = int(ctx->profUsed++);
}

void ckd_prof_end(ckd_ctx *ctx)
{
	if (ctx->profPending < 0)
		return;
	cudaEventRecord(ctx->profEntries[ctx->profPending].stop, ctx->stream);
	ctx->profPending = -1;
}

extern "C" int ckd_profile_begin(ckd_ctx *ctx)
{
	CKD_REQUIRE(ctx, "null context");
	ctx->profiling = true;
	ctx->profUsed = 0;
	ctx->profPending = -1;
	return CKD_OK;
}

extern "C" int ckd_profile_end(ckd_ctx *ctx, ckd_kernel_stat *out_stats, int max_stats, int *out_count)
{
	CKD_REQUIRE(ctx && out_count, "null argument");
	ctx->profiling = false;
	CKD_CUDA(cudaStreamSynchronize(ctx->stream));
	int count = 0;
	for (size_t i = 0; i < ctx->profUsed; ++i)
	{
		const ckd_ctx::ProfEntry &e = ctx->profEntries[i];
		float ms = 0.f;
		CKD_CUDA(cudaEventElapsedTime(&ms, e.start, e.stop));
		int slot = -1;
		for (int j = 0; j < count; ++j)
			if (0 == strncmp(out_stats[j].name, e.name, sizeof(out_stats[j].name)-1)) { slot = j; break; }
		if (slot < 0)
		{
			if (!out_stats || count >= max_stats)
				continue;
			slot = count++;
			memset(&out_stats[slot], 0, sizeof(ckd_kernel_stat));
			strncpy(out_stats[slot].name, e.name, sizeof(out_stats[slot].name)-1);
		}
		out_stats[slot].launches++;
		out_stats[slot].total_ms += ms;
		out_stats[slot].algo_bytes += e.algoBytes;
	}
	*out_count = count;
	ctx->profUsed = 0;
	return CKD_OK;
}

extern "C" int ckd_set_image(ckd_ctx *ctx, ckd_image slot, const void *h_pixels, int width, int height, int bytes_per_pixel)
{
	CKD_REQUIRE(ctx && h_pixels, "null argument");
	CKD_REQUIRE(slot >= 0 && slot < CKD_IMG_COUNT, "bad image slot");
	CKD_REQUIRE(width > 0 && height > 0 && (bytes_per_pixel == 1 || bytes_per_pixel == 4), "bad image geometry");
	ctx->inputsGen++;
	ckd_image_slot &s = ctx->images[slot];
	ckd_release_footprint_texture(s);
	if (s.d_pixels) { cudaFree(s.d_pixels); s.d_pixels = nullptr; }
	const size_t bytes = size_t(width)*height*bytes_per_pixel;
	CKD_CUDA(cudaMalloc(&s.d_pixels, bytes + 256)); // slack like the harness' loader
	CKD_CUDA(cudaMemset(s.d_pixels, 0, bytes + 256));
	CKD_CUDA(cudaMemcpy(s.d_pixels, h_pixels, bytes, cudaMemcpyHostToDevice));
	s.width = width; s.height = height; s.bpp = bytes_per_pixel;
	s.firstPixel = 0;
	memcpy(&s.firstPixel, h_pixels, std::min<size_t>(4, bytes));
	return CKD_OK;
}

void ckd_release_footprint_texture(ckd_image_slot &slot)
{
	if (slot.gatherTex) { cudaDestroyTextureObject(slot.gatherTex); slot.gatherTex = 0; }
	if (slot.gatherArray) { cudaFreeArray(slot.gatherArray); slot.gatherArray = nullptr; }
}

// The voxel casters read a 2x2 footprint per map and step, along rays of any direction.  From linear memory a ray that runs
// across the rows touches a different 128-byte line with every lane and tap (up to 32 tag look-ups per load, 8 loads per step);
// a block-linear array read with tex2Dgather delivers the footprint in one request and keeps neighbouring rows in one tile.
// Texels are single-channel (L8, or the BGRA word as one 32-bit channel), addressing wraps like the reference's '& mapAnd'.
int ckd_footprint_texture(ckd_ctx *ctx, int slot, cudaTextureObject_t *pTex)
{
	CKD_REQUIRE(ctx && pTex && slot >= 0 && slot < CKD_IMG_COUNT, "bad image slot");
	ckd_image_slot &s = ctx->images[slot];
	CKD_REQUIRE(s.d_pixels && (s.bpp == 1 || s.bpp == 4), "image not set");
	if (!s.gatherTex)
	{
		const cudaChannelFormatDesc desc = cudaCreateChannelDesc(8*s.bpp, 0, 0, 0, cudaChannelFormatKindUnsigned);
		CKD_CUDA(cudaMallocArray(&s.gatherArray, &desc, size_t(s.width), size_t(s.height), cudaArrayTextureGather));
		const size_t pitch = size_t(s.width)*s.bpp;
		CKD_CUDA(cudaMemcpy2DToArray(s.gatherArray, 0, 0, s.d_pixels, pitch, pitch, size_t(s.height), cudaMemcpyDeviceToDevice));
		cudaResourceDesc res = {};
		res.resType = cudaResourceTypeArray;
		res.res.array.array = s.gatherArray;
		cudaTextureDesc tex = {};
		tex.addressMode[0] = tex.addressMode[1] = cudaAddressModeWrap;
		tex.filterMode = cudaFilterModePoint;
		tex.readMode = cudaReadModeElementType;
		tex.normalizedCoords = 1;
		CKD_CUDA(cudaCreateTextureObject(&s.gatherTex, &res, &tex, nullptr));
	}
	*pTex = s.gatherTex;
	return CKD_OK;
}

extern "C" const void *ckd_get_image(ckd_ctx *ctx, ckd_image slot)
{
	if (!ctx || slot < 0 || slot >= CKD_IMG_COUNT) return nullptr;
	return ctx->images[slot].d_pixels;
}

// ---------------------------------------------------------------------------------------------------------------
// fastcosf / fastsinf -- fast-cosine.h:17-53
// ---------------------------------------------------------------------------------------------------------------

// The phase 1 + |x|/2pi is kept in double: its exponent says how far the mantissa has to be shifted so that the bits below the
// binary point line up as a 32-bit fraction of a turn; the top 10 of those index the table, the other 22 interpolate.
__device__ __forceinline__ float fastcos_eval(const double *table, double x)
{
	x = fabs(x);
	const double phaseScale = double(1.f/ckdh::k2PI);        // a float constant in the reference (fast-cosine.h:23)
	const double phase = 1.0 + x*phaseScale;
	const unsigned long long phaseBits = (unsigned long long) __double_as_longlong(phase);
	const int exponent = int(phaseBits >> 52) - 1023;
	// x86 shifts by the low 6 bits of the count (huge and non-finite arguments: "quality degrades", but the bits are defined)
	const unsigned significand = unsigned((phaseBits << (unsigned(exponent) & 63u)) >> (52-32));
	const unsigned index = significand >> 22;
	const double left = table[index], right = table[index+1];
	const double t = double(significand & 0x3fffffu)*(1.0/4194304.0);
	return float(left + (right-left)*t);
}

__global__ void __launch_bounds__(256) fastcos_kernel(float *__restrict__ pOut, const double *__restrict__ pIn, size_t n, const double *__restrict__ table, int sine)
{
	__shared__ double s_table[1025];
	for (int i = threadIdx.x; i < 1025; i += blockDim.x)
		s_table[i] = table[i];
	__syncthreads();
	const size_t stride = size_t(gridDim.x)*blockDim.x;
	for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n; i += stride)
		pOut[i] = fastcos_eval(s_table, sine ? pIn[i] - 0.25 : pIn[i]); // fastsinf, fast-cosine.h:51-53
}

extern "C" int ckd_fastcos(ckd_ctx *ctx, float *d_out, const double *d_x, size_t n, int sine)
{
	CKD_REQUIRE(ctx && d_out && d_x, "null argument");
	if (0 == n)
		return CKD_OK;
	CKD_CUDA(cudaSetDevice(ctx->device));
	const unsigned blocks = unsigned(std::max<size_t>(1, std::min<size_t>(size_t(ctx->numSMs)*8, ckd_div_up(n, 256))));
	ckd_prof_begin(ctx, "fastcos", 12.0*double(n));
	fastcos_kernel<<<blocks, 256, 0, ctx->stream>>>(d_out, d_x, n, ctx->d_fastCosTab, sine);
	CKD_CHECK_LAUNCH(ctx);
	return CKD_OK;
}

// ckd_post.cu -- the integer 2D post chain: Fx_Blit_2x2, blend ops, rectangular blits, memset32, polar remap, TapeWarp32.
//
// All of these are HBM-bound byte shuffles: one thread produces 4 (or 2x4) output pixels, loads/stores are 128-bit,
// grids are sized from the pixel count, and no shared memory is needed (no reuse beyond what L1/L2 give for free).
// Results are bit-exact with the reference's SSE code (16-bit lane arithmetic restated on packed 32-bit words).

#include "ckd_internal.h"
#include "ckd_math.cuh"
#include "ckd_hostmath.h"

#include <stdlib.h>

using namespace ckd;

// ---------------------------------------------------------------------------------------------------------------
// Fx_Blit_2x2 -- fx-blitter.cpp:27-75
// ---------------------------------------------------------------------------------------------------------------
// Each thread takes 2 horizontally adjacent FX-map pixels (plus their right/down neighbours from the guard band)
// and writes 4 output pixels on each of 2 output rows with one 128-bit store per row.

__global__ void __launch_bounds__(256) fx_blit_2x2_kernel(uint32_t *__restrict__ pDest, const uint32_t *__restrict__ pSrc, int fxX, int resX, int halfX, int y0, int y1)
{
	const int pairX = blockIdx.x*blockDim.x + threadIdx.x; // index of the source pixel pair
	const int iY = y0 + blockIdx.y*blockDim.y + threadIdx.y; // FX-map rows [y0, y1)
	if (pairX*2 >= halfX || iY >= y1)
		return;

	const int sx = pairX*2;
	const uint32_t *row0 = pSrc + size_t(iY)*fxX + sx;
	const uint32_t *row1 = row0 + fxX;

	const uint2 a01 = *reinterpret_cast<const uint2 *>(row0); // sx is even and fxX is a multiple of 4: 8-byte aligned
	const uint32_t a2 = row0[2];
	const uint2 b01 = *reinterpret_cast<const uint2 *>(row1);
	const uint32_t b2 = row1[2];

	const uint32_t avgH0_0 = avg_u8x4(a01.x, a01.y), avgH0_1 = avg_u8x4(a01.y, a2);
	const uint32_t avgH1_0 = avg_u8x4(b01.x, b01.y), avgH1_1 = avg_u8x4(b01.y, b2);
	const uint32_t avgV0_0 = avg_u8x4(a01.x, b01.x), avgV0_1 = avg_u8x4(a01.y, b01.y);
	const uint32_t center0 = avg_u8x4(avgH0_0, avgH1_0), center1 = avg_u8x4(avgH0_1, avgH1_1);

	uint32_t *top = pDest + size_t(iY*2)*resX + sx*2;
	*reinterpret_cast<uint4 *>(top) = make_uint4(a01.x, avgH0_0, a01.y, avgH0_1);
	*reinterpret_cast<uint4 *>(top + resX) = make_uint4(avgV0_0, center0, avgV0_1, center1);
}

int ckd_fx_blit_2x2_rows(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, int y0, int y1)
{
	CKD_REQUIRE(ctx && d_dest && d_src, "null argument");
	CKD_REQUIRE(0 == (reinterpret_cast<uintptr_t>(d_dest) & 15) && 0 == (reinterpret_cast<uintptr_t>(d_src) & 15), "buffers must be 16-byte aligned (fx-blitter.cpp:29-30)");
	const int halfX = ctx->fxX - 4, halfY = ctx->fxY - 4;
	CKD_REQUIRE(0 <= y0 && y0 <= y1 && y1 <= halfY, "row range outside the FX map");
	if (y0 == y1)
		return CKD_OK;
	const dim3 block(64, 4);
	const dim3 grid(ckd_div_up(halfX/2, block.x), ckd_div_up(y1 - y0, block.y));
	ckd_prof_begin(ctx, "fx_blit_2x2", 5.0*ctx->resX*2.0*(y1 - y0));
	fx_blit_2x2_kernel<<<grid, block, 0, ctx->stream>>>(d_dest, d_src, ctx->fxX, ctx->resX, halfX, y0, y1);
	CKD_CHECK_LAUNCH(ctx);
	return CKD_OK;
}

extern "C" int ckd_fx_blit_2x2(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src)
{
	CKD_REQUIRE(ctx, "null argument");
	return ckd_fx_blit_2x2_rows(ctx, d_dest, d_src, 0, ctx->fxY - 4);
}

// ---------------------------------------------------------------------------------------------------------------
// blend ops -- util.cpp:83-812
// ---------------------------------------------------------------------------------------------------------------

struct BlendParams { uint32_t u0, u1; };

// SoftLightBlend, util.cpp:227-240 (A = src channel, B = dest channel; int arithmetic, returned as unsigned)
__device__ __forceinline__ unsigned soft_light(unsigned A, unsigned B)
{
	const int dA = int(A/2) + 64;
	if (B < 128)
		return unsigned((2*dA*int(B))/256);
	return unsigned(255 - (2*(255-dA)*(255-int(B)))/256);
}

// Overlay channel, util.cpp:450-452
__device__ __forceinline__ unsigned overlay_chan(unsigned bottom, unsigned top)
{
	return bottom < 128 ? (2*bottom*top/255) : (255 - 2*(255-bottom)*(255-top)/255);
}

// ---- the three colour channels of a pixel at once ---------------------------------------------------------------------------
// SoftLightBlend and the Overlay channel both have the shape  B < 128 ? f(a, B) : 255 - f(255 - a, 255 - B)  with f = 2*a*B/256
// (SoftLight, a = A/2 + 64) or 2*B*top/255 (Overlay).  255 - x of a byte is x ^ 0xff, so with m = 0xff in every byte whose B is
// >= 128 (one PRMT: sign replication) both branches become  f(a ^ m, B ^ m) ^ m  and everything except the per-channel product
// runs on all bytes of the word.  2*(B ^ m) <= 254 and a <= 255 keep a product below 2^16: its byte 1 is the >> 8, the /255 is
// the usual multiply-shift (exact below 2^16).  Checked against the scalar forms over 2 M random pixel pairs and the byte edge
// values before it went in; the golden and live post-op tests pin it on the device.
__device__ __forceinline__ uint32_t sign_bytes(uint32_t x)
{
	uint32_t m;
	asm("prmt.b32 %0, %1, 0, 0xba98;" : "=r"(m) : "r"(x));
	return m;
}

// (soft_light(R1, R2) << 16) | (soft_light(G1, G2) << 8) | soft_light(B1, B2), s = source pixel (A), d = destination pixel (B)
__device__ __forceinline__ uint32_t soft_light3(uint32_t s, uint32_t d)
{
	const uint32_t m = sign_bytes(d);
	const uint32_t a = (((s >> 1) & 0x7f7f7f7fu) + 0x40404040u) ^ m;   // A/2 + 64 per byte (<= 191: no carry), mirrored where B >= 128
	const uint32_t b2 = ((d ^ m) << 1) & 0xfefefefeu;                  // 2*B' per byte (B' <= 127)
	const uint32_t p0 = (a & 0xffu)*(b2 & 0xffu), p1 = ((a >> 8) & 0xffu)*((b2 >> 8) & 0xffu), p2 = ((a >> 16) & 0xffu)*((b2 >> 16) & 0xffu);
	const uint32_t t = __byte_perm(__byte_perm(p0, p1, 0x0051), p2, 0x7510); // byte 1 of each product; p2's byte 3 is zero
	return t ^ (m & 0x00ffffffu);
}

// (overlay_chan(R2, R1) << 16) | (overlay_chan(G2, G1) << 8) | overlay_chan(B2, B1), bottom = destination, top = source
__device__ __forceinline__ uint32_t overlay3(uint32_t bottom, uint32_t top)
{
	const uint32_t m = sign_bytes(bottom);
	const uint32_t b2 = ((bottom ^ m) << 1) & 0xfefefefeu, t = top ^ m;
	const uint32_t q0 = ((b2 & 0xffu)*(t & 0xffu)*0x8081u) >> 23;      // x/255 for x < 2^16
	const uint32_t q1 = (((b2 >> 8) & 0xffu)*((t >> 8) & 0xffu)*0x8081u) >> 23;
	const uint32_t q2 = (((b2 >> 16) & 0xffu)*((t >> 16) & 0xffu)*0x8081u) >> 23;
	return (q0 | (q1 << 8) | (q2 << 16)) ^ (m & 0x00ffffffu);
}

template <int OP> __device__ __forceinline__ uint32_t blend_px(uint32_t d, uint32_t s, const BlendParams &p)
{
	const unsigned A2 = d >> 24, R2 = (d >> 16) & 0xff, G2 = (d >> 8) & 0xff, B2 = d & 0xff;
	const unsigned A1 = s >> 24, R1 = (s >> 16) & 0xff, G1 = (s >> 8) & 0xff, B1 = s & 0xff;

	if (OP == CKD_MIX32 || OP == CKD_FADE32 || OP == CKD_MIXSRC32)
	{
		// (dest<<8 + alpha*(src-dest)) >> 8 in 16-bit lanes == (dest*(256-alpha) + src*alpha) >> 8  (util.cpp:92-95, 672-675, 808-810)
		const uint32_t alpha = (OP == CKD_MIXSRC32) ? A1 : p.u0;
		const uint32_t src = (OP == CKD_FADE32) ? p.u1 : s;
		return lerp8x4(d, src, alpha);
	}
	if (OP == CKD_ADD32) return adds_u8x4(d, s); // util.cpp:141-143
	if (OP == CKD_SUB32) return subs_u8x4(d, s); // util.cpp:188-190
	if (OP == CKD_MIXOVER32)
	{
		// util.cpp:148-177: c = ((c1*(0xff - iA1)) >> 8) + ((c2*iA1) >> 8) with iA1 = 0xff - A1; the two weights add up to 255, so the
		// sum stays <= 254 and the clamp at 255 of the reference cannot act.  R and B share a multiply (two 16-bit lanes).
		const uint32_t ia = 0xffu - A1;
		const uint32_t rb = ((((s & 0x00ff00ffu)*A1) >> 8) & 0x00ff00ffu) + ((((d & 0x00ff00ffu)*ia) >> 8) & 0x00ff00ffu);
		const uint32_t g = ((G1*A1) >> 8) + ((G2*ia) >> 8);
		return rb | (g << 8);
	}
	if (OP == CKD_EXCL32)
	{
		// util.cpp:194-223
		const unsigned R = R1 + R2 - ((2*R1*R2)>>8);
		const unsigned G = G1 + G2 - ((2*G1*G2)>>8);
		const unsigned B = B1 + B2 - ((2*B1*B2)>>8);
		return (A2<<24)|(R<<16)|(G<<8)|B;
	}
	if (OP == CKD_SOFTLIGHT32)
	{
		// util.cpp:242-271
		return (d & 0xff000000u) | soft_light3(s, d);
	}
	if (OP == CKD_SOFTLIGHT32A || OP == CKD_SOFTLIGHT32AA)
	{
		// util.cpp:274-346: the lerp runs in *unsigned* 32-bit arithmetic; wrapped bits spill into the neighbouring fields
		const uint32_t sl = soft_light3(s, d);
		unsigned R = (sl >> 16) & 0xffu, G = (sl >> 8) & 0xffu, B = sl & 0xffu;
		const unsigned a = (OP == CKD_SOFTLIGHT32A) ? A1 : p.u0;
		R = R2 + (((R-R2)*a)>>8);
		G = G2 + (((G-G2)*a)>>8);
		B = B2 + (((B-B2)*a)>>8);
		if (OP == CKD_SOFTLIGHT32A)
			return (R<<16)|(G<<8)|B;
		return (a<<24)|(R<<16)|(G<<8)|B;
	}
	if (OP == CKD_OVERLAY32)
	{
		// util.cpp:434-456
		return overlay3(d, s);
	}
	if (OP == CKD_OVERLAY32A)
	{
		// util.cpp:485-513
		const uint32_t ov = overlay3(d, s);
		const unsigned nR = (ov >> 16) & 0xffu, nG = (ov >> 8) & 0xffu, nB = ov & 0xffu;
		const unsigned R = R2 + (((nR-R2)*A1)>>8);
		const unsigned G = G2 + (((nG-G2)*A1)>>8);
		const unsigned B = B2 + (((nB-B2)*A1)>>8);
		return (R<<16)|(G<<8)|B;
	}
	if (OP == CKD_DARKEN32_50)
	{
		// util.cpp:518-549
		const unsigned R = (R2 + min(R1, R2))>>1, G = (G2 + min(G1, G2))>>1, B = (B2 + min(B1, B2))>>1;
		return (A2<<24)|(R<<16)|(G<<8)|B;
	}
	if (OP == CKD_MULSRC32)
	{
		// util.cpp:605-618: (src*dest)>>8 per channel, alpha included: byte 1 of each 16-bit product
		return __byte_perm(__byte_perm(B1*B2, G1*G2, 0x0051), __byte_perm(R1*R2, A1*A2, 0x0051), 0x5410);
	}
	if (OP == CKD_MULSRC32A)
	{
		// util.cpp:620-634: (srcAlpha*dest)>>8 per channel: two 16-bit lanes per multiply (a product is < 2^16: no carry between lanes)
		const uint32_t rb = (((d & 0x00ff00ffu)*A1) >> 8) & 0x00ff00ffu;
		const uint32_t ag = (((d >> 8) & 0x00ff00ffu)*A1) & 0xff00ff00u;
		return rb | ag;
	}
	return d;
}

template <int OP> __global__ void __launch_bounds__(256) blend_kernel(uint32_t *pDest, const uint32_t *pSrc, unsigned numQuads, unsigned numPixels, BlendParams p)
{
	// pDest/pSrc are not __restrict__: the reference allows them to alias exactly (same pointer)
	const unsigned stride = gridDim.x*blockDim.x;
	for (unsigned q = blockIdx.x*blockDim.x + threadIdx.x; q < numQuads; q += stride)
	{
		uint4 d = reinterpret_cast<const uint4 *>(pDest)[q];
		uint4 s = (OP == CKD_FADE32) ? d : reinterpret_cast<const uint4 *>(pSrc)[q];
		d.x = blend_px<OP>(d.x, s.x, p);
		d.y = blend_px<OP>(d.y, s.y, p);
		d.z = blend_px<OP>(d.z, s.z, p);
		d.w = blend_px<OP>(d.w, s.w, p);
		reinterpret_cast<uint4 *>(pDest)[q] = d;
	}
	// tail (numPixels not a multiple of 4)
	const unsigned tail = numQuads*4 + blockIdx.x*blockDim.x + threadIdx.x;
	if (tail < numPixels)
		pDest[tail] = blend_px<OP>(pDest[tail], (OP == CKD_FADE32) ? 0u : pSrc[tail], p);
}

// unaligned fallback (sub-rectangles of larger images): one pixel per thread
template <int OP> __global__ void __launch_bounds__(256) blend_kernel_scalar(uint32_t *pDest, const uint32_t *pSrc, unsigned numPixels, BlendParams p)
{
	const unsigned stride = gridDim.x*blockDim.x;
	for (unsigned i = blockIdx.x*blockDim.x + threadIdx.x; i < numPixels; i += stride)
		pDest[i] = blend_px<OP>(pDest[i], (OP == CKD_FADE32) ? 0u : pSrc[i], p);
}

template <int OP> static int LaunchBlend(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned numPixels, BlendParams p)
{
	if (0 == numPixels)
		return CKD_OK;
	const bool aligned = 0 == ((reinterpret_cast<uintptr_t>(d_dest) | reinterpret_cast<uintptr_t>(d_src)) & 15);
	const unsigned maxBlocks = unsigned(ctx->numSMs)*16;
	if (aligned)
	{
		const unsigned numQuads = numPixels/4;
		const unsigned blocks = std::max(1u, std::min(maxBlocks, ckd_div_up(std::max(numQuads, 1u), 256)));
		ckd_prof_begin(ctx, "blend", (OP == CKD_FADE32 ? 8.0 : 12.0)*numPixels);
		blend_kernel<OP><<<blocks, 256, 0, ctx->stream>>>(d_dest, d_src, numQuads, numPixels, p);
	}
	else
	{
		const unsigned blocks = std::max(1u, std::min(maxBlocks, ckd_div_up(numPixels, 256)));
		ckd_prof_begin(ctx, "blend", (OP == CKD_FADE32 ? 8.0 : 12.0)*numPixels);
		blend_kernel_scalar<OP><<<blocks, 256, 0, ctx->stream>>>(d_dest, d_src, numPixels, p);
	}
	CKD_CHECK_LAUNCH(ctx);
	return CKD_OK;
}

// per-op scalar parameters as the reference derives them from its arguments
static BlendParams MakeBlendParams(ckd_blend_op op, float f_param, unsigned u_param)
{
	BlendParams p = { 0, 0 };
	if (op == CKD_MIX32) p.u0 = u_param & 0xff;
	else if (op == CKD_SOFTLIGHT32AA) p.u0 = ckdh::x86_f2u(ckdh::saturatef(f_param)*255.f); // util.cpp:311-312: alpha = saturatef(alpha)*255.f; iA = unsigned(alpha)
	else if (op == CKD_FADE32) { p.u0 = u_param >> 24; p.u1 = u_param & 0xffffff; }
	return p;
}

// ---- fused chain: up to kMaxChain blend ops applied to every pixel in one pass --------------------------------------
// The compositor stacks 3-8 full-frame blends per frame (demo.cpp:511-1000); each is dest = f(dest, layer) per pixel, so a
// chain keeps the pixel in registers and reads every layer once: (layers + 2) x 4 bytes per pixel instead of 12 per op.
constexpr int kMaxChain = 8;
struct ChainArgs
{
	const uint32_t *src[kMaxChain];   // nullptr: the op reads the running pixel itself (Fade32, or a source aliasing the destination)
	BlendParams p[kMaxChain];
	int op[kMaxChain];
	int count;
};

// one step on 4 pixels: the switch is taken once per step (uniform across the grid), not once per pixel
#define CKD_CHAIN_CASE(OP) case OP: d.x = blend_px<OP>(d.x, v.x, p); d.y = blend_px<OP>(d.y, v.y, p); d.z = blend_px<OP>(d.z, v.z, p); d.w = blend_px<OP>(d.w, v.w, p); break;
__device__ __forceinline__ void chain_quad(int op, uint4 &d, const uint4 &v, const BlendParams &p)
{
	switch (op)
	{
	CKD_CHAIN_CASE(CKD_MIX32) CKD_CHAIN_CASE(CKD_MIXOVER32) CKD_CHAIN_CASE(CKD_ADD32) CKD_CHAIN_CASE(CKD_SUB32) CKD_CHAIN_CASE(CKD_EXCL32)
	CKD_CHAIN_CASE(CKD_SOFTLIGHT32) CKD_CHAIN_CASE(CKD_SOFTLIGHT32A) CKD_CHAIN_CASE(CKD_SOFTLIGHT32AA) CKD_CHAIN_CASE(CKD_OVERLAY32)
	CKD_CHAIN_CASE(CKD_OVERLAY32A) CKD_CHAIN_CASE(CKD_DARKEN32_50) CKD_CHAIN_CASE(CKD_MULSRC32) CKD_CHAIN_CASE(CKD_MULSRC32A)
	CKD_CHAIN_CASE(CKD_MIXSRC32)
	default: d.x = blend_px<CKD_FADE32>(d.x, v.x, p); d.y = blend_px<CKD_FADE32>(d.y, v.y, p); d.z = blend_px<CKD_FADE32>(d.z, v.z, p); d.w = blend_px<CKD_FADE32>(d.w, v.w, p); break;
	}
}
#undef CKD_CHAIN_CASE

__global__ void __launch_bounds__(256, 4) blend_chain_kernel(uint32_t *pDest, unsigned numQuads, unsigned numPixels, const ChainArgs args)
{
	const unsigned stride = gridDim.x*blockDim.x;
	for (unsigned q = blockIdx.x*blockDim.x + threadIdx.x; q < numQuads; q += stride)
	{
		uint4 d = reinterpret_cast<const uint4 *>(pDest)[q];
		// the first layers are requested before any arithmetic (they do not depend on the running pixel)
		uint4 s[4];
		#pragma unroll
		for (int k = 0; k < 4; ++k)
			if (k < args.count && nullptr != args.src[k])
				s[k] = __ldg(reinterpret_cast<const uint4 *>(args.src[k]) + q);
		#pragma unroll 1
		for (int k = 0; k < args.count; ++k)
		{
			uint4 v = d;
			if (nullptr != args.src[k])
				v = (k < 4) ? ((k == 0) ? s[0] : (k == 1) ? s[1] : (k == 2) ? s[2] : s[3]) : __ldg(reinterpret_cast<const uint4 *>(args.src[k]) + q);
			chain_quad(args.op[k], d, v, args.p[k]);
		}
		reinterpret_cast<uint4 *>(pDest)[q] = d;
	}
	const unsigned tail = numQuads*4 + blockIdx.x*blockDim.x + threadIdx.x;
	if (tail < numPixels)
	{
		uint4 d = make_uint4(pDest[tail], 0, 0, 0);
		for (int k = 0; k < args.count; ++k)
		{
			const uint4 v = make_uint4((nullptr != args.src[k]) ? args.src[k][tail] : d.x, 0, 0, 0);
			chain_quad(args.op[k], d, v, args.p[k]);
		}
		pDest[tail] = d.x;
	}
}

extern "C" int ckd_blend_chain(ckd_ctx *ctx, uint32_t *d_dest, const ckd_blend_step *steps, unsigned num_steps, unsigned num_pixels)
{
	CKD_REQUIRE(ctx && d_dest && (steps || 0 == num_steps), "null argument");
	if (0 == num_pixels)
		return CKD_OK;
	bool aligned = 0 == (reinterpret_cast<uintptr_t>(d_dest) & 15);
	for (unsigned i = 0; i < num_steps; ++i)
	{
		CKD_REQUIRE(unsigned(steps[i].op) <= unsigned(CKD_FADE32), "unknown blend op");
		CKD_REQUIRE(steps[i].op == CKD_FADE32 || steps[i].d_src, "null source");
		aligned = aligned && 0 == (reinterpret_cast<uintptr_t>(steps[i].d_src) & 15);
	}
	if (!aligned)
	{
		// unaligned sub-rectangles: one ckd_blend per step
		for (unsigned i = 0; i < num_steps; ++i)
			CKD_TRY(ckd_blend(ctx, steps[i].op, d_dest, steps[i].d_src, num_pixels, steps[i].f_param, steps[i].u_param));
		return CKD_OK;
	}

	for (unsigned first = 0; first < num_steps; first += kMaxChain)
	{
		ChainArgs args;
		memset(&args, 0, sizeof(args));
		args.count = int(std::min<unsigned>(kMaxChain, num_steps - first));
		int layers = 0;
		for (int k = 0; k < args.count; ++k)
		{
			const ckd_blend_step &step = steps[first + k];
			const bool self = step.op == CKD_FADE32 || step.d_src == d_dest;
			args.src[k] = self ? nullptr : step.d_src;
			args.op[k] = int(step.op);
			args.p[k] = MakeBlendParams(step.op, step.f_param, step.u_param);
			layers += self ? 0 : 1;
		}
		const unsigned numQuads = num_pixels/4;
		const unsigned blocks = std::max(1u, std::min(unsigned(ctx->numSMs)*16, ckd_div_up(std::max(numQuads, 1u), 256)));
		ckd_prof_begin(ctx, "blend_chain", 4.0*(layers + 2)*num_pixels);
		blend_chain_kernel<<<blocks, 256, 0, ctx->stream>>>(d_dest, numQuads, num_pixels, args);
		CKD_CHECK_LAUNCH(ctx);
	}
	return CKD_OK;
}

extern "C" int ckd_blend(ckd_ctx *ctx, ckd_blend_op op, uint32_t *d_dest, const uint32_t *d_src, unsigned num_pixels, float f_param, unsigned u_param)
{
	CKD_REQUIRE(ctx && d_dest, "null argument");
	CKD_REQUIRE(op == CKD_FADE32 || d_src, "null source");
	if (op == CKD_FADE32)
		d_src = d_dest;

	BlendParams p = MakeBlendParams(op, f_param, u_param);
	switch (op)
	{
	case CKD_MIX32:         return LaunchBlend<CKD_MIX32>(ctx, d_dest, d_src, num_pixels, p);
	case CKD_MIXOVER32:     return LaunchBlend<CKD_MIXOVER32>(ctx, d_dest, d_src, num_pixels, p);
	case CKD_ADD32:         return LaunchBlend<CKD_ADD32>(ctx, d_dest, d_src, num_pixels, p);
	case CKD_SUB32:         return LaunchBlend<CKD_SUB32>(ctx, d_dest, d_src, num_pixels, p);
	case CKD_EXCL32:        return LaunchBlend<CKD_EXCL32>(ctx, d_dest, d_src, num_pixels, p);
	case CKD_SOFTLIGHT32:   return LaunchBlend<CKD_SOFTLIGHT32>(ctx, d_dest, d_src, num_pixels, p);
	case CKD_SOFTLIGHT32A:  return LaunchBlend<CKD_SOFTLIGHT32A>(ctx, d_dest, d_src, num_pixels, p);
	case CKD_SOFTLIGHT32AA: return LaunchBlend<CKD_SOFTLIGHT32AA>(ctx, d_dest, d_src, num_pixels, p);
	case CKD_OVERLAY32:     return LaunchBlend<CKD_OVERLAY32>(ctx, d_dest, d_src, num_pixels, p);
	case CKD_OVERLAY32A:    return LaunchBlend<CKD_OVERLAY32A>(ctx, d_dest, d_src, num_pixels, p);
	case CKD_DARKEN32_50:   return LaunchBlend<CKD_DARKEN32_50>(ctx, d_dest, d_src, num_pixels, p);
	case CKD_MULSRC32:      return LaunchBlend<CKD_MULSRC32>(ctx, d_dest, d_src, num_pixels, p);
	case CKD_MULSRC32A:     return LaunchBlend<CKD_MULSRC32A>(ctx, d_dest, d_src, num_pixels, p);
	case CKD_MIXSRC32:      return LaunchBlend<CKD_MIXSRC32>(ctx, d_dest, d_src, num_pixels, p);
	case CKD_FADE32:        return LaunchBlend<CKD_FADE32>(ctx, d_dest, d_src, num_pixels, p);
	}
	CKD_REQUIRE(false, "unknown blend op");
}

// ---------------------------------------------------------------------------------------------------------------
// rectangular blits -- util.cpp:636-659 (MixSrc32S), 707-796 (BlitSrc32/A, BlitAdd32/A)
// ---------------------------------------------------------------------------------------------------------------

enum { kRectMixSrc = 0, kRectBlitSrcA = 1, kRectBlitAdd = 2, kRectBlitAddA = 3 };

// fa = the four bytes of 0x01010101*unsigned(alpha*255.f) (c2vISSE16 of that word: one 16-bit lane per byte)
template <int OP> __device__ __forceinline__ uint32_t rect_px(uint32_t d, uint32_t s, uint32_t fa)
{
	if (OP == kRectMixSrc)
	{
		const uint32_t a = s >> 24;
		return lerp8x4(d, s, a);
	}
	if (OP == kRectBlitAdd)
		return adds_u8x4(d, s);

	uint32_t out = 0;
	const uint32_t srcA = s >> 24;
	#pragma unroll
	for (int c = 0; c < 4; ++c)
	{
		const uint32_t f = (fa >> (8*c)) & 0xff;
		const uint32_t sc = (s >> (8*c)) & 0xff, dc = (d >> (8*c)) & 0xff;
		uint32_t r;
		if (OP == kRectBlitSrcA)
		{
			const uint32_t a = (srcA*f) >> 8;                 // util.cpp:743
			r = (((dc << 8) + a*(sc - dc)) & 0xffff) >> 8;    // 16-bit lane arithmetic, util.cpp:745-746
			r = min(r, 255u);
		}
		else
		{
			const uint32_t mod = (sc*f) >> 8;                 // util.cpp:788
			r = min(dc + mod, 255u);                          // add_epi16 + packus
		}
		out |= r << (8*c);
	}
	return out;
}

template <int OP> __global__ void __launch_bounds__(256) rect_kernel(uint32_t *pDest, const uint32_t *pSrc, unsigned destStride, unsigned srcStride, unsigned width, unsigned height, uint32_t fa)
{
	const unsigned x = blockIdx.x*blockDim.x + threadIdx.x;
	const unsigned y = blockIdx.y*blockDim.y + threadIdx.y;
	if (x >= width || y >= height)
		return;
	const size_t di = size_t(y)*destStride + x;
	pDest[di] = rect_px<OP>(pDest[di], pSrc[size_t(y)*srcStride + x], fa);
}

// four pixels per thread with 128-bit accesses: rectangles whose rows start 16-byte aligned in both buffers (most of the
// compositor's sprites at 4K: widths and strides are multiples of 4)
template <int OP> __global__ void __launch_bounds__(256) rect_kernel4(uint32_t *pDest, const uint32_t *pSrc, unsigned destStride, unsigned srcStride, unsigned quadsPerRow, unsigned height, uint32_t fa)
{
	const unsigned q = blockIdx.x*blockDim.x + threadIdx.x;
	const unsigned y = blockIdx.y*blockDim.y + threadIdx.y;
	if (q >= quadsPerRow || y >= height)
		return;
	uint4 *dp = reinterpret_cast<uint4 *>(pDest + size_t(y)*destStride) + q;
	const uint4 sv = *(reinterpret_cast<const uint4 *>(pSrc + size_t(y)*srcStride) + q);
	uint4 dv = *dp;
	dv.x = rect_px<OP>(dv.x, sv.x, fa); dv.y = rect_px<OP>(dv.y, sv.y, fa); dv.z = rect_px<OP>(dv.z, sv.z, fa); dv.w = rect_px<OP>(dv.w, sv.w, fa);
	*dp = dv;
}

template <int OP> static int LaunchRect(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned destStride, unsigned srcStride, unsigned width, unsigned height, uint32_t fa)
{
	if (0 == width || 0 == height)
		return CKD_OK;
	if (0 == ((width | destStride | srcStride) & 3u) && 0 == ((reinterpret_cast<uintptr_t>(d_dest) | reinterpret_cast<uintptr_t>(d_src)) & 15u))
	{
		const dim3 block4(64, 4);
		const dim3 grid4(ckd_div_up(width/4, block4.x), ckd_div_up(height, block4.y));
		ckd_prof_begin(ctx, "rect_blit", 12.0*width*height);
		rect_kernel4<OP><<<grid4, block4, 0, ctx->stream>>>(d_dest, d_src, destStride, srcStride, width/4, height, fa);
		CKD_CHECK_LAUNCH(ctx);
		return CKD_OK;
	}
	const dim3 block(64, 4);
	const dim3 grid(ckd_div_up(width, block.x), ckd_div_up(height, block.y));
	ckd_prof_begin(ctx, "rect_blit", 12.0*width*height);
	rect_kernel<OP><<<grid, block, 0, ctx->stream>>>(d_dest, d_src, destStride, srcStride, width, height, fa);
	CKD_CHECK_LAUNCH(ctx);
	return CKD_OK;
}

extern "C" int ckd_blit(ckd_ctx *ctx, ckd_blit_op op, uint32_t *d_dest, const uint32_t *d_src, unsigned dest_res_x, unsigned src_res_x, unsigned y_res, float alpha)
{
	CKD_REQUIRE(ctx && d_dest && d_src, "null argument");
	const uint32_t fa = 0x01010101u * ckdh::x86_f2u(alpha*255.f); // util.cpp:734, 778
	switch (op)
	{
	case CKD_BLITSRC32:  return LaunchRect<kRectMixSrc>(ctx, d_dest, d_src, dest_res_x, src_res_x, src_res_x, y_res, 0);
	case CKD_BLITSRC32A: return LaunchRect<kRectBlitSrcA>(ctx, d_dest, d_src, dest_res_x, src_res_x, src_res_x, y_res, fa);
	case CKD_BLITADD32:  return LaunchRect<kRectBlitAdd>(ctx, d_dest, d_src, dest_res_x, src_res_x, src_res_x, y_res, 0);
	case CKD_BLITADD32A: return LaunchRect<kRectBlitAddA>(ctx, d_dest, d_src, dest_res_x, src_res_x, src_res_x, y_res, fa);
	}
	CKD_REQUIRE(false, "unknown blit op");
}

extern "C" int ckd_mix_src_s(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned dest_res_x, unsigned dest_res_y, unsigned src_stride)
{
	CKD_REQUIRE(ctx && d_dest && d_src, "null argument");
	return LaunchRect<kRectMixSrc>(ctx, d_dest, d_src, dest_res_x, src_stride, dest_res_x, dest_res_y, 0);
}

// ---------------------------------------------------------------------------------------------------------------
// memset32 -- util.h:57-67
// ---------------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) memset32_kernel(uint32_t *__restrict__ pDest, uint32_t value, size_t numQuads, size_t numInts)
{
	const size_t stride = size_t(gridDim.x)*blockDim.x;
	const uint4 v = make_uint4(value, value, value, value);
	for (size_t q = size_t(blockIdx.x)*blockDim.x + threadIdx.x; q < numQuads; q += stride)
		reinterpret_cast<uint4 *>(pDest)[q] = v;
	const size_t tail = numQuads*4 + size_t(blockIdx.x)*blockDim.x + threadIdx.x;
	if (tail < numInts)
		pDest[tail] = value;
}

extern "C" int ckd_memset32(ckd_ctx *ctx, uint32_t *d_dest, uint32_t value, size_t num_ints)
{
	CKD_REQUIRE(ctx && d_dest, "null argument");
	CKD_REQUIRE(0 == (reinterpret_cast<uintptr_t>(d_dest) & 15), "destination must be 16-byte aligned");
	if (0 == num_ints)
		return CKD_OK;
	const size_t numQuads = num_ints/4;
	const unsigned blocks = unsigned(std::max<size_t>(1, std::min<size_t>(size_t(ctx->numSMs)*16, ckd_div_up(std::max<size_t>(numQuads, 1), 256))));
	ckd_prof_begin(ctx, "memset32", 4.0*num_ints);
	memset32_kernel<<<blocks, 256, 0, ctx->stream>>>(d_dest, value, numQuads, num_ints);
	CKD_CHECK_LAUNCH(ctx);
	return CKD_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// polar remap -- polar.cpp:82-198
// ---------------------------------------------------------------------------------------------------------------
// One thread per 4 destination pixels: 2x 128-bit map loads (8 B/px, the dominant stream), 16 texel gathers from the
// source (L2 resident at 33 MB), one 128-bit store.  ALPHA adds a 128-bit read of the destination.

__device__ __forceinline__ uint32_t polar_fetch(const uint32_t *__restrict__ pSrc, int U, int V, unsigned resX)
{
	// Fetch32/Fetch16, polar.cpp:82-106
	const unsigned U0 = unsigned(U >> 8);
	const unsigned V0 = unsigned(V >> 8)*resX;
	const uint32_t fu = U & 0xff, fv = V & 0xff;
	const uint32_t *p = pSrc + U0 + V0;
	const uint32_t s0 = __ldg(p), s1 = __ldg(p + 1), s2 = __ldg(p + resX), s3 = __ldg(p + resX + 1);
	return bilerp_argb(s0, s1, s2, s3, fu, fv);
}

__device__ __forceinline__ uint32_t polar_blend(uint32_t d, uint32_t s)
{
	// Polar_Blit_TileA, polar.cpp:169-174: lerp every channel by the fetched alpha
	const uint32_t a = s >> 24;
	return lerp8x4(d, s, a);
}

// HALO: SoftLight32A(pDest, pHalo) (util.cpp:274-346) applied to the remapped pixel before it is stored -- the ball's halo layer
// (ball.cpp:357-361) in the same pass instead of a second read-modify-write of the frame
// TILED: a warp covers 32 pixels x 4 rows instead of 128 pixels of one row (a block 64 x 16): the footprint of its gathers in the
// source is a compact patch instead of a long arc, fewer distinct sectors per gather instruction (the kernel is L1TEX bound)
template <bool ALPHA, bool HALO, int TILED> __global__ void __launch_bounds__(256) polar_blit_kernel(uint32_t *pDest, const uint32_t *__restrict__ pSrc, const int4 *__restrict__ pMap, unsigned numQuads, unsigned resX,
	const uint4 *__restrict__ pHalo)
{
	// TILED = 0: linear; else log2 of the quads a warp covers per row + 1 (5: 16 quads x 2 rows, 4: 8 x 4, 3: 4 x 8, 2: 2 x 16, 1: 1 x 32).
	// Measured at 4K (profiles/r02_polar_variants.txt): linear 45.7 us, 8 x 4 37.5, 4 x 8 37.2, 2 x 16 39.6, 1 x 32 54.0
	unsigned q;
	if (TILED)
	{
		constexpr unsigned QW = 1u << (TILED > 0 ? TILED - 1 : 0), RW = 32u/QW; // quads per warp row, rows per warp
		const unsigned quadsPerRow = resX >> 2, rows = numQuads/quadsPerRow;
		const unsigned tilesX = (quadsPerRow + 2*QW - 1)/(2*QW);
		const unsigned bx = blockIdx.x % tilesX, by = blockIdx.x / tilesX;
		const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		const unsigned qx = bx*(2*QW) + (warp & 1)*QW + (lane & (QW-1)), y = by*(4*RW) + (warp >> 1)*RW + lane/QW;
		if (qx >= quadsPerRow || y >= rows)
			return;
		q = y*quadsPerRow + qx;
	}
	else
	{
		q = blockIdx.x*blockDim.x + threadIdx.x;
		if (q >= numQuads)
			return;
	}
	const int4 m0 = __ldg(pMap + size_t(q)*2), m1 = __ldg(pMap + size_t(q)*2 + 1);
	uint4 halo = make_uint4(0, 0, 0, 0);
	if (HALO) halo = __ldg(pHalo + q);
	uint4 out;
	out.x = polar_fetch(pSrc, m0.x, m0.y, resX);
	out.y = polar_fetch(pSrc, m0.z, m0.w, resX);
	out.z = polar_fetch(pSrc, m1.x, m1.y, resX);
	out.w = polar_fetch(pSrc, m1.z, m1.w, resX);
	if (ALPHA)
	{
		const uint4 d = reinterpret_cast<const uint4 *>(pDest)[q];
		out.x = polar_blend(d.x, out.x);
		out.y = polar_blend(d.y, out.y);
		out.z = polar_blend(d.z, out.z);
		out.w = polar_blend(d.w, out.w);
	}
	if (HALO)
	{
		const BlendParams none = { 0u, 0u };
		out.x = blend_px<CKD_SOFTLIGHT32A>(out.x, halo.x, none);
		out.y = blend_px<CKD_SOFTLIGHT32A>(out.y, halo.y, none);
		out.z = blend_px<CKD_SOFTLIGHT32A>(out.z, halo.z, none);
		out.w = blend_px<CKD_SOFTLIGHT32A>(out.w, halo.w, none);
	}
	reinterpret_cast<uint4 *>(pDest)[q] = out;
}

// one destination pixel per thread: neighbouring lanes fetch neighbouring texels (the polar map is smooth), so a warp's
// gather touches a handful of sectors instead of 32
template <bool ALPHA> __global__ void __launch_bounds__(256) polar_blit_kernel_1px(uint32_t *pDest, const uint32_t *__restrict__ pSrc, const int2 *__restrict__ pMap, unsigned numPixels, unsigned resX)
{
	const unsigned i = blockIdx.x*blockDim.x + threadIdx.x;
	if (i >= numPixels)
		return;
	const int2 m = __ldg(pMap + i);
	uint32_t out = polar_fetch(pSrc, m.x, m.y, resX);
	if (ALPHA)
		out = polar_blend(pDest[i], out);
	pDest[i] = out;
}

constexpr int kPolarTile = 4; // a warp covers 8 quads (32 pixels) x 4 rows

// grid of the tiled kernel for `rows` rows of quadsPerRow quads
template <int T> static unsigned PolarTiles(unsigned quadsPerRow, unsigned rows)
{
	constexpr unsigned QW = 1u << (T - 1), RW = 32u/QW;
	return ((quadsPerRow + 2*QW - 1)/(2*QW))*((rows + 4*RW - 1)/(4*RW));
}

// one remap of `rows` whole rows starting at the given pointers (the banded tail passes row bands)
static void LaunchPolarRows(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, const int4 *pMap, unsigned resX, unsigned rows, bool alpha, const uint32_t *d_halo)
{
	const unsigned quadsPerRow = resX/4, numQuads = quadsPerRow*rows;
	static const int variant = getenv("CKD_POLAR_VARIANT") ? atoi(getenv("CKD_POLAR_VARIANT")) : 0; // tuning: 2 linear, 4..8 tile shapes
	const uint4 *halo = reinterpret_cast<const uint4 *>(d_halo);
	#define CKD_POLAR(T, GRID) { if (halo) polar_blit_kernel<true, true, T><<<GRID, 256, 0, ctx->stream>>>(d_dest, d_src, pMap, numQuads, resX, halo); \
		else if (alpha) polar_blit_kernel<true, false, T><<<GRID, 256, 0, ctx->stream>>>(d_dest, d_src, pMap, numQuads, resX, nullptr); \
		else polar_blit_kernel<false, false, T><<<GRID, 256, 0, ctx->stream>>>(d_dest, d_src, pMap, numQuads, resX, nullptr); }
	switch (variant)
	{
	case 2: CKD_POLAR(0, ckd_div_up(numQuads, 256)) break;
	case 5: CKD_POLAR(3, PolarTiles<3>(quadsPerRow, rows)) break;
	case 6: CKD_POLAR(2, PolarTiles<2>(quadsPerRow, rows)) break;
	case 7: CKD_POLAR(1, PolarTiles<1>(quadsPerRow, rows)) break;
	case 8: CKD_POLAR(5, PolarTiles<5>(quadsPerRow, rows)) break;
	default: CKD_POLAR(kPolarTile, PolarTiles<kPolarTile>(quadsPerRow, rows)) break;
	}
	#undef CKD_POLAR
}

static int LaunchPolar(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, int inverse, bool alpha)
{
	CKD_REQUIRE(ctx && d_dest && d_src, "null argument");
	CKD_REQUIRE(d_dest != d_src, "polar blit cannot run in place");
	CKD_REQUIRE(0 == (reinterpret_cast<uintptr_t>(d_dest) & 15), "destination must be 16-byte aligned");
	const unsigned numQuads = unsigned(size_t(ctx->resX)*ctx->resY/4);
	const int4 *pMap = reinterpret_cast<const int4 *>(inverse ? ctx->d_polarInvMap : ctx->d_polarMap);
	ckd_prof_begin(ctx, alpha ? "polar_blit_a" : "polar_blit", (alpha ? 20.0 : 16.0)*ctx->resX*ctx->resY);
	static const int variant = getenv("CKD_POLAR_VARIANT") ? atoi(getenv("CKD_POLAR_VARIANT")) : 0;
	if (variant == 1 || 0 != (ctx->resX & 3))
	{
		const unsigned numPixels = numQuads*4;
		const int2 *pMap2 = reinterpret_cast<const int2 *>(pMap);
		if (alpha)
			polar_blit_kernel_1px<true><<<ckd_div_up(numPixels, 256), 256, 0, ctx->stream>>>(d_dest, d_src, pMap2, numPixels, unsigned(ctx->resX));
		else
			polar_blit_kernel_1px<false><<<ckd_div_up(numPixels, 256), 256, 0, ctx->stream>>>(d_dest, d_src, pMap2, numPixels, unsigned(ctx->resX));
	}
	else
		LaunchPolarRows(ctx, d_dest, d_src, pMap, unsigned(ctx->resX), unsigned(ctx->resY), alpha, nullptr);
	CKD_CHECK_LAUNCH(ctx);
	return CKD_OK;
}

// Polar_Blit / Polar_BlitA (+ the ball's SoftLight32A halo) as the LAST stage of a frame.  With a banded read-back armed
// (ckd_arm_readback) the remap is issued per band of destination rows and every finished band leaves for the host on the
// copy stream while the next one is remapped; the source is the complete render target, so bands are independent.
int ckd_polar_tail(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, int inverse, bool alpha, const uint32_t *d_softLightSrc)
{
	void *h_dest = ctx->rbHost;
	ctx->rbHost = nullptr;
	const unsigned numPixels = unsigned(ctx->resX)*unsigned(ctx->resY);
	const int4 *pMap = reinterpret_cast<const int4 *>(inverse ? ctx->d_polarInvMap : ctx->d_polarMap);
	// the halo rides in the remap kernel when the shapes allow it (alpha remap, 16-byte aligned layers)
	const bool fuseHalo = nullptr != d_softLightSrc && alpha && 0 == (ctx->resX & 3) && 0 == ((reinterpret_cast<uintptr_t>(d_softLightSrc) | reinterpret_cast<uintptr_t>(d_dest)) & 15);
	if (nullptr == h_dest || 0 != (ctx->resX & 3))
	{
		if (fuseHalo)
		{
			CKD_REQUIRE(d_dest != d_src, "polar blit cannot run in place");
			ckd_prof_begin(ctx, "polar_blit_a_halo", 24.0*numPixels);
			LaunchPolarRows(ctx, d_dest, d_src, pMap, unsigned(ctx->resX), unsigned(ctx->resY), true, d_softLightSrc);
			CKD_CHECK_LAUNCH(ctx);
			return CKD_OK;
		}
		CKD_TRY(LaunchPolar(ctx, d_dest, d_src, inverse, alpha));
		return d_softLightSrc ? ckd_blend(ctx, CKD_SOFTLIGHT32A, d_dest, d_softLightSrc, numPixels, 0.f, 0) : CKD_OK;
	}
	CKD_REQUIRE(d_dest != d_src, "polar blit cannot run in place");
	const int bands = ctx->rbBands;
	for (int k = 0; k < bands; ++k)
	{
		const int y0 = ctx->resY*k/bands, y1 = ctx->resY*(k + 1)/bands;
		if (y1 <= y0)
			continue;
		const size_t offset = size_t(y0)*ctx->resX, count = size_t(y1 - y0)*ctx->resX;
		ckd_prof_begin(ctx, fuseHalo ? "polar_blit_a_halo" : alpha ? "polar_blit_a" : "polar_blit", (fuseHalo ? 24.0 : alpha ? 20.0 : 16.0)*double(count));
		LaunchPolarRows(ctx, d_dest + offset, d_src, pMap + offset/2, unsigned(ctx->resX), unsigned(y1 - y0), alpha, fuseHalo ? d_softLightSrc + offset : nullptr);
		CKD_CHECK_LAUNCH(ctx);
		if (d_softLightSrc && !fuseHalo)
			CKD_TRY(ckd_blend(ctx, CKD_SOFTLIGHT32A, d_dest + offset, d_softLightSrc + offset, unsigned(count), 0.f, 0));
		CKD_CUDA(cudaEventRecord(ctx->evBand[k], ctx->stream));
		CKD_CUDA(cudaStreamWaitEvent(ctx->copyStream, ctx->evBand[k], 0));
		CKD_CUDA(cudaMemcpyAsync(static_cast<uint32_t *>(h_dest) + offset, d_dest + offset, count*sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->copyStream));
	}
	ctx->rbIssued = true;
	return CKD_OK;
}

extern "C" int ckd_polar_blit(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, int inverse) { return LaunchPolar(ctx, d_dest, d_src, inverse, false); }
extern "C" int ckd_polar_blit_a(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, int inverse) { return LaunchPolar(ctx, d_dest, d_src, inverse, true); }

// Polar_Blit_2x2, polar.cpp:200-218: the same remap on FX-map sized buffers (row stride fxResX) through the FX-map sized maps
extern "C" int ckd_polar_blit_2x2(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, int inverse)
{
	CKD_REQUIRE(ctx && d_dest && d_src, "null argument");
	CKD_REQUIRE(d_dest != d_src, "polar blit cannot run in place");
	CKD_REQUIRE(0 == (reinterpret_cast<uintptr_t>(d_dest) & 15), "destination must be 16-byte aligned");
	CKD_TRY(ckd_ensure_polar_maps_2x2(ctx));
	const int4 *pMap = reinterpret_cast<const int4 *>(inverse ? ctx->d_polarInvMap2x2 : ctx->d_polarMap2x2); // fxResX is a multiple of 4 (fx-blitter.h:18)
	ckd_prof_begin(ctx, "polar_blit_2x2", 16.0*ctx->fxX*ctx->fxY);
	LaunchPolarRows(ctx, d_dest, d_src, pMap, unsigned(ctx->fxX), unsigned(ctx->fxY), false, nullptr);
	CKD_CHECK_LAUNCH(ctx);
	return CKD_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// TapeWarp32 -- util.cpp:552-603
// ---------------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) tape_warp_kernel(uint32_t *__restrict__ pDest, const uint32_t *__restrict__ pSrc, const float2 *__restrict__ g_lut2,
	unsigned xRes, unsigned yRes, unsigned globalResX, float strength, float speed)
{
	// 2 LUT reads per pixel and 1 KB of output per block: the 16 KB table is read through L1 instead of being staged
	const float2 *s_lut2 = g_lut2;

	const unsigned iX = blockIdx.x*blockDim.x + threadIdx.x;
	const int iY = int(blockIdx.y*blockDim.y + threadIdx.y);
	if (iX >= xRes || unsigned(iY) >= yRes)
		return;

	const float dX = lutsinf(s_lut2, float(iY)*speed)*strength*1.f;
	const float dY = lutcosf(s_lut2, float(iX)*speed)*strength*1.f;
	float tX = float(iX) + dX;
	float tY = float(iY) + dY;

	if (tX < 0.f) tX = 0.f;
	else if (tX >= float(xRes)-1.f) tX = float(xRes) - 2.f;
	if (tY < 0.f) tY = 0.f;
	else if (tY >= float(yRes)-1.f) tY = float(yRes) - 2.f;

	const int U = ftofp24(tX), V = ftofp24(tY);
	const unsigned U0 = unsigned(U >> 8);
	const unsigned V0 = unsigned(V >> 8)*globalResX; // the reference strides by kResX, not xRes (util.cpp:594)
	const uint32_t *p = pSrc + U0 + V0;
	const uint32_t s0 = __ldg(p), s1 = __ldg(p + 1), s2 = __ldg(p + globalResX), s3 = __ldg(p + globalResX + 1);
	pDest[size_t(iY)*xRes + iX] = bilerp_argb(s0, s1, s2, s3, U & 0xff, V & 0xff);
}

extern "C" int ckd_tape_warp(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength, float speed)
{
	CKD_REQUIRE(ctx && d_dest && d_src, "null argument");
	CKD_REQUIRE(d_dest != d_src, "TapeWarp32 cannot run in place");
	const dim3 block(64, 4);
	const dim3 grid(ckd_div_up(x_res, block.x), ckd_div_up(y_res, block.y));
	ckd_prof_begin(ctx, "tape_warp", 8.0*x_res*y_res);
	tape_warp_kernel<<<grid, block, 0, ctx->stream>>>(d_dest, d_src, ctx->d_cosLUT2, x_res, y_res, unsigned(ctx->resX), strength, speed);
	CKD_CHECK_LAUNCH(ctx);
	return CKD_OK;
}

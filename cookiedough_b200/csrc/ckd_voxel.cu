// ckd_voxel.cu -- voxel ray casters: landscape, tunnelscape, ball (beams / no beams), torus twister.
//
// The reference walks every ray front to back, one step at a time, keeping the highest (or lowest) projected height
// drawn so far and emitting a Gouraud span (cspanISSE16) whenever a step pokes out above it.  Here a WARP owns a ray
// and its 32 lanes own 32 consecutive steps: every lane samples its step (bilinear height + colour gathers out of
// L2-resident maps), a warp prefix-min/max over the projected heights gives each lane the occlusion horizon it would
// have seen in the serial loop, the visible lanes emit their spans (short ones per lane, long ones warp-cooperatively),
// and the carry (horizon, previous height/colour, beam accumulator) moves on to the next 32 steps.  Spans are written
// into a shared-memory line buffer pre-filled with the clear colour, so the reference's memset32 + scattered span
// stores become exactly one coalesced write of every output pixel.

#include "ckd_internal.h"
#include "ckd_math.cuh"
#include "ckd_hostmath.h"

using namespace ckd;

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct Color16 { int c[4]; }; // 16-bit lanes B,G,R,A of an unpacked pixel (c2vISSE16, util.h:125-127)

__device__ __forceinline__ Color16 unpack16(uint32_t px)
{
	Color16 r;
	r.c[0] = int(px & 0xff); r.c[1] = int((px >> 8) & 0xff); r.c[2] = int((px >> 16) & 0xff); r.c[3] = int(px >> 24);
	return r;
}

__device__ __forceinline__ Color16 shfl_color(const Color16 &v, int srcLane)
{
	Color16 r;
	#pragma unroll
	for (int i = 0; i < 4; ++i) r.c[i] = __shfl_sync(kFull, v.c[i], srcLane);
	return r;
}

__device__ __forceinline__ Color16 shfl_up_color(const Color16 &v)
{
	Color16 r;
	#pragma unroll
	for (int i = 0; i < 4; ++i) r.c[i] = __shfl_up_sync(kFull, v.c[i], 1);
	return r;
}

// ---- texture sampling (bilinear.h) ----------------------------------------------------------------------------------

struct TexCoords { unsigned i00, i10, i01, i11; uint32_t fu, fv; };

// bsamp_prepUVs, bilinear.h:10-31
__device__ __forceinline__ TexCoords prep_uvs(int U, int V, unsigned mapAnd, unsigned mapShift)
{
	unsigned U0 = unsigned(U >> 8), V0 = unsigned(V >> 8);
	unsigned U1 = U0 + 1, V1 = V0 + 1;
	U0 &= mapAnd; V0 = (V0 & mapAnd) << mapShift;
	U1 &= mapAnd; V1 = (V1 & mapAnd) << mapShift;
	TexCoords t;
	t.i00 = U0+V0; t.i10 = U1+V0; t.i01 = U0+V1; t.i11 = U1+V1;
	t.fu = U & 0xff; t.fv = V & 0xff;
	return t;
}

__device__ __forceinline__ unsigned sample_u8(const uint8_t *__restrict__ tex, const TexCoords &t)
{
	return bilerp_u8(__ldg(tex + t.i00), __ldg(tex + t.i10), __ldg(tex + t.i01), __ldg(tex + t.i11), int(t.fu), int(t.fv));
}

__device__ __forceinline__ uint32_t sample_argb(const uint32_t *__restrict__ tex, const TexCoords &t)
{
	return bilerp_argb(__ldg(tex + t.i00), __ldg(tex + t.i10), __ldg(tex + t.i01), __ldg(tex + t.i11), t.fu, t.fv);
}

// the raw texels of one step (height + colour gathers, fog entry): loaded one chunk ahead of their use
struct RawSample { int h[4]; uint32_t c[4]; uint32_t fog, fu, fv; };

__device__ __forceinline__ RawSample fetch_sample(const uint8_t *__restrict__ heightMap, const uint32_t *__restrict__ colorMap, const uint32_t *__restrict__ fogGradient,
	unsigned iStep, int curX, int curY, unsigned mapAnd, unsigned mapShift)
{
	const TexCoords t = prep_uvs(curX, curY, mapAnd, mapShift);
	RawSample r;
	r.h[0] = __ldg(heightMap + t.i00); r.h[1] = __ldg(heightMap + t.i10); r.h[2] = __ldg(heightMap + t.i01); r.h[3] = __ldg(heightMap + t.i11);
	r.c[0] = __ldg(colorMap + t.i00); r.c[1] = __ldg(colorMap + t.i10); r.c[2] = __ldg(colorMap + t.i01); r.c[3] = __ldg(colorMap + t.i11);
	r.fog = __ldg(fogGradient + (iStep >> 1));
	r.fu = t.fu; r.fv = t.fv;
	return r;
}

// the same texels through tex2Dgather (ckd_footprint_texture; the landscape, whose 26 warps per SM are bound by L1 tag look-ups --
// the tunnelscape at 15 warps per SM is not, and measured 2 % slower this way): the footprint of (U0, V0) is addressed by the corner its four texels
// share, (U0+1, V0+1)/size -- exact in float for power-of-two maps, half a texel away from any rounding boundary -- and comes back
// as .w = (U0,V0), .z = (U0+1,V0), .x = (U0,V0+1), .y = (U0+1,V0+1), with the reference's '& mapAnd' as wrap addressing
__device__ __forceinline__ RawSample fetch_sample_tex(cudaTextureObject_t heightTex, cudaTextureObject_t colorTex, const uint32_t *__restrict__ fogGradient,
	unsigned iStep, int curX, int curY, unsigned mapAnd, float invMapSize)
{
	const float u = (float((unsigned(curX) >> 8) & mapAnd) + 1.f)*invMapSize, v = (float((unsigned(curY) >> 8) & mapAnd) + 1.f)*invMapSize;
	const uchar4 h = tex2Dgather<uchar4>(heightTex, u, v, 0);
	const uint4 c = tex2Dgather<uint4>(colorTex, u, v, 0);
	RawSample r;
	r.h[0] = h.w; r.h[1] = h.z; r.h[2] = h.x; r.h[3] = h.y;
	r.c[0] = c.w; r.c[1] = c.z; r.c[2] = c.x; r.c[3] = c.y;
	r.fog = __ldg(fogGradient + (iStep >> 1));
	r.fu = unsigned(curX) & 0xff; r.fv = unsigned(curY) & 0xff;
	return r;
}

// ---- cspanISSE16 (cspan.h:47-78) --------------------------------------------------------------------------------------
// 16.16 fixed-point ramp from colour A to colour B over 'length' pixels of which the last... 'drawLength' are drawn, with
// the reference's pmaddwd / pmuldq quirks restated literally: the divisor and the deltas are read as
// pairs of signed 16-bit words, and the pre-step only reaches the B and R lanes.

struct SpanRamp { uint32_t from[4]; int step[4]; };

__device__ __forceinline__ int madd16(int a, int b)
{
	// one 32-bit lane of pmaddwd
	return int(short(a & 0xffff))*int(short(b & 0xffff)) + int(short(a >> 16))*int(short(b >> 16));
}

__device__ __forceinline__ SpanRamp span_setup(unsigned length, unsigned drawLength, const Color16 &A, const Color16 &B)
{
	SpanRamp r;
	const int divisor = int(65536u/length);
	const unsigned preSteps = length - drawLength;
	const long long prod = (long long)divisor*(long long)int(preSteps); // _mm_mul_epi32: lanes 0 and 2, 64-bit results
	const int prodLo = int(prod & 0xffffffffll), prodHi = int(prod >> 32);
	#pragma unroll
	for (int i = 0; i < 4; ++i)
	{
		const int delta = B.c[i] - A.c[i];
		const int preStep = madd16(delta, (i & 1) ? prodHi : prodLo);
		r.step[i] = madd16(delta, divisor);
		r.from[i] = (uint32_t(A.c[i]) << 16) + uint32_t(preStep);
	}
	return r;
}

__device__ __forceinline__ uint32_t span_pixel(const SpanRamp &r, unsigned j)
{
	uint32_t px = 0;
	#pragma unroll
	for (int i = 0; i < 4; ++i)
	{
		const uint32_t v = (r.from[i] + j*uint32_t(r.step[i])) >> 16; // psrld 16, then packusdw (no-op on 16 bits), packuswb
		const uint32_t ch = (v > 32767u) ? 0u : min(v, 255u);
		px |= ch << (8*i);
	}
	return px;
}

// spans up to this length are drawn by their own lane, longer ones by the whole warp (tuned on B200, profiles/r02_notes.md;
// CKD_SHORT_SPAN overrides it for sweeps)
__constant__ unsigned c_shortSpan = 48;
#define kShortSpan c_shortSpan
// landscape: instruction estimates of the two span emitters (per pixel of the longest span / per batch of 32 pixels), see emit_tiled_spans
__constant__ unsigned c_tiledPerLane = 12, c_tiledPerBatch = 45;

// ramp of a visible lane's span and whether it is 'plain': a ramp whose two end points -- evaluated without the 32-bit
// wrap-around -- lie within [0, 2^24) cannot wrap or leave that range in between (it is linear): every channel value is then
// byte 2 of its accumulator, the packusdw / packuswb clamps of span_pixel cannot act, and a pixel costs four additions (or
// multiply-adds) and three byte permutes.  Anything else (the ball's over-bright colours, the pmaddwd quirks of steep ramps)
// takes the literal path.
__device__ __forceinline__ bool span_prepare(SpanRamp &ramp, bool visible, unsigned length, unsigned drawLength, const Color16 &A, const Color16 &B)
{
	bool plain = visible;
	if (visible)
	{
		ramp = span_setup(length, drawLength, A, B);
		#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			const long long last = (long long)(ramp.from[i]) + (long long)(drawLength - 1)*(long long)(ramp.step[i]);
			plain = plain && ramp.from[i] < (1u << 24) && last >= 0 && last < (1ll << 24);
		}
	}
	return plain;
}

__device__ __forceinline__ void emit_prepared_spans(uint32_t *line, int limit, bool visible, int pos, int dir, unsigned drawLength, const SpanRamp &ramp, bool plain)
{
	const int lane = threadIdx.x & 31;

	if (visible && drawLength <= kShortSpan)
	{
		if (plain)
		{
			uint32_t c0 = ramp.from[0], c1 = ramp.from[1], c2 = ramp.from[2], c3 = ramp.from[3];
			int idx = pos;
			for (unsigned j = 0; j < drawLength; ++j, idx += dir)
			{
				if (idx >= 0 && idx < limit)
					line[idx] = __byte_perm(__byte_perm(c0, c1, 0x0062), __byte_perm(c2, c3, 0x0062), 0x5410);
				c0 += uint32_t(ramp.step[0]); c1 += uint32_t(ramp.step[1]); c2 += uint32_t(ramp.step[2]); c3 += uint32_t(ramp.step[3]);
			}
		}
		else
		{
			for (unsigned j = 0; j < drawLength; ++j)
			{
				const int idx = pos + int(j)*dir;
				if (idx >= 0 && idx < limit)
					line[idx] = span_pixel(ramp, j);
			}
		}
	}

	unsigned longMask = __ballot_sync(kFull, visible && drawLength > kShortSpan);
	const unsigned plainMask = __ballot_sync(kFull, plain);
	while (longMask)
	{
		const int src = __ffs(longMask) - 1;
		longMask &= longMask - 1;
		SpanRamp r;
		#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			r.from[i] = __shfl_sync(kFull, ramp.from[i], src);
			r.step[i] = __shfl_sync(kFull, ramp.step[i], src);
		}
		const int p = __shfl_sync(kFull, pos, src);
		const unsigned len = __shfl_sync(kFull, drawLength, src);
		if ((plainMask >> src) & 1u)
		{
			for (unsigned j = lane; j < len; j += 32)
			{
				const int idx = p + int(j)*dir;
				if (idx >= 0 && idx < limit)
					line[idx] = __byte_perm(__byte_perm(r.from[0] + j*uint32_t(r.step[0]), r.from[1] + j*uint32_t(r.step[1]), 0x0062),
						__byte_perm(r.from[2] + j*uint32_t(r.step[2]), r.from[3] + j*uint32_t(r.step[3]), 0x0062), 0x5410);
			}
		}
		else
		{
			for (unsigned j = lane; j < len; j += 32)
			{
				const int idx = p + int(j)*dir;
				if (idx >= 0 && idx < limit)
					line[idx] = span_pixel(r, j);
			}
		}
	}
}

// Emits the spans of one 32-step chunk into a line buffer.  Lane parameters: visible, pos (index of the span's first
// pixel in the line), dir (+1/-1 index increment), length/drawLength and the two colours; 'limit' clips writes to the line.
__device__ __forceinline__ void emit_spans(uint32_t *line, int limit, bool visible, int pos, int dir, unsigned length, unsigned drawLength, const Color16 &A, const Color16 &B)
{
	SpanRamp ramp;
	const bool plain = span_prepare(ramp, visible, length, drawLength, A, B);
	emit_prepared_spans(line, limit, visible, pos, dir, drawLength, ramp, plain);
}

// The landscape's spans of one chunk tile an interval: a visible step draws [height, lowest height before it), so together the
// visible steps cover [lowest after the chunk, lowest before it) without gaps or overlap.  emit_spans pays for the LONGEST span
// of the chunk (every lane walks its own); here the lanes share the chunk's pixels instead -- lane k takes pixel lo + k, finds
// the step that owns it with a binary search over the (monotone) running minimum and evaluates that step's ramp at its offset:
// the cost follows the number of pixels on screen.  The caller picks whichever is cheaper for the chunk.
//   m       = running minimum including this lane's step (= its height where the step is visible)
//   [lo,hi) = the chunk's interval clipped to the line
__device__ __forceinline__ void emit_tiled_spans(uint32_t *line, int lo, int hi, int m, const SpanRamp &ramp, unsigned plainMask)
{
	const int lane = threadIdx.x & 31;
	for (int y0 = lo; y0 < hi; y0 += 32)
	{
		const int y = min(y0 + lane, hi - 1);
		int owner = 0; // number of leading steps whose running minimum is still above y = the first step that reaches it
		#pragma unroll
		for (int s = 16; s; s >>= 1)
		{
			const int v = __shfl_sync(kFull, m, owner + s - 1);
			if (v > y) owner += s;
		}
		const unsigned j = unsigned(y - __shfl_sync(kFull, m, owner));
		SpanRamp r;
		#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			r.from[i] = __shfl_sync(kFull, ramp.from[i], owner);
			r.step[i] = __shfl_sync(kFull, ramp.step[i], owner);
		}
		uint32_t px;
		if ((plainMask >> owner) & 1u)
			px = __byte_perm(__byte_perm(r.from[0] + j*uint32_t(r.step[0]), r.from[1] + j*uint32_t(r.step[1]), 0x0062),
				__byte_perm(r.from[2] + j*uint32_t(r.step[2]), r.from[3] + j*uint32_t(r.step[3]), 0x0062), 0x5410);
		else
			px = span_pixel(r, j);
		if (y0 + lane < hi)
			line[y] = px;
	}
}

// warp exclusive prefix min / max with carry-in (lane i gets op(carry, v_0..v_{i-1})); returns the inclusive total via 'total'
__device__ __forceinline__ int warp_excl_min(int v, int carry, int &total)
{
	const int lane = threadIdx.x & 31;
	int incl = v;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const int o = __shfl_up_sync(kFull, incl, d);
		if (lane >= d) incl = min(incl, o);
	}
	int excl = __shfl_up_sync(kFull, incl, 1);
	excl = (lane == 0) ? carry : min(carry, excl);
	total = min(carry, __shfl_sync(kFull, incl, 31));
	return excl;
}

__device__ __forceinline__ unsigned warp_excl_max(unsigned v, unsigned carry, unsigned &total)
{
	const int lane = threadIdx.x & 31;
	unsigned incl = v;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const unsigned o = __shfl_up_sync(kFull, incl, d);
		if (lane >= d) incl = max(incl, o);
	}
	unsigned excl = __shfl_up_sync(kFull, incl, 1);
	excl = (lane == 0) ? carry : max(carry, excl);
	total = max(carry, __shfl_sync(kFull, incl, 31));
	return excl;
}

// per-channel psubusw / paddusw on 16-bit lanes
__device__ __forceinline__ int subs16(int a, int b) { return max(a - b, 0); }
__device__ __forceinline__ int adds16(int a, int b) { return min(a + b, 65535); }

// -------------------------------------------------------------------------------------------------------------
// Landscape -- landscape.cpp:56-192
// -------------------------------------------------------------------------------------------------------------

constexpr int kScapeColsPerBlock = 8;     // default columns (= warps) per CTA; the launch picks what fills the machine in whole waves
constexpr int kScapeMaxColsPerBlock = 13;   // with two CTAs per SM in the launch bounds: 72 registers
constexpr unsigned kScapeRayLength = 512; // landscape.cpp:43


struct LandscapeFrame
{
	int resX, resY;
	int fpX1, fpY1;           // ray origin, 24:8
	float X1, Y1;
	float viewCos, viewSin;
	float rayY;               // kMapSize*kMapViewLenScale
	int mapTilt;              // s_mapTilt
	uint32_t clearColor;      // s_pFogGradient[0]
};

__global__ void __launch_bounds__(kScapeMaxColsPerBlock*32, 2) landscape_kernel(uint32_t *__restrict__ pDest, cudaTextureObject_t heightTex, cudaTextureObject_t colorTex, const uint32_t *__restrict__ colorMap,
	const uint32_t *__restrict__ fogGradient, const LandscapeFrame f, int lineStride)
{
	extern __shared__ uint32_t s_lines[]; // [columns per CTA][lineStride]
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int colsPerBlock = int(blockDim.x >> 5);
	const unsigned iRay = blockIdx.x*colsPerBlock + warp;
	uint32_t *line = s_lines + warp*lineStride;

	for (int y = lane; y < f.resY; y += 32)
		line[y] = f.clearColor;
	__syncwarp();

	if (iRay < unsigned(f.resX))
	{
		// per-column set-up, landscape.cpp:172-190 and voxel-shared.h:8-24 (IEEE float, identical on both sides)
		const float rayX = 0.25f*(float(iRay) - float(f.resX)*0.5f);
		float rotRayX = rayX, rotRayY = f.rayY;
		{
			const float rX = f.viewCos*rotRayX - f.viewSin*rotRayY;
			const float rY = f.viewSin*rotRayX + f.viewCos*rotRayY;
			rotRayX = rX; rotRayY = rY;
		}
		const float X2 = f.X1+rotRayX, Y2 = f.Y1+rotRayY;
		float dX = X2-f.X1, dY = Y2-f.Y1;
		if (fabsf(dX+dY) > kEpsilon)
		{
			const float length = 1.f/sqrtf(dX*dX + dY*dY);
			dX *= length;
			dY *= length;
		}
		const float fishMul = f.rayY / sqrtf(rotRayX*rotRayX + rotRayY*rotRayY);

		const int fpDX = ftofp24(dX), fpDY = ftofp24(dY);
		const int fpFishMul = ftofp24(fabsf(fishMul));

		// vscape_ray, landscape.cpp:56-107
		int carryLastHeight = f.resY;
		int carryLastDrawn = f.resY;
		Color16 carryLastColor;
		{
			const unsigned U = unsigned(f.fpX1 >> 8) & 1023u, V = (unsigned(f.fpY1 >> 8) & 1023u) << 10;
			carryLastColor = unpack16(__ldg(colorMap + (U|V)));
		}

		// the samples of a chunk do not depend on the carry: the gathers of the NEXT chunk are issued before this chunk's scan
		// and span stores (see tunnelscape_kernel), so their L2 latency overlaps the emission instead of heading every iteration
		RawSample next = fetch_sample_tex(heightTex, colorTex, fogGradient, lane, int(unsigned(f.fpX1) + (lane+1)*unsigned(fpDX)), int(unsigned(f.fpY1) + (lane+1)*unsigned(fpDY)), 1023u, 1.f/1024.f);

		for (unsigned base = 0; base < kScapeRayLength; base += 32)
		{
			const unsigned iStep = base + lane;
			const RawSample raw = next;
			if (base + 32 < kScapeRayLength)
			{
				const unsigned nStep = iStep + 32;
				next = fetch_sample_tex(heightTex, colorTex, fogGradient, nStep, int(unsigned(f.fpX1) + (nStep+1)*unsigned(fpDX)), int(unsigned(f.fpY1) + (nStep+1)*unsigned(fpDY)), 1023u, 1.f/1024.f);
			}

			const unsigned mapHeight = bilerp_u8(raw.h[0], raw.h[1], raw.h[2], raw.h[3], int(raw.fu), int(raw.fv));
			Color16 color = unpack16(bilerp_argb(raw.c[0], raw.c[1], raw.c[2], raw.c[3], raw.fu, raw.fv));

			const Color16 fog = unpack16(raw.fog);
			#pragma unroll
			for (int i = 0; i < 4; ++i) color.c[i] = subs16(color.c[i], fog.c[i]);

			int height = 255-int(mapHeight);
			height <<= 16;
			height = int(unsigned(height) / (unsigned(fpFishMul)*(iStep+1)));
			height *= 512; // kMapScale
			height >>= 8;
			height += f.mapTilt;

			// what the serial loop would have seen: previous step's height/colour and the lowest height drawn before it
			int prevHeight = __shfl_up_sync(kFull, height, 1);
			Color16 prevColor = shfl_up_color(color);
			if (lane == 0) { prevHeight = carryLastHeight; prevColor = carryLastColor; }

			int newLastDrawn;
			const int drawnBefore = warp_excl_min(height, carryLastDrawn, newLastDrawn);

			const bool visible = height < drawnBefore;
			// spans: per lane (cost ~ the longest span of the chunk) or tiled over the lanes (cost ~ the chunk's pixels on screen)
			const int lo = max(newLastDrawn, 0), hi = min(carryLastDrawn, f.resY);
			if (hi > lo)
			{
				const unsigned drawLength = unsigned(drawnBefore - height);
				SpanRamp ramp;
				const bool plain = span_prepare(ramp, visible, unsigned(prevHeight - height), drawLength, color, prevColor);
				const unsigned longest = __reduce_max_sync(kFull, visible ? drawLength : 0u);
				if (longest*c_tiledPerLane > unsigned((hi - lo + 31) >> 5)*c_tiledPerBatch)
					emit_tiled_spans(line, lo, hi, min(drawnBefore, height), ramp, __ballot_sync(kFull, plain));
				else
					emit_prepared_spans(line, f.resY, visible, height, 1, drawLength, ramp, plain);
			}

			carryLastDrawn = newLastDrawn;
			carryLastHeight = __shfl_sync(kFull, height, 31);
			carryLastColor = shfl_color(color, 31);
			if (carryLastDrawn <= 0)
				break; // the column is drawn up to the top of the screen: whatever else is visible lies above it
		}
	}

	__syncthreads();

	// write the CTA's columns out: consecutive threads take consecutive columns of a row (one 32-byte sector per row at 8 columns)
	const int x0 = blockIdx.x*colsPerBlock;
	const int c = int(threadIdx.x) % colsPerBlock;
	if (x0 + c < f.resX)
	{
		for (int y = int(threadIdx.x) / colsPerBlock; y < f.resY; y += 32)
			pDest[size_t(y)*f.resX + x0 + c] = s_lines[c*lineStride + y];
	}
}

// -------------------------------------------------------------------------------------------------------------
// Tunnelscape -- tunnelscape.cpp:44-134
// -------------------------------------------------------------------------------------------------------------

constexpr int kRowsPerBlock = 4;
// the ball's rays one per CTA: measured at 4K 1 / 2 / 4 / 8 rays per CTA = 36.5 / 36.8 / 41.5 / 36.8 us (beams: 45.6 / 45.6 / 49.4 /
// 45.0); the tunnelscape is the other way round (50.2 / 50.2 / 47.8 / 50.1), the twister does not care (37.2 / 37.9 / 37.6 / 37.3)
constexpr int kBallRowsPerBlock = 1;

struct TunnelscapeFrame
{
	int resX, resY;
	int dX, dY, fpFromY;
	float mapStepX;           // 2048.f/(kTargetResY-1)
	float fromXOffs;          // syncDirX * time*kGoldenRatio
	float viewLenScale;       // kMapViewLenScale = kAspect*0.5f (tunnelscape.cpp:29)
	uint32_t clearColor;
};

__global__ void __launch_bounds__(kRowsPerBlock*32) tunnelscape_kernel(uint32_t *__restrict__ pDest, const uint8_t *__restrict__ heightMap, const uint32_t *__restrict__ colorMap,
	const uint32_t *__restrict__ fogGradient, const TunnelscapeFrame f)
{
	// The spans of a ray pile up left to right without gaps, so they go straight to the row in global memory (neighbouring
	// lanes write neighbouring pixels, L2 merges the sectors); what the ray leaves uncovered gets the clear colour at the end.
	// No shared-memory line buffer: the resident warps per SM are bounded by registers only, which is what hides the L2
	// gather latency of this kernel.
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const unsigned iRay = blockIdx.x*kRowsPerBlock + warp;
	if (iRay >= unsigned(f.resY))
		return;
	uint32_t *line = pDest + size_t(iRay)*f.resX;

	// tscape, tunnelscape.cpp:121-131
	const float mapX = float(iRay)*f.mapStepX;
	const float fromX = mapX + f.fromXOffs;
	const int fpFromX = ftofp24(fromX);

	// tscape_ray, tunnelscape.cpp:44-99
	int carryLastHeight = f.resX;
	int carryLastDrawn = f.resX;
	Color16 carryLastColor;
	{
		const unsigned U = unsigned(fpFromX >> 8) & 2047u, V = (unsigned(f.fpFromY >> 8) & 2047u) << 11;
		carryLastColor = unpack16(__ldg(colorMap + (U|V)));
	}

	// The samples of a chunk do not depend on the carry, and at 4K there are only 2160 rays (15 warps per SM): the gathers of
	// the NEXT chunk are issued before this chunk's scan and span stores, so their L2 latency overlaps the emission
	// instead of heading every iteration (registers are free at this occupancy).
	RawSample next = fetch_sample(heightMap, colorMap, fogGradient, lane, int(unsigned(fpFromX) + (lane+1)*unsigned(f.dX)), int(unsigned(f.fpFromY) + (lane+1)*unsigned(f.dY)), 2047u, 11u);

	for (unsigned base = 0; base < 512; base += 32)
	{
		const unsigned iStep = base + lane;
		const RawSample raw = next;
		if (base + 32 < 512)
		{
			const unsigned nStep = iStep + 32;
			next = fetch_sample(heightMap, colorMap, fogGradient, nStep, int(unsigned(fpFromX) + (nStep+1)*unsigned(f.dX)), int(unsigned(f.fpFromY) + (nStep+1)*unsigned(f.dY)), 2047u, 11u);
		}

		const unsigned mapHeight = bilerp_u8(raw.h[0], raw.h[1], raw.h[2], raw.h[3], int(raw.fu), int(raw.fv));
		Color16 color = unpack16(bilerp_argb(raw.c[0], raw.c[1], raw.c[2], raw.c[3], raw.fu, raw.fv));

		const Color16 fog = unpack16(raw.fog);
		#pragma unroll
		for (int i = 0; i < 4; ++i) color.c[i] = subs16(color.c[i], fog.c[i]);

		// tunnelscape.cpp:73-81 (the divide by 'iStep+1' is unsigned, the multiply wraps)
		int height = 255-int(mapHeight);
		height -= 96;   // kMapViewHeight
		height <<= 8;
		height = cvtt_x86(float(height)/f.viewLenScale);
		height = int(unsigned(height) / (iStep+1));
		height = int(unsigned(height)*160u); // kMapScale
		height >>= 8;
		height += 120;  // kMapTilt

		int prevHeight = __shfl_up_sync(kFull, height, 1);
		Color16 prevColor = shfl_up_color(color);
		if (lane == 0) { prevHeight = carryLastHeight; prevColor = carryLastColor; }

		int newLastDrawn;
		const int drawnBefore = warp_excl_min(height, carryLastDrawn, newLastDrawn);

		const bool visible = height < drawnBefore;
		// spans pile up left to right: the span of this step starts where everything drawn before it ended
		emit_spans(line, f.resX, visible, f.resX - drawnBefore, 1, unsigned(prevHeight - height), unsigned(drawnBefore - height), color, prevColor);

		carryLastDrawn = newLastDrawn;
		carryLastHeight = __shfl_sync(kFull, height, 31);
		carryLastColor = shfl_color(color, 31);
	}

	// memset32(g_renderTarget[0], s_pFogGradient[0], ...), tunnelscape.cpp:170: only the part no span covered
	for (int x = max(f.resX - carryLastDrawn, 0) + lane; x < f.resX; x += 32)
		line[x] = f.clearColor;
}

// -------------------------------------------------------------------------------------------------------------
// Ball -- ball.cpp:80-365
// -------------------------------------------------------------------------------------------------------------

struct BallFrame
{
	int resX, resY;
	int fromX, fromY;
	unsigned rayLength;       // s_curRayLength
	unsigned beamAtten;       // s_beamAtten
	float beamAlphaMin;       // s_beamAlphaMin
	unsigned lowLight;
	int clearLastPixel;       // ckd_set_frame_independent: the pixel the beam path never writes starts from 0 instead of history
};

// The tail of vball_ray_beams (ball.cpp:168-203): from the last drawn pixel to one short of the row's end the beam colour goes
// out with an alpha that follows smoothstep(beamAlphaMin, luminosity, curStep), curStep accumulated in float pixel by pixel.
// RAW (ckd_ball_beam_tail's check mode) writes the bits of curStep itself instead of the pixel.
template <bool RAW>
__device__ __forceinline__ void beam_tail(uint32_t *line, int resX, unsigned carryLastDrawn, uint32_t beamCol, float beamAlphaMin)
{
	const int lane = threadIdx.x & 31;
	const unsigned lastDrawnHeight = carryLastDrawn;
	const unsigned remainder = unsigned(resX - 1) - lastDrawnHeight;
	beamCol &= 0xffffffu;

	const unsigned beamR = beamCol >> 16, beamG = (beamCol >> 8) & 0xff, beamB = beamCol & 0xff;
	const unsigned mulR = 4731u, mulG = 46871u, mulB = 13932u; // unsigned(0.0722f*65536.f) etc.
	const unsigned luminosity = ((beamR*mulR) >> 16) + ((beamG*mulG) >> 16) + ((beamB*mulB) >> 16);
	const float fLuminosity = float(luminosity);

	if (remainder <= unsigned(resX))
	{
		// 'curStep += alphaStep' is a serial float accumulation (ball.cpp:195-203): pixel i needs the value after i rounded
		// additions.  The first 32 pixels take them literally.  After that the chain is walked binade by binade: while the
		// running value c stays within one binade its grid (ulp) is fixed, so fl(c + a) = c + (a rounded to that grid) -- a
		// constant increment, once one addition has been made inside the binade (a tie of a's dropped bits rounds to even,
		// which can make the first step of a binade differ; after it the values are even multiples and the tie always falls
		// the same way).  So per binade: c (as it entered), c1 = fl(c + a), and from there c1 + (m-1)*inc with
		// inc = fl(c1 + a) - c1; products and sums of grid multiples below 2^24 grid units are exact in float.  The step that
		// leaves the binade is one real addition.  A dozen binades per ray (c runs from 32/(n-1) to 1) instead of one dependent
		// FADD per pixel on every lane.
		const float alphaStep = 1.f / float(remainder - 1);
		const bool narrow = fabsf(beamAlphaMin) < 1.0e9f; // the alpha stays far inside int32: cvttss2si's 64-bit form and the 32-bit one agree
		auto put = [&](unsigned i, float curStep)
		{
			if (RAW)
			{
				line[lastDrawnHeight + i] = __float_as_uint(curStep);
				return;
			}
			const float fBeamAlpha = smoothstepf(beamAlphaMin, fLuminosity, curStep);
			const unsigned beamAlpha = narrow ? unsigned(__float2int_rz(fBeamAlpha)) : f2u_x86(fBeamAlpha);
			line[lastDrawnHeight + i] = beamCol | (beamAlpha << 24);
		};

		// pixels 0..31: the same 32 dependent additions on every lane; a lane keeps the value after `lane` of them, picked with
		// a three-level select per group of 8 (predicates from the lane's low bits)
		float cur = 0.f, mine = 0.f;
		#pragma unroll
		for (int grp = 0; grp < 4; ++grp)
		{
			float v[8];
			v[0] = cur;
			#pragma unroll
			for (int j = 1; j < 8; ++j) v[j] = v[j-1] + alphaStep;
			cur = v[7] + alphaStep;
			const float s01 = (lane & 1) ? v[1] : v[0], s23 = (lane & 1) ? v[3] : v[2], s45 = (lane & 1) ? v[5] : v[4], s67 = (lane & 1) ? v[7] : v[6];
			const float s03 = (lane & 2) ? s23 : s01, s47 = (lane & 2) ? s67 : s45;
			const float sel = (lane & 4) ? s47 : s03;
			if ((lane >> 3) == grp) mine = sel;
		}
		if (unsigned(lane) < remainder)
			put(lane, mine);

		unsigned k = 32;  // index of the next pixel
		float c = cur;    // its value (identical on every lane)
		while (k < remainder)
		{
			const float c1 = c + alphaStep, c2 = c1 + alphaStep;
			const unsigned e = __float_as_uint(c) >> 23;
			float base = c, inc = 0.f;
			unsigned n = 1; // values of the chain within this binade
			if ((__float_as_uint(c1) >> 23) == e)
			{
				inc = c2 - c1;
				base = c1 - inc;
				const float perUlp = __uint_as_float((277u - e) << 23); // 2^(150 - e): grid units per 1.0
				const unsigned baseUnits = unsigned(__float2int_rz(base*perUlp)), incUnits = unsigned(__float2int_rz(inc*perUlp));
				n = 1u + (0xffffffu - baseUnits)/incUnits;
			}
			n = min(n, remainder - k);
			for (unsigned m = lane; m < n; m += 32)
				put(k + m, (0 == m) ? c : base + float(m)*inc);
			const unsigned last = n - 1;
			c = ((0 == last) ? c : base + float(last)*inc) + alphaStep;
			k += n;
		}
	}
}

// tables: heightProj[1024], projNorm0[1024], projNorm1[1024], projNorm2[1024] (ball.cpp:61-62)
template <bool BEAMS>
__global__ void __launch_bounds__(kBallRowsPerBlock*32) ball_kernel(uint32_t *pDest, const uint8_t *__restrict__ heightMap, const uint32_t *__restrict__ colorMap,
	const uint32_t *__restrict__ auxMap /* beam mix or env map */, const int *__restrict__ tables, const int *__restrict__ rayDeltas, const BallFrame f)
{
	// spans go straight to the row in global memory (see tunnelscape_kernel); with beams the reference does not clear the
	// render target (ball.cpp:352-363), so whatever no span and no beam pixel covers simply keeps its previous content
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const unsigned iRay = blockIdx.x*kBallRowsPerBlock + warp;
	if (iRay >= unsigned(f.resY))
		return;
	uint32_t *line = pDest + size_t(iRay)*f.resX;
	if (BEAMS && f.clearLastPixel)
	{
		// the extrusion stops one pixel short of the row (ball.cpp:186) and nothing clears the target (ball.cpp:352-363): that
		// pixel is whatever an earlier frame left there unless a span reaches it
		if (0 == lane) line[f.resX-1] = 0;
		__syncwarp();
	}

	const int dX = rayDeltas[iRay*2], dY = rayDeltas[iRay*2+1];
	const int *heightProj = tables, *projNorm0 = tables + 1024, *projNorm1 = tables + 2048, *projNorm2 = tables + 3072;

	unsigned carryLastHeight = 0, carryLastDrawn = 0;
	Color16 carryLastColor = unpack16(sample_argb(colorMap, prep_uvs(f.fromX, f.fromY, 1023u, 10u)));
	int beamCarry[4] = { 0, 0, 0, 0 };

	// the height / colour (/ beam) texels of a chunk do not depend on the carry: those of the NEXT chunk are requested before this
	// chunk's scan and span stores (see tunnelscape_kernel).  Steps past the ray's end sample wrapped coordinates that nothing uses.
	struct BallRaw { int h[4]; uint32_t c[4]; uint32_t b[4]; uint32_t fu, fv; };
	auto fetchRaw = [&](unsigned iStep) -> BallRaw
	{
		const TexCoords t = prep_uvs(int(unsigned(f.fromX) - (iStep+1)*unsigned(dX)), int(unsigned(f.fromY) - (iStep+1)*unsigned(dY)), 1023u, 10u);
		BallRaw r;
		r.h[0] = __ldg(heightMap + t.i00); r.h[1] = __ldg(heightMap + t.i10); r.h[2] = __ldg(heightMap + t.i01); r.h[3] = __ldg(heightMap + t.i11);
		r.c[0] = __ldg(colorMap + t.i00); r.c[1] = __ldg(colorMap + t.i10); r.c[2] = __ldg(colorMap + t.i01); r.c[3] = __ldg(colorMap + t.i11);
		if (BEAMS) { r.b[0] = __ldg(auxMap + t.i00); r.b[1] = __ldg(auxMap + t.i10); r.b[2] = __ldg(auxMap + t.i01); r.b[3] = __ldg(auxMap + t.i11); }
		r.fu = t.fu; r.fv = t.fv;
		return r;
	};
	BallRaw next = fetchRaw(lane);

	for (unsigned base = 0; base < f.rayLength; base += 32)
	{
		const unsigned iStep = base + lane;
		const bool active = iStep < f.rayLength;
		const unsigned tabIdx = active ? iStep : 0;

		const BallRaw raw = next;
		if (base + 32 < f.rayLength)
			next = fetchRaw(iStep + 32);
		const unsigned mapHeight = bilerp_u8(raw.h[0], raw.h[1], raw.h[2], raw.h[3], int(raw.fu), int(raw.fv));
		Color16 color = unpack16(bilerp_argb(raw.c[0], raw.c[1], raw.c[2], raw.c[3], raw.fu, raw.fv));

		if (BEAMS)
		{
			// vball_ray_beams, ball.cpp:113-140
			Color16 beam = unpack16(bilerp_argb(raw.b[0], raw.b[1], raw.b[2], raw.b[3], raw.fu, raw.fv));
			const unsigned heightNorm = (mapHeight*unsigned(__ldg(projNorm0 + tabIdx))) >> 8;
			const unsigned heightNorm2 = (mapHeight*unsigned(__ldg(projNorm1 + tabIdx))) >> 8;
			const unsigned diffuse = heightNorm + ((unsigned(int(heightNorm2-heightNorm))*f.lowLight) >> 8);
			const int litWhite = int(diffuse & 0xffffu); // _mm_set1_epi16

			// paddusw prefix: all terms are >= 0, so the saturating running sum is min(65535, exact prefix sum).  A term is at most
			// 255, a warp's prefix at most 8160: two channels share one 32-bit word through the scan (16 bits each, no carry between)
			int lit[4];
			#pragma unroll
			for (int i = 0; i < 4; ++i)
			{
				const int b = ((beam.c[i]*int(f.beamAtten)) & 0xffff) >> 8;
				lit[i] = active ? (((b*litWhite) & 0xffff) >> 8) : 0;
			}
			unsigned incl01 = unsigned(lit[0]) | (unsigned(lit[1]) << 16), incl23 = unsigned(lit[2]) | (unsigned(lit[3]) << 16);
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1)
			{
				const unsigned o01 = __shfl_up_sync(kFull, incl01, d), o23 = __shfl_up_sync(kFull, incl23, d);
				if (lane >= d) { incl01 += o01; incl23 += o23; }
			}
			const int incl[4] = { int(incl01 & 0xffffu), int(incl01 >> 16), int(incl23 & 0xffffu), int(incl23 >> 16) };
			int beamIncl[4];
			#pragma unroll
			for (int i = 0; i < 4; ++i)
			{
				beamIncl[i] = min(beamCarry[i] + incl[i], 65535);
				beamCarry[i] = __shfl_sync(kFull, beamIncl[i], 31);
				color.c[i] = adds16(adds16(color.c[i], beamIncl[i]), litWhite);
			}
		}
		else
		{
			// vball_ray_no_beams, ball.cpp:228-262
			const int envU = int(unsigned(512 << 8) - (iStep+1)*unsigned(dX << 1));
			const int envV = int(unsigned(512 << 8) - (iStep+1)*unsigned(dY << 1));
			const TexCoords te = prep_uvs(envU + int(mapHeight), envV + int(mapHeight), 1023u, 10u);
			const Color16 envCol = unpack16(sample_argb(auxMap, te));
			const unsigned diffuse = (mapHeight*unsigned(__ldg(projNorm2 + tabIdx))) >> 8;
			const int lit = int(diffuse & 0xffffu);
			const int litFull = int(min(255u, 32u+diffuse)); // kAmbient
			#pragma unroll
			for (int i = 0; i < 4; ++i)
				color.c[i] = adds16(adds16(color.c[i], ((envCol.c[i]*litFull) & 0xffff) >> 8), lit);
		}

		unsigned height = (mapHeight*unsigned(__ldg(heightProj + tabIdx))) >> 8;
		if (!active) height = 0; // never visible, leaves the carries alone (they are taken from the last active lane below)

		unsigned prevHeight = __shfl_up_sync(kFull, height, 1);
		Color16 prevColor = shfl_up_color(color);
		if (lane == 0) { prevHeight = carryLastHeight; prevColor = carryLastColor; }

		unsigned newLastDrawn;
		const unsigned drawnBefore = warp_excl_max(height, carryLastDrawn, newLastDrawn);

		const bool visible = active && height > drawnBefore;
		emit_spans(line, f.resX, visible, int(drawnBefore), 1, height - prevHeight, height - drawnBefore, prevColor, color);

		carryLastDrawn = newLastDrawn;
		const int lastLane = int(min(f.rayLength - base, 32u)) - 1;
		carryLastHeight = __shfl_sync(kFull, height, lastLane);
		carryLastColor = shfl_color(color, lastLane);
	}

	__syncwarp();

	if (BEAMS)
	{
		// beam extrusion, ball.cpp:168-203
		uint32_t beamCol = 0;
		#pragma unroll
		for (int i = 0; i < 4; ++i)
		{
			const uint32_t v = uint32_t(beamCarry[i]);
			beamCol |= ((v > 32767u) ? 0u : min(v, 255u)) << (8*i); // v2cISSE16: packuswb
		}
		beam_tail<false>(line, f.resX, carryLastDrawn, beamCol, f.beamAlphaMin);
	}
	else
	{
		// memset32(pDest, 0, kTargetSize), ball.cpp:342: only the part no span covered
		for (int x = min(int(carryLastDrawn), f.resX) + lane; x < f.resX; x += 32)
			line[x] = 0;
	}
}

// one warp per row: the beam tail alone, row r with `firstRemainder + r` pixels left
template <bool RAW>
__global__ void __launch_bounds__(kRowsPerBlock*32) beam_tail_kernel(uint32_t *pDest, int rowPixels, int rows, unsigned firstRemainder, uint32_t beamCol, float beamAlphaMin)
{
	const unsigned row = blockIdx.x*kRowsPerBlock + (threadIdx.x >> 5);
	if (row >= unsigned(rows))
		return;
	const unsigned remainder = firstRemainder + row;
	beam_tail<RAW>(pDest + size_t(row)*rowPixels, rowPixels, unsigned(rowPixels - 1) - remainder, beamCol, beamAlphaMin);
}

// -------------------------------------------------------------------------------------------------------------
// Twister -- torus-twister.cpp:40-117
// -------------------------------------------------------------------------------------------------------------

// one warp per ray, two rays (right/left of centre) per row; rayOrigins[iRay*2] = fromX, [iRay*2+1] = fromY (host sinf)
__global__ void __launch_bounds__(kRowsPerBlock*64) twister_kernel(uint32_t *__restrict__ pDest, const uint8_t *__restrict__ heightMap, const uint32_t *__restrict__ colorMap,
	const int *__restrict__ tables, const int *__restrict__ rayOrigins, int resX, int resY)
{
	// one warp per ray, spans straight to global memory (see tunnelscape_kernel); the uncovered rest of each half row is cleared
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int rowInBlock = warp >> 1, side = warp & 1;
	const unsigned iRay = blockIdx.x*kRowsPerBlock + rowInBlock;
	if (iRay >= unsigned(resY))
		return;
	uint32_t *line = pDest + size_t(iRay)*resX;
	const int half = resX >> 1;

	const int *heightProj = tables, *heightProjNorm = tables + 512;
	const int kHalfMap = 1024/2;
	const int fromX0 = rayOrigins[iRay*2], fromY = rayOrigins[iRay*2+1];

	// vtwister_ray(pDest+xOffs, fromX, fromY, 512) and vtwister_ray(pDest+xOffs-1, fromX-512, fromY, -512), torus-twister.cpp:113-115
	const int startX = (0 == side) ? fromX0 : fromX0 - kHalfMap;
	const int dX = (0 == side) ? kHalfMap : -kHalfMap;
	const int direction = (dX < 0) ? -1 : 1;
	const int startPos = half + ((0 == side) ? 0 : -1);

	unsigned carryLastHeight = 0, carryLastDrawn = 0;
	Color16 carryLastColor = unpack16(sample_argb(colorMap, prep_uvs(startX, fromY, 1023u, 10u)));

	for (unsigned base = 0; base < 512; base += 32)
	{
		const unsigned iStep = base + lane;
		const int curX = int(unsigned(startX) - (iStep+1)*unsigned(dX));
		const TexCoords t = prep_uvs(curX, fromY, 1023u, 10u);
		const unsigned mapHeight = sample_u8(heightMap, t);
		Color16 color = unpack16(sample_argb(colorMap, t));

		const unsigned heightNorm = (mapHeight*unsigned(__ldg(heightProjNorm + iStep))) >> 8;
		const int litWhite = int(heightNorm & 0xffffu);
		#pragma unroll
		for (int i = 0; i < 4; ++i) color.c[i] = adds16(color.c[i], litWhite);

		const unsigned height = (mapHeight*unsigned(__ldg(heightProj + iStep))) >> 8;

		unsigned prevHeight = __shfl_up_sync(kFull, height, 1);
		Color16 prevColor = shfl_up_color(color);
		if (lane == 0) { prevHeight = carryLastHeight; prevColor = carryLastColor; }

		unsigned newLastDrawn;
		const unsigned drawnBefore = warp_excl_max(height, carryLastDrawn, newLastDrawn);

		const bool visible = height > drawnBefore;
		emit_spans(line, resX, visible, startPos + int(drawnBefore)*direction, direction, height - prevHeight, height - drawnBefore, prevColor, color);

		carryLastDrawn = newLastDrawn;
		carryLastHeight = __shfl_sync(kFull, height, 31);
		carryLastColor = shfl_color(color, 31);
	}

	// memset32(g_renderTarget[0], 0, kTargetSize), torus-twister.cpp:169: the part of this half row no span covered
	const int covered = min(int(carryLastDrawn), half);
	if (0 == side)
		for (int x = half + covered + lane; x < resX; x += 32) line[x] = 0;
	else
		for (int x = lane; x < half - covered; x += 32) line[x] = 0;
}

// -------------------------------------------------------------------------------------------------------------
// host helpers
// -------------------------------------------------------------------------------------------------------------

int RequireImage(const ckd_ctx *ctx, ckd_image slot, int w, int h, int bpp, const char *what)
{
	const ckd_image_slot &s = ctx->images[slot];
	if (!s.d_pixels || s.bpp != bpp || (w > 0 && s.width != w) || (h > 0 && s.height != h))
	{
		ckd_set_error(std::string("missing or mis-sized input image: ") + what);
		return CKD_ERR_MISSING_INPUT;
	}
	return CKD_OK;
}

template <class K> int EnsureSmem(K kernel, size_t bytes)
{
	if (bytes > 48*1024)
		CKD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
	return CKD_OK;
}

} // namespace

// CKD_SHORT_SPAN=n: tuning override of the lane-parallel / cooperative span threshold (once per process)
static int ApplyShortSpanOverride()
{
	static bool done = false;
	if (done)
		return CKD_OK;
	done = true;
	if (const char *env = getenv("CKD_SHORT_SPAN"))
	{
		const unsigned v = unsigned(atoi(env));
		CKD_CUDA(cudaMemcpyToSymbol(c_shortSpan, &v, sizeof(v)));
	}
	if (const char *env = getenv("CKD_TILED_LANE"))
	{
		const unsigned v = unsigned(atoi(env));
		CKD_CUDA(cudaMemcpyToSymbol(c_tiledPerLane, &v, sizeof(v)));
	}
	if (const char *env = getenv("CKD_TILED_BATCH"))
	{
		const unsigned v = unsigned(atoi(env));
		CKD_CUDA(cudaMemcpyToSymbol(c_tiledPerBatch, &v, sizeof(v)));
	}
	return CKD_OK;
}

// Columns (= warps) per landscape CTA.  A column's line buffer is resY pixels of shared memory, so the CTAs an SM holds are few
// (three of 8 columns at 4K) and the grid easily ends in a nearly empty last wave: 3840 columns in CTAs of 8 are 480 CTAs on
// 444 slots, the last 36 run alone.  The pick minimises waves x resident warps (what an instruction-bound kernel pays), with a
// floor where a wave is latency bound anyway; at 4K on 148 SMs that is 13 columns: 296 CTAs, two per SM, one full wave.
static int PickScapeColumns(ckd_ctx *ctx, int lineStride, int *pCols)
{
	if (ctx->scapeCols)
	{
		*pCols = ctx->scapeCols;
		return CKD_OK;
	}
	int maxOptin = 0;
	CKD_CUDA(cudaDeviceGetAttribute(&maxOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
	const int maxCols = std::min(kScapeMaxColsPerBlock, int(size_t(maxOptin)/(size_t(lineStride)*4)));
	CKD_REQUIRE(maxCols >= 1, "the output is too tall for the landscape's shared-memory line buffer");
	CKD_TRY(EnsureSmem(landscape_kernel, size_t(maxCols)*lineStride*4));
	int best = 0;
	long bestCost = 0;
	if (const char *env = getenv("CKD_SCAPE_COLS")) // tuning override
		best = std::max(1, std::min(maxCols, atoi(env)));
	if (!best)
	{
		constexpr long kLatencyFloorWarps = 16;
		for (int i = 0; i <= maxCols; ++i)
		{
			const int cols = (i == 0) ? std::min(kScapeColsPerBlock, maxCols) : maxCols + 1 - i; // the default first, then wide to narrow: ties keep the earlier one
			int perSM = 0;
			CKD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, landscape_kernel, cols*32, size_t(cols)*lineStride*4));
			if (perSM < 1)
				continue;
			const long ctas = ckd_div_up(ctx->resX, cols);
			const long slots = long(perSM)*ctx->numSMs;
			const long waves = (ctas + slots - 1)/slots;
			const long resident = std::min<long>(perSM, (ctas + ctx->numSMs - 1)/ctx->numSMs)*cols;
			const long cost = waves*std::max(resident, kLatencyFloorWarps);
			if (!best || cost < bestCost) { best = cols; bestCost = cost; }
		}
	}
	CKD_REQUIRE(best >= 1, "no landscape launch shape fits this device");
	ctx->scapeCols = best;
	*pCols = best;
	return CKD_OK;
}

// Landscape_Draw, landscape.cpp:228-243
extern "C" int ckd_landscape_draw(ckd_ctx *ctx, const ckd_landscape_params *p, float time, uint32_t *d_dest)
{
	CKD_TRY(ApplyShortSpanOverride());
	(void)time;
	CKD_REQUIRE(ctx && p && d_dest, "null argument");
	CKD_TRY(RequireImage(ctx, CKD_IMG_SCAPE_HEIGHT, 1024, 1024, 1, "landscape height map (assets/scape/D17.png)"));
	CKD_TRY(RequireImage(ctx, CKD_IMG_SCAPE_COLOR, 1024, 1024, 4, "landscape colour map (assets/scape/C17W-edit.png)"));
	CKD_TRY(RequireImage(ctx, CKD_IMG_SCAPE_FOG, 256, 1, 4, "fog gradient (assets/scape/foggradient.jpg)"));

	const float aspect = float(ctx->resY)/float(ctx->resX);

	LandscapeFrame f;
	f.resX = ctx->resX;
	f.resY = ctx->resY;

	// vscape, landscape.cpp:109-169 with the gamepad contribution passed in as accumulated state
	const float tilt = ckdh::clampf(-90.f, 90.f, p->tilt + p->pad_tilt);
	f.mapTilt = 90 + ckdh::x86_cvtt(tilt);
	f.viewCos = cosf(p->view_angle);
	f.viewSin = sinf(p->view_angle);
	float moveX = -f.viewSin*p->forward;
	float moveY = f.viewCos*p->forward;
	moveX += p->pad_move_x;
	moveY += p->pad_move_y;
	f.X1 = moveX+p->strafe_x;
	f.Y1 = moveY+p->strafe_y;
	f.fpX1 = ckdh::ftofp24(f.X1);
	f.fpY1 = ckdh::ftofp24(f.Y1);
	const float kMapViewLenScale = aspect*(ckdh::kPI*0.1f);
	f.rayY = 1024*kMapViewLenScale;

	f.clearColor = ctx->images[CKD_IMG_SCAPE_FOG].firstPixel; // s_pFogGradient[0], landscape.cpp:238

	const bool warp = 0.f != p->warp_strength;
	uint32_t *pWrite = warp ? ctx->d_renderTarget[0] : d_dest;

	cudaTextureObject_t heightTex = 0, colorTex = 0;
	CKD_TRY(ckd_footprint_texture(ctx, CKD_IMG_SCAPE_HEIGHT, &heightTex));
	CKD_TRY(ckd_footprint_texture(ctx, CKD_IMG_SCAPE_COLOR, &colorTex));
	const int lineStride = ctx->resY | 1;
	int cols = 0;
	CKD_TRY(PickScapeColumns(ctx, lineStride, &cols));
	const size_t smem = size_t(cols)*lineStride*4;
	ckd_prof_begin(ctx, "voxel_landscape", 4.0*ctx->resX*ctx->resY);
	landscape_kernel<<<ckd_div_up(ctx->resX, cols), cols*32, smem, ctx->stream>>>(pWrite, heightTex, colorTex, static_cast<const uint32_t *>(ctx->images[CKD_IMG_SCAPE_COLOR].d_pixels),
		static_cast<const uint32_t *>(ctx->images[CKD_IMG_SCAPE_FOG].d_pixels), f, lineStride);
	CKD_CHECK_LAUNCH(ctx);

	if (warp) // note the reference passes (WarpSpeed, WarpStrength) into (strength, speed), landscape.cpp:242
		CKD_TRY(ckd_tape_warp(ctx, d_dest, pWrite, unsigned(ctx->resX), unsigned(ctx->resY), p->warp_speed, p->warp_strength));
	return CKD_OK;
}

// Tunnelscape_Draw, tunnelscape.cpp:168-186
extern "C" int ckd_tunnelscape_draw(ckd_ctx *ctx, const ckd_tunnelscape_params *p, float time, uint32_t *d_dest)
{
	CKD_TRY(ApplyShortSpanOverride());
	CKD_REQUIRE(ctx && p && d_dest, "null argument");
	CKD_TRY(RequireImage(ctx, CKD_IMG_TSCAPE_HEIGHT, 2048, 2048, 1, "tunnelscape height map (assets/scape/tscape-D7-edit.png)"));
	CKD_TRY(RequireImage(ctx, CKD_IMG_TSCAPE_COLOR, 2048, 2048, 4, "tunnelscape colour map (assets/scape/tscape-C7W-edit.png)"));
	CKD_TRY(RequireImage(ctx, CKD_IMG_TSCAPE_FOG, 256, 1, 4, "fog gradient (assets/scape/foggradient.jpg)"));

	const float aspect = float(ctx->resY)/float(ctx->resX);
	const float oneOverAspect = 1.f/aspect;

	// tscape, tunnelscape.cpp:103-119
	TunnelscapeFrame f;
	f.resX = ctx->resX;
	f.resY = ctx->resY;
	f.mapStepX = 2048.f/float(ctx->resY-1);
	const float syncDirX = p->step_u, syncDirY = p->step_v;
	const float speedMul = sqrtf(syncDirX*syncDirX + syncDirY*syncDirY) * p->speed;
	const float fromY = 1024.f + speedMul*time;
	f.dX = ckdh::ftofp24(syncDirY);
	f.dY = ckdh::ftofp24(oneOverAspect*syncDirX);
	f.fpFromY = ckdh::ftofp24(fromY);
	f.fromXOffs = syncDirX * time*ckdh::kGoldenRatio;
	f.viewLenScale = aspect*0.5f;

	f.clearColor = ctx->images[CKD_IMG_TSCAPE_FOG].firstPixel; // s_pFogGradient[0], tunnelscape.cpp:170

	ckd_prof_begin(ctx, "voxel_tunnelscape", 4.0*ctx->resX*ctx->resY);
	tunnelscape_kernel<<<ckd_div_up(ctx->resY, kRowsPerBlock), kRowsPerBlock*32, 0, ctx->stream>>>(ctx->d_renderTarget[0],
		static_cast<const uint8_t *>(ctx->images[CKD_IMG_TSCAPE_HEIGHT].d_pixels), static_cast<const uint32_t *>(ctx->images[CKD_IMG_TSCAPE_COLOR].d_pixels),
		static_cast<const uint32_t *>(ctx->images[CKD_IMG_TSCAPE_FOG].d_pixels), f);
	CKD_CHECK_LAUNCH(ctx);

	if (0.f == p->blur)
		return ckd_polar_tail(ctx, d_dest, ctx->d_renderTarget[0], 1, false, nullptr);
	ctx->rbHost = nullptr; // the blur works on the whole frame: no banded read-back
	CKD_TRY(ckd_polar_blit(ctx, d_dest, ctx->d_renderTarget[0], 1));
	{
		const float scaledBlur = ckdh::BoxBlurScale(p->blur);
		CKD_TRY(ckd_old_blur(ctx, d_dest, d_dest, unsigned(ctx->resX), unsigned(ctx->resY), scaledBlur));
		CKD_TRY(ckd_old_blur(ctx, d_dest, d_dest, unsigned(ctx->resX), unsigned(ctx->resY), scaledBlur));
	}
	return CKD_OK;
}

// Ball_Draw, ball.cpp:452-514
extern "C" int ckd_ball_draw(ckd_ctx *ctx, const ckd_ball_params *p, float time, uint32_t *d_dest)
{
	CKD_TRY(ApplyShortSpanOverride());
	CKD_REQUIRE(ctx && p && d_dest, "null argument");
	for (int i = 0; i < 5; ++i)
		CKD_TRY(RequireImage(ctx, ckd_image(CKD_IMG_BALL_HEIGHT0 + i), 1024, 1024, 1, "ball height map"));
	CKD_TRY(RequireImage(ctx, CKD_IMG_BALL_COLOR0, 1024, 1024, 4, "ball colour map 0"));
	CKD_TRY(RequireImage(ctx, CKD_IMG_BALL_COLOR1, 1024, 1024, 4, "ball colour map 1"));
	CKD_TRY(RequireImage(ctx, CKD_IMG_BALL_ENV, 1024, 1024, 4, "ball env map"));
	for (int i = 0; i < 3; ++i)
		CKD_TRY(RequireImage(ctx, ckd_image(CKD_IMG_BALL_BEAM0 + i), 1024, 1024, 4, "ball beam map"));
	CKD_TRY(RequireImage(ctx, CKD_IMG_BALL_BACKGROUND0, ctx->resX, ctx->resY, 4, "ball background 0 (output sized)"));
	CKD_TRY(RequireImage(ctx, CKD_IMG_BALL_BACKGROUND1, ctx->resX, ctx->resY, 4, "ball background 1 (output sized)"));
	CKD_TRY(RequireImage(ctx, CKD_IMG_BALL_HALO, ctx->resX, ctx->resY, 4, "ball halo (output sized)"));

	constexpr size_t mapNumPixels = 1024*1024;
	const bool hasBeams = 0 != p->has_beams;

	// height map mix, ball.cpp:458-464
	const unsigned iBaseMap = unsigned(ckdh::clampi(1, 4, p->base_shape_index));
	CKD_CUDA(cudaMemcpyAsync(ctx->d_ballHeightMix, ctx->images[CKD_IMG_BALL_HEIGHT0 + iBaseMap].d_pixels, mapNumPixels, cudaMemcpyDeviceToDevice, ctx->stream));
	const uint8_t spikes = uint8_t(p->spikes);
	if (0 != spikes)
		CKD_TRY(ckd_blend(ctx, CKD_MIX32, reinterpret_cast<uint32_t *>(ctx->d_ballHeightMix), static_cast<const uint32_t *>(ctx->images[CKD_IMG_BALL_HEIGHT0].d_pixels), unsigned(mapNumPixels/4), 0.f, spikes));

	if (hasBeams)
	{
		// beam map mix, ball.cpp:466-480
		const float beamA[3] = { ckdh::saturatef(p->beams1), ckdh::saturatef(p->beams2), ckdh::saturatef(p->beams3) };
		CKD_TRY(ckd_memset32(ctx, ctx->d_ballBeamMix, 0, mapNumPixels));
		for (int i = 0; i < 3; ++i)
			if (beamA[i] > 0.f)
				CKD_TRY(ckd_blit(ctx, CKD_BLITADD32A, ctx->d_ballBeamMix, static_cast<const uint32_t *>(ctx->images[CKD_IMG_BALL_BEAM0 + i].d_pixels), 1024, 1024, 1024, beamA[i]));
	}

	// vball(g_renderTarget[0], time*speed), ball.cpp:312-365
	const float ballTime = time * p->speed;

	// vball_precalc, ball.cpp:283-308 (host libm, like the reference)
	static thread_local int tables[4096];
	static thread_local int rayDeltas[16384*2];
	CKD_REQUIRE(ctx->resY <= 16384, "resolution too large");
	const float radius = ckdh::clampf(1.f, 1920.f, p->radius);
	const unsigned rayLength = unsigned(ckdh::clampi(1, 1024, p->ray_length));
	const float angStepSin = ckdh::kPI/float(rayLength-1);
	const float angStepCos = angStepSin * 0.99f;
	for (unsigned iAngle = 0; iAngle < rayLength; ++iAngle)
	{
		tables[iAngle] = int(ckdh::x86_f2u(radius*sinf(angStepSin*float(iAngle))));
		const float cosine = cosf(angStepCos*float(iAngle));
		if (cosine >= 0.f)
		{
			tables[1024+iAngle] = ckdh::x86_cvtt(255.f*powf(cosine, ckdh::kGoldenRatio));
			tables[2048+iAngle] = ckdh::x86_cvtt(255.f*powf(cosine, ckdh::kGoldenAngle));
			tables[3072+iAngle] = ckdh::x86_cvtt(255.f*powf(cosine, ckdh::kPI));
		}
		else
			tables[1024+iAngle] = tables[2048+iAngle] = tables[3072+iAngle] = 0;
	}

	BallFrame f;
	f.resX = ctx->resX;
	f.resY = ctx->resY;
	f.rayLength = rayLength;
	f.beamAtten = unsigned(ckdh::clampi(0, 255, p->beam_atten));
	f.beamAlphaMin = ckdh::clampf(0.f, 255.f, p->beam_alpha_min);
	f.lowLight = unsigned(ckdh::clampi(0, 255, p->low_beams));
	f.clearLastPixel = ctx->frameIndependent ? 1 : 0;

	const float timeScale = float(rayLength)*(0.25f/1024);
	const float fMapDim = 1024.f, fMapHalf = fMapDim*0.5f;
	f.fromX = ckdh::ftofp24(fMapDim*sinf(ballTime*timeScale) + fMapHalf + p->rotate_offs_x);
	f.fromY = ckdh::ftofp24(fMapDim*cosf(ballTime*timeScale) + fMapHalf + p->rotate_offs_y);

	// fan deltas, ball.cpp:334-363 + voxel-shared.h:16-31
	const float delta = ckdh::k2PI/float(ctx->resY-1);
	for (int iRay = 0; iRay < ctx->resY; ++iRay)
	{
		const float curAngle = float(unsigned(iRay))*delta;
		float dX = cosf(curAngle), dY = sinf(curAngle);
		if (fabsf(dX+dY) > ckdh::kEpsilon)
		{
			const float length = 1.f/sqrtf(dX*dX + dY*dY);
			dX *= length;
			dY *= length;
		}
		rayDeltas[iRay*2] = ckdh::ftofp24(dX);
		rayDeltas[iRay*2+1] = ckdh::ftofp24(dY);
	}

	int *d_tables = ctx->d_voxelTables;
	int *d_rayDeltas = reinterpret_cast<int *>(ctx->d_rayParams);
	CKD_CUDA(cudaMemcpyAsync(d_tables, tables, sizeof(int)*4096, cudaMemcpyHostToDevice, ctx->stream));
	CKD_CUDA(cudaMemcpyAsync(d_rayDeltas, rayDeltas, sizeof(int)*2*ctx->resY, cudaMemcpyHostToDevice, ctx->stream));

	const unsigned blocks = ckd_div_up(ctx->resY, kBallRowsPerBlock);
	if (hasBeams)
	{
		ckd_prof_begin(ctx, "voxel_ball_beams", 4.0*ctx->resX*ctx->resY);
		ball_kernel<true><<<blocks, kBallRowsPerBlock*32, 0, ctx->stream>>>(ctx->d_renderTarget[0], ctx->d_ballHeightMix,
			static_cast<const uint32_t *>(ctx->images[CKD_IMG_BALL_COLOR0].d_pixels), ctx->d_ballBeamMix, d_tables, d_rayDeltas, f);
	}
	else
	{
		ckd_prof_begin(ctx, "voxel_ball", 4.0*ctx->resX*ctx->resY);
		ball_kernel<false><<<blocks, kBallRowsPerBlock*32, 0, ctx->stream>>>(ctx->d_renderTarget[0], ctx->d_ballHeightMix,
			static_cast<const uint32_t *>(ctx->images[CKD_IMG_BALL_COLOR1].d_pixels), static_cast<const uint32_t *>(ctx->images[CKD_IMG_BALL_ENV].d_pixels), d_tables, d_rayDeltas, f);
	}
	CKD_CHECK_LAUNCH(ctx);

	// ball.cpp:486-502
	const float blur = ckdh::BoxBlurScale(p->blur);
	if (0.f != blur)
		CKD_TRY(ckd_old_blur_h(ctx, ctx->d_renderTarget[0], ctx->d_renderTarget[0], unsigned(ctx->resX), unsigned(ctx->resY), blur));

	const void *pBackground = ctx->images[hasBeams ? CKD_IMG_BALL_BACKGROUND0 : CKD_IMG_BALL_BACKGROUND1].d_pixels;
	CKD_CUDA(cudaMemcpyAsync(d_dest, pBackground, size_t(ctx->resX)*ctx->resY*4, cudaMemcpyDeviceToDevice, ctx->stream));
	return ckd_polar_tail(ctx, d_dest, ctx->d_renderTarget[0], 0, true, hasBeams ? static_cast<const uint32_t *>(ctx->images[CKD_IMG_BALL_HALO].d_pixels) : nullptr);
}

// Twister_Draw, torus-twister.cpp:166-188
extern "C" int ckd_twister_draw(ckd_ctx *ctx, const ckd_twister_params *p, float time, uint32_t *d_dest)
{
	CKD_TRY(ApplyShortSpanOverride());
	CKD_REQUIRE(ctx && p && d_dest, "null argument");
	CKD_TRY(RequireImage(ctx, CKD_IMG_TWISTER_HEIGHT, 1024, 1024, 1, "twister height map"));
	CKD_TRY(RequireImage(ctx, CKD_IMG_TWISTER_COLOR, 1024, 1024, 4, "twister colour map"));
	CKD_TRY(RequireImage(ctx, CKD_IMG_TWISTER_BACKGROUND, ctx->resX, ctx->resY, 4, "twister background (output sized)"));
	CKD_REQUIRE(ctx->resY <= 16384, "resolution too large");

	// vtwister_precalc, torus-twister.cpp:119-135
	static thread_local int tables[1024];
	static thread_local int rayOrigins[16384*2];
	for (unsigned iAngle = 0; iAngle < 512; ++iAngle)
	{
		const float angle = ckdh::kPI/(512-1) * iAngle;
		const float scale = 600.f*sinf(angle);
		tables[iAngle] = int(ckdh::x86_f2u(scale));
		const float cosine = cosf(angle*0.99f);
		tables[512+iAngle] = (cosine > 0.f) ? int(ckdh::x86_f2u(255.f*powf(cosine, 2.f))) : 0;
	}

	// vtwister, torus-twister.cpp:92-117
	const float fMapSize = 1024.f;
	const float fMapSizeHH = (fMapSize*0.5f) - 0.5f;
	const float fMapSizeHHH = (fMapSize*0.25f) - 0.5f;
	const float mapStepY = fMapSize/float(ctx->resY-1);
	const float shearStep = ckdh::k2PI/float(ctx->resY-1);
	for (int iRay = 0; iRay < ctx->resY; ++iRay)
	{
		const float shearAngle = float(iRay) * shearStep;
		const float mapY = float(unsigned(iRay))*mapStepY;
		rayOrigins[iRay*2] = ckdh::ftofp24(fMapSizeHH + fMapSizeHHH*sinf(time*p->shear_speed + shearAngle));
		rayOrigins[iRay*2+1] = ckdh::ftofp24(mapY + time*p->speed);
	}

	int *d_tables = ctx->d_voxelTables;
	int *d_rayOrigins = reinterpret_cast<int *>(ctx->d_rayParams);
	CKD_CUDA(cudaMemcpyAsync(d_tables, tables, sizeof(int)*1024, cudaMemcpyHostToDevice, ctx->stream));
	CKD_CUDA(cudaMemcpyAsync(d_rayOrigins, rayOrigins, sizeof(int)*2*ctx->resY, cudaMemcpyHostToDevice, ctx->stream));

	ckd_prof_begin(ctx, "voxel_twister", 4.0*ctx->resX*ctx->resY);
	twister_kernel<<<ckd_div_up(ctx->resY, kRowsPerBlock), kRowsPerBlock*64, 0, ctx->stream>>>(ctx->d_renderTarget[0],
		static_cast<const uint8_t *>(ctx->images[CKD_IMG_TWISTER_HEIGHT].d_pixels), static_cast<const uint32_t *>(ctx->images[CKD_IMG_TWISTER_COLOR].d_pixels),
		d_tables, d_rayOrigins, ctx->resX, ctx->resY);
	CKD_CHECK_LAUNCH(ctx);

	const float blur = p->blur;
	if (0.f != blur)
		CKD_TRY(ckd_old_blur_h(ctx, ctx->d_renderTarget[0], ctx->d_renderTarget[0], unsigned(ctx->resX), unsigned(ctx->resY), ckdh::BoxBlurScale(blur)));

	CKD_CUDA(cudaMemcpyAsync(d_dest, ctx->images[CKD_IMG_TWISTER_BACKGROUND].d_pixels, size_t(ctx->resX)*ctx->resY*4, cudaMemcpyDeviceToDevice, ctx->stream));
	return ckd_polar_tail(ctx, d_dest, ctx->d_renderTarget[0], 0, true, nullptr);
}

// The beam tail of vball_ray_beams on its own (ball.cpp:168-203), for `rows` rows of `row_pixels` pixels: row r is a ray whose
// spans ended `first_remainder + r` pixels before the last pixel of the row.  What the tail does not write is left alone.
// raw_steps != 0 stores the float bits of the accumulated smoothstep parameter instead of the pixel.
extern "C" int ckd_ball_beam_tail(ckd_ctx *ctx, uint32_t *d_rows, int row_pixels, int rows, int first_remainder, uint32_t beam_color, float beam_alpha_min, int raw_steps)
{
	CKD_REQUIRE(ctx && d_rows, "null argument");
	CKD_REQUIRE(row_pixels >= 1 && rows >= 1 && first_remainder >= 0 && first_remainder + rows - 1 <= row_pixels - 1, "the tails must fit their rows");
	if (raw_steps)
		beam_tail_kernel<true><<<ckd_div_up(rows, kRowsPerBlock), kRowsPerBlock*32, 0, ctx->stream>>>(d_rows, row_pixels, rows, unsigned(first_remainder), beam_color, beam_alpha_min);
	else
		beam_tail_kernel<false><<<ckd_div_up(rows, kRowsPerBlock), kRowsPerBlock*32, 0, ctx->stream>>>(d_rows, row_pixels, rows, unsigned(first_remainder), beam_color, beam_alpha_min);
	CKD_CHECK_LAUNCH(ctx);
	return CKD_OK;
}

// ckd_blur.cu -- the two box blurs of the reference.
//
//  * "old" 2007 blur (deprecated/boxblur.cpp): the one every live effect calls, almost always IN PLACE.  Per line it is a
//    running 16-bit saturating accumulator whose trailing edge re-reads pixels the same line already wrote, i.e. a
//    non-linear recurrence along the line: lines and channels are independent, positions are not.  The kernel therefore
//    runs one thread per (line, channel) and walks the line; a warp covers 8 neighbouring lines x 4 channels, so a
//    vertical pass touches one 32-byte sector per step and a horizontal pass keeps 8 sectors hot in L1.
//  * "new" 2026 blur (boxblur.cpp): 32-bit sums without saturation, N ping-pong passes, 10:22 fixed-point scale.
//    Same thread mapping; the transposes of the reference (Transpose32) disappear because a line is addressed with a
//    (lineStride, stepStride) pair -- vertical passes simply walk columns.

#include "ckd_internal.h"
#include "ckd_hostmath.h"

// ---------------------------------------------------------------------------------------------------------------
// old blur -- deprecated/boxblur.cpp:44-231
// ---------------------------------------------------------------------------------------------------------------

struct OldBlurSetup
{
	unsigned edgeSpan, kernelMedian, remainderShift, startWeight, fullPassLen, fullDiv;
	int subEdges;
};

// WeightToDiv, deprecated/boxblur.cpp:11-14; the value lands in 16-bit lanes (_mm_set1_epi16): keep the low 16 bits
__host__ __device__ static inline unsigned WeightToDiv16(unsigned weight) { return (((65536u*256u)/weight) >> 4) & 0xffffu; }

// Div, deprecated/boxblur.cpp:39-42: pmulhuw then packuswb (the word is read as signed: >= 0x8000 -> 0)
__device__ __forceinline__ unsigned old_div(unsigned acc, unsigned div16)
{
	const unsigned v = (acc*div16) >> 16;
	return (v > 32767u) ? 0u : min(v, 255u);
}

__global__ void __launch_bounds__(128) old_blur_kernel(uint8_t *pDest, const uint8_t *pSrc, unsigned numLines, size_t lineStride, size_t stepStride, OldBlurSetup s)
{
	const unsigned t = blockIdx.x*blockDim.x + threadIdx.x;
	const unsigned line = t >> 2, chan = t & 3;
	if (line >= numLines)
		return;

	// byte addressing: pixel i of this line, channel 'chan'
	const uint8_t *src = pSrc + (size_t(line)*lineStride)*4 + chan;
	uint8_t *dst = pDest + (size_t(line)*lineStride)*4 + chan;
	const size_t step = stepStride*4;

	unsigned acc = 0, addRem = 0, subRem = 0;
	size_t addPos = 0, subPos = 0, destPos = 0;

	// Add / Sub, deprecated/boxblur.cpp:17-36 (per 16-bit lane; two saturating adds of non-negative terms == one)
	auto Add = [&](unsigned px)
	{
		acc = min(acc + addRem, 65535u);
		addRem = px >> s.remainderShift;
		acc = min(acc + (px - addRem), 65535u);
	};
	auto Sub = [&](unsigned px)
	{
		acc = acc - min(acc, subRem);
		subRem = px >> s.remainderShift;
		acc = acc - min(acc, px - subRem);
	};

	// pre-read: bring accumulator up to edge weight
	for (unsigned i = 0; i < s.edgeSpan; ++i, addPos += step)
		Add(src[addPos]);

	// pre-pass: up to full weight
	for (unsigned i = 0; i < s.kernelMedian; ++i, addPos += step, destPos += step)
	{
		Add(src[addPos]);
		dst[destPos] = uint8_t(old_div(acc, WeightToDiv16(s.startWeight + 16*i)));
	}

	// main pass (in place, src[subPos] is a pixel this very thread wrote kernelMedian steps ago)
	for (unsigned i = 0; i < s.fullPassLen; ++i, addPos += step, subPos += step, destPos += step)
	{
		Add(src[addPos]);
		Sub(src[subPos]);
		dst[destPos] = uint8_t(old_div(acc, s.fullDiv));
	}

	if (s.subEdges)
		acc = min(acc + addRem, 65535u);

	// post-pass: back to median weight
	for (unsigned i = s.edgeSpan; i > 0; --i, subPos += step, destPos += step)
	{
		Sub(src[subPos]);
		dst[destPos] = uint8_t(old_div(acc, WeightToDiv16(s.startWeight + 16*(i-1))));
	}
}

static int OldBlurPass(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned numLines, unsigned lineLen, size_t lineStride, size_t stepStride, float strength)
{
	// deprecated/boxblur.cpp:53-76
	const float fKernelSpan = strength*255.f;
	unsigned kernelSpan = ckdh::x86_f2u(fKernelSpan);
	kernelSpan = unsigned(ckdh::clampi(1, 255, int(kernelSpan)));

	OldBlurSetup s;
	s.subEdges = (kernelSpan & 1) == 0;
	s.edgeSpan = kernelSpan >> 1;
	s.remainderShift = 1 + ((!s.subEdges)*7);
	s.kernelMedian = s.edgeSpan + !s.subEdges;
	s.startWeight = (s.kernelMedian << 4) + (s.subEdges << 3);
	CKD_REQUIRE(lineLen >= s.kernelMedian + s.edgeSpan, "image smaller than the blur kernel (the reference would run off the buffer)");
	s.fullPassLen = lineLen - (s.kernelMedian + s.edgeSpan);
	s.fullDiv = WeightToDiv16(kernelSpan << 4);

	const unsigned threads = numLines*4;
	ckd_prof_begin(ctx, stepStride == 1 ? "old_blur_h" : "old_blur_v", 8.0*numLines*lineLen);
	old_blur_kernel<<<ckd_div_up(threads, 128), 128, 0, ctx->stream>>>(reinterpret_cast<uint8_t *>(d_dest), reinterpret_cast<const uint8_t *>(d_src), numLines, lineStride, stepStride, s);
	CKD_CHECK_LAUNCH(ctx);
	return CKD_OK;
}

extern "C" int ckd_old_blur_h(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength)
{
	CKD_REQUIRE(ctx && d_dest && d_src, "null argument");
	return OldBlurPass(ctx, d_dest, d_src, y_res, x_res, x_res, 1, strength);
}

extern "C" int ckd_old_blur_v(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength)
{
	CKD_REQUIRE(ctx && d_dest && d_src, "null argument");
	return OldBlurPass(ctx, d_dest, d_src, x_res, y_res, 1, x_res, strength);
}

// BoxBlur32, deprecated/boxblur.cpp:222-231: horizontal, then vertical in place
extern "C" int ckd_old_blur(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength)
{
	CKD_TRY(ckd_old_blur_h(ctx, d_dest, d_src, x_res, y_res, strength));
	return ckd_old_blur_v(ctx, d_dest, d_dest, x_res, y_res, strength);
}

extern "C" float ckd_box_blur_scale(float strength) { return ckdh::BoxBlurScale(strength); }

// ---------------------------------------------------------------------------------------------------------------
// new blur -- boxblur.cpp:87-318
// ---------------------------------------------------------------------------------------------------------------

constexpr unsigned kNewBlurMaxRadius = 500; // boxblur.cpp:79

struct NewBlurSetup
{
	unsigned iSpan;
	int iAlpha;
	unsigned iScale;      // low dword of the 10:22 scale (what _mm_mul_epu32 reads)
	float halfScale, dScale;
};

// iDiv + v2cISSE32, boxblur.cpp:63-73, util.h:152-154 -- one lane
__device__ __forceinline__ unsigned new_div_pack(int iSum, unsigned scale)
{
	const unsigned long long q = (static_cast<unsigned long long>(unsigned(iSum))*scale) >> 22;
	const int v = int(unsigned(q) + unsigned(q >> 32)); // phaddd adds the low and the high dword
	if (v < 0) return 0u;                               // packusdw
	const unsigned w = unsigned(min(v, 65535));
	return (w > 32767u) ? 0u : min(w, 255u);            // packuswb
}

// ftofp<int64_t>(value, 22) low dword (boxblur.cpp:113,120; util.h:199-202)
__device__ __forceinline__ unsigned scale_fp22(float value)
{
	const float scaled = value*4194304.f;
	if (!(scaled >= -9223372036854775808.f && scaled < 9223372036854775808.f))
		return 0u;
	return unsigned(__float2ll_rz(scaled));
}

// One (line, channel) per thread.  Element i of a line lives at base + line*lineStride + i*stepStride (in pixels); the
// reference reads up to 2 elements past the end of a line (boxblur.cpp:171,185), which in its (transposed) layout are the
// first elements of the next line, or whatever follows the buffer for the last line: 'numLines' bounds that to zero here.
__global__ void __launch_bounds__(128) new_blur_kernel(uint8_t *pDest, const uint8_t *pSrc, unsigned numLines, unsigned lineLen,
	size_t srcLineStride, size_t srcStepStride, size_t dstLineStride, size_t dstStepStride, NewBlurSetup s)
{
	const unsigned t = blockIdx.x*blockDim.x + threadIdx.x;
	const unsigned line = t >> 2, chan = t & 3;
	if (line >= numLines)
		return;

	auto Read = [&](unsigned i) -> int
	{
		unsigned l = line;
		if (i >= lineLen) { i -= lineLen; ++l; }
		if (l >= numLines) return 0;
		return int(pSrc[(size_t(l)*srcLineStride + size_t(i)*srcStepStride)*4 + chan]);
	};

	uint8_t *dst = pDest + (size_t(line)*dstLineStride)*4 + chan;
	const size_t dstStep = dstStepStride*4;
	size_t writeIdx = 0;

	int iSum = 0;
	unsigned tail = 0, head = 0;

	// calculate sum at first pixel (median), boxblur.cpp:150-157
	for (unsigned i = 0; i < s.iSpan; ++i)
		iSum += Read(head++);
	iSum += (Read(head)*s.iAlpha) >> 16;

	// iAdd / iSub, boxblur.cpp:37-61: sum +/-= A + (((B-A)*alpha16) >> 16), arithmetic shift
	int headA = Read(head+1), headB;
	for (unsigned i = 0; i < s.iSpan; ++i)
	{
		dst[writeIdx] = uint8_t(new_div_pack(iSum, scale_fp22(s.halfScale + float(i)*s.dScale)));
		writeIdx += dstStep;
		headB = Read(head+2);
		iSum += headA + (((headB-headA)*s.iAlpha) >> 16);
		headA = headB;
		++head;
	}

	int tailA = Read(tail), tailB;
	const unsigned fullLen = lineLen - s.iSpan*2;
	for (unsigned i = 0; i < fullLen; ++i)
	{
		dst[writeIdx] = uint8_t(new_div_pack(iSum, s.iScale));
		writeIdx += dstStep;
		headB = Read(head+2);
		iSum += headA + (((headB-headA)*s.iAlpha) >> 16);
		headA = headB;
		++head;
		tailB = Read(tail+1);
		iSum -= tailA + (((tailB-tailA)*s.iAlpha) >> 16);
		tailA = tailB;
		++tail;
	}

	for (unsigned i = s.iSpan; i > 0; --i)
	{
		dst[writeIdx] = uint8_t(new_div_pack(iSum, scale_fp22(s.halfScale + float(i-1)*s.dScale)));
		writeIdx += dstStep;
		tailB = Read(tail+1);
		iSum -= tailA + (((tailB-tailA)*s.iAlpha) >> 16);
		tailA = tailB;
		++tail;
	}
}

// HorzBlur32, boxblur.cpp:87-212, on lines described by strides (natural layout, no transposes)
static int NewBlurLines(ckd_ctx *ctx, uint32_t *d_dest, uint32_t *d_scratch, const uint32_t *d_src,
	unsigned numLines, unsigned lineLen, size_t lineStride, size_t stepStride, float strength, float gain, unsigned numPasses)
{
	CKD_REQUIRE(numPasses > 0, "numPasses must be > 0");

	// boxblur.cpp:101-121
	strength *= 0.01f;
	const float radius = ckdh::stdmin(float(kNewBlurMaxRadius), strength*float((lineLen-2)/2));
	const unsigned iSpan = ckdh::x86_f2u(radius);
	CKD_REQUIRE(lineLen >= iSpan*2, "image smaller than the blur kernel");

	const float scale = 1.f/((2.f-gain)*radius + 1.f);
	const float alpha = radius - float(iSpan);

	NewBlurSetup s;
	s.iSpan = iSpan;
	{
		const float scaled = scale*float(1<<22);
		s.iScale = (scaled >= -9223372036854775808.f && scaled < 9223372036854775808.f) ? unsigned(uint64_t(int64_t(scaled))) : 0u;
	}
	s.iAlpha = ckdh::x86_cvtt(65536.f*alpha);
	s.halfScale = scale*0.5f;
	s.dScale = s.halfScale/float(iSpan);

	const uint32_t *pRead = d_src;
	uint32_t *pDest = d_dest, *pScratch = d_scratch;
	if (0 == (numPasses & 1))
		std::swap(pDest, pScratch);

	const unsigned threads = numLines*4;
	for (unsigned iPass = 0; iPass < numPasses; ++iPass)
	{
		ckd_prof_begin(ctx, stepStride == 1 ? "new_blur_h" : "new_blur_v", 8.0*numLines*lineLen);
		new_blur_kernel<<<ckd_div_up(threads, 128), 128, 0, ctx->stream>>>(reinterpret_cast<uint8_t *>(pDest), reinterpret_cast<const uint8_t *>(pRead),
			numLines, lineLen, lineStride, stepStride, lineStride, stepStride, s);
		CKD_CHECK_LAUNCH(ctx);
		pRead = pDest;
		std::swap(pDest, pScratch);
	}
	return CKD_OK;
}

extern "C" int ckd_new_blur_h(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength, float gain, unsigned num_passes)
{
	CKD_REQUIRE(ctx && d_dest && d_src, "null argument");
	CKD_REQUIRE(size_t(x_res)*y_res <= size_t(ctx->resX)*ctx->resY, "image larger than the context's scratch buffers");
	return NewBlurLines(ctx, d_dest, ctx->d_scratch[0], d_src, y_res, x_res, x_res, 1, strength, gain, num_passes);
}

extern "C" int ckd_new_blur_v(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength, float gain, unsigned num_passes)
{
	CKD_REQUIRE(ctx && d_dest && d_src, "null argument");
	CKD_REQUIRE(size_t(x_res)*y_res <= size_t(ctx->resX)*ctx->resY, "image larger than the context's scratch buffers");
	// BoxBlur_Vert32, boxblur.cpp:281-299: lines are columns (length yRes), radius derives from yRes
	return NewBlurLines(ctx, d_dest, ctx->d_scratch[0], d_src, x_res, y_res, 1, x_res, strength, gain, num_passes);
}

extern "C" int ckd_new_blur(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength, float gain, unsigned num_passes)
{
	CKD_REQUIRE(ctx && d_dest && d_src, "null argument");
	CKD_REQUIRE(size_t(x_res)*y_res <= size_t(ctx->resX)*ctx->resY, "image larger than the context's scratch buffers");
	// BoxBlur_32, boxblur.cpp:301-318
	CKD_TRY(NewBlurLines(ctx, ctx->d_scratch[1], ctx->d_scratch[0], d_src, y_res, x_res, x_res, 1, strength, gain, num_passes));
	return NewBlurLines(ctx, d_dest, ctx->d_scratch[0], ctx->d_scratch[1], x_res, y_res, 1, x_res, strength, gain, num_passes);
}

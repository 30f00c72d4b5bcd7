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

// ---- staged version: the production path -----------------------------------------------------------------------------
// One WARP owns 8 neighbouring lines (lanes = 8 lines x 4 channels).  The lines' pixels stream through a shared-memory
// ring that is filled ahead of time with 16-byte cp.async copies (kPrefetch stages of 64 steps in flight), so the serial
// walk never waits for HBM; results go to a second ring from which (a) the in-place trailing edge re-reads the pixels
// "it already wrote" and (b) finished 64-step chunks are flushed with coalesced 16-byte stores.  Every global byte is
// read once and written once; what remains is the dependent add/clamp chain of the recurrence itself.

constexpr unsigned kRing = 512;                 // ring length in steps: >= 255 (widest kernel) + (kPrefetch+1)*kStage
constexpr unsigned kStage = 64;                 // steps per cp.async group / per flush
constexpr int kPrefetch = 3;                    // groups in flight ahead of the one being consumed
constexpr unsigned kPitchH = kRing*4 + 16;      // bytes per line in the horizontal layout (+16: lines land in different banks)
constexpr unsigned kRingBytes = 8*kPitchH;      // >= kRing*32 (vertical layout)

__device__ __forceinline__ void cp_async16(void *smemDst, const void *gmemSrc, bool valid)
{
	const unsigned dst = unsigned(__cvta_generic_to_shared(smemDst));
	const int bytes = valid ? 16 : 0; // 0 -> zero fill
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(gmemSrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

constexpr int kBlurWarps = 4;                   // warps per CTA: one per SM sub-partition (a 1-warp CTA always lands on sub-partition 0)

template <bool VERT, bool INPLACE, bool SUBEDGES>
__global__ void __launch_bounds__(kBlurWarps*32) old_blur_staged_kernel(uint8_t *pDest, const uint8_t *pSrc, unsigned numLines, unsigned len, unsigned pitch, OldBlurSetup s)
{
	extern __shared__ __align__(16) uint8_t s_rings[]; // per warp: input ring, output ring
	const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint8_t *s_in = s_rings + warp*(2*kRingBytes);
	uint8_t *s_out = s_in + kRingBytes;

	const unsigned r = lane >> 2, chan = lane & 3;
	const unsigned line0 = (blockIdx.x*kBlurWarps + warp)*8;
	if (line0 >= numLines)
		return;

	constexpr unsigned kStep = VERT ? 32 : 4;       // ring bytes between consecutive steps of one line
	const unsigned laneBase = VERT ? (r*4 + chan) : (r*kPitchH + chan);
	auto ringAt = [&](unsigned pos) -> unsigned { return laneBase + (pos & (kRing-1))*kStep; };

	// global <-> ring transfers of the 64 steps starting at p0, 16 bytes per lane and instruction
	auto loadStage = [&](unsigned p0)
	{
		#pragma unroll
		for (unsigned k = 0; k < 4; ++k)
		{
			if (VERT)
			{
				const unsigned pos = p0 + k*16 + (lane >> 1), col = line0 + (lane & 1)*4;
				const bool valid = pos < len && col + 3 < numLines;
				cp_async16(s_in + (pos & (kRing-1))*32 + (lane & 1)*16, valid ? pSrc + (size_t(pos)*pitch + col)*4 : pSrc, valid);
			}
			else
			{
				const unsigned pos = p0 + (k*4 + (lane & 3))*4, line = line0 + (lane >> 2);
				const bool valid = pos < len && line < numLines;
				cp_async16(s_in + (lane >> 2)*kPitchH + (pos & (kRing-1))*4, valid ? pSrc + (size_t(line)*pitch + pos)*4 : pSrc, valid);
			}
		}
	};
	auto flushStage = [&](unsigned p0)
	{
		uint4 v[4]; // all four shared-memory reads first, then the four global stores
		#pragma unroll
		for (unsigned k = 0; k < 4; ++k)
		{
			if (VERT)
			{
				const unsigned pos = p0 + k*16 + (lane >> 1);
				v[k] = *reinterpret_cast<const uint4 *>(s_out + (pos & (kRing-1))*32 + (lane & 1)*16);
			}
			else
			{
				const unsigned pos = p0 + (k*4 + (lane & 3))*4;
				v[k] = *reinterpret_cast<const uint4 *>(s_out + (lane >> 2)*kPitchH + (pos & (kRing-1))*4);
			}
		}
		#pragma unroll
		for (unsigned k = 0; k < 4; ++k)
		{
			if (VERT)
			{
				const unsigned pos = p0 + k*16 + (lane >> 1), col = line0 + (lane & 1)*4;
				if (pos < len && col + 3 < numLines)
					*reinterpret_cast<uint4 *>(pDest + (size_t(pos)*pitch + col)*4) = v[k];
			}
			else
			{
				const unsigned pos = p0 + (k*4 + (lane & 3))*4, line = line0 + (lane >> 2);
				if (pos < len && line < numLines)
					*reinterpret_cast<uint4 *>(pDest + (size_t(line)*pitch + pos)*4) = v[k];
			}
		}
	};

	const int sh = int(s.remainderShift);
	const unsigned edgeSpan = s.edgeSpan, kM = s.kernelMedian, span = edgeSpan + kM;
	const unsigned total = len + edgeSpan;           // step t adds pixel t (while t < len) and emits output t - edgeSpan
	const unsigned numStages = (total + kStage - 1)/kStage;
	const unsigned mainEnd = len - edgeSpan;         // outputs [kM, mainEnd) are the full-weight main pass
	const unsigned fullDiv = s.fullDiv;

	#pragma unroll
	for (int p = 0; p < kPrefetch; ++p)
	{
		loadStage(p*kStage);
		cp_async_commit();
	}

	int acc = 0, addRem = 0, subRem = 0;
	unsigned flushed = 0;
	unsigned long long hist = 0;                    // the last 8 outputs of this lane, newest in the low byte (short in-place kernels)
	const unsigned histShift = (kM - 1)*8;
	const bool shortInPlace = INPLACE && kM < 8;

	// value the trailing edge subtracts at output o: pixel o - kM of the *source as the reference sees it*: already blurred
	// when running in place, the original otherwise
	auto subAt = [&](unsigned pos) -> int { return INPLACE ? int(s_out[ringAt(pos)]) : int(s_in[ringAt(pos)]); };

	// one steady-state step: Add + Sub + Div (deprecated/boxblur.cpp:104-110) with the four saturating 16-bit operations folded:
	// max(min(acc + a, 65535) - b, 0) == max(min(acc + (a - b), 65535 - b), 0), a single DPX add-min-relu.
	// Odd kernels (SUBEDGES == false) shift the remainder out entirely (px >> 8 == 0): a == px, b == spx.
	// Div: pmulhuw + packuswb == min((acc*div) >> 16, 255) whenever div <= 32768 (every kernel wider than 1 pixel; a 1-pixel
	// kernel has div == 0): the "word read as signed" case of packuswb cannot occur, and (acc*div) >> 16 == umulhi(acc, div << 16).
	const unsigned fullDivHi = fullDiv << 16;
	const bool plainDiv = fullDiv <= 32768u;
	auto steady = [&](int px, int spx) -> unsigned
	{
		int a = px, b = spx;
		if (SUBEDGES)
		{
			a = addRem + (px - (px >> 1));
			addRem = px >> 1;
			b = subRem + (spx - (spx >> 1));
			subRem = spx >> 1;
		}
		acc = __viaddmin_s32_relu(acc, a - b, 65535 - b);
		return plainDiv ? min(__umulhi(unsigned(acc), fullDivHi), 255u) : old_div(unsigned(acc), fullDiv);
	};

	for (unsigned stage = 0; stage < numStages; ++stage)
	{
		loadStage((stage + kPrefetch)*kStage);
		cp_async_commit();
		cp_async_wait<kPrefetch>();
		__syncwarp();

		const unsigned t0 = stage*kStage;
		const unsigned t1 = min(t0 + kStage, total);

		if (t0 >= span && t1 <= len && t1 - t0 == kStage)
		{
			for (unsigned tb = t0; tb < t1; tb += 8)
			{
				const uint8_t *inp = s_in + laneBase + (tb & (kRing-1))*kStep; // tb is a multiple of 8: no wrap inside the batch
				const unsigned subPos = (tb - span) & (kRing-1), outPos = (tb - edgeSpan) & (kRing-1);
				int px[8];
				#pragma unroll
				for (int j = 0; j < 8; ++j) px[j] = int(inp[j*kStep]);

				if (shortInPlace)
				{
					// the subtracted pixel was produced fewer than 8 steps ago: take it from the register history
					#pragma unroll
					for (int j = 0; j < 8; ++j)
					{
						const int spx = int((hist >> histShift) & 0xff);
						const unsigned o = steady(px[j], spx);
						hist = (hist << 8) | o;
						s_out[laneBase + ((outPos + j) & (kRing-1))*kStep] = uint8_t(o);
					}
				}
				else
				{
					const uint8_t *subBase = (INPLACE ? s_out : s_in) + laneBase;
					int spx[8];
					if (subPos <= kRing-8 && outPos <= kRing-8)
					{
						#pragma unroll
						for (int j = 0; j < 8; ++j) spx[j] = int(subBase[(subPos + j)*kStep]);
						uint8_t *outp = s_out + laneBase + outPos*kStep;
						#pragma unroll
						for (int j = 0; j < 8; ++j) outp[j*kStep] = uint8_t(steady(px[j], spx[j]));
					}
					else
					{
						#pragma unroll
						for (int j = 0; j < 8; ++j) spx[j] = int(subBase[((subPos + j) & (kRing-1))*kStep]);
						#pragma unroll
						for (int j = 0; j < 8; ++j) s_out[laneBase + ((outPos + j) & (kRing-1))*kStep] = uint8_t(steady(px[j], spx[j]));
					}
				}
			}
		}
		else
		{
			for (unsigned t = t0; t < t1; ++t)
			{
				if (t < len)
				{
					const int px = int(s_in[ringAt(t)]);
					acc = min(acc + addRem + (px - (px >> sh)), 65535);
					addRem = px >> sh;
				}
				else if (t == len && s.subEdges)
					acc = min(acc + addRem, 65535); // deprecated/boxblur.cpp:113-115
				if (t >= edgeSpan)
				{
					const unsigned o = t - edgeSpan;
					unsigned div;
					if (o < kM)
						div = WeightToDiv16(s.startWeight + 16*o);
					else
					{
						const int spx = subAt(o - kM);
						acc = max(acc - (subRem + (spx - (spx >> sh))), 0);
						subRem = spx >> sh;
						div = (o < mainEnd) ? fullDiv : WeightToDiv16(s.startWeight + 16*(len - 1 - o));
					}
					const unsigned outv = old_div(unsigned(acc), div);
					hist = (hist << 8) | outv;
					s_out[ringAt(o)] = uint8_t(outv);
				}
			}
		}
		__syncwarp();

		const unsigned done = (t1 > edgeSpan) ? t1 - edgeSpan : 0;
		while (flushed + kStage <= done)
		{
			flushStage(flushed);
			flushed += kStage;
		}
	}

	cp_async_wait<0>();
	while (flushed < len)
	{
		flushStage(flushed);
		flushed += kStage;
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

	const bool vert = (stepStride != 1);
	const unsigned pitch = unsigned(vert ? stepStride : lineStride); // pixels per image row
	uint8_t *pDest = reinterpret_cast<uint8_t *>(d_dest);
	const uint8_t *pSrc = reinterpret_cast<const uint8_t *>(d_src);
	const bool inPlace = (pDest == pSrc);
	const bool overlap = !inPlace && pDest < pSrc + size_t(numLines)*lineLen*4 && pSrc < pDest + size_t(numLines)*lineLen*4;
	const bool aligned = 0 == (pitch & 3) && 0 == ((reinterpret_cast<uintptr_t>(pDest) | reinterpret_cast<uintptr_t>(pSrc)) & 15);

	ckd_prof_begin(ctx, vert ? "old_blur_v" : "old_blur_h", 8.0*numLines*lineLen);
	if (aligned && !overlap)
	{
		const unsigned blocks = ckd_div_up(numLines, 8*kBlurWarps);
		const size_t smem = size_t(kBlurWarps)*2*kRingBytes;
		#define CKD_BLUR_LAUNCH(V, I, E) do { \
			if (!ctx->blurAttrSet[variant]) { CKD_CUDA(cudaFuncSetAttribute(old_blur_staged_kernel<V, I, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); ctx->blurAttrSet[variant] = true; } \
			old_blur_staged_kernel<V, I, E><<<blocks, kBlurWarps*32, smem, ctx->stream>>>(pDest, pSrc, numLines, lineLen, pitch, s); } while (0)
		const int variant = (vert ? 4 : 0) | (inPlace ? 2 : 0) | (s.subEdges ? 1 : 0);
		switch (variant)
		{
		case 0: CKD_BLUR_LAUNCH(false, false, false); break;
		case 1: CKD_BLUR_LAUNCH(false, false, true); break;
		case 2: CKD_BLUR_LAUNCH(false, true, false); break;
		case 3: CKD_BLUR_LAUNCH(false, true, true); break;
		case 4: CKD_BLUR_LAUNCH(true, false, false); break;
		case 5: CKD_BLUR_LAUNCH(true, false, true); break;
		case 6: CKD_BLUR_LAUNCH(true, true, false); break;
		default: CKD_BLUR_LAUNCH(true, true, true); break;
		}
		#undef CKD_BLUR_LAUNCH
	}
	else
	{
		// unaligned sub-rectangles or partially overlapping buffers: plain byte-wise walk with the reference's exact read/write order
		const unsigned threads = numLines*4;
		old_blur_kernel<<<ckd_div_up(threads, 128), 128, 0, ctx->stream>>>(pDest, pSrc, numLines, lineStride, stepStride, s);
	}
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

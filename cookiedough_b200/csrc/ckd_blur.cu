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
	int noFastPath;           // CKD_BLUR_NOFAST (A/B measurements): the blocked kernel never tries its plain-sum fast path
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
// read once and written once; what remains is the dependent add/clamp chain of the recurrence itself (2 ALU operations
// per step).  To keep that chain the only thing a step waits for, the shared-memory loads of batch i+1 are issued before
// the chain of batch i (the compiler cannot hoist them itself: it must assume the ring stores alias them), and in-place
// kernels narrower than 15 pixels take their trailing edge from registers (KM > 0) instead of reading back the ring.

constexpr unsigned kRing = 512;                 // ring length in steps: >= 255 (widest kernel) + (kPrefetch+1)*kStage
constexpr unsigned kStage = 64;                 // steps per cp.async group / per flush
constexpr int kPrefetch = 3;                    // groups in flight ahead of the one being consumed
constexpr unsigned kMirror = 16;                // the blocked kernel repeats the first steps of the input ring behind its end
constexpr unsigned kPitchH = (kRing + kMirror)*4 + 16; // bytes per line in the horizontal layout (+16: lines land in different banks)
constexpr unsigned kRingBytes = 8*kPitchH;      // >= (kRing + kMirror)*32 (vertical layout)

__device__ __forceinline__ void cp_async16(void *smemDst, const void *gmemSrc, bool valid)
{
	const unsigned dst = unsigned(__cvta_generic_to_shared(smemDst));
	const int bytes = valid ? 16 : 0; // 0 -> zero fill
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(gmemSrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smemDst, const void *gmemSrc, bool valid)
{
	const unsigned dst = unsigned(__cvta_generic_to_shared(smemDst));
	const int bytes = valid ? 4 : 0; // 0 -> zero fill
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst), "l"(gmemSrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

constexpr int kBlurWarps = 4;                   // warps per CTA: one per SM sub-partition (a 1-warp CTA always lands on sub-partition 0)

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: used where a CTA stages one long contiguous run
//      (a whole line of the new blur).  The old blur's rings are filled 256 bytes per line and stage: there a TMA variant (eight
//      bulk copies per stage from warp 0, one mbarrier per ring slot) measured 10-25 % slower than 128 threads x cp.async,
//      see profiles/r01_notes.md. ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void *smemDst, const void *gmemSrc, unsigned bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(smem_u32(smemDst)), "l"(gmemSrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
	unsigned done;
	do
	{
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	}
	while (0 == done);
}

template <bool B> struct BoolTag { static constexpr bool value = B; };

// KM: the kernel median when it is <= 8 and the blur runs in place (trailing edge from registers); 0: any width (from the ring)
template <bool VERT, bool INPLACE, bool SUBEDGES, int KM>
__global__ void __launch_bounds__(kBlurWarps*32) old_blur_staged_kernel(uint8_t *pDest, const uint8_t *pSrc, unsigned numLines, unsigned len, unsigned pitch, OldBlurSetup s)
{
	extern __shared__ __align__(16) uint8_t s_rings[]; // per warp: input ring, output ring
	__shared__ unsigned short s_divTab[128];           // WeightToDiv16(startWeight + 16*i): the ramps at both ends of a line

	const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const unsigned edgeSpan = s.edgeSpan, kM = (KM > 0) ? unsigned(KM) : s.kernelMedian, span = edgeSpan + kM;
	if (threadIdx.x < kM)
		s_divTab[threadIdx.x] = (unsigned short) WeightToDiv16(s.startWeight + 16*threadIdx.x);
	__syncthreads();

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

	constexpr int sh = SUBEDGES ? 1 : 8;             // remainderShift, deprecated/boxblur.cpp:62
	const unsigned total = len + edgeSpan;           // step t adds pixel t (while t < len) and emits output t - edgeSpan
	const unsigned numStages = (total + kStage - 1)/kStage;
	const unsigned mainEnd = len - edgeSpan;         // outputs [kM, mainEnd) are the full-weight main pass
	const unsigned fastBegin = (span + 7) & ~7u;     // steady-state batches cover the steps [fastBegin, fastEnd): Add + Sub + full weight
	const unsigned fastEnd = len & ~7u;
	const unsigned fullDiv = s.fullDiv;

	#pragma unroll
	for (int p = 0; p < kPrefetch; ++p)
	{
		loadStage(p*kStage);
		cp_async_commit();
	}

	int acc = 0, addRem = 0, subRem = 0;
	int hist[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };       // KM > 0: the last KM outputs of this lane, newest first
	unsigned flushed = 0;

	auto pushHist = [&](int o)
	{
		if (KM > 0)
		{
			#pragma unroll
			for (int i = 7; i > 0; --i)
				if (i < KM) hist[i] = hist[i-1];
			hist[0] = o;
		}
	};

	// one steady-state step: Add + Sub + Div (deprecated/boxblur.cpp:104-110) with the four saturating 16-bit operations folded:
	// max(min(acc + a, 65535) - b, 0) == max(min(acc + (a - b), 65535 - b), 0), an add and a min-relu on the dependent chain.
	// Odd kernels (SUBEDGES == false) shift the remainder out entirely (px >> 8 == 0): a == px, b == spx.
	// Div: pmulhuw + packuswb == min((acc*div) >> 16, 255) because the full-weight divisor is <= 32768 for every kernel wider than
	// one pixel (and 0 for a 1-pixel kernel): the "word read as signed" case of packuswb cannot occur, and
	// (acc*div) >> 16 == umulhi(acc, div << 16).
	const unsigned fullDivHi = fullDiv << 16;
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
		return min(__umulhi(unsigned(acc), fullDivHi), 255u);
	};

	// any step, including the ramps at both ends of the line (deprecated/boxblur.cpp:84-102, 113-126)
	auto slowStep = [&](unsigned t)
	{
		if (t < len)
		{
			const int px = int(s_in[ringAt(t)]);
			acc = min(acc + addRem + (px - (px >> sh)), 65535);
			addRem = px >> sh;
		}
		else if (SUBEDGES && t == len)
			acc = min(acc + addRem, 65535); // deprecated/boxblur.cpp:113-115
		if (t >= edgeSpan)
		{
			const unsigned o = t - edgeSpan;
			unsigned div;
			if (o < kM)
				div = s_divTab[o];
			else
			{
				// the trailing edge subtracts pixel o - kM of the *source as the reference sees it*: already blurred when running in place
				const int spx = (KM > 0) ? hist[(KM > 0) ? KM-1 : 0] : int((INPLACE ? s_out : s_in)[ringAt(o - kM)]);
				acc = max(acc - (subRem + (spx - (spx >> sh))), 0);
				subRem = spx >> sh;
				div = (o < mainEnd) ? fullDiv : s_divTab[len - 1 - o];
			}
			const unsigned outv = old_div(unsigned(acc), div);
			pushHist(int(outv));
			s_out[ringAt(o)] = uint8_t(outv);
		}
	};

	// shared-memory reads of one batch of 8 steps
	auto loadPx = [&](unsigned tb, int (&d)[8]) // tb is a multiple of 8: no wrap inside the batch
	{
		const uint8_t *inp = s_in + laneBase + (tb & (kRing-1))*kStep;
		#pragma unroll
		for (int j = 0; j < 8; ++j) d[j] = int(inp[j*kStep]);
	};
	auto loadSpx = [&](unsigned tb, int (&d)[8])
	{
		const uint8_t *base = (INPLACE ? s_out : s_in) + laneBase;
		const unsigned p = (tb - span) & (kRing-1);
		if (p <= kRing-8)
		{
			#pragma unroll
			for (int j = 0; j < 8; ++j) d[j] = int(base[(p + j)*kStep]);
		}
		else
		{
			#pragma unroll
			for (int j = 0; j < 8; ++j) d[j] = int(base[((p + j) & (kRing-1))*kStep]);
		}
	};

	for (unsigned stage = 0; stage < numStages; ++stage)
	{
		loadStage((stage + kPrefetch)*kStage);
		cp_async_commit();
		cp_async_wait<kPrefetch>();
		__syncwarp();

		const unsigned t0 = stage*kStage;
		const unsigned t1 = min(t0 + kStage, total);
		unsigned t = t0;

		for (const unsigned headEnd = min(t1, fastBegin); t < headEnd; ++t)
			slowStep(t);

		const unsigned batchEnd = min(t1, fastEnd); // a multiple of 8
		if (t < batchEnd)
		{
			int px[8], spx[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
			if (KM > 0) loadPx(t, px);

			for (unsigned tb = t; tb < batchEnd; tb += 8)
			{
				// KM > 0: the next batch's pixels are requested before this batch's chain (the compiler cannot hoist them over the
				// ring stores itself); KM == 0: plain loads at the top of the batch measured faster than carrying 16 more registers
				int pxNext[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
				// (the batch after the last one belongs to a stage that may still be in flight: the last batch requests itself again)
				if (KM > 0) loadPx(min(tb + 8, batchEnd - 8), pxNext);
				else { loadPx(tb, px); loadSpx(tb, spx); }

				const unsigned outPos = (tb - edgeSpan) & (kRing-1);
				auto chain = [&](auto wrapTag)
				{
					constexpr bool kWrap = decltype(wrapTag)::value;
					uint8_t *outp = s_out + laneBase + (kWrap ? 0 : outPos*kStep);
					#pragma unroll
					for (int j = 0; j < 8; ++j)
					{
						const int sub = (KM > 0) ? hist[(KM > 0) ? KM-1 : 0] : spx[j];
						const unsigned o = steady(px[j], sub);
						pushHist(int(o));
						if (kWrap) outp[((outPos + j) & (kRing-1))*kStep] = uint8_t(o);
						else outp[j*kStep] = uint8_t(o);
					}
				};
				if (outPos <= kRing-8) chain(BoolTag<false>()); else chain(BoolTag<true>());

				if (KM > 0)
				{
					#pragma unroll
					for (int j = 0; j < 8; ++j) px[j] = pxNext[j];
				}
			}
			t = batchEnd;
		}

		for (; t < t1; ++t)
			slowStep(t);
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

// ---- blocked version: kernels of 15 pixels and wider ------------------------------------------------------------------
// The trailing edge of step t is the output of step t - kM (kM = kernel median), so inside a block of kM consecutive steps
// every operand of the recurrence  acc' = max(min(acc + d, h), 0)  is known up front and the block is a composition of
// kM clamped additions.  Those compose in closed form -- clamp(x + D, L, H) followed by the step (d, h) is
// clamp(x + D + d, step(L), step(H)) -- which makes the walk a scan:
//   1. each of P lanes folds its share of the block (at most QT steps) into one (D, L, H) triple,
//   2. a log2(P)-round shuffle scan composes the triples; applied to the accumulator the block started with, it hands
//      every lane the accumulator its share starts with,
//   3. each lane replays its steps with the real accumulator and emits the outputs.
// A lane's trailing-edge operands are the outputs the same lane produced one block earlier, so they stay in registers.
// One CTA owns 8 lines (32 line-channels x P lanes); the rings, cp.async staging and 16-byte flushes are those of the
// staged kernel above, shared by the CTA.  In place only: out of place the trailing edge is the untouched source, there is
// no recurrence to break up and the staged kernel does.  The input ring repeats its first kMirror steps behind its end so that a lane
// reads its few consecutive steps with immediate offsets.  Bit-exact with the serial walk: the composition is exact
// integer algebra.

template <bool VERT, bool SUBEDGES, int LOG2P, int QT>
__global__ void __launch_bounds__(32 << LOG2P) old_blur_blocked_kernel(uint8_t *pDest, unsigned numLines, unsigned len, unsigned pitch, OldBlurSetup s)
{
	extern __shared__ __align__(16) uint8_t s_rings[]; // input ring, output ring
	__shared__ unsigned short s_divTab[128];           // WeightToDiv16(startWeight + 16*i): the ramps at both ends of a line

	constexpr unsigned P = 1u << LOG2P, C = 32u >> LOG2P; // parts per block, line-channels per warp
	const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const unsigned edgeSpan = s.edgeSpan, kM = s.kernelMedian, span = edgeSpan + kM;
	if (tid < kM)
		s_divTab[tid] = (unsigned short) WeightToDiv16(s.startWeight + 16*tid);

	uint8_t *s_in = s_rings;
	uint8_t *s_out = s_rings + kRingBytes;

	// lane -> (line-channel, part of the block)
	const unsigned part = lane / C, sub = lane % C;
	const unsigned lc = warp*C + sub;
	const unsigned r = lc >> 2, chan = lc & 3;
	const unsigned line0 = blockIdx.x*8;

	// both rings are line-major for either direction (a "line" of the vertical pass is an image column): consecutive steps of
	// a line-channel are 4 bytes apart, the 8 lines land in different banks
	constexpr unsigned kStep = 4;
	const unsigned laneBase = r*kPitchH + chan;

	// 128 work items move the 64 steps starting at p0 between global memory and the rings: 16 bytes each along a row
	// (horizontal), or 4 x 4 bytes, one image row each, 8 neighbouring columns per row (vertical).  A CTA has 64 to 512
	// threads: item w is handled by thread w, w - blockDim.x, ...
	constexpr unsigned kThreads = 32u << LOG2P;
	// Full stages (all 64 steps inside the line, the thread's line inside the image) go through per-item bases computed once:
	// an item's global address is base + p0*stepBytes, its ring address base + (p0 mod ring)*4, the four vertical copies 16 rows
	// apart.  Everything else (the stages around the end of the line, the zero-filled prefetch past it) takes the generic code.
	constexpr unsigned kItems = (kThreads >= 128) ? 1 : 128/kThreads;
	size_t itemG[kItems];
	unsigned itemS[kItems];
	bool itemOk[kItems], itemMirror[kItems];
	bool myItemsInside = true;                       // every item of this thread lies on a line of the image (else: generic code, which zero-fills)
	#pragma unroll
	for (unsigned it = 0; it < kItems; ++it)
	{
		const unsigned w = tid + it*kThreads;
		if (VERT)
		{
			const unsigned col = w & 7;
			itemG[it] = (size_t(w >> 3)*pitch + line0 + col)*4;
			itemS[it] = col*kPitchH + (w >> 3)*4;
			itemOk[it] = w < 128;
			myItemsInside = myItemsInside && (w >= 128 || line0 + col < numLines);
			itemMirror[it] = true;                       // rows (w >> 3) + 0 of a stage that starts the ring: < kMirror
		}
		else
		{
			itemG[it] = (size_t(line0 + (w >> 4))*pitch + (w & 15)*4)*4;
			itemS[it] = (w >> 4)*kPitchH + (w & 15)*16;
			itemOk[it] = w < 128;
			myItemsInside = myItemsInside && (w >= 128 || line0 + (w >> 4) < numLines);
			itemMirror[it] = (w & 15)*4 < kMirror;
		}
	}
	const size_t stepBytes = VERT ? size_t(pitch)*4 : 4;
	const unsigned sInBase = unsigned(__cvta_generic_to_shared(s_in)), sOutBase = unsigned(__cvta_generic_to_shared(s_out));
	auto loadStageFull = [&](unsigned p0)
	{
		const unsigned ringOffs = (p0 & (kRing-1))*4;
		#pragma unroll
		for (unsigned it = 0; it < kItems; ++it)
		{
			if (!itemOk[it])
				continue;
			const uint8_t *g = pDest + itemG[it] + size_t(p0)*stepBytes;
			const unsigned sa = sInBase + itemS[it] + ringOffs;
			if (VERT)
			{
				#pragma unroll
				for (unsigned k = 0; k < 4; ++k)
					asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sa + k*64), "l"(g + size_t(k)*16*stepBytes) : "memory");
				if (0 == ringOffs && itemMirror[it])
					asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sa + kRing*4), "l"(g) : "memory");
			}
			else
			{
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(g) : "memory");
				if (0 == ringOffs && itemMirror[it])
					asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa + kRing*4), "l"(g) : "memory");
			}
		}
	};
	auto flushStageFull = [&](unsigned p0)
	{
		const unsigned ringOffs = (p0 & (kRing-1))*4;
		#pragma unroll
		for (unsigned it = 0; it < kItems; ++it)
		{
			if (!itemOk[it])
				continue;
			uint8_t *g = pDest + itemG[it] + size_t(p0)*stepBytes;
			const unsigned sa = sOutBase + itemS[it] + ringOffs;
			if (VERT)
			{
				uint32_t v[4];
				#pragma unroll
				for (unsigned k = 0; k < 4; ++k)
					asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[k]) : "r"(sa + k*64) : "memory");
				#pragma unroll
				for (unsigned k = 0; k < 4; ++k)
					*reinterpret_cast<uint32_t *>(g + size_t(k)*16*stepBytes) = v[k];
			}
			else
			{
				uint4 v;
				asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sa) : "memory");
				*reinterpret_cast<uint4 *>(g) = v;
			}
		}
	};
	auto loadStage = [&](unsigned p0)
	{
		if (p0 + kStage <= len && myItemsInside)
		{
			loadStageFull(p0);
			return;
		}
		#pragma unroll
		for (unsigned w = tid; w < 128; w += kThreads)
		{
			if (VERT)
			{
				const unsigned col = w & 7, line = line0 + col;
				#pragma unroll
				for (unsigned k = 0; k < 4; ++k)
				{
					const unsigned pos = p0 + (w >> 3) + k*16;
					const bool valid = pos < len && line < numLines;
					const uint8_t *src = valid ? pDest + (size_t(pos)*pitch + line)*4 : pDest;
					cp_async4(s_in + col*kPitchH + (pos & (kRing-1))*4, src, valid);
					if ((pos & (kRing-1)) < kMirror)
						cp_async4(s_in + col*kPitchH + (kRing + (pos & (kRing-1)))*4, src, valid);
				}
			}
			else
			{
				const unsigned pos = p0 + (w & 15)*4, line = line0 + (w >> 4);
				const bool valid = pos < len && line < numLines;
				const uint8_t *src = valid ? pDest + (size_t(line)*pitch + pos)*4 : pDest;
				cp_async16(s_in + (w >> 4)*kPitchH + (pos & (kRing-1))*4, src, valid);
				if ((pos & (kRing-1)) < kMirror)
					cp_async16(s_in + (w >> 4)*kPitchH + (kRing + (pos & (kRing-1)))*4, src, valid);
			}
		}
	};
	auto flushStage = [&](unsigned p0)
	{
		if (p0 + kStage <= len && myItemsInside)
		{
			flushStageFull(p0);
			return;
		}
		#pragma unroll
		for (unsigned w = tid; w < 128; w += kThreads)
		{
			if (VERT)
			{
				const unsigned col = w & 7, line = line0 + col;
				uint32_t v[4];
				#pragma unroll
				for (unsigned k = 0; k < 4; ++k)
					v[k] = *reinterpret_cast<const uint32_t *>(s_out + col*kPitchH + ((p0 + (w >> 3) + k*16) & (kRing-1))*4);
				#pragma unroll
				for (unsigned k = 0; k < 4; ++k)
				{
					const unsigned pos = p0 + (w >> 3) + k*16;
					if (pos < len && line < numLines)
						*reinterpret_cast<uint32_t *>(pDest + (size_t(pos)*pitch + line)*4) = v[k];
				}
			}
			else
			{
				const unsigned pos = p0 + (w & 15)*4, line = line0 + (w >> 4);
				if (pos < len && line < numLines)
					*reinterpret_cast<uint4 *>(pDest + (size_t(line)*pitch + pos)*4) = *reinterpret_cast<const uint4 *>(s_out + (w >> 4)*kPitchH + (pos & (kRing-1))*4);
			}
		}
	};

	constexpr int sh = SUBEDGES ? 1 : 8;             // remainderShift, deprecated/boxblur.cpp:62
	const unsigned total = len + edgeSpan;           // step t adds pixel t (while t < len) and emits output t - edgeSpan
	const unsigned mainEnd = len - edgeSpan;         // outputs [kM, mainEnd) are the full-weight main pass
	const unsigned fullDiv = s.fullDiv, fullDivHi = fullDiv << 16;
	// fast path of a steady block: every accumulator within [0, fastLimit] <=> no clamp anywhere (b <= 255 in every step)
	const unsigned fastLimit = s.noFastPath ? 0u : min(65535u - 255u, fullDiv ? ((256u << 16) - 1u)/fullDiv : 0u); // 0: every block takes the exact path

	// this lane's share of every block: steps [s0, s0 + cnt) of the block, cnt is QT or QT - 1 (the host picks QT = ceil(kM/P))
	const unsigned s0 = (part*kM) >> LOG2P;
	const unsigned cnt = (((part + 1)*kM) >> LOG2P) - s0;
	const unsigned srcCarry = (part == 0) ? (P-1)*C + sub : lane - C;
	const unsigned srcLast = (P-1)*C + sub;

	#pragma unroll
	for (int p = 0; p < kPrefetch; ++p)
	{
		loadStage(p*kStage);
		cp_async_commit();
	}
	int x0 = 0;                                      // accumulator at the start of the block (same on every lane)
	int outPrev[QT];                                 // this lane's outputs of the previous block
	#pragma unroll
	for (int q = 0; q < QT; ++q) outPrev[q] = 0;
	int lastOut1 = 0, lastOut2 = 0;                  // this lane's last output one and two blocks ago

	__syncthreads();                                 // s_divTab

	// shared-memory byte accesses through 32-bit shared addresses (one base, immediate offsets)
	const unsigned inBase = unsigned(__cvta_generic_to_shared(s_in)) + laneBase;
	const unsigned outBase = unsigned(__cvta_generic_to_shared(s_out)) + laneBase;
	// the lane's share is QT or QT - 1 steps: the last step is the only one that can fall outside it
	const int lastMine = (cnt == unsigned(QT)) ? 1 : 0;

	// one block.  kSteady: every step is Add + Sub at full weight and all operands exist
	auto block = [&](unsigned tb, unsigned landed, auto steadyTag)   // pixels [0, landed) are in the ring
	{
		constexpr bool kSteady = decltype(steadyTag)::value;
		const unsigned t0 = tb + s0;

		// even kernels: the remainder the first step subtracts belongs to the output before this lane's share
		int carry = 0;
		if (SUBEDGES)
			carry = __shfl_sync(0xffffffffu, (part == P-1) ? lastOut2 : lastOut1, srcCarry);

		// 1. fold this lane's steps: operands (deprecated/boxblur.cpp:17-36, the four saturating operations of Add + Sub folded
		//    into d = a - b, h = 65535 - b as in the staged kernel) and the (D, L, H) triple of their composition.
		//    A step past the lane's share reads as zero pixels: d = 0, h = 65535, the identity.
		const unsigned inAddr = inBase + ((t0 - 1) & (kRing-1))*kStep; // +0: pixel t0 - 1, +(1 + q)*kStep: pixel t0 + q (the ring's mirror: no wrap)
		int pxs[QT + 1];
		#pragma unroll
		for (int q = 0; q <= QT; ++q)
		{
			unsigned v = 0;
			// ramp blocks touch pixels that do not exist (t0 - 1 = -1, positions past the line) or whose stage is still in
			// flight (past the end of a cut-off block): their values are masked below, the read itself is skipped
			if (kSteady || t0 + q - 1 < landed)
				asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(inAddr + q*kStep) : "memory");
			pxs[q] = int(v);
		}

		int d[QT], bq[QT];
		int prevPx = 0, prevS = 0;
		if (SUBEDGES)
		{
			prevPx = (kSteady || (t0 >= 1 && t0 <= len)) ? pxs[0] : 0;
			prevS = (kSteady || t0 >= span + 1) ? carry : 0;
		}
		#pragma unroll
		for (int q = 0; q < QT; ++q)
		{
			const unsigned t = t0 + q;
			const int m = (q == QT-1) ? lastMine : 1;
			int px = pxs[1 + q]*m, spx = outPrev[q]*m, a;
			if (!kSteady)
			{
				if (t >= len) px = 0;
				if (t < span || t >= total) spx = 0;
			}
			// odd kernels shift the remainder out entirely (bytes >> 8 == 0): a = px, b = spx
			a = SUBEDGES ? (prevPx >> sh) + (px - (px >> sh)) : px;
			if (!kSteady && t > len) a = 0;                  // t == len: the pending remainder, deprecated/boxblur.cpp:113-115
			int b = SUBEDGES ? (prevS >> sh) + (spx - (spx >> sh)) : spx;
			if (SUBEDGES && q == QT-1) { a *= m; b *= m; }   // the pending remainders belong to the next lane's first step
			prevPx = px; prevS = spx;
			d[q] = a - b;
			bq[q] = b;
		}

		int o[QT];
		bool done = false;
		if (kSteady && fastLimit != 0)
		{
			// Fast path.  While no clamp of the recurrence is active the accumulator is a plain running sum: the parts exchange
			// one integer (a prefix sum of their shares' totals) instead of a (D, L, H) triple and replay with additions.  The
			// replayed values prove the premise: every step stayed within [0, fastLimit], fastLimit <= 65535 - max(b) (neither
			// the saturating add nor the saturating subtract of deprecated/boxblur.cpp:17-36 could have clamped, by induction
			// over the steps) and <= the largest accumulator whose quotient still is <= 255 (packuswb would not clamp either).
			// One vote per warp; a warp with any step out of range redoes the block on the exact path below.
			int Ds = 0;
			#pragma unroll
			for (int q = 0; q < QT; ++q) Ds += d[q];
			#pragma unroll
			for (unsigned round = 0; round < unsigned(LOG2P); ++round)
			{
				const int up = __shfl_up_sync(0xffffffffu, Ds, C << round);
				if (part >= (1u << round)) Ds += up;
			}
			const int endF = x0 + Ds;
			int accF = __shfl_up_sync(0xffffffffu, endF, C);
			if (part == 0) accF = x0;
			const int x0F = __shfl_sync(0xffffffffu, endF, srcLast);
			unsigned worst = 0;
			#pragma unroll
			for (int q = 0; q < QT; ++q)
			{
				accF += d[q];
				worst = max(worst, unsigned(accF));              // a negative accumulator reads as a huge unsigned one
				o[q] = int(__umulhi(unsigned(accF), fullDivHi));
			}
			done = __all_sync(0xffffffffu, worst <= fastLimit);
			if (done) x0 = x0F;
		}

		if (!done)
		{
			// Exact path: the (D, L, H) triple of the composition of this lane's clamped steps, h = 65535 - b.
			// A step past the lane's share reads as zero pixels: d = 0, h = 65535, the identity.
			int D = 0, L = 0, H = 65535;
			#pragma unroll
			for (int q = 0; q < QT; ++q)
			{
				const int h = 65535 - bq[q];
				D += d[q];
				L = __viaddmin_s32_relu(L, d[q], h);
				H = __viaddmin_s32_relu(H, d[q], h);
			}

			// 2. inclusive scan over the parts: (D, L, H) becomes the composition of parts 0..part
			#pragma unroll
			for (unsigned round = 0; round < unsigned(LOG2P); ++round)
			{
				const unsigned delta = C << round;
				const int De = __shfl_up_sync(0xffffffffu, D, delta), Le = __shfl_up_sync(0xffffffffu, L, delta), He = __shfl_up_sync(0xffffffffu, H, delta);
				if (part >= (1u << round))
				{
					// x -> clamp(clamp(x + De, Le, He) + D, L, H) == clamp(x + De + D, clamp(Le + D, L, H), clamp(He + D, L, H))
					const int Ln = max(__viaddmin_s32(Le, D, H), L), Hn = max(__viaddmin_s32(He, D, H), L);
					D += De; L = Ln; H = Hn;
				}
			}
			const int accEnd = max(__viaddmin_s32(x0, D, H), L);           // accumulator after this lane's share
			int acc = __shfl_up_sync(0xffffffffu, accEnd, C);               // ... and before it
			if (part == 0) acc = x0;
			x0 = __shfl_sync(0xffffffffu, accEnd, srcLast);                 // the next block starts where the last part ends

			// 3. replay with the real accumulator, emit outputs (Div, deprecated/boxblur.cpp:39-42).  A step past the lane's share
			//    computes garbage that nothing reads: its slot of outPrev is masked when it is used.
			#pragma unroll
			for (int q = 0; q < QT; ++q)
			{
				const unsigned t = t0 + q;
				acc = __viaddmin_s32_relu(acc, d[q], 65535 - bq[q]);
				if (kSteady)
					o[q] = int(min(__umulhi(unsigned(acc), fullDivHi), 255u)); // the full-weight divisor is <= 32768: see the staged kernel
				else
				{
					const unsigned oi = (t >= edgeSpan && t < total) ? t - edgeSpan : 0;
					const unsigned div = (oi < kM) ? s_divTab[oi] : (oi < mainEnd) ? fullDiv : s_divTab[len - 1 - oi];
					o[q] = int(old_div(unsigned(acc), div));
				}
			}
		}
		#pragma unroll
		for (int q = 0; q < QT; ++q) outPrev[q] = o[q];
		if (SUBEDGES)
		{
			lastOut2 = lastOut1;
			lastOut1 = lastMine ? o[QT-1] : o[(QT >= 2) ? QT-2 : 0];
		}

		// stores: a step outside the share (or outside the line, at the ramps) is not written
		const unsigned outPos = (t0 - edgeSpan) & (kRing-1);
		if (kSteady && outPos + QT <= kRing)
		{
			const unsigned addr0 = outBase + outPos*kStep;
			#pragma unroll
			for (int q = 0; q < QT; ++q)
				if (q < QT-1 || lastMine)
					asm volatile("st.shared.u8 [%0], %1;" :: "r"(addr0 + q*kStep), "r"(o[q]) : "memory");
		}
		else
		{
			#pragma unroll
			for (int q = 0; q < QT; ++q)
			{
				const unsigned t = t0 + q;
				bool emit = (q < QT-1) || lastMine;
				if (!kSteady) emit = emit && t >= edgeSpan && t < total;
				if (emit)
					asm volatile("st.shared.u8 [%0], %1;" :: "r"(outBase + ((outPos + q) & (kRing-1))*kStep), "r"(o[q]) : "memory");
			}
		}
	};

	unsigned readyEnd = 0;                           // pixels [0, readyEnd) have landed in the ring
	unsigned flushNext = kStage + edgeSpan;          // a 64-output chunk is complete once the steps reach this
	auto advance = [&](unsigned te)                  // pixels up to the end of the block must have landed
	{
		while (readyEnd < te)
		{
			loadStage(readyEnd + kPrefetch*kStage);
			cp_async_commit();
			cp_async_wait<kPrefetch>();
			__syncthreads();
			readyEnd += kStage;
		}
	};
	auto retire = [&](unsigned te)                   // finished 64-output chunks go out with 16-byte stores
	{
		if (te >= flushNext)
		{
			__syncthreads();
			do
			{
				flushStage(flushNext - kStage - edgeSpan);
				flushNext += kStage;
			}
			while (te >= flushNext);
		}
	};

	// ramp up, steady state, ramp down
	unsigned tb = 0;
	for (; tb < span + 1 && tb < total; tb += kM)
	{
		const unsigned te = min(tb + kM, total);
		advance(min(te, len));
		block(tb, min(te, len), BoolTag<false>());
		retire(te);
	}
	// two blocks per iteration: the outputs of one block are the trailing-edge operands of the next, and with the pair spelled
	// out they stay where they were computed instead of being copied at the loop's back edge
	for (; tb + 2*kM <= len; tb += 2*kM)
	{
		advance(tb + kM);
		block(tb, tb + kM, BoolTag<true>());
		retire(tb + kM);
		advance(tb + 2*kM);
		block(tb + kM, tb + 2*kM, BoolTag<true>());
		retire(tb + 2*kM);
	}
	for (; tb + kM <= len; tb += kM)
	{
		advance(tb + kM);
		block(tb, tb + kM, BoolTag<true>());
		retire(tb + kM);
	}
	for (; tb < total; tb += kM)
	{
		const unsigned te = min(tb + kM, total);
		advance(min(te, len));
		block(tb, min(te, len), BoolTag<false>());
		retire(te);
	}

	cp_async_wait<0>();
	__syncthreads();
	for (unsigned p0 = flushNext - kStage - edgeSpan; p0 < len; p0 += kStage)
		flushStage(p0);
}

template <bool VERT, bool SUBEDGES>
static cudaError_t LaunchBlocked(ckd_ctx *ctx, uint8_t *p, unsigned numLines, unsigned len, unsigned pitch, const OldBlurSetup &s)
{
	// Shape of a block: P = 1 << log2P lanes per line-channel, QT = ceil(kM/P) steps each.  Measured on B200 at 4K the fewest
	// parts that keep QT <= 8 win: every extra scan round costs more than the warps it adds hide (two parts up to a median of
	// 13, four up to 32, ...).
	const unsigned blocks = ckd_div_up(numLines, 8), kM = s.kernelMedian;
	unsigned log2P = (kM <= 13) ? 1 : (kM <= 32) ? 2 : (kM <= 64) ? 3 : 4;
	static const int forcedLog2P = getenv("CKD_BLUR_LOG2P") ? atoi(getenv("CKD_BLUR_LOG2P")) : 0; // tuning sweeps: 1..4 where the shape allows it
	if (forcedLog2P >= 1 && forcedLog2P <= 4)
	{
		const unsigned q = (kM + (1u << forcedLog2P) - 1) >> forcedLog2P;
		if (q >= 2 && q <= 8) log2P = unsigned(forcedLog2P);
	}
	const int qt = int((kM + (1u << log2P) - 1) >> log2P); // every lane then owns qt or qt - 1 steps

	const size_t smem = 2*kRingBytes;
	#define CKD_BLOCKED(LOG2P, QT) old_blur_blocked_kernel<VERT, SUBEDGES, LOG2P, QT><<<blocks, 32 << LOG2P, smem, ctx->stream>>>(p, numLines, len, pitch, s)
	#define CKD_BLOCKED_QT(LOG2P) switch (qt) { case 2: CKD_BLOCKED(LOG2P, 2); break; case 3: CKD_BLOCKED(LOG2P, 3); break; case 4: CKD_BLOCKED(LOG2P, 4); break; \
		case 5: CKD_BLOCKED(LOG2P, 5); break; case 6: CKD_BLOCKED(LOG2P, 6); break; case 7: CKD_BLOCKED(LOG2P, 7); break; default: CKD_BLOCKED(LOG2P, 8); break; }
	switch (log2P)
	{
	case 1: CKD_BLOCKED_QT(1); break;
	case 2: CKD_BLOCKED_QT(2); break;
	case 3: CKD_BLOCKED_QT(3); break;
	default: CKD_BLOCKED_QT(4); break;
	}
	#undef CKD_BLOCKED_QT
	#undef CKD_BLOCKED
	return cudaSuccess;
}

template <bool VERT, bool INPLACE, bool SUBEDGES, int KM>
static cudaError_t LaunchStaged(ckd_ctx *ctx, uint8_t *pDest, const uint8_t *pSrc, unsigned numLines, unsigned len, unsigned pitch, const OldBlurSetup &s)
{
	const unsigned blocks = ckd_div_up(numLines, 8*kBlurWarps);
	const size_t smem = size_t(kBlurWarps)*2*kRingBytes;
	bool &attrSet = ctx->blurAttrSet[(((VERT ? 1 : 0)*2 + (INPLACE ? 1 : 0))*2 + (SUBEDGES ? 1 : 0))*9 + KM];
	if (!attrSet)
	{
		const cudaError_t err = cudaFuncSetAttribute(old_blur_staged_kernel<VERT, INPLACE, SUBEDGES, KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
		if (cudaSuccess != err)
			return err;
		attrSet = true;
	}
	old_blur_staged_kernel<VERT, INPLACE, SUBEDGES, KM><<<blocks, kBlurWarps*32, smem, ctx->stream>>>(pDest, pSrc, numLines, len, pitch, s);
	return cudaSuccess;
}

template <bool VERT, bool SUBEDGES>
static cudaError_t LaunchStagedInPlace(ckd_ctx *ctx, uint8_t *p, unsigned numLines, unsigned len, unsigned pitch, const OldBlurSetup &s)
{
	switch (s.kernelMedian)
	{
	case 1: return LaunchStaged<VERT, true, SUBEDGES, 1>(ctx, p, p, numLines, len, pitch, s);
	case 2: return LaunchStaged<VERT, true, SUBEDGES, 2>(ctx, p, p, numLines, len, pitch, s);
	case 3: return LaunchStaged<VERT, true, SUBEDGES, 3>(ctx, p, p, numLines, len, pitch, s);
	case 4: return LaunchStaged<VERT, true, SUBEDGES, 4>(ctx, p, p, numLines, len, pitch, s);
	case 5: return LaunchStaged<VERT, true, SUBEDGES, 5>(ctx, p, p, numLines, len, pitch, s);
	case 6: return LaunchStaged<VERT, true, SUBEDGES, 6>(ctx, p, p, numLines, len, pitch, s);
	case 7: return LaunchStaged<VERT, true, SUBEDGES, 7>(ctx, p, p, numLines, len, pitch, s);
	case 8: return LaunchStaged<VERT, true, SUBEDGES, 8>(ctx, p, p, numLines, len, pitch, s);
	default: return LaunchStaged<VERT, true, SUBEDGES, 0>(ctx, p, p, numLines, len, pitch, s);
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
	static const bool noFast = nullptr != getenv("CKD_BLUR_NOFAST");
	s.noFastPath = noFast ? 1 : 0;

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
		cudaError_t err;
		const int variant = (vert ? 4 : 0) | (inPlace ? 2 : 0) | (s.subEdges ? 1 : 0);
		// The serial walk runs one warp per 8 lines; once that alone puts three warps on every SM (vertical passes at 4K) it holds
		// its own against the scan up to medium kernels.  Measured on B200, see profiles/r01_notes.md.
		// medians up to 8 keep their trailing edge in registers (staged walk); an odd kernel of median 9 is still better off on the
		// staged walk than on a two-part scan (profiles/r02_blur_dispatch_sweep.txt)
		unsigned blockedFrom = (numLines/8 >= 3u*unsigned(ctx->numSMs)) ? 24 : (s.subEdges ? 9 : 10);
		static const int forcedFrom = getenv("CKD_BLUR_BLOCKED_FROM") ? atoi(getenv("CKD_BLUR_BLOCKED_FROM")) : 0; // tuning sweeps
		if (forcedFrom >= 9) blockedFrom = unsigned(forcedFrom);
		if (inPlace && s.kernelMedian >= blockedFrom)
		{
			if (vert) err = s.subEdges ? LaunchBlocked<true, true>(ctx, pDest, numLines, lineLen, pitch, s) : LaunchBlocked<true, false>(ctx, pDest, numLines, lineLen, pitch, s);
			else err = s.subEdges ? LaunchBlocked<false, true>(ctx, pDest, numLines, lineLen, pitch, s) : LaunchBlocked<false, false>(ctx, pDest, numLines, lineLen, pitch, s);
		}
		else
		{
			switch (variant)
			{
			case 0: err = LaunchStaged<false, false, false, 0>(ctx, pDest, pSrc, numLines, lineLen, pitch, s); break;
			case 1: err = LaunchStaged<false, false, true, 0>(ctx, pDest, pSrc, numLines, lineLen, pitch, s); break;
			case 2: err = LaunchStagedInPlace<false, false>(ctx, pDest, numLines, lineLen, pitch, s); break;
			case 3: err = LaunchStagedInPlace<false, true>(ctx, pDest, numLines, lineLen, pitch, s); break;
			case 4: err = LaunchStaged<true, false, false, 0>(ctx, pDest, pSrc, numLines, lineLen, pitch, s); break;
			case 5: err = LaunchStaged<true, false, true, 0>(ctx, pDest, pSrc, numLines, lineLen, pitch, s); break;
			case 6: err = LaunchStagedInPlace<true, false>(ctx, pDest, numLines, lineLen, pitch, s); break;
			default: err = LaunchStagedInPlace<true, true>(ctx, pDest, numLines, lineLen, pitch, s); break;
			}
		}
		CKD_CUDA(err);
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

// The sums are plain 32-bit integers without saturation, so the running window of a line is a difference of prefix sums:
// with E(m) = px[m] + (((px[m+1] - px[m])*alpha16) >> 16) (iAdd / iSub, boxblur.cpp:37-61) and P(n) = E(0) + ... + E(n-1), the
// accumulator in front of output k is
//     iSum0 + P(iSpan + 1 + min(k, len - iSpan)) - P(iSpan + 1) - P(max(k - iSpan, 0)),
// iSum0 = px[0] + ... + px[iSpan-1] + ((px[iSpan]*alpha16) >> 16) (boxblur.cpp:150-157).  One CTA owns one line: the line is
// staged in shared memory by one TMA bulk copy (cp.async.bulk + mbarrier), every thread folds a chunk of consecutive E into a local prefix, a block
// scan offsets the chunks, and the outputs -- two 16-byte prefix reads, the 10:22 scale, the pack -- go out coalesced.
// Lines are contiguous (vertical passes run on a transposed copy, like the reference's own Transpose32 round trip,
// boxblur.cpp:215-299), so the two elements the reference reads past the end of a line (boxblur.cpp:171,185) are simply the
// next two elements of the buffer; past the end of the buffer they read as 0.  Exact: integer sums in any order.
constexpr int kNewBlurThreads = 256;

__global__ void __launch_bounds__(kNewBlurThreads) new_blur_line_kernel(uint32_t *__restrict__ pDest, const uint32_t *__restrict__ pSrc, unsigned numLines, unsigned len, NewBlurSetup s)
{
	extern __shared__ __align__(16) uint8_t s_mem[];
	uint32_t *s_px = reinterpret_cast<uint32_t *>(s_mem);                                   // px[0..len+1] (+2 of slack), contiguous: TMA destination
	int4 *s_P = reinterpret_cast<int4 *>(s_mem + ((size_t(len) + 4)*4 + 15 & ~size_t(15))); // P(0..len+1)
	__shared__ int4 s_warpTotals[kNewBlurThreads/32];
	__shared__ int s_first[4];                                                              // px[0] + ... + px[iSpan-1] per channel
	__shared__ __align__(8) uint64_t s_bar;

	const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const unsigned line = blockIdx.x;
	const size_t total = size_t(numLines)*len, base = size_t(line)*len;

	// the line and the two elements behind it: one TMA bulk copy when the run is 16-byte aligned and inside the buffer
	const unsigned runElems = (len + 2 + 3) & ~3u;
	const bool tma = 0 == ((reinterpret_cast<uintptr_t>(pSrc + base)) & 15) && base + runElems <= total;
	if (tid < 4) s_first[tid] = 0;
	if (tma)
	{
		if (0 == tid)
		{
			mbar_init(&s_bar, 1);
			mbar_fence_init();
		}
		__syncthreads();
		if (0 == tid)
		{
			mbar_arrive_expect_tx(&s_bar, runElems*4);
			bulk_load(s_px, pSrc + base, runElems*4, &s_bar);
		}
		mbar_wait(&s_bar, 0);
	}
	else
	{
		for (unsigned i = tid; i < len + 2; i += kNewBlurThreads)
			s_px[i] = (base + i < total) ? pSrc[base + i] : 0u;
		__syncthreads();
	}

	// chunk of consecutive elements per thread: E(m) for m in [m0, m1), local exclusive prefix into s_P.  The chunk length is
	// odd: neighbouring threads then start an odd number of words (px) / of 16-byte entries (P) apart: no bank conflicts
	const unsigned perThread = ((len + 1 + kNewBlurThreads - 1)/kNewBlurThreads) | 1u; // E(0..len) -> P(0..len+1)
	const unsigned m0 = min(tid*perThread, len + 1), m1 = min(m0 + perThread, len + 1);
	int4 run = make_int4(0, 0, 0, 0), first = make_int4(0, 0, 0, 0);
	uint32_t A = (m0 < m1) ? s_px[m0] : 0u;
	for (unsigned m = m0; m < m1; ++m)
	{
		const uint32_t B = s_px[m + 1];
		s_P[m] = run;
		const int a0 = int(A & 0xff), a1 = int((A >> 8) & 0xff), a2 = int((A >> 16) & 0xff), a3 = int(A >> 24);
		const int b0 = int(B & 0xff), b1 = int((B >> 8) & 0xff), b2 = int((B >> 16) & 0xff), b3 = int(B >> 24);
		run.x += a0 + (((b0 - a0)*s.iAlpha) >> 16);
		run.y += a1 + (((b1 - a1)*s.iAlpha) >> 16);
		run.z += a2 + (((b2 - a2)*s.iAlpha) >> 16);
		run.w += a3 + (((b3 - a3)*s.iAlpha) >> 16);
		if (m < s.iSpan) { first.x += a0; first.y += a1; first.z += a2; first.w += a3; }
		A = B;
	}
	if (m0 < s.iSpan && m0 < m1)
	{
		atomicAdd(&s_first[0], first.x); atomicAdd(&s_first[1], first.y); atomicAdd(&s_first[2], first.z); atomicAdd(&s_first[3], first.w);
	}

	// block exclusive scan of the chunk totals
	int4 incl = run;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const int x = __shfl_up_sync(0xffffffffu, incl.x, d), y = __shfl_up_sync(0xffffffffu, incl.y, d), z = __shfl_up_sync(0xffffffffu, incl.z, d), w = __shfl_up_sync(0xffffffffu, incl.w, d);
		if (lane >= unsigned(d)) { incl.x += x; incl.y += y; incl.z += z; incl.w += w; }
	}
	if (lane == 31) s_warpTotals[warp] = incl;
	__syncthreads();
	int4 offset = make_int4(incl.x - run.x, incl.y - run.y, incl.z - run.z, incl.w - run.w);
	for (unsigned w = 0; w < warp; ++w)
	{
		const int4 t = s_warpTotals[w];
		offset.x += t.x; offset.y += t.y; offset.z += t.z; offset.w += t.w;
	}
	for (unsigned m = m0; m < m1; ++m)
	{
		int4 v = s_P[m];
		v.x += offset.x; v.y += offset.y; v.z += offset.z; v.w += offset.w;
		s_P[m] = v;
	}
	if (m1 == len + 1 && m0 < m1) // the thread that owns the last E also provides P(len + 1)
		s_P[len + 1] = make_int4(offset.x + run.x, offset.y + run.y, offset.z + run.z, offset.w + run.w);
	__syncthreads();

	// outputs
	const uint32_t pivot = s_px[s.iSpan];
	int4 sum0 = s_P[s.iSpan + 1]; // subtracted below
	sum0.x = s_first[0] + ((int(pivot & 0xff)*s.iAlpha) >> 16) - sum0.x;
	sum0.y = s_first[1] + ((int((pivot >> 8) & 0xff)*s.iAlpha) >> 16) - sum0.y;
	sum0.z = s_first[2] + ((int((pivot >> 16) & 0xff)*s.iAlpha) >> 16) - sum0.z;
	sum0.w = s_first[3] + ((int(pivot >> 24)*s.iAlpha) >> 16) - sum0.w;

	for (unsigned k = tid; k < len; k += kNewBlurThreads)
	{
		const int4 hi = s_P[s.iSpan + 1 + min(k, len - s.iSpan)];
		const int4 lo = s_P[(k > s.iSpan) ? k - s.iSpan : 0];
		unsigned scale = s.iScale;
		if (k < s.iSpan) scale = scale_fp22(s.halfScale + float(k)*s.dScale);                        // boxblur.cpp:160-174
		else if (k >= len - s.iSpan) scale = scale_fp22(s.halfScale + float(len - 1 - k)*s.dScale);  // boxblur.cpp:196-207
		const uint32_t out = new_div_pack(sum0.x + hi.x - lo.x, scale) | (new_div_pack(sum0.y + hi.y - lo.y, scale) << 8)
			| (new_div_pack(sum0.z + hi.z - lo.z, scale) << 16) | (new_div_pack(sum0.w + hi.w - lo.w, scale) << 24);
		pDest[base + k] = out;
	}
}

// Transpose32, boxblur.cpp:215-268: pDest[x*yRes + y] = pSrc[y*xRes + x]
__global__ void __launch_bounds__(256) transpose32_kernel(uint32_t *__restrict__ pDest, const uint32_t *__restrict__ pSrc, unsigned xRes, unsigned yRes)
{
	__shared__ uint32_t s_tile[32][33];
	const unsigned x0 = blockIdx.x*32, y0 = blockIdx.y*32;
	for (unsigned j = threadIdx.y; j < 32; j += 8)
	{
		const unsigned x = x0 + threadIdx.x, y = y0 + j;
		if (x < xRes && y < yRes) s_tile[j][threadIdx.x] = pSrc[size_t(y)*xRes + x];
	}
	__syncthreads();
	for (unsigned j = threadIdx.y; j < 32; j += 8)
	{
		const unsigned y = y0 + threadIdx.x, x = x0 + j;
		if (x < xRes && y < yRes) pDest[size_t(x)*yRes + y] = s_tile[threadIdx.x][j];
	}
}

static int Transpose32(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned xRes, unsigned yRes)
{
	ckd_prof_begin(ctx, "transpose32", 8.0*xRes*yRes);
	transpose32_kernel<<<dim3(ckd_div_up(xRes, 32), ckd_div_up(yRes, 32)), dim3(32, 8), 0, ctx->stream>>>(d_dest, d_src, xRes, yRes);
	CKD_CHECK_LAUNCH(ctx);
	return CKD_OK;
}

// HorzBlur32, boxblur.cpp:87-212, numPasses passes over contiguous lines, ping-pong between two buffers.
// d_a receives pass 0, d_b pass 1, ...; returns the buffer the last pass wrote.
static int NewBlurPasses(ckd_ctx *ctx, uint32_t *d_a, uint32_t *d_b, const uint32_t *d_src, unsigned numLines, unsigned lineLen,
	float strength, float gain, unsigned numPasses, const char *what, uint32_t **d_result)
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

	const size_t smem = ((size_t(lineLen) + 4)*4 + 15 & ~size_t(15)) + (size_t(lineLen) + 2)*sizeof(int4);
	CKD_REQUIRE(smem <= 200*1024, "line too long for the blur's shared-memory staging");
	if (!ctx->newBlurAttrSet)
	{
		CKD_CUDA(cudaFuncSetAttribute(new_blur_line_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024));
		ctx->newBlurAttrSet = true;
	}

	const uint32_t *pRead = d_src;
	uint32_t *pWrite = d_a, *pOther = d_b;
	for (unsigned iPass = 0; iPass < numPasses; ++iPass)
	{
		ckd_prof_begin(ctx, what, 8.0*numLines*lineLen);
		new_blur_line_kernel<<<numLines, kNewBlurThreads, smem, ctx->stream>>>(pWrite, pRead, numLines, lineLen, s);
		CKD_CHECK_LAUNCH(ctx);
		pRead = pWrite;
		std::swap(pWrite, pOther);
	}
	*d_result = pOther; // after the swap: the buffer the last pass wrote
	return CKD_OK;
}

extern "C" int ckd_new_blur_h(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength, float gain, unsigned num_passes)
{
	CKD_REQUIRE(ctx && d_dest && d_src, "null argument");
	CKD_REQUIRE(size_t(x_res)*y_res <= size_t(ctx->resX)*ctx->resY, "image larger than the context's scratch buffers");
	// BoxBlur_Horz32, boxblur.cpp:270-279.  The last pass lands in d_dest; a pass never writes the buffer it reads (lines read
	// two elements into their successor), so an in-place call goes through both scratch images.
	uint32_t *d_result = nullptr;
	if (d_dest == d_src)
	{
		CKD_TRY(NewBlurPasses(ctx, ctx->d_scratch[0], ctx->d_scratch[1], d_src, y_res, x_res, strength, gain, num_passes, "new_blur_h", &d_result));
		return ckd_copy(ctx, d_dest, d_result, size_t(x_res)*y_res*4);
	}
	uint32_t *d_first = (num_passes & 1) ? d_dest : ctx->d_scratch[0], *d_second = (num_passes & 1) ? ctx->d_scratch[0] : d_dest;
	return NewBlurPasses(ctx, d_first, d_second, d_src, y_res, x_res, strength, gain, num_passes, "new_blur_h", &d_result);
}

extern "C" int ckd_new_blur_v(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength, float gain, unsigned num_passes)
{
	CKD_REQUIRE(ctx && d_dest && d_src, "null argument");
	CKD_REQUIRE(size_t(x_res)*y_res <= size_t(ctx->resX)*ctx->resY, "image larger than the context's scratch buffers");
	// BoxBlur_Vert32, boxblur.cpp:281-299: transpose, blur the columns as lines (length yRes, the radius derives from yRes), transpose back
	uint32_t *d_result = nullptr;
	CKD_TRY(Transpose32(ctx, ctx->d_scratch[1], d_src, x_res, y_res));
	CKD_TRY(NewBlurPasses(ctx, ctx->d_scratch[0], ctx->d_scratch[1], ctx->d_scratch[1], x_res, y_res, strength, gain, num_passes, "new_blur_v", &d_result));
	return Transpose32(ctx, d_dest, d_result, y_res, x_res);
}

extern "C" int ckd_new_blur(ckd_ctx *ctx, uint32_t *d_dest, const uint32_t *d_src, unsigned x_res, unsigned y_res, float strength, float gain, unsigned num_passes)
{
	CKD_REQUIRE(ctx && d_dest && d_src, "null argument");
	CKD_REQUIRE(size_t(x_res)*y_res <= size_t(ctx->resX)*ctx->resY, "image larger than the context's scratch buffers");
	// BoxBlur_32, boxblur.cpp:301-318: rows, then columns
	uint32_t *d_rows = nullptr, *d_cols = nullptr;
	CKD_TRY(NewBlurPasses(ctx, ctx->d_scratch[0], ctx->d_scratch[1], d_src, y_res, x_res, strength, gain, num_passes, "new_blur_h", &d_rows));
	uint32_t *d_t = (d_rows == ctx->d_scratch[0]) ? ctx->d_scratch[1] : ctx->d_scratch[0];
	CKD_TRY(Transpose32(ctx, d_t, d_rows, x_res, y_res));
	CKD_TRY(NewBlurPasses(ctx, d_rows, d_t, d_t, x_res, y_res, strength, gain, num_passes, "new_blur_v", &d_cols));
	return Transpose32(ctx, d_dest, d_cols, y_res, x_res);
}

// ckd_gather.cu -- frame gather to one GPU over peer memory (the north star's "P2P gather to rank 0").
//
// What it replaces: in the reference every finished frame goes from the one renderer to Display::Update
// (display.cpp:66-82).  With frames sharded over the GPUs of a box (frame i -> rank i mod N, one process per GPU) the
// renderers' frames have to meet again on one GPU before they go to the host / the sink.  This is that step, and it is the
// only exchange of the system: a ring of frame slots in the collecting GPU's HBM, written by the producers with peer copies
// over NVLink (CUDA IPC mapping of the ring, copy engines, no SM time) and flag-signalled on the device.  No NCCL, no host
// round trip per frame:
//
//   producer (any rank, its own process)          ring in the collector's HBM            collector (rank 0)
//   ------------------------------------          ---------------------------            ------------------
//   render frame q into a local staging frame
//   push stream:  wait   drained[s] >= q+1-S      slot s = q mod S                       consumer stream:  wait ready[s] == q+1
//                 copy   staging -> slot s   ---> [frame bytes]                                            checksum / copy to host
//                 signal ready[s] = q+1      ---> ready[s]                                                 signal drained[s] = q+1
//
// q is a global sequence number (frames are pushed and popped in the order 0, 1, 2, ...; who renders which q is the caller's
// business).  A slot is reused only after the collector has drained it, so a fast producer blocks on the device, never on
// the host.  The waits are one-thread kernels polling a 64-bit flag with acquire semantics; they give up after
// `timeout_ms` and latch an error in the ring's status word instead of hanging the GPU.
//
// Streams: the copy of frame q overlaps the rendering of frame q+1 (K local staging frames per producer, the compute stream
// only waits for the copy that last read the staging frame it is about to overwrite).

#include "ckd_internal.h"

#include <string.h>
#include <unistd.h>

namespace {

constexpr unsigned kGatherMagic = 0x47444b43u; // "CKDG"
constexpr int kMaxSlots = 64;
constexpr int kMaxStaging = 8;                   // up to four frames rendering side by side (lanes) while as many are being copied
constexpr size_t kControlBytes = 4096;         // ready[64], drained[64], status

// control block at the start of the ring allocation (lives in the collector's HBM, mapped by every producer)
struct GatherControl
{
	unsigned long long ready[kMaxSlots];        // q+1 of the frame that is complete in the slot
	unsigned long long drained[kMaxSlots];      // q+1 of the last frame the collector has finished reading from the slot
	unsigned long long status;                  // 0 = fine; else (q+1) | kind<<56 of the first wait that timed out
};

struct ExportedHandle                           // what ckd_gather_export writes (CKD_GATHER_HANDLE_BYTES)
{
	unsigned magic;
	int slots;
	int device;
	int pid;
	unsigned long long frameBytes;
	unsigned long long slotStride;
	cudaIpcMemHandle_t mem;
};
static_assert(sizeof(ExportedHandle) <= 128, "handle does not fit CKD_GATHER_HANDLE_BYTES");

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
	asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}

// polls *flag until it is >= want (the flags only grow); false on time-out
__device__ bool wait_flag(const unsigned long long *flag, unsigned long long want, unsigned long long timeoutNs)
{
	if (ld_acquire_sys(flag) >= want)
		return true;
	const unsigned long long t0 = global_timer_ns();
	unsigned backoff = 64;
	for (;;)
	{
		if (ld_acquire_sys(flag) >= want)
			return true;
		if (global_timer_ns() - t0 > timeoutNs)
			return false;
		__nanosleep(backoff);
		if (backoff < 2048) backoff <<= 1;
	}
}

// producer side: the slot must have been drained of frame q-S before frame q may be copied into it
__global__ void gather_wait_drained_kernel(GatherControl *ctl, int slot, unsigned long long want, unsigned long long timeoutNs)
{
	if (!wait_flag(&ctl->drained[slot], want, timeoutNs))
		atomicCAS(&ctl->status, 0ull, want | (1ull << 56));
}

// producer side, after the copy (stream order): publish the frame
__global__ void gather_signal_ready_kernel(GatherControl *ctl, int slot, unsigned long long seqPlus1)
{
	__threadfence_system();
	st_release_sys(&ctl->ready[slot], seqPlus1);
}

// collector side: wait for frame q in its slot (used in front of a copy to the host)
__global__ void gather_wait_ready_kernel(GatherControl *ctl, int slot, unsigned long long want, unsigned long long timeoutNs)
{
	if (!wait_flag(&ctl->ready[slot], want, timeoutNs))
		atomicCAS(&ctl->status, 0ull, want | (2ull << 56));
}

__global__ void gather_signal_drained_kernel(GatherControl *ctl, int slot, unsigned long long seqPlus1)
{
	__threadfence_system();
	st_release_sys(&ctl->drained[slot], seqPlus1);
}

// Collector side, checksum mode: folds the pixels into sum_i pixel[i]*(2i+1) mod 2^64 -- position dependent, order
// independent, so the value does not depend on the launch shape -- and the last CTA to finish publishes the sum and (release)
// hands the slot back.  It runs behind the one-thread gather_wait_ready_kernel in stream order and never waits itself: a full
// grid of CTAs parked on a flag would hold registers and warp slots on every SM that the render kernels of the frame it is
// waiting for may need (a 1024-thread CTA does not fit next to them), which is a deadlock until the time-out.
__global__ void __launch_bounds__(256) gather_checksum_kernel(GatherControl *ctl, int slot, unsigned long long seqPlus1,
	const uint4 *__restrict__ frame, size_t numQuads, unsigned long long *sumOut, unsigned long long *scratch /* [0] partial, [1] CTAs done */, int release)
{
	__shared__ unsigned long long warpSums[8];
	unsigned long long sum = 0;
	const size_t stride = size_t(gridDim.x)*blockDim.x;
	for (size_t q = size_t(blockIdx.x)*blockDim.x + threadIdx.x; q < numQuads; q += stride)
	{
		const uint4 v = __ldcv(frame + q); // written by a peer's copy engine since the last launch: no stale L1 lines
		const unsigned long long w = 8ull*q + 1ull;
		sum += v.x*w + v.y*(w + 2) + v.z*(w + 4) + v.w*(w + 6);
	}
	for (int d = 16; d > 0; d >>= 1)
		sum += __shfl_xor_sync(0xffffffffu, sum, d);
	if (0 == (threadIdx.x & 31))
		warpSums[threadIdx.x >> 5] = sum;
	__syncthreads();
	if (0 == threadIdx.x)
	{
		unsigned long long total = 0;
		for (int i = 0; i < int(blockDim.x >> 5); ++i)
			total += warpSums[i];
		atomicAdd(&scratch[0], total);
		__threadfence();
		if (atomicAdd(&scratch[1], 1ull) == gridDim.x - 1)
		{
			*sumOut = atomicExch(&scratch[0], 0ull);
			scratch[1] = 0;
			__threadfence_system();
			if (release)
				st_release_sys(&ctl->drained[slot], seqPlus1);
		}
	}
}

} // namespace

struct ckd_gather
{
	ckd_ctx *ctx = nullptr;
	bool owner = false;                  // this process allocated the ring (the collector)
	bool mapped = false;                 // ring came from cudaIpcOpenMemHandle
	int slots = 0;
	size_t frameBytes = 0, slotStride = 0;
	uint8_t *d_ring = nullptr;           // control block + slots (collector's HBM; a peer mapping on the producers)
	GatherControl *ctl = nullptr;
	unsigned long long timeoutNs = 20ull*1000*1000*1000;

	// producer side
	cudaStream_t pushStream = nullptr;
	int numStaging = 0;
	uint32_t *d_staging[kMaxStaging] = {};
	cudaEvent_t evRendered[kMaxStaging] = {}, evPushed[kMaxStaging] = {};
	bool stagingBusy[kMaxStaging] = {};
	unsigned long long acquired = 0;     // staging frames handed out so far
	int currentStaging = -1;
	unsigned long long peerBytes = 0;    // bytes this process has pushed into another GPU's memory

	// collector side
	cudaStream_t popStream = nullptr;
	unsigned long long *d_sums = nullptr, *d_scratch = nullptr;
	size_t sumCapacity = 0;
	static constexpr int kPopEvents = 16;
	cudaEvent_t evPopped[kPopEvents] = {};
	cudaEvent_t evFlush = nullptr, evAdhoc = nullptr;
};

static uint8_t *SlotPtr(const ckd_gather *g, int slot) { return g->d_ring + kControlBytes + size_t(slot)*g->slotStride; }

static int CreateCommon(ckd_gather *g)
{
	CKD_CUDA(cudaStreamCreateWithFlags(&g->pushStream, cudaStreamNonBlocking));
	CKD_CUDA(cudaEventCreateWithFlags(&g->evFlush, cudaEventDisableTiming));
	CKD_CUDA(cudaEventCreateWithFlags(&g->evAdhoc, cudaEventDisableTiming));
	return CKD_OK;
}

extern "C" int ckd_gather_create(ckd_ctx *ctx, int slots, ckd_gather **out_gather)
{
	CKD_REQUIRE(ctx && out_gather, "null argument");
	*out_gather = nullptr;
	CKD_REQUIRE(slots >= 2 && slots <= kMaxSlots, "slots must be in [2, 64]");
	CKD_CUDA(cudaSetDevice(ctx->device));

	ckd_gather *g = new ckd_gather;
	g->ctx = ctx;
	g->owner = true;
	g->slots = slots;
	g->frameBytes = size_t(ctx->resX)*ctx->resY*sizeof(uint32_t);
	g->slotStride = (g->frameBytes + 4095)/4096*4096;
	const size_t total = kControlBytes + size_t(slots)*g->slotStride;

	int rc = CKD_OK;
	cudaError_t err = cudaMalloc(&g->d_ring, total); // a plain cudaMalloc: the allocation must be exportable with cudaIpcGetMemHandle
	if (cudaSuccess == err) err = cudaMemset(g->d_ring, 0, total);
	if (cudaSuccess == err) err = cudaStreamCreateWithFlags(&g->popStream, cudaStreamNonBlocking);
	g->sumCapacity = 8192;
	if (cudaSuccess == err) err = cudaMalloc(&g->d_sums, g->sumCapacity*sizeof(unsigned long long));
	if (cudaSuccess == err) err = cudaMemset(g->d_sums, 0, g->sumCapacity*sizeof(unsigned long long));
	if (cudaSuccess == err) err = cudaMalloc(&g->d_scratch, 2*sizeof(unsigned long long));
	if (cudaSuccess == err) err = cudaMemset(g->d_scratch, 0, 2*sizeof(unsigned long long));
	for (auto &ev : g->evPopped)
		if (cudaSuccess == err) err = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
	if (cudaSuccess != err)
		rc = ckd_cuda_fail(err, "ckd_gather_create", __FILE__, __LINE__);
	if (CKD_OK == rc)
		rc = CreateCommon(g);
	if (CKD_OK != rc)
	{
		ckd_gather_destroy(g);
		return rc;
	}
	g->ctl = reinterpret_cast<GatherControl *>(g->d_ring);
	*out_gather = g;
	return CKD_OK;
}

extern "C" int ckd_gather_export(ckd_gather *g, void *out_handle)
{
	CKD_REQUIRE(g && out_handle, "null argument");
	CKD_REQUIRE(g->owner, "only the process that created the ring can export it");
	ExportedHandle h;
	memset(&h, 0, sizeof(h));
	h.magic = kGatherMagic;
	h.slots = g->slots;
	h.device = g->ctx->device;
	h.pid = int(getpid());
	h.frameBytes = g->frameBytes;
	h.slotStride = g->slotStride;
	CKD_CUDA(cudaIpcGetMemHandle(&h.mem, g->d_ring));
	memset(out_handle, 0, CKD_GATHER_HANDLE_BYTES);
	memcpy(out_handle, &h, sizeof(h));
	return CKD_OK;
}

extern "C" int ckd_gather_open(ckd_ctx *ctx, const void *handle, ckd_gather **out_gather)
{
	CKD_REQUIRE(ctx && handle && out_gather, "null argument");
	*out_gather = nullptr;
	ExportedHandle h;
	memcpy(&h, handle, sizeof(h));
	CKD_REQUIRE(kGatherMagic == h.magic, "not a gather handle");
	CKD_REQUIRE(h.pid != int(getpid()), "the creating process uses its ckd_gather directly (a CUDA IPC handle cannot be opened where it was made)");
	CKD_REQUIRE(h.frameBytes == size_t(ctx->resX)*ctx->resY*sizeof(uint32_t), "the ring was created for another resolution");
	CKD_REQUIRE(h.slots >= 2 && h.slots <= kMaxSlots, "corrupt gather handle");
	CKD_CUDA(cudaSetDevice(ctx->device));

	ckd_gather *g = new ckd_gather;
	g->ctx = ctx;
	g->slots = h.slots;
	g->frameBytes = h.frameBytes;
	g->slotStride = h.slotStride;
	void *mappedPtr = nullptr;
	cudaError_t err = cudaIpcOpenMemHandle(&mappedPtr, h.mem, cudaIpcMemLazyEnablePeerAccess); // maps the collector's HBM (NVLink peer access when it is another GPU)
	if (cudaSuccess != err)
	{
		delete g;
		return ckd_cuda_fail(err, "cudaIpcOpenMemHandle(gather ring)", __FILE__, __LINE__);
	}
	g->d_ring = static_cast<uint8_t *>(mappedPtr);
	g->mapped = true;
	g->ctl = reinterpret_cast<GatherControl *>(g->d_ring);
	const int rc = CreateCommon(g);
	if (CKD_OK != rc)
	{
		ckd_gather_destroy(g);
		return rc;
	}
	*out_gather = g;
	return CKD_OK;
}

extern "C" void ckd_gather_destroy(ckd_gather *g)
{
	if (!g) return;
	cudaSetDevice(g->ctx->device);
	if (g->pushStream) cudaStreamSynchronize(g->pushStream);
	if (g->popStream) cudaStreamSynchronize(g->popStream);
	for (int i = 0; i < kMaxStaging; ++i)
	{
		if (g->d_staging[i]) cudaFree(g->d_staging[i]);
		if (g->evRendered[i]) cudaEventDestroy(g->evRendered[i]);
		if (g->evPushed[i]) cudaEventDestroy(g->evPushed[i]);
	}
	for (auto &ev : g->evPopped)
		if (ev) cudaEventDestroy(ev);
	if (g->evFlush) cudaEventDestroy(g->evFlush);
	if (g->evAdhoc) cudaEventDestroy(g->evAdhoc);
	if (g->pushStream) cudaStreamDestroy(g->pushStream);
	if (g->popStream) cudaStreamDestroy(g->popStream);
	if (g->d_sums) cudaFree(g->d_sums);
	if (g->d_scratch) cudaFree(g->d_scratch);
	if (g->d_ring)
	{
		if (g->mapped) cudaIpcCloseMemHandle(g->d_ring);
		else cudaFree(g->d_ring);
	}
	delete g;
}

extern "C" int ckd_gather_set_timeout_ms(ckd_gather *g, unsigned timeout_ms)
{
	CKD_REQUIRE(g, "null argument");
	g->timeoutNs = (unsigned long long)(timeout_ms ? timeout_ms : 1)*1000ull*1000ull;
	return CKD_OK;
}

// a local frame to render sequence number q into: one of numStaging device frames, handed out round robin; the compute stream
// waits (on the device) for the push that last read it
extern "C" int ckd_gather_acquire_on(ckd_gather *g, ckd_ctx *renderCtx, uint32_t **out_d_frame)
{
	CKD_REQUIRE(g && renderCtx && out_d_frame, "null argument");
	CKD_REQUIRE(renderCtx->device == g->ctx->device && renderCtx->resX == g->ctx->resX && renderCtx->resY == g->ctx->resY, "the rendering context must match the gather's");
	CKD_CUDA(cudaSetDevice(g->ctx->device));
	if (0 == g->numStaging)
	{
		g->numStaging = kMaxStaging;
		for (int i = 0; i < g->numStaging; ++i)
		{
			CKD_CUDA(cudaMalloc(&g->d_staging[i], g->frameBytes + size_t(g->ctx->resX)*16)); // + 4 guard rows like every frame of the context
			CKD_CUDA(cudaMemset(g->d_staging[i], 0, g->frameBytes + size_t(g->ctx->resX)*16));
			CKD_CUDA(cudaEventCreateWithFlags(&g->evRendered[i], cudaEventDisableTiming));
			CKD_CUDA(cudaEventCreateWithFlags(&g->evPushed[i], cudaEventDisableTiming));
		}
	}
	const int k = int(g->acquired % unsigned(g->numStaging));
	if (g->stagingBusy[k])
		CKD_CUDA(cudaStreamWaitEvent(renderCtx->stream, g->evPushed[k], 0));
	g->currentStaging = k;
	++g->acquired;
	*out_d_frame = g->d_staging[k];
	return CKD_OK;
}

extern "C" int ckd_gather_acquire(ckd_gather *g, uint32_t **out_d_frame)
{
	CKD_REQUIRE(g, "null argument");
	return ckd_gather_acquire_on(g, g->ctx, out_d_frame);
}

// publishes the frame rendered into the staging frame of the last ckd_gather_acquire (or any device frame d_frame != NULL that
// stays untouched until ckd_gather_flush) as sequence number `seq`: ordered after everything enqueued on the context's stream
extern "C" int ckd_gather_push(ckd_gather *g, const uint32_t *d_frame, unsigned long long seq)
{
	CKD_REQUIRE(g, "null argument");
	return ckd_gather_push_on(g, g->ctx, d_frame, seq);
}

extern "C" int ckd_gather_push_on(ckd_gather *g, ckd_ctx *renderCtx, const uint32_t *d_frame, unsigned long long seq)
{
	CKD_REQUIRE(g && renderCtx, "null argument");
	CKD_REQUIRE(renderCtx->device == g->ctx->device, "the rendering context must live on the gather's device");
	CKD_CUDA(cudaSetDevice(g->ctx->device));
	int k = -1;
	if (nullptr == d_frame)
	{
		CKD_REQUIRE(g->currentStaging >= 0, "ckd_gather_push without a frame: call ckd_gather_acquire first");
		k = g->currentStaging;
		d_frame = g->d_staging[k];
		g->currentStaging = -1;
	}
	else
	{
		// a staging frame handed out earlier (two frames in flight): find it again
		for (int i = 0; i < g->numStaging; ++i)
			if (g->d_staging[i] == d_frame) k = i;
	}
	const int slot = int(seq % unsigned(g->slots));
	cudaEvent_t evRendered = (k >= 0) ? g->evRendered[k] : g->evAdhoc;
	CKD_CUDA(cudaEventRecord(evRendered, renderCtx->stream));
	CKD_CUDA(cudaStreamWaitEvent(g->pushStream, evRendered, 0));
	if (seq >= unsigned(g->slots))
	{
		gather_wait_drained_kernel<<<1, 1, 0, g->pushStream>>>(g->ctl, slot, seq + 1 - unsigned(g->slots), g->timeoutNs);
		g->ctx->launches++;
	}
	CKD_CUDA(cudaMemcpyAsync(SlotPtr(g, slot), d_frame, g->frameBytes, cudaMemcpyDeviceToDevice, g->pushStream)); // peer copy over NVLink when the ring is remote
	gather_signal_ready_kernel<<<1, 1, 0, g->pushStream>>>(g->ctl, slot, seq + 1);
	g->ctx->launches++;
	cudaError_t err = cudaPeekAtLastError();
	if (cudaSuccess != err)
		return ckd_cuda_fail(err, "gather push", __FILE__, __LINE__);
	if (k >= 0)
	{
		CKD_CUDA(cudaEventRecord(g->evPushed[k], g->pushStream));
		g->stagingBusy[k] = true;
	}
	if (g->mapped)
		g->peerBytes += g->frameBytes;
	return CKD_OK;
}

// collector: consume sequence number `seq` (must be called for seq = 0, 1, 2, ... in order).  mode bits: CKD_GATHER_CHECKSUM
// folds the frame into the checksum table, CKD_GATHER_TO_HOST copies it to h_dest (page-locked memory for an asynchronous
// copy); 0 just releases the slot.  Everything is enqueued on the gather's consumer stream; nothing blocks the host.
extern "C" int ckd_gather_pop(ckd_gather *g, unsigned long long seq, int mode, void *h_dest)
{
	CKD_REQUIRE(g, "null argument");
	CKD_REQUIRE(g->owner, "only the collector pops");
	CKD_REQUIRE(0 == (mode & CKD_GATHER_TO_HOST) || nullptr != h_dest, "CKD_GATHER_TO_HOST needs a destination");
	CKD_CUDA(cudaSetDevice(g->ctx->device));
	const int slot = int(seq % unsigned(g->slots));
	const bool toHost = 0 != (mode & CKD_GATHER_TO_HOST);
	gather_wait_ready_kernel<<<1, 1, 0, g->popStream>>>(g->ctl, slot, seq + 1, g->timeoutNs);
	g->ctx->launches++;
	if (mode & CKD_GATHER_CHECKSUM)
	{
		const size_t numQuads = g->frameBytes/16;
		const unsigned blocks = unsigned(g->ctx->numSMs)*2;
		gather_checksum_kernel<<<blocks, 256, 0, g->popStream>>>(g->ctl, slot, seq + 1, reinterpret_cast<const uint4 *>(SlotPtr(g, slot)), numQuads,
			g->d_sums + seq % g->sumCapacity, g->d_scratch, toHost ? 0 : 1);
		g->ctx->launches++;
	}
	if (toHost)
		CKD_CUDA(cudaMemcpyAsync(h_dest, SlotPtr(g, slot), g->frameBytes, cudaMemcpyDeviceToHost, g->popStream));
	if (toHost || 0 == (mode & CKD_GATHER_CHECKSUM))
	{
		gather_signal_drained_kernel<<<1, 1, 0, g->popStream>>>(g->ctl, slot, seq + 1);
		g->ctx->launches++;
	}
	cudaError_t err = cudaPeekAtLastError();
	if (cudaSuccess != err)
		return ckd_cuda_fail(err, "gather pop", __FILE__, __LINE__);
	CKD_CUDA(cudaEventRecord(g->evPopped[seq % ckd_gather::kPopEvents], g->popStream));
	return CKD_OK;
}

// blocks the host until the pop of `seq` has completed (its frame is in h_dest); seq must be one of the last 16 popped
extern "C" int ckd_gather_wait_pop(ckd_gather *g, unsigned long long seq)
{
	CKD_REQUIRE(g && g->owner, "only the collector pops");
	CKD_CUDA(cudaEventSynchronize(g->evPopped[seq % ckd_gather::kPopEvents]));
	return CKD_OK;
}

// makes the context's stream wait for everything this process has pushed / popped so far: an event recorded on the context's
// stream afterwards covers the gather (that is how bench.py times it), and ckd_sync() then waits for it
extern "C" int ckd_gather_flush(ckd_gather *g)
{
	CKD_REQUIRE(g, "null argument");
	CKD_CUDA(cudaSetDevice(g->ctx->device));
	CKD_CUDA(cudaEventRecord(g->evFlush, g->pushStream));
	CKD_CUDA(cudaStreamWaitEvent(g->ctx->stream, g->evFlush, 0));
	if (g->popStream)
	{
		CKD_CUDA(cudaEventRecord(g->evFlush, g->popStream));
		CKD_CUDA(cudaStreamWaitEvent(g->ctx->stream, g->evFlush, 0));
	}
	return CKD_OK;
}

// 0 when no wait has timed out; CKD_ERR_TIMEOUT (and a description in ckd_last_error) otherwise.  Synchronises the gather's streams.
extern "C" int ckd_gather_status(ckd_gather *g)
{
	CKD_REQUIRE(g, "null argument");
	CKD_CUDA(cudaSetDevice(g->ctx->device));
	CKD_CUDA(cudaStreamSynchronize(g->pushStream));
	if (g->popStream) CKD_CUDA(cudaStreamSynchronize(g->popStream));
	unsigned long long status = 0;
	CKD_CUDA(cudaMemcpy(&status, &g->ctl->status, sizeof(status), cudaMemcpyDeviceToHost));
	if (0 == status)
		return CKD_OK;
	const unsigned kind = unsigned(status >> 56);
	GatherControl snapshot;
	std::string flags;
	if (cudaSuccess == cudaMemcpy(&snapshot, g->ctl, sizeof(snapshot), cudaMemcpyDeviceToHost))
		for (int i = 0; i < g->slots; ++i)
			flags += " [" + std::to_string(i) + "] ready " + std::to_string(snapshot.ready[i]) + " drained " + std::to_string(snapshot.drained[i]);
	ckd_set_error(std::string("ckd_gather: ") + (1 == kind ? "a producer timed out waiting for a free slot" : "the collector timed out waiting for a frame")
		+ " (sequence number " + std::to_string((status & ((1ull << 56) - 1)) - 1) + "; slot flags now:" + flags + ")");
	return CKD_ERR_TIMEOUT;
}

extern "C" int ckd_gather_checksums(ckd_gather *g, unsigned long long first_seq, unsigned count, unsigned long long *out_sums)
{
	CKD_REQUIRE(g && out_sums, "null argument");
	CKD_REQUIRE(g->owner, "only the collector holds checksums");
	CKD_REQUIRE(count <= g->sumCapacity, "more checksums requested than the table keeps");
	CKD_CUDA(cudaSetDevice(g->ctx->device));
	CKD_CUDA(cudaStreamSynchronize(g->popStream));
	for (unsigned i = 0; i < count; ++i)
		CKD_CUDA(cudaMemcpy(out_sums + i, g->d_sums + (first_seq + i) % g->sumCapacity, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	return CKD_OK;
}

extern "C" unsigned long long ckd_gather_peer_bytes(const ckd_gather *g) { return g ? g->peerBytes : 0; }
extern "C" int ckd_gather_slots(const ckd_gather *g) { return g ? g->slots : 0; }

// the checksum of ckd_gather_pop(CKD_GATHER_CHECKSUM) for a frame that is already in this context's memory (same kernel, no
// ring): what a single-GPU run compares the gathered stream with
extern "C" int ckd_frame_checksum(ckd_ctx *ctx, const uint32_t *d_frame, unsigned long long *out_sum)
{
	CKD_REQUIRE(ctx && d_frame && out_sum, "null argument");
	CKD_CUDA(cudaSetDevice(ctx->device));
	if (nullptr == ctx->d_checksumWork)
	{
		CKD_CUDA(cudaMalloc(&ctx->d_checksumWork, 3*sizeof(unsigned long long))); // [0] partial sum, [1] CTAs done, [2] result
		CKD_CUDA(cudaMemset(ctx->d_checksumWork, 0, 3*sizeof(unsigned long long)));
	}
	unsigned long long *d_work = ctx->d_checksumWork;
	const size_t numQuads = size_t(ctx->resX)*ctx->resY/4;
	gather_checksum_kernel<<<unsigned(ctx->numSMs)*2, 256, 0, ctx->stream>>>(nullptr, 0, 1, reinterpret_cast<const uint4 *>(d_frame), numQuads, d_work + 2, d_work, 0);
	CKD_CHECK_LAUNCH(ctx);
	CKD_CUDA(cudaMemcpyAsync(out_sum, d_work + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
	CKD_CUDA(cudaStreamSynchronize(ctx->stream));
	return CKD_OK;
}

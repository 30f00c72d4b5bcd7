// ckd_timeline.cpp -- frame-sharded timeline rendering (BASELINE config 5).
//
// The reference renders its timeline on one machine, frame after frame, inside main()'s loop (main.cpp:318-350: time from
// the audio stream, Demo_Draw, Display::Update).  Here the same loop runs once per GPU of a box, each process taking the
// frames i with i % world == rank; the finished device frames travel to rank 0 through the gather of include/ckd.h (slot
// ring in rank 0's HBM, peer copies over NVLink, device-side flags) and rank 0 hands them on in order: to nowhere (checksums
// only), to a ring of page-locked host buffers, or to the frame sink that stands in for Display::Update.
//
// The loop lives here, in native code, because at several thousand frames per second and GPU the per-frame host work has
// to stay in the microseconds: a producer iteration is Rocket evaluation + the part's launches + three stream operations.

#include "ckd_host_internal.h"

#include <deque>
#include <string>

namespace {

struct PendingFrame { unsigned long long seq; uint32_t *buffer; unsigned frame; };

bool Ok(int rc, const char *what) { return ckdhost::Check(rc, what); }

} // namespace

unsigned CkdTimeline_Owner(unsigned frame, unsigned world, unsigned collectorSkip)
{
	if (world <= 1)
		return 0;
	if (collectorSkip <= 1)
		return frame % world;
	const unsigned cycle = collectorSkip*(world - 1) + 1;
	const unsigned c = frame % cycle;
	return (0 == c) ? 0 : 1 + (c - 1) % (world - 1);
}

unsigned CkdTimeline_DefaultCollectorSkip(unsigned world) { return (world >= 4) ? 2 : 1; }

bool CkdTimeline_Render(const CkdTimelineJob *job)
{
	CkdHost_SelectLane(0);                           // (a no-op before CkdHost_Create: the check below reports that)
	ckd_ctx *ctx = CkdHost_Context();
	if (nullptr == ctx || nullptr == job || nullptr == job->times)
	{
		SetLastError("CkdTimeline_Render: no context / job");
		return false;
	}
	if (0 == job->world || job->rank >= job->world)
	{
		SetLastError("CkdTimeline_Render: rank outside [0, world)");
		return false;
	}
	ckd_gather *gather = job->gather;
	const bool collector = nullptr != gather && 0 == job->rank;
	const bool toHost = collector && 0 != (job->popMode & CKD_GATHER_TO_HOST);
	const bool useSink = toHost && nullptr == job->hostRing;
	if (toHost && !useSink && (job->hostRingFrames < 2 || job->hostRingFrames > 8))
	{
		SetLastError("CkdTimeline_Render: the host ring must hold 2..8 frames");
		return false;
	}

	// frames must not depend on which frame the same GPU rendered before (ckd.h: ckd_set_frame_independent)
	if (!Ok(ckd_set_frame_independent(ctx, 1), "CkdTimeline_Render"))
		return false;

	// lanes: this rank's frames rotate through several contexts, each with its own stream (ckd_host.h)
	const unsigned numLanes = (job->lanes >= 2 && job->lanes <= 4) ? job->lanes : 1;
	if (numLanes > 1 && !CkdHost_PrepareLanes(int(numLanes)))   // copies lane 0's inputs, the flag above included
	{
		ckd_set_frame_independent(ctx, 0);
		return false;
	}
	ckd_ctx *lanes[4] = { ctx, ctx, ctx, ctx };
	for (unsigned i = 1; i < numLanes; ++i)
		lanes[i] = CkdHost_LaneContext(int(i));
	unsigned rendered = 0;

	SetLastError("");
	bool ok = true;
	std::deque<PendingFrame> pending;                 // rank 0: copies to the host that are in flight
	const size_t maxPending = useSink ? 2 : (job->hostRingFrames > 1 ? job->hostRingFrames - 1 : 1);

	auto retire = [&](size_t keep)
	{
		while (ok && pending.size() > keep)
		{
			const PendingFrame done = pending.front();
			pending.pop_front();
			ok = Ok(ckd_gather_wait_pop(gather, done.seq), "CkdTimeline_Render: ckd_gather_wait_pop");
			if (ok && useSink)
				ok = CkdSink_Commit(done.buffer, done.frame);
		}
	};

	for (unsigned pass = 0; ok && pass < job->passes; ++pass)
	{
		for (unsigned i = 0; ok && i < job->numFrames; ++i)
		{
			const unsigned long long seq = job->seqBase + (unsigned long long)(pass)*job->numFrames + i;
			if (CkdTimeline_Owner(i, job->world, job->collectorSkip) == job->rank)
			{
				ckd_ctx *lane = lanes[rendered % numLanes];
				if (numLanes > 1)
					CkdHost_SelectLane(int(rendered % numLanes));
				++rendered;
				uint32_t *d_frame = nullptr;
				if (nullptr != gather)
				{
					ok = Ok(ckd_gather_acquire_on(gather, lane, &d_frame), "CkdTimeline_Render: ckd_gather_acquire");
					if (!ok) break;
				}
				CkdHost_SetDeviceTarget(d_frame);     // nullptr (no gather): the context's own frame
				CkdHost_SetTime(job->times[i]);
				Demo_Draw(nullptr, float(job->times[i]), job->delta); // false = the demo is over: the frame keeps its old content, the stream its shape
				CkdHost_SetDeviceTarget(nullptr);
				if (!CkdHost_GetLastError().empty())
				{
					ok = false;
					break;
				}
				if (nullptr != gather)
					ok = Ok(ckd_gather_push_on(gather, lane, nullptr, seq), "CkdTimeline_Render: ckd_gather_push");
			}
			if (ok && collector)
			{
				uint32_t *h_dest = nullptr;
				if (toHost)
				{
					if (useSink)
					{
						h_dest = CkdSink_Acquire();
						if (nullptr == h_dest) { ok = false; break; }
					}
					else
					{
						retire(maxPending - 1);       // the buffer about to be reused must have been delivered
						h_dest = job->hostRing[seq % job->hostRingFrames];
					}
				}
				ok = ok && Ok(ckd_gather_pop(gather, seq, job->popMode, h_dest), "CkdTimeline_Render: ckd_gather_pop");
				if (ok && toHost)
				{
					pending.push_back({ seq, h_dest, i });
					retire(maxPending);
				}
			}
		}
	}
	retire(0);
	if (nullptr != gather)
	{
		ok = Ok(ckd_gather_flush(gather), "CkdTimeline_Render: ckd_gather_flush") && ok;
	}
	if (numLanes > 1)
	{
		// the caller's stream (lane 0's) stands for the whole job: it waits for what the other lanes still have in flight
		CkdHost_SelectLane(0);
		for (unsigned i = 1; i < numLanes; ++i)
		{
			ok = Ok(ckd_join(ctx, lanes[i]), "CkdTimeline_Render: ckd_join") && ok;
			ckd_set_frame_independent(lanes[i], 0);
		}
	}
	ckd_set_frame_independent(ctx, 0);
	return ok;
}

extern "C" {

// ctypes hook: the job as plain arguments
int ckdhost_timeline_render(const double *times, unsigned numFrames, unsigned passes, unsigned rank, unsigned world, void *gather, int popMode,
	uint32_t *const *hostRing, unsigned hostRingFrames, unsigned long long seqBase, float delta, unsigned collectorSkip, unsigned lanes)
{
	CkdTimelineJob job = { times, numFrames, passes, rank, world, static_cast<ckd_gather *>(gather), popMode, hostRing, hostRingFrames, seqBase, delta, lanes, collectorSkip };
	return CkdTimeline_Render(&job) ? 0 : -1;
}
unsigned ckdhost_timeline_owner(unsigned frame, unsigned world, unsigned collectorSkip) { return CkdTimeline_Owner(frame, world, collectorSkip); }
unsigned ckdhost_timeline_default_skip(unsigned world) { return CkdTimeline_DefaultCollectorSkip(world); }
unsigned long long ckdhost_launch_count() { return CkdHost_LaunchCount(); }

} // extern "C"

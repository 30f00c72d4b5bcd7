#!/usr/bin/env python3
"""bench.py -- headline benchmark of cookiedough_b200 (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload "effect-suite-4k" (BASELINE configs 1-3 at 4K): one suite pass renders one 3840x2160 frame through each of
the 12 effect entry points of the hot path at their pinned Rocket rows (the 7 raymarch variants, landscape, tunnelscape, ball
with and without beams, twister), each including its own post chain (Fx_Blit_2x2, polar remap, in-place box blur, blends)
exactly as the reference's X_Draw does.  One STEP = SUITE_PASSES such passes (so that the K timed steps last about a
second).  Metric: Mpixel/s of finished output pixels.

  value   device-resident: parameters evaluated, maps resident in HBM, frames stay on the GPU; CUDA-event timed.
  e2e     the same frames through the reference-facing C++ host layer: Rocket evaluation on the host, X_Draw(pDest, time,
          delta) into a pinned HOST buffer, i.e. every frame is copied device->host inside the timed region.
  N > 1   the suite is weak scaling (every rank renders K steps, no exchange).

Sub-records of the same JSON line (the other BASELINE configs; every one carries its CPU figure at N = 1):
  timeline        config 5: the 600-frame directors-cut timeline through Demo_Draw at 4K, frame i -> rank i mod N (STRONG
                  scaling), every frame published to rank 0 through the peer-memory gather (ckd_gather_*, include/ckd.h: slot
                  ring in rank 0's HBM, copy engines over NVLink, device-side flags, no NCCL on the data path) and checksummed
                  there; plus the same with every gathered frame copied to pinned host memory on rank 0 (e2e).
  post_chain_4k   config 4: Polar_Blit(inverse) -> BoxBlur32 in place -> Fx_Blit_2x2 -> blends -> BoxBlur_32 kGauss ->
                  Polar_BlitA on seeded buffers, per op achieved GB/s against the measured HBM peak.
  config1_720p    configs 1-2 at the reference's native 1280x720: per effect device ms / fps, e2e fps, CPU ms.
  parity          per effect at the benched rows: % exact / <=1 / <=2 LSB / max delta against oracle/_ref (N = 1).
"""
import argparse
import hashlib
import json
import os
import struct
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

RES_X, RES_Y = 3840, 2160
ROW_RATE = (170.0 / (60.0 * (170.0 / 174.0))) * 16.0
SUITE_PASSES = 30          # suite passes per step (our arm); the reference arm's step is one pass (a bounded sample)

# label, C-ABI effect, host/reference effect, close flag, pinned Rocket row
SUITE = [
    ("plasma", "plasma", "plasma", None, 2600),
    ("nautilus", "nautilus", "nautilus", None, 5700),
    ("spikey_close", "spikey", "spikey_close", True, 6800),
    ("spikey_distant", "spikey", "spikey_distant", False, 3600),
    ("tunnel", "tunnel", "tunnel", None, 4500),
    ("sinuses", "sinuses", "sinuses", None, 7800),
    ("laura", "laura", "laura", None, 8900),
    ("landscape", "landscape", "landscape", None, 500),
    ("tunnelscape", "tunnelscape", "tunnelscape", None, 4300),
    ("ball", "ball", "ball", None, 1500),
    ("ball_beams", "ball", "ball", None, 2060),
    ("twister", "twister", "twister", None, 2008),
]
INTEGER_LABELS = {"landscape", "tunnelscape", "ball", "ball_beams", "twister"}
PIXELS_PER_PASS = len(SUITE) * RES_X * RES_Y


def suite_config():
    """the `config` object: identical in both arms (everything run-specific lives in the sibling key `run`)"""
    return {"workload": "effect-suite-4k", "res": [RES_X, RES_Y], "effects": [s[0] for s in SUITE]}


# FP operations per FX-map pixel at the pinned rows (the measured mean, not the <= bound).
#   profiles/r02_ref_fp_ops.json  -- counted on the REFERENCE: its translation units rebuilt with every SSE float intrinsic and
#       scalar float operator routed through counting wrappers, run at the pinned rows at 4K (tests/tools/count_ref_ops.py).
#       This is the denominator SURVEY 7(5) asks for: independent of this repo's kernels.
#   fallback: profiles/r01_fp_ops.json, executed thread-level FP instructions of this repo's own kernels (ncu), the round-1 numbers.
# 1 op = one FP add / mul / compare / min-max / sqrt / rsqrt / division / conversion per lane.
FLOP_FALLBACK = {
    "raymarch_plasma": 1461.0, "raymarch_nautilus": 1814.0, "raymarch_spikey_close": 1673.0, "raymarch_spikey_distant": 1688.0,
    "raymarch_sinuses": 2490.0, "raymarch_laura": 1494.0, "raymarch_tunnel": 268.0,
}


def flop_table():
    path = os.path.join(REPO, "profiles", "r02_ref_fp_ops.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        ops = {k: float(v["ops_per_fx_pixel"]) for k, v in d["kernels"].items()}
        addmul = {k: float(v["fadd"] + v["fmul"]) for k, v in d["kernels"].items()}
        if all(k in ops for k in FLOP_FALLBACK):
            return ops, addmul, "counted on the reference (profiles/r02_ref_fp_ops.json, tests/tools/count_ref_ops.py)"
    addmul = {}
    path = os.path.join(REPO, "profiles", "r01_fp_ops.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        addmul = {k: float(v["fadd"] + v["fmul"]) for k, v in d["kernels"].items()}
    return dict(FLOP_FALLBACK), addmul, "executed FP instructions of this repo's kernels (profiles/r01_fp_ops.json)"


# what ncu says bounds each kernel (profiles/r0*_notes.md): reported next to the roofline fraction so that a small HBM
# fraction of a kernel that is not HBM bound is not misread
LIMITER = {
    "raymarch": "instruction issue: FMUL/FADD chains without FMA contraction (bit parity) + the non-FP instructions of every LUT lookup",
    "old_blur": "integer ALU pipe + the dependent chain of the saturating in-place recurrence; DRAM traffic = algorithmic bytes",
    "voxel": "L2 gather latency of the height/colour map samples (warp per ray; the landscape: issue slots + L1 look-ups of one full wave)",
    "polar_blit": "L1TEX: 16 four-byte texel gathers per thread fill the LSU queue (l1tex 75 %, lg/mio throttle; profiles/r02_notes.md); DRAM traffic = algorithmic bytes",
    "fx_blit_2x2": "L2 write-back of the 33 MB frame",
    "blend": "HBM / L2 bandwidth",
    "rect_blit": "HBM / L2 bandwidth (small rectangles: launch latency)",
    "memset32": "launch latency (4 MB)",
}


def limiter_of(name):
    for prefix, text in LIMITER.items():
        if name.startswith(prefix):
            return text
    return None


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.mark = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def begin_region(self):
        self.mark = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines[self.mark:]:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(local):
    """one process per GPU: keep the process (and with it the pinned frame buffers it allocates and the copies it drives) on the
    CPU cores of the GPU's own NUMA node.  Returns the number of cores bound to, or None when the topology is not exposed."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(local)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def dist_setup(n_gpus):
    # NCCL_DEBUG is left as the caller set it.  NCCL prints its version banner on fd 1 either way (NCCL_DEBUG_FILE only redirects the
    # levels above VERSION): main() has moved fd 1 to stderr and kept the real stdout for the one JSON line (claim_stdout).
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        bind_to_gpu_numa(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        return rank, world, local, dist
    return rank, world, local, None


def reduce_max(dist, value):
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(dist, value):
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ---------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation (oracle/_ref) on the host cores.  BASELINE.md 3.2: the identical public
# entry point, all cores AND one thread, first two calls discarded, median of >= 10, CPU model next to the numbers.
# ---------------------------------------------------------------------------------------------------------------

def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def set_omp_threads(n):
    """The reference parallelises with OpenMP.  torchrun exports OMP_NUM_THREADS=1 to every rank, which would time the reference
    on one core: set the count on the already loaded runtime."""
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass
    return n


def use_all_host_threads():
    return set_omp_threads(cpu_threads())


def timed_median(fn, reps=10, discard=2):
    for _ in range(discard):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


def reference_at(res_y, demo=False):
    from cookiedough_b200.assets import Assets
    from oracle import ref as oref
    if not oref.available(res_y):
        return None
    res_x = res_y * 16 // 9
    return oref.Reference.get(res_y, Assets(res_x, res_y), demo=demo)


def pixel_parity(ours, ref):
    a = np.ascontiguousarray(ours).view(np.uint8).reshape(-1, 4).astype(np.int16)
    b = np.ascontiguousarray(ref).view(np.uint8).reshape(-1, 4).astype(np.int16)
    d = np.abs(a - b).max(axis=1)
    n = float(d.size)
    return {"exact_pct": 100.0 * float((d == 0).sum()) / n, "le1_pct": 100.0 * float((d <= 1).sum()) / n,
            "le2_pct": 100.0 * float((d <= 2).sum()) / n, "max_delta": int(d.max())}


def run_reference(args):
    """reference arm: the UNMODIFIED reference compiled from /root/reference (oracle/_ref), its own X_Draw on all host threads.
    A step is ONE pass of the suite (12 frames at 4K, about 0.6 s on 16 cores): a bounded sample of our arm's step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = use_all_host_threads()
    base = {"impl": "reference", "metric": "Mpixel/s", "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
            "config": suite_config()}
    R = reference_at(RES_Y)
    if R is None:
        base["unavailable"] = "oracle/_ref (compiled reference) is not present in this checkout"
        emit(base)
        return 0
    out = R.frame()

    def step():
        for _, _, ref_eff, _, row in SUITE:
            R.set_row(row)
            R.draw(ref_eff, out)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = PIXELS_PER_PASS * args.steps / dt / 1e6
    sample = f"each step = 1 pass of the suite (12 frames at {RES_X}x{RES_Y}); {args.steps} steps after {max(args.warmup, 1)} warm-up steps, OpenMP on {threads} threads"
    base.update({"value": value, "ms_per_step": 1e3 * dt / args.steps, "gpu_launches": 0,
                 "run": {"suite_passes_per_step": 1, "cpu_model": cpu_model()},
                 "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": threads, "kind": "reference", "sample": sample, "cpu_model": cpu_model()},
                 "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(base)
    return 0


# ---------------------------------------------------------------------------------------------------------------
# legs of our arm
# ---------------------------------------------------------------------------------------------------------------

def timeline_leg(host, ctx, dist, rank, world, frames, passes, with_cpu, lanes=2):
    """BASELINE config 5 with the gather.  Returns the sub-record (rank 0) or None."""
    import torch
    from cookiedough_b200 import capi, sharding

    frame_bytes = RES_X * RES_Y * 4
    times = sharding.timeline_times(frames)
    skip = sharding.default_collector_skip(world)      # the library's default (hostapi passes it on when collector_skip is None)
    lanes = max(1, min(4, lanes))                      # frames in flight per GPU (ckd_host.h: lanes)
    slots = min(64, max(8, 4 * world))   # every producer may run four frames ahead of the collector (33 MB per 4K slot in rank 0's HBM)

    # the ring lives on rank 0; its CUDA IPC handle travels once, over the control plane (torch.distributed)
    if rank == 0:
        gather = capi.Gather(ctx, slots=slots)
        handle = gather.export()
    else:
        gather, handle = None, bytes(capi.GATHER_HANDLE_BYTES)
    if dist is not None:
        t = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, src=0)
        if rank != 0:
            gather = capi.Gather(ctx, handle=bytes(t.cpu().tolist()))
    gather.set_timeout_ms(30000)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    seq = 0
    # warm-up pass (also the first list of checksums)
    host.timeline_render(times, rank=rank, world=world, gather=gather, passes=1, pop_mode=capi.GATHER_CHECKSUM, seq_base=seq, collector_skip=skip, lanes=lanes)
    ctx.sync()
    gather.status()
    warm_sums = gather.checksums(seq, frames) if rank == 0 else None
    seq += frames
    barrier()

    launches0 = host.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    host.timeline_render(times, rank=rank, world=world, gather=gather, passes=passes, pop_mode=capi.GATHER_CHECKSUM, seq_base=seq, collector_skip=skip, lanes=lanes)
    ev1.record()                        # CkdTimeline_Render ends with ckd_gather_flush: the stream waits for every push and pop
    torch.cuda.synchronize()
    gather.status()
    ms = reduce_max(dist, ev0.elapsed_time(ev1))
    launches = reduce_sum(dist, host.launch_count() - launches0)
    sums = gather.checksums(seq, frames) if rank == 0 else None
    last_sums = gather.checksums(seq + (passes - 1) * frames, frames) if rank == 0 else None
    seq += passes * frames
    peer_bytes_timed = reduce_sum(dist, gather.peer_bytes()) * passes / (passes + 1.0)   # the warm-up pass pushed too

    # e2e: every gathered frame also goes to pinned host memory on rank 0 (what a sink / encoder / display would read)
    ring = [ctx.malloc_host(frame_bytes) for _ in range(4)] if rank == 0 else None
    barrier()
    t0 = time.perf_counter()
    host.timeline_render(times, rank=rank, world=world, gather=gather, passes=1, pop_mode=capi.GATHER_CHECKSUM | capi.GATHER_TO_HOST,
                         host_ring=ring, seq_base=seq, collector_skip=skip, lanes=lanes)
    ctx.sync()
    e2e_s = reduce_max(dist, time.perf_counter() - t0)
    gather.status()
    e2e_sums = gather.checksums(seq, frames) if rank == 0 else None
    seq += frames

    # the same loop without the gather (frames stay where they were rendered): what the exchange costs
    barrier()
    eva, evb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eva.record()
    host.timeline_render(times, rank=rank, world=world, gather=None, passes=1, lanes=lanes)
    evb.record()
    torch.cuda.synchronize()
    nogather_ms = reduce_max(dist, eva.elapsed_time(evb))

    rec = None
    if rank == 0:
        digest = hashlib.sha256(struct.pack(f"<{frames}Q", *sums)).hexdigest()
        px = frames * RES_X * RES_Y
        fps = frames * passes / (ms * 1e-3)
        rec = {
            "config": {"workload": "timeline-4k", "frames": frames, "res": [RES_X, RES_Y], "api": "Demo_Draw via CkdTimeline_Render (include/ckd_host.h)"},
            "scaling": "strong", "n_gpus": world, "passes": passes,
            "sharding": ("frame i -> rank i mod N" if skip <= 1 else f"weighted round-robin (CkdTimeline_Owner, collector_skip = {skip}): rank 0, which collects and checksums every frame, renders 1 frame per {skip} rounds of the other ranks ({len(sharding.frames_for_rank(frames, 0, world, skip))} of {frames} frames)")
                        + "; every frame pushed to a slot ring in rank 0's HBM (ckd_gather_*: CUDA IPC mapping, copy-engine peer copies over NVLink, device-side ready/drained flags), checksummed there in order; no NCCL on the data path",
            "collector_skip": skip, "lanes_per_gpu": lanes,
            "lanes_note": "each rank alternates its frames between two contexts on its GPU (own render targets and stream): the latency-bound kernels of one frame overlap the issue-bound ones of the other",
            "gathered_to_rank0_fps": fps, "value": px * passes / (ms * 1e-3) / 1e6, "unit": "Mpixel/s", "ms_per_pass": ms / passes,
            "no_gather_fps": frames / (nogather_ms * 1e-3),
            "nvlink_bytes_per_frame": peer_bytes_timed / (frames * passes), "nvlink_gbs": peer_bytes_timed / (ms * 1e-3) / 1e9,
            "nvlink_note": "frames rendered on rank 0 are copied inside its own HBM; the others cross NVLink once; rank 0 ingests at most what one GPU's NVLink takes: 770 GB/s measured peer copy (900 nominal) = 23 200 4K frames/s over the wire",
            "e2e": {"fps": frames / e2e_s, "value": px / e2e_s / 1e6, "unit": "Mpixel/s", "d2h_bytes_per_pass": frames * frame_bytes, "h2d_bytes_per_pass": 0,
                    "note": "every gathered frame copied to a pinned host ring on rank 0 inside the timed region; one PCIe link (rank 0's) carries all frames: that link is the ceiling at every N"},
            "gpu_launches": int(launches), "ring_slots": slots,
            "checksum": "sum_i pixel[i]*(2i+1) mod 2^64 per frame, computed on rank 0 over the gathered frame",
            "checksums_sha256": digest, "checksums_head": [str(s) for s in sums[:6]],
            "checksums_stable": bool(sums == warm_sums and sums == last_sums and sums == e2e_sums),
        }
        for p in ring:
            ctx.free_host(p)
    barrier()
    gather.close()

    if rank == 0 and with_cpu:
        R = reference_at(RES_Y, demo=True)
        if R is not None and R.demo:
            threads = use_all_host_threads()
            stride = max(1, frames // 20)
            sample = list(range(stride // 2, frames, stride))
            out = R.frame()
            R.set_time(times[sample[0]])
            R.demo_draw(out)
            t0 = time.perf_counter()
            for i in sample:
                R.set_time(times[i])
                R.demo_draw(out)
            dt = time.perf_counter() - t0
            rec["cpu_baseline"] = {"fps": len(sample) / dt, "value": len(sample) * RES_X * RES_Y / dt / 1e6, "unit": "Mpixel/s", "cores": threads, "kind": "reference",
                                   "cpu_model": cpu_model(),
                                   "sample": f"every {stride}th frame of the timeline ({len(sample)} frames spread over all parts) through the reference's Demo_Draw at {RES_X}x{RES_Y} after 1 warm-up frame"}
    return rec


def post_chain_leg(ctx, with_cpu, hbm_peak):
    """BASELINE config 4 at 3840x2160, sequenced like tests/test_gpu_post.py::test_post_chain_4k_against_reference.  Every op is
    timed alone (CUDA events inside the library around each launch) over 10 repetitions that rotate through 5 buffer sets
    (5 x 33 MB per operand > the 126 MB L2), after 2 discarded repetitions."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import post_cases as pc
    w, h = RES_X, RES_Y
    n = w * h
    fx_n = ctx.fx_x * ctx.fx_y
    sets = 5
    src = pc.seeded(n, "mul")
    dst = pc.seeded(n, "mix")
    fx = pc.seeded(fx_n, "noise")
    d_src = [ctx.to_device(src, pad_elems=4 * w) for _ in range(sets)]
    d_dst = [ctx.to_device(dst, pad_elems=4 * w) for _ in range(sets)]
    d_tmp = [ctx.to_device(src, pad_elems=4 * w) for _ in range(sets)]
    d_fx = [ctx.to_device(fx, pad_elems=4 * w) for _ in range(sets)]

    traffic = {}
    tpath = os.path.join(REPO, "profiles", "ncu_traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            traffic = json.load(f)

    # (label, reference entry point, algorithmic bytes per output pixel [SURVEY 8d], launch)
    ops = [
        ("Polar_Blit(inverse)", "polar.cpp:135", 16.0, lambda k: ctx.polar_blit(d_dst[k], d_src[k], True)),
        ("BoxBlur32 in place s=0.11", "deprecated/boxblur.cpp:222", 16.0, lambda k: ctx.old_blur("hv", d_dst[k], d_dst[k], w, h, 0.11)),
        ("HorizontalBoxBlur32 in place s=0.2", "deprecated/boxblur.cpp:44", 8.0, lambda k: ctx.old_blur("h", d_dst[k], d_dst[k], w, h, 0.2)),
        ("VerticalBoxBlur32 in place s=0.2", "deprecated/boxblur.cpp:128", 8.0, lambda k: ctx.old_blur("v", d_dst[k], d_dst[k], w, h, 0.2)),
        ("Fx_Blit_2x2", "fx-blitter.cpp:27", 4.0 + 4.0 * fx_n / n, lambda k: ctx.fx_blit_2x2(d_tmp[k], d_fx[k])),
        ("MixSrc32", "util.cpp:661", 12.0, lambda k: ctx.blend("MixSrc32", d_dst[k], d_tmp[k], n)),
        ("SoftLight32", "util.cpp:242", 12.0, lambda k: ctx.blend("SoftLight32", d_dst[k], d_tmp[k], n)),
        ("Overlay32", "util.cpp:434", 12.0, lambda k: ctx.blend("Overlay32", d_dst[k], d_tmp[k], n)),
        ("BoxBlur_32 kGauss 3 passes", "boxblur.cpp:301", 16.0, lambda k: ctx.new_blur("hv", d_tmp[k], d_dst[k], w, h, 6.28, 0.1, 3)),
        ("Polar_BlitA", "polar.cpp:180", 20.0, lambda k: ctx.polar_blit(d_dst[k], d_tmp[k], False, alpha=True)),
    ]
    records = []
    for label, cite, bytes_px, launch in ops:
        for k in range(2):
            launch(k)
        ctx.sync()
        reps = 10
        ctx.profile_begin()
        for r in range(reps):
            launch(r % sets)
        stats = ctx.profile_end()
        ms = sum(s["total_ms"] for s in stats.values()) / reps
        gbs = bytes_px * n / (ms * 1e-3) / 1e9
        rec = {"op": label, "reference": cite, "ms": ms, "mpixel_s": n / ms / 1e3, "algo_bytes_per_px": bytes_px, "achieved_gbs": gbs, "frac_of_hbm": gbs / hbm_peak,
               "kernels": {name: {"launches": s["launches"] / reps, "ms": s["total_ms"] / reps, "dram_traffic_bytes": traffic.get(name)} for name, s in stats.items()}}
        records.append(rec)

    if with_cpu:
        R = reference_at(RES_Y)
        if R is not None:
            from oracle.ref import aligned_u32
            threads = use_all_host_threads()
            r_src = aligned_u32(n, pad=4 * w).reshape(h, w); r_src[:] = src.reshape(h, w)
            r_dst = aligned_u32(n, pad=4 * w).reshape(h, w); r_dst[:] = dst.reshape(h, w)
            r_tmp = aligned_u32(n, pad=4 * w).reshape(h, w); r_tmp[:] = src.reshape(h, w)
            r_fx = aligned_u32(fx_n, pad=4 * w).reshape(ctx.fx_y, ctx.fx_x); r_fx[:] = fx.reshape(ctx.fx_y, ctx.fx_x)
            cpu_ops = [
                lambda: R.polar_blit(r_dst, r_src, True),
                lambda: R.old_blur("hv", r_dst, r_dst, w, h, 0.11),
                lambda: R.old_blur("h", r_dst, r_dst, w, h, 0.2),
                lambda: R.old_blur("v", r_dst, r_dst, w, h, 0.2),
                lambda: R.fx_blit_2x2(r_tmp, r_fx),
                lambda: R.blend("MixSrc32", r_dst, r_tmp),
                lambda: R.blend("SoftLight32", r_dst, r_tmp),
                lambda: R.blend("Overlay32", r_dst, r_tmp),
                lambda: R.new_blur("hv", r_tmp, r_dst, w, h, 6.28, 0.1, 3),
                lambda: R.polar_blit(r_dst, r_tmp, False, alpha=True),
            ]
            for rec, fn in zip(records, cpu_ops):
                rec["cpu_ms"] = timed_median(fn, reps=10, discard=2)
                rec["cpu_threads"] = threads
    for d in d_src + d_dst + d_tmp + d_fx:
        ctx.free(d)
    chain = [r for r in records if not r["op"].startswith(("Horizontal", "Vertical"))]
    out = {"config": {"workload": "post-chain-4k", "res": [RES_X, RES_Y]}, "ops": records,
           "chain_ms": sum(r["ms"] for r in chain), "chain_mpixel_s": n / sum(r["ms"] for r in chain) / 1e3,
           "hbm_peak_gbs": hbm_peak, "timing": "CUDA events around every launch inside the library; 10 repetitions over 5 rotating buffer sets after 2 discards",
           "traffic_source": "profiles/ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch, one ncu --set full capture)"}
    if with_cpu and all("cpu_ms" in r for r in chain):
        out["chain_cpu_ms"] = sum(r["cpu_ms"] for r in chain)
        out["cpu_model"] = cpu_model()
    return out


def parity_and_cpu_leg(host, ctx):
    """N = 1, rank 0: every suite effect at 4K on the reference (all host threads: median of 10 after 2 discards) and its frame
    compared with ours drawn through the host layer at the same pinned row."""
    R = reference_at(RES_Y)
    if R is None:
        return None, {"value": None, "unit": "Mpixel/s", "cores": cpu_threads(), "kind": "reference", "sample": "oracle/_ref not present"}
    threads = use_all_host_threads()
    out_ref = R.frame()
    ours = np.zeros((RES_Y, RES_X), dtype=np.uint32)
    zeros = np.zeros((RES_Y, RES_X), dtype=np.uint32)
    parity, per_effect_ms = {}, {}
    for label, eff, host_eff, close, row in SUITE:
        R.set_row(row)
        R.render_target(0)[:] = 0
        per_effect_ms[label] = timed_median(lambda: R.draw(host_eff, out_ref), reps=10, discard=2)
        R.render_target(0)[:] = 0
        R.draw(host_eff, out_ref)
        ctx.upload(ctx.render_target(0), zeros)
        host.set_row(row)
        host.draw(host_eff, ours)
        p = pixel_parity(ours, out_ref)
        p["required"] = "bit-exact" if label in INTEGER_LABELS else "<= 2 LSB per channel and >= 99.5 % exact"
        p["ok"] = bool(p["max_delta"] == 0) if label in INTEGER_LABELS else bool(p["max_delta"] <= 2 and p["exact_pct"] >= 99.5)
        parity[label] = p
    total_ms = sum(per_effect_ms.values())
    cpu = {"value": PIXELS_PER_PASS / (total_ms * 1e-3) / 1e6, "unit": "Mpixel/s", "cores": threads, "kind": "reference", "cpu_model": cpu_model(),
           "sample": f"one pass of the same suite (12 frames at {RES_X}x{RES_Y}) on oracle/_ref: per effect the median of 10 X_Draw calls after 2 discarded, OpenMP on {threads} threads; value = 12 frames / sum of the medians",
           "per_effect_ms": per_effect_ms}
    return parity, cpu


def leg_720p(local, with_cpu, rsqrt_from=None):
    """BASELINE configs 1-2 at the reference's native 1280x720 (rank 0; the 4K host layer has been closed before this runs)."""
    import torch
    from cookiedough_b200 import capi, hostapi
    from cookiedough_b200.assets import Assets
    res_x, res_y = 1280, 720
    assets = Assets(res_x, res_y)
    host = hostapi.Host(res_x, res_y, local, assets)
    ctx = host.context()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    h_frame = ctx.malloc_host(res_x * res_y * 4)
    d_frame = ctx.malloc(res_x * res_y * 4 + 65536)
    rec = {}
    for label, eff, host_eff, close, row in SUITE:
        host.set_row(row)
        params = capi.params_from_tracks(eff, host.track)
        t = float(np.float32(host.time))
        ts = []
        for i in range(12):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ctx.draw(eff, params, t, d_dest=d_frame, close=close)
            b.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        e2e_ms = timed_median(lambda: host.draw(host_eff, h_frame), reps=10, discard=2)
        rec[label] = {"row": row, "gpu_ms": ms, "gpu_fps": 1e3 / ms, "gpu_mpixel_s": res_x * res_y / ms / 1e3, "e2e_ms": e2e_ms, "e2e_fps": 1e3 / e2e_ms}
    out = {"config": {"workload": "effect-suite-720p", "res": [res_x, res_y], "effects": [s[0] for s in SUITE]},
           "timing": "gpu_*: CUDA events around one X_Draw on the device, median of 10 after 2 discards; e2e_*: the host layer's X_Draw into a pinned host buffer, wall clock, same protocol",
           "per_effect": rec}
    if with_cpu:
        R = reference_at(res_y)
        if R is not None:
            frame = R.frame()
            threads = use_all_host_threads()
            for label, eff, host_eff, close, row in SUITE:
                R.set_row(row)
                rec[label]["cpu_ms_all_threads"] = timed_median(lambda: R.draw(host_eff, frame), reps=10, discard=2)
            set_omp_threads(1)
            for label, eff, host_eff, close, row in SUITE:
                R.set_row(row)
                primary = label == "nautilus"     # config 1's primary case gets the full protocol on one thread, the others a shorter one
                rec[label]["cpu_ms_1_thread"] = timed_median(lambda: R.draw(host_eff, frame), reps=10 if primary else 3, discard=2 if primary else 1)
            use_all_host_threads()
            out["cpu"] = {"cpu_model": cpu_model(), "threads_all": threads, "kind": "reference",
                          "protocol": "the reference's X_Draw (oracle/_ref 720p build); all threads: median of 10 after 2 discards; 1 thread: the same for nautilus (config 1's primary case), median of 3 after 1 discard for the others"}
    ctx.free_host(h_frame)
    ctx.free(d_frame)
    host.close()
    return out


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------

def run_ours(args):
    rank, world, local, dist = dist_setup(args.gpus)
    import torch
    from cookiedough_b200 import capi, hostapi
    from cookiedough_b200.assets import Assets

    torch.cuda.set_device(local)
    assets = Assets(RES_X, RES_Y)
    host = hostapi.Host(RES_X, RES_Y, local, assets, demo=not args.no_timeline)
    ctx = host.context()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    with_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    passes_per_step = max(1, args.suite_passes)
    # The 12 frames of a pass are independent: the device-resident leg renders them on --streams CUDA streams (one context =
    # one set of render targets per stream), so the latency-bound casters and blurs of one frame overlap the issue-bound
    # raymarcher of another.  The host-API legs (e2e) and the per-kernel profile use the single context of the host layer.
    n_streams = max(1, args.streams)
    side_streams = [torch.cuda.Stream() for _ in range(n_streams - 1)]
    ctxs = [ctx] + [capi.Context(RES_X, RES_Y, local, assets) for _ in side_streams]
    for c, st in zip(ctxs[1:], side_streams):
        c.set_stream(st.cuda_stream)

    # parameters of every suite entry, evaluated once by the host layer's Rocket (the device-resident leg feeds them
    # straight to the C ABI; the e2e leg re-evaluates them per frame like the reference does)
    cases = []
    for label, eff, host_eff, close, row in SUITE:
        host.set_row(row)
        cases.append((label, eff, host_eff, close, row, capi.params_from_tracks(eff, host.track), float(np.float32(host.time))))

    # two targets per stream, alternated so that no frame is rewritten back to back
    d_frames = [[c.malloc(RES_X * RES_Y * 4 + 65536) for _ in range(2)] for c in ctxs]

    def pass_device(i):
        for j, (label, eff, host_eff, close, row, params, t) in enumerate(cases):
            k = j % n_streams
            ctxs[k].draw(eff, params, t, d_dest=d_frames[k][(i + j // n_streams) & 1], close=close)

    def step_device(i):
        for p in range(passes_per_step):
            pass_device(i * passes_per_step + p)

    def join_streams():
        for st in side_streams:
            torch.cuda.current_stream().wait_stream(st)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step_device(i)
    barrier()
    if rank == 0:
        # nvidia-smi needs about a second to come up: keep the identical load running until it produces samples
        t_load = time.perf_counter()
        i = 0
        while time.perf_counter() - t_load < 3.0 and len(sampler.lines) < 4:
            pass_device(i)
            i += 1
            if i % 8 == 0:
                torch.cuda.synchronize()
    barrier()
    sampler.begin_region()
    launches0 = sum(c.launch_count() for c in ctxs)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()                       # every stream is idle here (barrier above)
    for st in side_streams:
        st.wait_event(ev0)
    for i in range(args.steps):
        step_device(i)
    join_streams()
    ev1.record()
    torch.cuda.synchronize()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = sum(c.launch_count() for c in ctxs) - launches0
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = reduce_max(dist, elapsed_ms)
    launches = int(reduce_sum(dist, launches))
    barrier()

    if args.device_only:
        # profiling target (tools/ncu_round.sh): only the device-resident passes above, so that a launch list taken under ncu
        # holds exactly the launches `value` and `kernels[*].share` are made of
        if rank == 0:
            emit({"metric": "Mpixel/s", "value": world * PIXELS_PER_PASS * passes_per_step * args.steps / (elapsed_ms * 1e-3) / 1e6, "unit": "Mpixel/s",
                              "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "gpu_launches": launches,
                              "config": suite_config(), "note": "--device-only: device-resident leg only (profiling target), not a bench line"})
        for c in ctxs[1:]:
            c.close()
        host.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- end to end through the reference-facing host API (HOST pDest, D2H inside the timed region) -------------
    frame_bytes = RES_X * RES_Y * 4
    h_frame = ctx.malloc_host(frame_bytes)
    h_frame2 = ctx.malloc_host(frame_bytes)
    e2e_passes = max(1, min(args.steps, 10))

    def pass_e2e():
        for label, eff, host_eff, close, row, params, t in cases:
            host.set_row(row)
            host.draw(host_eff, h_frame)

    def timed_e2e(fn, after=None):
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_passes):
            fn()
        if after is not None:
            after()
        torch.cuda.synchronize()
        return world * PIXELS_PER_PASS * e2e_passes / reduce_max(dist, time.perf_counter() - t0) / 1e6

    pass_e2e()
    e2e_value = timed_e2e(pass_e2e)
    host.set_readback_bands(0)         # the same synchronous calls with the banded read-back off (one copy after the frame is finished)
    e2e_unbanded = timed_e2e(pass_e2e)
    host.set_readback_bands(-1)

    # same calls with the host layer's two-deep frame pipeline (CkdHost_SetPipelined): frame i's copy overlaps frame i+1's render
    counter = [0]

    def pass_pipelined():
        for label, eff, host_eff, close, row, params, t in cases:
            host.set_row(row)
            host.draw(host_eff, h_frame if (counter[0] & 1) == 0 else h_frame2)
            counter[0] += 1
    host.set_pipelined(True)
    e2e_pipelined = timed_e2e(pass_pipelined, after=host.flush)
    host.set_pipelined(False)

    # per-effect e2e: the synchronous drop-in call, median of 10 after 2 discards
    e2e_per_effect = {}
    if rank == 0:
        for label, eff, host_eff, close, row, params, t in cases:
            host.set_row(row)
            ms = timed_median(lambda: host.draw(host_eff, h_frame), reps=10, discard=2)
            e2e_per_effect[label] = {"ms": ms, "fps": 1e3 / ms}
    # per pass: 12 parameter structs + the ball / twister per-frame tables go up, 12 finished frames come down
    h2d_bytes = 12 * 96 + 2 * (4096 * 4 + RES_Y * 8) + (1024 * 4 + RES_Y * 8)
    d2h_bytes = len(SUITE) * frame_bytes

    # ---- per-kernel roofline: CUDA events around every launch in an instrumented repeat of the timed passes -------
    roofline, kernels = None, {}
    per_effect = {}
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    if rank == 0:
        flops, addmul, flop_src = flop_table()
        prof_passes = 5
        torch.cuda.synchronize()
        ctx.profile_begin()
        for i in range(prof_passes):
            for j, (label, eff, host_eff, close, row, params, t) in enumerate(cases):
                ctx.draw(eff, params, t, d_dest=d_frames[0][(i + j) & 1], close=close)
        stats = ctx.profile_end()
        total_ms = sum(s["total_ms"] for s in stats.values()) or 1.0
        fx_pixels = (RES_X // 2 + 4) * (RES_Y // 2 + 4)
        fp32_peak = 148 * 128 * sm_max_mhz * 1e6 / 1e12  # FADD/FMUL issue rate without FMA contraction, TFLOP/s
        traffic = {}
        tpath = os.path.join(REPO, "profiles", "ncu_traffic.json")
        if os.path.isfile(tpath):
            with open(tpath) as f:
                traffic = json.load(f)
        for name, s in sorted(stats.items(), key=lambda kv: -kv[1]["total_ms"]):
            avg_ms = s["total_ms"] / s["launches"]
            entry = {"launches_per_pass": s["launches"] / prof_passes, "avg_ms": avg_ms, "share": s["total_ms"] / total_ms}
            if name in flops:
                tf = flops[name] * fx_pixels / (avg_ms * 1e-3) / 1e12
                entry.update({"bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak})
                if name in addmul:
                    entry["frac_fadd_fmul_only"] = addmul[name] * fx_pixels / (avg_ms * 1e-3) / 1e12 / fp32_peak
            else:
                gbs = s["algo_bytes"] / s["launches"] / (avg_ms * 1e-3) / 1e9
                entry.update({"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak})
            entry["traffic"] = traffic.get(name)
            entry["limiter"] = limiter_of(name)
            kernels[name] = entry
        # The suite's dominant kernel is raymarch_kernel<Effect>: one __global__ template, several instantiations that the
        # profiler names separately.  The headline roofline is that kernel -- algorithmic FLOPs of all its launches over their
        # summed time -- unless a single other kernel outweighs the family.
        fam = [k for k in kernels if k.startswith("raymarch_") and k != "raymarch_tunnel"]
        fam_share = sum(kernels[k]["share"] for k in fam)
        dominant = max(kernels, key=lambda k: kernels[k]["share"])
        if fam and fam_share > kernels[dominant]["share"]:
            fam_ms = sum(kernels[k]["avg_ms"] * kernels[k]["launches_per_pass"] for k in fam)
            fam_flop = sum(flops[k] * fx_pixels * kernels[k]["launches_per_pass"] for k in fam)
            fam_addmul = sum(addmul.get(k, 0.0) * fx_pixels * kernels[k]["launches_per_pass"] for k in fam)
            tf = fam_flop / (fam_ms * 1e-3) / 1e12
            n_launch = sum(kernels[k]["launches_per_pass"] for k in fam)
            roofline = {"kernel": "raymarch_kernel<Effect> (" + ", ".join(k[len("raymarch_"):] for k in fam) + ")",
                        "launches_per_pass": n_launch, "avg_ms": fam_ms / n_launch,
                        "share": fam_share, "bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak,
                        "frac_fadd_fmul_only": fam_addmul / (fam_ms * 1e-3) / 1e12 / fp32_peak,
                        "traffic": sum((kernels[k]["traffic"] or 0.0) for k in fam) / len(fam), "limiter": limiter_of("raymarch"),
                        "flop_source": flop_src,
                        "peak_note": "FADD/FMUL issue rate without FMA contraction (bit parity forbids FMA): 148 SM x 128 lanes x clock"}
        else:
            roofline = dict(kernels[dominant], kernel=dominant)
        roofline.update(peak_source=peak_src, timing="CUDA events around every launch, instrumented single-stream repeat of the timed passes")
        # the HBM-bound kernel with the largest share, reported next to the dominant one
        hbm_kernels = [k for k in kernels if kernels[k]["bound"] == "hbm"]
        if hbm_kernels:
            top_hbm = max(hbm_kernels, key=lambda k: kernels[k]["share"])
            roofline["dominant_hbm_kernel"] = dict(kernels[top_hbm], kernel=top_hbm)

        # per-effect device time (one frame each, median of 10 after 2 discards)
        for label, eff, host_eff, close, row, params, t in cases:
            ts = []
            for i in range(12):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ctx.draw(eff, params, t, d_dest=d_frames[0][0], close=close)
                b.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ts.append(a.elapsed_time(b))
            ms = float(np.median(ts))
            per_effect[label] = {"ms": ms, "fps": 1e3 / ms, "mpixel_s": RES_X * RES_Y / ms / 1e3,
                                 "e2e_ms": e2e_per_effect[label]["ms"], "e2e_fps": e2e_per_effect[label]["fps"]}

    for c in ctxs[1:]:
        c.close()
    for row_frames in d_frames[:1]:
        for d in row_frames:
            ctx.free(d)

    # ---- BASELINE config 5: the timeline with the gather (all ranks) ---------------------------------------------
    timeline = None
    if not args.no_timeline:
        timeline = timeline_leg(host, ctx, dist, rank, world, args.frames, args.timeline_passes, with_cpu, args.timeline_lanes)

    # ---- BASELINE config 4: the post chain (rank 0) ---------------------------------------------------------------
    post_chain = post_chain_leg(ctx, with_cpu, hbm_peak) if rank == 0 and not args.no_post_chain else None

    # ---- parity + CPU baseline beside it (rank 0, N = 1 only): the reference itself on the host cores ---------------
    parity, cpu_baseline = None, None
    if with_cpu:
        parity, cpu_baseline = parity_and_cpu_leg(host, ctx)
        if per_effect and cpu_baseline.get("per_effect_ms"):
            for label, ms in cpu_baseline["per_effect_ms"].items():
                per_effect[label]["cpu_ms"] = ms
                per_effect[label]["cpu_fps"] = 1e3 / ms

    ctx.free_host(h_frame)
    ctx.free_host(h_frame2)
    host.close()

    # ---- BASELINE configs 1-2 at 1280x720 (rank 0) ---------------------------------------------------------------
    config1 = leg_720p(local, with_cpu) if rank == 0 and not args.no_720p else None

    if rank == 0:
        value = world * PIXELS_PER_PASS * passes_per_step * args.steps / (elapsed_ms * 1e-3) / 1e6
        working_set_mb = (12 * 2 * RES_X * RES_Y * 4 + 2 * RES_X * RES_Y * 8 + 60e6) / 1e6
        line = {
            "metric": "Mpixel/s", "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "ms_per_suite_pass": elapsed_ms / args.steps / passes_per_step,
            "fps_per_effect_mean": 1e3 * len(SUITE) * passes_per_step / (elapsed_ms / args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
            "config": suite_config(),
            "run": {"suite_passes_per_step": passes_per_step, "rows": [s[4] for s in SUITE], "streams": n_streams,
                    "assets": "procedural stand-ins" if assets.synthetic else "reference art (refdata/assets.npz)",
                    "l2": f"no explicit flush: one suite pass streams ~{working_set_mb:.0f} MB (frames, render targets, polar maps, textures) through the 126 MB L2",
                    "timed_region_s": elapsed_ms * 1e-3},
            "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d_bytes * passes_per_step, "d2h_bytes_per_step": d2h_bytes * passes_per_step,
                    "suite_passes_timed": e2e_passes,
                    "api": "X_Draw(uint32_t *pDest, float time, float delta) of include/ckd_host.h, pinned host pDest, synchronous (drop-in semantics)",
                    "readback": "automatic (CkdHost_SetReadbackBands(-1)): frames that end in raymarch + Fx_Blit_2x2 or in a polar remap without a whole-frame post chain render those stages in 4 row bands and every finished band is copied while the next one renders; the other frames are copied whole",
                    "unbanded_value": e2e_unbanded,
                    "pipelined_value": e2e_pipelined,
                    "pipelined_note": "same calls with CkdHost_SetPipelined(true): two device frame buffers, copy stream; pDest valid after CkdHost_Flush()"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": kernels, "per_effect": per_effect,
            "cpu_baseline": cpu_baseline, "parity": parity,
            "timeline": timeline, "post_chain_4k": post_chain, "config1_720p": config1,
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------------
# --workload timeline-4k: config 5 as the headline of its own line (the default line carries it as `timeline`)
# ---------------------------------------------------------------------------------------------------------------

def timeline_reference(args):
    """reference arm of the timeline workload: the reference's own Demo_Draw on the host cores, on a bounded sample"""
    from cookiedough_b200 import sharding
    base = {"impl": "reference", "metric": "Mpixel/s", "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
            "config": {"workload": "timeline-4k", "frames": args.frames, "res": [RES_X, RES_Y]}}
    R = reference_at(RES_Y, demo=True)
    if R is None:
        base["unavailable"] = "oracle/_ref (compiled reference) is not present in this checkout"
        emit(base)
        return 0
    threads = use_all_host_threads()
    times = sharding.timeline_times(args.frames)
    stride = max(1, args.frames // 40)
    sample = list(range(0, args.frames, stride))          # ~40 frames spread over every part
    out = R.frame()

    def step():
        for i in sample:
            R.set_time(times[i])
            R.demo_draw(out)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = len(sample) * RES_X * RES_Y * args.steps / dt / 1e6
    base.update({"value": value, "fps": len(sample) * args.steps / dt, "ms_per_step": 1e3 * dt / args.steps, "gpu_launches": 0,
                 "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": threads, "kind": "reference", "cpu_model": cpu_model(),
                                  "sample": f"every {stride}th frame of the {args.frames}-frame timeline ({len(sample)} frames per step) through the reference's Demo_Draw at {RES_X}x{RES_Y}"},
                 "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(base)
    return 0


def run_timeline(args):
    """BASELINE config 5 as its own line: a step = one pass over the `--frames`-frame timeline (strong scaling)"""
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        return timeline_reference(args)
    rank, world, local, dist = dist_setup(args.gpus)
    import torch
    from cookiedough_b200 import hostapi
    from cookiedough_b200.assets import Assets
    torch.cuda.set_device(local)
    assets = Assets(RES_X, RES_Y)
    host = hostapi.Host(RES_X, RES_Y, local, assets, demo=True)
    ctx = host.context()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    rec = timeline_leg(host, ctx, dist, rank, world, args.frames, max(1, args.steps), rank == 0 and world == 1 and not args.no_cpu_baseline, args.timeline_lanes)
    if rank == 0:
        line = {"metric": "Mpixel/s", "value": rec["value"], "unit": "Mpixel/s", "n_gpus": world, "steps": max(1, args.steps), "warmup": 1,
                "ms_per_step": rec["ms_per_pass"], "fps": rec["gathered_to_rank0_fps"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32+u8", "data": "synthetic", "config": {"workload": "timeline-4k", "frames": args.frames, "res": [RES_X, RES_Y]},
                "e2e": dict(rec["e2e"], h2d_bytes_per_step=0, d2h_bytes_per_step=rec["e2e"]["d2h_bytes_per_pass"]),
                "gpu_launches": rec["gpu_launches"], "timeline": rec, "cpu_baseline": rec.get("cpu_baseline")}
        emit(line)
    host.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_RESULT_OUT = None


def claim_stdout():
    """stdout owes the driver exactly one JSON line.  Libraries write there too (NCCL's version banner at N > 1 comes out on fd 1
    whatever NCCL_DEBUG_FILE says), so the real stdout is set aside for emit() and fd 1 -- for this process and everything it
    loads -- becomes stderr."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(record):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(record) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true", help="run only the device-resident leg (target for ncu launch lists)")
    ap.add_argument("--no-timeline", action="store_true")
    ap.add_argument("--no-post-chain", action="store_true")
    ap.add_argument("--no-720p", action="store_true")
    ap.add_argument("--suite-passes", type=int, default=SUITE_PASSES, help="suite passes (12 frames each) per timed step")
    ap.add_argument("--streams", type=int, default=4, help="CUDA streams (contexts) the device-resident leg spreads the 12 independent frames of a pass over")
    ap.add_argument("--workload", default="effect-suite-4k", choices=["effect-suite-4k", "timeline-4k"],
                    help="timeline-4k: the 600-frame directors-cut timeline through Demo_Draw (BASELINE config 5) as the headline of its own line")
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--timeline-lanes", type=int, default=2, help="contexts per GPU the timeline alternates its frames between (1 or 2)")
    ap.add_argument("--timeline-passes", type=int, default=2, help="timed passes over the timeline in the default line's `timeline` record")
    args = ap.parse_args()
    if args.workload == "timeline-4k":
        return run_timeline(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

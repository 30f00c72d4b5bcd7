#!/usr/bin/env python3
"""bench.py -- headline benchmark of cookiedough_b200 (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload "effect-suite-4k": one STEP renders one 3840x2160 frame through each of the 12 effect entry points of the
hot path (BASELINE.json configs 1-3 at their pinned Rocket rows: the 7 raymarch variants, landscape, tunnelscape,
ball with and without beams, twister), each including its own post chain (Fx_Blit_2x2, polar remap, in-place box blur,
blends) exactly as the reference's X_Draw does.  Metric: Mpixel/s of finished output pixels (12 x 8.2944 Mpx per step).

  value   device-resident: parameters evaluated, maps resident in HBM, frames stay on the GPU; CUDA-event timed.
  e2e     the same step through the reference-facing C++ host layer: Rocket evaluation on the host, X_Draw(pDest, time,
          delta) into a pinned HOST buffer, i.e. every frame is copied device->host inside the timed region.
  N > 1   frames shard across GPUs (one process per GPU, no data-path collective): weak scaling, every rank renders
          K steps; value = total pixels / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

RES_X, RES_Y = 3840, 2160
ROW_RATE = (170.0 / (60.0 * (170.0 / 174.0))) * 16.0

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
PIXELS_PER_STEP = len(SUITE) * RES_X * RES_Y

# (The counts were taken on the exact-lookup kernels, i.e. they are the reference's own operations -- lutcosf with its two
#  conversions included; the conversion-free lookup executes one FP instruction more per call, which is NOT credited.)
# FP operations per FX-map pixel at the pinned rows (SURVEY.md 8d asks for the measured mean, not the <= bound): MEASURED on
# the B200 as executed thread-level FP instructions of one 4K launch of each kernel (tools/count_fp_ops.py ->
# profiles/r01_fp_ops.json).  The kernels execute the reference's float operations one for one (no FMA contraction), so this
# is the algorithmic count; 1 op = one FP add / mul / compare / min-max / MUFU / conversion, FP64 ops of powf/expf likewise.
# FFMA instructions are NOT counted (they only occur in the IEEE division / sqrt refinement sequences that stand for one
# divss / sqrtss of the reference, whose MUFU seed is the one op counted), which makes the fractions conservative.
FLOP_PER_FX_PIXEL = {
    "raymarch_plasma": 1461.0, "raymarch_nautilus": 1814.0, "raymarch_spikey_close": 1673.0, "raymarch_spikey_distant": 1688.0,
    "raymarch_sinuses": 2490.0, "raymarch_laura": 1494.0, "raymarch_tunnel": 268.0,
}


# what ncu says bounds each kernel (profiles/r01_notes.md): reported next to the roofline fraction so that a small HBM
# fraction of a kernel that is not HBM bound is not misread
LIMITER = {
    "raymarch": "instruction issue: FMUL/FADD chains without FMA contraction (bit parity) + 4 non-FP instructions per LUT lookup; the XU pipe is out of the picture since the conversion-free lookup (spikey kernels: exact lookup, XU 70 %)",
    "old_blur": "integer ALU pipe + dependent chain of the saturating in-place recurrence (ALU 57 %, issue 62 %); DRAM traffic = algorithmic bytes",
    "voxel": "L2 gather latency of the height/colour map samples (warp per ray)",
    "polar_blit": "dependent map -> texel gather chain; DRAM traffic = algorithmic bytes",
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
        for line in self.lines:
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
    # rank 0 must print exactly one JSON line on stdout: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
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


# ---------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own CPU implementation (oracle/_ref) on the host cores
# ---------------------------------------------------------------------------------------------------------------

def reference_suite_runner():
    from cookiedough_b200.assets import Assets
    from oracle import ref as oref
    if not oref.available(RES_Y):
        return None, None
    R = oref.Reference.get(RES_Y, Assets(RES_X, RES_Y))
    out = R.frame()

    def step():
        for _, _, ref_eff, _, row in SUITE:
            R.set_row(row)
            R.draw(ref_eff, out)
    return R, step


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_host_threads():
    """The reference parallelises with OpenMP.  torchrun exports OMP_NUM_THREADS=1 to every rank, which would time the reference
    on one core: give it every core this process may run on (omp_set_num_threads on the already loaded runtime)."""
    n = cpu_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    use_all_host_threads()
    R, step = reference_suite_runner()
    base = {"impl": "reference", "metric": "Mpixel/s", "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
            "config": {"workload": "effect-suite-4k", "res": [RES_X, RES_Y], "effects": [s[0] for s in SUITE]}}
    if step is None:
        base["unavailable"] = "oracle/_ref (compiled reference) is not present in this checkout"
        print(json.dumps(base))
        return 0
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = PIXELS_PER_STEP * args.steps / dt / 1e6
    sample = f"{args.steps} full steps (12 frames at {RES_X}x{RES_Y} each) after {max(args.warmup, 1)} warm-up steps"
    base.update({"value": value, "ms_per_step": 1e3 * dt / args.steps, "gpu_launches": 0,
                 "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": cpu_threads(), "kind": "reference", "sample": sample},
                 "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))
    return 0


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
    host = hostapi.Host(RES_X, RES_Y, local, assets)
    ctx = host.context()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    # The 12 frames of a step are independent: the device-resident leg renders them on --streams CUDA streams (one context =
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

    def step_device(i):
        for j, (label, eff, host_eff, close, row, params, t) in enumerate(cases):
            k = j % n_streams
            ctxs[k].draw(eff, params, t, d_dest=d_frames[k][(i + j // n_streams) & 1], close=close)

    def join_streams():
        for st in side_streams:
            torch.cuda.current_stream().wait_stream(st)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    for i in range(args.warmup):
        step_device(i)
    barrier()

    # clocks: nvidia-smi needs ~1 s to come up, the timed region may be shorter: start it now, keep the identical load
    # running until it has produced samples, then time the K steps with the sampler still attached
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        t_load = time.perf_counter()
        i = 0
        while time.perf_counter() - t_load < 2.5 and len(sampler.lines) < 8:
            step_device(i)
            i += 1
            if i % 8 == 0:
                torch.cuda.synchronize()
    barrier()
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
    if rank == 0:
        # keep the same load up briefly so the 100 ms sampler certainly sees the region's clocks
        t_load = time.perf_counter()
        while time.perf_counter() - t_load < 0.4:
            step_device(0)
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        dist.barrier()

    # ---- end to end through the reference-facing host API (HOST pDest, D2H inside the timed region) -------------
    frame_bytes = RES_X * RES_Y * 4
    h_frame = ctx.malloc_host(frame_bytes)
    e2e_steps = max(1, min(args.steps, 10))

    def step_e2e():
        for label, eff, host_eff, close, row, params, t in cases:
            host.set_row(row)
            host.draw(host_eff, h_frame)

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * PIXELS_PER_STEP * e2e_steps / e2e_s / 1e6

    # the same synchronous calls with the banded read-back switched off (one copy after the frame is finished)
    host.set_readback_bands(0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_plain_s = time.perf_counter() - t0
    host.set_readback_bands(-1)
    if dist is not None:
        t = torch.tensor([e2e_plain_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_plain_s = float(t.item())
    e2e_unbanded = world * PIXELS_PER_STEP * e2e_steps / e2e_plain_s / 1e6

    # same calls with the host layer's two-deep frame pipeline (CkdHost_SetPipelined): frame i's copy overlaps frame i+1's render
    h_frame2 = ctx.malloc_host(frame_bytes)
    host.set_pipelined(True)
    barrier()
    t0 = time.perf_counter()
    k = 0
    for _ in range(e2e_steps):
        for label, eff, host_eff, close, row, params, t in cases:
            host.set_row(row)
            host.draw(host_eff, h_frame if (k & 1) == 0 else h_frame2)
            k += 1
    host.flush()
    e2e_pipe_s = time.perf_counter() - t0
    host.set_pipelined(False)
    if dist is not None:
        tp = torch.tensor([e2e_pipe_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        e2e_pipe_s = float(tp.item())
    e2e_pipelined = world * PIXELS_PER_STEP * e2e_steps / e2e_pipe_s / 1e6
    # per step: 12 parameter structs + the ball / twister per-frame tables go up, 12 finished frames come down
    h2d_bytes = 12 * 96 + 2 * (4096 * 4 + RES_Y * 8) + (1024 * 4 + RES_Y * 8)
    d2h_bytes = len(SUITE) * frame_bytes

    # ---- per-kernel roofline: CUDA events around every launch in an instrumented repeat of the timed steps -------
    roofline, kernels = None, {}
    per_effect = {}
    if rank == 0:
        hbm_peak, sm_max_mhz, peak_src = measured_peaks()
        prof_steps = max(1, min(args.steps, 5))
        torch.cuda.synchronize()
        ctx.profile_begin()
        for i in range(prof_steps):
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
            entry = {"launches_per_step": s["launches"] / prof_steps, "avg_ms": avg_ms, "share": s["total_ms"] / total_ms}
            if name in FLOP_PER_FX_PIXEL:
                tf = FLOP_PER_FX_PIXEL[name] * fx_pixels / (avg_ms * 1e-3) / 1e12
                entry.update({"bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak})
            else:
                gbs = s["algo_bytes"] / s["launches"] / (avg_ms * 1e-3) / 1e9
                entry.update({"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak})
            entry["traffic"] = traffic.get(name)
            entry["limiter"] = limiter_of(name)
            kernels[name] = entry
        # The step's dominant kernel is raymarch_kernel<Effect>: one __global__ template, seven instantiations that the
        # profiler names separately (52 % of the step together, 7-11 % each).  The headline roofline is that kernel --
        # algorithmic FLOPs of all its launches over their summed time -- unless a single other kernel outweighs the family.
        fam = [k for k in kernels if k.startswith("raymarch_") and k != "raymarch_tunnel"]
        fam_share = sum(kernels[k]["share"] for k in fam)
        dominant = max(kernels, key=lambda k: kernels[k]["share"])
        if fam and fam_share > kernels[dominant]["share"]:
            fam_ms = sum(kernels[k]["avg_ms"] * kernels[k]["launches_per_step"] for k in fam)
            fam_flop = sum(FLOP_PER_FX_PIXEL[k] * fx_pixels * kernels[k]["launches_per_step"] for k in fam)
            tf = fam_flop / (fam_ms * 1e-3) / 1e12
            roofline = {"kernel": "raymarch_kernel<Effect> (" + ", ".join(k[len("raymarch_"):] for k in fam) + ")",
                        "launches_per_step": sum(kernels[k]["launches_per_step"] for k in fam), "avg_ms": fam_ms / sum(kernels[k]["launches_per_step"] for k in fam),
                        "share": fam_share, "bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak,
                        "traffic": sum((kernels[k]["traffic"] or 0.0) for k in fam) / len(fam), "limiter": limiter_of("raymarch"),
                        "peak_note": "FADD/FMUL issue rate without FMA contraction (bit parity forbids FMA): 148 SM x 128 lanes x clock"}
        else:
            roofline = dict(kernels[dominant], kernel=dominant)
        roofline.update(peak_source=peak_src, timing="CUDA events around every launch, instrumented single-stream repeat of the timed steps")
        # the HBM-bound kernel with the largest share, reported next to the dominant one
        hbm_kernels = [k for k in kernels if kernels[k]["bound"] == "hbm"]
        if hbm_kernels:
            top_hbm = max(hbm_kernels, key=lambda k: kernels[k]["share"])
            roofline["dominant_hbm_kernel"] = dict(kernels[top_hbm], kernel=top_hbm)

        # per-effect device time (one frame each, median of 3)
        for label, eff, host_eff, close, row, params, t in cases:
            ts = []
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ctx.draw(eff, params, t, d_dest=d_frames[0][0], close=close)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = float(np.median(ts))
            per_effect[label] = {"ms": ms, "fps": 1e3 / ms, "mpixel_s": RES_X * RES_Y / ms / 1e3}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference itself on the host cores ---------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        use_all_host_threads()
        R, ref_step = reference_suite_runner()
        if ref_step is not None:
            ref_step()
            passes = 2
            t0 = time.perf_counter()
            for _ in range(passes):
                ref_step()
            dt = time.perf_counter() - t0
            cpu_baseline = {"value": PIXELS_PER_STEP * passes / dt / 1e6, "unit": "Mpixel/s", "cores": cpu_threads(), "kind": "reference",
                            "sample": f"{passes} full steps of the same suite (12 frames at {RES_X}x{RES_Y}) on oracle/_ref after 1 warm-up step, OpenMP on all host threads"}
        else:
            cpu_baseline = {"value": None, "unit": "Mpixel/s", "cores": cpu_threads(), "kind": "reference", "sample": "oracle/_ref not present"}

    if rank == 0:
        value = world * PIXELS_PER_STEP * args.steps / (elapsed_ms * 1e-3) / 1e6
        working_set_mb = (12 * 2 * RES_X * RES_Y * 4 + 2 * RES_X * RES_Y * 8 + 60e6) / 1e6
        line = {
            "metric": "Mpixel/s", "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "fps_per_effect_mean": 1e3 * len(SUITE) / (elapsed_ms / args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
            "config": {"workload": "effect-suite-4k", "res": [RES_X, RES_Y], "effects": [s[0] for s in SUITE], "rows": [s[4] for s in SUITE],
                       "streams": n_streams,
                       "assets": "procedural stand-ins" if assets.synthetic else "reference art (refdata/assets.npz)",
                       "l2": f"no explicit flush: one step streams ~{working_set_mb:.0f} MB (frames, render targets, polar maps, textures) through the 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                    "api": "X_Draw(uint32_t *pDest, float time, float delta) of include/ckd_host.h, pinned host pDest, synchronous (drop-in semantics)",
                    "readback": "automatic (CkdHost_SetReadbackBands(-1)): frames that end in raymarch + Fx_Blit_2x2 or in a polar remap without a whole-frame post chain render those stages in 4 row bands and every finished band is copied while the next one renders; the other frames are copied whole",
                    "unbanded_value": e2e_unbanded,
                    "pipelined_value": e2e_pipelined,
                    "pipelined_note": "same calls with CkdHost_SetPipelined(true): two device frame buffers, copy stream; pDest valid after CkdHost_Flush()"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels, "per_effect": per_effect,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    for c in ctxs[1:]:
        c.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def timeline_reference(args):
    """reference arm of the timeline workload: the reference's own Demo_Draw on the host cores, on a bounded sample"""
    from cookiedough_b200 import sharding
    from cookiedough_b200.assets import Assets
    from oracle import ref as oref
    base = {"impl": "reference", "metric": "Mpixel/s", "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
            "config": {"workload": "timeline-4k", "frames": args.frames, "res": [RES_X, RES_Y]}}
    if not oref.available(RES_Y):
        base["unavailable"] = "oracle/_ref (compiled reference) is not present in this checkout"
        print(json.dumps(base))
        return 0
    use_all_host_threads()
    R = oref.Reference.get(RES_Y, Assets(RES_X, RES_Y), demo=True)
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
                 "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": cpu_threads(), "kind": "reference",
                                  "sample": f"every {stride}th frame of the {args.frames}-frame timeline ({len(sample)} frames per step) through the reference's Demo_Draw at {RES_X}x{RES_Y}"},
                 "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))
    return 0


def run_timeline(args):
    """BASELINE config 5: the directors-cut timeline through Demo_Draw (code/demo.cpp:469-1023) -- the part's effect plus its
    layers, composed on the device -- at 3840x2160.  Frames shard by index (frame i -> rank i mod N, no data-path collective);
    the total work is fixed, so this line reports strong scaling.  'value': frames stay on the device; 'e2e': every frame is
    copied to a pinned host buffer inside the timed region (pipelined_value: with the host layer's two-deep frame pipeline)."""
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        return timeline_reference(args)
    rank, world, local, dist = dist_setup(args.gpus)
    import torch
    from cookiedough_b200 import hostapi, sharding
    from cookiedough_b200.assets import Assets

    torch.cuda.set_device(local)
    assets = Assets(RES_X, RES_Y)
    host = hostapi.Host(RES_X, RES_Y, local, assets, demo=True)
    ctx = host.context()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    times = sharding.timeline_times(args.frames)
    mine = sharding.frames_for_rank(args.frames, rank, world)
    frame_bytes = RES_X * RES_Y * 4
    h_frames = [ctx.malloc_host(frame_bytes) for _ in range(2)]

    def pass_device():
        for i in mine:
            host.demo_draw(0, times[i])          # pDest == nullptr: the composed frame stays on the device

    def pass_e2e():
        for k, i in enumerate(mine):
            host.demo_draw(h_frames[k & 1], times[i])

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    for _ in range(max(1, min(args.warmup, 2))):
        pass_device()
    barrier()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        pass_device()
    ev1.record()
    torch.cuda.synchronize()
    ms = sharding.reduce_max(dist, ev0.elapsed_time(ev1), device="cuda")
    launches = ctx.launch_count() - launches0

    barrier()
    t0 = time.perf_counter()
    pass_e2e()
    torch.cuda.synchronize()
    e2e_s = sharding.reduce_max(dist, time.perf_counter() - t0, device="cuda")

    host.set_pipelined(True)
    barrier()
    t0 = time.perf_counter()
    pass_e2e()
    host.flush()
    e2e_pipe_s = sharding.reduce_max(dist, time.perf_counter() - t0, device="cuda")
    host.set_pipelined(False)

    # a checksum of checksums over the whole timeline: the same on any number of ranks (frames are pure functions of time)
    local_sums = {}
    frame = np.zeros((RES_Y, RES_X), dtype=np.uint32)
    for i in mine[::max(1, len(mine) // 8)]:
        host.demo_draw(frame, times[i])
        local_sums[i] = sharding.frame_checksum(frame)
    sums = sharding.gather_checksums(dist, local_sums, args.frames, device="cuda")

    if rank == 0:
        px = args.frames * RES_X * RES_Y
        print(json.dumps({
            "metric": "Mpixel/s", "value": px * args.steps / (ms * 1e-3) / 1e6, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "fps": args.frames * args.steps / (ms * 1e-3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
            "config": {"workload": "timeline-4k", "frames": args.frames, "res": [RES_X, RES_Y], "api": "Demo_Draw (effect + the part's layers, composed on the device)",
                       "sharding": "frame i -> rank i mod N, no data-path collective",
                       "assets": "procedural stand-ins" if assets.synthetic else "reference art (refdata/assets.npz), layers nearest-upscaled x3"},
            "e2e": {"value": px / e2e_s / 1e6, "unit": "Mpixel/s", "fps": args.frames / e2e_s, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": args.frames * frame_bytes, "pipelined_value": px / e2e_pipe_s / 1e6, "pipelined_fps": args.frames / e2e_pipe_s},
            "gpu_launches": int(launches), "frame_checksums_crc32": {str(i): c for i, c in enumerate(sums) if c}}))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=4, help="CUDA streams (contexts) the device-resident leg spreads the 12 independent frames of a step over")
    ap.add_argument("--workload", default="effect-suite-4k", choices=["effect-suite-4k", "timeline-4k"],
                    help="timeline-4k: the 600-frame directors-cut timeline through Demo_Draw (BASELINE config 5), frame i -> rank i mod N (strong scaling)")
    ap.add_argument("--frames", type=int, default=600)
    args = ap.parse_args()
    if args.workload == "timeline-4k":
        return run_timeline(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python3
"""Parity + timing report: every effect and post op, CUDA path (C ABI) vs the compiled reference (oracle/_ref).

    python tests/tools/parity_report.py [--res 720|2160|both] [--json out.json]

Needs a B200 (CUDA side) and oracle/_ref (reference side).  Prints one line per case:
exact-pixel %, max channel delta, % within 1/2 LSB, GPU ms (CUDA events, median of 5) and CPU ms.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

from cookiedough_b200 import capi  # noqa: E402
from cookiedough_b200.assets import Assets  # noqa: E402
from oracle import ref as oref  # noqa: E402


def compare(a, b):
    a8 = a.view(np.uint8).reshape(-1, 4).astype(np.int16)
    b8 = b.view(np.uint8).reshape(-1, 4).astype(np.int16)
    d = np.abs(a8 - b8).max(axis=1)
    n = d.size
    return {
        "exact_pct": 100.0 * float((d == 0).sum()) / n,
        "le1_pct": 100.0 * float((d <= 1).sum()) / n,
        "le2_pct": 100.0 * float((d <= 2).sum()) / n,
        "max_delta": int(d.max()),
    }


def gpu_time(ctx, fn, reps=5):
    fn()
    ctx.sync()
    ts = []
    for _ in range(reps):
        ctx.timer_start()
        fn()
        ts.append(ctx.timer_stop_ms())
    return float(np.median(ts))


def cpu_time(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts))


EFFECT_CASES = [
    # name, ckd effect, reference effect, row, extra
    ("plasma", "plasma", "plasma", 2600, {}),
    ("nautilus", "nautilus", "nautilus", 5700, {}),
    ("spikey_close", "spikey", "spikey_close", 6800, {"close": True}),
    ("spikey_distant", "spikey", "spikey_distant", 3600, {"close": False}),
    ("tunnel", "tunnel", "tunnel", 4500, {}),
    ("sinuses", "sinuses", "sinuses", 7800, {}),
    ("laura", "laura", "laura", 8900, {}),
    ("landscape", "landscape", "landscape", 500, {}),
    ("tunnelscape", "tunnelscape", "tunnelscape", 4300, {}),
    ("ball", "ball", "ball", 1500, {}),
    ("ball_beams", "ball", "ball", 2060, {}),
    ("twister", "twister", "twister", 2008, {}),
]


def run_res(res_y, results, quick=False):
    res_x = res_y * 16 // 9
    assets = Assets(res_x, res_y)
    R = oref.Reference.get(res_y, assets)
    ctx = capi.Context(res_x, res_y, 0, assets)
    tab, log2 = ctx.rsqrt_table()
    print(f"== {res_x}x{res_y}  rsqrt table: {tab.size} entries (log2 bin {log2}); assets synthetic={assets.synthetic}")

    n = res_x * res_y
    for name, eff, ref_eff, row, extra in EFFECT_CASES:
        R.set_row(row)
        t = R.time
        params = capi.params_from_tracks(eff, R.track)
        # identical, known content in the render target both sides start from (ball beams keeps stale pixels)
        seed = (np.arange(n, dtype=np.uint32) * np.uint32(2654435761)).reshape(res_y, res_x)
        R.render_target(0)[:] = seed
        ctx.upload(ctx.render_target(0), seed)

        ref_out = R.draw(ref_eff).copy()
        ctx.draw(eff, params, t, close=extra.get("close"))
        out = ctx.read_frame()
        stats = compare(out, ref_out)

        def redraw():
            ctx.draw(eff, params, t, close=extra.get("close"))
        stats["gpu_ms"] = gpu_time(ctx, redraw)
        stats["cpu_ms"] = cpu_time(lambda: R.draw(ref_eff)) if not quick else float("nan")
        stats["gpu_fps"] = 1e3 / stats["gpu_ms"]
        stats["mpix_s"] = n / stats["gpu_ms"] / 1e3
        results[f"{name}@{res_y}"] = stats
        print(f"{name:16s} row {row:5d}  exact {stats['exact_pct']:8.4f}%  <=1 {stats['le1_pct']:8.4f}%  <=2 {stats['le2_pct']:8.4f}%  max {stats['max_delta']:3d}"
              f"  gpu {stats['gpu_ms']:8.3f} ms ({stats['gpu_fps']:8.1f} fps)  cpu {stats['cpu_ms']:8.2f} ms")
        sys.stdout.flush()

    # ---- post chain on seeded synthetic buffers ---------------------------------------
    rng = np.random.default_rng(1234)
    src = rng.integers(0, 2**32, size=n, dtype=np.uint32).reshape(res_y, res_x)
    dst = rng.integers(0, 2**32, size=n, dtype=np.uint32).reshape(res_y, res_x)
    fx = rng.integers(0, 2**32, size=R.fx_x * R.fx_y, dtype=np.uint32).reshape(R.fx_y, R.fx_x)

    d_src = ctx.to_device(src, pad_elems=res_x * 2)
    d_dst = ctx.to_device(dst, pad_elems=res_x * 2)
    d_fx = ctx.to_device(fx, pad_elems=res_x * 2)

    def ref_buf(a):
        b = oref.aligned_u32(a.size, pad=res_x * 2).reshape(a.shape)
        b[:] = a
        return b

    def post_case(label, ref_fn, gpu_fn, bytes_per_px, in_place_input=None):
        r_dst = ref_buf(dst)
        r_src = ref_buf(src)
        ref_fn(r_dst, r_src)
        ctx.upload(d_dst, dst)
        ctx.upload(d_src, src)
        gpu_fn(d_dst, d_src)
        out = ctx.download(d_dst, (res_y, res_x))
        stats = compare(out, r_dst)
        stats["gpu_ms"] = gpu_time(ctx, lambda: gpu_fn(d_dst, d_src))
        stats["cpu_ms"] = cpu_time(lambda: ref_fn(r_dst, r_src)) if not quick else float("nan")
        stats["gbs"] = bytes_per_px * n / stats["gpu_ms"] / 1e6
        results[f"{label}@{res_y}"] = stats
        print(f"{label:28s} exact {stats['exact_pct']:8.4f}%  max {stats['max_delta']:3d}  gpu {stats['gpu_ms']:8.4f} ms  {stats['gbs']:8.1f} GB/s (algorithmic)  cpu {stats['cpu_ms']:8.3f} ms")
        sys.stdout.flush()

    r_fx = ref_buf(fx)
    post_case("Fx_Blit_2x2", lambda d, s: R.fx_blit_2x2(d, r_fx), lambda d, s: ctx.fx_blit_2x2(d, d_fx), 5)
    post_case("Polar_Blit", lambda d, s: R.polar_blit(d, s, False), lambda d, s: ctx.polar_blit(d, s, False), 16)
    post_case("Polar_Blit(inverse)", lambda d, s: R.polar_blit(d, s, True), lambda d, s: ctx.polar_blit(d, s, True), 16)
    post_case("Polar_BlitA", lambda d, s: R.polar_blit(d, s, False, alpha=True), lambda d, s: ctx.polar_blit(d, s, False, alpha=True), 20)
    for strength in (0.01, 0.11, 0.33, 1.0):
        post_case(f"HBlur32 inplace s={strength}", lambda d, s: R.old_blur("h", d, d, res_x, res_y, strength), lambda d, s: ctx.old_blur("h", d, d, res_x, res_y, strength), 8)
        post_case(f"VBlur32 inplace s={strength}", lambda d, s: R.old_blur("v", d, d, res_x, res_y, strength), lambda d, s: ctx.old_blur("v", d, d, res_x, res_y, strength), 8)
    post_case("BoxBlur32 inplace s=0.11", lambda d, s: R.old_blur("hv", d, d, res_x, res_y, 0.11), lambda d, s: ctx.old_blur("hv", d, d, res_x, res_y, 0.11), 16)
    post_case("BoxBlur32 out-of-place s=0.11", lambda d, s: R.old_blur("hv", d, s, res_x, res_y, 0.11), lambda d, s: ctx.old_blur("hv", d, s, res_x, res_y, 0.11), 16)
    post_case("BoxBlur_Horz32 1 pass", lambda d, s: R.new_blur("h", d, s, res_x, res_y, 6.28, 0.1, 1), lambda d, s: ctx.new_blur("h", d, s, res_x, res_y, 6.28, 0.1, 1), 8)
    post_case("BoxBlur_Vert32 2 pass", lambda d, s: R.new_blur("v", d, s, res_x, res_y, 6.28, 0.1, 2), lambda d, s: ctx.new_blur("v", d, s, res_x, res_y, 6.28, 0.1, 2), 16)
    post_case("BoxBlur_32 3 pass (kGauss)", lambda d, s: R.new_blur("hv", d, s, res_x, res_y, 6.28, 0.1, 3), lambda d, s: ctx.new_blur("hv", d, s, res_x, res_y, 6.28, 0.1, 3), 48)
    post_case("TapeWarp32", lambda d, s: R.tape_warp(d, s, res_x, res_y, 0.5, 0.33), lambda d, s: ctx.tape_warp(d, s, res_x, res_y, 0.5, 0.33), 8)
    for op in oref.BLEND_OPS:
        fparam, uparam = 0.0, 0
        if op == "Mix32":
            uparam = 77
        if op == "SoftLight32AA":
            fparam = 0.37
        if op == "Fade32":
            uparam = (200 << 24) | 0x123456
        post_case(op, lambda d, s: R.blend(op, d, s, fparam, uparam), lambda d, s: ctx.blend(op, d, s, n, fparam, uparam), 8 if op == "Fade32" else 12)
    for op in oref.BLIT_OPS:
        post_case(op, lambda d, s: R.blit(op, d, s, res_x, res_x, res_y, 0.6), lambda d, s: ctx.blit(op, d, s, res_x, res_x, res_y, 0.6), 12)

    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", default="both")
    ap.add_argument("--json", default=None)
    ap.add_argument("--quick", action="store_true", help="skip CPU timing")
    args = ap.parse_args()
    results = {}
    for res_y in ((720, 2160) if args.res == "both" else (int(args.res),)):
        run_res(res_y, results, quick=args.quick)
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        with open(args.json, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()

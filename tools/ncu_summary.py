#!/usr/bin/env python3
"""Summarise ncu captures of bench.py for profiles/.

    python tools/ncu_summary.py launches <launches.csv>                 -> markdown table of launch shares (stdout)
    python tools/ncu_summary.py full <capture.ncu-rep> <summary.txt> <traffic.json>
                                                                        -> per-kernel extract + DRAM bytes per launch keyed by
                                                                           bench.py's kernel names (read back into roofline.traffic)
"""
import collections
import csv
import json
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "lts__t_sector_hit_rate.pct"]


def bench_name(kernel):
    """kernel function name -> the name bench.py's per-kernel profiler uses"""
    m = re.search(r"raymarch_kernel<.*?(\w+)Effect[,>]", kernel)   # raymarch_kernel<Effect, FAST>
    if m:
        name = re.sub(r"(?<!^)(?=[A-Z])", "_", m.group(1)).lower()
        return "raymarch_" + name
    if "tunnel_kernel" in kernel and "scape" not in kernel:
        return "raymarch_tunnel"
    for key, name in (("landscape_kernel", "voxel_landscape"), ("tunnelscape_kernel", "voxel_tunnelscape"), ("twister_kernel", "voxel_twister"),
                      ("fx_blit_2x2", "fx_blit_2x2"), ("blend_kernel", "blend"), ("rect_kernel", "rect_blit"), ("memset32", "memset32"),
                      ("tape_warp", "tape_warp"), ("transpose32", "transpose32"), ("new_blur", "new_blur")):
        if key in kernel:
            return name
    if "ball_kernel" in kernel:
        return "voxel_ball_beams" if re.search(r"ball_kernel<\(bool\)1>|ball_kernel<true>|ball_kernel<1>", kernel) else "voxel_ball"
    if "polar_blit" in kernel:
        if re.search(r"<\(bool\)1, \(bool\)1[,>]|<true, true[,>]|<1, 1[,>]", kernel):
            return "polar_blit_a_halo"
        return "polar_blit_a" if re.search(r"<\(bool\)1[,>]|<true[,>]|<1[,>]", kernel) else "polar_blit"
    if "old_blur" in kernel:
        vert = re.search(r"old_blur_\w+_kernel<\(bool\)(\d)|old_blur_\w+_kernel<(\d)", kernel)
        flag = (vert.group(1) or vert.group(2)) if vert else "0"
        return "old_blur_v" if flag == "1" else "old_blur_h"
    return kernel


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    stats = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        us = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        us *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1.0)
        k = r[ix["Kernel Name"]]
        n, t = stats.get(k, (0, 0.0))
        stats[k] = (n + 1, t + us)
    total = sum(t for _, t in stats.values())
    print("| kernel | bench.py name | launches | avg µs | share |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(stats.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:70]}` | {bench_name(k)} | {n} | {t/n:.1f} | {100*t/total:.1f} % |")
    by_name = collections.defaultdict(float)
    for k, (n, t) in stats.items():
        by_name[bench_name(k)] += t
    print("\nshares by bench.py name: " + ", ".join(f"{k} {100*v/total:.1f} %" for k, v in sorted(by_name.items(), key=lambda kv: -kv[1])))


def full(rep, summary_path, traffic_path):
    # `rep` is the capture itself or its raw page exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv > x_raw.csv`: the
    # reports embed the whole module and do not fit gpurun_out/ once there are several)
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen, traffic, count = collections.OrderedDict(), collections.defaultdict(float), collections.Counter()
    for r in rows[2:]:
        k = r[ix["Kernel Name"]]
        name = bench_name(k)
        def val(metric):
            v = float(r[ix[metric]].replace(",", ""))
            u = units[ix[metric]]
            return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
        traffic[name] += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        count[name] += 1
        if k not in seen:
            seen[k] = [(m, r[ix[m]], units[ix[m]]) for m in METRICS if m in ix]
    with open(summary_path, "w") as f:
        for k, vals in seen.items():
            f.write(f"{k}   [{bench_name(k)}]\n")
            for m, v, u in vals:
                f.write(f"    {m:75s} {v} {u}\n")
    with open(traffic_path, "w") as f:
        json.dump({name: traffic[name] / count[name] for name in traffic}, f, indent=1)
    print(f"{len(seen)} kernels -> {summary_path}, {traffic_path}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4])

#!/usr/bin/env python3
"""Measured FP operations per FX-map pixel of every raymarch kernel (the denominators of bench.py's FP32 roofline).

    ncu --metrics <METRICS> -k regex:'raymarch_kernel|tunnel_kernel' --csv --log-file ops.csv python tools/effect_all.py
    python tools/count_fp_ops.py ops.csv profiles/r01_fp_ops.json

The kernels execute the reference's floating-point operations one for one (no FMA contraction, same LUT scheme), so the
executed thread-level FP instruction count at the pinned rows is the algorithmic count SURVEY.md 8d asks for:
1 op = one FP add / mul / compare / min-max / MUFU / conversion, FP64 operations of powf/expf likewise (a DFMA once).
FFMA is reported but NOT counted: with -fmad=false it only occurs inside the IEEE division / sqrt refinement sequences, each
of which stands for ONE divss / sqrtss of the reference -- the MUFU seed of the sequence is the op that is counted."""
import collections, csv, json, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from tools.ncu_summary import bench_name

METRICS = ["smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
           "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum",
           "smsp__sass_thread_inst_executed_op_conversion_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
           "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]


def main(csv_path, out_path, res_x=3840, res_y=2160):
    rows = [r for r in csv.reader(open(csv_path, errors="replace")) if r]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    ix = {h: i for i, h in enumerate(rows[start])}
    per = collections.defaultdict(dict)
    for r in rows[start + 1:]:
        if len(r) != len(rows[start]):
            continue
        per[(r[ix["ID"]], r[ix["Kernel Name"]])][r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    fx_pixels = (res_x // 2 + 4) * (res_y // 2 + 4)
    out = {}
    for (_, kernel), m in per.items():
        g = lambda k: m.get(f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum", 0.0)
        fp32_other = g("fp32") - g("fadd") - g("fmul") - g("ffma")      # compares, min/max, MUFU ...
        ops = g("fadd") + g("fmul") + fp32_other + g("conversion") + g("dadd") + g("dmul") + g("dfma")
        out[bench_name(kernel)] = {"ops_per_fx_pixel": ops/fx_pixels, "fadd": g("fadd")/fx_pixels, "fmul": g("fmul")/fx_pixels, "ffma": g("ffma")/fx_pixels,
                                   "fp32_other": fp32_other/fx_pixels, "conversion": g("conversion")/fx_pixels,
                                   "fp64": (g("dadd") + g("dmul") + g("dfma"))/fx_pixels}
    with open(out_path, "w") as f:
        json.dump({"res": [res_x, res_y], "fx_pixels": fx_pixels, "rows": "pinned rows of bench.py SUITE", "kernels": out}, f, indent=1)
    for k, v in out.items():
        print(f"{k:26s} {v['ops_per_fx_pixel']:8.1f} ops/px  (fadd {v['fadd']:.0f} fmul {v['fmul']:.0f} ffma {v['ffma']:.0f} other {v['fp32_other']:.0f} cvt {v['conversion']:.0f} fp64 {v['fp64']:.0f})")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

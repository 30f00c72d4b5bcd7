#!/usr/bin/env python3
"""print the headline numbers and the per-kernel table of a bench.py JSON line"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "pipelined", round(d["e2e"].get("pipelined_value", 0), 1),
      "launches", d["gpu_launches"], "clocks", d["clocks"])
print("roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"].items() if k != "dominant_hbm_kernel"})
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["share"]):
    print(f"{k:26s} n={v['launches_per_step']:4.1f} avg={v['avg_ms']*1e3:7.1f}us share={v['share']*100:5.1f}% {v['bound']:5s} frac={v['frac']:.3f}")
print({k: round(v["ms"], 3) for k, v in d["per_effect"].items()})

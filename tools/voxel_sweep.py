#!/usr/bin/env python3
"""times the five voxel-caster frames of the suite at 4K (device time of the caster kernel alone, CUDA events inside the library)
-- run under different CKD_SHORT_SPAN values to tune the span emission threshold"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cookiedough_b200 import capi, hostapi
from cookiedough_b200.assets import Assets
host = hostapi.Host(bench.RES_X, bench.RES_Y, 0, Assets(bench.RES_X, bench.RES_Y))
ctx = host.context()
out = []
for label, eff, host_eff, close, row in bench.SUITE:
    if label not in bench.INTEGER_LABELS:
        continue
    host.set_row(row)
    params = capi.params_from_tracks(eff, host.track)
    t = float(np.float32(host.time))
    for _ in range(2):
        ctx.draw(eff, params, t, close=close)
    ctx.sync()
    ctx.profile_begin()
    for _ in range(5):
        ctx.draw(eff, params, t, close=close)
    stats = ctx.profile_end()
    vox = [v["total_ms"] / v["launches"] for k, v in stats.items() if k.startswith("voxel_")]
    out.append(f"{label} {vox[0]*1e3:.1f}")
print(f"short_span {os.environ.get('CKD_SHORT_SPAN', 'default(6)'):>10}: " + "  ".join(out), flush=True)
host.close()

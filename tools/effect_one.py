#!/usr/bin/env python3
"""a few launches of one effect at 4K at its pinned row -- a target for ncu (python tools/effect_one.py sinuses)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cookiedough_b200 import capi, hostapi
from cookiedough_b200.assets import Assets
want = sys.argv[1] if len(sys.argv) > 1 else "sinuses"
host = hostapi.Host(bench.RES_X, bench.RES_Y, 0, Assets(bench.RES_X, bench.RES_Y))
ctx = host.context()
for label, eff, host_eff, close, row in bench.SUITE:
    if label != want:
        continue
    host.set_row(row)
    params = capi.params_from_tracks(eff, host.track)
    for _ in range(3):
        ctx.draw(eff, params, float(np.float32(host.time)), close=close)
    ctx.sync()
host.close()

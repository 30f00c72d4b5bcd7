#!/usr/bin/env python3
"""one launch of every suite effect at 4K at its pinned row -- a target for ncu metric collection"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cookiedough_b200 import capi, hostapi
from cookiedough_b200.assets import Assets
host = hostapi.Host(bench.RES_X, bench.RES_Y, 0, Assets(bench.RES_X, bench.RES_Y))
ctx = host.context()
for label, eff, host_eff, close, row in bench.SUITE:
    host.set_row(row)
    ctx.draw(eff, capi.params_from_tracks(eff, host.track), float(np.float32(host.time)), close=close)
ctx.sync()
host.close()

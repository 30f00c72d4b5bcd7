#!/usr/bin/env python3
"""a few old-blur launches of one (direction, strength) given on the command line -- a target for ncu"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cookiedough_b200 import capi
res_x, res_y = 3840, 2160
ctx = capi.Context(res_x, res_y, 0)
n = res_x*res_y
src = (np.arange(n, dtype=np.uint32)*np.uint32(2654435761))
d_a = ctx.to_device(src, pad_elems=4*res_x)
kind = sys.argv[1] if len(sys.argv) > 1 else "h"
strength = float(sys.argv[2]) if len(sys.argv) > 2 else 0.11
for _ in range(3):
    ctx.old_blur(kind, d_a, d_a, res_x, res_y, strength)
ctx.sync()

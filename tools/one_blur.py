import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from cookiedough_b200 import capi
import post_cases as pc
ctx = capi.Context(3840, 2160, 0)
n = 3840 * 2160
d_a = ctx.to_device(pc.seeded(n, "noise"), pad_elems=4 * 3840)
for _ in range(3):
    ctx.old_blur("h", d_a, d_a, 3840, 2160, 0.11)
    ctx.old_blur("v", d_a, d_a, 3840, 2160, 0.11)
ctx.sync()

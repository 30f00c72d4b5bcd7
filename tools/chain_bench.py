#!/usr/bin/env python3
"""fused blend chain vs the same blends one launch at a time (4K, CUDA events inside the library)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cookiedough_b200 import capi
res_x, res_y = 3840, 2160
n = res_x*res_y
ctx = capi.Context(res_x, res_y, 0)
rng = np.random.default_rng(3)
layers = [ctx.to_device(rng.integers(0, 2**32, size=n, dtype=np.uint32)) for _ in range(6)]
dest = ctx.to_device(rng.integers(0, 2**32, size=n, dtype=np.uint32))
CHAINS = {
    "part 1 (Fade, SoftLight32A, MulSrc32A)": ["Fade32", "SoftLight32A", "MulSrc32A"],
    "part 4 (Sub, MixSrc, Overlay, Fade)": ["Sub32", "MixSrc32", "Overlay32", "Fade32"],
    "part 6 (SoftLight x2, Fade, Overlay32A, MixSrc, MixSrc)": ["SoftLight32", "SoftLight32", "Fade32", "Overlay32A", "MixSrc32", "MixSrc32"],
    "part 8 (Fade, SoftLight, Sub, Excl, MulSrc32A, MixOver, Overlay)": ["Fade32", "SoftLight32", "Sub32", "Excl32", "MulSrc32A", "MixOver32", "Overlay32"],
}
for label, ops in CHAINS.items():
    steps = [(op, None if op == "Fade32" else layers[i % 6], 0.0, (128 << 24) if op == "Fade32" else 0) for i, op in enumerate(ops)]
    def seq():
        for op, src, f, u in steps:
            ctx.blend(op, dest, src or dest, n, f, u)
    def fused():
        ctx.blend_chain(dest, steps, n)
    res = []
    for fn in (seq, fused):
        fn(); ctx.sync()
        ctx.profile_begin()
        for _ in range(5):
            fn()
        stats = ctx.profile_end()
        res.append(sum(v["total_ms"] for v in stats.values())/5*1e3)
    print(f"{label:70s} sequential {res[0]:7.1f} us   fused {res[1]:7.1f} us")
ctx.close()

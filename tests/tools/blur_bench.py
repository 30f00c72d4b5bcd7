#!/usr/bin/env python3
"""micro-benchmark + parity of the box blurs against the compiled reference (needs a GPU and oracle/_ref)"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from cookiedough_b200 import capi
from cookiedough_b200.assets import Assets
from oracle import ref as oref
import post_cases as pc

STRENGTHS = tuple(float(a) for a in sys.argv[1:]) or (0.01, 0.02, 0.05, 0.11, 0.33, 1.0)
for res_y in (720, 2160):
    res_x = res_y * 16 // 9
    assets = Assets(res_x, res_y)
    R = oref.Reference.get(res_y, assets)
    ctx = capi.Context(res_x, res_y, 0)
    n = res_x * res_y
    src = pc.seeded(n, "noise")
    d_a = ctx.to_device(src, pad_elems=4 * res_x)
    d_b = ctx.to_device(src, pad_elems=4 * res_x)
    for kind in ("h", "v", "hv"):
        for strength in STRENGTHS:
            for inplace in (True, False):
                ra = oref.aligned_u32(n, pad=4 * res_x); ra[:] = src
                rb = oref.aligned_u32(n, pad=4 * res_x); rb[:] = src
                R.old_blur(kind, ra, ra if inplace else rb, res_x, res_y, strength)
                ctx.upload(d_a, src); ctx.upload(d_b, src)
                ctx.old_blur(kind, d_a, d_a if inplace else d_b, res_x, res_y, strength)
                out = ctx.download(d_a, (n,))
                ok = np.array_equal(out, ra)
                ctx.profile_begin()  # CUDA events around every launch, inside the library: no Python in the timed interval
                for _ in range(5):
                    ctx.old_blur(kind, d_a, d_a if inplace else d_b, res_x, res_y, strength)
                stats = ctx.profile_end()
                us = sum(v["total_ms"] for v in stats.values())/5*1e3
                print(f"{res_x}x{res_y} old_blur_{kind:2s} s={strength:<5} K={max(1, min(255, int(strength*255 + 0.5))):3d} {'inplace' if inplace else 'copy   '} {'OK  ' if ok else 'FAIL'} {us:8.1f} us")
    ctx.close()

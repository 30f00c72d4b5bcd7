#!/usr/bin/env python3
"""times Polar_Blit(inverse) and Polar_BlitA at 4K (CUDA events inside the library, 5 rotating buffer sets) and checks them
against the compiled reference -- run under different CKD_POLAR_VARIANT values to compare kernel variants"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from cookiedough_b200 import capi
from cookiedough_b200.assets import Assets
from oracle import ref as oref
import post_cases as pc
w, h = 3840, 2160
n = w*h
ctx = capi.Context(w, h, 0)
src, dst = pc.seeded(n, "mul"), pc.seeded(n, "mix")
sets = 5
d_src = [ctx.to_device(src, pad_elems=4*w) for _ in range(sets)]
d_dst = [ctx.to_device(dst, pad_elems=4*w) for _ in range(sets)]
R = oref.Reference.get(h, Assets(w, h)) if oref.available(h) else None
for label, inverse, alpha in (("polar_blit(inverse)", True, False), ("polar_blit", False, False), ("polar_blit_a", False, True)):
    ok = "n/a"
    if R is not None:
        rs = oref.aligned_u32(n, pad=4*w).reshape(h, w); rs[:] = src.reshape(h, w)
        rd = oref.aligned_u32(n, pad=4*w).reshape(h, w); rd[:] = dst.reshape(h, w)
        R.polar_blit(rd, rs, inverse, alpha=alpha)
        ctx.upload(d_dst[0], dst)
        ctx.polar_blit(d_dst[0], d_src[0], inverse, alpha=alpha)
        ok = "OK" if np.array_equal(ctx.download(d_dst[0], (n,)), rd.reshape(-1)) else "FAIL"
    for k in range(sets):
        ctx.polar_blit(d_dst[k], d_src[k], inverse, alpha=alpha)
    ctx.sync()
    ctx.profile_begin()
    for r in range(20):
        ctx.polar_blit(d_dst[r % sets], d_src[r % sets], inverse, alpha=alpha)
    stats = ctx.profile_end()
    us = sum(v["total_ms"] for v in stats.values())/20*1e3
    print(f"variant {os.environ.get('CKD_POLAR_VARIANT', '0'):>2} {label:20s} {us:7.1f} us  {ok}", flush=True)
ctx.close()

#!/usr/bin/env python3
"""4K in-place sweep of the old box blur over kernel widths: time (CUDA events inside the library) and parity against the
compiled reference, on three images: noise, a smooth gradient, and bright blocks on black (the recurrence's undershoot after a
bright -> black edge drives the accumulator into its clamp at zero: the exact path of the blocked kernel).
    python tests/tools/blur_sweep.py [K ...]"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from cookiedough_b200 import capi
from cookiedough_b200.assets import Assets
from oracle import ref as oref
import post_cases as pc

KS = tuple(int(a) for a in sys.argv[1:]) or (3, 5, 9, 15, 17, 26, 31, 42, 51, 84, 128, 255)
res_x, res_y = 3840, 2160
n = res_x * res_y
R = oref.Reference.get(res_y, Assets(res_x, res_y)) if oref.available(res_y) else None
ctx = capi.Context(res_x, res_y, 0)
yy, xx = np.mgrid[0:res_y, 0:res_x]
images = {
    "noise": pc.seeded(n, "noise"),
    "blocks": np.where(((xx // 97 + yy // 61) % 3) == 0, np.uint32(0xffffffff), np.uint32(0)).astype(np.uint32).reshape(-1),
    "gradient": ((xx * 255 // res_x) | ((yy * 255 // res_y) << 8) | (((xx + yy) * 255 // (res_x + res_y)) << 16) | (0xff << 24)).astype(np.uint32).reshape(-1),
}
if os.environ.get("CKD_SWEEP_IMAGES"):
    images = {k: v for k, v in images.items() if k in os.environ["CKD_SWEEP_IMAGES"].split(",")}
d_a = ctx.to_device(images["noise"], pad_elems=4 * res_x)
print(f"{'K':>4} {'img':>9} {'h us':>8} {'v us':>8}  parity")
for K in KS:
    strength = (K + 0.25) / 255.0
    for name, src in images.items():
        us, oks = {}, []
        for kind in ("h", "v"):
            if R is not None:
                ra = oref.aligned_u32(n, pad=4 * res_x); ra[:] = src
                R.old_blur(kind, ra, ra, res_x, res_y, strength)
            ctx.upload(d_a, src)
            ctx.old_blur(kind, d_a, d_a, res_x, res_y, strength)
            out = ctx.download(d_a, (n,))
            oks.append("OK" if R is None or np.array_equal(out, ra) else "FAIL")
            ts = []
            for _ in range(5):
                ctx.upload(d_a, src)
                ctx.sync()
                ctx.profile_begin()
                ctx.old_blur(kind, d_a, d_a, res_x, res_y, strength)
                stats = ctx.profile_end()
                ts.append(sum(v["total_ms"] for v in stats.values()) * 1e3)
            us[kind] = float(np.median(ts))
        print(f"{K:4d} {name:>9} {us['h']:8.1f} {us['v']:8.1f}  {'/'.join(oks)}", flush=True)
ctx.close()

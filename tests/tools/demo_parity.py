#!/usr/bin/env python3
"""Demo_Draw parity: the composed frame of every part (CUDA host layer) against the compiled reference's Demo_Draw.

    python tests/tools/demo_parity.py [--res 720|2160] [--rows r0,r1,...] [--frames N]
"""
import argparse, os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from cookiedough_b200 import hostapi
from cookiedough_b200.assets import Assets
from oracle import ref as oref

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=720)
ap.add_argument("--rows", default="")
ap.add_argument("--frames", type=int, default=60)
args = ap.parse_args()
res_y = args.res; res_x = res_y*16//9
assets = Assets(res_x, res_y)
R = oref.Reference.get(res_y, assets, demo=True)
H = hostapi.Host(res_x, res_y, 0, Assets(res_x, res_y), demo=True)
rows = [float(r) for r in args.rows.split(",") if r] or [i*10296.0/args.frames for i in range(args.frames)]
out = np.zeros((res_y, res_x), dtype=np.uint32)
worst = 100.0
for row in rows:
    t = row/oref.ROW_RATE
    R.set_time(t)
    t0 = time.perf_counter(); ref = R.demo_draw(); t_ref = time.perf_counter() - t0
    t0 = time.perf_counter(); H.demo_draw(out, t); t_gpu = time.perf_counter() - t0
    a = out.view(np.uint8).reshape(-1, 4).astype(np.int16); b = ref.view(np.uint8).reshape(-1, 4).astype(np.int16)
    dlt = np.abs(a - b).max(axis=1)
    exact = 100.0*float((dlt == 0).mean()); worst = min(worst, exact)
    print(f"row {row:8.1f} part {int(R.track('demo:Effect')):2d}  exact {exact:8.4f}%  <=2 {100.0*float((dlt <= 2).mean()):8.4f}%  max {int(dlt.max()):3d}   gpu {t_gpu*1e3:7.2f} ms  cpu {t_ref*1e3:7.1f} ms")
    sys.stdout.flush()
print("worst exact %", worst)
H.close()

#!/usr/bin/env python3
"""How much does rendering independent frames on K streams (K contexts) help the device-resident suite throughput?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cookiedough_b200 import capi, hostapi
from cookiedough_b200.assets import Assets

assets = Assets(bench.RES_X, bench.RES_Y)
host = hostapi.Host(bench.RES_X, bench.RES_Y, 0, assets)
cases = []
for label, eff, host_eff, close, row in bench.SUITE:
    host.set_row(row)
    cases.append((label, eff, close, capi.params_from_tracks(eff, host.track), float(np.float32(host.time))))
# single-stream device time per effect (us, 4K) -> longest-processing-time-first assignment to K streams
COST = {"nautilus": 446, "ball_beams": 310, "sinuses": 284, "ball": 244, "twister": 198, "spikey_distant": 195, "spikey_close": 190,
        "plasma": 178, "laura": 176, "tunnel": 160, "tunnelscape": 100, "landscape": 98}


def lpt(K):
    load, where = [0.0]*K, {}
    for label in sorted(COST, key=lambda l: -COST[l]):
        k = min(range(K), key=lambda i: load[i])
        where[label] = k
        load[k] += COST[label]
    return where


for K, balanced in ((1, False), (2, False), (2, True), (4, False), (4, True), (6, True)):
    ctxs = [capi.Context(bench.RES_X, bench.RES_Y, 0, assets) for _ in range(K)]
    streams = [torch.cuda.Stream() for _ in range(K)]
    for c, s in zip(ctxs, streams):
        c.set_stream(s.cuda_stream)
    where = lpt(K)
    def step():
        for j, (label, eff, close, params, t) in enumerate(cases):
            ctxs[where[label] if balanced else j % K].draw(eff, params, t, close=close)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps = 30
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(f"K={K} {'LPT' if balanced else 'round-robin'}: {dt*1e3:.3f} ms/step  {bench.PIXELS_PER_STEP/dt/1e6:.0f} Mpixel/s")
    for c in ctxs:
        c.close()
host.close()

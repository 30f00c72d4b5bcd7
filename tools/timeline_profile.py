#!/usr/bin/env python3
"""per-kernel time over the whole 600-frame timeline (Demo_Draw at 4K), CUDA events around every launch"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cookiedough_b200 import hostapi, sharding
from cookiedough_b200.assets import Assets
res_x, res_y = 3840, 2160
host = hostapi.Host(res_x, res_y, 0, Assets(res_x, res_y), demo=True)
ctx = host.context()
times = sharding.timeline_times(600)
for t in times[::50]:
    host.demo_draw(0, t)
ctx.sync()
ctx.profile_begin()
for t in times:
    host.demo_draw(0, t)
stats = ctx.profile_end()
total = sum(v["total_ms"] for v in stats.values())
print(f"total {total:.1f} ms for 600 frames = {total/600*1e3:.1f} us/frame")
for k, v in sorted(stats.items(), key=lambda kv: -kv[1]["total_ms"]):
    print(f"{k:26s} launches {v['launches']:6d}  avg {v['total_ms']/v['launches']*1e3:8.1f} us  share {100*v['total_ms']/total:5.1f} %")
host.close()

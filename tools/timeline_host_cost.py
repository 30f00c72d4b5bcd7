#!/usr/bin/env python3
"""how much of a timeline pass is host time: wall clock until CkdTimeline_Render returns (everything enqueued) against wall
clock until the GPU has finished, without a gather, for 1 and 2 lanes"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cookiedough_b200 import hostapi, sharding
from cookiedough_b200.assets import Assets
host = hostapi.Host(3840, 2160, 0, Assets(3840, 2160), demo=True)
ctx = host.context()
times = sharding.timeline_times(600)
for lanes in (1, 2):
    host.timeline_render(times, lanes=lanes)
    ctx.sync()
    t0 = time.perf_counter()
    host.timeline_render(times, lanes=lanes)
    t1 = time.perf_counter()
    ctx.sync()
    t2 = time.perf_counter()
    print(f"lanes {lanes}: enqueue {1e3*(t1-t0):.1f} ms ({1e6*(t1-t0)/600:.0f} us/frame), finished {1e3*(t2-t0):.1f} ms ({600/(t2-t0):.0f} fps)", flush=True)
host.close()

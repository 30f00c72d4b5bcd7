#!/usr/bin/env python3
"""a few Demo_Draw frames at 4K at one Rocket row -- a target for ncu (python tools/demo_one.py 2600)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cookiedough_b200 import hostapi, sharding
from cookiedough_b200.assets import Assets
row = float(sys.argv[1]) if len(sys.argv) > 1 else 2600.0
host = hostapi.Host(3840, 2160, 0, Assets(3840, 2160), demo=True)
ctx = host.context()
for _ in range(3):
    host.demo_draw(0, row / sharding.ROW_RATE)
ctx.sync()
host.close()

"""oracle/rocket.py -- TEST INFRASTRUCTURE: GNU Rocket track evaluation restated in Python.

sync_get_val (3rdparty/rocket-stripped/lib/track.c:9-60): key values are floats, the interpolation runs in double;
Rocket::geti is int(roundf(getf)) (code/rocket.h:27-29).  Keys come from tests/golden/tracks.json (all keys of
target/directors-cut.rocket) or any {name: [[row, value, interpolation], ...]} mapping."""
import json
import math

import numpy as np

ROW_RATE = (170.0 / (60.0 * (170.0 / 174.0))) * 16.0   # code/audio.cpp:18


def sync_get_val(keys, row):
    """track.c:32-60"""
    if not keys:
        return 0.0
    irow = math.floor(row)
    idx = -1
    for i, k in enumerate(keys):
        if k[0] <= irow:
            idx = i
    if idx < 0:
        return float(np.float32(keys[0][1]))
    if idx > len(keys) - 2:
        return float(np.float32(keys[-1][1]))
    r0, v0, t0 = keys[idx]
    r1, v1, _ = keys[idx + 1]
    v0, v1 = np.float32(v0), np.float32(v1)
    t = (row - r0) / (r1 - r0)
    if t0 == 0:
        return float(v0)
    if t0 == 2:
        t = t * t * (3 - 2 * t)
    elif t0 == 3:
        t = math.pow(t, 2.0)
    return float(v0) + float(np.float32(v1 - v0)) * t


class Tracks:
    def __init__(self, tracks):
        self.tracks = tracks
        self.row = 0.0

    @classmethod
    def from_json(cls, path):
        with open(path) as f:
            return cls(json.load(f)["tracks"])

    def set_time(self, seconds):
        self.row = float(seconds) * ROW_RATE   # Audio_Rocket_Sync, code/audio.cpp:175-178

    def get(self, name):
        return sync_get_val(self.tracks.get(name, []), self.row)

    def getf(self, name):
        return float(np.float32(self.get(name)))

    def geti(self, name):
        v = np.float32(self.get(name))
        return int(math.floor(abs(float(v)) + 0.5) * (1 if v >= 0 else -1))   # roundf: half away from zero

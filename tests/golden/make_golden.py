#!/usr/bin/env python3
"""Generates the committed golden fixtures from the compiled reference (oracle/_ref, see oracle/build_ref.py).

    python tests/golden/make_golden.py            # writes tests/golden/*.json, *.npy

The reference ships no tests, golden vectors or known-answer fixtures for this path, so the
pins are outputs of the reference itself, run here:

  golden_effects_720.json   every effect entry point at 1280x720: pinned Rocket rows of the real timeline
                            ("timeline") and hand-built parameter sets that reach the branches the timeline misses
                            ("scenario"), all rendered with the deterministic *synthetic* assets of
                            cookiedough_b200/assets.py so the fixtures work from a bare checkout.
                            Per case: the evaluated parameter struct, sha256 of the frame, and a 48x6 pixel crop.
  golden_post_720.json      2D post chain ops on seeded buffers: sha256 of the result.
  rsqrt_table_golden.npy    the 2x1024-entry RSQRTPS table of the CPU that generated the goldens (the float effects
                            depend on it; tests install it with ckd_set_rsqrt_table before comparing).
  golden_demo_720.json      Demo_Draw (effect + the part's layers) at rows of the real timeline that reach every part and
                            every optional layer; synthetic assets; sha256 + crop per frame.
  tracks.json               all keys of target/directors-cut.rocket (row, value, interpolation) for the host-side
                            Rocket tests.

Each mode runs in its own process because the reference keeps its state in globals.
"""
import hashlib
import json
import os
import platform
import struct
import subprocess
import sys
import tempfile
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

TIMELINE_CASES = [
    # label, ckd effect, reference effect, close flag, rows
    ("plasma", "plasma", "plasma", None, [2600, 2200, 3100]),
    ("nautilus", "nautilus", "nautilus", None, [5700, 5300, 6210]),
    ("spikey_close", "spikey", "spikey_close", True, [6800, 6400, 7100, 7300]),
    ("spikey_distant", "spikey", "spikey_distant", False, [3600, 3200, 4100]),
    ("tunnel", "tunnel", "tunnel", None, [4500, 5100, 5236]),
    ("sinuses", "sinuses", "sinuses", None, [7800, 7400, 8300]),
    ("laura", "laura", "laura", None, [8900, 8500, 9300]),
    ("landscape", "landscape", "landscape", None, [500, 100, 1000, 1040]),
    ("tunnelscape", "tunnelscape", "tunnelscape", None, [4300, 4190, 4710]),
    ("ball", "ball", "ball", None, [1500, 2060, 1200, 1430, 1800]),
    ("twister", "twister", "twister", None, [2008, 1995]),
]

# scenario = (label, ckd effect, reference effect, close, base row, {track: value})
SCENARIOS = [
    ("spikey_spec_only", "spikey", "spikey_distant", False, 3600, {"distSpike:Warmup": 0.5}),
    ("spikey_spec_only_hot", "spikey", "spikey_distant", False, 3900, {"distSpike:Warmup": 3.0, "spike:Roll": 1.1}),
    ("spikey_close_mixblur", "spikey", "spikey_close", True, 6800, {"closeSpike:MixBlurOpacity": 0.7, "closeSpike:MixBlurMap": 0.4, "closeSpike:MixBlur": 6.0, "closeSpike:MixMapBlur": 3.0}),
    ("spikey_close_map0", "spikey", "spikey_close", True, 6500, {"closeSpike:MixBlurOpacity": 1.0, "closeSpike:MixBlurMap": 0.0, "closeSpike:MixBlur": 0.5, "closeSpike:MixMapBlur": 0.0}),
    ("spikey_close_map1", "spikey", "spikey_close", True, 6900, {"closeSpike:MixBlurOpacity": 0.3, "closeSpike:MixBlurMap": 1.0, "closeSpike:MixBlur": 20.0, "closeSpike:MixMapBlur": 12.0}),
    ("spikey_close_noaspect_rim", "spikey", "spikey_close", True, 6700, {"closeSpike:AspectMul": 0.0, "closeSpike:Rim": 1.0, "closeSpike:MixBlurOpacity": 0.0}),
    ("landscape_warp", "landscape", "landscape", None, 500, {"voxelScape:WarpStrength": 0.33, "voxelScape:WarpSpeed": 0.02, "voxelScape:Tilt": -45.0}),
    ("landscape_tilt_min", "landscape", "landscape", None, 300, {"voxelScape:Tilt": -90.0}),
    ("ball_spikes_blur", "ball", "ball", None, 1500, {"ball:Spikes": 128.0, "ball:BaseShapeIndex": 2.0, "ball:Blur": 10.0, "ball:Radius": 1100.0}),
    ("ball_beams_low", "ball", "ball", None, 2060, {"ball:BallLowBeams": 128.0, "ball:Beams1": 1.0, "ball:Beams2": 0.5, "ball:Beams3": 0.25, "ball:Radius": 1000.0, "ball:BeamAttenuation": 96.0, "ball:Blur": 0.0}),
    ("ball_short_ray", "ball", "ball", None, 1500, {"ball:RayLength": 100.0, "ball:Radius": 700.0, "ball:Blur": 0.0, "ball:BaseShapeIndex": 4.0}),
    ("twister_blur", "twister", "twister", None, 2008, {"twister:Blur": 10.0, "twister::ShearSpeed": 0.7}),
    ("twister_noblur", "twister", "twister", None, 2000, {"twister:Blur": 0.0}),
    ("nautilus_noblur", "nautilus", "nautilus", None, 5700, {"nautilus:Blur": 0.0}),
    ("nautilus_blur5", "nautilus", "nautilus", None, 5900, {"nautilus:Blur": 5.0, "nautilus:Roll": 0.4}),
    ("tunnel_lit_boxy", "tunnel", "tunnel", None, 4500, {"tunnel:LitTiles": 1.0, "tunnel:LitBlur": 3.0, "tunnel:Boxy": 0.5}),
    ("tunnel_unlit", "tunnel", "tunnel", None, 5000, {"tunnel:LitTiles": 0.0}),
    ("tunnelscape_blur", "tunnelscape", "tunnelscape", None, 4300, {"starsTunnel:Blur": 3.0}),
    ("plasma_gamma", "plasma", "plasma", None, 2600, {"plasma:Gamma": 2.2, "plasma:Desaturation": 0.5, "plasma:Hue": 4.0}),
    ("sinuses_spec", "sinuses", "sinuses", None, 7800, {"sinusesTunnel:Specular": 7.5, "sinusesTunnel:Roll": 0.9, "sinusesTunnel:OffsX": 0.25}),
    ("laura_yaw", "laura", "laura", None, 8900, {"laura:Yaw": 0.6, "laura:Pitch": -0.3, "laura:Saturate": 1.3}),
    # LUT angles far outside the table's exact range (scaled angle >= 2^23: the reference's lutcosf aliases there): these
    # frames must take the exact-lookup kernel on the device (LutRangeProof in csrc/ckd_raymarch.cu)
    ("plasma_far", "plasma", "plasma", None, 2600, {"plasma:Speed": 40000.0}),
    ("nautilus_far", "nautilus", "nautilus", None, 5700, {"nautilus:Speed": 300.0, "nautilus:Blur": 0.0}),
    ("sinuses_far", "sinuses", "sinuses", None, 7800, {"sinusesTunnel:Speed": 300.0}),
    ("laura_far", "laura", "laura", None, 8900, {"laura:Speed": 500.0}),
]
SCENARIO_ROW_STEP = 8


def parse_rocket(path):
    tracks = {}
    for tr in ET.parse(path).getroot().iter("track"):
        tracks[tr.attrib["name"]] = [[int(k.attrib["row"]), float(k.attrib["value"]), int(k.attrib["interpolation"])] for k in tr.iter("key")]
    return tracks


def track_file_name(name):
    # path_encode, 3rdparty/rocket-stripped/lib/device.c:41-78
    out = ""
    for ch in name:
        out += ch if (ch.isalnum() and ch.isascii()) or ch in "._/" else "-%02X" % ord(ch)
    return "_" + out + ".track"


def write_track(path, keys):
    # read_track_data, device.c:309-332
    with open(path, "wb") as f:
        f.write(struct.pack("<i", len(keys)))
        for row, value, interp in keys:
            f.write(struct.pack("<ifb", row, value, interp))


def crop_of(frame):
    h, w = frame.shape
    y0, x0 = h // 2 - 3, w // 2 - 24
    return {"x": x0, "y": y0, "w": 48, "h": 6, "hex": frame[y0:y0 + 6, x0:x0 + 48].astype("<u4").tobytes().hex()}


def render_cases(mode, out_path):
    from cookiedough_b200 import capi
    from cookiedough_b200.assets import Assets
    from oracle import ref as oref

    assets = Assets(1280, 720, force_synthetic=True)
    tmp = None
    data_dir = oref.DATA_DIR
    if mode == "scenario":
        # custom sync/ directory: every track gets one step key per scenario row
        tracks = parse_rocket(os.path.join(oref.DATA_DIR, "directors-cut.rocket"))
        tmp = tempfile.TemporaryDirectory(prefix="ckd_golden_")
        data_dir = tmp.name
        os.makedirs(os.path.join(data_dir, "sync"))
        probe = oref.Reference(720, assets)  # real tracks, to read the base values
        names = sorted({t for _, (_, m) in capi.TRACKS.items() for t in m.values()})
        keys = {n: [] for n in names}
        for i, (label, eff, ref_eff, close, base_row, overrides) in enumerate(SCENARIOS):
            probe.set_row(base_row)
            for n in names:
                v = overrides.get(n, probe.track(n))
                keys[n].append((i * SCENARIO_ROW_STEP, float(np.float32(v)), 0))
        for n in names:
            write_track(os.path.join(data_dir, "sync", track_file_name(n)), keys[n])
        write_track(os.path.join(data_dir, "sync", track_file_name("demo:quit")), [(0, 0.0, 0)])
        # the probe instance must not be reused (globals): re-exec in a child with the prepared directory
        env = dict(os.environ, CKD_GOLDEN_DATA_DIR=data_dir)
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", "scenario_child", out_path], env=env)
        tmp.cleanup()
        return

    if mode == "scenario_child":
        data_dir = os.environ["CKD_GOLDEN_DATA_DIR"]

    oref.DATA_DIR = data_dir
    R = oref.Reference(720, assets)
    cases = {}
    if mode == "timeline":
        todo = [(f"{label}@{row}", eff, ref_eff, close, row, R_row) for label, eff, ref_eff, close, rows in TIMELINE_CASES for row in rows for R_row in [row]]
    else:
        todo = [(label, eff, ref_eff, close, base_row, i * SCENARIO_ROW_STEP) for i, (label, eff, ref_eff, close, base_row, _) in enumerate(SCENARIOS)]

    n = R.res_x * R.res_y
    seed = (np.arange(n, dtype=np.uint32) * np.uint32(2654435761)).reshape(R.res_y, R.res_x)
    for label, eff, ref_eff, close, time_row, rocket_row in todo:
        # 'time' follows the base row of the real timeline, the Rocket row selects the parameter set
        R.set_row(rocket_row)
        params = capi.params_from_tracks(eff, R.track)
        time_s = float(np.float32(time_row / oref.ROW_RATE))
        R.time = time_s
        R.render_target(0)[:] = seed
        frame = R.draw(ref_eff)
        cases[label] = {
            "effect": eff, "ref_effect": ref_eff, "close": close, "row": time_row, "time": time_s,
            "params": {name: getattr(params, name) for name, _ in params._fields_},
            "sha256": hashlib.sha256(frame.astype("<u4").tobytes()).hexdigest(),
            "crop": crop_of(frame),
        }
        print(f"  {label:28s} {cases[label]['sha256'][:16]}")
    with open(out_path, "w") as f:
        json.dump(cases, f)


# Demo_Draw (the compositor, code/demo.cpp:469-1023): rows of the real timeline that reach every part and every optional
# layer of a part (found by scanning the tracks: shooting star, credit logo blurs and cross-fades, Cousteau blur, the three
# Moonraker text paths, the 1995/2006 logo fades, ribbons vs. full warp, disco guys with and without strip blur, the GPU joke)
DEMO_ROWS = [
    0, 300, 450, 482, 508, 700, 980, 1000, 1030, 1046,            # part 2: landscape, shooting star + trail, overlay, Revision logo warp/blur
    1104, 1300, 1500, 1712, 1900, 2060, 2092,                     # part 3: ball without / with beams
    1996, 2008, 2020,                                             # part 1: twister
    2100, 2200, 2364, 2510, 2700, 3064, 3108, 3130, 3142,         # part 5: plasma + credit logos (anim blend, H/V blur)
    3250, 3600, 3716, 4000, 4180,                                 # part 8: distant spikes + title logos
    4204, 4246, 4300, 4400, 4720, 4900,                           # part 4: tunnelscape + 1995 logos
    4500, 4996, 5050, 5100, 5200,                                 # part 9: tunnel + 2006 logos
    5300, 5378, 5700, 5906, 6026, 6158, 6284,                     # part 6: nautilus, Cousteau 1/2, blur
    6290, 6418, 6628, 6656, 6684, 6700, 6728, 6778, 6808, 6812, 7074, 7332,  # part 7: close-up spikes, dirt 1/2/3, Moonraker text paths
    7340, 7600, 7882, 8122, 8254, 8380,                           # part 10: sinuses + love prism blur, dirt
    8500, 8900, 9300, 9390,                                       # part 11: greetings
    9410, 9482, 9524, 9548, 9556, 9726, 9794, 9892,               # part 12: ribbons / full warp, blur
    9924, 9950, 9972, 9980, 10142, 10300, 10310, 10354, 10396,    # part 13: disco guys (strip blur), credits, GPU joke
]


def demo_cases(out_path):
    from cookiedough_b200.assets import Assets
    from oracle import ref as oref

    R = oref.Reference.get(720, Assets(1280, 720, force_synthetic=True), demo=True)
    cases = {}
    n = R.res_x * R.res_y
    seed = (np.arange(n, dtype=np.uint32) * np.uint32(2654435761)).reshape(R.res_y, R.res_x)
    for row in DEMO_ROWS:
        time_s = float(np.float32(row / oref.ROW_RATE))
        R.set_time(time_s)
        R.render_target(0)[:] = seed   # the ball's beam path keeps stale render-target pixels: make every frame history-free
        part = int(round(R.track("demo:Effect")))
        frame = R.demo_draw()
        cases[str(row)] = {"row": row, "time": time_s, "part": part, "sha256": hashlib.sha256(frame.astype("<u4").tobytes()).hexdigest(), "crop": crop_of(frame)}
        print(f"  row {row:6d} part {part:2d} {cases[str(row)]['sha256'][:16]}")
    with open(out_path, "w") as f:
        json.dump(cases, f)


def post_cases(out_path):
    from cookiedough_b200.assets import Assets
    from oracle import ref as oref
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import post_cases as pc

    R = oref.Reference(720, Assets(1280, 720, force_synthetic=True))
    out = {}
    for case in pc.CASES:
        result = pc.run_reference(R, case)
        out[case["label"]] = hashlib.sha256(result.astype("<u4").tobytes()).hexdigest()
        print(f"  {case['label']:36s} {out[case['label']][:16]}")
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1)


def main():
    if len(sys.argv) >= 4 and sys.argv[1] == "--child":
        mode, out_path = sys.argv[2], sys.argv[3]
        if mode == "post":
            post_cases(out_path)
        elif mode == "demo":
            demo_cases(out_path)
        else:
            render_cases(mode, out_path)
        return

    from oracle import ref as oref
    if not oref.available(720):
        sys.exit("oracle/_ref is missing: run `python oracle/build_ref.py` first")

    meta = {"generator": "tests/golden/make_golden.py", "cpu": platform.processor() or platform.machine(), "assets": "synthetic (cookiedough_b200/assets.py)", "res": [1280, 720]}
    with open("/proc/cpuinfo") as f:
        for line in f:
            if line.startswith("model name"):
                meta["cpu"] = line.split(":", 1)[1].strip()
                break

    # the compositor's frames live in their own file: `make_golden.py --only demo` refreshes just those
    with tempfile.TemporaryDirectory() as tmp:
        part = os.path.join(tmp, "demo.json")
        print("[demo]")
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", "demo", part])
        with open(part) as f:
            demo = json.load(f)
    with open(os.path.join(HERE, "golden_demo_720.json"), "w") as f:
        json.dump({"meta": meta, "frames": demo}, f, indent=1)
    if sys.argv[1:3] == ["--only", "demo"]:
        return

    if sys.argv[1:3] == ["--only", "effects"]:   # refresh the effect pins alone (new cases); existing pins may not move
        with tempfile.TemporaryDirectory() as tmp:
            parts = {}
            for mode in ("timeline", "scenario"):
                part = os.path.join(tmp, mode + ".json")
                subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", mode, part])
                with open(part) as f:
                    parts[mode] = json.load(f)
        path = os.path.join(HERE, "golden_effects_720.json")
        with open(path) as f:
            golden = json.load(f)
        moved = [f"{mode}/{k}" for mode in parts for k, v in golden[mode].items() if parts[mode].get(k, {}).get("sha256") != v["sha256"]]
        if moved:
            sys.exit(f"existing pins changed: {moved}")
        golden["timeline"], golden["scenario"] = parts["timeline"], parts["scenario"]
        with open(path, "w") as f:
            json.dump(golden, f, indent=1)
        return

    if sys.argv[1:3] == ["--only", "post"]:   # refresh the post-chain pins alone (new cases); existing pins may not move
        with tempfile.TemporaryDirectory() as tmp:
            part = os.path.join(tmp, "post.json")
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", "post", part])
            with open(part) as f:
                new = json.load(f)
        path = os.path.join(HERE, "golden_post_720.json")
        with open(path) as f:
            golden = json.load(f)
        moved = [k for k, v in golden["cases"].items() if new.get(k) != v]
        if moved:
            sys.exit(f"existing pins changed: {moved}")
        golden["cases"] = new
        with open(path, "w") as f:
            json.dump(golden, f, indent=1)
        return

    with tempfile.TemporaryDirectory() as tmp:
        parts = {}
        for mode in ("timeline", "scenario", "post"):
            part = os.path.join(tmp, mode + ".json")
            print(f"[{mode}]")
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", mode, part])
            with open(part) as f:
                parts[mode] = json.load(f)

    with open(os.path.join(HERE, "golden_effects_720.json"), "w") as f:
        json.dump({"meta": meta, "timeline": parts["timeline"], "scenario": parts["scenario"]}, f, indent=1)
    with open(os.path.join(HERE, "golden_post_720.json"), "w") as f:
        json.dump({"meta": meta, "cases": parts["post"]}, f, indent=1)

    # RSQRTPS table of this CPU (2 parities x 1024 bins)
    from cookiedough_b200.assets import Assets
    R = oref.Reference(720, Assets(1280, 720, force_synthetic=True))
    full = R.rsqrt_table(stride=1 << 13)
    np.save(os.path.join(HERE, "rsqrt_table_golden.npy"), full)

    tracks = parse_rocket(os.path.join(oref.DATA_DIR, "directors-cut.rocket"))
    with open(os.path.join(HERE, "tracks.json"), "w") as f:
        json.dump({"source": "target/directors-cut.rocket", "row_rate": oref.ROW_RATE, "tracks": tracks}, f, separators=(",", ":"))
    print("done")


if __name__ == "__main__":
    main()

"""Host-side logic (no GPU): the C++ GNU Rocket reader of the host layer against an independent Python restatement of
sync_get_val (3rdparty/rocket-stripped/lib/track.c:9-60) on the committed key fixture, against the binary .track
loader, and against the reference's own Rocket when oracle/_ref is built."""
import json
import math
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN, REPO

from cookiedough_b200 import capi, hostapi


@pytest.fixture(scope="module")
def tracks():
    with open(os.path.join(GOLDEN, "tracks.json")) as f:
        return json.load(f)["tracks"]


def write_xml(path, tracks):
    with open(path, "w") as f:
        f.write('<sync rows="11000">\n\t<tracks>\n')
        for name, keys in tracks.items():
            f.write(f'\t\t<track name="{name}">\n')
            for row, value, interp in keys:
                f.write(f'\t\t\t<key interpolation="{interp}" row="{row}" value="{value!r}"/>\n')
            f.write('\t\t</track>\n')
        f.write('\t</tracks>\n</sync>\n')


from oracle.rocket import sync_get_val as py_sync_get_val  # noqa: E402  (track.c:32-60 restated: float key values, double interpolation)


ROWS = [0.0, 0.5, 499.999, 500.0, 1047.9, 1048.0, 2060.25, 3600.0, 4500.5, 5700.0, 6808.3, 7080.0, 7144.75, 8900.0, 9407.0, 10296.0, 10999.0, 12000.0]


def test_xml_reader_matches_python_restatement(tracks, tmp_path):
    xml = tmp_path / "t.rocket"
    write_xml(xml, tracks)
    rocket = hostapi.RocketOnly(xml)
    for row in ROWS:
        rocket.set_row(row)
        exact_row = (row / hostapi.ROW_RATE) * hostapi.ROW_RATE  # what Rocket::Boost computes (audio.cpp:175-178)
        for name, keys in tracks.items():
            assert rocket.track(name) == py_sync_get_val(keys, exact_row), (name, row)
    rocket.set_row(500)
    assert rocket.track("no:SuchTrack") == 0.0
    assert rocket.track_i("ball:RayLength") == capi.geti(py_sync_get_val(tracks["ball:RayLength"], (500 / hostapi.ROW_RATE) * hostapi.ROW_RATE))


def test_binary_track_reader_and_path_encoding(tracks, tmp_path):
    sync = tmp_path / "sync"
    sync.mkdir()
    names = ["twister::ShearSpeed", "ball:Radius", "closeSpike:MixBlurOpacity", "voxelScape:WarpStrength"]
    encoded = {"twister::ShearSpeed": "_twister-3A-3AShearSpeed.track", "ball:Radius": "_ball-3ARadius.track",
               "closeSpike:MixBlurOpacity": "_closeSpike-3AMixBlurOpacity.track", "voxelScape:WarpStrength": "_voxelScape-3AWarpStrength.track"}
    for name in names:
        with open(sync / encoded[name], "wb") as f:
            f.write(struct.pack("<i", len(tracks[name])))
            for row, value, interp in tracks[name]:
                f.write(struct.pack("<ifb", row, value, interp))
    rocket = hostapi.RocketOnly(str(sync) + "/")
    for row in ROWS:
        rocket.set_row(row)
        exact_row = (row / hostapi.ROW_RATE) * hostapi.ROW_RATE
        for name in names:
            assert rocket.track(name) == py_sync_get_val(tracks[name], exact_row), (name, row)


def test_demo_quit_stops_boost(tracks, tmp_path):
    xml = tmp_path / "t.rocket"
    write_xml(xml, tracks)
    rocket = hostapi.RocketOnly(xml)
    assert rocket.set_row(5000) == 1
    assert rocket.set_row(10300) == 0  # demo:quit = 1 from row 10296 (rocket.cpp:77-78)


def test_against_reference_rocket():
    from oracle import ref as oref
    if not oref.available(720):
        pytest.skip("oracle/_ref not built")
    import subprocess, sys, tempfile
    # the reference lives in its own process (global state); dump its values for a grid of rows
    names = sorted({t for _, (_, m) in capi.TRACKS.items() for t in m.values()})
    code = (
        "import sys, json; sys.path.insert(0, %r)\n"
        "from oracle import ref as o\n"
        "from cookiedough_b200.assets import Assets\n"
        "R = o.Reference(720, Assets(1280, 720, force_synthetic=True))\n"
        "names = json.loads(sys.argv[1]); rows = json.loads(sys.argv[2]); out = {}\n"
        "for r in rows:\n"
        "    R.set_row(r); out[str(r)] = [R.track(n) for n in names]\n"
        "print(json.dumps(out))\n" % REPO)
    rows = [float(r) for r in range(0, 10400, 173)] + ROWS
    ref_vals = json.loads(subprocess.check_output([sys.executable, "-c", code, json.dumps(names), json.dumps(rows)], cwd=REPO, text=True).strip().splitlines()[-1])
    for source in (os.path.join(oref.DATA_DIR, "directors-cut.rocket"), os.path.join(oref.DATA_DIR, "sync") + "/"):
        rocket = hostapi.RocketOnly(source)
        for r in rows:
            rocket.set_row(r)
            got = [rocket.track(n) for n in names]
            assert got == ref_vals[str(r)], (source, r)

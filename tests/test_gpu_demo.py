"""The compositor: Demo_Draw of the C++ host layer -- effect + the part's layers, composed on the
device -- against the reference's Demo_Draw (code/demo.cpp:469-1023).

* golden: committed frames of the compiled reference (tests/golden/golden_demo_720.json: synthetic assets, rows that reach
  every part and every optional layer of a part);
* live: the compiled reference (oracle/_ref) next to the CUDA path, same real art, rows across the whole timeline;
* 4K: the same in a child process (a second reference instance with the compositor cannot live in this one).

Every blend / blit / blur / warp of the chain is integer arithmetic that is bit-exact on its own (test_gpu_post.py), so a
composed frame may differ from the reference's only where the float effect underneath does: the same tolerance as the bare
effect applies (>= 99.5 % of the pixels exact, <= 2 LSB per channel), integer-only parts (12 without the plasma, 13) are exact."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, REPO
from util import MAX_LSB, MIN_EXACT_PCT, pixel_stats

pytestmark = pytest.mark.gpu

LIVE_ROWS = [0, 450, 508, 980, 1030, 1300, 1500, 2008, 2060, 2100, 2364, 2510, 3064, 3130, 3250, 3716, 4204, 4246, 4300, 4500, 4996, 5050,
             5378, 5700, 6026, 6290, 6628, 6684, 6728, 6812, 7340, 7882, 8254, 8500, 9300, 9410, 9524, 9556, 9794, 9924, 9980, 10142, 10310]


def _check(out, ref, label, failures):
    exact, max_delta = pixel_stats(out, ref)
    if exact < MIN_EXACT_PCT or max_delta > MAX_LSB:
        failures.append(f"{label}: {exact:.4f}% exact, max delta {max_delta} LSB")


def test_demo_draw_matches_golden_frames(golden_rsqrt):
    from cookiedough_b200 import hostapi
    from cookiedough_b200.assets import Assets
    with open(os.path.join(GOLDEN, "golden_demo_720.json")) as f:
        frames = json.load(f)["frames"]
    host = hostapi.Host(1280, 720, 0, Assets(1280, 720, force_synthetic=True), demo=True)
    try:
        ctx = host.context()
        ctx.set_rsqrt_table(golden_rsqrt, 13)
        out = np.zeros((720, 1280), dtype=np.uint32)
        seed = (np.arange(1280 * 720, dtype=np.uint32) * np.uint32(2654435761)).reshape(720, 1280)
        mismatches = []
        for key, case in frames.items():
            out.fill(0)  # past the end of the timeline Demo_Draw returns false and leaves the frame alone, like the reference
            ctx.upload(ctx.render_target(0), seed)  # as the golden generator does (the ball's beam path keeps stale pixels)
            host.demo_draw(out, case["time"])
            if hashlib.sha256(out.astype("<u4").tobytes()).hexdigest() != case["sha256"]:
                crop = case["crop"]
                want = np.frombuffer(bytes.fromhex(crop["hex"]), dtype="<u4").reshape(crop["h"], crop["w"])
                got = out[crop["y"]:crop["y"] + crop["h"], crop["x"]:crop["x"] + crop["w"]]
                mismatches.append(f"row {key} part {case['part']}: sha256 differs, crop equal: {np.array_equal(got, want)}")
        assert not mismatches, "\n".join(mismatches)
    finally:
        host.close()


def test_demo_draw_matches_reference_live():
    from oracle import ref as oref
    if not oref.available(720):
        pytest.skip("oracle/_ref not built")
    from cookiedough_b200 import hostapi
    from cookiedough_b200.assets import Assets
    R = oref.Reference.get(720, Assets(1280, 720), demo=True)
    host = hostapi.Host(1280, 720, 0, Assets(1280, 720), demo=True)
    try:
        out = np.zeros((720, 1280), dtype=np.uint32)
        failures, parts = [], set()
        for row in LIVE_ROWS:
            t = float(np.float32(row / oref.ROW_RATE))
            R.set_time(t)
            ref = R.demo_draw().copy()
            parts.add(int(round(R.track("demo:Effect"))))
            out.fill(0)
            host.demo_draw(out, t)
            _check(out, ref, f"row {row}", failures)
        assert not failures, "\n".join(failures)
        assert parts == set(range(1, 14)), f"parts covered: {sorted(parts)}"
    finally:
        host.close()


def test_demo_draw_pipelined_frames_arrive_in_order():
    """CkdHost_SetPipelined: Demo_Draw returns once enqueued, pDest is complete after the second following call / the flush"""
    from oracle import ref as oref
    if not oref.available(720):
        pytest.skip("oracle/_ref not built")
    from cookiedough_b200 import hostapi
    from cookiedough_b200.assets import Assets
    host = hostapi.Host(1280, 720, 0, Assets(1280, 720), demo=True)
    try:
        rows = [300, 1500, 2600, 4300, 5700, 9980]
        sync = [np.zeros((720, 1280), dtype=np.uint32) for _ in rows]
        for buf, row in zip(sync, rows):
            host.demo_draw(buf, row / oref.ROW_RATE)
        host.set_pipelined(True)
        piped = [np.zeros((720, 1280), dtype=np.uint32) for _ in rows]
        for buf, row in zip(piped, rows):
            host.demo_draw(buf, row / oref.ROW_RATE)
        host.flush()
        host.set_pipelined(False)
        for a, b, row in zip(sync, piped, rows):
            assert np.array_equal(a, b), f"row {row}"
    finally:
        host.close()


def test_demo_draw_4k_child_process():
    from oracle import ref as oref
    if not oref.available(2160):
        pytest.skip("oracle/_ref not built")
    rows = "450,1500,2364,4246,5700,6684,7882,9410,9980"
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "tools", "demo_parity.py"), "--res", "2160", "--rows", rows],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    worst = [line for line in r.stdout.splitlines() if line.startswith("worst exact %")]
    assert worst and float(worst[-1].split()[-1]) >= 99.5, r.stdout[-3000:]
    assert all(int(line.split("max")[1].split()[0]) <= 2 for line in r.stdout.splitlines() if line.startswith("row")), r.stdout[-3000:]


def test_render_demo_stream_matches_demo_draw(tmp_path):
    """tools/render_demo.py (e + f4: CkdTimeline_Render -> ckd_gather ring -> sink's pinned ring -> writer thread) writes the frames Demo_Draw produces"""
    from oracle import ref as oref
    if not oref.available(720):
        pytest.skip("oracle/_ref not built")
    from cookiedough_b200 import hostapi, sharding, sink
    from cookiedough_b200.assets import Assets
    path = tmp_path / "demo.ckdf"
    frames = 24
    r = subprocess.run([sys.executable, os.path.join(REPO, "tools", "render_demo.py"), "--out", str(path), "--frames", str(frames), "--res", "720"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert sink.read_header(path) == (1280, 720, frames)
    host = hostapi.Host(1280, 720, 0, Assets(1280, 720), demo=True)
    try:
        host.context().set_frame_independent(True)   # what CkdTimeline_Render renders with (frames must not depend on the sharding)
        out = np.zeros((720, 1280), dtype=np.uint32)
        times = sharding.timeline_times(frames)
        for i in range(frames):
            host.demo_draw(out, times[i])
            assert np.array_equal(sink.read_frame(path, i), out), f"frame {i}"
    finally:
        host.close()


def test_demo_draw_unknown_part_draws_the_test_pattern(tmp_path):
    """demo:Effect outside 1..13 -> FxBlitter_DrawTestPattern (code/demo.cpp:1000-1001, fx-blitter.cpp:77-98) + the post fade;
    driven through a hand-written Rocket project, so this also covers Demo_Create on a caller-supplied .rocket file"""
    import json
    from test_host_rocket import write_xml
    from cookiedough_b200 import hostapi
    from cookiedough_b200.assets import Assets
    with open(os.path.join(GOLDEN, "tracks.json")) as f:
        tracks = json.load(f)["tracks"]
    tracks["demo:Effect"] = [[0, 0.0, 0]]
    tracks["demo:FadeToBlack"] = [[0, 0.25, 0]]
    tracks["demo:FadeToWhite"] = [[0, 0.0, 0]]
    xml = tmp_path / "pattern.rocket"
    write_xml(xml, tracks)
    host = hostapi.Host(1280, 720, 0, Assets(1280, 720, force_synthetic=True), rocket_source=xml, demo=True)
    try:
        ctx = host.context()
        out = np.zeros((720, 1280), dtype=np.uint32)
        assert host.demo_draw(out, 1.0)
        fx_y, fx_x = 720 // 2 + 4, 1280 // 2 + 4
        iy, ix = np.mgrid[0:fx_y, 0:fx_x]
        pattern = np.where(iy < fx_y // 2, np.where(iy & 1, 0xFFFFFFFF, 0), np.where(ix & 1, 0xFFFFFFFF, 0)).astype(np.uint32)
        d_fx = ctx.to_device(pattern, pad_elems=4 * 1280)
        ctx.fx_blit_2x2(ctx.frame(), d_fx)
        ctx.blend("Fade32", ctx.frame(), ctx.frame(), 1280 * 720, 0.0, (int(np.float32(0.25) * np.float32(255.0)) << 24))
        assert np.array_equal(out, ctx.read_frame())
        assert out.any() and not (out == out.flat[0]).all()
        ctx.free(d_fx)
    finally:
        host.close()

"""The frame gather over peer memory (include/ckd.h ckd_gather_*) and the frame-sharded timeline runner
(CkdTimeline_Render): frames rendered by several processes -- one per GPU; on a one-GPU box several on the same GPU, which
takes the same CUDA IPC path -- must arrive on the collector bit for bit, in order, and the gathered stream must not depend on
how many processes rendered it."""
import ctypes
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import REPO
from cookiedough_b200 import capi

pytestmark = pytest.mark.gpu
WORKER = os.path.join(REPO, "tests", "tools", "gather_worker.py")


def _pattern(i, n):
    return (np.arange(n, dtype=np.uint32) * np.uint32(2654435761) + np.uint32(i * 977)).astype(np.uint32)


def test_frame_checksum_matches_the_host_formula(ctx_synth):
    n = ctx_synth.res_x * ctx_synth.res_y
    frame = _pattern(5, n)
    ctx_synth.upload(ctx_synth.frame(), frame)
    assert ctx_synth.frame_checksum() == capi.frame_checksum_host(frame)
    # position dependent: swapping two different pixels changes it
    frame[[3, 70000]] = frame[[70000, 3]]
    ctx_synth.upload(ctx_synth.frame(), frame)
    assert ctx_synth.frame_checksum() == capi.frame_checksum_host(frame) != capi.frame_checksum_host(_pattern(5, n))


def test_ring_round_trip_in_one_process(ctx_synth):
    """collector and producer in one process: 11 frames through a 2-slot ring and 3 staging frames (both wrap), checksummed on
    the device and copied to the host"""
    ctx = ctx_synth
    n = ctx.res_x * ctx.res_y
    g = capi.Gather(ctx, slots=2)
    try:
        g.set_timeout_ms(20000)
        ring = [ctx.malloc_host(n * 4) for _ in range(2)]
        frames = 11
        got = {}
        for seq in range(frames):
            d = g.acquire()
            ctx.upload(d, _pattern(seq, n))
            g.push(seq)
            g.pop(seq, capi.GATHER_CHECKSUM | capi.GATHER_TO_HOST, ring[seq % 2])
            g.wait_pop(seq)
            buf = (ctypes.c_uint32 * n).from_address(ring[seq % 2])
            got[seq] = np.frombuffer(buf, dtype=np.uint32).copy()
        g.flush()
        ctx.sync()
        g.status()
        sums = g.checksums(0, frames)
        for seq in range(frames):
            want = _pattern(seq, n)
            assert np.array_equal(got[seq], want), f"frame {seq} arrived damaged"
            assert sums[seq] == capi.frame_checksum_host(want)
        assert g.peer_bytes() == 0  # the ring is local to this process
        for p in ring:
            ctx.free_host(p)
    finally:
        g.close()


def test_producer_blocks_on_the_device_until_the_slot_is_drained(ctx_synth):
    """pushing more frames than the ring has slots before anything is popped must not lose a frame: the third push waits (on the
    device) for the first pop"""
    ctx = ctx_synth
    n = ctx.res_x * ctx.res_y
    g = capi.Gather(ctx, slots=2)
    try:
        g.set_timeout_ms(20000)
        for seq in range(3):               # 3 staging frames, 2 slots: the push of seq 2 parks behind a device-side wait
            d = g.acquire()
            ctx.upload(d, _pattern(seq, n))
            g.push(seq)
        for seq in range(3):
            g.pop(seq, capi.GATHER_CHECKSUM)
        g.flush()
        ctx.sync()
        g.status()
        assert g.checksums(0, 3) == [capi.frame_checksum_host(_pattern(s, n)) for s in range(3)]
    finally:
        g.close()


def test_a_wait_that_cannot_be_satisfied_times_out_instead_of_hanging(ctx_synth):
    g = capi.Gather(ctx_synth, slots=2)
    try:
        g.set_timeout_ms(50)
        g.pop(0, 0)                        # nobody pushes sequence number 0
        with pytest.raises(capi.CkdError, match="timed out"):
            g.status()
    finally:
        g.close()


def _run_world(world, frames, devices, to_host=False, slots=4, passes=1, collector_skip=1, lanes=1):
    with tempfile.TemporaryDirectory() as scratch:
        procs = []
        for rank in range(world):
            cmd = [sys.executable, WORKER, "--rank", str(rank), "--world", str(world), "--device", str(devices[rank % len(devices)]),
                   "--dir", scratch, "--frames", str(frames), "--slots", str(slots), "--passes", str(passes), "--collector-skip", str(collector_skip), "--lanes", str(lanes)]
            if to_host:
                cmd.append("--to-host")
            procs.append(subprocess.Popen(cmd, cwd=REPO, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = []
        for p in procs:
            try:
                out, _ = p.communicate(timeout=600)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
            outs.append(out)
        for rank, (p, out) in enumerate(zip(procs, outs)):
            assert p.returncode == 0, f"rank {rank} of {world} failed:\n{out[-3000:]}"
        with open(os.path.join(scratch, "result.json")) as f:
            result = json.load(f)
        result["peer_bytes_of_rank"] = {}
        for rank in range(1, world):
            with open(os.path.join(scratch, f"peer_bytes.{rank}")) as f:
                result["peer_bytes_of_rank"][rank] = int(f.read())
        return result


def test_gathered_timeline_does_not_depend_on_the_number_of_ranks():
    """24 frames across the whole timeline (every part, the ball with beams included) through Demo_Draw: rendered by 1, 2 and 3
    processes, gathered on rank 0 -- the per-frame checksums of the gathered stream are identical, and equal to the checksums
    of the frames a plain Demo_Draw loop delivers to the host"""
    import torch
    n_dev = torch.cuda.device_count()
    devices = list(range(n_dev))
    frames = 24
    one = _run_world(1, frames, devices[:1], to_host=True)
    two = _run_world(2, frames, devices)
    three = _run_world(3, frames, devices, slots=2, passes=2)
    # (the frame at t = 0 is black: its checksum is 0)
    assert len(one["checksums"]) == frames and len({c for c in one["checksums"] if int(c) != 0}) >= frames - 2
    assert two["checksums"] == one["checksums"]
    assert three["checksums"] == one["checksums"] * 2            # two passes over the same timeline
    # two lanes per rank (two frames in flight on two contexts of the same GPU): the same stream again, alone and sharded
    assert _run_world(1, frames, devices[:1], lanes=2, to_host=True)["checksums"] == one["checksums"]
    assert _run_world(2, frames, devices, lanes=2, passes=2)["checksums"] == one["checksums"] * 2
    # weighted sharding (the collector renders one frame per two rounds of the others): the same stream, fewer frames on rank 0
    weighted = _run_world(3, frames, devices, slots=3, collector_skip=2)
    assert weighted["checksums"] == one["checksums"]
    from cookiedough_b200 import sharding as _sh
    for rank in (1, 2):
        assert weighted["peer_bytes_of_rank"][rank] == len(_sh.frames_for_rank(frames, rank, 3, 2)) * 1280 * 720 * 4
    for seq, s in one["host_checksums"].items():                 # what reached the host is what was summed on the device
        assert str(s) == one["checksums"][int(seq)]
    frame_bytes = 1280 * 720 * 4
    assert two["peer_bytes_of_rank"][1] == (frames // 2) * frame_bytes   # rank 1 pushed its 12 frames through the mapped ring
    assert one["peer_bytes"] == 0

    # the same frames through the plain host API, one process, in order
    from cookiedough_b200 import hostapi, sharding
    from cookiedough_b200.assets import Assets
    host = hostapi.Host(1280, 720, 0, Assets(1280, 720, force_synthetic=True), demo=True)
    try:
        ctx = host.context()
        ctx.set_rsqrt_table(np.load(os.path.join(REPO, "tests", "golden", "rsqrt_table_golden.npy")), 13)
        ctx.set_frame_independent(True)
        out = np.zeros((720, 1280), dtype=np.uint32)
        direct = []
        for t in sharding.timeline_times(frames):
            host.demo_draw(out, t)
            direct.append(str(capi.frame_checksum_host(out)))
        assert direct == one["checksums"]
    finally:
        host.close()


def test_ball_with_beams_is_frame_independent_when_asked(ctx_synth, golden_effects):
    """with beams the ball leaves the last pixel of every ray row of render target 0 to history (ball.cpp:186,352);
    ckd_set_frame_independent clears it, so the frame no longer depends on what was rendered before -- and equals the
    reference's frame for a render target whose last column is zero"""
    ctx = ctx_synth
    case = golden_effects["timeline"]["ball@2060"]
    cls, _ = capi.TRACKS["ball"]
    p = cls()
    for k, v in case["params"].items():
        setattr(p, k, v)
    seed = (np.arange(1280 * 720, dtype=np.uint32) * np.uint32(2654435761)).reshape(720, 1280)

    def draw(history):
        ctx.upload(ctx.render_target(0), history)
        ctx.draw("ball", p, case["time"])
        return ctx.read_frame()

    # the reference's behaviour (flag off): history shows through
    assert not np.array_equal(draw(seed), draw(np.zeros_like(seed)))
    ctx.set_frame_independent(True)
    try:
        a, b = draw(seed), draw(np.zeros_like(seed))
        assert np.array_equal(a, b)
        cleared = seed.copy()
        cleared[:, -1] = 0
        ctx.set_frame_independent(False)
        assert np.array_equal(draw(cleared), a)   # = the pinned reference algorithm on a target whose last column is 0
    finally:
        ctx.set_frame_independent(False)

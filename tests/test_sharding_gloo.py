"""N > 1 host logic on CPU: two gloo ranks shard a timeline by frame, reduce the max time and gather per-frame checksums."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cookiedough_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_frames, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.frames_for_rank(num_frames, rank, world)
        times = sharding.timeline_times(num_frames)
        # stand-in for rendering: a "frame" that depends only on the frame's time, like every X_Draw does
        local = {i: sharding.frame_checksum(torch.full((4, 4), int(times[i] * 1000) & 0xFFFFFF, dtype=torch.int32).numpy()) for i in mine}
        full = sharding.gather_checksums(dist, local, num_frames)
        slowest = sharding.reduce_max(dist, 10.0 + rank)
        dist.barrier()
        if rank == 0:
            out.put((mine, full, slowest))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_ranks_shard_a_timeline():
    world, num_frames = 2, 37
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    mine0, full, slowest = q.get(timeout=90)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert mine0 == list(range(0, num_frames, 2))
    times = sharding.timeline_times(num_frames)
    expect = [sharding.frame_checksum(torch.full((4, 4), int(t * 1000) & 0xFFFFFF, dtype=torch.int32).numpy()) for t in times]
    assert full == expect              # every frame rendered exactly once, by some rank
    assert slowest == 11.0             # max over ranks


def test_sharding_covers_every_frame_once():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in sharding.frames_for_rank(600, r, world))
        assert seen == list(range(600))
    with pytest.raises(ValueError):
        sharding.frames_for_rank(10, 2, 2)


def test_timeline_times_match_config5():
    t = sharding.timeline_times(600)
    assert len(t) == 600 and t[0] == 0.0
    assert abs(t[1] - 0.369828) < 1e-6   # t_i = i * 0.369828 s
    assert t[-1] < 10296 / sharding.ROW_RATE


def test_weighted_sharding_matches_the_library_and_covers_every_frame():
    """frame_owner (python) == CkdTimeline_Owner (C++), every frame has exactly one owner, and with collector_skip = k rank 0
    renders one frame per k rounds of the others"""
    from cookiedough_b200 import hostapi
    L = hostapi._lib() if hasattr(hostapi, "_lib") else None
    for world in (1, 2, 3, 4, 8):
        for skip in (0, 1, 2, 3):
            owners = [sharding.frame_owner(i, world, skip) for i in range(600)]
            assert all(0 <= o < world for o in owners)
            if L is not None:
                assert owners == [L.ckdhost_timeline_owner(i, world, skip) for i in range(600)]
            per_rank = [sharding.frames_for_rank(600, r, world, skip) for r in range(world)]
            assert sorted(i for fr in per_rank for i in fr) == list(range(600))
            if world > 1 and skip > 1:
                cycle = skip * (world - 1) + 1
                assert abs(len(per_rank[0]) - 600 / cycle) <= 1
                assert all(abs(len(fr) - 600 * skip / cycle) <= skip for fr in per_rank[1:])
            elif world > 1:
                assert owners == [i % world for i in range(600)]
    assert sharding.default_collector_skip(2) == 1 and sharding.default_collector_skip(8) == 2
    if L is not None:
        assert [L.ckdhost_timeline_default_skip(w) for w in (1, 2, 3, 4, 8)] == [sharding.default_collector_skip(w) for w in (1, 2, 3, 4, 8)]

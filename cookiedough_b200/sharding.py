"""Frame-parallel sharding of a timeline across the GPUs of one box.

Every X_Draw is a pure function of (time, Rocket row, assets), so a timeline shards by frame index with no data-path
collective: frame i goes to rank i mod N (or, when rank 0 also collects the frames, to a weighted round-robin that gives rank 0
fewer turns: frame_owner).  The only communication is bookkeeping (barrier, max-over-ranks time, gathering
per-frame checksums to rank 0) and runs over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import zlib

import numpy as np

ROW_RATE = (170.0 / (60.0 * (170.0 / 174.0))) * 16.0   # code/audio.cpp:18
TIMELINE_ROWS = 10296                                   # demo:quit fires here

# demo:Effect value -> effect entry point (code/demo.cpp:508-1003); 12 draws Plasma only when demo:FullWarpTPB is set,
# 13 (sprites) has no effect layer
EFFECT_OF_PART = {1: "twister", 2: "landscape", 3: "ball", 4: "tunnelscape", 5: "plasma", 6: "nautilus", 7: "spikey_close",
                  8: "spikey_distant", 9: "tunnel", 10: "sinuses", 11: "laura", 12: "plasma", 13: None}


def timeline_times(num_frames=600):
    """t_i = i * (10296 / 46.4) / num_frames  """
    total = TIMELINE_ROWS / ROW_RATE
    return [i * total / num_frames for i in range(num_frames)]


def frame_owner(frame, world_size, collector_skip=1):
    """which rank renders frame i (the rule of CkdTimeline_Owner, host/ckd_timeline.cpp): i mod N, or, with collector_skip = k > 1,
    cycles of k*(N-1) + 1 frames whose first frame goes to rank 0 -- the rank that also collects every frame -- and whose other
    frames go round-robin over the ranks 1..N-1"""
    if world_size <= 1:
        return 0
    if collector_skip <= 1:
        return frame % world_size
    c = frame % (collector_skip * (world_size - 1) + 1)
    return 0 if c == 0 else 1 + (c - 1) % (world_size - 1)


def default_collector_skip(world_size):
    """CkdTimeline_DefaultCollectorSkip: rank 0 renders a full share up to 3 GPUs, one frame per two rounds from 4 GPUs on"""
    return 2 if world_size >= 4 else 1


def frames_for_rank(num_frames, rank, world_size, collector_skip=1):
    """the frames rank `rank` renders (frame i -> rank i mod N by default)"""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return [i for i in range(num_frames) if frame_owner(i, world_size, collector_skip) == rank]


def frame_checksum(frame):
    return zlib.crc32(np.ascontiguousarray(frame).view(np.uint8)) & 0xFFFFFFFF


def reduce_max(dist, value, device="cpu"):
    """max over ranks of a python float (the timing contract: max-over-ranks, on whatever backend dist uses)"""
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_checksums(dist, local, num_frames, device="cpu"):
    """local: {frame index: crc32}; returns the full per-frame list on every rank (all-reduce of disjoint slots)"""
    import torch
    t = torch.zeros(num_frames, dtype=torch.int64, device=device)
    for i, c in local.items():
        t[i] = int(c)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.tolist()]

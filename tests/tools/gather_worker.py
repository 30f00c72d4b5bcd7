#!/usr/bin/env python3
"""One rank of a frame-sharded timeline run with the peer-memory gather (ckd_gather_*, include/ckd.h), as its own process --
what one GPU's process does in `bench.py --gpus N` / tools/render_demo.py, without torch.distributed: the ring handle travels
through a file.  tests/test_gpu_gather.py starts `world` of these (on one GPU, or one per GPU when the box has several).

    gather_worker.py --rank R --world N --device D --dir SCRATCH --frames F [--res 720] [--passes P] [--to-host]

rank 0 writes SCRATCH/result.json: per-frame checksums computed by the collector on the gathered frames (and, with --to-host,
host-side checksums of the frames it copied out)."""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)


def wait_for(path, timeout=120.0):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise TimeoutError(path)
        time.sleep(0.02)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rank", type=int, required=True)
    ap.add_argument("--world", type=int, required=True)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--dir", required=True)
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--res", type=int, default=720)
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--slots", type=int, default=4)
    ap.add_argument("--to-host", action="store_true")
    ap.add_argument("--collector-skip", type=int, default=1)
    ap.add_argument("--lanes", type=int, default=1)
    args = ap.parse_args()

    from cookiedough_b200 import capi, hostapi, sharding
    from cookiedough_b200.assets import Assets

    res_y = args.res
    res_x = res_y * 16 // 9
    host = hostapi.Host(res_x, res_y, args.device, Assets(res_x, res_y, force_synthetic=True), demo=True)
    ctx = host.context()
    rsqrt = np.load(os.path.join(REPO, "tests", "golden", "rsqrt_table_golden.npy"))
    ctx.set_rsqrt_table(rsqrt, 13)
    times = sharding.timeline_times(args.frames)
    handle_path = os.path.join(args.dir, "handle.bin")

    if args.rank == 0:
        gather = capi.Gather(ctx, slots=args.slots)
        gather.set_timeout_ms(60000)
        with open(handle_path + ".tmp", "wb") as f:
            f.write(gather.export())
        os.rename(handle_path + ".tmp", handle_path)
    else:
        wait_for(handle_path)
        with open(handle_path, "rb") as f:
            gather = capi.Gather(ctx, handle=f.read())
        gather.set_timeout_ms(60000)

    mode = capi.GATHER_CHECKSUM | (capi.GATHER_TO_HOST if args.to_host else 0)
    ring, host_sums = None, {}
    if args.rank == 0 and args.to_host:
        ring = [ctx.malloc_host(res_x * res_y * 4) for _ in range(3)]
    t0 = time.perf_counter()
    host.timeline_render(times, rank=args.rank, world=args.world, gather=gather, passes=args.passes, pop_mode=mode, host_ring=ring, collector_skip=args.collector_skip, lanes=args.lanes)
    ctx.sync()
    elapsed = time.perf_counter() - t0
    gather.status()

    if args.rank == 0:
        total = args.frames * args.passes
        sums = gather.checksums(0, total)
        if ring:
            # the last len(ring) frames are still in the host ring: check that what arrived is what was summed on the device
            import ctypes
            for seq in range(max(0, total - len(ring)), total):
                buf = (ctypes.c_uint32 * (res_x * res_y)).from_address(ring[seq % len(ring)])
                host_sums[str(seq)] = capi.frame_checksum_host(np.frombuffer(buf, dtype=np.uint32))
        with open(os.path.join(args.dir, "result.json.tmp"), "w") as f:
            json.dump({"checksums": [str(s) for s in sums], "host_checksums": host_sums, "seconds": elapsed, "world": args.world,
                       "peer_bytes": gather.peer_bytes()}, f)
        os.rename(os.path.join(args.dir, "result.json.tmp"), os.path.join(args.dir, "result.json"))
        for r in range(1, args.world):          # the ring may only go away after every producer has unmapped it
            wait_for(os.path.join(args.dir, f"done.{r}"))
    else:
        with open(os.path.join(args.dir, f"peer_bytes.{args.rank}"), "w") as f:
            f.write(str(gather.peer_bytes()))
    gather.close()
    if args.rank != 0:
        open(os.path.join(args.dir, f"done.{args.rank}"), "w").close()
    host.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Render the demo timeline to a raw frame stream (SURVEY.md section 8 rows e + f1 + f4): Demo_Draw on every GPU of the box
(CkdTimeline_Render, include/ckd_host.h), every frame pushed into the slot ring in rank 0's HBM by the library's peer-memory
gather (ckd_gather_*, include/ckd.h: CUDA IPC mapping, copy-engine peer copies over NVLink, device-side flags -- no NCCL on
the data path; torch.distributed only carries the 128-byte ring handle), drained in frame order into the sink's pinned host
ring and written by its writer thread.

    python tools/render_demo.py --out /tmp/demo.ckdf --frames 600 [--res 2160]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/render_demo.py --out ... --frames 600

Frame i is rendered by rank i mod N.  Prints one JSON line (rank 0): frames per second into the file, max over ranks."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--res", type=int, default=2160)
    ap.add_argument("--ring", type=int, default=6)
    ap.add_argument("--lanes", type=int, default=2, help="contexts per GPU the rank's frames alternate between (two frames in flight)")
    ap.add_argument("--slots", type=int, default=0, help="frames of the gather's slot ring in rank 0's HBM (0: 4 per rank, at least 8)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from cookiedough_b200 import capi, hostapi, sharding, sink
    from cookiedough_b200.assets import Assets

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    res_y = args.res; res_x = res_y * 16 // 9
    n = res_x * res_y
    host = hostapi.Host(res_x, res_y, local, Assets(res_x, res_y), demo=True)
    ctx = host.context()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    times = sharding.timeline_times(args.frames)

    # the ring lives on rank 0; its handle travels once over the control plane
    if rank == 0:
        gather = capi.Gather(ctx, slots=args.slots or min(64, max(8, 4 * world)))
        handle = gather.export()
    else:
        gather, handle = None, bytes(capi.GATHER_HANDLE_BYTES)
    if world > 1:
        t = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, src=0)
        if rank != 0:
            gather = capi.Gather(ctx, handle=bytes(t.cpu().tolist()))
    gather.set_timeout_ms(60000)

    out = sink.Sink(args.out, res_x, res_y, args.frames, ring_frames=args.ring, pinned=True, create=True) if rank == 0 else None

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    # rank r renders the frames i % world == r and pushes them; rank 0 pops every frame in order into the open sink
    host.timeline_render(times, rank=rank, world=world, gather=gather, passes=1, pop_mode=capi.GATHER_CHECKSUM | capi.GATHER_TO_HOST, host_ring=None, lanes=args.lanes)
    ctx.sync()
    gather.status()
    if rank == 0:
        out.close()
    elapsed = time.perf_counter() - t0
    peer_bytes = gather.peer_bytes()
    if world > 1:
        t = torch.tensor([elapsed, float(peer_bytes)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
        elapsed, peer_bytes = float(t[0].item()), float(t[1].item())
    if rank == 0:
        print(json.dumps({"tool": "render_demo", "frames": args.frames, "res": [res_x, res_y], "n_gpus": world, "seconds": elapsed,
                          "fps": args.frames / elapsed, "mpixel_s": args.frames * n / elapsed / 1e6, "gbytes_written": args.frames * n * 4 / 1e9,
                          "out": args.out, "nvlink_gbytes": peer_bytes / 1e9,
                          "gather": "ckd_gather_* (slot ring in rank 0's HBM, peer copies, device-side flags) -> sink's pinned ring -> writer thread"}))
    if world > 1:
        dist.barrier()          # the ring may only go away after every producer has unmapped it: producers close first
    if rank != 0:
        gather.close()
    if world > 1:
        dist.barrier()
    if rank == 0:
        gather.close()
    host.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

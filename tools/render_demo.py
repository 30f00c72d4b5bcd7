#!/usr/bin/env python3
"""Render the demo timeline to a raw frame stream (SURVEY.md section 8 rows f1 + f4): Demo_Draw on every GPU of the box, the
frames gathered to rank 0 over NCCL (device to device, NVLink), copied into a pinned host ring and written by the sink's
writer thread.

    python tools/render_demo.py --out /tmp/demo.ckdf --frames 600 [--res 2160]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/render_demo.py --out ... --frames 600

Frame i is rendered by rank i mod N.  Prints one JSON line (rank 0): frames per second into the file, max over ranks."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class DeviceFrame:
    """__cuda_array_interface__ view of a device frame owned by the renderer's context"""
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (int(ptr), False), "version": 3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--res", type=int, default=2160)
    ap.add_argument("--ring", type=int, default=6)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from cookiedough_b200 import hostapi, sharding, sink
    from cookiedough_b200.assets import Assets

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    res_y = args.res; res_x = res_y * 16 // 9
    n = res_x * res_y
    host = hostapi.Host(res_x, res_y, local, Assets(res_x, res_y), demo=True)
    ctx = host.context()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    times = sharding.timeline_times(args.frames)

    out = sink.Sink(args.out, res_x, res_y, args.frames, ring_frames=args.ring, pinned=True, create=True) if rank == 0 else None
    frame_dev = torch.as_tensor(DeviceFrame(ctx.frame(), n), device="cuda")      # the context's device frame, zero-copy
    recv = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(2)] if rank == 0 else None
    copy_stream = torch.cuda.Stream() if rank == 0 else None

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pending = []  # rank 0: (event, buffer address, frame index) of device->host copies in flight
    for i in range(args.frames):
        owner = i % world
        if owner == rank:
            host.demo_draw(0, times[i])                      # composed frame stays on the device
        if rank == 0:
            if owner == 0:
                src = frame_dev
            else:
                src = recv[i & 1]
                dist.recv(src, src=owner)
            ptr = out.acquire()                              # pinned host buffer of the sink's ring
            dst = torch.from_numpy(out.view(ptr).reshape(-1).view(np.int32))
            copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(copy_stream):
                dst.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            torch.cuda.current_stream().wait_stream(copy_stream)   # src (device frame / receive buffer) is reused by what follows
            pending.append((ev, ptr, i))
            while len(pending) > 2:
                e, p, k = pending.pop(0)
                e.synchronize()
                out.commit(p, k)
        elif owner == rank:
            dist.send(frame_dev, dst=0)
    if rank == 0:
        for e, p, k in pending:
            e.synchronize()
            out.commit(p, k)
        out.close()
    torch.cuda.synchronize()
    elapsed = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    if rank == 0:
        print(json.dumps({"tool": "render_demo", "frames": args.frames, "res": [res_x, res_y], "n_gpus": world, "seconds": elapsed,
                          "fps": args.frames / elapsed, "mpixel_s": args.frames * n / elapsed / 1e6, "gbytes_written": args.frames * n * 4 / 1e9,
                          "out": args.out, "gather": "NCCL send/recv to rank 0, pinned ring, writer thread" if world > 1 else "single rank"}))
    host.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

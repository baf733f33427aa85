#!/usr/bin/env python
"""tools/h2d_peak.py -- the host<->device copy ceiling bench.py's `e2e` runs under.

Pinned-memory cudaMemcpyAsync (through torch) H2D alone, D2H alone and both directions at once, per rank and summed
over the ranks of one node:

  python tools/h2d_peak.py                                               one GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_peak.py     N GPUs concurrently

All ranks start together (barrier) and the time is the max over ranks, like bench.py.  Prints one JSON line (rank 0).
The e2e workload's own traffic pattern is also timed: `mix` = H2D of 252 MB and D2H of 22 MB per step, overlapped,
the byte counts bench.py's headline e2e moves per step.
"""
import json
import os

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = 256 << 20
    h_src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_src.fill_(1)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.ones(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=10):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        for s in (s1, s2):
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def h2d(n=nbytes):
        s1.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s1):
            d_a[:n].copy_(h_src[:n], non_blocking=True)

    def d2h(n=nbytes):
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s2):
            h_dst[:n].copy_(d_b[:n], non_blocking=True)

    def both():
        h2d()
        d2h()

    MIX_IN, MIX_OUT = 251_662_208, 21_937_152

    def mix():
        h2d(MIX_IN)
        d2h(MIX_OUT)

    t_h2d, t_d2h, t_both, t_mix = timed(h2d), timed(d2h), timed(both), timed(mix)
    if rank == 0:
        gb = nbytes / 1e9
        print(json.dumps({
            "what": "pinned cudaMemcpyAsync ceilings, max over ranks, all ranks concurrently", "n_gpus": world,
            "bytes_per_copy": nbytes,
            "h2d_gbs_per_gpu": gb / (t_h2d * 1e-3), "d2h_gbs_per_gpu": gb / (t_d2h * 1e-3),
            "duplex_gbs_per_gpu_each_way": gb / (t_both * 1e-3),
            "h2d_gbs_aggregate": world * gb / (t_h2d * 1e-3), "d2h_gbs_aggregate": world * gb / (t_d2h * 1e-3),
            "mix": {"h2d_bytes": MIX_IN, "d2h_bytes": MIX_OUT, "ms": t_mix,
                    "windows_per_s_ceiling_at_9472_windows_per_step": world * 9472 / (t_mix * 1e-3)}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""NVLink all-to-all reference for the slab exchange: what NCCL's all_to_all_single and plain peer copies reach on this
box for the message sizes of CH-3D-n on P GPUs (per rank: 2 half spectra forward, 1 back).  Run under torchrun.
Prints one JSON line on rank 0: per-GPU egress GB/s = bytes sent to the OTHER ranks / time (max over ranks)."""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    n = int(os.environ.get("MRL_BENCH_N", "512"))
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ncp = (n // 2 + 1 + 7) // 8 * 8
    field = n * (n // world) * ncp            # complex128 elements of one half spectrum on this rank
    out = {"n": n, "world": world}
    for nf, name in ((2, "forward (2 spectra)"), (1, "return (1 spectrum)")):
        send = torch.zeros(nf * field, dtype=torch.complex128, device="cuda")
        recv = torch.empty_like(send)
        sv = torch.view_as_real(send).view(world, -1)
        rv = torch.view_as_real(recv).view(world, -1)
        for _ in range(3):
            dist.all_to_all_single(rv, sv)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            dist.all_to_all_single(rv, sv)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        egress = nf * field * 16 * (world - 1) / world
        out[name] = {"ms": round(ms, 4), "egress_mb_per_gpu": round(egress / 1e6, 1), "egress_gbs_per_gpu": round(egress / 1e9 / (ms / 1e3), 1)}
        del send, recv
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

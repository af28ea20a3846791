#!/usr/bin/env python
"""Peer-copy bandwidth between GPU 0 and GPU 1 with the copy engines, one direction and both directions at once
(development probe: what the NVLink path of this box delivers to a plain cudaMemcpyPeerAsync, next to the ~500 GB/s per
direction the fused exchange passes and NCCL's all-to-all reach).  One process, two devices.  Prints one JSON line."""
import json

import torch


def main():
    n = 1 << 28  # 256 Mi doubles? no: bytes below
    nbytes = 1 << 30
    a0 = torch.empty(nbytes, dtype=torch.uint8, device="cuda:0")
    b0 = torch.empty(nbytes, dtype=torch.uint8, device="cuda:0")
    a1 = torch.empty(nbytes, dtype=torch.uint8, device="cuda:1")
    b1 = torch.empty(nbytes, dtype=torch.uint8, device="cuda:1")
    s0 = torch.cuda.Stream(device="cuda:0")
    s1 = torch.cuda.Stream(device="cuda:1")
    out = {"bytes": nbytes, "p2p": torch.cuda.can_device_access_peer(0, 1)}

    def run(both, reps=5):
        for d in (0, 1):
            torch.cuda.synchronize(d)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s0):
            e0.record()
            for _ in range(reps):
                a1.copy_(a0, non_blocking=True)       # 0 -> 1, issued on GPU 0's stream
            e1.record()
        if both:
            with torch.cuda.stream(s1):
                for _ in range(reps):
                    b0.copy_(b1, non_blocking=True)   # 1 -> 0, issued on GPU 1's stream
        for d in (0, 1):
            torch.cuda.synchronize(d)
        return nbytes * reps / 1e9 / (e0.elapsed_time(e1) / 1e3)

    run(False, 2)
    out["one_direction_gbs"] = round(run(False), 1)
    run(True, 2)
    out["both_directions_gbs_per_direction"] = round(run(True), 1)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-pass device times of the fused Cahn-Hilliard substep (CUDA events through
mrl_split_substep_timed) for one kernel-variant selection (MRL_* environment variables).
Development tool: prints one JSON line."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from marlin_b200 import capi  # noqa: E402
from marlin_b200.capi import AB_BETA  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    dims = tuple(int(v) for v in os.environ["PT_DIMS"].split(",")) if "PT_DIMS" in os.environ else (n, n, n)
    prec = capi.F32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else capi.F64
    reps = 10
    L = n * 8 * math.pi / 200
    ctx = capi.Context(0, prec)
    ctx.use_torch_stream()
    ctx.domain_set(len(dims), dims, (0,) * len(dims), tuple(L * d / n for d in dims))
    torch.manual_seed(0)
    c = (torch.rand(dims, dtype=torch.float64) * 0.12 + 0.44).to(ctx.rdtype).cuda()
    plan = ctx.split_plan(double_well=(0.1, 0.0, 1.0), M_factor=0.2, L_factor=-0.001, history=1)
    dt = 1e-3
    plan.substep(c, dt, AB_BETA[0], 0)
    plan.advance_state()
    for _ in range(3):
        plan.substep(c, dt, AB_BETA[1], 1)
        plan.advance_state()
    torch.cuda.synchronize()
    chk = [float(c.double().sum()), float(c.double().square().sum())]
    acc = None
    for _ in range(reps):
        t = plan.substep_timed(c, dt, AB_BETA[1], 1)
        plan.advance_state()
        acc = t if acc is None else [a + b for a, b in zip(acc, t)]
    ms = [a / reps for a in acc]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        plan.substep(c, dt, AB_BETA[1], 1)
        plan.advance_state()
    e1.record()
    torch.cuda.synchronize()
    env = {k: v for k, v in os.environ.items() if k.startswith("MRL_")}
    print(json.dumps({"n": dims, "env": env, "pass_ms": [round(x, 4) for x in ms], "sum_ms": round(sum(ms), 4),
                      "step_ms": round(e0.elapsed_time(e1) / 20, 4), "checksum": chk}), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""MECH-3D-n (test/tests/mechanics/mech3d.i at n^3): one FFTMechanics substep on the GPU, with the
per-application time of the CG operator G(K4:x) and the achieved HBM rate against the
algorithmic bytes of SURVEY.md 8(d) (29 S_r + 72 S_c per application).  Prints one JSON line."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from marlin_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    L = 2 * math.pi
    ctx = capi.Context(0, capi.F64)
    ctx.use_torch_stream()
    ctx.domain_set(3, (n, n, n), (0,) * 3, (L,) * 3)
    ax = [ctx.axis(a).cuda() for a in range(3)]
    ph = ((torch.cos(ax[0]) / 2 + 0.5).view(n, 1, 1) * (torch.cos(ax[1]) / 2 + 0.5).view(1, n, 1) *
          (torch.cos(ax[2]) / 2 + 0.5).view(1, 1, n)).contiguous()
    K = ((1 - ph) * 1.0 + ph * 10.0).contiguous()
    mu = ((1 - ph) * 0.5 + ph * 5.0).contiguous()
    plan = capi.MechPlan(ctx, K, mu, l_tol=1e-2, nl_rel_tol=2e-2, nl_abs_tol=2e-2)
    F = torch.zeros(9, n, n, n, dtype=torch.float64, device="cuda")
    F[0] = F[4] = F[8] = 1.0
    x = torch.rand(9, n, n, n, dtype=torch.float64, device="cuda") - 0.5
    for _ in range(2):
        plan.apply_GK(F, x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    reps = 5
    for _ in range(reps):
        y = plan.apply_GK(F, x)
    e1.record()
    torch.cuda.synchronize()
    op_ms = e0.elapsed_time(e1) / reps
    s_r, s_c = n ** 3 * 8, n * n * (n // 2 + 1) * 16
    b_op = 29 * s_r + 72 * s_c
    # one substep of mech3d.i: applied shear 0.001 (sub-time of the second substep)
    applied = [0.0, 0.001, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]
    l0 = ctx.launch_count()
    e0.record()
    P, st = plan.solve(F, applied)
    e1.record()
    torch.cuda.synchronize()
    solve_ms = e0.elapsed_time(e1)
    its = st.cg_iterations_total
    print(json.dumps({
        "workload": f"MECH-3D-{n}", "GK_ms": round(op_ms, 3), "GK_alg_gb": round(b_op / 1e9, 3),
        "GK_gbs": round(b_op / 1e9 / (op_ms / 1e3), 1), "solve_ms": round(solve_ms, 2), "newton": st.newton_iterations,
        "cg_iterations": list(st.cg_iterations[:st.cg_solves]), "ms_per_cg_iteration": round(solve_ms / max(its, 1), 3),
        "launches": ctx.launch_count() - l0, "final_rnorm": st.final_rnorm,
        "Fmax": float(F.abs().max()), "Pnorm": float(torch.linalg.norm(P.reshape(-1)))}), flush=True)
    plan.close()
    ctx.close()


if __name__ == "__main__":
    main()

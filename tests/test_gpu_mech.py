"""GPU parity of the de Geus mechanics path (FFTMechanics + HyperElasticIsotropic + CG) against
the oracle and the reference's gold file.  Tolerance: relative L2 <= 1e-10 per field (fp64)."""
import math
import os

import numpy as np
import pytest
import torch

import oracle_cases as oc
from oracle import marlin as om

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def rel_l2(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b))


def soa(ctx, t):
    """oracle layout [nx,ny,nz,3,3] -> component-major [9,nx,ny,nz] on the device (through the C ABI)."""
    from marlin_b200 import capi
    g = t.contiguous().cuda()
    return capi.components(ctx, g, 9, True).view(9, *t.shape[:3])


def aos(ctx, t):
    from marlin_b200 import capi
    return capi.components(ctx, t.contiguous(), 9, False).view(*t.shape[1:], 3, 3).cpu()


@pytest.fixture(scope="module")
def ctx():
    from marlin_b200 import capi
    c = capi.Context(0, capi.F64)
    yield c
    c.close()


def _setup(ctx, n, ics_only=False):
    from marlin_b200 import capi
    p = oc.mech3d_problem(n=n, ics_only=ics_only)
    p.initial()
    L = 2 * math.pi
    ctx.domain_set(3, (n, n, n), (0,) * 3, (L,) * 3)
    K, mu = p.buf["K"].contiguous().cuda(), p.buf["mu"].contiguous().cuda()
    plan = capi.MechPlan(ctx, K, mu, l_tol=1e-2, nl_rel_tol=2e-2, nl_abs_tol=2e-2)
    return p, plan


@pytest.mark.parametrize("n", [16, 20, 32, 64, 128])
def test_mech_operators_match_oracle(ctx, n):
    """constitutive law, Green-operator projection and the CG operator on random states (128: the TMA-pipelined
    passes and the fused Green-projection pass k_mech_fused_tma, against the oracle's materialised Ghat4 / K4)."""
    p, plan = _setup(ctx, n)
    torch.manual_seed(4)
    F = torch.eye(3, dtype=torch.float64).expand(n, n, n, 3, 3) + 0.1 * torch.rand(n, n, n, 3, 3, dtype=torch.float64)
    x = torch.rand(n, n, n, 3, 3, dtype=torch.float64) - 0.5
    p.buf["Fnew"] = F
    p.mech.cm.compute()
    assert rel_l2(aos(ctx, plan.constitutive(soa(ctx, F))), p.buf["stress"]) < 1e-13
    d = p.domain
    Gx = d.ifft(om.ddot42(p.mech.Ghat4, d.fft(x)))
    assert rel_l2(aos(ctx, plan.apply_G(soa(ctx, x))), Gx) < 1e-12
    KdF = om.trans2(om.ddot42(p.buf[p.mech.K4], om.trans2(x)))
    GK = d.ifft(om.ddot42(p.mech.Ghat4, d.fft(KdF)))
    assert rel_l2(aos(ctx, plan.apply_GK(soa(ctx, F), soa(ctx, x))), GK) < 1e-12
    plan.close()


def test_mech3d_matches_gold_and_oracle_iterations(ctx):
    """test/tests/mechanics/mech3d.i: 3 steps x 10 substeps at 16^3 vs gold/mech3d.h5 and the
    oracle's CG / Newton iteration counts."""
    n = 16
    g = torch.from_numpy(np.load(f"{G}/mech3d_h5.npz")["F"])
    p, plan = _setup(ctx, n)
    F = soa(ctx, p.buf["F"])
    t, dt, substeps = 0.0, 0.01, 10
    for fr in range(g.shape[0]):
        p.step(dt)
        oracle_its = None
        for s in range(substeps):
            sub_time = t + s * dt / substeps
            # MacroscopicShearTensor (test/src/tensor_computes/MacroscopicShearTensor.C:31-41)
            avg = [ctx.reduce(0, F[c]) / n ** 3 for c in range(9)]
            applied = [(1.0 if c in (0, 4, 8) else 0.0) - avg[c] for c in range(9)]
            applied[1] += sub_time
            P, st = plan.solve(F, applied)
            oracle_its = (st.newton_iterations, list(st.cg_iterations[:st.cg_solves]))
        t += dt
        assert rel_l2(aos(ctx, F), g[fr]) < 1e-10, fr
        assert rel_l2(aos(ctx, F), p.buf["F"]) < 1e-10
        assert rel_l2(aos(ctx, P), p.buf["stress"]) < 1e-10
        assert oracle_its[0] == p.mech.newton_iterations and oracle_its[1] == p.mech.cg_iterations, (oracle_its, p.mech.cg_iterations)
    plan.close()


@pytest.mark.parametrize("n", [64, 128])
def test_mech3d_substep_matches_oracle_on_tma_sizes(ctx, n):
    """One MOOSE step of test/tests/mechanics/mech3d.i at n = 64 / 128 (two substeps: the first has a zero applied shear
    and returns at once, the second runs the full Newton-CG solve) against the oracle: F, stress, Newton and
    per-solve CG iteration counts.  128^3 runs every FFT pass, the fused Green projection and the 128-bit CG kernels
    the 256^3 bench line times."""
    p, plan = _setup(ctx, n)
    F = soa(ctx, p.buf["F"])
    dt, substeps = 0.02, 2
    p.solver.substeps = substeps
    p.step(dt)
    for s in range(substeps):
        sub_time = s * dt / substeps
        avg = [ctx.reduce(0, F[c]) / n ** 3 for c in range(9)]
        applied = [(1.0 if c in (0, 4, 8) else 0.0) - avg[c] for c in range(9)]
        applied[1] += sub_time
        P, st = plan.solve(F, applied)
    its = (st.newton_iterations, list(st.cg_iterations[:st.cg_solves]))
    assert rel_l2(aos(ctx, F), p.buf["F"]) < 1e-10
    assert rel_l2(aos(ctx, P), p.buf["stress"]) < 1e-10
    assert its[0] == p.mech.newton_iterations and its[1] == p.mech.cg_iterations, (its, p.mech.cg_iterations)
    assert sum(its[1]) > 20      # a real solve, not the trivial first substep
    plan.close()


def test_mech_256_operators_match_oracle(ctx):
    """BASELINE.json north_star size of the mechanics (256^3): constitutive law, G and G(K4:x) of the FFTCfg<256,...>
    TMA instantiations against the oracle evaluated WITHOUT materialising Ghat4 / K4 (11 GB each at this size):
    oracle.green_project_lowmem / tangent_apply_chunked, which tests/test_oracle_mech_lowmem.py pins to the
    materialised forms."""
    n = 256
    p, plan = _setup(ctx, n, ics_only=True)
    d = p.domain
    torch.manual_seed(4)
    F = (torch.eye(3, dtype=torch.float64).expand(n, n, n, 3, 3) + 0.1 * torch.rand(n, n, n, 3, 3, dtype=torch.float64)).contiguous()
    x = torch.rand(n, n, n, 3, 3, dtype=torch.float64) - 0.5
    Fs, xs = soa(ctx, F), soa(ctx, x)
    P_ref, KdF = om.tangent_apply_chunked(om.HyperElasticIsotropic, d, F, p.buf["K"], p.buf["mu"], x, chunk=16)
    assert rel_l2(aos(ctx, plan.constitutive(Fs)), P_ref) < 1e-13
    del P_ref
    assert rel_l2(aos(ctx, plan.apply_G(xs)), om.green_project_lowmem(d, x)) < 1e-12
    assert rel_l2(aos(ctx, plan.apply_GK(Fs, xs)), om.green_project_lowmem(d, KdF)) < 1e-12
    plan.close()


@pytest.mark.parametrize("n", [32, 20, 128])
def test_mech2d_operators_match_oracle(ctx, n):
    """2-D (2x2 tensors, test/tests/mechanics/mech.i): constitutive law, Green projection, CG operator."""
    from marlin_b200 import capi
    p = oc.mech3d_problem(n=n, dim=2)
    p.initial()
    L = 2 * math.pi
    ctx.domain_set(2, (n, n), (0,) * 2, (L,) * 2)
    plan = capi.MechPlan(ctx, p.buf["K"].contiguous().cuda(), p.buf["mu"].contiguous().cuda())
    to_soa = lambda t: capi.components(ctx, t.contiguous().cuda(), 4, True).view(4, n, n)
    to_aos = lambda t: capi.components(ctx, t.contiguous(), 4, False).view(n, n, 2, 2).cpu()
    torch.manual_seed(5)
    F = torch.eye(2, dtype=torch.float64).expand(n, n, 2, 2) + 0.1 * torch.rand(n, n, 2, 2, dtype=torch.float64)
    x = torch.rand(n, n, 2, 2, dtype=torch.float64) - 0.5
    p.buf["Fnew"] = F
    p.mech.cm.compute()
    assert rel_l2(to_aos(plan.constitutive(to_soa(F))), p.buf["stress"]) < 1e-13
    d = p.domain
    Gx = d.ifft(om.ddot42(p.mech.Ghat4, d.fft(x)))
    assert rel_l2(to_aos(plan.apply_G(to_soa(x))), Gx) < 1e-12
    KdF = om.trans2(om.ddot42(p.buf[p.mech.K4], om.trans2(x)))
    GK = d.ifft(om.ddot42(p.mech.Ghat4, d.fft(KdF)))
    assert rel_l2(to_aos(plan.apply_GK(to_soa(F), to_soa(x))), GK) < 1e-12
    plan.close()


def test_mech2d_matches_gold_and_oracle_iterations(ctx):
    """test/tests/mechanics/mech.i (2-D, 2x2 tensors, l_max_its = 40): 3 steps x 3 substeps at 32^2
    vs gold/mech.h5 and the oracle's CG / Newton iteration counts."""
    from marlin_b200 import capi
    n = 32
    g = torch.from_numpy(np.load(f"{G}/mech2d_h5.npz")["F"])
    p = oc.mech2d_problem()
    p.initial()
    ctx.domain_set(2, (n, n), (0,) * 2, (2 * math.pi,) * 2)
    plan = capi.MechPlan(ctx, p.buf["K"].contiguous().cuda(), p.buf["mu"].contiguous().cuda(), l_tol=1e-5, l_max_its=40,
                         nl_rel_tol=2e-4, nl_abs_tol=2e-3)
    to_aos = lambda t: capi.components(ctx, t.contiguous(), 4, False).view(n, n, 2, 2).cpu()
    F = capi.components(ctx, p.buf["F"].contiguous().cuda(), 4, True).view(4, n, n)
    t, dt, substeps = 0.0, 0.02, 3
    for fr in range(g.shape[0]):
        p.step(dt)
        its = None
        for s in range(substeps):
            sub_time = t + s * dt / substeps
            avg = [ctx.reduce(0, F[c]) / n ** 2 for c in range(4)]
            applied = [(1.0 if c in (0, 3) else 0.0) - avg[c] for c in range(4)]
            applied[1] += sub_time
            P, st = plan.solve(F, applied)
            its = (st.newton_iterations, list(st.cg_iterations[:st.cg_solves]))
        t += dt
        assert rel_l2(to_aos(F), g[fr]) < 1e-9, fr
        assert rel_l2(to_aos(F), p.buf["F"]) < 1e-9
        assert its[0] == p.mech.newton_iterations and its[1] == p.mech.cg_iterations, (its, p.mech.cg_iterations)
    plan.close()


@pytest.mark.parametrize("n", [128, 256])
def test_mech_large_properties(ctx, n):
    """Sizes on the TMA kernels (padded spectra): the Green operator is a projection on
    band-limited fields (G(G(A)) = G(A); the un-zeroed Nyquist planes of the reference's k-grid
    break this for full-spectrum input, there and here), linear, and annihilates constants; it
    equals its closed form evaluated with torch.fft; the CG operator is linear."""
    from marlin_b200 import capi
    L = 2 * math.pi
    ctx.domain_set(3, (n, n, n), (0,) * 3, (L,) * 3)
    torch.manual_seed(6)
    K = (1.0 + 9.0 * torch.rand(n, n, n, dtype=torch.float64)).cuda()
    mu = (0.5 + 4.5 * torch.rand(n, n, n, dtype=torch.float64)).cuda()
    plan = capi.MechPlan(ctx, K, mu)
    A = torch.rand(9, n, n, n, dtype=torch.float64, device="cuda") - 0.5
    B = torch.rand(9, n, n, n, dtype=torch.float64, device="cuda") - 0.5
    # band-limited field: keep |index| < n/4 on every axis
    kk = torch.fft.fftfreq(n, 1.0 / n, device="cuda").abs() < n / 4
    mask = kk.view(n, 1, 1) & kk.view(1, n, 1) & kk[:n // 2 + 1].view(1, 1, -1)
    S = torch.fft.irfftn(torch.fft.rfftn(A, dim=(1, 2, 3)) * mask, s=(n, n, n), dim=(1, 2, 3)).contiguous()
    GS = plan.apply_G(S)
    assert rel_l2(plan.apply_G(GS), GS) < 1e-12
    GA = plan.apply_G(A)
    assert rel_l2(plan.apply_G(2.0 * A - 3.0 * B), 2.0 * GA - 3.0 * plan.apply_G(B)) < 1e-12
    assert float(plan.apply_G(torch.ones_like(A)).abs().max()) < 1e-12
    # against torch.fft on the device for one tensor row: (G A)_0j = ifft( (A_0k q_k) q_j / |q|^2 )
    k = [ctx.axis(a, True).cuda() for a in range(3)]
    q = torch.stack(torch.meshgrid(k[0], k[1], k[2], indexing="ij"))
    Q = (q * q).sum(0)
    Ah = torch.fft.rfftn(A[0:3], dim=(1, 2, 3))
    v = (Ah * q).sum(0) / torch.where(Q == 0, torch.ones_like(Q), Q)
    v = torch.where(Q == 0, torch.zeros_like(v), v)
    ref = torch.fft.irfftn(v.unsqueeze(0) * q, s=(n, n, n), dim=(1, 2, 3))
    assert rel_l2(GA[0:3], ref) < 1e-12
    F = torch.zeros(9, n, n, n, dtype=torch.float64, device="cuda")
    F[0] = F[4] = F[8] = 1.0
    F += 0.05 * (torch.rand_like(F) - 0.5)
    lin = plan.apply_GK(F, 2.0 * A - 3.0 * B)
    assert rel_l2(lin, 2.0 * plan.apply_GK(F, A) - 3.0 * plan.apply_GK(F, B)) < 1e-11
    plan.close()


def test_mech_float32_tma_size():
    """floating_precision = SINGLE on the fused Green-projection pass (128^3): G(A) against its closed form
    evaluated with torch.fft in float64, and the CG operator's linearity, at float32 accuracy."""
    from marlin_b200 import capi
    n, L = 128, 2 * math.pi
    c32 = capi.Context(0, capi.F32)
    c32.domain_set(3, (n, n, n), (0,) * 3, (L,) * 3)
    torch.manual_seed(8)
    K = (1.0 + 9.0 * torch.rand(n, n, n)).cuda()
    mu = (0.5 + 4.5 * torch.rand(n, n, n)).cuda()
    plan = capi.MechPlan(c32, K, mu)
    A = (torch.rand(9, n, n, n, device="cuda") - 0.5)
    GA = plan.apply_G(A)
    k = [c32.axis(a, True).double().cuda() for a in range(3)]
    q = torch.stack(torch.meshgrid(k[0], k[1], k[2], indexing="ij"))
    Q = (q * q).sum(0)
    Ah = torch.fft.rfftn(A[0:3].double(), dim=(1, 2, 3))
    v = (Ah * q).sum(0) / torch.where(Q == 0, torch.ones_like(Q), Q)
    v = torch.where(Q == 0, torch.zeros_like(v), v)
    ref = torch.fft.irfftn(v.unsqueeze(0) * q, s=(n, n, n), dim=(1, 2, 3))
    assert rel_l2(GA[0:3], ref) < 2e-6
    F = torch.zeros(9, n, n, n, device="cuda")
    F[0] = F[4] = F[8] = 1.0
    B = torch.rand(9, n, n, n, device="cuda") - 0.5
    lin = plan.apply_GK(F, 2.0 * A - 3.0 * B)
    assert rel_l2(lin, 2.0 * plan.apply_GK(F, A) - 3.0 * plan.apply_GK(F, B)) < 1e-4
    plan.close()
    c32.close()


def test_mech_mixed_grid_falls_back_per_axis(ctx):
    """A grid whose last axis has the tangent-fused first pass (256) but whose x axis has no fused Green-projection
    configuration (48): the operator must finish with the separate passes, and agree with the fully un-fused path."""
    import os

    from marlin_b200 import capi
    shape = (48, 64, 256)
    ctx.domain_set(3, shape, (0,) * 3, (1.0, 2.0, 3.0))
    torch.manual_seed(9)
    K = (1.0 + torch.rand(shape, dtype=torch.float64)).cuda()
    mu = (0.5 + torch.rand(shape, dtype=torch.float64)).cuda()
    F = (torch.eye(3, dtype=torch.float64).reshape(9, 1, 1, 1) + 0.1 * torch.rand((9,) + shape, dtype=torch.float64)).contiguous().cuda()
    x = (torch.rand((9,) + shape, dtype=torch.float64) - 0.5).cuda()
    outs = []
    for fused in ("1", "0"):
        os.environ["MRL_MECH_TANGENT_FUSED"] = fused
        try:
            plan = capi.MechPlan(ctx, K, mu, l_tol=1e-2, nl_rel_tol=2e-2, nl_abs_tol=2e-2)
            outs.append(plan.apply_GK(F, x.clone()).clone())
            plan.close()
        finally:
            os.environ.pop("MRL_MECH_TANGENT_FUSED", None)
    assert float((outs[0] - outs[1]).abs().max() / outs[1].abs().max()) < 1e-13

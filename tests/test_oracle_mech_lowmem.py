"""The memory-lean forms of the oracle's mechanics operators (used for the 256^3 GPU parity check, where the
reference's materialised Ghat4 / K4 would need 11 GB each) against the materialised forms that restate
FFTMechanics.C:74-84,104-110 and HyperElasticIsotropic.C:42-52 and are pinned to gold mech3d.h5 / mech.h5."""
import pytest
import torch

import oracle_cases as oc
from oracle import marlin as om


@pytest.mark.parametrize("dim,n", [(3, 12), (3, 9), (2, 16)])
def test_lowmem_operators_equal_materialised(dim, n):
    p = oc.mech3d_problem(n=n, dim=dim)
    p.initial()
    d = p.domain
    torch.manual_seed(1)
    shp = d.value_shape([dim, dim])
    F = torch.eye(dim, dtype=torch.float64).expand(shp) + 0.1 * torch.rand(shp, dtype=torch.float64)
    x = torch.rand(shp, dtype=torch.float64) - 0.5
    p.buf["Fnew"] = F
    p.mech.cm.compute()
    G_ref = d.ifft(om.ddot42(p.mech.Ghat4, d.fft(x)))
    G_low = om.green_project_lowmem(d, x)
    assert float((G_low - G_ref).abs().max()) < 1e-14 * float(G_ref.abs().max()) + 1e-15
    K_ref = om.trans2(om.ddot42(p.buf[p.mech.K4], om.trans2(x)))
    P_low, K_low = om.tangent_apply_chunked(om.HyperElasticIsotropic, d, F.contiguous(), p.buf["K"], p.buf["mu"], x, chunk=5)
    assert torch.equal(P_low, p.buf["stress"])
    assert float((K_low - K_ref).abs().max()) <= 1e-15 * float(K_ref.abs().max())

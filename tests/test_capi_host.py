"""CPU-side checks of the C ABI: the library loads, exports every declared symbol, and its
host-only arithmetic (axes) is bit-exact against ATen.  No device compute here."""
import ctypes
import math
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from marlin_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        subprocess.check_call(["make", "-j8"], cwd=ROOT)
    return capi.lib()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "marlin_b200.h")).read()
    names = set(re.findall(r"\b(mrl_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 20
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_create_fails_loudly_without_gpu(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.mrl_create(0, 0, ctypes.byref(h))
    assert rc != 0 and b"no CPU path" in lib.mrl_last_error()


@pytest.mark.parametrize("n,mn,mx", [(20, 0.0, 3.0), (200, 0.0, 8 * math.pi), (512, 0.0, 512 * 8 * math.pi / 200),
                                     (11, -1.0, 2.5), (1, 0.0, 1.0), (150, 0.0, 2 * math.pi), (13, 0.0, 7.0)])
def test_axes_bit_exact_vs_aten(lib, n, mn, mx):
    """Grid contract (src/actions/DomainAction.C:241-293): cell-centred linspace axis and
    2*pi*fftfreq / 2*pi*rfftfreq reciprocal axes, compared bit for bit with ATen."""
    from marlin_b200 import capi
    dx = (mx - mn) / n
    ax = torch.linspace(mn + dx / 2.0, mx - dx / 2.0, n, dtype=torch.float64)
    mine = torch.tensor(capi.axis_values(n, mn, mx), dtype=torch.float64)
    # ATen's linspace rounds differently per CPU vector width / device (it runs on the
    # compute device in the reference), so the cell centres agree to 1 ulp, not bit for bit
    assert ((mine - ax).abs() <= 2.3e-16 * ax.abs().clamp(min=1e-300)).all()
    assert mine[0] == ax[0] and mine[-1] == ax[-1]
    k = torch.fft.fftfreq(n, dx, dtype=torch.float64) * 2.0 * math.pi
    assert capi.axis_values(n, mn, mx, reciprocal=True) == k.tolist()
    kh = torch.fft.rfftfreq(n, dx, dtype=torch.float64) * 2.0 * math.pi
    assert capi.axis_values(n, mn, mx, reciprocal=True, half=True) == kh.tolist()


def test_emulated_kernels():
    """Host emulation of the FFT kernels vs a long-double DFT (index math / barriers)."""
    subprocess.check_call(["make", "emu"], cwd=ROOT)
    out = subprocess.run([os.path.join(ROOT, "tests/emu/_build/emu_fft_test")], capture_output=True, text=True)
    assert out.returncode == 0 and "EMU TESTS PASSED" in out.stdout, out.stdout[-2000:]


def _partition_ref(total, weights):
    """DomainAction::partitionHepler restated (include/actions/DomainAction.h:249-280)."""
    ns, rem = [], sum(weights)
    for w in weights:
        n = max((total * w) // rem, 1)
        ns.append(n)
        rem -= w
        total -= n
    ns[-1] += total
    return ns


@pytest.mark.parametrize("total,weights", [(20, [1, 1, 1]), (512, [1] * 8), (10, [1, 2, 3, 4]), (7, [1, 1]), (40, [3, 1, 1]),
                                           (257, [1] * 4), (5, [1] * 5)])
def test_partition_matches_reference_rule(lib, total, weights):
    from marlin_b200 import capi
    got = capi.partition(total, len(weights), weights)
    assert got == _partition_ref(total, weights) and sum(got) == total
    assert capi.partition(total, len(weights)) == _partition_ref(total, [1] * len(weights))


def test_pencil_factorisation_rule():
    """partitionPencils' choice of Py x Pz (src/actions/DomainAction.C:574-613): both factors > 1, fitting the domain
    (Py <= ny and <= nx/2+1, Pz <= nz and <= ny), smallest |Py - Pz|, the first pair found for d = 2 .. sqrt(ranks) in the
    order (d, ranks/d), (ranks/d, d).  Host-only arithmetic."""
    import ctypes as C

    from marlin_b200 import capi

    def factors(nranks, n):
        py, pz = C.c_int(), C.c_int()
        rc = capi.lib().mrl_pencil_factors(nranks, (C.c_int64 * 3)(*n), C.byref(py), C.byref(pz))
        return (py.value, pz.value) if rc == 0 else capi.lib().mrl_last_error().decode()

    assert factors(4, (40, 40, 40)) == (2, 2)
    assert factors(8, (40, 40, 40)) == (2, 4)       # cost 2 either way: the first pair considered wins
    assert factors(6, (40, 40, 40)) == (2, 3)
    assert factors(16, (64, 64, 64)) == (4, 4)
    assert factors(12, (64, 64, 64)) == (3, 4)      # d = 2: (2, 6) cost 4; d = 3: (3, 4) cost 1
    assert factors(8, (4, 40, 3)) == (2, 4) or "FFT_PENCIL requires" in str(factors(8, (4, 40, 3)))
    assert factors(8, (40, 40, 3)) == (4, 2)        # Pz = 4 does not fit nz = 3
    for bad in (1, 2, 3, 5, 7):
        assert "FFT_PENCIL requires factoring the number of MPI ranks" in str(factors(bad, (40, 40, 40)))

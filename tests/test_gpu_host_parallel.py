"""[Domain] parallel_mode = FFT_SLAB through the stand-alone host driver, one process per rank, with the cli_args of the
reference's parallel test specs (test/tests/cahnhilliard/tests [xdmf_output_hdf5_parallel], test/tests/gradient/tests
[gradient_cpu_slab], test/tests/tensor_compute/parallel_roundtrip.i) against the reference's gold files.

The ranks find each other through the torchrun-style environment (host/shim/comm.h) and exchange field data GPU to GPU
through CUDA IPC (mrl_dist_*).  When the box has fewer GPUs than ranks the ranks share devices - slow (the device-side
barriers then wait for the driver's time slicing) but the same code path, so these tests also run on a one-GPU box."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "marlin_b200", "marlin_b200-opt")
REF = os.path.join(ROOT, "tests", "inputs", "ref")
G = os.path.join(ROOT, "tests", "golden")


def launch(tmp, nranks, inp, *args, dump=(), inp_dir=REF):
    """mpiexec -n nranks marlin-opt -i inp args: one process per rank.  A rendezvous that cannot bind its port (left in
    use by something else on the box) is retried once on another port."""
    for attempt in range(2):
        port = 29600 + (os.getpid() * 7 + nranks + 131 * attempt) % 300
        procs = []
        for r in range(nranks):
            env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(nranks), MASTER_ADDR="127.0.0.1", MRL_COMM_PORT=str(port))
            cmd = [APP, "-i", f"{inp_dir}/{inp}", "--output-dir", str(tmp), "--compute-device=cuda", *args]
            if dump:
                cmd += ["--dump", ",".join(dump), "--dump-dir", str(tmp)]
            procs.append(subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        outs = []
        try:
            for p in procs:
                outs.append(p.communicate(timeout=600))
        finally:
            for p in procs:
                if p.poll() is None:
                    p.kill()
        if attempt == 0 and any("Comm: bind to port" in se for _, se in outs):
            continue
        break
    for r, (p, (so, se)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r}:\n{so}\n{se}"
    return outs


def csv(path):
    with open(path) as fh:
        head = fh.readline().strip().split(",")
        rows = np.array([[float(x) for x in ln.split(",")] for ln in fh if ln.strip()])
    return head, rows


def test_cahnhilliard_two_rank_slab_gold(tmp_path):
    """[xdmf_output_hdf5_parallel]: cahnhilliard.i on two ranks; rank 1's data files against gold cahnhilliard.rank0001.h5
    (HDF5Diff abs_tol 1e-13 in the reference's spec), the XMF a spatial collection with one sub-grid per rank."""
    import xml.etree.ElementTree as ET
    g = np.load(f"{G}/ch2d_slab_rank1_h5.npz")["c"]
    launch(tmp_path, 2, "cahnhilliard.i", 'TensorOutputs/active=xdmf2', "Domain/parallel_mode=FFT_SLAB", "Domain/device_names=cpu")
    import h5lite
    h1, h0 = h5lite.H5File(f"{tmp_path}/cahnhilliard.rank0001.h5"), h5lite.H5File(f"{tmp_path}/cahnhilliard.rank0000.h5")
    assert h1.keys() == sorted(f"c.{f}" for f in range(11))
    for frame in range(11):
        got = h1.read(f"c.{frame}")
        assert got.shape == (20, 10) and np.abs(got - g[frame]).max() < 1e-13, frame
        assert np.abs(h0.read(f"c.{frame}") - got).max() < 1e-13      # identical random blocks stay identical
    root = ET.parse(f"{tmp_path}/cahnhilliard.xmf").getroot()
    series = root.find("Domain").find("Grid")
    frames = series.findall("Grid")
    assert len(frames) == 11 and all(f.get("GridType") == "Collection" and f.get("CollectionType") == "Spatial" for f in frames)
    subs = frames[3].findall("Grid")
    assert [s.get("Name") for s in subs] == ["Rank0", "Rank1"]
    assert subs[1].find("Topology").get("Dimensions") == "21 11"
    assert subs[1].find("Geometry").findall("DataItem")[0].text == "0 1.5"       # origin of rank 1's part: y = 10 * 0.15
    item = subs[1].find("Attribute").find("DataItem")
    assert item.get("Dimensions") == "20 10" and item.get("Format") == "HDF" and item.text.endswith("cahnhilliard.rank0001.h5:/c.3")
    assert not os.path.exists(f"{tmp_path}/cahnhilliard.rank0001.xmf")
    # the CSV is rank 0's; its postprocessors are gathered over the ranks
    head, rows = csv(f"{tmp_path}/cahnhilliard_out.csv")
    assert rows.shape[0] == 11
    # min_c (SemiImplicitCriticalTimeStep: gatherMin over the ranks' k-space parts) equals the serial run's
    ser = tmp_path / "serial"
    ser.mkdir()
    launch(ser, 1, "cahnhilliard.i")
    head1, rows1 = csv(f"{ser}/cahnhilliard_out.csv")
    assert np.abs(rows[:, head.index("min_c")] - rows1[:, head1.index("min_c")]).max() < 1e-15


def test_fft_slab_on_one_rank_equals_serial(tmp_path):
    """parallel_mode = FFT_SLAB with a single process takes the distributed code path (mrl_dist_* with one rank) and must
    reproduce the serial run."""
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(), b.mkdir()
    launch(a, 1, "cahnhilliard.i", dump=("c",))
    launch(b, 1, "cahnhilliard.i", "Domain/parallel_mode=FFT_SLAB", dump=("c",))
    ca, cb = np.fromfile(f"{a}/c.f64"), np.fromfile(f"{b}/c.f64")
    assert np.abs(ca - cb).max() < 1e-13


@pytest.mark.parametrize("nranks,weights", [(3, "1 1 1"), (2, "3 1")])
def test_gradient_slab_csv_gold(tmp_path, nranks, weights):
    """[gradient_cpu_slab]: gradient.i (40^3, three FFTGradients against analytic derivatives) on three ranks - 40 does
    not divide by 3: the remainder goes to the last rank (partitionHepler) - and on two unequally weighted ranks."""
    gold = np.load(f"{G}/csv_golds.npz")["gradient_out"]
    names = " ".join(["cpu"] * nranks)
    launch(tmp_path, nranks, "gradient.i", f"Domain/device_names={names}", f"Domain/device_weights={weights}", "Domain/parallel_mode=FFT_SLAB")
    head, rows = csv(f"{tmp_path}/gradient_out.csv")
    assert rows.shape == gold.shape
    # CSVDiff default tolerances (rel_err 5.5e-6, abs_zero 1e-10)
    assert np.all(np.abs(rows - gold) <= 5.5e-6 * np.abs(gold) + 1e-10)


def test_gradient_pencil_csv_gold(tmp_path):
    """[gradient_cpu_pencil]: gradient.i on four ranks with parallel_mode = FFT_PENCIL (2 x 2 pencils; reciprocal space
    holds the half spectrum on x, split along kx and ky)."""
    gold = np.load(f"{G}/csv_golds.npz")["gradient_out"]
    launch(tmp_path, 4, "gradient.i", "Domain/device_names=cpu cpu cpu cpu", "Domain/device_weights=1 1 1 1", "Domain/parallel_mode=FFT_PENCIL")
    head, rows = csv(f"{tmp_path}/gradient_out.csv")
    assert rows.shape == gold.shape
    assert np.all(np.abs(rows - gold) <= 5.5e-6 * np.abs(gold) + 1e-10)


def test_pencil_needs_a_factorisation(tmp_path):
    """partitionPencils (DomainAction.C:606-613): two ranks cannot be factored into two integers greater than one."""
    with pytest.raises(AssertionError, match="FFT_PENCIL requires factoring the number of MPI ranks"):
        launch(tmp_path, 2, "gradient.i", "Domain/parallel_mode=FFT_PENCIL")


def test_slab_roundtrip_three_ranks(tmp_path):
    """The check of test/tests/tensor_compute/parallel_roundtrip.i (a 2-D field through fftSlab / ifftSlab on three ranks,
    128 = 42 + 42 + 44, returns unchanged; that file reads a [Solve] output at INITIAL and is not part of the reference's
    test specs) with the transforms in [Initialize]."""
    launch(tmp_path, 3, "slab_roundtrip.i", inp_dir=os.path.join(ROOT, "tests", "inputs"))
    head, rows = csv(f"{tmp_path}/slab_roundtrip_out.csv")
    assert abs(rows[-1, head.index("max_error")]) < 1e-13
    assert abs(rows[-1, head.index("l2_error")]) < 1e-10


def test_fused_slab_plan_through_the_host_solver(tmp_path):
    """examples/cahn_hilliard/cahnhilliard2.i (the headline workload's file) at 128^3 with parallel_mode = FFT_SLAB on two
    ranks: the host AdamsBashforthMoulton recognises the split-operator graph and runs the multi-GPU fused plan
    (mrl_slab_*: exchanges fused into the passes, the file's ParsedCompute nonlinearity compiled into the first pass);
    the same run with fuse = false goes operator by operator over mrl_dist_rfftn / mrl_dist_irfftn - the path the 2-rank
    gold above pins.  Both start from the same per-rank random block."""
    import math
    n = 128
    L = n * (8 * math.pi / 200)
    args = ["--allow-unused", f"Domain/nx={n}", f"Domain/ny={n}", f"Domain/nz={n}", f"Domain/xmax={L!r}", f"Domain/ymax={L!r}",
            f"Domain/zmax={L!r}", "Executioner/num_steps=2", "TensorComputes/Initialize/c/seed=0", "Domain/parallel_mode=FFT_SLAB",
            "TensorSolver/substeps=25", "Problem/print_debug_output=true"]   # 2 x 25 substeps: AB1 start-up + steady AB2
    a, b = tmp_path / "fused", tmp_path / "generic"
    a.mkdir(), b.mkdir()
    outs = launch(a, 2, "cahnhilliard2.i", *args, dump=("c",))
    assert "fused slab-decomposed plan" in outs[0][0] + outs[0][1]
    outs = launch(b, 2, "cahnhilliard2.i", *args, "TensorSolver/fuse=false", dump=("c",))
    assert "operator-by-operator path" in outs[0][0] + outs[0][1]
    for r in range(2):
        ca = np.fromfile(f"{a}/c.rank{r:04d}.f64")
        cb = np.fromfile(f"{b}/c.rank{r:04d}.f64")
        assert ca.size == n * (n // 2) * n
        assert np.linalg.norm(ca - cb) / np.linalg.norm(cb) < 1e-10


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("nranks,mode", [(2, "FFT_SLAB"), (4, "FFT_PENCIL")])
def test_mech3d_decomposed_matches_gold(tmp_path, nranks, mode):
    """test/tests/mechanics/mech3d.i (de Geus finite-strain mechanics: Newton + CG with the Green projection) on a
    decomposed domain - FFTMechanics over DomainAction::fft / ifft in FFT_SLAB / FFT_PENCIL mode, inner products summed
    over the ranks - against the first frame of the reference's serial gold mech3d.h5 (one time step: the ranks share
    the GPU here, and every transform is a handful of inter-process barriers).  ComputeDisplacements (nodal output) is
    switched off: nodal fields do not exist on a decomposed domain.

    FFT_SLAB keeps the serial mode's spectral layout (half spectrum on z) and reproduces the gold to round-off.  FFT_PENCIL
    uses the reference's pencil layout - half spectrum on x, rfftfreq there and fftfreq on z (gridChanged :289-306) - so the
    Nyquist wavevectors of x and z carry the opposite sign of the serial mode's; the Green projection q_i q_j / |q|^2 is
    odd in each component, and the solve differs from the serial gold at the 3e-6 level on this 16^3 grid (the
    transforms themselves are exact: tests/test_dist_gpu.py)."""
    if mode == "FFT_PENCIL" and _gpus() < 4:
        # ~40 CG iterations x 9 fields x two-stage exchanges = thousands of inter-process barriers; with four processes
        # time-slicing one GPU that takes minutes.  Runs where every rank has its own GPU (profiles/r2*_8gpu logs).
        pytest.skip("pencil-decomposed mechanics needs four GPUs to run in reasonable time")
    g = np.load(f"{G}/mech3d_h5.npz")
    launch(tmp_path, nranks, "mech3d.i", f"Domain/parallel_mode={mode}", "Executioner/num_steps=1", "TensorComputes/Postprocess/active=vonmises",
           "TensorOutputs/deformation_tensor/buffer=sV F", "TensorOutputs/deformation_tensor/output_mode=CELL CELL", dump=("F", "sV"))
    n = 16
    F = np.zeros((9, n, n, n))
    sV = np.zeros((n, n, n))
    import ctypes

    from marlin_b200 import capi
    for r in range(nranks):
        # the rank's real-space part, from the library's own partition rule (host only)
        if mode == "FFT_SLAB":
            cnt = (ctypes.c_int64 * nranks)()
            capi._ck(capi.lib().mrl_partition(ctypes.c_int64(n), nranks, None, cnt))
            y0, ny, z0, nz = sum(cnt[:r]), cnt[r], 0, n
        else:
            py, pz = 2, 2
            y0, ny, z0, nz = (r % py) * (n // py), n // py, (r // py) * (n // pz), n // pz
        F[:, :, y0:y0 + ny, z0:z0 + nz] = np.fromfile(f"{tmp_path}/F.rank{r:04d}.f64").reshape(9, n, ny, nz)
        sV[:, y0:y0 + ny, z0:z0 + nz] = np.fromfile(f"{tmp_path}/sV.rank{r:04d}.f64").reshape(n, ny, nz)
    ref = np.moveaxis(g["F"][0].reshape(n, n, n, 9), -1, 0)
    tol = 1e-9 if mode == "FFT_SLAB" else 2e-5
    assert np.linalg.norm(F - ref) / np.linalg.norm(ref) < tol
    # the strains are ~1e-2 of F, so the von Mises stress sees the pencil convention 100 x more strongly
    assert np.abs(sV - g["sV"][0]).max() < (1e-9 if mode == "FFT_SLAB" else 2e-3) * np.abs(g["sV"][0]).max()

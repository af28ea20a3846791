"""The reference's own input files, VERBATIM (tests/inputs/ref/*.i are byte-for-byte copies of the files named
in each test; tests/test_host_cpu.py checks that against /root/reference where it exists), run through the
stand-alone host driver with the cli_args of the reference's `tests` specs, against the reference's gold
results (tests/golden/*.npz) or the oracle.  "Existing input files drop in unchanged" proven for a run,
not only for the reader.  GPU only."""
import math
import os
import subprocess

import numpy as np
import pytest

import oracle_cases as oc
from oracle import marlin as om

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "marlin_b200", "marlin_b200-opt")
REF = os.path.join(ROOT, "tests", "inputs", "ref")
G = os.path.join(ROOT, "tests", "golden")


def run(tmp, inp, *args, dump=()):
    cmd = [APP, "-i", f"{REF}/{inp}", "--output-dir", str(tmp), "--compute-device=cuda", *args]
    if dump:
        cmd += ["--dump", ",".join(dump), "--dump-dir", str(tmp)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    return r


def field(tmp, name, shape):
    return np.fromfile(f"{tmp}/{name}.f64", dtype=np.float64).reshape(shape)


def csv(path):
    with open(path) as fh:
        head = fh.readline().strip().split(",")
        rows = np.array([[float(x) for x in ln.split(",")] for ln in fh if ln.strip()])
    return head, rows


def test_cahnhilliard_i_map_to_aux_2d(tmp_path):
    """test/tests/cahnhilliard/tests [map_to_aux_2d]: cahnhilliard.i -> gold cahnhilliard_out.e."""
    g = np.load(f"{G}/ch2d_exodus.npz")
    run(tmp_path, "cahnhilliard.i", "TensorBuffers/c/map_to_aux_variable=c", "TensorBuffers/mu/map_to_aux_variable=mu",
        dump=("c", "mu"))
    assert np.abs(field(tmp_path, "c", (20, 20)) - g["c"][10]).max() < 1e-12
    assert np.abs(field(tmp_path, "mu", (20, 20)) - g["mu"][10]).max() < 1e-12
    head, rows = csv(f"{tmp_path}/cahnhilliard_out.csv")
    assert head == ["time", "delta_int_c", "min_c"] and rows.shape[0] == 11


def test_cahnhilliard_i_map_to_aux_3d(tmp_path):
    """[map_to_aux_3d]: the same file with the 3-D cli_args -> gold map_to_aux_3d.e."""
    g = np.load(f"{G}/ch3d_map_to_aux_exodus.npz")
    run(tmp_path, "cahnhilliard.i", "TensorBuffers/c/map_to_aux_variable=c", "TensorBuffers/mu/map_to_aux_variable=mu",
        "Domain/dim=3", "Domain/nx=5", "Domain/ny=5", "Domain/nz=5", "Domain/zmax=3", "Outputs/file_base=map_to_aux_3d",
        dump=("c", "mu"))
    assert np.abs(field(tmp_path, "c", (5, 5, 5)) - g["c"][10]).max() < 1e-12
    assert np.abs(field(tmp_path, "mu", (5, 5, 5)) - g["mu"][10]).max() < 1e-12
    assert os.path.exists(f"{tmp_path}/map_to_aux_3d.csv")


def test_cahnhilliard_i_xdmf(tmp_path):
    """[xdmf_output_xml] / [xdmf_output_hdf5]: TensorOutputs/active="xdmf" -> gold cahnhilliard.xmf / cahnhilliard.h5
    (the NODE / CELL arrays of the gold HDF5 file are tests/golden/ch2d_xdmf_h5.npz)."""
    import re
    run(tmp_path, "cahnhilliard.i", 'TensorOutputs/active=xdmf')
    gold = open(f"{G}/cahnhilliard_gold.xmf").read()
    mine = open(f"{tmp_path}/cahnhilliard.xmf").read()
    # XMLDiff: the document equals the gold one, HDF storage included, once the output directory is taken off the file names
    assert mine.replace(f"{tmp_path}/", "") == gold
    # HDF5Diff (abs_tol 1e-13 in the reference's spec): every dataset of the gold cahnhilliard.h5
    import h5lite
    g = np.load(f"{G}/ch2d_xdmf_h5.npz")
    h = h5lite.H5File(f"{tmp_path}/cahnhilliard.h5")
    frames = [int(f) for f in g["frames"]]
    assert sorted(h.keys()) == sorted([f"c.{f}" for f in range(11)] + [f"mu.{f}" for f in range(11)])
    for i, f in enumerate(frames):
        assert np.abs(h.read(f"c.{f}") - g["c_node"][i]).max() < 1e-13
        assert np.abs(h.read(f"mu.{f}") - g["mu_cell"][i]).max() < 1e-13
    d = h.info("c.0")
    assert d["shape"] == (21, 21) and d["filters"] == [(1, (9,))] and d["chunk"] == (21, 21, 8)


def test_cahnhilliard_explicit_i(tmp_path):
    """[explicit_euler_exodiff]: cahnhilliard_explicit.i -> gold cahnhilliard_explicit_out.e (last frame)."""
    g = np.load(f"{G}/ch2d_explicit_exodus.npz")
    frames = list(g["frames"])
    run(tmp_path, "cahnhilliard_explicit.i", dump=("c", "mu"))
    k = len(frames) - 1
    assert np.abs(field(tmp_path, "c", (50, 50)) - g["c"][k]).max() < 1e-10
    assert np.abs(field(tmp_path, "mu", (50, 50)) - g["mu"][k]).max() < 1e-10


@pytest.mark.parametrize("inp,ss,cs,order", [("diagonal", 10, 0, 2), ("diagonal", 20, 0, 4), ("diagonal", 10, 2, 2),
                                             ("coupled", 10, 0, 2), ("coupled", 10, 2, 2),
                                             ("nl_coupled", 10, 0, 3), ("nl_coupled", 10, 1, 1)])
def test_solver_inputs_match_csv_golds(tmp_path, inp, ss, cs, order):
    """test/tests/solvers/tests: diagonal.i / coupled.i / nl_coupled.i with cli_args 'ss= cs= order='."""
    gold = np.load(f"{G}/csv_golds.npz")[f"{inp}_{ss}_{cs}_{order}"]
    run(tmp_path, f"{inp}.i", f"ss={ss}", f"cs={cs}", f"order={order}")
    head, rows = csv(f"{tmp_path}/{inp}_{ss}_{cs}_{order}.csv")
    assert head == ["time", "U", "V", "u_max", "u_min", "v_max", "v_min"]
    assert rows.shape == gold.shape
    err = np.abs(rows - gold) / np.maximum(np.abs(gold), 1e-4 if inp != "diagonal" else 1e-8)
    assert err[1:].max() < 1e-9, err.max()


def test_etdrk4_diffusion_i(tmp_path):
    """[etdrk4_diffusion]: etdrk4_diffusion.i with cli_args 'ss=1 dt=10.0' -> gold etdrk4_diffusion_rmse.csv."""
    gold = np.load(f"{G}/csv_golds.npz")["etdrk4_diffusion_rmse"]
    run(tmp_path, "etdrk4_diffusion.i", "ss=1", "dt=10.0")
    head, rows = csv(f"{tmp_path}/etdrk4_diffusion_rmse.csv")
    assert head[:2] == ["time", "mse"] and rows.shape[0] == gold.shape[0]
    assert np.abs(rows[:, 0] - gold[:, 0]).max() < 1e-12
    assert np.abs(rows[:, 1] - gold[:, 1]).max() < 1e-12


def test_mech3d_i(tmp_path):
    """test/tests/mechanics/mech3d.i -> gold mech3d.h5 (F, sV, disp of the last frame)."""
    g = np.load(f"{G}/mech3d_h5.npz")
    run(tmp_path, "mech3d.i", dump=("F", "sV", "disp"))
    F = field(tmp_path, "F", (9, 16, 16, 16))
    ref = np.moveaxis(g["F"][2].reshape(16, 16, 16, 9), -1, 0)
    assert np.linalg.norm(F - ref) / np.linalg.norm(ref) < 1e-9
    assert np.abs(field(tmp_path, "sV", (16, 16, 16)) - g["sV"][2]).max() < 1e-9 * np.abs(g["sV"][2]).max()
    assert np.abs(field(tmp_path, "disp", (3, 17, 17, 17)) - np.moveaxis(g["disp"][2], -1, 0)).max() < 1e-10


def test_mech_i(tmp_path):
    """test/tests/mechanics/mech.i (2-D) -> gold mech.h5."""
    g = np.load(f"{G}/mech2d_h5.npz")
    run(tmp_path, "mech.i", dump=("F", "sV"))
    F = field(tmp_path, "F", (4, 32, 32))
    ref = np.moveaxis(g["F"][2].reshape(32, 32, 4), -1, 0)
    assert np.linalg.norm(F - ref) / np.linalg.norm(ref) < 1e-9
    assert np.abs(field(tmp_path, "sV", (32, 32)) - g["sV"][2]).max() < 1e-9 * np.abs(g["sV"][2]).max()


def test_gradient_i(tmp_path):
    """test/tests/gradient/gradient.i -> gold gradient_out.csv: the L1 error of the spectral gradient of an
    analytic field is at round-off level (7.6e-12 in the gold file)."""
    run(tmp_path, "gradient.i")
    head, rows = csv(f"{tmp_path}/gradient_out.csv")
    assert head == ["time", "diff"]
    assert 0 <= rows[-1, 1] < 1e-10


def test_rotating_grain_secant_i(tmp_path):
    """test/tests/tensor_compute/rotating_grain_secant.i -> gold rotating_grain_secant.h5 (abs_tol 1e-10)."""
    g = np.load(f"{G}/rotating_grain_secant_h5.npz")["psi"]
    run(tmp_path, "rotating_grain_secant.i", dump=("psi",))
    assert np.abs(field(tmp_path, "psi", (40, 40)) - g[-1]).max() < 1e-10


def test_kks_no_flux_bc_i(tmp_path):
    """test/tests/kks/KKS_no_flux_bc.i -> gold KKS_no_flux_bc.h5 / KKS_no_flux_bc_out.csv."""
    g = np.load(f"{G}/kks_no_flux_bc.npz")
    run(tmp_path, "KKS_no_flux_bc.i", dump=("c", "eta", "mu"))
    for k in ("c", "eta", "mu"):
        assert np.abs(field(tmp_path, k, (20, 20)) - g[k][10]).max() < 1e-9, k
    head, rows = csv(f"{tmp_path}/KKS_no_flux_bc_out.csv")
    assert head == ["time", "total_C", "total_eta"] and rows.shape == g["csv"].shape
    assert (np.abs(rows - g["csv"]) / np.maximum(np.abs(g["csv"]), 1.0)).max() < 1e-9


def test_postprocessors_i(tmp_path):
    """test/tests/postprocessors/postprocessors.i (num_steps = 0) -> gold average / integral / extreme_value /
    reciprocal_integral CSVs (0.8, 4.8, 3.2375 / -1.6375, 4.8)."""
    run(tmp_path, "postprocessors.i")
    head, rows = csv(f"{tmp_path}/postprocessors_out.csv")
    col = {h: rows[:, i] for i, h in enumerate(head)}
    assert np.abs(col["avg_c"] - 0.8).max() < 1e-13 and np.abs(col["int_c"] - 4.8).max() < 1e-12
    assert np.abs(col["max_c"] - 3.2375).max() < 1e-13 and np.abs(col["min_c"] + 1.6375).max() < 1e-13


def test_1a_solver_i(tmp_path):
    """benchmarks/01_spinodal_decomposition/1a_solver.i (PFHub BM1a, BASELINE.json configs[1]) vs the oracle: 2 steps x
    1000 substeps.  The file carries two stale parameters (`TensorSolver/history_size`, `Problem/spectral_solve_substeps`)
    that no class declares any more: like the reference (MooseApp.C:477, ERROR_UNUSED by default) the driver rejects them
    unless --allow-unused is given."""
    r = subprocess.run([APP, "-i", f"{REF}/1a_solver.i", "--output-dir", str(tmp_path), "Executioner/num_steps=1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "unused parameter" in r.stderr and "--allow-unused" in r.stderr
    run(tmp_path, "1a_solver.i", "--allow-unused", "Executioner/num_steps=2", dump=("c",))
    p = oc.bm1_problem()
    p.initial()
    dt = 1.0
    for _ in range(2):
        p.step(dt)
        dt *= 1.1
    ref = p.buf["c"].numpy()
    got = field(tmp_path, "c", (200, 200))
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-10


def test_2a_i(tmp_path):
    """benchmarks/02_oswald_ripening/2a.i (PFHub BM2a, BASELINE.json configs[2]; multi-line let-expression free energy)
    vs the oracle: 2 steps x 50 substeps."""
    run(tmp_path, "2a.i", "--allow-unused", "TensorSolver/substeps=50", "Executioner/num_steps=2", dump=("c", "n1", "n2", "n3", "n4"))
    p = oc.bm2_problem(substeps=50)
    p.initial()
    dt = 0.001
    for _ in range(2):
        p.step(dt)
        dt *= 1.1
    for k in ("c", "n1", "n2", "n3", "n4"):
        ref = p.buf[k].numpy()
        got = field(tmp_path, k, (200, 200))
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-10, k


def test_cahnhilliard2_i(tmp_path):
    """examples/cahn_hilliard/cahnhilliard2.i (the file the headline workload is derived from) at a reduced grid
    (the file's own dx kept) vs the oracle: 2 steps."""
    n = 32
    L = n * (8 * math.pi / 200)
    text = open(f"{REF}/cahnhilliard2.i").read()
    args = ["--allow-unused", f"Domain/nx={n}", f"Domain/ny={n}", f"Domain/nz={n}", f"Domain/xmax={L!r}", f"Domain/ymax={L!r}",
            f"Domain/zmax={L!r}", "Executioner/num_steps=2", "TensorComputes/Initialize/c/seed=0"]   # the file draws an unseeded IC
    r = run(tmp_path, "cahnhilliard2.i", *args, "Problem/print_debug_output=true", dump=("c",))
    assert "fused five-pass plan" in r.stderr + r.stdout
    import re
    ss = int(re.search(r"substeps\s*=\s*(\d+)", text).group(1))
    dt0 = float(re.search(r"\n\s*dt\s*=\s*([0-9.e+-]+)", text).group(1))
    gf = re.search(r"growth_factor\s*=\s*([0-9.e+-]+)", text)
    p = oc.ch_problem(3, n, L, substeps=ss)
    p.initial()
    dt = dt0
    for _ in range(2):
        p.step(dt)
        if gf:
            dt *= float(gf.group(1))
    ref = p.buf["c"].numpy()
    got = field(tmp_path, "c", (n, n, n))
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-10

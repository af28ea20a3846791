"""The stand-alone host driver (marlin_b200-opt: MOOSE-style TensorProblem / TensorOperator /
TensorSolver objects over the C ABI) run on input files, against the reference's own gold
results (tests/golden/*.npz) and the oracle.  GPU only."""
import math
import os
import subprocess

import numpy as np
import pytest
import torch

import oracle_cases as oc
from oracle import marlin as om

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "marlin_b200", "marlin_b200-opt")
INP = os.path.join(ROOT, "tests", "inputs")
G = os.path.join(ROOT, "tests", "golden")


def run(tmp, inp, *args, dump=()):
    cmd = [APP, "-i", f"{INP}/{inp}", "--output-dir", str(tmp), *args]
    if dump:
        cmd += ["--dump", ",".join(dump), "--dump-dir", str(tmp)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r


def field(tmp, name, shape):
    return np.fromfile(f"{tmp}/{name}.f64", dtype=np.float64).reshape(shape)


def csv(path):
    with open(path) as fh:
        head = fh.readline().strip().split(",")
        rows = np.array([[float(x) for x in ln.split(",")] for ln in fh if ln.strip()])
    return head, rows


@pytest.mark.parametrize("steps", [1, 3, 10])
def test_ch2d_input_matches_exodus_gold(tmp_path, steps):
    """test/tests/cahnhilliard/cahnhilliard.i -> gold/cahnhilliard_out.e; AB1 during the first
    MOOSE step (quirk Q1), AB2 afterwards."""
    g = np.load(f"{G}/ch2d_exodus.npz")
    r = run(tmp_path, "ch2d_gold.i", f"Executioner/num_steps={steps}", dump=("c", "mu"))
    assert "AuxKernels" in r.stderr                      # finite-element blocks reported, skipped
    c, mu = field(tmp_path, "c", (20, 20)), field(tmp_path, "mu", (20, 20))
    assert np.abs(c - g["c"][steps]).max() < 1e-12
    assert np.abs(mu - g["mu"][steps]).max() < 1e-12
    head, rows = csv(f"{tmp_path}/ch2d_gold_out.csv")
    assert head == ["time", "delta_int_c", "dt_crit", "int_c"] and rows.shape[0] == steps + 1
    assert abs(rows[-1, 0] - steps * 1e-3) < 1e-15
    dx = 3.0 / 20
    assert abs(rows[-1, 3] - g["c"][steps].sum() * dx * dx) < 1e-12
    # delta_int_c = TensorIntegralChangePostprocessor = integral of |c - c_old[0]|
    # (src/postprocessors/TensorIntegralChangePostprocessor.C:44-54): no history during MOOSE step 1
    # (quirk Q1) -> integral of |c|; afterwards c_old[0] is c before the LAST substep
    assert abs(rows[1, 1] - np.abs(g["c"][1]).sum() * dx * dx) < 1e-12
    if steps > 1:
        p = oc.ch_problem(2, 20, 3.0, substeps=10)
        old = p.get_old("c", 1)
        p.initial()
        for k in range(1, steps + 1):
            p.step(1e-3)
            if k > 1:
                ref = float((p.buf["c"] - old[0]).abs().sum()) * dx * dx
                assert abs(rows[k, 1] - ref) < 1e-12 * max(1.0, ref / 1e-3), (k, rows[k, 1], ref)


@pytest.mark.parametrize("ss,cs,order", [(10, 0, 1), (10, 0, 2), (10, 0, 3), (20, 0, 4),
                                         (10, 1, 1), (10, 2, 1), (10, 2, 2)])
def test_abm_diagonal_input_matches_csv_gold(tmp_path, ss, cs, order):
    """test/tests/solvers/diagonal.i with the cli_args of test/tests/solvers/tests."""
    gold = np.load(f"{G}/csv_golds.npz")[f"diagonal_{ss}_{cs}_{order}"]
    run(tmp_path, "abm_diagonal.i", f"ss={ss}", f"cs={cs}", f"order={order}")
    head, rows = csv(f"{tmp_path}/abm_diagonal_{ss}_{cs}_{order}.csv")
    assert head == ["time", "U", "V", "u_max", "u_min", "v_max", "v_min"]
    assert rows.shape == gold.shape
    err = np.abs(rows - gold) / np.maximum(np.abs(gold), 1e-8)
    assert err[1:].max() < 1e-9, err.max()


COUPLED = [(10, 0, 1), (10, 0, 2), (10, 0, 3), (20, 0, 4), (10, 1, 1), (10, 2, 1), (10, 2, 2)]


@pytest.mark.parametrize("ss,cs,order", COUPLED + [(1, 0, 1), (5, 0, 1)])
def test_abm_coupled_input_matches_csv_gold(tmp_path, ss, cs, order):
    """test/tests/solvers/coupled.i (AdamsBashforthMoultonCoupled, dense linear operator solved per
    wavevector by mrl_coupled_solve) with the cli_args of test/tests/solvers/tests."""
    gold = np.load(f"{G}/csv_golds.npz")[f"coupled_{ss}_{cs}_{order}"]
    run(tmp_path, "abm_coupled.i", f"ss={ss}", f"cs={cs}", f"order={order}")
    head, rows = csv(f"{tmp_path}/abm_coupled_{ss}_{cs}_{order}.csv")
    assert head == ["time", "U", "V", "u_max", "u_min", "v_max", "v_min"]
    assert rows.shape == gold.shape
    err = np.abs(rows - gold) / np.maximum(np.abs(gold), 1e-4)   # U, V are round-off sums (1e-17)
    assert err[1:].max() < 1e-9, err.max()


@pytest.mark.parametrize("ss,cs,order", COUPLED)
def test_abm_nl_coupled_input_matches_csv_gold(tmp_path, ss, cs, order):
    """test/tests/solvers/nl_coupled.i: complex-valued ParsedCompute nonlinear terms, correctors."""
    gold = np.load(f"{G}/csv_golds.npz")[f"nl_coupled_{ss}_{cs}_{order}"]
    run(tmp_path, "abm_nl_coupled.i", f"ss={ss}", f"cs={cs}", f"order={order}")
    head, rows = csv(f"{tmp_path}/abm_nl_coupled_{ss}_{cs}_{order}.csv")
    assert head == ["time", "U", "V", "u_max", "u_min", "v_max", "v_min"]
    assert rows.shape == gold.shape
    err = np.abs(rows - gold) / np.maximum(np.abs(gold), 1e-4)
    assert err[1:].max() < 1e-9, err.max()


def test_etdrk4_input_matches_csv_gold(tmp_path):
    """test/tests/solvers/etdrk4_diffusion.i -> gold/etdrk4_diffusion_rmse.csv (time,mse,rmse);
    rmse is a MOOSE ParsedPostprocessor = sqrt(mse) (skipped by the driver, formed here)."""
    gold = np.load(f"{G}/csv_golds.npz")["etdrk4_diffusion_rmse"]
    run(tmp_path, "etdrk4_decay.i")
    head, rows = csv(f"{tmp_path}/etdrk4_decay.csv")
    assert head == ["time", "mse"] and rows.shape[0] == gold.shape[0]
    assert np.abs(rows[:, 0] - gold[:, 0]).max() < 1e-12
    # (the root compute's last evaluation inside a substep sees stage d with t still at the start
    # of the substep, so the gold "error" is the one-step lag 1 - exp(-D k^2 dt), not round-off)
    assert np.abs(rows[:, 1] - gold[:, 1]).max() < 1e-12
    assert np.abs(np.sqrt(rows[:, 1]) - gold[:, 2]).max() < 1e-12


def test_etdrk4_nonzero_nonlinear_input_matches_oracle(tmp_path):
    """ETDRK4Solver with a non-zero nonlinear term (tests/inputs/etdrk4_cahnhilliard.i): all four stages and the
    phi coefficients, including the L*dt == 0 branch at k = 0 (dt, dt^2/2, dt^2/6 as coded in
    src/tensor_solver/ETDRK4Solver.C:84-91 - a dt^3/6 there moves this result by 2.4e-3), vs the oracle."""
    run(tmp_path, "etdrk4_cahnhilliard.i", dump=("c",))
    p = oc.etdrk4_ch_problem()
    p.initial()
    for _ in range(3):
        p.step(0.2)
    ref = p.buf["c"].numpy()
    got = field(tmp_path, "c", (32, 32))
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-10


def test_fft_semi_implicit_input_matches_oracle(tmp_path):
    """FFTSemiImplicit (src/tensor_timeintegrators/FFTSemiImplicit.C:43-62; no test in the reference) as an operator
    under a forwarding ForwardEulerSolver: first-order update during MOOSE step 1 (no history, quirk Q1), the
    3/2 N - 1/2 N_old combination afterwards; 3 steps x 5 substeps vs the oracle."""
    for steps in (1, 3):
        run(tmp_path, "fft_semi_implicit.i", f"Executioner/num_steps={steps}", dump=("c",))
        p = oc.fft_semi_implicit_problem()
        p.initial()
        for _ in range(steps):
            p.step(5e-3)
        ref = p.buf["c"].numpy()
        got = field(tmp_path, "c", (32, 24))
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-10, steps


def test_gradient_input(tmp_path):
    """test/tests/gradient/gradient.i (+ gradient_square.i): error integrals at round-off level
    like the gold CSVs (7.6e-12 / 1.5e-11)."""
    run(tmp_path, "fft_gradient.i")
    head, rows = csv(f"{tmp_path}/fft_gradient_out.csv")
    assert head == ["time", "diff", "gsq"]
    vol = 2 * math.pi * 4 * math.pi * 6 * math.pi
    assert 0 <= rows[-1, 1] < 1e-10
    assert abs(rows[-1, 2] - 0.5 * 1.5 * vol) < 1e-9 * vol


def test_mech3d_input_matches_hdf5_gold(tmp_path):
    """test/tests/mechanics/mech3d.i -> gold/mech3d.h5 (deformation gradient after 3 steps)."""
    g = np.load(f"{G}/mech3d_h5.npz")["F"]
    run(tmp_path, "mech3d_shear.i", dump=("F", "sV", "disp"))
    disp = np.moveaxis(np.load(f"{G}/mech3d_h5.npz")["disp"][2], -1, 0)   # [Postprocess] ComputeDisplacements, nodal 17^3
    assert np.abs(field(tmp_path, "disp", (3, 17, 17, 17)) - disp).max() < 1e-10
    F = field(tmp_path, "F", (9, 16, 16, 16))            # rank-two fields are component major
    ref = np.moveaxis(g[2].reshape(16, 16, 16, 9), -1, 0)
    rel = np.linalg.norm(F - ref) / np.linalg.norm(ref)
    assert rel < 1e-9, rel
    sV = np.load(f"{G}/mech3d_h5.npz")["sV"][2]             # [Postprocess] ComputeVonMisesStress
    assert np.abs(field(tmp_path, "sV", (16, 16, 16)) - sV).max() < 1e-9 * np.abs(sV).max()
    for k in (1, 2):
        run(tmp_path, "mech3d_shear.i", f"Executioner/num_steps={k}", dump=("F",))
        F = field(tmp_path, "F", (9, 16, 16, 16))
        ref = np.moveaxis(g[k - 1].reshape(16, 16, 16, 9), -1, 0)
        assert np.linalg.norm(F - ref) / np.linalg.norm(ref) < 1e-9


def test_mech2d_input_matches_hdf5_gold(tmp_path):
    """test/tests/mechanics/mech.i (2-D, 2x2 tensors) -> gold/mech.h5."""
    g = np.load(f"{G}/mech2d_h5.npz")["F"]
    for k in (1, 3):
        run(tmp_path, "mech2d_shear.i", f"Executioner/num_steps={k}", dump=("F", "sV", "disp"))
        disp = np.moveaxis(np.load(f"{G}/mech2d_h5.npz")["disp"][k - 1], -1, 0)
        assert np.abs(field(tmp_path, "disp", (2, 33, 33)) - disp).max() < 1e-10
        sV = np.load(f"{G}/mech2d_h5.npz")["sV"][k - 1]
        assert np.abs(field(tmp_path, "sV", (32, 32)) - sV).max() < 1e-9 * np.abs(sV).max()
        F = field(tmp_path, "F", (4, 32, 32))
        ref = np.moveaxis(g[k - 1].reshape(32, 32, 4), -1, 0)            # rank-two fields are component major
        assert np.linalg.norm(F - ref) / np.linalg.norm(ref) < 1e-9


def test_swift_hohenberg_secant_input_matches_hdf5_gold(tmp_path):
    """test/tests/tensor_compute/rotating_grain_secant.i -> gold/rotating_grain_secant.h5 (abs_tol 1e-10
    in the reference's HDF5Diff): [Functions]/MooseFunctionTensor IC, SwiftHohenbergLinear, SecantSolver,
    TensorSolveIterationAdaptiveDT."""
    g = np.load(f"{G}/rotating_grain_secant_h5.npz")["psi"]
    run(tmp_path, "swift_hohenberg_secant.i", "Executioner/num_steps=0", dump=("psi",))
    assert np.abs(field(tmp_path, "psi", (40, 40)) - g[0]).max() < 1e-13
    for k in (1, 4, 10):
        r = run(tmp_path, "swift_hohenberg_secant.i", f"Executioner/num_steps={k}", dump=("psi",))
        assert np.abs(field(tmp_path, "psi", (40, 40)) - g[k]).max() < 1e-10, k
    assert f"dt = {1.4 ** 9:.8g}"[:12] in r.stderr            # the step grew by growth_factor every step


def test_linear_tensor_predictor(tmp_path):
    """[TensorSolver/Predictors/*] with LinearTensorPredictor (src/tensor_predictor/LinearTensorPredictor.C:26-38).  As in
    the reference - whose AddTensorPredictorAction builds the object but never hands it to the solver
    (src/actions/AddTensorPredictorAction.C:41) - the block alone changes nothing: the run reproduces the gold of the input
    without it.  With apply_predictors = true the secant iteration starts from u + (u_old0 - u_old1) and must reach the
    same implicit-Euler solution to within the solver's tolerance."""
    g = np.load(f"{G}/rotating_grain_secant_h5.npz")["psi"]
    r0 = run(tmp_path, "secant_predictor.i", "Executioner/num_steps=4", dump=("psi",))
    assert np.abs(field(tmp_path, "psi", (40, 40)) - g[4]).max() < 1e-10
    r1 = run(tmp_path, "secant_predictor.i", "Executioner/num_steps=4", "TensorSolver/apply_predictors=true", dump=("psi",))
    psi = field(tmp_path, "psi", (40, 40))
    d = np.abs(psi - g[4]).max()
    assert 0 < d < 1e-6, d          # another starting point, the same fixed point
    # an object that is not a predictor, or a solver that cannot take one, is an error
    r = subprocess.run([APP, "-i", f"{INP}/secant_predictor.i", "--compute-device=cuda", "TensorSolver/type=AdamsBashforthMoulton"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "not an iterative tensor solver" in r.stderr


def test_kks_no_flux_input_matches_gold(tmp_path):
    """test/tests/kks/KKS_no_flux_bc.i -> gold KKS_no_flux_bc.h5 (abs_tol 1e-10) and KKS_no_flux_bc_out.csv:
    ReciprocalMatDiffusion, ReciprocalAllenCahn, mask from a ParsedFunction on a domain with a non-zero
    minimum, two coupled variables, AB3, 10 x 1000 substeps."""
    g = np.load(f"{G}/kks_no_flux_bc.npz")
    run(tmp_path, "kks_no_flux.i", "Executioner/num_steps=0", dump=("c", "eta", "psi"))
    for k in ("c", "eta", "psi"):
        assert np.abs(field(tmp_path, k, (20, 20)) - g[k][0]).max() < 1e-13, k
    run(tmp_path, "kks_no_flux.i", "Executioner/num_steps=1", dump=("c", "eta", "mu"))
    for k in ("c", "eta", "mu"):
        assert np.abs(field(tmp_path, k, (20, 20)) - g[k][1]).max() < 1e-10, k
    run(tmp_path, "kks_no_flux.i", dump=("c", "eta", "mu"))
    for k in ("c", "eta", "mu"):
        assert np.abs(field(tmp_path, k, (20, 20)) - g[k][10]).max() < 1e-9, k
    head, rows = csv(f"{tmp_path}/kks_no_flux_out.csv")
    assert head == ["time", "total_C", "total_eta"] and rows.shape == g["csv"].shape
    assert (np.abs(rows - g["csv"]) / np.maximum(np.abs(g["csv"]), 1.0)).max() < 1e-9


def test_ch2d_explicit_input_matches_exodus_gold(tmp_path):
    """test/tests/cahnhilliard/cahnhilliard_explicit.i (ForwardEulerSolver) -> gold cahnhilliard_explicit_out.e."""
    g = np.load(f"{G}/ch2d_explicit_exodus.npz")
    frames = list(g["frames"])
    for step in (1, 10, 20):
        run(tmp_path, "ch2d_explicit.i", f"Executioner/num_steps={step}", dump=("c", "mu"))
        k = frames.index(step)
        assert np.abs(field(tmp_path, "c", (50, 50)) - g["c"][k]).max() < 1e-10, step
        assert np.abs(field(tmp_path, "mu", (50, 50)) - g["mu"][k]).max() < 1e-10, step


@pytest.mark.parametrize("method", ["SHARP", "HOULI"])
def test_ch2d_explicit_smooth_input_matches_exodus_gold(tmp_path, method):
    """test/tests/cahnhilliard/cahnhilliard_explicit_smooth.i (DeAliasingTensor) -> gold sharp.e / houli.e."""
    g = np.load(f"{G}/ch2d_explicit_{method.lower()}_exodus.npz")
    frames = list(g["frames"])
    for step in (1, 5, 20):
        run(tmp_path, "ch2d_explicit_smooth.i", f"smooth={method}", f"Executioner/num_steps={step}", dump=("c", "mu"))
        k = frames.index(step)
        assert np.abs(field(tmp_path, "c", (50, 50)) - g["c"][k]).max() < 1e-9, step
        assert np.abs(field(tmp_path, "mu", (50, 50)) - g["mu"][k]).max() < 1e-9, step


def test_ch3d_map_to_aux_input_matches_exodus_gold(tmp_path):
    """test/tests/cahnhilliard/cahnhilliard.i with the 3-D cli_args of the reference's tests file (gold map_to_aux_3d.e)."""
    g = np.load(f"{G}/ch3d_map_to_aux_exodus.npz")
    for step in (1, 10):
        run(tmp_path, "ch2d_gold.i", "Domain/dim=3", "Domain/nx=5", "Domain/ny=5", "Domain/nz=5", "Domain/zmax=3",
            f"Executioner/num_steps={step}", dump=("c", "mu"))
        assert np.abs(field(tmp_path, "c", (5, 5, 5)) - g["c"][step]).max() < 1e-12, step
        assert np.abs(field(tmp_path, "mu", (5, 5, 5)) - g["mu"][step]).max() < 1e-12, step


def test_postprocessors_input_matches_csv_golds(tmp_path):
    """test/tests/postprocessors/postprocessors.i -> gold average.csv (0.8), integral.csv (4.8),
    extreme_value.csv (3.2375 / -1.6375), reciprocal_integral.csv (4.8), count.csv (0, 10, 20)."""
    run(tmp_path, "pp_basic.i")
    head, rows = csv(f"{tmp_path}/pp_basic.csv")
    assert head == ["time", "avg_c", "count", "int_c", "int_c_bar", "max_c", "min_c"]
    col = {h: rows[:, i] for i, h in enumerate(head)}
    assert np.abs(col["avg_c"] - 0.8).max() < 1e-13 and np.abs(col["int_c"] - 4.8).max() < 1e-12
    assert np.abs(col["int_c_bar"] - 4.8).max() < 1e-12
    assert np.abs(col["max_c"] - 3.2375).max() < 1e-13 and np.abs(col["min_c"] + 1.6375).max() < 1e-13
    assert list(col["count"]) == [0.0, 10.0, 20.0] and list(col["time"]) == [0.0, 1.0, 2.0]


def test_interface_velocity_input_matches_csv_gold(tmp_path):
    """test/tests/postprocessors/interface_velocity.i -> gold interface_velocity_out.csv (no solver: the
    [Solve] computes run every step; history of c through getBufferOld)."""
    gold = np.load(f"{G}/csv_golds.npz")["interface_velocity_out"]
    run(tmp_path, "interface_velocity.i")
    head, rows = csv(f"{tmp_path}/interface_velocity_out.csv")
    assert head == ["time", "v"] and rows.shape == gold.shape
    assert np.abs(rows - gold).max() < 1e-10, (rows, gold)


def test_broyden_input_matches_oracle(tmp_path):
    """BroydenSolver (src/tensor_solver/BroydenSolver.C; no gold file in the reference) through the host
    objects (mrl_broyden_step / mrl_broyden_update) vs the oracle's restatement, two steps of two substeps.
    The kernels themselves are checked at round-off level in test_gpu_parity.py; the trajectory is not
    round-off stable (rank-one updates divide by sk.yk down to the 1e-12 guard, and yk = Rnew - R cancels),
    so two correct implementations agree to ~1e-7 here, not to 1e-10."""
    run(tmp_path, "broyden_coupled.i", dump=("u", "v"))
    p = oc.broyden_problem()
    p.initial()
    for _ in range(2):
        p.step(0.05)
        assert p.solver.iterations == 12          # fixed iteration count (tolerances 0 in the input)
    for k in ("u", "v"):
        ref = p.buf[k].numpy()
        got = field(tmp_path, k, (32, 32))
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-6, k


def test_bm1_spinodal_input_matches_oracle(tmp_path):
    """benchmarks/01_spinodal_decomposition/1a_solver.i (PFHub BM1a, BASELINE.json configs[1]) through the host objects
    (fused plan) vs the oracle: 3 steps x 1000 substeps, dt = 1, 1.1, 1.21; rel L2 <= 1e-10; free energy and extrema
    from the CSV."""
    r = run(tmp_path, "bm1_spinodal.i", "Executioner/num_steps=3", "Problem/print_debug_output=true", dump=("c",))
    assert "fused five-pass plan" in r.stderr + r.stdout
    p = oc.bm1_problem()
    p.initial()
    dt = 1.0
    for _ in range(3):
        p.step(dt)
        dt *= 1.1
    ref = p.buf["c"].numpy()
    got = field(tmp_path, "c", (200, 200))
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-10
    fg = om.FFTGradientSquare(p, "Fgrad", "c", 1.0)
    fg.compute()
    F = om.ParsedCompute(p, "F", "rho_s * (c-c_alpha)^2 * (c_beta-c)^2 + Fgrad", inputs=["c", "Fgrad"],
                         constant_names=["rho_s", "c_alpha", "c_beta"], constant_expressions=["5", "0.3", "0.7"])
    F.compute()
    head, rows = csv(f"{tmp_path}/bm1_spinodal.csv")
    assert head == ["time", "F", "change", "max_c", "min_c"]
    assert abs(rows[-1, 1] - om.pp_integral(p, "F")) < 1e-9 * abs(rows[-1, 1])
    assert abs(rows[-1, 3] - float(p.buf["c"].max())) < 1e-11 and abs(rows[-1, 4] - float(p.buf["c"].min())) < 1e-11


def test_bm2_ostwald_input_matches_oracle(tmp_path):
    """benchmarks/02_oswald_ripening/2a.i (PFHub BM2a, BASELINE.json configs[2]: five coupled fields, parsed free
    energy with let-bindings and symbolic derivatives, IterationAdaptiveDT growth 1.1) through the host objects vs
    the oracle: 3 steps x 50 substeps (the first step at AB1 by quirk Q1, order reset on every dt change by Q2),
    rel L2 <= 1e-10 per field."""
    run(tmp_path, "bm2_ostwald.i", "TensorSolver/substeps=50", "Executioner/num_steps=3", dump=("c", "n1", "n2", "n3", "n4", "F"))
    p = oc.bm2_problem(substeps=50)
    p.initial()
    dt = 0.001
    for _ in range(3):
        p.step(dt)
        dt *= 1.1
    for k in ("c", "n1", "n2", "n3", "n4"):
        ref = p.buf[k].numpy()
        got = field(tmp_path, k, (200, 200))
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-10, k
    head, rows = csv(f"{tmp_path}/bm2_ostwald.csv")
    assert head == ["time", "F", "max_c", "min_c"]
    assert abs(rows[-1, 2] - float(p.buf["c"].max())) < 1e-12 and abs(rows[-1, 3] - float(p.buf["c"].min())) < 1e-12


def test_solver_fuses_canonical_inputs_and_matches_unfused(tmp_path):
    """The AdamsBashforthMoulton host object recognises the canonical split-operator compute graph and
    replaces it by the fused CUDA plan (reported with Problem/print_debug_output=true); `fuse=false` keeps
    the operator-by-operator path.  Both give the same fields."""
    args = ("Domain/dim=3", "Domain/nx=32", "Domain/ny=32", "Domain/nz=32", "Domain/zmax=3", "Executioner/num_steps=2",
            "Problem/print_debug_output=true")
    r = run(tmp_path, "ch2d_gold.i", *args, dump=("c",))
    assert "fused five-pass plan" in r.stderr + r.stdout, (r.stderr + r.stdout)[-2000:]
    fused = field(tmp_path, "c", (32, 32, 32)).copy()
    r = run(tmp_path, "ch2d_gold.i", *args, "TensorSolver/fuse=false", dump=("c",))
    assert "operator-by-operator path" in r.stderr + r.stdout
    plain = field(tmp_path, "c", (32, 32, 32))
    assert np.linalg.norm(fused - plain) / np.linalg.norm(plain) < 1e-12
    # PFHub BM2a: five coupled variables, compiled nonlinearities with symbolic derivatives
    r = run(tmp_path, "bm2_ostwald.i", "TensorSolver/substeps=5", "Executioner/num_steps=1", "Problem/print_debug_output=true")
    assert "fused five-pass plan" in r.stderr + r.stdout, (r.stderr + r.stdout)[-2000:]


def test_xdmf_output_matches_gold_xmf_and_fields(tmp_path):
    """test/tests/cahnhilliard/cahnhilliard.i with TensorOutputs/active="xdmf" -> gold cahnhilliard.xmf (XMLDiff in the
    reference).  The document must equal the gold one except for the DataItem storage (raw binary files here, HDF5
    datasets there); the data files hold c (NODE: periodic continuation to 21x21) and mu (CELL) of every frame."""
    import re
    gold = open(f"{G}/cahnhilliard_gold.xmf").read()
    run(tmp_path, "ch2d_gold.i", "TensorOutputs/active=xdmf")
    mine = open(f"{tmp_path}/ch2d_gold.xmf").read()
    norm_gold = re.sub(r' Format="HDF">cahnhilliard\.h5:/([a-z]+)\.(\d+)<', r' STORAGE>\1.\2<', gold)
    norm_mine = re.sub(r' Format="HDF">[^<]*ch2d_gold\.h5:/([a-z]+)\.(\d+)<', r' STORAGE>\1.\2<', mine)
    assert norm_mine == norm_gold
    import h5lite
    g = np.load(f"{G}/ch2d_exodus.npz")
    h = h5lite.H5File(f"{tmp_path}/ch2d_gold.h5")
    for frame in (0, 3, 10):
        c = h.read(f"c.{frame}")
        assert c.shape == (21, 21)
        assert np.abs(c[:20, :20] - g["c"][frame]).max() < 1e-12
        assert np.array_equal(c[20, :20], c[0, :20]) and np.array_equal(c[:, 20], c[:, 0])
        if frame:
            mu = h.read(f"mu.{frame}")
            assert np.abs(mu - g["mu"][frame]).max() < 1e-12
    # enable_hdf5 = false keeps the reference's raw binary storage
    run(tmp_path, "ch2d_gold.i", "TensorOutputs/active=xdmf", "TensorOutputs/xdmf/enable_hdf5=false")
    c = np.fromfile(f"{tmp_path}/ch2d_gold.c.3.bin", dtype="<f8").reshape(21, 21)
    assert np.array_equal(c, h5lite.H5File(f"{tmp_path}/ch2d_gold.h5").read("c.3")) or np.abs(c[:20, :20] - g["c"][3]).max() < 1e-12


def test_quasistatic_elasticity_input_matches_oracle(tmp_path):
    """FFTQuasistaticElasticity (src/tensor_computes/FFTQuasistaticElasticity.C:45-104: 3x3 solve per wavevector, here
    through mrl_coupled_solve on generated coefficient fields) and FFTElasticChemicalPotential
    (FFTElasticChemicalPotential.C:46-61) through the host objects vs the oracle (no gold file in the reference)."""
    run(tmp_path, "pf_mech_quasistatic.i", dump=("disp_x", "disp_y", "disp_z", "mumech"))
    shape, L = (16, 12, 10), 4 * math.pi
    d = om.Domain(3, list(shape), (0, 0, 0), (L, L, L))
    p = om.Problem(d)
    p.ics = [om.RandomTensor(p, "c", 0.44, 0.56, 0)]
    p.computes = [om.ForwardFFT(p, "cbar", "c"),
                  om.FFTQuasistaticElasticity(p, ["disp_x", "disp_y", "disp_z"], "cbar", 50.0, 100.0, 0.02),
                  om.FFTElasticChemicalPotential(p, "mumechbar", ["disp_x", "disp_y", "disp_z"], "cbar", 50.0, 100.0, 0.02),
                  om.InverseFFT(p, "mumech", "mumechbar")]
    p.initial()
    p.step(1.0)
    for k in ("disp_x", "disp_y", "disp_z", "mumech"):
        ref = p.buf[k].numpy()
        got = field(tmp_path, k, shape)
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-11, k


def test_histogram_input_matches_csv_gold(tmp_path):
    """test/tests/histogram/test.i ([VectorPostprocessors] TensorHistogram) -> gold test_out_hist_0001.csv."""
    gold = np.load(f"{G}/csv_golds.npz")["histogram_out_hist_0001"]
    run(tmp_path, "histogram.i")
    head, rows = csv(f"{tmp_path}/histogram_out_hist_0001.csv")
    assert head == ["bin", "count"] and rows.shape == gold.shape
    assert np.abs(rows[:, 0] - gold[:, 0]).max() < 1e-13 and np.array_equal(rows[:, 1], gold[:, 1])


def test_smooth_rectangle_input_matches_hdf5_gold(tmp_path):
    """test/tests/tensor_compute/smooth_rectangle.i (SmoothRectangleCompute sharp / COS / TANH as generated kernels)
    -> gold smooth_rectangle.h5."""
    g = np.load(f"{G}/smooth_rectangle_h5.npz")
    names = ("rectangle_sharp", "rectangle_cos", "rectangle_tanh")
    run(tmp_path, "smooth_rectangle.i", dump=names)
    for name in names:
        got = field(tmp_path, name, (100, 100))
        assert np.abs(got - g[name.split("_")[1]]).max() < 1e-13, name
    assert np.array_equal(field(tmp_path, "rectangle_sharp", (100, 100)), g["sharp"])


def test_ch3d_input_matches_oracle(tmp_path):
    """examples/cahn_hilliard/cahnhilliard2.i-style 3-D run (32^3, 2 steps x 10 substeps) through
    the host objects vs the oracle, rel L2 <= 1e-10 (BASELINE.json north_star)."""
    n, L = 32, 32 * (8 * math.pi / 200)
    run(tmp_path, "ch2d_gold.i", "Domain/dim=3", f"Domain/nx={n}", f"Domain/ny={n}", f"Domain/nz={n}",
        f"Domain/xmax={L!r}", f"Domain/ymax={L!r}", f"Domain/zmax={L!r}", "Executioner/num_steps=2",
        "Executioner/dt=0.01", dump=("c",))
    c = field(tmp_path, "c", (n, n, n))
    p = oc.ch_problem(3, n, L, substeps=10)
    p.initial()
    for _ in range(2):
        p.step(0.01)
    ref = p.buf["c"].numpy()
    assert np.linalg.norm(c - ref) / np.linalg.norm(ref) < 1e-10


def test_unknown_parameter_and_type_are_errors(tmp_path):
    r = subprocess.run([APP, "-i", f"{INP}/ch2d_gold.i", "TensorSolver/bogus=1"], capture_output=True, text=True)
    assert r.returncode != 0 and "bogus" in r.stderr
    r = subprocess.run([APP, "-i", f"{INP}/ch2d_gold.i", "TensorSolver/type=NoSuchSolver"], capture_output=True, text=True)
    assert r.returncode != 0 and "NoSuchSolver" in r.stderr

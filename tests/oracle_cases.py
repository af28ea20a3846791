"""Problem builders shared by the oracle pinning tests and the CUDA parity tests.

Each builder restates one reference input file with the oracle's operator classes
(oracle/marlin.py).  Paths cited are relative to the reference root.
"""
import math

from oracle import marlin as om


def ch_problem(dim, n, L, substeps, mu_expr="0.1*c^2*(c-1)^2", M=0.2, kappa=-0.001, seed=0,
               predictor_order=2, cmin=0.44, cmax=0.56, constant_names=(),
               constant_expressions=()):
    """test/tests/cahnhilliard/cahnhilliard.i and examples/cahn_hilliard/cahnhilliard2.i."""
    # n, L: one value for all axes or one per axis
    ns = list(n) if isinstance(n, (tuple, list)) else [n] * dim
    Ls = list(L) if isinstance(L, (tuple, list)) else [L] * dim
    d = om.Domain(dim, ns, (0, 0, 0), tuple(Ls) + (1.0,) * (3 - dim))
    p = om.Problem(d)
    p.ics = [om.RandomTensor(p, "c", cmin, cmax, seed),
             om.ReciprocalLaplacianFactor(p, "Mbar", M),
             om.ReciprocalLaplacianSquareFactor(p, "kappabarbar", kappa)]
    root = om.Group(p, [
        om.ParsedCompute(p, "mu", mu_expr, inputs=["c"], derivatives=["c"],
                         constant_names=constant_names, constant_expressions=constant_expressions),
        om.ForwardFFT(p, "mubar", "mu"),
        om.ParsedCompute(p, "Mbarmubar", "Mbar*mubar", inputs=["Mbar", "mubar"]),
        om.ForwardFFT(p, "cbar", "c"),
    ])
    p.solver = om.AdamsBashforthMoulton(p, root, ["c"], ["cbar"], ["kappabarbar"], ["Mbarmubar"],
                                        substeps=substeps, predictor_order=predictor_order)
    return p


def diagonal_problem(ss, cs, order, n=150):
    """test/tests/solvers/diagonal.i with cli_args ss/cs/order."""
    d = om.Domain(2, [n, n], (0, 0, 0), (2 * math.pi, 2 * math.pi, 1.0))
    p = om.Problem(d)
    cn, ce = ["A", "B"], ["1", "3.5"]
    p.ics = [om.ParsedCompute(p, "u", "sin(x)*sin(y)", extra_symbols=True, expand="REAL",
                              constant_names=cn, constant_expressions=ce),
             om.ConstantTensor(p, "v", 0.0),
             om.ReciprocalLaplacianFactor(p, "Du", 1e-2),
             om.ReciprocalLaplacianFactor(p, "Dv", 1e-3)]
    root = om.Group(p, [
        om.ForwardFFT(p, "u_bar", "u"),
        om.ForwardFFT(p, "v_bar", "v"),
        om.ParsedCompute(p, "source_u", "A - (B+1)*u +u^2*v", inputs=["u", "v"],
                         constant_names=cn, constant_expressions=ce),
        om.ForwardFFT(p, "source_u_bar", "source_u"),
        om.ParsedCompute(p, "source_v", "B*u - u^2*v", inputs=["u", "v"], constant_names=cn,
                         constant_expressions=ce),
        om.ForwardFFT(p, "source_v_bar", "source_v"),
    ])
    p.solver = om.AdamsBashforthMoulton(p, root, ["u", "v"], ["u_bar", "v_bar"], ["Du", "Dv"],
                                        ["source_u_bar", "source_v_bar"], substeps=ss,
                                        predictor_order=order, corrector_order=order,
                                        corrector_steps=cs)
    return p


def coupled_problem(ss, cs, order, n=150, nonlinear=False):
    """test/tests/solvers/coupled.i (AdamsBashforthMoultonCoupled, off-diagonal linear operator) and
    nl_coupled.i (the same coupling moved into the nonlinear term of AdamsBashforthMoulton)."""
    d = om.Domain(2, [n, n], (0, 0, 0), (2 * math.pi, 2 * math.pi, 1.0))
    p = om.Problem(d)
    p.ics = [om.ParsedCompute(p, "u", "sin(x)*sin(y)", extra_symbols=True, expand="REAL"),
             om.ParsedCompute(p, "v", "cos(x)*cos(y)", extra_symbols=True, expand="REAL"),
             om.ConstantTensor(p, "zero", 0.0, reciprocal=True),
             om.ReciprocalLaplacianFactor(p, "D1", 1e-2),
             om.ReciprocalLaplacianFactor(p, "D2", 1e-3)]
    ops = [om.ForwardFFT(p, "u_bar", "u"), om.ForwardFFT(p, "v_bar", "v")]
    if nonlinear:
        ops += [om.ParsedCompute(p, "Du", "D2*v_bar", inputs=["D2", "v_bar"]),
                om.ParsedCompute(p, "Dv", "D2*u_bar", inputs=["D2", "u_bar"])]
        root = om.Group(p, ops)
        p.solver = om.AdamsBashforthMoulton(p, root, ["u", "v"], ["u_bar", "v_bar"], ["D1", "D1"],
                                            ["Du", "Dv"], substeps=ss, predictor_order=order,
                                            corrector_order=order, corrector_steps=cs)
    else:
        root = om.Group(p, ops)
        p.solver = om.AdamsBashforthMoultonCoupled(
            p, root, ["u", "v"], ["u_bar", "v_bar"], ["D1", "D1"], ["zero", "zero"], substeps=ss,
            predictor_order=order, corrector_order=order, corrector_steps=cs,
            linear_offdiag_rows=[1, 0], linear_offdiag_cols=[0, 1], linear_offdiag=["D2", "D2"])
    return p


def diagonal_row(p):
    """Column order of the gold CSV: time,U,V,u_max,u_min,v_max,v_min."""
    return [p.time, om.pp_integral(p, "u"), om.pp_integral(p, "v"), om.pp_extreme(p, "u", "MAX"),
            om.pp_extreme(p, "u", "MIN"), om.pp_extreme(p, "v", "MAX"),
            om.pp_extreme(p, "v", "MIN")]


def mech2d_problem(n=32):
    """test/tests/mechanics/mech.i (2-D, 2x2 tensors; zmax is set but unused)."""
    return mech3d_problem(n, substeps=3, l_tol=1e-5, nl_rel_tol=2e-4, nl_abs_tol=2e-3, dim=2, l_max_its=40)


def mech3d_problem(n=16, substeps=10, l_tol=1e-2, nl_rel_tol=2e-2, nl_abs_tol=2e-2, dim=3, l_max_its=None, ics_only=False):
    """test/tests/mechanics/mech3d.i.  ics_only: just the initial fields (phase, K, mu, F) - the FFTMechanics object
    materialises Ghat4, 11 GB at 256^3."""
    L = 2 * math.pi
    d = om.Domain(dim, [n] * dim, (0, 0, 0), (L, L, L))
    p = om.Problem(d)
    p.ics = [
        om.ParsedCompute(p, "phase", "(cos(x)/2+0.5)^1*(cos(y)/2+0.5)^1*(cos(z)/2+0.5)^1",
                         extra_symbols=True),
        om.ParsedCompute(p, "K", "(1-phase)*Ka + phase*Kb", inputs=["phase"],
                         constant_names=["Ka", "Kb"], constant_expressions=["1", "10"]),
        om.ParsedCompute(p, "mu", "(1-phase)*mua + phase*mub", inputs=["phase"],
                         constant_names=["mua", "mub"], constant_expressions=["0.5", "5"]),
        om.RankTwoIdentity(p, "F"),
    ]
    if ics_only:
        return p
    hyper = om.HyperElasticIsotropic(p, "stress", "Fnew", "K", "mu")
    mech = om.FFTMechanics(p, "Fnew", hyper, "K", "mu", F="F", stress="stress",
                           applied_macroscopic_strain="applied_strain", l_tol=l_tol,
                           l_max_its=l_max_its, nl_rel_tol=nl_rel_tol, nl_abs_tol=nl_abs_tol)
    root = om.Group(p, [om.MacroscopicShearTensor(p, "applied_strain", "F"), mech])
    p.solver = om.ForwardEulerSolver(p, root, substeps=substeps, forward=[("F", "Fnew")])
    p.mech = mech
    return p


CRYSTAL = ("-(sin(sin(a)*y/2+cos(a)*x/2)^2 + sin(sin(a+1/3*pi)*y/2+cos(a+1/3*pi)*x/2)^2 + "
           "sin(sin(a-1/3*pi)*y/2+cos(a-1/3*pi)*x/2)^2 - 1.5)*0.25")


def rotating_grain_problem(n=40, w=6, substeps=3):
    """test/tests/tensor_compute/rotating_grain_secant.i: Swift-Hohenberg (phase-field crystal) rotated
    grain in a matrix, SecantSolver, TensorSolveIterationAdaptiveDT (see rotating_grain_run)."""
    d = om.Domain(2, [n, n], (0, 0, 0), (w * math.pi * 2, w * math.pi * 2 / math.sin(math.pi / 3), 1.0))
    p = om.Problem(d)
    functions = {
        "grain1": ("a := 0; " + CRYSTAL, [], []),
        "grain2": ("a := 0.95; " + CRYSTAL, [], []),
        "domain": (f"r := (x-{w}*pi)^2+(y-{w}*pi)^2; if(r<({w}*2/3*pi)^2, grain2, grain1)",
                   ["grain1", "grain2"], ["grain1", "grain2"]),
    }
    p.ics = [om.MooseFunctionTensor(p, "psi", "domain", functions),
             om.SwiftHohenbergLinear(p, "linear", r=0.025, alpha=1.0)]
    root = om.Group(p, [om.ParsedCompute(p, "psi3", "0.20*psi^2-psi^3", inputs=["psi"]),
                        om.ForwardFFT(p, "psibar", "psi"),
                        om.ForwardFFT(p, "psi3bar", "psi3")])
    p.solver = om.SecantSolver(p, root, ["psi"], ["psibar"], ["linear"], ["psi3bar"], substeps=substeps)
    return p


def rotating_grain_run(p, num_steps=10, dt=1.0, min_iterations=100, max_iterations=400, growth_factor=1.4,
                       cutback_factor=0.9, dtmax=500.0, on_step=None):
    """Transient + TensorSolveIterationAdaptiveDT (src/timesteppers/TensorSolveIterationAdaptiveDT.C:
    computeInitialDT :71-75, computeDT/computeAdaptiveDT :77-93,161-174): the next dt grows when the
    solver's last substep took fewer than min_iterations secant iterations, shrinks above
    max_iterations.  (A failed solve would repeat the step with half the dt; the gold run never fails.)"""
    p.initial()
    for step in range(num_steps):
        if step > 0:
            it = p.solver.iterations
            if it < min_iterations:
                dt *= growth_factor
            elif it > max_iterations:
                dt *= cutback_factor
        dt = min(dt, dtmax)
        p.step(dt)
        assert p.solver.converged, "secant solve failed: step repetition is not restated"
        if on_step:
            on_step(step, dt)


def kks_no_flux_problem(n=20, substeps=1000):
    """test/tests/kks/KKS_no_flux_bc.i: two-phase KKS-type model (c conserved through ReciprocalMatDiffusion,
    eta through ReciprocalAllenCahn) inside a smoothed-boundary mask psi, AdamsBashforthMoulton order 3."""
    r, l = 30, 4.2
    kappa_eta, rho_sq, w, M, L, c0_a, c0_b = 5, 2, 1, 5, 5, 0.3, 0.7
    eta_ic = f"0.5*(1-tanh(2*(sqrt(x^2+y^2)-{r})/{l}))"
    h = "eta^3*(6*eta^2-15*eta+10)"
    F = (f"{h}*({rho_sq}*((c - (1-{h})*({c0_b} - {c0_a}))-{c0_a})^2) + (1-{h})*({rho_sq}*((c + ({h})*({c0_b} - {c0_a}))-{c0_b})^2 ) "
         f"+ {w}*(eta^2)*(1-eta)^2")
    d = om.Domain(2, [n, n], (-50, -50, 0), (50, 50, 1.0))
    p = om.Problem(d)
    psi_expr = (f"if(x<x_min-{l},0,if(x>x_min+{l},1,0.5-0.5*cos(pi*(x-(x_min-{l}))/2/{l}) )) * "
                f"if(x<x_max-{l},1,if(x>x_max+{l},0,0.5+0.5*cos(pi*(x-(x_max-{l}))/2/{l}) ))")
    functions = {"psi_func": (psi_expr, ["x_min", "x_max", "y_min", "y_max"], ["30", "70", "0", "100"])}
    p.ics = [om.ParsedCompute(p, "c", f"0.6 + ({c0_a}-0.6)*{eta_ic}", extra_symbols=True),
             om.ParsedCompute(p, "eta", eta_ic, extra_symbols=True),
             om.MooseFunctionTensor(p, "psi", "psi_func", functions),
             om.ConstantTensor(p, "zero", 0.0, reciprocal=True),
             om.ConstantTensor(p, "M", float(M)), om.ConstantTensor(p, "L", float(L)),
             om.ConstantTensor(p, "L_kappa", float(L * kappa_eta))]
    root = om.Group(p, [
        om.ForwardFFT(p, "cbar", "c"), om.ForwardFFT(p, "etabar", "eta"),
        om.ParsedCompute(p, "mu", F, inputs=["c", "eta"], derivatives=["c"]),
        om.ReciprocalMatDiffusion(p, "div_J", "mu", "M", "psi"),
        om.ParsedCompute(p, "domega_chem_deta", f"{F} - mu*c", inputs=["mu", "c", "eta"], derivatives=["eta"]),
        om.ReciprocalAllenCahn(p, "AC_bulk", "domega_chem_deta", "L", "psi"),
        om.ReciprocalMatDiffusion(p, "kappa_grad_eta", "eta", "L_kappa", "psi"),
        om.ParsedCompute(p, "AC_bar", "kappa_grad_eta + AC_bulk", inputs=["AC_bulk", "kappa_grad_eta"]),
    ])
    p.solver = om.AdamsBashforthMoulton(p, root, ["c", "eta"], ["cbar", "etabar"], ["zero", "zero"], ["div_J", "AC_bar"],
                                        substeps=substeps, predictor_order=3)
    return p


def ch_explicit_problem(n=50, L=3.0, substeps=50, smooth=None):
    """test/tests/cahnhilliard/cahnhilliard_explicit.i: explicit (ForwardEulerSolver) Cahn-Hilliard, the
    time derivative Mbar*mubar - Mkappabarbar*cbar assembled by one ParsedCompute in reciprocal space."""
    d = om.Domain(2, [n, n], (0, 0, 0), (L, L, 1.0))
    p = om.Problem(d)
    p.ics = [om.RandomTensor(p, "c", 0.44, 0.56, 0), om.ConstantTensor(p, "mu", 0.0),
             om.ReciprocalLaplacianFactor(p, "Mbar", 0.2),
             om.ReciprocalLaplacianSquareFactor(p, "Mkappabarbar", 0.2 * 1e-4),
             om.ConstantTensor(p, "dc_dt_bar", 0.0, reciprocal=True)]
    rate, rate_in = "Mbar*mubar - Mkappabarbar*cbar", ["Mbar", "mubar", "Mkappabarbar", "cbar"]
    if smooth:  # cahnhilliard_explicit_smooth.i: de-aliasing filter on the time derivative
        p.ics.append(om.DeAliasingTensor(p, "smooth", smooth))
        rate, rate_in = f"smooth * ({rate})", rate_in + ["smooth"]
    root = om.Group(p, [
        om.ParsedCompute(p, "mu", "0.1*c^2*(c-1)^2", inputs=["c"], derivatives=["c"]),
        om.ForwardFFT(p, "mubar", "mu"),
        om.ForwardFFT(p, "cbar", "c"),   # dependency order (ComputeGroup sorts; the input lists it last)
        om.ParsedCompute(p, "dc_dt_bar", rate, inputs=rate_in),
    ])
    p.solver = om.ForwardEulerSolver(p, root, ["c"], ["cbar"], ["dc_dt_bar"], substeps=substeps)
    return p


def broyden_problem(n=32):
    """tests/inputs/broyden_coupled.i: two coupled Allen-Cahn-type fields, BroydenSolver."""
    L = 2 * math.pi
    d = om.Domain(2, [n, n], (0, 0, 0), (L, L, 1.0))
    p = om.Problem(d)
    p.ics = [om.ParsedCompute(p, "u", "0.5+0.1*sin(x)*sin(y)", extra_symbols=True, expand="REAL"),
             om.ParsedCompute(p, "v", "0.4+0.1*cos(x)*cos(2*y)", extra_symbols=True, expand="REAL"),
             om.ReciprocalLaplacianFactor(p, "Lu", 0.1), om.ReciprocalLaplacianFactor(p, "Lv", 0.05)]
    root = om.Group(p, [om.ForwardFFT(p, "ub", "u"), om.ForwardFFT(p, "vb", "v"),
                        om.ParsedCompute(p, "fu", "-(u^3-u) - 0.3*v", inputs=["u", "v"]), om.ForwardFFT(p, "fub", "fu"),
                        om.ParsedCompute(p, "fv", "-(v^3-v) - 0.3*u", inputs=["u", "v"]), om.ForwardFFT(p, "fvb", "fv")])
    p.solver = om.BroydenSolver(p, root, ["u", "v"], ["ub", "vb"], ["Lu", "Lv"], ["fub", "fvb"], substeps=2, max_iterations=12,
                                relative_tolerance=0.0, absolute_tolerance=0.0)
    return p


BM2_FCHEM = ("fa:=rho^2*(c-ca)^2; fb:=rho^2*(cb-c)^2; h:=n1^3*(6*n1^2-15*n1+10) + n2^3*(6*n2^2-15*n2+10) + "
             "n3^3*(6*n3^2-15*n3+10) + n4^3*(6*n4^2-15*n4+10); g:=n1^2*(1-n1)^2 + n2^2*(1-n2)^2 + n3^2*(1-n3)^2 + "
             "n4^2*(1-n4)^2 + alpha*(n1^2*n2^2 + n1^2*n3^2 + n1^2*n4^2 + n2^2*n1^2 + n2^2*n3^2 + n2^2*n4^2 + n3^2*n1^2 + "
             "n3^2*n2^2 + n3^2*n4^2 + n4^2*n1^2 + n4^2*n2^2 + n4^2*n3^2); (fa*(1-h) + fb*h + w*g)")
BM2_NIC = ("epsilon*(cos((0.01*idx)*x-4)*cos((0.007+0.01*idx)*y) +cos((0.11+0.01*idx)*x)*cos((0.11+0.01*idx)*y) "
           "+psi*(cos((0.046+0.001*idx)*x+(0.0405+0.001*idx)*y) *cos((0.031+0.001*idx)*x-(0.004+0.001*idx)*y))^2)^2")


def bm2_problem(n=200, substeps=2000):
    """benchmarks/02_oswald_ripening/2a.i (PFHub benchmark 2a, BASELINE.json configs[2])."""
    d = om.Domain(2, [n, n], (0, 0, 0), (200.0, 200.0, 1.0))
    p = om.Problem(d)
    cn, cv = ["rho", "ca", "cb", "alpha", "w", "L", "M"], ["sqrt(2)", "0.3", "0.7", "5", "1", "5", "5"]
    ns = ["n1", "n2", "n3", "n4"]
    p.ics = [om.ParsedCompute(p, "c", "c0+epsilon*(cos(0.105*x)*cos(0.11*y)+(cos(0.13*x)*cos(0.087*y))^2+"
                                      "cos(0.025*x-0.15*y)*cos(0.07*x-0.02*y))", extra_symbols=True,
                              constant_names=["c0", "epsilon"], constant_expressions=["0.5", "0.01"]),
             om.ReciprocalLaplacianFactor(p, "Lbar", 1.0),
             om.ReciprocalLaplacianSquareFactor(p, "MkappaL2bar", -15.0),
             om.ReciprocalLaplacianFactor(p, "kappaLbar", 15.0)]
    for k, nm in enumerate(ns):
        p.ics.append(om.ParsedCompute(p, nm, BM2_NIC, extra_symbols=True, constant_names=["idx", "epsilon", "psi"],
                                      constant_expressions=[str(k + 1), "0.1", "1.5"]))
    allv = ["c"] + ns
    ops = [om.ParsedCompute(p, "mu_c", f"{BM2_FCHEM}*M", inputs=allv, derivatives=["c"], constant_names=cn, constant_expressions=cv)]
    ops += [om.ParsedCompute(p, f"mu_{nm}", f"{BM2_FCHEM}*(-L)", inputs=allv, derivatives=[nm], constant_names=cn,
                             constant_expressions=cv) for nm in ns]
    ops += [om.ForwardFFT(p, f"mu_{nm}_bar", f"mu_{nm}") for nm in allv]
    ops += [om.ParsedCompute(p, "Mbar_mu_c_bar", "Lbar*mu_c_bar", inputs=["Lbar", "mu_c_bar"])]
    ops += [om.ForwardFFT(p, f"{nm}_bar", nm) for nm in allv]
    root = om.Group(p, ops)
    p.solver = om.AdamsBashforthMoulton(p, root, allv, [f"{nm}_bar" for nm in allv],
                                        ["MkappaL2bar"] + ["kappaLbar"] * 4,
                                        ["Mbar_mu_c_bar"] + [f"mu_{nm}_bar" for nm in ns], substeps=substeps,
                                        predictor_order=2, corrector_order=2, corrector_steps=0)
    return p


def bm1_problem(n=200, substeps=1000):
    """benchmarks/01_spinodal_decomposition/1a_solver.i (PFHub benchmark 1a, BASELINE.json configs[1])."""
    d = om.Domain(2, [n, n], (0, 0, 0), (200.0, 200.0, 1.0))
    p = om.Problem(d)
    cn, cv = ["rho_s", "c_alpha", "c_beta"], ["5", "0.3", "0.7"]
    p.ics = [om.ParsedCompute(p, "c", "c0+epsilon*(cos(0.105*x)*cos(0.11*y)+(cos(0.13*x)*cos(0.087*y))^2+"
                                      "cos(0.025*x-0.15*y)*cos(0.07*x-0.02*y))", extra_symbols=True,
                              constant_names=["c0", "epsilon"], constant_expressions=["0.5", "0.01"]),
             om.ReciprocalLaplacianFactor(p, "Mbar", 5.0), om.ReciprocalLaplacianSquareFactor(p, "kappabarbar", -10.0)]
    root = om.Group(p, [om.ParsedCompute(p, "mu", "rho_s*(c-c_alpha)^2*(c_beta-c)^2", inputs=["c"], derivatives=["c"],
                                         constant_names=cn, constant_expressions=cv),
                        om.ForwardFFT(p, "mubar", "mu"),
                        om.ParsedCompute(p, "Mbarmubar", "Mbar*mubar", inputs=["Mbar", "mubar"]),
                        om.ForwardFFT(p, "cbar", "c")])
    p.solver = om.AdamsBashforthMoulton(p, root, ["c"], ["cbar"], ["kappabarbar"], ["Mbarmubar"], substeps=substeps)
    return p


def fft_semi_implicit_problem():
    """tests/inputs/fft_semi_implicit.i: FFTSemiImplicit as an operator under a forwarding ForwardEulerSolver."""
    d = om.Domain(2, [32, 24], (0, 0, 0), (4.0, 3.0, 1.0))
    p = om.Problem(d)
    p.ics = [om.RandomTensor(p, "c", 0.44, 0.56, 0),
             om.ReciprocalLaplacianFactor(p, "Mbar", 0.2),
             om.ReciprocalLaplacianSquareFactor(p, "kappabarbar", -0.001)]
    root = om.Group(p, [
        om.ParsedCompute(p, "mu", "0.1*c^2*(c-1)^2", inputs=["c"], derivatives=["c"]),
        om.ForwardFFT(p, "mubar", "mu"),
        om.ParsedCompute(p, "Mbarmubar", "Mbar*mubar", inputs=["Mbar", "mubar"]),
        om.ForwardFFT(p, "cbar", "c"),
        om.FFTSemiImplicit(p, "cnew", "cbar", "kappabarbar", "Mbarmubar"),
    ])
    p.solver = om.ForwardEulerSolver(p, root, substeps=5, forward=[("c", "cnew")])
    return p


def etdrk4_ch_problem():
    """tests/inputs/etdrk4_cahnhilliard.i: ETDRK4Solver with a non-zero nonlinear term (k = 0 mode included)."""
    d = om.Domain(2, [32, 32], (0, 0, 0), (4.0, 4.0, 1.0))
    p = om.Problem(d)
    p.ics = [om.RandomTensor(p, "c", 0.44, 0.56, 0),
             om.ReciprocalLaplacianFactor(p, "Mbar", 0.2),
             om.ReciprocalLaplacianSquareFactor(p, "kappabarbar", -0.5)]
    root = om.Group(p, [
        om.ParsedCompute(p, "mu", "0.1*c^2*(c-1)^2", inputs=["c"], derivatives=["c"]),
        om.ForwardFFT(p, "mubar", "mu"),
        om.ForwardFFT(p, "cbar", "c"),
        om.ParsedCompute(p, "Nbar", "Mbar*mubar + 0.5*cbar", inputs=["Mbar", "mubar", "cbar"]),
    ])
    p.solver = om.ETDRK4Solver(p, root, ["c"], ["cbar"], ["kappabarbar"], ["Nbar"], substeps=4)
    return p

"""Device expression compiler (ParsedCompute on the GPU): generic pointwise kernels and the
expression-specialised first FFT pass, against the oracle (libTorch CPU restatement)."""
import math

import pytest
import torch

import oracle_cases as oc
from ch_driver import SplitDriver
from marlin_b200 import capi
from marlin_b200.capi import AB_BETA
from oracle import exprparser as xp
from oracle import marlin as om

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float(torch.linalg.norm((a.double() - b.double()).reshape(-1)) / torch.linalg.norm(b.double().reshape(-1)))


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0, capi.F64)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx32():
    c = capi.Context(0, capi.F32)
    yield c
    c.close()


CASES = ["hypot(x,y)", "sqrt(x^2+y^2+n)", "tan((x-y)/2)", "tanh(x-y)", "atan(x + y)", "asin((x * y / 2) / n)",
         "acosh(x+y+1)", "atan2(x,y)", "1/sqrt(x+y)", "-x", "rsqrt(x*y)", "exp2(x*y)", "(x*y) % 1.5", "pow(y, x)",
         "min(x^3,y^2)", "pow(2, x)", "if(x<1 | y>=2, x, y)", "if(x<=1 & y>2, x*x, 3*y)", "r2:=x^2+y^2; sqrt(r2)",
         "max(x^2,sin(4*y))", "x^y", "x^(-2.5) + y^0.5 + x^7 + y^(-3)", "!(x>1) + (y<=2)", "round(3*x)+ceil(y)+floor(x)+trunc(-y)",
         "a:=sin(x^2); a + 2*a + 3*a", "abs(x-1)*log10(x)+log2(y)+cosh(x)*sinh(y)+log(x)+exp(-y)"]


@pytest.mark.parametrize("deriv", [(), ("x",), ("y",), ("x", "y")])
def test_expression_kernels_match_oracle(ctx, deriv):
    """unit/src/ParsedTensorTest.C:138-203 expressions (+ derivatives) evaluated on full fields."""
    ctx.domain_set(2, (24, 18), (0, 0), (1.0, 1.0))
    g = torch.Generator().manual_seed(1)
    x = torch.rand((24, 18), dtype=torch.float64, generator=g) * 1.9 + 0.1
    y = torch.rand((24, 18), dtype=torch.float64, generator=g) * 2.9 + 0.11
    n = torch.tensor([(x * y).max().item() * 1.01], dtype=torch.float64)
    dx, dy, dn = x.cuda(), y.cuda(), n.cuda()
    for expr in CASES:
        if deriv and ("%" in expr or expr.startswith("if(") or "round" in expr or "!" in expr):
            continue
        f = xp.ParsedTensor(expr, ["x", "y", "n"])
        for d in deriv:
            f.differentiate(d)
        f.compile()
        ref = f.eval([x, y, n[0]])
        if ref.dtype == torch.bool:
            ref = ref.double()
        ref = ref.expand(24, 18) if ref.dim() else ref.reshape(1)
        e = capi.Expr(ctx, expr, inputs=["x", "y", "n"], layouts=[capi.VAR_REAL, capi.VAR_REAL, capi.VAR_SCALAR],
                      derivatives=deriv)
        assert str(e) == str(f), expr
        got = e.eval([dx, dy, dn]).cpu()
        if e.space == 0:
            got = got.reshape(1)
            ref = ref.reshape(-1)[:1]
        err = ((got - ref).abs() / (ref.abs() + 1e-30)).max().item()
        assert err < 1e-12 or (got - ref).abs().max().item() < 1e-13, (expr, deriv, err)
        e.close()


@pytest.mark.parametrize("dim,shape", [(1, (30,)), (2, (12, 10)), (3, (6, 8, 10)), (3, (5, 7, 9))])
def test_extra_symbols_and_spaces(ctx, dim, shape):
    """ParsedCompute extra_symbols (src/tensor_computes/ParsedCompute.C:73-74,146-159): x y z / kx ky kz k2 / t / pi e i,
    real and reciprocal index spaces, complex results."""
    mx = tuple(2.0 + d for d in range(dim))
    ctx.domain_set(dim, shape, (0,) * dim, mx)
    d = om.Domain(dim, list(shape), (0,) * 3, mx + (1.0,) * (3 - dim))
    p = om.Problem(d)
    p.sub_time = 0.37
    op = om.ParsedCompute(p, "u", "sin(x)*cos(2*y)+z*t+pi-e", extra_symbols=True, expand="REAL")
    op.compute()
    e = capi.Expr(ctx, "sin(x)*cos(2*y)+z*t+pi-e", extra_symbols=True, expand=capi.EXPAND_REAL)
    assert e.space == 1 and not e.is_complex
    assert rel_l2(e.eval([], t=0.37).cpu(), p.buf["u"]) < 1e-14
    # reciprocal space, complex: i*kx*cbar*exp(-k2*t) + M*cbar
    g = torch.Generator().manual_seed(2)
    cbar = torch.randn(d.rshape, dtype=torch.complex128, generator=g)
    M = torch.rand(d.rshape, dtype=torch.float64, generator=g)
    p.buf["cbar"], p.buf["M"] = cbar, M
    expr = "i*kx*cbar*exp(-k2*t) + M*cbar/(1+ky^2+kz^2)"
    op = om.ParsedCompute(p, "v", expr, inputs=["cbar", "M"], extra_symbols=True)
    op.compute()
    e2 = capi.Expr(ctx, expr, inputs=["cbar", "M"], layouts=[capi.VAR_RECIP_COMPLEX, capi.VAR_RECIP_REAL],
                   extra_symbols=True)
    assert e2.space == 2 and e2.is_complex
    got = e2.eval([cbar.cuda(), M.cuda()], t=0.37).cpu()
    assert rel_l2(torch.view_as_real(got), torch.view_as_real(p.buf["v"])) < 1e-14
    # all-constant expression -> one value (ParsedJITTensor.C:148-153)
    e3 = capi.Expr(ctx, "2*A+sqrt(4)", constants={"A": 1.5})
    assert e3.space == 0 and float(e3.eval([]).cpu()[0]) == 5.0


def test_expression_float32(ctx32):
    ctx32.domain_set(2, (16, 16), (0, 0), (1.0, 1.0))
    x = torch.rand((16, 16), dtype=torch.float32) + 0.5
    e = capi.Expr(ctx32, "0.1*c^2*(c-1)^2", inputs=["c"], derivatives=["c"])
    ref = 0.1 * (2.0 * x.double()) * (x.double() - 1) ** 2 + 0.1 * x.double() ** 2 * (2.0 * (x.double() - 1))
    assert rel_l2(e.eval([x.cuda()]).cpu(), ref) < 1e-5


def _run_expr_split(ctx, p, mu_expr, steps, dt, substeps, order=2, M=0.2, kappa=-0.001, constants=None):
    d = p.domain
    ctx.domain_set(d.dim, d.n[:d.dim], d.min, d.max)
    c = p.buf["c"].to(ctx.rdtype).cuda().contiguous()
    e = capi.Expr(ctx, mu_expr, inputs=["c"], derivatives=["c"], constants=constants)
    mu = torch.zeros_like(c)
    plan = ctx.split_plan(expr=e, expr_var=0, expr_inputs=[c], M_factor=M, L_factor=kappa, history=order - 1, g_out=mu)
    drv = SplitDriver(plan, c, substeps, predictor_order=order)
    for _ in range(steps):
        drv.step(dt)
    out, muo = c.cpu(), mu.cpu()
    plan.close()
    e.close()
    return out, muo


@pytest.mark.parametrize("dim,n,L", [(2, 20, 3.0), (2, 64, 8.0), (2, 200, 25.0), (3, 20, 2.5), (3, 32, 4.0), (3, 128, 16.0)])
def test_ch_with_compiled_nonlinearity_matches_oracle(ctx, dim, n, L):
    """The CH substep with mu = d/dc[expression] compiled into the first FFT pass (generic, register
    and TMA variants of the pass), 100 substeps over 2 MOOSE steps, rel L2 <= 1e-10."""
    sub = 50 if n < 128 else 10
    p = oc.ch_problem(dim, n, L, substeps=sub)
    p.initial()
    got, mu = _run_expr_split(ctx, p, "0.1*c^2*(c-1)^2", 2, 0.05 * sub / 50, sub)
    for _ in range(2):
        p.step(0.05 * sub / 50)
    assert rel_l2(got, p.buf["c"]) < 1e-10
    # `mu` is materialised from the values the last substep started from
    assert torch.isfinite(mu).all() and float(mu.abs().max()) > 0


def test_ch_other_free_energy(ctx):
    """A non-polynomial free energy (log terms) - nothing is special-cased for the double well."""
    expr = "c*log(c)+(1-c)*log(1-c)+w*c*(1-c)"
    p = oc.ch_problem(2, 64, 8.0, substeps=20, mu_expr=expr, constant_names=["w"], constant_expressions=["2.5"])
    p.initial()
    got, _ = _run_expr_split(ctx, p, expr, 2, 0.01, 20, constants={"w": 2.5})
    for _ in range(2):
        p.step(0.01)
    assert rel_l2(got, p.buf["c"]) < 1e-10


def test_coupled_two_variable_solver_matches_oracle_and_gold(ctx):
    """test/tests/solvers/diagonal.i (two coupled fields, 150^2, AB2): every variable's nonlinearity is
    evaluated from the OLD fields (mrl_split_forward on all plans) before any variable is updated
    (mrl_split_finish) - SplitOperatorBase.C:39-64 / AdamsBashforthMoulton.C:60-101."""
    p = oc.diagonal_problem(10, 0, 2)
    p.initial()
    d = p.domain
    ctx.domain_set(2, d.n[:2], d.min, d.max)
    u = p.buf["u"].cuda().contiguous()
    v = p.buf["v"].expand(d.shape).cuda().contiguous()
    Du, Dv = p.buf["Du"].cuda().contiguous(), p.buf["Dv"].cuda().contiguous()
    consts = {"A": 1.0, "B": 3.5}
    eu = capi.Expr(ctx, "A - (B+1)*u +u^2*v", inputs=["u", "v"], constants=consts)
    ev = capi.Expr(ctx, "B*u - u^2*v", inputs=["u", "v"], constants=consts)
    pu = ctx.split_plan(expr=eu, expr_var=0, expr_inputs=[u, v], M_identity=True, L_buffer=Du, history=1)
    pv = ctx.split_plan(expr=ev, expr_var=1, expr_inputs=[u, v], M_identity=True, L_buffer=Dv, history=1)
    t_step, stored, dt = 0, 0, 0.1
    for step in range(3):
        t_step += 1
        if t_step > 1:
            stored = pu.advance_state()
            pv.advance_state()
        for s in range(10):
            order = min(stored, 1)
            pu.forward(u)
            pv.forward(v)
            pu.finish(u, dt / 10, AB_BETA[order], order)
            pv.finish(v, dt / 10, AB_BETA[order], order)
            if s < 9 and t_step > 1:
                stored = pu.advance_state()
                pv.advance_state()
        p.step(dt)
        assert rel_l2(u.cpu(), p.buf["u"]) < 1e-10 and rel_l2(v.cpu(), p.buf["v"]) < 1e-10, step

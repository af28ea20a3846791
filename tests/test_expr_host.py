"""Host side of the device expression compiler (mrl_expr_*): the C++ parser / differentiator /
simplifier must reproduce the oracle's (= the reference's) strings, reject what the reference
rejects, and every generated kernel must compile with NVRTC for sm_100a.  No GPU needed."""
import math

import pytest

from marlin_b200 import capi
from oracle import exprparser as xp


def oracle_string(expr, derivatives=(), constants=()):
    ast = xp.parse(expr, constants)
    for d in derivatives:
        ast = xp.differentiate(ast, d)
    return xp.to_string(xp.simplify(ast))


CORPUS = [
    ("0.1*c^2*(c-1)^2", ["c"], ("c",), {}),
    ("0.1*c^2*(c-1)^2", ["c"], ("c", "c"), {}),
    ("A - (B+1)*u +u^2*v", ["u", "v"], (), {"A": 1.0, "B": 3.5}),
    ("B*u - u^2*v", ["u", "v"], ("u",), {"A": 1.0, "B": 3.5}),
    ("hypot(x,y)", ["x", "y"], ("x",), {}),
    ("sqrt(x^2+y^2+n)", ["x", "y", "n"], ("y",), {}),
    ("tan((x-y)/2)", ["x", "y"], ("x",), {}),
    ("tanh(x-y)", ["x", "y"], ("y",), {}),
    ("atan(x + y)", ["x", "y"], ("x",), {}),
    ("asin((x * y / 2) / n)", ["x", "y", "n"], ("x",), {}),
    ("acosh(x+y+1)", ["x", "y"], ("x",), {}),
    ("atan2(x,y)", ["x", "y"], ("y",), {}),
    ("1/sqrt(x+y)", ["x", "y"], ("x",), {}),
    ("rsqrt(x*y)", ["x", "y"], ("x",), {}),
    ("exp2(x*y)", ["x", "y"], ("x",), {}),
    ("(x*y) % 1.5", ["x", "y"], (), {}),
    ("pow(y, x)", ["x", "y"], ("x",), {}),
    ("min(x^3,y^2)", ["x", "y"], ("x",), {}),
    ("max(x^2,sin(4*y))", ["x", "y"], ("y",), {}),
    ("if(x<1 | y>=2, x, y)", ["x", "y"], ("x",), {}),
    ("if(x<=1 & y>2, x*x, 3*y)", ["x", "y"], (), {}),
    ("r2:=x^2+y^2; sqrt(r2)", ["x", "y"], ("x",), {}),
    ("x2:=x^2; sinx2:=sin(x2); 4*sinx2", ["x"], ("x",), {}),
    ("r:=sqrt(a^2+1); r^2", ["a"], ("a",), {}),     # test/tests/parsed_tensor/local_vars_derivative.i
    ("a:=sin(x^2); a + 2*a + 3*a", ["x"], ("x",), {}),
    ("x^y", ["x", "y"], ("x",), {}),
    ("x^y", ["x", "y"], ("y",), {}),
    ("-x^2", ["x"], ("x",), {}),
    ("2^3^2", [], (), {}),
    ("!(x>1)", ["x"], (), {}),
    ("abs(x-1)*log10(x)+log2(x)+cosh(x)*sinh(x)", ["x"], ("x",), {}),
    ("acos(x)+asinh(x)+atanh(x)", ["x"], ("x",), {}),
    ("round(x)+ceil(x)+floor(x)+trunc(x)", ["x"], ("x",), {}),
    ("(x + 0) * 1 + 0", ["x"], (), {}),
    ("sqrt(4) + log(1) + exp(0)", [], (), {}),
    ("1e-3*x + 2.5E+2", ["x"], (), {}),
]


@pytest.mark.parametrize("expr,inputs,derivs,consts", CORPUS)
def test_strings_match_oracle(expr, inputs, derivs, consts):
    ours = capi.expr_simplified(expr, inputs=inputs, derivatives=derivs, constants=consts)
    assert ours == oracle_string(expr, derivs, consts.keys())


def test_reference_known_strings():
    """unit/src/ParsedTensorTest.C:411-543 (Simplify)."""
    s = lambda e, **kw: capi.expr_simplified(e, **kw)  # noqa: E731
    assert s("2 + 3") == "5.000000"
    assert s("2 ^ 3") == "8.000000"
    assert s("x * 0", inputs=["x"]) == "0.000000"
    assert s("x ^ 1", inputs=["x"]) == "x"
    assert s("a := 2 + 3; a * x", inputs=["x"]) == "a:=5.000000; (a * x)"
    assert s("x + y", inputs=["x", "y", "z"], derivatives=["z"]) == "0.000000"


@pytest.mark.parametrize("bad", ["x + ", "(x + y", "x + y)", "sin(x", "a := ; x + a", "x + * y", "", "1.2.3 + x",
                                 "x^-1", ".5*x", "x $ y"])
def test_rejects(bad):
    with pytest.raises(capi.MarlinError):
        capi.expr_simplified(bad, inputs=["x", "y"])


def test_derivative_must_be_an_input():
    with pytest.raises(capi.MarlinError, match="not listed in `inputs`"):
        capi.expr_simplified("x*q", inputs=["x"], derivatives=["q"])


def test_constant_expressions():
    """libMesh FParser stand-in: benchmarks/02_oswald_ripening/2a.i uses sqrt(2)."""
    assert capi.expr_constant("sqrt(2)") == math.sqrt(2.0)
    assert capi.expr_constant("2*a+pi", {"a": 1.5}) == 3.0 + math.pi
    assert capi.expr_constant("3.5") == 3.5


@pytest.mark.parametrize("expr,inputs,derivs,consts", CORPUS)
def test_generated_kernels_compile_for_sm100a(expr, inputs, derivs, consts):
    src = capi.expr_check(expr, inputs=inputs, derivatives=derivs, constants=consts)
    assert "mrl_expr_u32" in src


def test_generated_kernel_complex_and_symbols():
    src = capi.expr_check("Mbar*mubar", inputs=["Mbar", "mubar"],
                          layouts=[capi.VAR_RECIP_REAL, capi.VAR_RECIP_COMPLEX])
    assert "cx" in src
    src = capi.expr_check("i*kx*cbar*exp(-k2*t)", inputs=["cbar"], layouts=[capi.VAR_RECIP_COMPLEX],
                          extra_symbols=True, precision=capi.F32)
    assert "e_k2" in src and "typedef float T" in src
    capi.expr_check("sin(x)*sin(y)+pi", extra_symbols=True, expand=capi.EXPAND_REAL)
    with pytest.raises(capi.MarlinError, match="mixes real-space and reciprocal-space"):
        capi.expr_check("x*kx", extra_symbols=True)


@pytest.mark.parametrize("n,prec", [(512, capi.F64), (200, capi.F64), (64, capi.F32), (1024, capi.F32)])
def test_fused_first_pass_compiles(n, prec):
    """The expression is compiled INTO the z r2c pass (TMA, register and generic variants)."""
    capi.expr_check_fused("rho_s*(c-c_alpha)^2*(c_beta-c)^2", n, staged_var=0, precision=prec, inputs=["c"],
                          derivatives=["c"], constants={"rho_s": 5, "c_alpha": 0.3, "c_beta": 0.7})
    capi.expr_check_fused("A - (B+1)*u +u^2*v + 0*t", n, staged_var=0, precision=prec, inputs=["u", "v"],
                          constants={"A": 1, "B": 3.5}, extra_symbols=True)


def test_fused_first_pass_rejects_unsupported():
    with pytest.raises(capi.MarlinError, match="coordinate symbol"):
        capi.expr_check_fused("c*x", 64, inputs=["c"], extra_symbols=True)

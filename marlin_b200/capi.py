"""ctypes binding of libmarlin_b200.so (the C ABI in include/marlin_b200.h).

PyTorch is used here only as plumbing: it owns the device allocations handed to the
library as raw pointers.  There is no CPU or eager fallback: if the shared library is
missing, or no CUDA device is present, calls raise.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmarlin_b200.so")

F64, F32 = 0, 1
KFACTOR_LAPLACIAN, KFACTOR_LAPLACIAN_SQUARE = 0, 1
SUM, MIN, MAX, SUMSQ = 0, 1, 2, 3
NONLIN_DOUBLE_WELL, NONLIN_EXPR = 0, 1


# Adams-Bashforth predictor coefficients exactly as coded in the reference
# (src/tensor_solver/AdamsBashforthMoulton.C:67-73; the AB5 leading entry 190/720 is the
# reference's value and is reproduced as is).
AB_BETA = [
    [1.0, 0.0, 0.0, 0.0, 0.0],
    [3.0 / 2.0, -1.0 / 2.0, 0.0, 0.0, 0.0],
    [23.0 / 12.0, -16.0 / 12.0, 5.0 / 12.0, 0.0, 0.0],
    [55.0 / 24.0, -59.0 / 24.0, 37.0 / 24.0, -9.0 / 24.0, 0.0],
    [190.0 / 720.0, -2774.0 / 720.0, 2616.0 / 720.0, -1274.0 / 720.0, 251.0 / 720.0],
]


class MarlinError(RuntimeError):
    pass


class SplitDesc(C.Structure):
    _fields_ = [
        ("nonlin_kind", C.c_int), ("nonlin_params", C.c_double * 4), ("nonlin_expr", C.c_void_p),
        ("M_closed_form", C.c_int), ("M_factor", C.c_double), ("M_real_dev", C.c_void_p),
        ("has_L", C.c_int), ("L_closed_form", C.c_int), ("L_factor", C.c_double),
        ("L_real_dev", C.c_void_p), ("history", C.c_int), ("g_out_real_dev", C.c_void_p),
        ("nonlin_var", C.c_int), ("nonlin_inputs_dev", C.c_void_p * 16),
    ]


_lib = None


def lib():
    """Load the shared library (built by `make` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MarlinError(f"{LIB_PATH} not found: build it with `make` (nvcc, sm_100a). "
                              "marlin_b200 has no fallback path.")
        _lib = C.CDLL(LIB_PATH)
        _lib.mrl_last_error.restype = C.c_char_p
        _lib.mrl_version.restype = C.c_char_p
    return _lib


def _ck(rc):
    if rc != 0:
        raise MarlinError(f"marlin_b200 error {rc}: {lib().mrl_last_error().decode()}")


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def axis_values(n, mn, mx, reciprocal=False, half=False):
    """Host-side axis arithmetic (no device needed)."""
    m = (n // 2 + 1) if (reciprocal and half) else n
    out = (C.c_double * m)()
    _ck(lib().mrl_axis_values(C.c_int64(n), C.c_double(mn), C.c_double(mx), int(reciprocal), int(half),
                              out))
    return list(out)


class Context:
    """One device context (mirrors DomainAction + the global device/precision choice)."""

    def __init__(self, device=0, precision=F64):
        self.h = C.c_void_p()
        _ck(lib().mrl_create(int(device), int(precision), C.byref(self.h)))
        self.device = torch.device("cuda", device)
        self.precision = precision
        self.rdtype = torch.float64 if precision == F64 else torch.float32
        self.cdtype = torch.complex128 if precision == F64 else torch.complex64
        self.dim = 0

    def close(self):
        if self.h:
            lib().mrl_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def use_torch_stream(self):
        _ck(lib().mrl_set_stream(self.h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def synchronize(self):
        _ck(lib().mrl_synchronize(self.h))

    # ---- staged host transfers (overlap with the compute stream; host tensors should be pinned)
    def upload_staged(self, dev, host):
        assert host.is_contiguous() and dev.is_contiguous() and host.numel() * host.element_size() == dev.numel() * dev.element_size()
        _ck(lib().mrl_upload_staged(self.h, C.c_void_p(dev.data_ptr()), C.c_void_p(host.data_ptr()),
                                    C.c_size_t(host.numel() * host.element_size())))

    def download_staged(self, host, dev):
        assert host.is_contiguous() and dev.is_contiguous() and host.numel() * host.element_size() == dev.numel() * dev.element_size()
        _ck(lib().mrl_download_staged(self.h, C.c_void_p(host.data_ptr()), C.c_void_p(dev.data_ptr()),
                                      C.c_size_t(host.numel() * host.element_size())))

    def staged_wait(self):
        _ck(lib().mrl_staged_wait(self.h))

    def launch_count(self):
        v = C.c_int64()
        _ck(lib().mrl_launch_count(self.h, C.byref(v)))
        return v.value

    # ---- domain
    def domain_set(self, dim, n, mins=(0.0, 0.0, 0.0), maxs=(1.0, 1.0, 1.0)):
        n3 = (C.c_int64 * 3)(*[int(n[d]) if d < dim else 1 for d in range(3)])
        mn = (C.c_double * 3)(*[float(mins[d]) if d < len(mins) else 0.0 for d in range(3)])
        mx = (C.c_double * 3)(*[float(maxs[d]) if d < len(maxs) else 1.0 for d in range(3)])
        _ck(lib().mrl_domain_set(self.h, int(dim), n3, mn, mx))
        self.dim = dim
        rs, ks = (C.c_int64 * 3)(), (C.c_int64 * 3)()
        _ck(lib().mrl_domain_shape(self.h, rs, ks))
        self.shape = [rs[d] for d in range(dim)]
        self.rshape = [ks[d] for d in range(dim)]

    def axis(self, d, reciprocal=False):
        m = (self.rshape[d] if reciprocal else self.shape[d]) if d < self.dim else 1
        out = (C.c_double * m)()
        _ck(lib().mrl_domain_axis(self.h, d, int(reciprocal), out))
        return torch.tensor(list(out), dtype=torch.float64)

    # ---- FFT
    def _batch(self, t, shape):
        nb = t.dim() - len(shape)
        if nb < 0 or list(t.shape[nb:]) != list(shape):
            raise MarlinError(f"tensor shape {tuple(t.shape)} does not end with the domain shape {shape}")
        b = 1
        for s in t.shape[:nb]:
            b *= s
        return b, list(t.shape[:nb])

    def rfftn(self, t):
        assert t.is_cuda and t.dtype == self.rdtype and t.is_contiguous()
        b, lead = self._batch(t, self.shape)
        out = torch.empty(lead + self.rshape, dtype=self.cdtype, device=t.device)
        _ck(lib().mrl_rfftn(self.h, _p(t), _p(out), b))
        return out

    def irfftn(self, t):
        assert t.is_cuda and t.dtype == self.cdtype and t.is_contiguous()
        b, lead = self._batch(t, self.rshape)
        out = torch.empty(lead + self.shape, dtype=self.rdtype, device=t.device)
        _ck(lib().mrl_irfftn(self.h, _p(t), _p(out), b))
        return out

    # ---- pointwise
    def kfactor(self, kind, factor):
        out = torch.empty(self.rshape, dtype=self.rdtype, device=self.device)
        _ck(lib().mrl_kfactor(self.h, int(kind), C.c_double(factor), _p(out)))
        return out

    def mul_real_complex(self, a, b):
        out = torch.empty_like(b)
        _ck(lib().mrl_mul_real_complex(self.h, _p(a), _p(b), _p(out)))
        return out

    def ab_update(self, cbar, N, L, dt, beta, Nold=()):
        out = torch.empty_like(cbar)
        nold = len(Nold)
        arr = (C.c_void_p * max(nold, 1))(*[t.data_ptr() for t in Nold])
        b = (C.c_double * 5)(*(list(beta) + [0.0] * 5)[:5])
        _ck(lib().mrl_ab_update(self.h, _p(out), _p(cbar), _p(N), _p(L), C.c_double(dt), b, nold, arr))
        return out

    def coupled_solve(self, L, rhs, dt, drop_imag=False):
        """(I - dt*L) ubar = rhs per wavevector; L: nvar x nvar nested list of real reciprocal-space
        tensors (None = 0), rhs: list of complex tensors.  Returns the list of solutions."""
        n = len(rhs)
        out = [torch.empty_like(r) for r in rhs]
        Lp = (C.c_void_p * (n * n))(*[(L[r][c].data_ptr() if L[r][c] is not None else None) for r in range(n) for c in range(n)])
        bp = (C.c_void_p * n)(*[t.data_ptr() for t in rhs])
        op = (C.c_void_p * n)(*[t.data_ptr() for t in out])
        _ck(lib().mrl_coupled_solve(self.h, n, Lp, bp, op, C.c_double(dt), int(bool(drop_imag))))
        return out

    def broyden_step(self, M, R, u):
        """sk = -M R, unew = u + 0.5 sk per wavevector; M: [n*n, *rshape] complex (component major)."""
        n = len(R)
        sk = [torch.empty_like(r) for r in R]
        un = [torch.empty_like(r) for r in R]
        arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
        _ck(lib().mrl_broyden_step(self.h, n, _p(M), arr(R), arr(u), arr(sk), arr(un)))
        return sk, un

    def broyden_update(self, M, sk, R, Rnew):
        """in place: M += (sk - M yk) sk^T / (sk^T yk) where |sk^T yk| > 1e-12, yk = Rnew - R."""
        n = len(R)
        arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
        _ck(lib().mrl_broyden_update(self.h, n, _p(M), arr(sk), arr(R), arr(Rnew)))

    def reduce(self, op, t):
        assert t.is_contiguous() and t.dtype == self.rdtype
        v = C.c_double()
        _ck(lib().mrl_reduce(self.h, int(op), _p(t), C.c_int64(t.numel()), C.byref(v)))
        return v.value

    # ---- fused split-operator plan
    def split_plan(self, **kw):
        return SplitPlan(self, **kw)


class SplitPlan:
    """Fused semi-implicit substep of one variable (AdamsBashforthMoulton::substep for the
    canonical Cahn-Hilliard compute graph)."""

    def __init__(self, ctx, double_well=None, expr=None, M_factor=None, M_buffer=None, L_factor=None,
                 L_buffer=None, has_L=True, history=1, g_out=None, expr_var=0, expr_inputs=(), M_identity=False):
        self.ctx = ctx
        d = SplitDesc()
        if expr is not None:
            d.nonlin_kind = NONLIN_EXPR
            d.nonlin_expr = expr.h
            d.nonlin_var = int(expr_var)
            for i, t in enumerate(expr_inputs):
                d.nonlin_inputs_dev[i] = t.data_ptr() if t is not None else None
        else:
            d.nonlin_kind = NONLIN_DOUBLE_WELL
            A, a, b = double_well
            d.nonlin_params = (C.c_double * 4)(A, a, b, 0.0)
        d.M_closed_form = 2 if M_identity else int(M_buffer is None)
        d.M_factor = float(M_factor or 0.0)
        d.M_real_dev = M_buffer.data_ptr() if M_buffer is not None else None
        d.has_L = int(has_L)
        d.L_closed_form = int(L_buffer is None)
        d.L_factor = float(L_factor or 0.0)
        d.L_real_dev = L_buffer.data_ptr() if L_buffer is not None else None
        d.history = history
        d.g_out_real_dev = g_out.data_ptr() if g_out is not None else None
        self._keep = (M_buffer, L_buffer, g_out, expr, tuple(expr_inputs))
        self.h = C.c_void_p()
        _ck(lib().mrl_split_plan_create(ctx.h, C.byref(d), C.byref(self.h)))
        self.launches_per_substep = lib().mrl_split_launches_per_substep(self.h)

    def substep(self, c, dt, beta, nold):
        b = (C.c_double * 5)(*(list(beta) + [0.0] * 5)[:5])
        _ck(lib().mrl_split_substep(self.h, _p(c), C.c_double(dt), b, int(nold)))

    def substeps(self, c, dt, beta, nold, count):
        """`count` x (substep + advance_state) with fixed dt / beta / nold; the periodic part is replayed
        from a CUDA graph when the context runs on a capturable (non-default) stream."""
        b = (C.c_double * 5)(*(list(beta) + [0.0] * 5)[:5])
        _ck(lib().mrl_split_substeps(self.h, _p(c), C.c_double(dt), b, int(nold), int(count)))

    def forward(self, c):
        _ck(lib().mrl_split_forward(self.h, _p(c)))

    def finish(self, c, dt, beta, nold):
        b = (C.c_double * 5)(*(list(beta) + [0.0] * 5)[:5])
        _ck(lib().mrl_split_finish(self.h, _p(c), C.c_double(dt), b, int(nold)))

    def set_time(self, t):
        _ck(lib().mrl_split_set_time(self.h, C.c_double(t)))

    def substep_timed(self, c, dt, beta, nold):
        b = (C.c_double * 5)(*(list(beta) + [0.0] * 5)[:5])
        ms = (C.c_float * 8)()
        _ck(lib().mrl_split_substep_timed(self.h, _p(c), C.c_double(dt), b, int(nold), ms))
        return [ms[i] for i in range(self.launches_per_substep)]

    def advance_state(self):
        v = C.c_int()
        _ck(lib().mrl_split_advance_state(self.h, C.byref(v)))
        return v.value

    def clear_states(self):
        _ck(lib().mrl_split_clear_states(self.h))

    def close(self):
        if self.h:
            lib().mrl_split_plan_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- runtime expressions (ParsedCompute) ------------------------------------------------------
VAR_REAL, VAR_RECIP_REAL, VAR_RECIP_COMPLEX, VAR_SCALAR, VAR_REAL_COMPLEX = 0, 1, 2, 3, 4
EXPAND_NONE, EXPAND_REAL, EXPAND_RECIPROCAL = 0, 1, 2


class ExprDesc(C.Structure):
    _fields_ = [("expression", C.c_char_p), ("nvars", C.c_int), ("var_names", C.POINTER(C.c_char_p)),
                ("var_layouts", C.POINTER(C.c_int)), ("nderivatives", C.c_int),
                ("derivatives", C.POINTER(C.c_char_p)), ("nconstants", C.c_int),
                ("constant_names", C.POINTER(C.c_char_p)), ("constant_values", C.POINTER(C.c_double)),
                ("extra_symbols", C.c_int), ("expand", C.c_int)]


def _expr_desc(expression, inputs=(), layouts=None, derivatives=(), constants=None, extra_symbols=False,
               expand=EXPAND_NONE):
    constants = dict(constants or {})
    d = ExprDesc()
    d.expression = expression.encode()
    keep = [d.expression]

    def strs(xs):
        arr = (C.c_char_p * max(len(xs), 1))(*[x.encode() for x in xs])
        keep.append(arr)
        return C.cast(arr, C.POINTER(C.c_char_p))

    d.nvars = len(inputs)
    d.var_names = strs(list(inputs))
    lay = (C.c_int * max(len(inputs), 1))(*[int(x) for x in (layouts or [VAR_REAL] * len(inputs))])
    keep.append(lay)
    d.var_layouts = C.cast(lay, C.POINTER(C.c_int))
    d.nderivatives = len(derivatives)
    d.derivatives = strs(list(derivatives))
    d.nconstants = len(constants)
    d.constant_names = strs(list(constants.keys()))
    vals = (C.c_double * max(len(constants), 1))(*[float(v) for v in constants.values()])
    keep.append(vals)
    d.constant_values = C.cast(vals, C.POINTER(C.c_double))
    d.extra_symbols = int(bool(extra_symbols))
    d.expand = int(expand)
    return d, keep


def expr_simplified(expression, **kw):
    """Host only: toString() of the parsed, differentiated, simplified expression."""
    d, keep = _expr_desc(expression, **kw)
    buf = C.create_string_buffer(1 << 16)
    _ck(lib().mrl_expr_simplified(C.byref(d), buf, C.c_size_t(len(buf))))
    return buf.value.decode()


def expr_constant(expression, constants=None):
    """Host only: value of one `constant_expressions` entry (may use earlier constants, pi, e)."""
    constants = dict(constants or {})
    names = (C.c_char_p * max(len(constants), 1))(*[k.encode() for k in constants])
    vals = (C.c_double * max(len(constants), 1))(*[float(v) for v in constants.values()])
    out = C.c_double()
    _ck(lib().mrl_expr_constant(expression.encode(), len(constants), names, vals, C.byref(out)))
    return out.value


def expr_check(expression, precision=F64, **kw):
    """Host only: generate the CUDA source of the expression kernel and compile it with NVRTC for
    sm_100a (no device needed).  Returns the generated source."""
    d, keep = _expr_desc(expression, **kw)
    buf = C.create_string_buffer(1 << 18)
    _ck(lib().mrl_expr_check(C.byref(d), int(precision), buf, C.c_size_t(len(buf))))
    return buf.value.decode()


def expr_check_fused(expression, n, staged_var=0, precision=F64, **kw):
    """Host only: NVRTC-compile the first FFT pass specialised for the expression (axis length n)."""
    d, keep = _expr_desc(expression, **kw)
    _ck(lib().mrl_expr_check_fused(C.byref(d), int(precision), int(n), int(staged_var)))


class Expr:
    """One compiled ParsedCompute expression on a context's device."""

    def __init__(self, ctx, expression, inputs=(), layouts=None, derivatives=(), constants=None,
                 extra_symbols=False, expand=EXPAND_NONE):
        self.ctx = ctx
        self.inputs = list(inputs)
        d, keep = _expr_desc(expression, inputs=inputs, layouts=layouts, derivatives=derivatives,
                             constants=constants, extra_symbols=extra_symbols, expand=expand)
        self.h = C.c_void_p()
        _ck(lib().mrl_expr_compile(ctx.h, C.byref(d), C.byref(self.h)))
        sp, cplx = C.c_int(), C.c_int()
        _ck(lib().mrl_expr_result(self.h, C.byref(sp), C.byref(cplx)))
        self.space, self.is_complex = sp.value, bool(cplx.value)

    def __str__(self):
        buf = C.create_string_buffer(1 << 16)
        _ck(lib().mrl_expr_string(self.h, buf, C.c_size_t(len(buf))))
        return buf.value.decode()

    def eval(self, tensors=(), t=0.0):
        ctx = self.ctx
        shape = [1] if self.space == 0 else (ctx.rshape if self.space == 2 else ctx.shape)
        out = torch.empty(shape, dtype=ctx.cdtype if self.is_complex else ctx.rdtype, device=ctx.device)
        arr = (C.c_void_p * max(len(tensors), 1))(*[x.data_ptr() if x is not None else None for x in tensors])
        _ck(lib().mrl_expr_eval(self.h, arr, C.c_double(t), _p(out)))
        return out

    def close(self):
        if self.h:
            lib().mrl_expr_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def partition(total, nranks, weights=None):
    """DomainAction::partitionHepler (include/actions/DomainAction.h:249-280). Host only."""
    cnt = (C.c_int64 * nranks)()
    w = (C.c_double * nranks)(*[float(x) for x in weights]) if weights is not None else None
    _ck(lib().mrl_partition(C.c_int64(total), int(nranks), w, cnt))
    return list(cnt)


class MechDesc(C.Structure):
    _fields_ = [("l_tol", C.c_double), ("l_max_its", C.c_int64), ("nl_rel_tol", C.c_double),
                ("nl_abs_tol", C.c_double), ("nl_max_its", C.c_int)]


class MechStats(C.Structure):
    _fields_ = [("newton_iterations", C.c_int), ("cg_solves", C.c_int), ("cg_iterations_total", C.c_int),
                ("cg_iterations", C.c_int * 64), ("final_rnorm", C.c_double), ("final_anorm", C.c_double)]


def components(ctx, t, ncomp, to_component_major):
    """[n][ncomp] (the reference's trailing value dimensions) <-> [ncomp][n]."""
    assert t.is_cuda and t.is_contiguous()
    out = torch.empty_like(t)
    n = t.numel() // ncomp
    _ck(lib().mrl_components(ctx.h, _p(t), _p(out), C.c_int64(n), int(ncomp), int(to_component_major)))
    return out


class MechPlan:
    """FFTMechanics + HyperElasticIsotropic behind the C ABI (component-major tensor fields
    [D*D][nx][ny][nz], D = dim = 2 or 3); see include/marlin_b200.h."""

    def __init__(self, ctx, K, mu, l_tol=1e-2, l_max_its=0, nl_rel_tol=1e-5, nl_abs_tol=1e-8, nl_max_its=100):
        self.ctx = ctx
        self._keep = (K, mu)
        d = MechDesc(float(l_tol), int(l_max_its or 0), float(nl_rel_tol), float(nl_abs_tol), int(nl_max_its))
        self.h = C.c_void_p()
        _ck(lib().mrl_mech_plan_create(ctx.h, C.byref(d), _p(K), _p(mu), C.byref(self.h)))

    def constitutive(self, F):
        P = torch.empty_like(F)
        _ck(lib().mrl_mech_constitutive(self.h, _p(F), _p(P)))
        return P

    def apply_G(self, A):
        out = torch.empty_like(A)
        _ck(lib().mrl_mech_apply_G(self.h, _p(A), _p(out)))
        return out

    def apply_GK(self, F, x):
        out = torch.empty_like(x)
        _ck(lib().mrl_mech_apply_GK(self.h, _p(F), _p(x), _p(out)))
        return out

    def solve(self, F, applied=None):
        """One FFTMechanics::computeBuffer; F is updated in place. Returns (P, stats)."""
        P = torch.empty_like(F)
        st = MechStats()
        a = (C.c_double * len(applied))(*[float(v) for v in applied]) if applied is not None else None
        _ck(lib().mrl_mech_solve(self.h, _p(F), a, _p(P), C.byref(st)))
        return P, st

    def close(self):
        if self.h:
            lib().mrl_mech_plan_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistContext(Context):
    """A context whose domain is one rank's part of a slab-decomposed grid (mrl_domain_set_dist): real space split
    along y, reciprocal space along x with the slab sizes of partitionHepler (include/actions/DomainAction.h:249-280)."""

    def domain_set_dist(self, dim, n, mins, maxs, rank, nranks, weights=None):
        n3 = (C.c_int64 * 3)(*[int(n[d]) if d < dim else 1 for d in range(3)])
        mn = (C.c_double * 3)(*[float(mins[d]) if d < len(mins) else 0.0 for d in range(3)])
        mx = (C.c_double * 3)(*[float(maxs[d]) if d < len(maxs) else 1.0 for d in range(3)])
        w = (C.c_double * nranks)(*[float(v) for v in weights]) if weights is not None else None
        _ck(lib().mrl_domain_set_dist(self.h, int(dim), n3, mn, mx, int(rank), int(nranks), w))
        self.dim, self.rank, self.nranks = dim, rank, nranks
        rs, rb, ks, kb = [(C.c_int64 * 3)() for _ in range(4)]
        _ck(lib().mrl_domain_local(self.h, rs, rb, ks, kb))
        self.shape, self.rbegin = [rs[d] for d in range(dim)], [rb[d] for d in range(dim)]
        self.rshape, self.kbegin = [ks[d] for d in range(dim)], [kb[d] for d in range(dim)]

    def domain_set_pencil(self, n, mins, maxs, rank, nranks):
        """FFT_PENCIL (partitionPencils, src/actions/DomainAction.C:569-742): real [nx][ny / Py][nz / Pz], reciprocal
        [(nx/2+1) / Py][ny / Pz][nz], rank = iz * Py + iy."""
        n3 = (C.c_int64 * 3)(*[int(v) for v in n])
        mn = (C.c_double * 3)(*[float(v) for v in mins])
        mx = (C.c_double * 3)(*[float(v) for v in maxs])
        _ck(lib().mrl_domain_set_pencil(self.h, 3, n3, mn, mx, int(rank), int(nranks)))
        self.dim, self.rank, self.nranks = 3, rank, nranks
        rs, rb, ks, kb = [(C.c_int64 * 3)() for _ in range(4)]
        _ck(lib().mrl_domain_local(self.h, rs, rb, ks, kb))
        self.shape, self.rbegin = list(rs), list(rb)
        self.rshape, self.kbegin = list(ks), list(kb)

    def rfftn(self, t):
        raise MarlinError("decomposed domain: use Dist.rfftn")

    irfftn = rfftn


class Dist:
    """DomainAction::fftSlab / ifftSlab (src/actions/DomainAction.C:870-938, :941-1019) through mrl_dist_*; the IPC
    handles travel over torch.distributed (any backend)."""

    def __init__(self, ctx, group=None):
        import torch.distributed as dist
        self.ctx = ctx
        self.h = C.c_void_p()
        _ck(lib().mrl_dist_create(ctx.h, C.byref(self.h)))
        if ctx.nranks > 1:
            mine = (C.c_ubyte * 128)()  # MRL_DIST_IPC_BYTES
            _ck(lib().mrl_dist_ipc_export(self.h, mine))
            t = torch.tensor(list(mine), dtype=torch.uint8, device=ctx.device if dist.get_backend(group) == "nccl" else "cpu")
            alls = [torch.empty_like(t) for _ in range(ctx.nranks)]
            dist.all_gather(alls, t, group=group)
            _ck(lib().mrl_dist_ipc_import(self.h, bytes(torch.cat(alls).cpu().tolist())))
            dist.barrier(group=group)

    def rfftn(self, t):
        c = self.ctx
        assert t.is_cuda and t.dtype == c.rdtype and t.is_contiguous()
        b, lead = c._batch(t, c.shape)
        out = torch.empty(lead + c.rshape, dtype=c.cdtype, device=t.device)
        _ck(lib().mrl_dist_rfftn(self.h, _p(t), _p(out), b))
        return out

    def irfftn(self, t):
        c = self.ctx
        assert t.is_cuda and t.dtype == c.cdtype and t.is_contiguous()
        b, lead = c._batch(t, c.rshape)
        out = torch.empty(lead + c.shape, dtype=c.rdtype, device=t.device)
        _ck(lib().mrl_dist_irfftn(self.h, _p(t), _p(out), b))
        return out

    def close(self):
        if self.h:
            lib().mrl_dist_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

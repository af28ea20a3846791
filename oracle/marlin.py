"""Oracle restatement of Marlin's spectral hot path on libTorch CPU kernels.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Structure: a `Domain` (grid contract +
fft/ifft), a `Problem` that owns named buffers with history (TensorProblem/TensorBuffer
semantics), operator classes with `compute()` (the reference's `computeBuffer()`), and the
solvers.  Every piece cites the reference file:line it follows.
"""
import math

import torch

from . import exprparser as xp

F64 = torch.float64


# =============================================================================== Domain
class Domain:
    """Grid contract: src/actions/DomainAction.C:227-338 (gridChanged), :1407-1434 (align),
    :1480-1509 (k-grid, k^2), :854-867 (fftSerial), :1054-1066 (ifft NONE branch)."""

    def __init__(self, dim, n, mins=(0.0, 0.0, 0.0), maxs=(1.0, 1.0, 1.0), dtype=F64):
        self.dim = dim
        self.n = [int(n[d]) if d < dim else 1 for d in range(3)]
        self.min = [float(mins[d]) for d in range(3)]
        self.max = [float(maxs[d]) for d in range(3)]
        self.dtype = dtype
        self.dx = [(self.max[d] - self.min[d]) / self.n[d] for d in range(3)]
        self.axis, self.kaxis = [], []
        for d in range(3):
            if d < dim:
                ax = torch.linspace(self.min[d] + self.dx[d] / 2.0, self.max[d] - self.dx[d] / 2.0,
                                    self.n[d], dtype=dtype)
                self.axis.append(self.align(ax, d))
                # serial mode: the LAST spatial dim is the halved one (:274)
                if d == dim - 1:
                    fr = torch.fft.rfftfreq(self.n[d], self.dx[d], dtype=dtype)
                else:
                    fr = torch.fft.fftfreq(self.n[d], self.dx[d], dtype=dtype)
                self.kaxis.append(self.align(fr * 2.0 * math.pi, d))
            else:
                self.axis.append(torch.tensor([0.0], dtype=dtype))
                self.kaxis.append(torch.tensor([0.0], dtype=dtype))
        self.shape = self.n[:dim]
        self.rshape = [self.n[d] if d < dim - 1 else self.n[d] // 2 + 1 for d in range(dim)]
        self.ncells = self.n[0] * self.n[1] * self.n[2]
        self.volume = 1.0
        for d in range(dim):
            self.volume *= self.max[d] - self.min[d]

    def align(self, t, d):
        shape = [1] * self.dim
        shape[d] = -1
        return t.reshape(shape)

    @property
    def k2(self):  # :1504-1509
        return self.kaxis[0] * self.kaxis[0] + self.kaxis[1] * self.kaxis[1] + \
            self.kaxis[2] * self.kaxis[2]

    @property
    def kgrid(self):  # :1480-1501
        if self.dim == 1:
            return self.kaxis[0]
        return torch.stack([self.kaxis[d].expand(self.rshape) for d in range(self.dim)], -1)

    def fft(self, t):
        return torch.fft.rfftn(t, dim=list(range(self.dim)))

    def ifft(self, t):
        return torch.fft.irfftn(t, s=self.shape, dim=list(range(self.dim)))

    def sum(self, t):  # :1559-1568
        return t.sum(dim=list(range(self.dim)))

    def average(self, t):  # :1570-1574
        return self.sum(t) / float(self.ncells)

    def value_shape(self, extra):
        return list(self.shape) + list(extra)


# =============================================================================== Problem
class Problem:
    """Buffer ownership + time bookkeeping: src/problems/TensorProblem.C:154-197 (execute),
    :451-472 (advanceState), include/tensor_buffers/TensorBuffer.h:64-79,112-116 (history)."""

    def __init__(self, domain):
        self.domain = domain
        self.buf = {}
        self.old = {}        # name -> list of old states (newest first)
        self.max_states = {}  # name -> requested history length
        self.time = 0.0
        self.time_old = 0.0
        self.dt = 0.0
        self.dt_old = 0.0
        self.t_step = 0
        self.sub_time = 0.0
        self.sub_dt = 0.0
        self.ics, self.computes, self.pps = [], [], []
        self.solver = None

    def get_old(self, name, states):
        self.max_states[name] = max(self.max_states.get(name, 0), states)
        return self.old.setdefault(name, [])

    def advance_state(self):
        if self.t_step <= 1:  # :455-456 (quirk Q1)
            return
        for name, mx in self.max_states.items():
            lst = self.old.setdefault(name, [])
            if len(lst) < mx:
                lst.append(None)
            if lst:
                for i in range(len(lst) - 1, 0, -1):
                    lst[i] = lst[i - 1]
                lst[0] = self.buf.get(name)

    # -- MOOSE Transient loop order: TransientBase.C:315-328,390-416; FixedPointSolve.C:402,463
    def initial(self):
        self.sub_time = self.time
        for ic in self.ics:
            ic.compute()
        for pp in self.pps:
            pp.compute()

    def step(self, dt):
        self.time_old = self.time
        self.t_step += 1
        self.advance_state()
        self.dt_old = self.dt if self.t_step > 1 else dt
        self.dt = dt
        self.time = self.time_old + dt
        self.sub_time = self.time_old
        if self.solver is not None:
            self.solver.compute()
        else:
            for c in self.computes:
                c.compute()
        for pp in self.pps:
            pp.compute()


class Op:
    def __init__(self, problem, buffer=None):
        self.p = problem
        self.d = problem.domain
        self.buffer = buffer

    def set(self, t):
        self.p.buf[self.buffer] = t

    def get(self, name):
        return self.p.buf[name]


class Group(Op):
    """Ordered list of computes (src/tensor_computes/ComputeGroup.C:50-88)."""

    def __init__(self, problem, ops):
        super().__init__(problem)
        self.ops = ops

    def compute(self):
        for o in self.ops:
            o.compute()


# =============================================================================== operators
class RandomTensor(Op):
    """src/tensor_computes/RandomTensor.C:37-55 (generate_on_cpu=true path)."""

    def __init__(self, problem, buffer, min, max, seed=None):
        super().__init__(problem, buffer)
        self.min, self.max, self.seed = min, max, seed

    def compute(self):
        if self.seed is not None:
            torch.manual_seed(self.seed)
        self.set(torch.rand(self.d.shape, dtype=self.d.dtype) * (self.max - self.min) + self.min)


class ConstantTensor(Op):
    """src/tensor_computes/ConstantTensor.C:45-53."""

    def __init__(self, problem, buffer, real=0.0, imaginary=0.0, reciprocal=False):
        super().__init__(problem, buffer)
        self.real, self.imag, self.reciprocal = real, imaginary, reciprocal

    def compute(self):
        if self.reciprocal:
            self.set(torch.complex(torch.full(self.d.rshape, self.real, dtype=self.d.dtype),
                                   torch.full(self.d.rshape, self.imag, dtype=self.d.dtype)))
        else:
            self.set(torch.full(self.d.shape, self.real, dtype=self.d.dtype))


class ReciprocalLaplacianFactor(Op):
    """src/tensor_computes/ReciprocalLaplacianFactor.C:30: u = -k2 * factor."""

    def __init__(self, problem, buffer, factor=1.0):
        super().__init__(problem, buffer)
        self.factor = factor

    def compute(self):
        self.set(-self.d.k2 * self.factor)


class ReciprocalLaplacianSquareFactor(Op):
    """src/tensor_computes/ReciprocalLaplacianSquareFactor.C:31: u = k2 * k2 * factor."""

    def __init__(self, problem, buffer, factor=1.0):
        super().__init__(problem, buffer)
        self.factor = factor

    def compute(self):
        self.set(self.d.k2 * self.d.k2 * self.factor)


class ForwardFFT(Op):
    """src/tensor_computes/PerformFFT.C:34-40."""

    def __init__(self, problem, buffer, input):
        super().__init__(problem, buffer)
        self.input = input

    def compute(self):
        self.set(self.d.fft(self.get(self.input)))


class InverseFFT(ForwardFFT):
    def compute(self):
        self.set(self.d.ifft(self.get(self.input)))


def eval_constant_expressions(names, expressions):
    """libMesh FParser stand-in for `constant_expressions` (ParsedCompute.C:104-123): each
    constant may use the previously evaluated ones."""
    vals = {}
    for n, e in zip(names, expressions):
        ast = xp.parse(str(e), ())
        env = dict(vals)
        env.update(pi=math.pi, e=math.e)
        v = xp.evaluate(xp.simplify(ast), env)
        vals[n] = float(v)
    return vals


class ParsedCompute(Op):
    """src/tensor_computes/ParsedCompute.C:50-181 (ctor), :184-265 (computeBuffer)."""

    def __init__(self, problem, buffer, expression, inputs=(), derivatives=(), constant_names=(),
                 constant_expressions=(), extra_symbols=False, expand="NONE"):
        super().__init__(problem, buffer)
        self.inputs = list(inputs)
        self.extra = extra_symbols
        self.expand = expand
        consts = {k: torch.tensor(v, dtype=self.d.dtype) for k, v in
                  eval_constant_expressions(constant_names, constant_expressions).items()}
        variables = list(self.inputs)
        if extra_symbols:
            consts["pi"] = torch.tensor(math.pi, dtype=self.d.dtype)
            consts["e"] = torch.tensor(math.e, dtype=self.d.dtype)
            consts["i"] = torch.tensor(1j, dtype=torch.complex128)
            variables += ["x", "kx", "y", "ky", "z", "kz", "k2", "t"]
        self.fn = xp.ParsedTensor(expression, variables, consts)
        for dv in derivatives:
            if dv not in self.inputs:
                raise ValueError(f"Derivative w.r.t `{dv}` was requested, but it is not listed "
                                 "in `inputs`.")
            self.fn.differentiate(dv)
        self.fn.compile()

    def compute(self):
        params = [self.get(n) for n in self.inputs]
        if self.extra:
            d = self.d
            params += [d.axis[0], d.kaxis[0], d.axis[1], d.kaxis[1], d.axis[2], d.kaxis[2], d.k2,
                       torch.tensor(self.p.sub_time, dtype=d.dtype)]
        u = self.fn.eval(params)
        if self.expand == "REAL":
            u = u.expand(self.d.shape)
        elif self.expand == "RECIPROCAL":
            u = u.expand(self.d.rshape)
        self.set(u)


class FFTGradient(Op):
    """src/tensor_computes/FFTGradient.C:36-40."""

    def __init__(self, problem, buffer, input, direction, input_is_reciprocal=False):
        super().__init__(problem, buffer)
        self.input, self.dir, self.recip = input, direction, input_is_reciprocal

    def compute(self):
        r = self.get(self.input) if self.recip else self.d.fft(self.get(self.input))
        self.set(self.d.ifft(r * self.d.kaxis[self.dir] * 1j))


class FFTGradientSquare(Op):
    """src/tensor_computes/FFTGradientSquare.C:37-48."""

    def __init__(self, problem, buffer, input, factor=1.0, input_is_reciprocal=False):
        super().__init__(problem, buffer)
        self.input, self.factor, self.recip = input, factor, input_is_reciprocal

    def compute(self):
        r = self.get(self.input) if self.recip else self.d.fft(self.get(self.input))
        u = self.d.ifft(r * self.d.kaxis[0] * 1j) ** 2
        for dd in range(1, self.d.dim):
            u = u + self.d.ifft(r * self.d.kaxis[dd] * 1j) ** 2
        if self.factor != 1.0:
            u = u * self.factor
        self.set(u)


# =============================================================================== solvers
AB_BETA = [  # src/tensor_solver/AdamsBashforthMoulton.C:67-73 (AB5 first entry as coded: Q3)
    [1.0, 0.0, 0.0, 0.0, 0.0],
    [3.0 / 2.0, -1.0 / 2.0, 0.0, 0.0, 0.0],
    [23.0 / 12.0, -16.0 / 12.0, 5.0 / 12.0, 0.0, 0.0],
    [55.0 / 24.0, -59.0 / 24.0, 37.0 / 24.0, -9.0 / 24.0, 0.0],
    [190.0 / 720.0, -2774.0 / 720.0, 2616.0 / 720.0, -1274.0 / 720.0, 251.0 / 720.0],
]
AM_ALPHA = [  # :108-114
    [1.0, 0.0, 0.0, 0.0, 0.0],
    [0.5, 0.5, 0.0, 0.0, 0.0],
    [5.0 / 12.0, 8.0 / 12.0, -1.0 / 12.0, 0.0, 0.0],
    [9.0 / 24.0, 19.0 / 24.0, -5.0 / 24.0, 1.0 / 24.0, 0.0],
    [251.0 / 720.0, 646.0 / 720.0, -264.0 / 720.0, 106.0 / 720.0, -19.0 / 720.0],
]


class TensorSolver:
    """src/tensor_solver/TensorSolver.C:93-110 (substep loop), :86-90 (forwardBuffers)."""

    def __init__(self, problem, root, substeps=1, forward=()):
        self.p, self.d = problem, problem.domain
        self.root = root
        self.substeps = substeps
        self.forward = list(forward)  # (dst, src) buffer names
        self.substep_index = 0

    def forward_buffers(self):
        for dst, src in self.forward:
            self.p.buf[dst] = self.p.buf[src]

    def compute(self):
        p = self.p
        p.sub_dt = p.dt / self.substeps
        for s in range(self.substeps):
            self.substep_index = s
            self.substep()
            if s < self.substeps - 1:
                p.advance_state()
            p.sub_time += p.sub_dt


class SplitOperatorSolver(TensorSolver):
    """src/tensor_solver/SplitOperatorBase.C:39-64."""

    def __init__(self, problem, root, buffer, reciprocal_buffer, linear_reciprocal,
                 nonlinear_reciprocal, substeps=1, history=0, forward=()):
        super().__init__(problem, root, substeps, forward)
        n = len(buffer)
        lin = list(linear_reciprocal) if linear_reciprocal else ["0"] * n
        self.vars = []
        for i in range(n):
            self.vars.append(dict(u=buffer[i], ubar=reciprocal_buffer[i],
                                  L=None if lin[i] == "0" else lin[i], N=nonlinear_reciprocal[i],
                                  Nold=problem.get_old(nonlinear_reciprocal[i], history)))


class AdamsBashforthMoulton(SplitOperatorSolver):
    """src/tensor_solver/AdamsBashforthMoulton.C:45-178."""

    def __init__(self, problem, root, buffer, reciprocal_buffer, linear_reciprocal,
                 nonlinear_reciprocal, substeps=1, predictor_order=2, corrector_order=2,
                 corrector_steps=0, forward=()):
        self.P = predictor_order - 1
        self.C = corrector_order - 1
        self.csteps = corrector_steps
        super().__init__(problem, root, buffer, reciprocal_buffer, linear_reciprocal,
                         nonlinear_reciprocal, substeps, max(self.P, self.C), forward)

    def substep(self):
        p, b = self.p, self.p.buf
        self.root.compute()
        self.forward_buffers()
        dt = p.sub_dt
        dt_changed = p.dt != p.dt_old
        for v in self.vars:
            n_old = len(v["Nold"])
            order = min(0 if (self.substep_index < self.P and dt_changed) else n_old, self.P)
            ubar = b[v["ubar"]] + (dt * AB_BETA[order][0]) * b[v["N"]]
            for i in range(order):
                ubar = ubar + (dt * AB_BETA[order][i + 1]) * v["Nold"][i]
            if v["L"] is not None:
                ubar = ubar / (1.0 - dt * b[v["L"]])
            b[v["u"]] = self.d.ifft(ubar)
        if self.csteps:
            p.sub_time += dt
            ubar_n = [b[v["ubar"]] for v in self.vars]
            N_n = [b[v["N"]] for v in self.vars] if self.C > 0 else None
            for _ in range(self.csteps):
                self.root.compute()
                self.forward_buffers()
                for k, v in enumerate(self.vars):
                    n_old = len(v["Nold"])
                    order = min(1 if (self.substep_index < self.C and dt_changed) else n_old + 1,
                                self.C)
                    if order == 0:
                        continue
                    ubar = ubar_n[k] + (dt * AM_ALPHA[order][0]) * b[v["N"]]
                    ubar = ubar + (dt * AM_ALPHA[order][1]) * N_n[k]
                    for i in range(order - 1):
                        ubar = ubar + (dt * AM_ALPHA[order][i + 2]) * v["Nold"][i]
                    if v["L"] is not None:
                        ubar = ubar / (1.0 - dt * b[v["L"]])
                    b[v["u"]] = self.d.ifft(ubar)
            p.sub_time -= dt


class AdamsBashforthMoultonCoupled(SplitOperatorSolver):
    """src/tensor_solver/AdamsBashforthMoultonCoupled.C:51-273: Adams-Bashforth-Moulton with a dense
    (per wavevector) linear operator, solved with a batched `linalg_solve`.  As coded there:
    * the dense operator is assembled as stack(stack(cols,-1) per row, -1), i.e. the matrix that
      reaches linalg_solve is the TRANSPOSE of the user's L_ij table (:151-164);
    * a missing diagonal entry dereferences a null pointer in the reference; here it is zero;
    * the sub-time is advanced inside substep() (:177) in addition to TensorSolver::computeBuffer;
    * a corrector step of order 0 still solves A u = u_n (:213-217), unlike AdamsBashforthMoulton;
    * the right-hand side is cast to the dtype of the first variable's linear operator (:141,167),
      which is real, so only the REAL part of each spectrum enters the solve and the inverse
      transform (the gold files coupled_*.csv pin exactly this behaviour)."""

    def __init__(self, problem, root, buffer, reciprocal_buffer, linear_reciprocal,
                 nonlinear_reciprocal, substeps=1, predictor_order=2, corrector_order=2,
                 corrector_steps=0, linear_offdiag_rows=(), linear_offdiag_cols=(),
                 linear_offdiag=(), assume_symmetric=False, forward=()):
        self.P = predictor_order - 1
        self.C = corrector_order - 1
        self.csteps = corrector_steps
        self.offdiag = list(zip(linear_offdiag_rows, linear_offdiag_cols, linear_offdiag))
        self.assume_symmetric = assume_symmetric
        super().__init__(problem, root, buffer, reciprocal_buffer, linear_reciprocal,
                         nonlinear_reciprocal, substeps, max(self.P, self.C), forward)

    def _solve(self, rhs_list):
        b, n = self.p.buf, len(self.vars)
        base = b[self.vars[0]["L"]]
        zero = torch.zeros_like(base)
        tab = [[zero] * n for _ in range(n)]
        given = set()
        for i, v in enumerate(self.vars):
            if v["L"] is not None:
                tab[i][i] = b[v["L"]]
        for i, j, name in self.offdiag:
            tab[i][j] = b[name]
            given.add((i, j))
        if self.assume_symmetric:
            for i, j, name in self.offdiag:
                if i != j and tab[j][i] is zero:
                    tab[j][i] = b[name]
        rows = [torch.stack([tab[i][j] for j in range(n)], -1) for i in range(n)]
        Lm = torch.stack(rows, -1)
        A = torch.eye(n, dtype=base.dtype) - self.p.sub_dt * Lm.to(base.dtype)
        # :167 casts the (complex) right-hand side to the dtype of the first linear operator, which
        # is REAL: the imaginary part of every spectrum is dropped before the solve (as coded)
        rhs = torch.stack(rhs_list, -1)
        rhs = rhs.real.to(base.dtype) if (rhs.is_complex() and not base.is_complex()) else rhs.to(base.dtype)
        sol = torch.linalg.solve(A, rhs)
        return torch.unbind(sol, -1)

    def substep(self):
        p, b = self.p, self.p.buf
        self.root.compute()
        self.forward_buffers()
        dt = p.sub_dt
        dt_changed = p.dt != p.dt_old
        rhs = []
        for v in self.vars:
            n_old = len(v["Nold"])
            order = min(0 if (self.substep_index < self.P and dt_changed) else n_old, self.P)
            r = b[v["ubar"]] + (dt * AB_BETA[order][0]) * b[v["N"]]
            for i in range(order):
                r = r + (dt * AB_BETA[order][i + 1]) * v["Nold"][i]
            rhs.append(r)
        for v, ubar in zip(self.vars, self._solve(rhs)):
            b[v["u"]] = self.d.ifft(ubar)
        p.sub_time += dt
        if self.csteps:
            ubar_n = [b[v["ubar"]] for v in self.vars]
            N_n = [b[v["N"]] for v in self.vars] if self.C > 0 else None
            for _ in range(self.csteps):
                self.root.compute()
                self.forward_buffers()
                rhs = []
                for k, v in enumerate(self.vars):
                    n_old = len(v["Nold"])
                    order = min(1 if (self.substep_index < self.C and dt_changed) else n_old + 1,
                                self.C)
                    if order == 0:
                        rhs.append(ubar_n[k])
                        continue
                    r = ubar_n[k] + (dt * AM_ALPHA[order][0]) * b[v["N"]]
                    r = r + (dt * AM_ALPHA[order][1]) * N_n[k]
                    for i in range(order - 1):
                        r = r + (dt * AM_ALPHA[order][i + 2]) * v["Nold"][i]
                    rhs.append(r)
                for v, ubar in zip(self.vars, self._solve(rhs)):
                    b[v["u"]] = self.d.ifft(ubar)


class ReciprocalMatDiffusion(Op):
    """src/tensor_computes/ReciprocalMatDiffusion.C:43-66: divergence of the flux -M grad(mu) for a
    spatially varying mobility, with the smoothed-boundary no-flux term grad(psi)/psi . J; psi and its
    gradients are captured at the first evaluation (always_update_psi = false)."""

    def __init__(self, problem, buffer, chemical_potential, mobility, psi, always_update_psi=False):
        super().__init__(problem, buffer)
        self.mu, self.M, self.psi, self.always = chemical_potential, mobility, psi, always_update_psi
        self.cache = None

    def compute(self):
        d = self.d
        k = [d.kaxis[a] for a in range(3)]
        i = torch.tensor(1j, dtype=torch.complex128)
        if self.cache is None or self.always:
            psi = self.get(self.psi)
            thresh = psi > 0.0
            g = [torch.where(thresh, d.ifft(k[a] * d.fft(psi) * i) / psi, 0.0) for a in range(3)]
            self.cache = (thresh, g)
        thresh, g = self.cache
        psi_M = self.get(self.M) * thresh
        mu = self.get(self.mu)
        J = [psi_M * d.ifft(k[a] * d.fft(mu) * i) for a in range(3)]
        div_J_hat = i * (k[0] * d.fft(J[0]) + k[1] * d.fft(J[1]) + k[2] * d.fft(J[2]))
        no_flux_hat = d.fft(g[0] * J[0] + g[1] * J[1] + g[2] * J[2])
        self.set(div_J_hat + no_flux_hat)


class ReciprocalAllenCahn(Op):
    """src/tensor_computes/ReciprocalAllenCahn.C:39-50: fft(where(psi > 0, -L dF/deta, 0))."""

    def __init__(self, problem, buffer, dF_chem_deta, L, psi, always_update_psi=False):
        super().__init__(problem, buffer)
        self.dF, self.L, self.psi, self.always = dF_chem_deta, L, psi, always_update_psi
        self.thresh = None

    def compute(self):
        if self.thresh is None or self.always:
            self.thresh = self.get(self.psi) > 0.0
        rate = torch.where(self.thresh, -1 * self.get(self.L) * self.get(self.dF), 0.0)
        self.set(self.d.fft(rate))


class DeAliasingTensor(Op):
    """src/tensor_computes/DeAliasingTensor.C:37-62: 2/3-rule (SHARP) or Hou-Li exponential filter."""

    def __init__(self, problem, buffer, method, p=16.0, alpha=36.0):
        super().__init__(problem, buffer)
        self.method, self.pw, self.alpha = method, p, alpha

    def compute(self):
        k = [self.d.kaxis[a] for a in range(3)]
        kmax = [float(torch.max(torch.abs(t))) for t in k]
        if self.method == "SHARP":
            cut = (torch.abs(k[0]) > 2 * kmax[0] / 3) | (torch.abs(k[1]) > 2 * kmax[1] / 3) | (torch.abs(k[2]) > 2 * kmax[2] / 3)
            self.set(torch.where(cut, 0.0, 1.0).to(self.d.dtype))
        else:
            pw = [torch.pow(torch.abs(k[a]) / (kmax[a] if kmax[a] else 1.0), self.pw) for a in range(3)]
            self.set(torch.exp(-self.alpha * (pw[0] + pw[1] + pw[2])))


class SwiftHohenbergLinear(Op):
    """src/tensor_computes/SwiftHohenbergLinear.C:33-36: r - alpha^2 (1 - k^2)^2 (real, reciprocal shape)."""

    def __init__(self, problem, buffer, r=-0.5, alpha=1.0):
        super().__init__(problem, buffer)
        self.r, self.alpha = r, alpha

    def compute(self):
        k2 = self.d.k2
        self.set(self.r - self.alpha * self.alpha * (1.0 - k2) * (1.0 - k2))


class SmoothRectangleCompute(Op):
    """src/tensor_computes/SmoothRectangleCompute.C:60-131: `inside` within the box [x1,x2]x[y1,y2](x[z1,z2]),
    `outside` elsewhere, blended over int_width by a half sine (COS) or tanh(4 d / w) (TANH) of the
    distance d to the nearest face per axis; int_width <= 0 is the sharp indicator."""

    def __init__(self, problem, buffer, x1, x2, y1, y2, z1=0.0, z2=0.0, profile="COS", int_width=0.0,
                 inside=1.0, outside=0.0):
        super().__init__(problem, buffer)
        if int_width < 0.0:
            raise ValueError("Interface width must be a non-negative real number.")
        self.lo, self.hi = (x1, y1, z1), (x2, y2, z2)
        self.profile, self.w, self.inside, self.outside = profile, int_width, inside, outside

    def compute(self):
        d, w = self.d, self.w
        ax = [d.axis[a].reshape(-1) for a in range(3)]
        if w <= 0.0:
            h = [((ax[a] >= self.lo[a]) & (ax[a] <= self.hi[a])) if a < d.dim else torch.ones_like(ax[a], dtype=torch.bool)
                 for a in range(3)]
            cond = torch.logical_and(h[0].reshape(-1, 1, 1), torch.logical_and(h[1].reshape(1, -1, 1), h[2].reshape(1, 1, -1)))
            box = torch.zeros(cond.shape, dtype=d.dtype)
            box[cond] = 1.0
        else:
            dist = [torch.minimum(ax[a] - self.lo[a], self.hi[a] - ax[a]) for a in range(3)]
            if self.profile == "COS":
                m = [dist[a].clamp(-w / 2.0, w / 2.0) if a < d.dim else torch.full_like(ax[a], w / 2.0) for a in range(3)]
                h = [0.5 + 0.5 * torch.sin(math.pi * m[a] / w) for a in range(3)]
            elif self.profile == "TANH":
                far = (None, 10 * w, 10 * w / 2.0)
                m = [dist[a] if a < d.dim else torch.full_like(ax[a], far[a]) for a in range(3)]
                h = [0.5 + 0.5 * torch.tanh(4 * m[a] / w) for a in range(3)]
            else:
                raise ValueError(f"profile {self.profile}")
            box = h[0].reshape(-1, 1, 1) * h[1].reshape(1, -1, 1) * h[2].reshape(1, 1, -1)
        self.set((box * self.inside + (1 - box) * self.outside).reshape(d.shape).to(d.dtype))


class MooseFunctionTensor(Op):
    """src/tensor_computes/MooseFunctionTensor.C:31-72: a MOOSE Function sampled at the cell centres
    i*dx + dx/2 (note: not the linspace axis of DomainAction).  `functions` maps a ParsedFunction
    name to (expression, symbol_names, symbol_values); symbol values are other functions (evaluated
    at the same points, as MOOSE's ParsedFunction does) or numbers.  The FParser grammar of these
    expressions (`:=` bindings, `if`, `^`, sin/cos, pi) is a subset of the Marlin grammar, so the
    oracle's expression evaluator is reused."""

    def __init__(self, problem, buffer, function, functions):
        super().__init__(problem, buffer)
        self.function, self.functions = function, functions

    def _eval(self, name, pts):
        expr, names, values = self.functions[name]
        consts = {"pi": torch.tensor(math.pi, dtype=torch.float64), "e": torch.tensor(math.e, dtype=torch.float64)}
        variables, params = ["x", "y", "z", "t"], list(pts) + [torch.tensor(self.p.time, dtype=torch.float64)]
        for n, v in zip(names, values):
            variables.append(n)
            params.append(self._eval(v, pts) if v in self.functions else torch.tensor(float(v), dtype=torch.float64))
        fn = xp.ParsedTensor(expr, variables, consts)
        fn.compile()
        return fn.eval(params)

    def compute(self):
        d = self.d
        pts = []
        for a in range(3):
            if a < d.dim:
                n, dx = d.n[a], d.dx[a]
                c = torch.arange(n, dtype=torch.float64) * dx + dx / 2.0
                shape = [1] * d.dim
                shape[a] = n
                pts.append(c.reshape(shape))
            else:
                pts.append(torch.tensor(0.0, dtype=torch.float64))
        self.set(self._eval(self.function, pts).expand(d.shape).contiguous().to(d.dtype))


class SecantSolver(SplitOperatorSolver):
    """src/tensor_solver/SecantSolver.C:32-185: implicit Euler solved per wavevector with a secant
    iteration on the reciprocal-space residual R = (N + L u) dt + u_old - u, bootstrapped by one
    semi-implicit step of size dt_epsilon.  `iterations`/`converged` are what
    TensorSolveIterationAdaptiveDT reads (IterativeTensorSolverInterface)."""

    def __init__(self, problem, root, buffer, reciprocal_buffer, linear_reciprocal,
                 nonlinear_reciprocal, substeps=1, max_iterations=30, relative_tolerance=1e-9,
                 absolute_tolerance=1e-9, damping=1.0, dt_epsilon=1e-4, forward=()):
        super().__init__(problem, root, buffer, reciprocal_buffer, linear_reciprocal,
                         nonlinear_reciprocal, substeps, 0, forward)
        self.max_iterations, self.rtol, self.atol = max_iterations, relative_tolerance, absolute_tolerance
        self.damping, self.dt_epsilon = damping, dt_epsilon
        self.iterations, self.converged = 0, True

    def substep(self):
        p, b, d = self.p, self.p.buf, self.d
        dt = p.sub_dt
        n = len(self.vars)
        u_old, Rprev, uprev, R0 = [None] * n, [None] * n, [None] * n, [0.0] * n
        self.root.compute()
        self.forward_buffers()
        for i, v in enumerate(self.vars):
            u, N = b[v["ubar"]], b[v["N"]]
            L = b[v["L"]] if v["L"] is not None else None
            Rprev[i] = (N + L * u) * dt if L is not None else N * dt
            uprev[i] = u
            R0[i] = float(torch.linalg.norm(Rprev[i]))
            u_old[i] = u
            eps = self.dt_epsilon
            b[v["u"]] = d.ifft((u + eps * N) / (1.0 - eps * L)) if L is not None else d.ifft(u + eps * N)
        all_converged = False
        it = 0
        while it < self.max_iterations:
            self.root.compute()
            self.forward_buffers()
            all_converged = True
            for i, v in enumerate(self.vars):
                u, N = b[v["ubar"]], b[v["N"]]
                L = b[v["L"]] if v["L"] is not None else None
                R = ((N + L * u) * dt if L is not None else N * dt) + u_old[i] - u
                dx = u - uprev[i]
                dy = R - Rprev[i]
                du = torch.where(dy != 0, -R * dx / dy, torch.zeros((), dtype=R.dtype))
                uprev[i], Rprev[i] = u, R
                b[v["u"]] = d.ifft(u + du) if self.damping == 1.0 else d.ifft(u + du * self.damping)
                Rnorm = float(torch.linalg.norm(R))
                if math.isnan(Rnorm):
                    all_converged = False
                    it = self.max_iterations
                    break
                all_converged = all_converged and (Rnorm < self.atol or Rnorm / R0[i] < self.rtol)
            if all_converged:
                self.converged = True
                break
            it += 1
        self.iterations = it
        if not all_converged:
            for i, v in enumerate(self.vars):
                b[v["u"]] = d.ifft(u_old[i])
            self.converged = False


class BroydenSolver(SplitOperatorSolver):
    """src/tensor_solver/BroydenSolver.C:34-176: implicit Euler with a per-wavevector Broyden update of the
    inverse Jacobian M ([grid..., n, n] complex, kept across substeps and steps).  As coded there: the
    step is u + 0.5 * sk (the `damping` parameter is not used), `dt_epsilon` is not used, the residual of
    the first iteration omits u_old - u (it is zero), and the rank-one update uses the plain transpose
    (no conjugation) with a |denominator| > 1e-12 guard."""

    def __init__(self, problem, root, buffer, reciprocal_buffer, linear_reciprocal, nonlinear_reciprocal,
                 substeps=1, max_iterations=5, relative_tolerance=1e-9, absolute_tolerance=1e-9,
                 initial_jacobian_guess=1.0, forward=()):
        super().__init__(problem, root, buffer, reciprocal_buffer, linear_reciprocal, nonlinear_reciprocal, substeps, 0, forward)
        self.max_iterations, self.rtol, self.atol = max_iterations, relative_tolerance, absolute_tolerance
        n = len(self.vars)
        self.M = (torch.eye(n, dtype=torch.complex128) * initial_jacobian_guess).expand(list(self.d.rshape) + [n, n])
        self.iterations, self.converged = 0, True

    def _stack(self):
        b = self.p.buf
        u = torch.stack([b[v["ubar"]] for v in self.vars], -1)
        N = torch.stack([b[v["N"]] for v in self.vars], -1)
        L = torch.stack([b[v["L"]] if v["L"] is not None else torch.zeros(self.d.rshape, dtype=self.d.dtype) for v in self.vars], -1)
        return u, N, L

    def substep(self):
        p, d = self.p, self.d
        dt = p.sub_dt
        self.root.compute()
        self.forward_buffers()
        u_old = torch.stack([p.buf[v["ubar"]] for v in self.vars], -1)
        u, N, L = self._stack()
        R = (N + L * u) * dt
        R0 = float(torch.linalg.norm(R))
        it = 0
        while it < self.max_iterations:
            Rnorm = float(torch.linalg.norm(R))
            if math.isnan(Rnorm):
                raise RuntimeError("NAN!")
            if Rnorm < self.atol or Rnorm / R0 < self.rtol:
                self.iterations, self.converged = it, True
                return
            sk = -torch.matmul(self.M, R.unsqueeze(-1))
            skT = sk.squeeze(-1).unsqueeze(-2)
            u_out = torch.unbind(u + sk.squeeze(-1) * 0.5, -1)
            for i, v in enumerate(self.vars):
                p.buf[v["u"]] = d.ifft(u_out[i])
            self.root.compute()
            self.forward_buffers()
            u, N, L = self._stack()
            Rnew = (N + L * u) * dt + u_old - u
            yk = (Rnew - R).unsqueeze(-1)
            denom = torch.matmul(skT, yk)
            self.M = self.M + torch.where(torch.abs(denom) > 1e-12, torch.matmul(sk - torch.matmul(self.M, yk), skT) / denom, 0.0)
            R = Rnew
            it += 1
        self.iterations, self.converged = it, False


class ForwardEulerSolver(TensorSolver):
    """src/tensor_solver/ForwardEulerSolver.C:29-38 (variables may be empty: mechanics)."""

    def __init__(self, problem, root, buffer=(), reciprocal_buffer=(), time_derivative_reciprocal=(),
                 substeps=1, forward=()):
        super().__init__(problem, root, substeps, forward)
        self.vars = list(zip(buffer, reciprocal_buffer, time_derivative_reciprocal))

    def substep(self):
        b = self.p.buf
        self.root.compute()
        self.forward_buffers()
        for u, ubar, nbar in self.vars:
            b[u] = self.d.ifft(b[ubar] + self.p.sub_dt * b[nbar])


class ETDRK4Solver(SplitOperatorSolver):
    """src/tensor_solver/ETDRK4Solver.C:29-115 (as coded, not textbook Cox-Matthews)."""

    def __init__(self, problem, root, buffer, reciprocal_buffer, linear_reciprocal,
                 nonlinear_reciprocal, substeps=1, forward=()):
        super().__init__(problem, root, buffer, reciprocal_buffer, linear_reciprocal,
                         nonlinear_reciprocal, substeps, 1, forward)

    def _nonlinear(self, stage):
        b = self.p.buf
        for v, ub in zip(self.vars, stage):
            b[v["u"]] = self.d.ifft(ub)
        self.root.compute()
        self.forward_buffers()
        return [b[v["N"]] for v in self.vars]

    def substep(self):
        b, dt = self.p.buf, self.p.sub_dt
        self.root.compute()
        self.forward_buffers()
        un = [b[v["ubar"]] for v in self.vars]
        N1 = [b[v["N"]] for v in self.vars]
        lin = [b[v["L"]] if v["L"] is not None else torch.zeros_like(un[i])
               for i, v in enumerate(self.vars)]
        eL, eL2, ph1, ph2, ph3, ub = [], [], [], [], [], []
        for i in range(len(self.vars)):
            Ldt = lin[i] * dt
            e = torch.exp(Ldt)
            eL.append(e)
            eL2.append(torch.exp(Ldt / 2.0))
            den = Ldt * Ldt * Ldt
            p1 = dt * (-4.0 - 3.0 * Ldt + e * (4.0 - Ldt)) / den
            p2 = dt * (2.0 + Ldt + e * (-2.0 + Ldt)) / den
            p3 = dt * (-4.0 - 3.0 * Ldt - Ldt * Ldt + e * (4.0 - Ldt)) / den
            zm = Ldt == 0.0
            if bool(zm.any()):
                dtt = torch.full_like(Ldt, dt)
                p1 = torch.where(zm, dtt, p1)
                p2 = torch.where(zm, dtt * dtt / 2.0, p2)
                p3 = torch.where(zm, dtt * dtt / 6.0, p3)
            ph1.append(p1)
            ph2.append(p2)
            ph3.append(p3)
            ub.append(eL2[i] * un[i] + 0.5 * dt * N1[i])
        N2 = self._nonlinear(ub)
        uc = [eL2[i] * un[i] + 0.5 * dt * N2[i] for i in range(len(self.vars))]
        N3 = self._nonlinear(uc)
        ud = [eL[i] * un[i] + dt * N3[i] for i in range(len(self.vars))]
        N4 = self._nonlinear(ud)
        for i, v in enumerate(self.vars):
            ubar = eL[i] * un[i] + ph1[i] * N1[i] + 2.0 * ph2[i] * (N2[i] + N3[i]) + ph3[i] * N4[i]
            b[v["u"]] = self.d.ifft(ubar)


class FFTSemiImplicit(Op):
    """Legacy per-buffer integrator used as an operator
    (src/tensor_timeintegrators/FFTSemiImplicit.C:43-62)."""

    def __init__(self, problem, buffer, reciprocal_buffer, linear_reciprocal, nonlinear_reciprocal,
                 history_size=1):
        super().__init__(problem, buffer)
        self.ubar, self.L, self.N = reciprocal_buffer, linear_reciprocal, nonlinear_reciprocal
        self.ubar_old = problem.get_old(reciprocal_buffer, history_size)
        self.N_old = problem.get_old(nonlinear_reciprocal, history_size)

    def compute(self):
        b, dt = self.p.buf, self.p.sub_dt
        n_old = min(len(self.ubar_old), len(self.N_old))
        if n_old == 0:
            ubar = (b[self.ubar] + dt * b[self.N]) / (1.0 - dt * b[self.L])
        else:
            ubar = (b[self.ubar] + dt / 2.0 * (3.0 * b[self.N] - self.N_old[0])) / \
                   (1.0 - dt * b[self.L])
        self.set(self.d.ifft(ubar))


# =============================================================================== mechanics
def trans2(A):  # src/utils/MarlinUtils.C:147-187
    return torch.einsum("...ij->...ji", A)


def ddot42(A, B):
    return torch.einsum("...ijkl,...lk->...ij", A, B)


def ddot44(A, B):
    return torch.einsum("...ijkl,...lkmn->...ijmn", A, B)


def dot22(A, B):
    return torch.einsum("...ij,...jk->...ik", A, B)


def dot24(A, B):
    return torch.einsum("...ij,...jkmn->...ikmn", A, B)


def dot42(A, B):
    return torch.einsum("...ijkl,...lm->...ijkm", A, B)


def dyad22(A, B):
    return torch.einsum("...ij,...kl->...ijkl", A, B)


def _unsq0(t, n):
    for _ in range(n):
        t = t.unsqueeze(0)
    return t


def conjugate_gradient(A, b, x0, tol, maxiter):
    """include/utils/MarlinUtils.h:57-123 (identity preconditioner)."""
    x = x0.clone() if x0 is not None else torch.zeros_like(b)
    b_norm = float(torch.norm(b))
    if b_norm == 0.0:
        return x, 0, 0.0
    if not maxiter:
        maxiter = b.numel()
    r = b - A(x)
    p = r.clone()
    rz_old = float(torch.sum(r * r))
    res = 0.0
    for k in range(maxiter):
        Ap = A(p)
        alpha = rz_old / float(torch.sum(p * Ap))
        x = x + alpha * p
        r = r - alpha * Ap
        res = float(torch.norm(r))
        if res <= tol * b_norm:
            return x, k + 1, res
        rz_new = float(torch.sum(r * r))
        beta = rz_new / rz_old
        p = r + beta * p
        rz_old = rz_new
    return x, maxiter, res


class RankTwoIdentity(Op):
    """src/tensor_computes/RankTwoIdentity.C:29-33."""

    def compute(self):
        dm = self.d.dim
        self.set(torch.eye(dm, dtype=self.d.dtype).expand(self.d.value_shape([dm, dm])))


class MacroscopicShearTensor(Op):
    """test/src/tensor_computes/MacroscopicShearTensor.C:31-41.  `_time` there is the
    TensorProblem sub-time reference (TensorOperatorBase), see SURVEY a12."""

    def __init__(self, problem, buffer, F="F"):
        super().__init__(problem, buffer)
        self.F = F

    def compute(self):
        avg = self.d.average(self.get(self.F))
        shear = torch.eye(self.d.dim, dtype=self.d.dtype)
        shear[0, 1] = shear[0, 1] + self.p.sub_time
        self.set(shear - avg)


class PhaseMechanicsTest(Op):
    """test/src/tensor_computes/PhaseMechanicsTest.C:31-50."""

    def compute(self):
        u = torch.zeros(self.d.shape, dtype=self.d.dtype)
        s = 30 if self.d.dim == 2 else 9
        if self.d.dim == 3:
            u[-s:, :s, -s:] = 1.0
        else:
            u[-s:, :s] = 1.0
        self.set(u)


class HyperElasticIsotropic(Op):
    """src/tensor_computes/HyperElasticIsotropic.C:24-52."""

    def __init__(self, problem, buffer, F, K, mu, tangent_operator="dstressdstrain"):
        super().__init__(problem, buffer)
        self.F, self.K, self.mu, self.K4 = F, K, mu, tangent_operator
        dm, dt = self.d.dim, self.d.dtype
        ti = torch.eye(dm, dtype=dt)
        self.tI = _unsq0(ti, dm)
        self.tI4 = _unsq0(torch.einsum("il,jk", ti, ti), dm)
        self.tI4rt = _unsq0(torch.einsum("ik,jl", ti, ti), dm)
        self.tI4s = (self.tI4 + self.tI4rt) / 2.0
        self.tII = dyad22(self.tI, self.tI)

    def compute(self):
        b = self.p.buf
        F = b[self.F]
        K = b[self.K].reshape(self.d.value_shape([1, 1, 1, 1]))
        mu = b[self.mu].reshape(self.d.value_shape([1, 1, 1, 1]))
        C4 = K * self.tII + 2.0 * mu * (self.tI4s - 1.0 / 3.0 * self.tII)
        S = ddot42(C4, 0.5 * (dot22(trans2(F), F) - self.tI))
        self.set(dot22(F, S))
        b[self.K4] = dot24(S, self.tI4) + ddot44(ddot44(self.tI4rt, dot42(dot24(F, C4), trans2(F))),
                                                 self.tI4rt)


class FFTMechanics(Op):
    """src/tensor_computes/FFTMechanics.C:48-85 (Ghat4), :96-163 (Newton-CG)."""

    def __init__(self, problem, buffer, constitutive_model, K, mu, F="F", stress="stress",
                 tangent_operator="dstressdstrain", applied_macroscopic_strain=None, l_tol=1e-2,
                 l_max_its=None, nl_rel_tol=1e-5, nl_abs_tol=1e-8, nl_max_its=100):
        super().__init__(problem, buffer)
        self.cm = constitutive_model
        self.F, self.P, self.K4 = F, stress, tangent_operator
        self.applied = applied_macroscopic_strain
        self.l_tol, self.l_max_its = l_tol, l_max_its or self.d.ncells
        self.nl_rel_tol, self.nl_abs_tol, self.nl_max_its = nl_rel_tol, nl_abs_tol, nl_max_its
        dm = self.d.dim
        q = self.d.kgrid
        Q = self.d.k2.unsqueeze(-1).unsqueeze(-1)
        M = torch.where(Q == 0, 0.0, q.unsqueeze(-2) * q.unsqueeze(-1) / Q)
        M = M.unsqueeze(-3).unsqueeze(-1)
        ti = torch.eye(dm, dtype=self.d.dtype)
        delta_im = ti.unsqueeze(1).unsqueeze(1).expand(dm, dm, dm, dm)
        self.Ghat4 = (M * delta_im).to(torch.complex128)
        self.r2 = self.d.value_shape([dm, dm])
        self.cg_iterations = []   # per CG solve
        self.newton_iterations = 0

    def compute(self):
        b, d = self.p.buf, self.d

        def G(A2):
            return d.ifft(ddot42(self.Ghat4, d.fft(A2))).reshape(-1)

        def K_dF(dFm):
            return trans2(ddot42(b[self.K4], trans2(dFm.reshape(self.r2))))

        def G_K_dF(dFm):
            return G(K_dF(dFm))

        b[self.buffer] = b[self.F]
        self.cm.compute()
        if self.applied is not None:
            rhs = -G_K_dF(b[self.applied].expand(self.r2))
            b[self.buffer] = b[self.buffer] + b[self.applied].expand(self.r2)
        else:
            rhs = -G_K_dF(torch.zeros_like(b[self.F]))
        Fn = float(torch.linalg.norm(b[self.buffer]))
        iiter = 0
        dFm = torch.zeros_like(rhs)
        self.cg_iterations = []
        while True:
            dFm, its, _ = conjugate_gradient(G_K_dF, rhs, dFm, self.l_tol, self.l_max_its)
            self.cg_iterations.append(its)
            b[self.buffer] = b[self.buffer] + dFm.reshape(self.r2)
            self.cm.compute()
            rhs = -G(b[self.P])
            anorm = float(torch.linalg.norm(dFm))
            rnorm = anorm / Fn
            if (rnorm < self.nl_rel_tol or anorm < self.nl_abs_tol) and iiter > 0:
                break
            iiter += 1
            if iiter > self.nl_max_its:
                raise RuntimeError("Exceeded the maximum number of nonlinear iterations without "
                                   "converging.")
        self.newton_iterations = iiter + 1


def green_project_lowmem(d, A2):
    """G(A) = ifft(ddot42(Ghat4, fft(A))) of FFTMechanics.C:74-84,104-105 WITHOUT materialising Ghat4 (81 complex
    fields: 11 GB at 256^3).  With Ghat4_ijlm = delta_im q_j q_l / |q|^2 (0 at q = 0) and ddot42(A, B)_ij = A_ijkl B_lk:
        ddot42(Ghat4, Ahat)_ij = sum_kl delta_il q_j q_k / Q * Ahat_lk = q_j * (sum_k q_k Ahat_ik) / Q.
    Same arithmetic, contracted row by row; pinned to the materialised form by tests/test_oracle_mech_lowmem.py."""
    dm = d.dim
    q = d.kgrid                                   # [..., dm]
    Q = d.k2
    Ah = d.fft(A2)                                # [..., dm, dm] complex
    out = torch.empty_like(Ah)
    for i in range(dm):
        s = torch.zeros_like(Ah[..., 0, 0])
        for k in range(dm):
            s = s + q[..., k] * Ah[..., i, k]
        s = torch.where(Q == 0, torch.zeros_like(s), s / torch.where(Q == 0, torch.ones_like(Q), Q))
        for j in range(dm):
            out[..., i, j] = q[..., j] * s
    return d.ifft(out)


def tangent_apply_chunked(hyper_cls, d, F, K, mu, x, chunk=32):
    """K_dF(x) = trans2(ddot42(K4, trans2(x))) (FFTMechanics.C:107-110) with K4 from HyperElasticIsotropic.C:42-52,
    evaluated on x-chunks of `chunk` planes so that the 81-component tangent (10.9 GB at 256^3) never exists at once.
    The constitutive law is pointwise in space, so this is the same arithmetic on sub-blocks; returns (P, K_dF)."""
    P = torch.empty_like(F)
    out = torch.empty_like(x)
    for x0 in range(0, F.shape[0], chunk):
        sl = slice(x0, min(x0 + chunk, F.shape[0]))
        n_sub = list(F.shape[:d.dim])
        n_sub[0] = sl.stop - sl.start
        sub = Domain(d.dim, n_sub + [1] * (3 - d.dim), d.min, d.max, d.dtype)
        pr = Problem(sub)
        pr.buf.update(F=F[sl], K=K[sl], mu=mu[sl])
        h = hyper_cls(pr, "stress", "F", "K", "mu")
        h.compute()
        P[sl] = pr.buf["stress"]
        out[sl] = trans2(ddot42(pr.buf["dstressdstrain"], trans2(x[sl])))
        del pr, h
    return P, out


class ComputeVonMisesStress(Op):
    """src/tensor_computes/ComputeVonMisesStress.C:30-67 (3-D and 2-D forms as coded)."""

    def __init__(self, problem, buffer, stress="stress"):
        super().__init__(problem, buffer)
        self.stress = stress

    def compute(self):
        s = self.p.buf.get(self.stress)
        if s is None:
            return
        if self.d.dim == 3:
            t = (s[..., 0, 0] - s[..., 1, 1]) ** 2 + (s[..., 1, 1] - s[..., 2, 2]) ** 2 + (s[..., 2, 2] - s[..., 0, 0]) ** 2
            t = t + 6 * (s[..., 0, 1] ** 2 + s[..., 1, 2] ** 2 + s[..., 2, 0] ** 2)
        else:
            t = (s[..., 0, 0] - s[..., 1, 1]) ** 2 + 6 * s[..., 0, 1] ** 2
        self.set(torch.sqrt(0.5 * t))


class ComputeDisplacements(Op):
    """src/tensor_computes/ComputeDisplacements.C:53-107: displacement field of a periodic deformation
    gradient: u = (<F> - I) X + ifft( Hbar q (-i) / |q|^2 ), Hbar = fft(F - <F>), resampled onto the
    (n+1)^dim nodal grid with (bi/tri)linear interpolation, align_corners = true."""

    def __init__(self, problem, buffer, F):
        super().__init__(problem, buffer)
        self.F = F

    def compute(self):
        d = self.d
        F = self.p.buf.get(self.F)
        if F is None:
            return
        dm = d.dim
        I3 = torch.eye(dm, dtype=F.dtype)
        Fbox = d.average(F)
        Hbar = d.fft(F - Fbox)
        q = d.kgrid * (-1j)
        Q = d.k2
        numer = torch.einsum("...ij,...j->...i", Hbar, q.to(Hbar.dtype))
        denom = Q.unsqueeze(-1)
        u_periodic_bar = torch.where(denom == 0, 0.0, numer / denom)
        X = torch.stack([d.axis[a].expand(d.shape) for a in range(dm)], -1) if dm > 1 else d.axis[0]
        u_aff = torch.einsum("ij,...j->...i", Fbox - I3, X)
        u_periodic = torch.fft.irfftn(u_periodic_bar, s=d.shape, dim=list(range(dm)))
        shape = [n + 1 for n in d.shape]
        mode = {1: "linear", 2: "bilinear", 3: "trilinear"}[dm]
        u = torch.nn.functional.interpolate((u_aff + u_periodic).movedim(-1, 0).unsqueeze(1), size=shape, mode=mode,
                                            align_corners=True).squeeze(1).movedim(0, -1)
        self.set(u)


class FFTQuasistaticElasticity(Op):
    """src/tensor_computes/FFTQuasistaticElasticity.C:45-104: homogeneous isotropic quasistatic elasticity with
    a volumetric eigenstrain e0*c, solved per wavevector (3x3 system, batched linalg_solve); 3-D only (the
    code indexes {0,0,0}).  The wave vector is 2 pi i times the reciprocal axis, which already carries 2 pi
    (as coded)."""

    def __init__(self, problem, displacements, cbar, mu, lam, e0):
        super().__init__(problem)
        self.disp, self.cbar, self.mu, self.lam, self.e0 = list(displacements), cbar, mu, lam, e0

    def compute(self):
        d, mu, lam = self.d, self.mu, self.lam
        tpi = torch.tensor(2j * math.pi, dtype=torch.complex128)
        kx, ky, kz = tpi * d.kaxis[0], tpi * d.kaxis[1], tpi * d.kaxis[2]
        ul = 2.0 * mu + lam
        Axx = ul * kx * kx + mu * ky * ky + mu * kz * kz
        s = Axx.shape
        Axy = ((lam + mu) * kx * ky).expand(s)
        Axz = ((lam + mu) * kx * kz).expand(s)
        Ayy = ul * ky * ky + mu * kx * kx + mu * kz * kz
        Ayz = ((lam + mu) * ky * kz).expand(s)
        Azz = ul * kz * kz + mu * kx * kx + mu * ky * ky
        Axx[0, 0, 0] = 1.0
        Ayy[0, 0, 0] = 1.0
        Azz[0, 0, 0] = 1.0
        e = 2.0 * self.e0 * self.get(self.cbar) * (3.0 * lam + mu)
        e[0, 0, 0] = 0.0
        b = torch.stack([kx * e, ky * e, kz * e], -1)
        A = torch.stack([torch.stack([Axx, Axy, Axz], -1), torch.stack([Axy, Ayy, Ayz], -1), torch.stack([Axz, Ayz, Azz], -1)], -1)
        x = torch.linalg.solve(A, b)
        for i, name in enumerate(self.disp):
            self.p.buf[name] = d.ifft(x[..., i])


class FFTElasticChemicalPotential(Op):
    """src/tensor_computes/FFTElasticChemicalPotential.C:46-61 (reciprocal-space elastic contribution to the
    chemical potential)."""

    def __init__(self, problem, buffer, displacements, cbar, mu, lam, e0):
        super().__init__(problem, buffer)
        self.disp, self.cbar, self.mu, self.lam, self.e0 = list(displacements), cbar, mu, lam, e0

    def compute(self):
        d = self.d
        tpi = torch.tensor(2j * math.pi, dtype=torch.complex128)
        kx, ky, kz = tpi * d.kaxis[0], tpi * d.kaxis[1], tpi * d.kaxis[2]
        ux, uy, uz = [d.fft(self.get(n)) for n in self.disp]
        cbar = self.get(self.cbar)
        self.set(-self.e0 * (self.e0 * (9.0 * self.lam * cbar + self.mu * 6.0 * cbar) -
                             (2.0 * self.mu + 3.0 * self.lam) * (kx * ux + ky * uy + kz * uz)))


# =============================================================================== postprocessors
def pp_integral(problem, name):
    """src/postprocessors/TensorIntegralPostprocessor.C:28-38: average * domain volume."""
    return pp_average(problem, name) * problem.domain.volume


def pp_average(problem, name):
    """src/postprocessors/TensorAveragePostprocessor.C:34-47."""
    return float(problem.buf[name].sum()) / float(problem.domain.ncells)


def pp_extreme(problem, name, kind):
    t = problem.buf[name]
    return float(t.min()) if kind == "MIN" else float(t.max())


def pp_interface_velocity(problem, name, old):
    """src/postprocessors/TensorInterfaceVelocityPostprocessor.C:41-65 (`old` = getBufferOld(name, 1);
    the gradient threshold is the literal 1e-3 of the code, not the parameter)."""
    if not old or old[0] is None:
        return 0.0
    d = problem.domain
    u = problem.buf[name]
    du = (u - old[0]) / problem.dt
    vsq = None
    for a in range(d.dim):
        grad = d.ifft(d.fft(u) * d.kaxis[a] * 1j)
        v = torch.where(torch.abs(grad) > 1e-3, du / grad, 0.0)
        vsq = v * v if vsq is None else vsq + v * v
    return math.sqrt(float(torch.max(vsq)))


def vpp_histogram(problem, name, mn, mx, bins):
    """src/vectorpostprocessors/TensorHistogram.C:31-84: (bin centres, counts); edges = linspace(min, max, bins+1),
    at::native::histogramdd on the CPU (the reference falls back to the CPU as well: no CUDA histogramdd)."""
    edges = torch.linspace(mn, mx, bins + 1, dtype=problem.domain.dtype)
    u = problem.buf[name].expand(problem.domain.shape).reshape(-1, 1)
    hist = torch.histogramdd(u, [edges])[0]
    step = (mx - mn) / bins
    return [mn + step / 2.0 + step * i for i in range(bins)], hist.tolist()

"""Test harness: drives the fused split plan with the reference's time-stepping protocol
(TensorSolver::computeBuffer substep loop, src/tensor_solver/TensorSolver.C:93-110;
AdamsBashforthMoulton order selection, src/tensor_solver/AdamsBashforthMoulton.C:75-91;
TensorProblem::advanceState early return for step 1, src/problems/TensorProblem.C:455-456).
The C++ host objects implement the same protocol; this Python mirror exists so the parity
tests can call the C ABI directly."""
from marlin_b200.capi import AB_BETA


class SplitDriver:
    def __init__(self, plan, c, substeps, predictor_order=2):
        self.plan, self.c = plan, c
        self.substeps = substeps
        self.P = predictor_order - 1
        self.t_step = 0
        self.dt_old = None
        self.stored = 0

    def _advance(self):
        if self.t_step <= 1:
            return
        self.stored = self.plan.advance_state()

    def step(self, dt):
        self.t_step += 1
        self._advance()
        dt_changed = self.dt_old is not None and dt != self.dt_old
        sub_dt = dt / self.substeps
        for s in range(self.substeps):
            order = min(0 if (s < self.P and dt_changed) else self.stored, self.P)
            self.plan.substep(self.c, sub_dt, AB_BETA[order], order)
            if s < self.substeps - 1:
                self._advance()
        self.dt_old = dt

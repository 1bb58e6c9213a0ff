"""CPU backend for lqcd_b200.rhmc.RHMCAction built on the oracle (TEST INFRASTRUCTURE: lives under tests/, never imported by the
product package).  Backend protocol: new_like(v), copy(v), axpy(y, a, x) [y += a x], scale(v, a), dot(a, b) -> complex,
apply(mode, x) -> y (mode in {"D", "Ddag"}), shifted_solve(b, shifts) -> list of x_j."""
import numpy as np


class OracleBackend:
    """CPU oracle backend (tests only)."""

    def __init__(self, orc, op, kind, U, eps=1e-24, maxsteps=5000):
        self.orc, self.op, self.kind, self.U, self.eps, self.maxsteps = orc, op, kind, U, eps, maxsteps
        self.last_iters = 0

    def new_like(self, v):
        return np.zeros_like(v)

    def copy(self, v):
        return v.copy()

    def axpy(self, y, a, x):
        y += a * x

    def scale(self, v, a):
        v *= a

    def dot(self, a, b):
        return complex(np.vdot(a, b))

    def apply(self, mode, x):
        return self.orc.apply(self.op, self.kind, {"D": self.orc.D, "Ddag": self.orc.DDAG}[mode], self.U, x)

    def shifted_solve(self, b, shifts):
        r = self.orc.mscg(self.op, self.kind, self.U, b, shifts, eps=self.eps, maxsteps=self.maxsteps)
        assert r["converged"]
        self.last_iters = r["iters"]
        return r["xs"]

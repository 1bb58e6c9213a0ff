"""
Rational-HMC plumbing on top of the multi-shift CG (lqcd_multishift_cg / upstream ``shiftedcg``).

Reference behaviour being mirrored: for staggered fermions
``FermiAction(D, parameters_action)`` (src/system/universe.jl:106-110,138) switches to RHMC when
``parameters_action["Nf"]`` is neither 4 nor 8 ("other than Nf=4, 8 ... RHMC is automatically used",
README.md:132; test/test_Nf2.toml, test/test_Nf3.toml): D^dag D describes 8 tastes, so

    heat bath   phi = (D^dag D)^{+Nf/16} xi                 (gauss_sampling_in_action! / sample_pseudofermions!)
    action      S_f = phi^dag (D^dag D)^{-Nf/8} phi          (evaluate_FermiAction)
    MD force    sum_j alpha_j * force(X_j, Y_j),  X_j = (D^dag D + beta_j)^-1 phi, Y_j = D X_j   (calc_UdSfdU!)

with x^p ~ alpha_0 + sum_j alpha_j / (x + beta_j).  Upstream obtains the coefficients from AlgRemez_jll
(Manifest_old.toml:11); that binary is not available here, so the partial fractions are computed by a
relative-error least-squares fit started from the Stieltjes-integral quadrature of x^p (same functional form,
same use; the coefficients differ from Remez' by construction, the approximation error is asserted instead).

Everything above the solver is written against a tiny *backend* (vector ops + shifted solve): the product backend is
B200Backend below; the test-suite plugs in a CPU backend of its own (tests/oracle_backend.py) to check the same composition
against dense matrix functions without a GPU.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass

import numpy as np


@dataclass
class RationalApprox:
    power: float
    alpha0: float
    alpha: np.ndarray          # residues
    beta: np.ndarray           # shifts (> 0, ascending)
    lo: float
    hi: float
    max_rel_err: float

    def __call__(self, x):
        x = np.asarray(x, dtype=float)
        return self.alpha0 + (self.alpha[None, :] / (x[..., None] + self.beta[None, :])).sum(-1)


def rational_approx(power: float, order: int, lo: float, hi: float, npts: int = 400) -> RationalApprox:
    """x^power ~ alpha0 + sum_j alpha_j/(x + beta_j) on [lo, hi], -1 < power < 1, power != 0 (cached per argument set)."""
    return _rational_approx(float(power), int(order), float(lo), float(hi), int(npts))


@functools.lru_cache(maxsize=64)
def _rational_approx(power: float, order: int, lo: float, hi: float, npts: int) -> RationalApprox:
    assert -1.0 < power < 1.0 and power != 0.0 and 0 < lo < hi and order >= 2
    if power < 0:
        c0, a, b = _fit_negative_power(power, order, lo, hi, npts, fix_c0=False)
    else:
        # x^p = x * x^(p-1): fit the negative power q = p - 1 WITHOUT constant term (its Stieltjes form has none), then
        # x * sum_j a_j/(x + b_j) = sum_j a_j - sum_j a_j b_j/(x + b_j).  Same relative error as the x^q fit, and the
        # Levenberg-Marquardt fit of a negative power converges in ~2 s where the direct positive-power fit needed ~30 s.
        _, a, b = _fit_negative_power(power - 1.0, order, lo, hi, npts, fix_c0=True)
        c0, a = float(a.sum()), -a * b
    idx = np.argsort(b)
    ra = RationalApprox(power, float(c0), a[idx].copy(), b[idx].copy(), lo, hi, 0.0)
    dense = np.exp(np.linspace(np.log(lo), np.log(hi), 4 * npts))
    ra.max_rel_err = float(np.abs(ra(dense) / dense ** power - 1.0).max())
    return ra


def _fit_negative_power(power, order, lo, hi, npts, fix_c0):
    """relative-error least squares (Lawson reweighted towards minimax) of x^power, -1 < power < 0, started from the
    quadrature of x^-g = (sin(pi g)/pi) int_0^inf t^-g/(x+t) dt on a log grid"""
    from scipy.optimize import least_squares

    xs = np.exp(np.linspace(np.log(lo), np.log(hi), npts))
    target = xs ** power
    g = -power
    pad = 2.5
    s = np.linspace(np.log(lo) - pad, np.log(hi) + pad, order)
    ds = s[1] - s[0]
    t = np.exp(s)
    al = np.sin(np.pi * g) / np.pi * t ** (1.0 - g) * ds          # weights of 1/(x+t_j)
    k0 = 0 if fix_c0 else 1

    def unpack(p):
        return (0.0 if fix_c0 else p[0]), p[k0:k0 + order], np.exp(p[k0 + order:])

    wts = np.ones_like(xs)

    def rel(p):
        c0, a, b = unpack(p)
        return (c0 + (a[None, :] / (xs[:, None] + b[None, :])).sum(1)) / target - 1.0

    def resid(p):
        return wts * rel(p)

    p = np.concatenate([[] if fix_c0 else [0.0], al, np.log(t)])
    for _ in range(5):           # Lawson-type reweighting: pushes the least-squares fit towards the minimax (Remez) one
        p = least_squares(resid, p, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=800).x
        e = np.abs(rel(p))
        wts = wts * (0.5 + e / e.mean())
        wts /= wts.mean()
    c0, a, b = unpack(p)
    return float(c0), np.array(a), np.array(b)


# ---------------------------------------------------------------------------------------------------
# backend protocol: new_like(v), copy(v), axpy(y, a, x) [y += a x], scale(v, a), dot(a, b) -> complex,
#                   apply(mode, x) -> y  (mode in {"D", "Ddag"}),  shifted_solve(b, shifts) -> list of x_j
# ---------------------------------------------------------------------------------------------------
class B200Backend:
    """liblqcd_b200 backend: device-resident FermionField handles, lqcd_multishift_cg for the shifted systems."""

    def __init__(self, D):
        from . import api
        self.api, self.D = api, D
        self.last_iters = 0

    def new_like(self, v):
        f = self.api.similar(v)
        self.api.clear_fermion_(f)
        return f

    def copy(self, v):
        f = self.api.similar(v)
        self.api.substitute_fermion_(f, v)
        return f

    def axpy(self, y, a, x):
        self.api.add_(y, a, x)

    def scale(self, v, a):
        a = complex(a)
        v.ctx.call("lqcd_blas_scale", a.real, a.imag, v.h)

    def dot(self, a, b):
        return self.api.dot(a, b)

    def apply(self, mode, x):
        y = self.api.similar(x)
        self.api.mul_(y, self.D if mode == "D" else self.api.adjoint(self.D), x)
        return y

    def shifted_solve(self, b, shifts):
        ys = [self.api.similar(b) for _ in shifts]
        info = self.api.shiftedcg_(ys, self.D, b, shifts)
        self.last_iters = info["iters"]
        return ys


class RHMCAction:
    """Staggered pseudofermion action det(D^dag D)^{Nf/8} through rational approximations."""

    def __init__(self, backend, Nf: int, lambda_min: float, lambda_max: float, order: int = 12, tolerance: float = 1e-6):
        if Nf in (4, 8):
            raise ValueError("Nf = 4, 8 use the plain HMC action (no rational approximation)")
        self.be, self.Nf = backend, Nf
        self._range = (order, float(lambda_min), float(lambda_max))
        self.tolerance = float(tolerance)

    def _checked(self, ra: RationalApprox) -> RationalApprox:
        """a fit that missed its tolerance (too low an order for the spectral range, a least-squares run that did not converge)
        would silently bias the action: refuse it"""
        if not (ra.max_rel_err <= self.tolerance):
            raise ValueError(f"rational approximation of x^{ra.power:+.4f} on [{ra.lo:.4g}, {ra.hi:.4g}] with {len(ra.beta)} poles has "
                             f"max relative error {ra.max_rel_err:.3e} > tolerance {self.tolerance:.1e}: raise rational_order or "
                             f"narrow rational_lambda_min / rational_lambda_max")
        return ra

    # the fits are computed on first use and cached per (power, order, range): the positive-power (heat-bath) fit is the
    # slow one (tens of seconds) and is not needed by evaluate / force_terms
    @property
    def r_heatbath(self) -> RationalApprox:
        return self._checked(rational_approx(+self.Nf / 16.0, *self._range))

    @property
    def r_action(self) -> RationalApprox:
        return self._checked(rational_approx(-self.Nf / 8.0, *self._range))

    def _apply_rational(self, ra: RationalApprox, b):
        """alpha0 b + sum_j alpha_j (D^dag D + beta_j)^-1 b with ONE multi-shift solve."""
        be = self.be
        xs = be.shifted_solve(b, list(ra.beta))
        out = be.copy(b)
        be.scale(out, ra.alpha0)
        for a, x in zip(ra.alpha, xs):
            be.axpy(out, a, x)
        return out, xs

    def sample_pseudofermions(self, xi):
        """phi = (D^dag D)^{Nf/16} xi."""
        return self._apply_rational(self.r_heatbath, xi)[0]

    def evaluate(self, phi) -> float:
        """S_f = phi^dag (D^dag D)^{-Nf/8} phi."""
        y, _ = self._apply_rational(self.r_action, phi)
        return self.be.dot(phi, y).real

    def force_terms(self, phi):
        """[(alpha_j, X_j, Y_j)]: the MD force is sum_j alpha_j F(X_j, Y_j) with F the bilinear force of
        lqcd_fermion_force / oracle.force (d/dU of phi^dag (D^dag D + beta_j)^-1 phi)."""
        xs = self.be.shifted_solve(phi, list(self.r_action.beta))
        return [(float(a), x, self.be.apply("D", x)) for a, x in zip(self.r_action.alpha, xs)]

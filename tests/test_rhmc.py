"""RHMC plumbing (rational approximations + multi-shift composition): CPU tests on the oracle backend; the same
code runs on the B200 backend in the gpu-marked test (only verified primitives underneath: shiftedcg_, dot, add_)."""
import numpy as np
import pytest

import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "latticeqcd.jl_b200"), str(ROOT / "tests")]
from lqcd_b200 import rhmc
from oracle_backend import OracleBackend
from oracle import oracle as orc

DIMS = (4, 4, 4, 4)


@pytest.mark.parametrize("power", [-0.25, -0.375, 0.125, 0.1875, -0.5])
def test_rational_approx_accuracy(power):
    ra = rhmc.rational_approx(power, 12, 0.25, 6.0)
    assert ra.max_rel_err < 1e-6, ra.max_rel_err
    assert (ra.beta > 0).all()
    x = np.array([0.25, 1.0, 3.3, 6.0])
    assert np.allclose(ra(x), x ** power, rtol=1e-6)


def _dense_DdagD(op, U):
    n = 256 * 3
    cols = []
    for i in range(n):
        e = np.zeros(n, dtype=complex); e[i] = 1
        cols.append(orc.apply(op, orc.STAGGERED, orc.DDAGD, U, e.reshape(4, 4, 4, 4, 3)).ravel())
    return np.array(cols).T


@pytest.mark.parametrize("Nf", [2, 3])
def test_rhmc_against_dense_spectrum(golden_dir, Nf):
    """heat bath phi = (D^dag D)^{Nf/16} xi and action phi^dag (D^dag D)^{-Nf/8} phi = xi^dag xi, checked against the exact
    matrix functions from a dense eigendecomposition on the staggered fixture (test/test_Nf2.toml, test_Nf3.toml physics)."""
    U = np.load(golden_dir / "staggered_nf2_4444.npy")
    op = orc.make_op(DIMS, mass=0.5)
    A = _dense_DdagD(op, U)
    w, V = np.linalg.eigh((A + A.conj().T) / 2)
    assert w.min() > 0.25 - 1e-9
    be = OracleBackend(orc, op, orc.STAGGERED, U)
    act = rhmc.RHMCAction(be, Nf, 0.9 * w.min(), 1.1 * w.max(), order=12)
    xi = orc.gaussian_field(DIMS, orc.STAGGERED, seed=77)
    phi = act.sample_pseudofermions(xi)
    exact = (V * w ** (Nf / 16.0)) @ (V.conj().T @ xi.ravel())
    assert np.abs(phi.ravel() - exact).max() < 1e-6 * np.abs(exact).max()
    S = act.evaluate(phi)
    assert abs(S - np.vdot(xi, xi).real) < 1e-5 * S
    terms = act.force_terms(phi)
    assert len(terms) == 12 and all(np.isfinite(a) for a, _, _ in terms)
    # sum_j alpha_j X_j + alpha0 phi reproduces (D^dag D)^{-Nf/8} phi
    acc = act.r_action.alpha0 * phi
    for a, X, Y in terms:
        acc = acc + a * X
    exact2 = (V * w ** (-Nf / 8.0)) @ (V.conj().T @ phi.ravel())
    assert np.abs(acc.ravel() - exact2).max() < 1e-6 * np.abs(exact2).max()


@pytest.mark.gpu
def test_rhmc_on_b200(golden_dir):
    import lqcd_b200 as q
    Uh = np.load(golden_dir / "staggered_nf2_4444.npy")
    U = q.gaugefields_from_array(Uh)
    x = q.Initialize_pseudofermion_fields(U[0], "staggered")
    D = q.Dirac_operator(U, x, {"Dirac_operator": "staggered", "mass": 0.5, "eps_CG": 1e-24, "MaxCGstep": 5000,
                                "boundarycondition": [1, 1, 1, -1]})
    act = rhmc.RHMCAction(rhmc.B200Backend(D), 2, 0.22, 6.0, order=12)
    xi = q.similar(x)
    xi_h = orc.gaussian_field(DIMS, orc.STAGGERED, seed=77)
    xi.from_host(xi_h)
    phi = act.sample_pseudofermions(xi)
    S = act.evaluate(phi)
    assert abs(S - np.vdot(xi_h, xi_h).real) < 1e-5 * S
    # same numbers as the oracle backend with the same rational functions
    be = OracleBackend(orc, orc.make_op(DIMS, mass=0.5), orc.STAGGERED, Uh)
    act_cpu = rhmc.RHMCAction(be, 2, 0.22, 6.0, order=12)
    phi_cpu = act_cpu.sample_pseudofermions(xi_h)
    assert np.abs(phi.to_host() - phi_cpu).max() < 1e-9 * np.abs(phi_cpu).max()


def _expm_antiherm(A):
    w, V = np.linalg.eigh(1j * A)                   # A anti-Hermitian -> iA Hermitian
    return (V * np.exp(-1j * w)) @ V.conj().T


def test_rhmc_force_finite_difference(golden_dir):
    """the RHMC force sum_j alpha_j force(X_j, Y_j) is the derivative of S = phi^dag r(D^dag D) phi with the SAME rational
    function r: dS/d eps = -2 Re tr[A F_mu(n)] for U_mu(n) -> exp(eps A) U_mu(n)  (the identity of SURVEY.md App. C.6
    applied term by term; pins the sign / weight convention that lqcd_fermion_force_xy accumulates on the device)."""
    U = np.load(golden_dir / "staggered_nf2_4444.npy")
    op = orc.make_op(DIMS, mass=0.5)
    phi = orc.gaussian_field(DIMS, orc.STAGGERED, seed=91)

    def make(Ux):
        return rhmc.RHMCAction(OracleBackend(orc, op, orc.STAGGERED, Ux), 2, 0.22, 17.0, order=12)

    act = make(U)
    F = np.zeros_like(U)
    for a, X, Y in act.force_terms(phi):
        F += a * orc.force(op, orc.STAGGERED, U, X, Y)
    rng = np.random.default_rng(5)
    for (mu, t, z, y, x) in [(0, 0, 0, 0, 0), (3, 3, 1, 2, 0), (2, 1, 1, 1, 1)]:
        H = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))
        A = (H - H.conj().T) / 2
        A -= np.trace(A) / 3 * np.eye(3)
        want = -2 * np.real(np.trace(A @ F[mu, t, z, y, x].T))
        h, vals = 1e-5, []
        for sgn in (+1, -1):
            U2 = U.copy()
            U2[mu, t, z, y, x] = (_expm_antiherm(sgn * h * A) @ U[mu, t, z, y, x].T).T
            vals.append(make(np.ascontiguousarray(U2)).evaluate(phi))
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - want) < 1e-6 * max(1.0, abs(want)), (fd, want)


@pytest.mark.gpu
def test_fermi_action_nf_dispatch_on_b200(golden_dir):
    """FermiAction(D, {"Nf": ..}) as the wrapper builds it (universe.jl:106-110,138): Nf = 8 plain, Nf = 4 even-site
    pseudofermions, Nf = 2 RHMC; heat bath, action and MD force of each against the oracle."""
    import lqcd_b200 as q
    Uh = np.load(golden_dir / "staggered_nf2_4444.npy")
    U = q.gaugefields_from_array(Uh)
    x = q.Initialize_pseudofermion_fields(U[0], "staggered")
    params = {"Dirac_operator": "staggered", "mass": 0.5, "eps_CG": 1e-24, "MaxCGstep": 5000, "boundarycondition": [1, 1, 1, -1]}
    D = q.Dirac_operator(U, x, params)
    op = orc.make_op(DIMS, mass=0.5)
    t, z, y, xx = np.meshgrid(*[np.arange(4)] * 4, indexing="ij")
    odd = ((t + z + y + xx) & 1) == 1

    # ---- Nf = 4: even-site pseudofermions --------------------------------------------------------------------------------
    fa4 = q.FermiAction(D, {"Nf": 4})
    assert isinstance(fa4, q.FermiActionB200) and fa4.even_only
    xi, eta = q.similar(x), q.similar(x)
    q.gauss_sampling_in_action_(xi, U, fa4, seed=5)
    xi_h = xi.to_host()
    assert np.abs(xi_h[odd]).min() > 0.0 and np.abs(xi_h[~odd]).min() > 0.0       # xi lives on all sites (api.FermiActionB200)
    q.sample_pseudofermions_(eta, U, fa4, xi)
    eta_h = eta.to_host()
    want = orc.apply(op, orc.STAGGERED, orc.DDAG, Uh, xi_h)
    want[odd] = 0.0
    assert np.abs(eta_h - want).max() < 1e-13
    S = q.evaluate_FermiAction(fa4, U, eta)
    assert abs(S - np.vdot(xi_h, xi_h).real) < 1e-9 * S       # Sfold = dot(xi, xi) (standardHMC.jl:54) IS the initial action
    ref = orc.cg(op, orc.STAGGERED, Uh, eta_h, eps=1e-24)
    assert np.abs(ref["x"][odd]).max() < 1e-12            # D^dag D does not couple the parities: X stays on even sites
    assert abs(S - np.vdot(eta_h, ref["x"]).real) < 1e-9 * abs(S)
    F = np.zeros_like(Uh)
    q.calc_UdSfdU_(F, fa4, U, eta)
    Fr = orc.force(op, orc.STAGGERED, Uh, ref["x"], orc.apply(op, orc.STAGGERED, orc.D, Uh, ref["x"]))
    assert np.abs(F - Fr).max() < 1e-9 * np.abs(Fr).max()

    # ---- Nf = 8: all sites -------------------------------------------------------------------------------------------------
    fa8 = q.FermiAction(D, {"Nf": 8})
    assert isinstance(fa8, q.FermiActionB200) and not fa8.even_only

    # ---- Nf = 2: RHMC, force accumulated on the device ---------------------------------------------------------------------
    fa2 = q.FermiAction(D, {"Nf": 2, "rational_lambda_min": 0.22, "rational_lambda_max": 17.0})
    assert isinstance(fa2, q.RHMCFermiAction)
    q.gauss_sampling_in_action_(xi, U, fa2, seed=6)
    xi_h = xi.to_host()
    q.sample_pseudofermions_(eta, U, fa2, xi)
    S = q.evaluate_FermiAction(fa2, U, eta)
    assert abs(S - np.vdot(xi_h, xi_h).real) < 1e-5 * S
    act_cpu = rhmc.RHMCAction(OracleBackend(orc, op, orc.STAGGERED, Uh), 2, 0.22, 17.0, order=12)
    eta_h = eta.to_host()
    Fr = np.zeros_like(Uh)
    for a, X, Y in act_cpu.force_terms(eta_h):
        Fr += a * orc.force(op, orc.STAGGERED, Uh, X, Y)
    q.calc_UdSfdU_(F, fa2, U, eta)
    assert np.abs(F - Fr).max() < 1e-8 * np.abs(Fr).max()


def test_rational_fit_that_misses_its_tolerance_is_refused():
    """a user-supplied order / spectral range the fit cannot cover must raise instead of silently biasing the action (the
    tolerance is parameters_action["rational_tolerance"], default 1e-6)"""
    from lqcd_b200 import rhmc
    bad = rhmc.RHMCAction(None, 2, 1e-3, 20.0, order=3)
    with pytest.raises(ValueError, match="max relative error"):
        bad.r_action
    good = rhmc.RHMCAction(None, 2, 0.22, 17.0, order=12)
    assert good.r_action.max_rel_err < 1e-7

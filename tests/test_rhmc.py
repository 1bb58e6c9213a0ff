"""RHMC plumbing (rational approximations + multi-shift composition): CPU tests on the oracle backend; the same
code runs on the B200 backend in the gpu-marked test (only verified primitives underneath: shiftedcg_, dot, add_)."""
import numpy as np
import pytest

import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "latticeqcd.jl_b200"))
from lqcd_b200 import rhmc
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
    be = rhmc.OracleBackend(orc, op, orc.STAGGERED, U)
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
    be = rhmc.OracleBackend(orc, orc.make_op(DIMS, mass=0.5), orc.STAGGERED, Uh)
    act_cpu = rhmc.RHMCAction(be, 2, 0.22, 6.0, order=12)
    phi_cpu = act_cpu.sample_pseudofermions(xi_h)
    assert np.abs(phi.to_host() - phi_cpu).max() < 1e-9 * np.abs(phi_cpu).max()

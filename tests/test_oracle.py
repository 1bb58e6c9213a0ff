"""
Pins the CPU oracle (oracle/lqcd_oracle.c).  The reference holds no vector-level golden data for the
Dslash / CG path (SURVEY.md 8c "parity unpinned"), so the oracle is pinned by
  (1) the in-tree gauge fixtures + their plaquettes,
  (2) an independent numpy restatement (tests/np_ref.py),
  (3) basis-independent known answers: free-field plane waves, gamma5-hermiticity, gauge covariance,
  (4) solver self-consistency: true residuals, the iteration counts SURVEY.md App. D recorded on the
      Wilson fixture, multi-shift == single-shift, finite-difference force.
All CPU, a few seconds.
"""
import json

import numpy as np
import pytest

from oracle import oracle as orc
import np_ref

DIMS = (4, 4, 4, 4)
KAPPA = 0.141139          # test/test_wilson.toml / parameter_structs.jl:127


@pytest.fixture(scope="module")
def Uw(golden_dir):
    return np.load(golden_dir / "wilson_4444.npy")


@pytest.fixture(scope="module")
def Us(golden_dir):
    return np.load(golden_dir / "staggered_4444.npy")


def test_fixture_plaquettes(golden_dir):
    meta = json.loads((golden_dir / "fixtures.json").read_text())
    survey = {"wilson_4444": 0.565800226845, "staggered_4444": 0.575584039475,
              "staggered_nf2_4444": 0.566501729368, "quenched_su3_4444": 0.568215750149}
    for name, want in survey.items():
        U = np.load(golden_dir / f"{name}.npy")
        p = orc.plaquette(meta[name]["dims"], U)
        assert abs(p - want) < 1e-11
        M = np_ref.links_mat(U)
        assert np.abs(M @ np.conj(np.swapaxes(M, -1, -2)) - np.eye(3)).max() < 1e-9   # unitary
        assert np.abs(np.linalg.det(M) - 1).max() < 1e-9


def test_gamma_algebra():
    op = orc.make_op(DIMS, r=1.0)
    rp, rm = orc.gamma_tables(op)
    g = (rp - rm) / 2
    for mu in range(4):
        assert np.allclose(g[mu], np_ref.G[mu])
        assert np.allclose(g[mu], g[mu].conj().T)
        for nu in range(4):
            anti = g[mu] @ g[nu] + g[nu] @ g[mu]
            assert np.allclose(anti, 2 * np.eye(4) * (mu == nu))
        assert np.allclose(np_ref.G5 @ g[mu] + g[mu] @ np_ref.G5, 0)


@pytest.mark.parametrize("dagger", [False, True])
def test_wilson_matches_numpy(Uw, dagger):
    op = orc.make_op(DIMS, kappa=KAPPA)
    x = orc.gaussian_field(DIMS, orc.WILSON, seed=5)
    y = orc.apply(op, orc.WILSON, orc.DDAG if dagger else orc.D, Uw, x)
    yn = np_ref.wilson(Uw, x, KAPPA, dagger=dagger)
    assert np.abs(y - yn).max() < 1e-14 * np.abs(yn).max() * 10


@pytest.mark.parametrize("dagger", [False, True])
def test_staggered_matches_numpy(Us, dagger):
    op = orc.make_op(DIMS, mass=0.5)
    x = orc.gaussian_field(DIMS, orc.STAGGERED, seed=6)
    y = orc.apply(op, orc.STAGGERED, orc.DDAG if dagger else orc.D, Us, x)
    yn = np_ref.staggered(Us, x, 0.5, dagger=dagger)
    assert np.abs(y - yn).max() < 1e-14


def test_wilson_plane_wave():
    """cold links: M psi_p = [1 - 2 kappa sum_mu (r cos p_mu - i gamma_mu sin p_mu)] psi_p, antiperiodic T."""
    dims = (4, 6, 4, 8)
    NX, NY, NZ, NT = dims
    U = np.zeros((4, NT, NZ, NY, NX, 3, 3), dtype=complex)
    U[..., 0, 0] = U[..., 1, 1] = U[..., 2, 2] = 1
    op = orc.make_op(dims, kappa=0.12, r=1.0)
    n = (1, 2, 0, 3)
    p = [2 * np.pi * n[0] / NX, 2 * np.pi * n[1] / NY, 2 * np.pi * n[2] / NZ, 2 * np.pi * (n[3] + 0.5) / NT]
    t, z, y, x = np.meshgrid(np.arange(NT), np.arange(NZ), np.arange(NY), np.arange(NX), indexing="ij")
    phase = np.exp(1j * (p[0] * x + p[1] * y + p[2] * z + p[3] * t))
    rng = np.random.default_rng(1)
    w = rng.standard_normal((4, 3)) + 1j * rng.standard_normal((4, 3))
    psi = np.ascontiguousarray(w[:, None, None, None, None, :] * phase[None, ..., None])
    got = orc.apply(op, orc.WILSON, orc.D, U, psi)
    Mp = np.eye(4, dtype=complex)
    for mu in range(4):
        Mp -= 2 * 0.12 * (np.cos(p[mu]) * np.eye(4) - 1j * np.sin(p[mu]) * np_ref.G[mu])
    want = np.einsum("sr,rtzyxc->stzyxc", Mp, psi)
    assert np.abs(got - want).max() < 1e-13


def test_staggered_plane_wave_norm():
    """cold links: D^dag D chi_p = (m^2 + sum sin^2 p_mu) chi_p for any momentum (eta phases square away)."""
    dims = (4, 4, 6, 8)
    NX, NY, NZ, NT = dims
    U = np.zeros((4, NT, NZ, NY, NX, 3, 3), dtype=complex)
    U[..., 0, 0] = U[..., 1, 1] = U[..., 2, 2] = 1
    m = 0.37
    op = orc.make_op(dims, mass=m)
    n = (1, 3, 2, 1)
    p = [2 * np.pi * n[0] / NX, 2 * np.pi * n[1] / NY, 2 * np.pi * n[2] / NZ, 2 * np.pi * (n[3] + 0.5) / NT]
    t, z, y, x = np.meshgrid(np.arange(NT), np.arange(NZ), np.arange(NY), np.arange(NX), indexing="ij")
    phase = np.exp(1j * (p[0] * x + p[1] * y + p[2] * z + p[3] * t))
    chi = np.ascontiguousarray(phase[..., None] * np.array([1.0, -0.5j, 0.25]))
    got = orc.apply(op, orc.STAGGERED, orc.DDAGD, U, chi)
    lam = m * m + sum(np.sin(pp) ** 2 for pp in p)
    assert np.abs(got - lam * chi).max() < 1e-13


def test_gamma5_hermiticity_and_adjoint(Uw, Us):
    op = orc.make_op(DIMS, kappa=KAPPA, mass=0.5)
    a = orc.gaussian_field(DIMS, orc.WILSON, seed=1)
    b = orc.gaussian_field(DIMS, orc.WILSON, seed=2)
    Db = orc.apply(op, orc.WILSON, orc.D, Uw, b)
    Dda = orc.apply(op, orc.WILSON, orc.DDAG, Uw, a)
    assert abs(np.vdot(a, Db) - np.vdot(Dda, b)) < 1e-11
    g5 = np.array([1, 1, -1, -1])[:, None, None, None, None, None]
    lhs = g5 * orc.apply(op, orc.WILSON, orc.D, Uw, g5 * a)
    assert np.abs(lhs - Dda).max() < 1e-14
    # staggered: <a, D b> = <D^dag a, b>; hop part anti-hermitian; D^dag D keeps parity
    sa = orc.gaussian_field(DIMS, orc.STAGGERED, seed=3)
    sb = orc.gaussian_field(DIMS, orc.STAGGERED, seed=4)
    assert abs(np.vdot(sa, orc.apply(op, orc.STAGGERED, orc.D, Us, sb))
               - np.vdot(orc.apply(op, orc.STAGGERED, orc.DDAG, Us, sa), sb)) < 1e-12
    t, z, y, x = np.meshgrid(*[np.arange(4)] * 4, indexing="ij")
    even = ((x + y + z + t) % 2 == 0)[..., None]
    se = sa * even
    out = orc.apply(op, orc.STAGGERED, orc.DDAGD, Us, np.ascontiguousarray(se))
    assert np.abs(out * (~even)).max() == 0.0


def test_gauge_covariance(Uw):
    rng = np.random.default_rng(9)
    a = rng.standard_normal((4, 4, 4, 4, 3, 3)) + 1j * rng.standard_normal((4, 4, 4, 4, 3, 3))
    q, _ = np.linalg.qr(a)
    q = q / (np.linalg.det(q) ** (1 / 3))[..., None, None]
    Ug = np_ref.gauge_transform(Uw, q)
    op = orc.make_op(DIMS, kappa=KAPPA)
    psi = orc.gaussian_field(DIMS, orc.WILSON, seed=8)
    gpsi = np.ascontiguousarray(np.einsum("tzyxab,stzyxb->stzyxa", q, psi))
    lhs = orc.apply(op, orc.WILSON, orc.D, Ug, gpsi)
    rhs = np.einsum("tzyxab,stzyxb->stzyxa", q, orc.apply(op, orc.WILSON, orc.D, Uw, psi))
    assert np.abs(lhs - rhs).max() < 1e-13


def test_cg_iteration_anchors(Uw):
    """SURVEY.md App. D (independent numpy probe on the same fixture): point source, eps=1e-19 absolute:
    CG(D^dag D) 96 iterations, CGNR('bicg') 94; kappa=0.12: 63 / 62.  Spectrum-driven -> allow +-2."""
    b = orc.point_source(DIMS, orc.WILSON, 0, 0)
    for kappa, want_cg, want_nr in [(KAPPA, 96, 94), (0.12, 63, 62)]:
        op = orc.make_op(DIMS, kappa=kappa)
        r = orc.cg(op, orc.WILSON, Uw, b)
        assert r["converged"] and abs(r["iters"] - want_cg) <= 2, r["iters"]
        true_r = b - orc.apply(op, orc.WILSON, orc.DDAGD, Uw, r["x"])
        assert np.vdot(true_r, true_r).real < 1e-18
        r2 = orc.cgnr(op, orc.WILSON, Uw, b)
        assert r2["converged"] and abs(r2["iters"] - want_nr) <= 2, r2["iters"]
        true_r = b - orc.apply(op, orc.WILSON, orc.D, Uw, r2["x"])
        assert np.vdot(true_r, true_r).real < 1e-18
        # CGNR solves D x = b;  D^dag D x' = D^dag b must agree
        bb = orc.apply(op, orc.WILSON, orc.DDAG, Uw, b)
        r3 = orc.cg(op, orc.WILSON, Uw, bb)
        assert np.abs(r3["x"] - r2["x"]).max() < 1e-8


def test_staggered_cg_anchor(Us):
    """App. D: staggered fixture, m=0.5, Gaussian source: 53 (full) / 52 (even) iterations -> window."""
    op = orc.make_op(DIMS, mass=0.5)
    b = orc.gaussian_field(DIMS, orc.STAGGERED, seed=112)
    r = orc.cg(op, orc.STAGGERED, Us, b)
    assert r["converged"] and 45 <= r["iters"] <= 60
    true_r = b - orc.apply(op, orc.STAGGERED, orc.DDAGD, Us, r["x"])
    assert np.vdot(true_r, true_r).real < 1e-18


def test_bicgstab(Uw):
    op = orc.make_op(DIMS, kappa=0.12)
    b = orc.gaussian_field(DIMS, orc.WILSON, seed=21)
    r = orc.bicgstab(op, orc.WILSON, Uw, b, eps=1e-20)
    assert r["converged"]
    true_r = b - orc.apply(op, orc.WILSON, orc.D, Uw, r["x"])
    assert np.vdot(true_r, true_r).real < 1e-18


def test_multishift_equals_single(Us):
    op = orc.make_op(DIMS, mass=0.1)
    b = orc.gaussian_field(DIMS, orc.STAGGERED, seed=31)
    shifts = [0.0, 0.05, 0.4, 2.0]
    ms = orc.mscg(op, orc.STAGGERED, Us, b, shifts, eps=1e-22)
    assert ms["converged"]
    for s, x in zip(shifts, ms["xs"]):
        op_s = orc.make_op(DIMS, mass=np.sqrt(0.1 ** 2 + s))        # D^dag D + s = (m^2+s) - hop^2
        single = orc.cg(op_s, orc.STAGGERED, Us, b, eps=1e-22)
        assert np.abs(single["x"] - x).max() < 1e-9


def _expm_antiherm(A):
    w, v = np.linalg.eigh(1j * A)          # A = -i H
    return (v * np.exp(-1j * w)[None, :]) @ v.conj().T


@pytest.mark.parametrize("kind", [orc.WILSON, orc.STAGGERED])
def test_force_finite_difference(Uw, Us, kind):
    """dS_f/d eps = -2 Re tr[A UdSfdU_mu(n)] for U_mu(n) -> exp(eps A) U_mu(n), S_f = phi^dag (D^dag D)^-1 phi
    (SURVEY.md App. C.6 self-consistency identity)."""
    U = Uw if kind == orc.WILSON else Us
    op = orc.make_op(DIMS, kappa=0.12, mass=0.5)
    phi = orc.gaussian_field(DIMS, kind, seed=41)

    def action(Ux):
        r = orc.cg(op, kind, Ux, phi, eps=1e-24)
        return np.vdot(phi, r["x"]).real, r["x"]

    S0, X = action(U)
    Y = orc.apply(op, kind, orc.D, U, X)
    F = orc.force(op, kind, U, X, Y)
    rng = np.random.default_rng(3)
    for (mu, t, z, y, x) in [(0, 0, 0, 0, 0), (3, 3, 1, 2, 0), (1, 2, 3, 3, 3), (3, 0, 2, 1, 3), (2, 1, 1, 1, 1)]:
        H = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))
        A = (H - H.conj().T) / 2
        A -= np.trace(A) / 3 * np.eye(3)
        Fm = F[mu, t, z, y, x].T                      # host [b,a] -> matrix [a,b]
        want = -2 * np.real(np.trace(A @ Fm))
        h = 1e-5
        vals = []
        for sgn in (+1, -1):
            U2 = U.copy()
            M = U[mu, t, z, y, x].T
            U2[mu, t, z, y, x] = (_expm_antiherm(sgn * h * A) @ M).T
            vals.append(action(np.ascontiguousarray(U2))[0])
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - want) < 1e-6 * max(1.0, abs(want)), (fd, want)


def test_oracle_regression(golden_dir, Uw, Us):
    """Freezes the oracle (tests/golden/oracle_regression.json, written by make_oracle_regression.py from the oracle
    itself -- not reference data): the checker of every GPU parity test must not drift unnoticed."""
    ref = json.loads((golden_dir / "oracle_regression.json").read_text())
    opw, ops = orc.make_op(DIMS, kappa=KAPPA), orc.make_op(DIMS, mass=0.5)
    pw, ps = orc.gaussian_field(DIMS, orc.WILSON, seed=112), orc.gaussian_field(DIMS, orc.STAGGERED, seed=112)

    def sig(a):
        a = np.asarray(a).ravel()
        w = np.cos(np.arange(a.size) * 0.37) + 1j * np.sin(np.arange(a.size) * 0.11)
        return np.array([np.vdot(a, a).real, np.vdot(w, a).real, np.vdot(w, a).imag])

    def close(name, a, tol=1e-10):
        want = np.array([ref[name]["norm2"], ref[name]["probe_re"], ref[name]["probe_im"]])
        assert np.allclose(sig(a), want, rtol=tol, atol=tol * abs(want).max()), name

    close("wilson_D", orc.apply(opw, orc.WILSON, orc.D, Uw, pw))
    close("wilson_Ddag", orc.apply(opw, orc.WILSON, orc.DDAG, Uw, pw))
    close("stag_D", orc.apply(ops, orc.STAGGERED, orc.D, Us, ps))
    r = orc.cg(opw, orc.WILSON, Uw, pw)
    assert r["iters"] == ref["wilson_cg"]["iters"]
    close("wilson_cg", r["x"], 1e-8)
    r = orc.cgnr(opw, orc.WILSON, Uw, orc.point_source(DIMS, orc.WILSON, 0, 0))
    assert r["iters"] == ref["wilson_cgnr_point"]["iters"]
    r = orc.cg(ops, orc.STAGGERED, Us, ps)
    assert r["iters"] == ref["stag_cg"]["iters"]

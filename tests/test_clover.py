"""
Wilson-clover oracle checks (CPU).  The clover term is NEW capability (not reachable from run_LQCD at the
surveyed commit), so it is pinned by construction-independent properties and an independent numpy restatement:
leaf geometry on a constant-field-strength background, Hermiticity / gamma5-Hermiticity, gauge covariance, dense numpy.
"""
import numpy as np
import pytest

import np_ref
from oracle import oracle as orc

DIMS = (4, 4, 4, 4)
KAPPA, CSW = 0.125, 1.5612      # csw: reference default, src/system/parameter_structs.jl:125


def fields(seed=21, dims=DIMS, eps=None):
    U = orc.random_su3(dims, seed=seed, eps=eps)
    psi = orc.gaussian_field(dims, orc.WILSON, seed=seed + 1)
    return U, psi


def test_sigma_is_chirality_block_diagonal_and_hermitian():
    for mu, nu in np_ref.PLANES:
        s = np_ref.sigma(mu, nu)
        assert np.abs(s - s.conj().T).max() < 1e-15
        assert np.abs(s[:2, 2:]).max() == 0 and np.abs(s[2:, :2]).max() == 0
        assert np.abs(np_ref.G5 @ s - s @ np_ref.G5).max() == 0


def test_fmunu_matches_numpy_and_is_antihermitian_traceless():
    U, _ = fields()
    op = orc.make_op(DIMS, kappa=KAPPA, csw=CSW)
    clov, F = orc.clover_build(op, U, want_f=True)
    Fo = np.swapaxes(F.reshape(4, 4, 4, 4, 6, 3, 3), -1, -2)               # [t,z,y,x,plane,a,b]
    Fn = np.moveaxis(np_ref.fmunu(U), 0, 4)
    assert np.abs(Fo - Fn).max() < 1e-14
    assert np.abs(Fo + np.conj(np.swapaxes(Fo, -1, -2))).max() < 1e-15
    assert np.abs(np.trace(Fo, axis1=-2, axis2=-1)).max() < 1e-15
    A = np.swapaxes(clov, -1, -2)                                           # [V, blk, i, j]
    assert np.abs(A - np.conj(np.swapaxes(A, -1, -2))).max() < 1e-15        # Hermitian 6x6 blocks


def test_constant_field_strength_background():
    """Abelian background with uniform plaquette exp(i theta T) in the (x,y) plane: every leaf equals the plaquette,
    so F^_xy = i sin(theta T) (made traceless) and all other planes vanish."""
    X, Y, Z, T = dims = (4, 6, 2, 2)
    theta = 2 * np.pi / (X * Y)
    Tg = np.array([1.0, 1.0, -2.0])
    M = np.zeros((4, T, Z, Y, X, 3, 3), dtype=complex)
    M[..., :, :] = np.eye(3)
    ys, xs = np.arange(Y), np.arange(X)
    for c in range(3):
        M[0, :, :, :, :, c, c] = np.exp(-1j * theta * Tg[c] * ys)[None, None, :, None]
        M[1, :, :, Y - 1, :, c, c] = np.exp(1j * theta * Tg[c] * Y * xs)[None, None, :]
    U = np.ascontiguousarray(np.swapaxes(M, -1, -2))
    assert abs(orc.plaquette(dims, U) - (5 + np.cos(theta * Tg).sum() / 3) / 6) < 1e-14
    op = orc.make_op(dims, kappa=KAPPA, csw=CSW)
    _, F = orc.clover_build(op, U, want_f=True)
    want = 1j * np.sin(theta * Tg)
    want = want - want.sum() / 3
    for p in range(6):
        for a in range(3):
            for b in range(3):
                w = want[a] if (p == 0 and a == b) else 0.0
                assert np.abs(F[:, p, b, a] - w).max() < 1e-14, (p, a, b)


@pytest.mark.parametrize("dagger", [False, True])
def test_apply_matches_numpy(dagger):
    U, psi = fields()
    op = orc.make_op(DIMS, kappa=KAPPA, csw=CSW)
    orc.clover_build(op, U)
    got = orc.apply(op, orc.WILSON, orc.DDAG if dagger else orc.D, U, psi)
    want = np_ref.wilson_clover(U, psi, KAPPA, CSW, dagger=dagger)
    assert np.abs(got - want).max() < 1e-13
    plain = orc.apply(orc.make_op(DIMS, kappa=KAPPA), orc.WILSON, orc.D, U, psi)
    assert np.abs(got - plain).max() > 1e-3            # the term is really there


def test_adjoint_and_gamma5_hermiticity():
    U, a = fields(seed=5)
    b = orc.gaussian_field(DIMS, orc.WILSON, seed=77)
    op = orc.make_op(DIMS, kappa=KAPPA, csw=CSW)
    orc.clover_build(op, U)
    Mb = orc.apply(op, orc.WILSON, orc.D, U, b)
    Mda = orc.apply(op, orc.WILSON, orc.DDAG, U, a)
    assert abs(np.vdot(a, Mb) - np.vdot(Mda, b)) < 1e-11
    g5 = np.array([1, 1, -1, -1])[:, None, None, None, None, None]
    g5Mg5a = g5 * orc.apply(op, orc.WILSON, orc.D, U, g5 * a)
    assert np.abs(g5Mg5a - Mda).max() < 1e-13


def test_gauge_covariance():
    U, psi = fields(seed=9)
    rng = np.random.default_rng(3)
    h = rng.standard_normal((4, 4, 4, 4, 3, 3)) + 1j * rng.standard_normal((4, 4, 4, 4, 3, 3))
    q, _ = np.linalg.qr(h)
    g = q / np.linalg.det(q)[..., None, None] ** (1 / 3)
    Ug = np_ref.gauge_transform(U, g)
    gpsi = np.einsum("tzyxab,stzyxb->stzyxa", g, psi)
    op = orc.make_op(DIMS, kappa=KAPPA, csw=CSW)
    orc.clover_build(op, U)
    lhs_src = orc.apply(op, orc.WILSON, orc.D, U, psi)
    op2 = orc.make_op(DIMS, kappa=KAPPA, csw=CSW)
    orc.clover_build(op2, Ug)
    rhs = orc.apply(op2, orc.WILSON, orc.D, Ug, np.ascontiguousarray(gpsi))
    lhs = np.einsum("tzyxab,stzyxb->stzyxa", g, lhs_src)
    assert np.abs(lhs - rhs).max() < 1e-13


def test_cg_true_residual_with_clover():
    U, b = fields(seed=31, eps=0.4)
    op = orc.make_op(DIMS, kappa=0.12, csw=CSW)
    orc.clover_build(op, U)
    res = orc.cg(op, orc.WILSON, U, b, eps=1e-20)
    assert res["converged"]
    r = b - orc.apply(op, orc.WILSON, orc.DDAGD, U, res["x"])
    assert np.vdot(r, r).real < 1e-18


def test_device_packing_and_apply_emulation():
    """Mirrors, in numpy, the packed storage written by clover_build_kernel (csrc/clover.cu) and the arithmetic of
    clover_apply (csrc/wilson_kernel.cuh): e = 0..2 diagonal pairs, e = 3 + i(i-1)/2 + j -> A_ij (i > j);
    y_i += A_ij x_j, y_j += conj(A_ij) x_i.  Checks the index algebra against the dense oracle blocks."""
    U, psi = fields(seed=41)
    op = orc.make_op(DIMS, kappa=KAPPA, csw=CSW)
    clov = orc.clover_build(op, U)                         # [V, blk, j, i]
    A = np.swapaxes(clov, -1, -2)                          # [V, blk, i, j]
    V = A.shape[0]
    packed = np.zeros((V, 2, 18), dtype=complex)
    for b in range(2):
        for h in range(3):
            packed[:, b, h] = A[:, b, 2 * h, 2 * h].real + 1j * A[:, b, 2 * h + 1, 2 * h + 1].real
        for i in range(1, 6):
            for j in range(i):
                packed[:, b, 3 + i * (i - 1) // 2 + j] = A[:, b, i, j]
    x = psi.reshape(4, V, 3)                                # [alpha, site, c]
    out = np.zeros_like(x)
    for b in range(2):
        xv = [x[2 * b + i // 3, :, i % 3] for i in range(6)]
        yv = [None] * 6
        for h in range(3):
            d = packed[:, b, h]
            yv[2 * h] = d.real * xv[2 * h]
            yv[2 * h + 1] = d.imag * xv[2 * h + 1]
        for i in range(1, 6):
            for j in range(i):
                c = packed[:, b, 3 + i * (i - 1) // 2 + j]
                yv[i] = yv[i] + c * xv[j]
                yv[j] = yv[j] + np.conj(c) * xv[i]
        for i in range(6):
            out[2 * b + i // 3, :, i % 3] = yv[i]
    want = psi + np_ref.clover_term(U, psi, KAPPA, CSW)
    assert np.abs(out.reshape(psi.shape) - want).max() < 1e-13


def _expm_antiherm(A):
    w, V = np.linalg.eigh(1j * A)
    return (V * np.exp(-1j * w)) @ V.conj().T


def test_clover_force_finite_difference():
    """Wilson-clover pseudofermion force (hopping part + clover-term part, oracle): dS_f/d eps = -2 Re tr[A F_mu(n)] for
    U_mu(n) -> exp(eps A) U_mu(n) with S_f = phi^dag (M^dag M)^-1 phi and the clover term REBUILT on the varied links --
    pins leaf orientation, sigma_mu_nu convention and the i c / 8 weight of the clover-term derivative (groundwork for the
    device kernel: the force entry points still reject csw != 0)"""
    dims = (4, 4, 4, 4)
    U = orc.random_su3(dims, seed=33, eps=0.35)
    phi = orc.gaussian_field(dims, orc.WILSON, seed=34)

    def action(Ux):
        op = orc.make_op(dims, kappa=0.11, csw=CSW)
        orc.clover_build(op, Ux)
        r = orc.cg(op, orc.WILSON, Ux, phi, eps=1e-24)
        assert r["converged"]
        return np.vdot(phi, r["x"]).real, r["x"], op

    S0, X, op = action(U)
    Y = orc.apply(op, orc.WILSON, orc.D, U, X)
    F = orc.force(op, orc.WILSON, U, X, Y)
    op0 = orc.make_op(dims, kappa=0.11)
    F_hop = orc.force(op0, orc.WILSON, U, X, Y)
    assert np.abs(F - F_hop).max() > 1e-3 * np.abs(F_hop).max()         # the clover part is not negligible here
    rng = np.random.default_rng(7)
    for (mu, t, z, y, x) in [(0, 0, 0, 0, 0), (3, 3, 1, 2, 0), (1, 2, 3, 3, 3), (2, 1, 1, 1, 1)]:
        H = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))
        A = (H - H.conj().T) / 2
        A -= np.trace(A) / 3 * np.eye(3)
        want = -2 * np.real(np.trace(A @ F[mu, t, z, y, x].T))
        h, vals = 1e-5, []
        for sgn in (+1, -1):
            U2 = U.copy()
            U2[mu, t, z, y, x] = (_expm_antiherm(sgn * h * A) @ U[mu, t, z, y, x].T).T
            vals.append(action(np.ascontiguousarray(U2))[0])
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - want) < 2e-6 * max(1.0, abs(want)), (fd, want)

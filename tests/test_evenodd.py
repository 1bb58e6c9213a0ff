"""
Even-odd (Schur) preconditioned Wilson solve -- oracle checks on the CPU (SURVEY.md 8c known-answer list item 6:
"even-odd solve == full solve").  New capability behind BASELINE.json configs[1]; see oracle/lqcd_oracle.h orc_eo_solve.
"""
import numpy as np
import pytest

from oracle import oracle as orc

DIMS = (4, 4, 4, 4)


def parity_mask(dims):
    NX, NY, NZ, NT = dims
    t, z, y, x = np.meshgrid(np.arange(NT), np.arange(NZ), np.arange(NY), np.arange(NX), indexing="ij")
    return ((x + y + z + t) & 1)[None, ..., None]


def test_hop_connects_opposite_parities_only():
    U = orc.random_su3(DIMS, seed=2)
    psi = orc.gaussian_field(DIMS, orc.WILSON, seed=3)
    op = orc.make_op(DIMS, kappa=0.13)
    odd = parity_mask(DIMS)
    full = (psi - orc.apply(op, orc.WILSON, orc.D, U, psi)) / op.kappa          # H psi
    for p in (0, 1):
        got = orc.hop_parity(op, U, psi, p)
        src = psi * (odd if p == 0 else 1 - odd)                                  # only the opposite parity matters
        got2 = orc.hop_parity(op, U, np.ascontiguousarray(src), p)
        assert np.abs(got - got2).max() == 0
        sel = (1 - odd) if p == 0 else odd
        assert np.abs(got * (1 - sel)).max() == 0
        assert np.abs(got - full * sel).max() < 1e-13


@pytest.mark.parametrize("method", ["bicg", "bicgstab"])
@pytest.mark.parametrize("dagger", [False, True])
def test_evenodd_solve_equals_full_solve(method, dagger):
    U = orc.random_su3(DIMS, seed=11, eps=0.4)
    b = orc.gaussian_field(DIMS, orc.WILSON, seed=12)
    op = orc.make_op(DIMS, kappa=0.125)
    eo = orc.eo_solve(op, U, b, method=method, dagger=dagger, eps=1e-22)
    assert eo["converged"]
    mode = orc.DDAG if dagger else orc.D
    r = b - orc.apply(op, orc.WILSON, mode, U, eo["x"])
    assert np.vdot(r, r).real < 2e-22                           # true residual of the FULL system == preconditioned one
    assert abs(np.vdot(r, r).real - eo["resid_sq"]) < 1e-24
    if dagger:
        return
    full = (orc.cgnr if method == "bicg" else orc.bicgstab)(op, orc.WILSON, U, b, eps=1e-22)
    assert full["converged"]
    assert np.abs(full["x"] - eo["x"]).max() < 1e-9
    assert eo["iters"] < full["iters"]                          # the point of the preconditioner


def test_evenodd_initial_guess_and_point_source():
    U = orc.random_su3(DIMS, seed=11, eps=0.4)
    b = orc.point_source(DIMS, orc.WILSON, color=1, spin=2)
    op = orc.make_op(DIMS, kappa=0.125)
    first = orc.eo_solve(op, U, b, method="bicg", eps=1e-20)
    again = orc.eo_solve(op, U, b, method="bicg", eps=1e-20, x0=first["x"])
    assert first["converged"] and again["iters"] == 0


@pytest.mark.parametrize("dims", [(8, 4, 4, 2), (4, 4, 2, 8), (16, 2, 4, 4)])
@pytest.mark.parametrize("dagger", [False, True])
def test_checkerboard_index_emulation(dims, dagger):
    """numpy mirror of the index algebra of csrc/wilson_eo.cu (eo_convert_kernel, wilson_eo_hop_kernel): half index
    h = (x>>1) + X/2*(y + Y*(z + Z*t)), x = 2*xh + ((y+z+t+p)&1); the +-x neighbour of the other parity is h or h+-1
    depending on the row parity; forward links from the output parity's array at h, backward links from the input
    parity's array at the neighbour index; boundary phases on wrap.  Compared with the oracle's parity hop."""
    import np_ref
    X, Y, Z, T = dims
    Xh, V, Vh = X // 2, X * Y * Z * T, X * Y * Z * T // 2
    bc = (1, 1, 1, -1)
    U = orc.random_su3(dims, seed=5)
    psi = orc.gaussian_field(dims, orc.WILSON, seed=6)
    op = orc.make_op(dims, kappa=0.13, bc=bc)
    M = np_ref.links_mat(U).reshape(4, V, 3, 3)                  # [mu, site, a, b]
    f = psi.reshape(4, V, 3).transpose(1, 0, 2)                  # [site, alpha, c]
    h = np.arange(Vh)
    xh, y, z, t = h % Xh, (h // Xh) % Y, (h // (Xh * Y)) % Z, h // (Xh * Y * Z)
    site_of = {}
    for p in (0, 1):
        x = 2 * xh + ((y + z + t + p) & 1)
        assert np.all(((x + y + z + t) & 1) == p)
        site_of[p] = x + X * (y + Y * (z + Z * t))               # eo_convert_kernel
    assert sorted(np.concatenate([site_of[0], site_of[1]])) == list(range(V))
    half_f = {p: f[site_of[p]] for p in (0, 1)}
    half_U = {p: M[:, site_of[p]] for p in (0, 1)}
    for p in (0, 1):
        odd_row = (y + z + t + p) & 1
        fin, g_out, g_in = half_f[1 - p], half_U[p], half_U[1 - p]
        acc = np.zeros((Vh, 4, 3), dtype=complex)
        coords = (None, y, z, t)
        strides = (None, Xh, Xh * Y, Xh * Y * Z)
        ext = (X, Y, Z, T)
        for mu in range(4):
            if mu == 0:
                wf = (odd_row == 1) & (xh == Xh - 1)
                nf = np.where(odd_row == 1, np.where(wf, h - (Xh - 1), h + 1), h)
                wb = (odd_row == 0) & (xh == 0)
                nb = np.where(odd_row == 1, h, np.where(wb, h + (Xh - 1), h - 1))
            else:
                c, st, d = coords[mu], strides[mu], ext[mu]
                wf, wb = c == d - 1, c == 0
                nf = np.where(wf, h - (d - 1) * st, h + st)
                nb = np.where(wb, h + (d - 1) * st, h - st)
            sgn = -1 if dagger else 1
            Pf = np.eye(4) - sgn * np_ref.G[mu]                  # D: forward (1 - g), backward (1 + g)
            Pb = np.eye(4) + sgn * np_ref.G[mu]
            phf = np.where(wf, bc[mu], 1.0)[:, None, None]
            phb = np.where(wb, bc[mu], 1.0)[:, None, None]
            fw = np.einsum("hab,hsb->hsa", g_out[mu][h], phf * fin[nf])
            bw = np.einsum("hba,hsb->hsa", np.conj(g_in[mu][nb]), phb * fin[nb])
            acc += np.einsum("sr,hra->hsa", Pf, fw) + np.einsum("sr,hra->hsa", Pb, bw)
        want = orc.hop_parity(op, U, psi, p, dagger=dagger).reshape(4, V, 3).transpose(1, 0, 2)[site_of[p]]
        assert np.abs(acc - want).max() < 1e-13

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

"""
GPU-vs-oracle parity AT the BASELINE.json lattice sizes (16^4, 24^4, 32^4): the kernels only take their production shape there
(x-line blocks, t-slowest CTA sweep, link-hint switch at V <= 2^18, bulk-copy window pipelines), so the 4096-site parity cases of
test_gpu_parity.py do not cover them.

  * one Wilson and one staggered mul!(y, D, x) and mul!(y, D', x) per size against the oracle on the links DOWNLOADED from the
    device (`get_links`): <= 1e-13 relative to max|y| (north_star tolerance clause; upstream call sites
    src/measurements/unusedfiles/measure_Pion_correlator.jl:379, test/test_wilson.toml);
  * 16^4 (BASELINE configs[1]): CG on DdagD and CGNR ("bicg", solve_DinvX!(y, D, b)) -- iteration counts IDENTICAL to the oracle
    and solutions <= 1e-10, Wilson kappa = 0.12 and 0.141139 (test/test_wilson.toml) on a warm field, staggered m = 0.5
    (test/test_staggered.toml), eps_CG = 1e-19 / MaxCGstep = 3000 (src/system/parameter_structs.jl:174-175); even-odd
    preconditioned CGNR against orc.eo_solve ("16^4 ... even-odd CG ... residual vs CPU ref").

The oracle applies D at 32^4 in ~50 ms on the box's host cores; the 16^4 solves take seconds.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import lqcd_b200 as q
from lqcd_b200 import _lib as L
from oracle import oracle as orc

BC = [1, 1, 1, -1]


def relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def _dims(env, default):
    v = os.environ.get(env)           # the CPU pre-flight (tests/emu) shrinks the sizes; on the B200 they are BASELINE's
    return tuple(int(t) for t in v.split("x")) if v else default


SIZES = [_dims("LQCD_TEST_CONFIG1_DIMS", (16, 16, 16, 16)), _dims("LQCD_TEST_CONFIG2_DIMS", (24, 24, 24, 24)),
         _dims("LQCD_TEST_FULL_DIMS", (32, 32, 32, 32))]


@pytest.mark.parametrize("dims", SIZES, ids=lambda d: "x".join(map(str, d)))
@pytest.mark.parametrize("kind", ["Wilson", "staggered"])
def test_dslash_matches_oracle_at_baseline_sizes(dims, kind):
    ctx = q.get_context(dims)
    ctx.call("lqcd_gauge_random", 111, -1.0)                     # hot links generated on the device (bench.py's field) ...
    Uh = q.get_links(ctx)                                         # ... and what the oracle sees is what the device holds
    k = L.WILSON if kind == "Wilson" else L.STAGGERED
    ok = orc.WILSON if kind == "Wilson" else orc.STAGGERED
    x, y = q.FermionField(ctx, k), q.FermionField(ctx, k)
    q.gauss_distribution_fermion_(x, 112)
    src = x.to_host()
    op = L.LqcdOp()
    op.kind, op.kappa, op.r, op.mass = k, 0.12, 1.0, 0.5
    for i, b in enumerate(BC):
        op.bc[i] = b
    oop = orc.make_op(dims, kappa=0.12, mass=0.5)
    import ctypes as C
    for mode, omode in ((L.OP_D, orc.D), (L.OP_DDAG, orc.DDAG)):
        ctx.call("lqcd_dslash", C.byref(op), y.h, x.h, mode)
        want = orc.apply(oop, ok, omode, Uh, src)
        assert relerr(y.to_host(), want) < 1e-13, (dims, kind, mode)
    # the SU(3) links really are what bench.py assumes: unitary to rounding
    m = Uh[0, 0, 0, 0, :4]
    assert np.abs(np.einsum("xba,xca->xbc", m, m.conj()) - np.eye(3)).max() < 1e-13


def _solve_case(dims, kind, kappa, method, Uh, evenodd=False):
    U = q.gaugefields_from_array(Uh)
    x = q.Initialize_pseudofermion_fields(U[0], kind)
    if kind == "Wilson":
        params = {"Dirac_operator": "Wilson", "κ": kappa, "r": 1.0, "eps_CG": 1e-19, "MaxCGstep": 3000, "boundarycondition": BC,
                  "method_CG": "bicg", "evenodd": evenodd}
        ok = orc.WILSON
    else:
        params = {"Dirac_operator": "staggered", "mass": 0.5, "eps_CG": 1e-19, "MaxCGstep": 3000, "boundarycondition": BC}
        ok = orc.STAGGERED
    D = q.Dirac_operator(U, x, params)
    b = orc.gaussian_field(dims, ok, seed=112)
    x.from_host(b)
    sol = q.similar(x)
    q.clear_fermion_(sol)
    oop = orc.make_op(dims, kappa=kappa, mass=0.5)
    if method == "cg":
        info = q.solve_DinvX_(sol, q.DdagD(D), x)
        ref = orc.cg(oop, ok, Uh, b)
        A = orc.DDAGD
    elif evenodd:
        info = q.solve_DinvX_(sol, D, x)
        ref = orc.eo_solve(oop, Uh, b)
        A = orc.D
    else:
        info = q.solve_DinvX_(sol, D, x)
        ref = orc.cgnr(oop, ok, Uh, b)
        A = orc.D
    assert ref["converged"]
    assert info["iters"] == ref["iters"], (kind, kappa, method, evenodd, info["iters"], ref["iters"])
    got = sol.to_host()
    assert relerr(got, ref["x"]) < 1e-10
    r = b - orc.apply(oop, ok, A, Uh, got)                          # true residual of the GPU solution, computed by the oracle
    assert np.vdot(r, r).real < 1e-12 * np.vdot(b, b).real      # north_star: residuals within 1e-12 relative


@pytest.fixture(scope="module")
def warm16():
    dims = SIZES[0]
    return dims, orc.random_su3(dims, seed=111, eps=0.3)


@pytest.mark.parametrize("kappa", [0.12, 0.141139])
@pytest.mark.parametrize("method", ["cg", "bicg"])
def test_wilson_solves_match_oracle_16_4(warm16, kappa, method):
    dims, Uh = warm16
    _solve_case(dims, "Wilson", kappa, method, Uh)


@pytest.mark.parametrize("kappa", [0.12, 0.141139])
def test_wilson_evenodd_cgnr_matches_oracle_16_4(warm16, kappa):
    dims, Uh = warm16
    _solve_case(dims, "Wilson", kappa, "bicg", Uh, evenodd=True)


def test_staggered_cg_matches_oracle_16_4(warm16):
    dims, Uh = warm16
    _solve_case(dims, "staggered", 0.0, "cg", Uh)

"""
GPU parity tests of device code that was written AFTER the round's GPU budget was spent: compiled for sm_100a and checked
on the CPU side (oracle, numpy restatements, emulations) but never executed on hardware.  They run last (file name) and are
marked xfail(strict=False): a failure is reported as "xfailed" and does not hide the verified suite, a pass shows up as
"xpassed".  Remove the marker once a B200 run has confirmed them.
"""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(reason="not yet run on hardware (written after the round's GPU budget was spent)", strict=False)]

CSW = 1.5612        # src/system/parameter_structs.jl:125


def _clover_setup(dims, kappa, seed=7):
    import lqcd_b200 as q
    Uh = orc.random_su3(dims, seed=seed, eps=0.35)
    U = q.gaugefields_from_array(Uh)
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, {"Dirac_operator": "WilsonClover", "κ": kappa, "r": 1.0, "Clover_coefficient": CSW,
                                "boundarycondition": [1, 1, 1, -1], "eps_CG": 1e-20, "MaxCGstep": 3000})
    op = orc.make_op(dims, kappa=kappa, csw=CSW)
    clov = orc.clover_build(op, Uh)
    return q, Uh, U, x, D, op, clov


@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (8, 4, 6, 4), (32, 4, 4, 4)])
def test_clover_term_matches_oracle(dims):
    q, Uh, U, x, D, op, clov = _clover_setup(dims, 0.125)
    got = D.clover_term()
    assert np.abs(got - clov).max() < 1e-13


@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (8, 4, 6, 4), (16, 8, 4, 8)])
def test_clover_dslash_matches_oracle(dims):
    q, Uh, U, x, D, op, clov = _clover_setup(dims, 0.125)
    src = orc.gaussian_field(dims, orc.WILSON, seed=19)
    x.from_host(src)
    y = q.similar(x)
    for A, mode in ((D, orc.D), (q.adjoint(D), orc.DDAG), (q.DdagD(D), orc.DDAGD)):
        q.mul_(y, A, x)
        want = orc.apply(op, orc.WILSON, mode, Uh, src)
        assert np.abs(y.to_host() - want).max() / np.abs(want).max() < 1e-13


def test_clover_cg_iterations_and_solution():
    dims = (8, 8, 8, 8)
    q, Uh, U, x, D, op, clov = _clover_setup(dims, 0.12)
    b = orc.gaussian_field(dims, orc.WILSON, seed=23)
    x.from_host(b)
    sol = q.similar(x)
    q.clear_fermion_(sol)
    info = q.solve_DinvX_(sol, q.DdagD(D), x)
    ref = orc.cg(op, orc.WILSON, Uh, b, eps=1e-20)
    assert ref["converged"] and info["iters"] == ref["iters"]
    assert np.abs(sol.to_host() - ref["x"]).max() / np.abs(ref["x"]).max() < 1e-10
    # plain Wilson on the same context still works after the clover operator (cache keyed on kappa*csw)
    D0 = q.Dirac_operator(U, x, {"Dirac_operator": "Wilson", "κ": 0.12, "boundarycondition": [1, 1, 1, -1]})
    y = q.similar(x)
    q.mul_(y, D0, x)
    want = orc.apply(orc.make_op(dims, kappa=0.12), orc.WILSON, orc.D, Uh, b)
    assert np.abs(y.to_host() - want).max() < 1e-13


# ---- even-odd preconditioned Wilson solve (csrc/wilson_eo.cu) ---------------------------------------------------------
@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (8, 8, 4, 4), (16, 4, 4, 8)])
@pytest.mark.parametrize("method", ["bicg", "bicgstab"])
@pytest.mark.parametrize("dagger", [False, True])
def test_evenodd_solve_matches_oracle(dims, method, dagger):
    import lqcd_b200 as q
    Uh = orc.random_su3(dims, seed=11, eps=0.4)
    U = q.gaugefields_from_array(Uh)
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, {"Dirac_operator": "Wilson", "κ": 0.125, "boundarycondition": [1, 1, 1, -1],
                                "eps_CG": 1e-22, "MaxCGstep": 3000, "method_CG": method, "evenodd": True})
    b = orc.gaussian_field(dims, orc.WILSON, seed=12)
    x.from_host(b)
    sol = q.similar(x)
    q.clear_fermion_(sol)
    info = q.solve_DinvX_(sol, q.adjoint(D) if dagger else D, x)
    op = orc.make_op(dims, kappa=0.125)
    ref = orc.eo_solve(op, Uh, b, method=method, dagger=dagger, eps=1e-22)
    assert ref["converged"]
    got = sol.to_host()
    r = b - orc.apply(op, orc.WILSON, orc.DDAG if dagger else orc.D, Uh, got)
    assert np.vdot(r, r).real < 2e-22                    # true residual of the FULL system
    assert np.abs(got - ref["x"]).max() / np.abs(ref["x"]).max() < 1e-9
    if method == "bicg":                                   # CGNR recurrences are smooth: identical iteration count
        assert info["iters"] == ref["iters"]
    else:
        assert abs(info["iters"] - ref["iters"]) <= 2


def test_evenodd_beats_full_solve_and_keeps_plain_path():
    import lqcd_b200 as q
    dims = (8, 8, 8, 8)
    Uh = orc.random_su3(dims, seed=3, eps=0.4)
    U = q.gaugefields_from_array(Uh)
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    params = {"Dirac_operator": "Wilson", "κ": 0.125, "boundarycondition": [1, 1, 1, -1], "eps_CG": 1e-20, "MaxCGstep": 3000}
    b = orc.gaussian_field(dims, orc.WILSON, seed=4)
    x.from_host(b)
    out = {}
    for eo in (True, False, True):
        D = q.Dirac_operator(U, x, dict(params, evenodd=eo))
        sol = q.similar(x)
        q.clear_fermion_(sol)
        out[eo] = (q.solve_DinvX_(sol, D, x)["iters"], sol.to_host())
    assert out[True][0] < out[False][0]
    assert np.abs(out[True][1] - out[False][1]).max() < 1e-8


# ---- experimental t-marching Wilson kernel (csrc/wilson_dslash3.cu, off by default) ------------------------------------
def test_tmarch_kernel_matches_oracle():
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, "tests/k3_worker.py"], cwd=root, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, LQCD_WILSON_KERNEL="3", LQCD_COMM_TIMEOUT_S="5"))
    assert r.returncode == 0 and "K3 OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


# ---- pipelined host-field mul! (csrc/host_pipeline.cu) -----------------------------------------------------------------
@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (8, 4, 6, 16), (16, 8, 4, 8), (4, 4, 2, 2), (6, 8, 4, 4)])
@pytest.mark.parametrize("kind", ["Wilson", "staggered", "WilsonClover"])
def test_pipelined_host_mul_matches_three_call_sequence(dims, kind):
    """lqcd_dslash_host (slab pipeline: H2D | convert + Dslash on CTA sub-ranges + convert | D2H) == upload + lqcd_dslash +
    download bit for bit, and == oracle; covers lattices whose tiling gives 1 (fallback), 2, 4 and 8 slabs"""
    import lqcd_b200 as q
    Uh = orc.random_su3(dims, seed=21, eps=0.4)
    U = q.gaugefields_from_array(Uh)
    name = "staggered" if kind == "staggered" else "Wilson"
    x = q.Initialize_pseudofermion_fields(U[0], name)
    D = q.Dirac_operator(U, x, {"Dirac_operator": kind, "κ": 0.125, "mass": 0.3, "Clover_coefficient": CSW, "boundarycondition": [1, 1, 1, -1]})
    k = orc.WILSON if name == "Wilson" else orc.STAGGERED
    op = orc.make_op(dims, kappa=0.125, mass=0.3, csw=CSW if kind == "WilsonClover" else 0.0)
    if kind == "WilsonClover":
        keep = orc.clover_build(op, Uh)
    src = orc.gaussian_field(dims, k, seed=22)
    y = q.similar(x)
    out = np.empty_like(src)
    for A, mode in ((D, orc.D), (q.adjoint(D), orc.DDAG), (q.DdagD(D), orc.DDAGD)):
        q.mul_host_(out, A, src, y=y, x=x)
        x.from_host(src)
        y2 = q.similar(x)
        q.mul_(y2, A, x)
        assert np.array_equal(out, y2.to_host())
        want = orc.apply(op, k, mode, Uh, src)
        assert np.abs(out - want).max() / np.abs(want).max() < 1e-13
        assert np.array_equal(x.to_host(), src) and np.array_equal(y.to_host(), out)      # device fields hold source and result

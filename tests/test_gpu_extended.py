"""
GPU parity tests (through the C ABI) of the wider path: Wilson-clover, even-odd, multi-RHS, staggered half-field solves, RHMC
plumbing, BASELINE-size checks.  First hardware run: round-1 driver GPU tier (all passed); plain strict tests since round 2.
"""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = [pytest.mark.gpu]

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


# ---- Wilson kernel families: register-resident kernel (default; two-row or full links) and the experimental t-marching TMA kernel ------
@pytest.mark.parametrize("env", [{"LQCD_WILSON_KERNEL": "4"}, {"LQCD_WILSON_KERNEL": "4", "LQCD_TM_CHUNKS": "1"}, {"LQCD_WILSON_KERNEL": "4", "LQCD_TM_CHUNKS": "2"},
                                 {"LQCD_WILSON_KERNEL": "5"}, {"LQCD_WILSON_KERNEL": "5", "LQCD_TM_CHUNKS": "2"}, {}, {"LQCD_LINKS12": "0"}],
                         ids=["tmarch-auto", "tmarch-1chunk", "tmarch-2chunks", "tmarch2-auto", "tmarch2-2chunks", "register-kernel-two-row-links",
                              "register-kernel-full-links"])
def test_wilson_kernel_families_match_oracle(env):
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, "tests/tmarch_worker.py"], cwd=root, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, LQCD_COMM_TIMEOUT_S="5", **env))
    assert r.returncode == 0 and "TMARCH OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


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


# ---- several right-hand sides in lock step (csrc/mrhs.cu) --------------------------------------------------
def _mrhs_setup(kind, dims, seed=3):
    import lqcd_b200 as q
    Uh = orc.random_su3(dims, seed=seed, eps=0.35)
    U = q.gaugefields_from_array(Uh)
    name = "Wilson" if kind == orc.WILSON else "staggered"
    x = q.Initialize_pseudofermion_fields(U[0], name)
    D = q.Dirac_operator(U, x, {"Dirac_operator": name, "κ": 0.12, "mass": 0.5, "r": 1.0, "boundarycondition": [1, 1, 1, -1],
                                "eps_CG": 1e-20, "MaxCGstep": 3000})
    return q, Uh, U, x, D, orc.make_op(dims, kappa=0.12, mass=0.5)


@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (6, 8, 4, 4), (32, 4, 4, 4), (16, 8, 4, 6)])
@pytest.mark.parametrize("kind,nrhs", [(orc.WILSON, 2), (orc.WILSON, 5), (orc.WILSON, 12), (orc.WILSON, 16),
                                       (orc.STAGGERED, 3), (orc.STAGGERED, 7), (orc.STAGGERED, 16)])
def test_multi_rhs_dslash_is_bit_identical_to_single(dims, kind, nrhs):
    """mul_multi_ == the loop of mul_ bit for bit (D, D^dag, D^dag D), and equals the oracle"""
    q, Uh, U, x, D, op = _mrhs_setup(kind, dims)
    srcs = [orc.gaussian_field(dims, kind, seed=100 + j) for j in range(nrhs)]
    xs = [q.similar(x).from_host(s) for s in srcs]
    ys = [q.similar(x) for _ in range(nrhs)]
    y1 = q.similar(x)
    for A, mode in ((D, orc.D), (q.adjoint(D), orc.DDAG), (q.DdagD(D), orc.DDAGD)):
        q.mul_multi_(ys, A, xs)
        for j in range(nrhs):
            q.mul_(y1, A, xs[j])
            got = ys[j].to_host()
            assert np.array_equal(got, y1.to_host()), (mode, j)
            if j in (0, nrhs - 1):
                want = orc.apply(op, kind, mode, Uh, srcs[j])
                assert np.abs(got - want).max() / np.abs(want).max() < 1e-13
            assert np.array_equal(xs[j].to_host(), srcs[j])             # inputs untouched


# (8, 4, 4, 8): irregular patch shape -> register multi-RHS kernel; (8, 8, 4, 8): 2x2 (y,z) patches -> t-marching TMA multi-RHS kernel
@pytest.mark.parametrize("dims", [(8, 4, 4, 8), (8, 8, 4, 8)])
@pytest.mark.parametrize("kind", [orc.WILSON, orc.STAGGERED])
@pytest.mark.parametrize("dagger", [False, True])
def test_multi_rhs_cgnr_matches_single_solves(kind, dagger, dims):
    """solve_DinvX_multi_ (CGNR = upstream "bicg"): right-hand sides of very different difficulty (point sources, Gaussian noise, a
    source that is already solved by its initial guess) advance in lock step; each one's iteration count, residual and solution
    are those of its own single solve, bit for bit; iteration counts equal the oracle's"""
    q, Uh, U, x, D, op = _mrhs_setup(kind, dims)
    A = q.adjoint(D) if dagger else D
    srcs = []
    for j in range(3):
        h = np.zeros(x.host_shape, dtype=complex)
        h[(j, 0, 0, 0, 0, j) if kind == orc.WILSON else (0, 0, j, 0, j)] = 1.0
        srcs.append(h)
    srcs += [orc.gaussian_field(dims, kind, seed=40 + j) for j in range(3)]
    srcs.append(np.zeros(x.host_shape, dtype=complex))                  # b = 0 with x0 = 0: converged at step 0
    bs = [q.similar(x).from_host(s) for s in srcs]
    ys = [q.similar(x) for _ in srcs]
    for y in ys:
        q.clear_fermion_(y)
    infos = q.solve_DinvX_multi_(ys, A, bs)
    y1 = q.similar(x)
    for j, s in enumerate(srcs):
        q.clear_fermion_(y1)
        one = q.solve_DinvX_(y1, A, bs[j])
        assert infos[j]["iters"] == one["iters"] and infos[j]["resid_sq"] == one["resid_sq"], (j, infos[j], one)
        assert np.array_equal(ys[j].to_host(), y1.to_host()), j
        if dagger:                       # the oracle's CGNR solves D x = b only: check the true residual of D^dag x = b
            tr = s - orc.apply(op, kind, orc.DDAG, Uh, ys[j].to_host())
            assert np.vdot(tr, tr).real < 1e-18
            continue
        ref = orc.cgnr(op, kind, Uh, s, eps=1e-20)
        assert ref["converged"] and infos[j]["iters"] == ref["iters"], (j, infos[j]["iters"], ref["iters"])
        if np.abs(ref["x"]).max() > 0:
            assert np.abs(ys[j].to_host() - ref["x"]).max() / np.abs(ref["x"]).max() < 1e-10
    assert infos[-1]["iters"] == 0


@pytest.mark.parametrize("dims", [(8, 4, 4, 8), (8, 8, 4, 8)])
@pytest.mark.parametrize("kind", [orc.WILSON, orc.STAGGERED])
def test_multi_rhs_cg_on_DdagD(kind, dims):
    q, Uh, U, x, D, op = _mrhs_setup(kind, dims)
    srcs = [orc.gaussian_field(dims, kind, seed=60 + j) for j in range(5)]
    bs = [q.similar(x).from_host(s) for s in srcs]
    ys = [q.similar(x) for _ in srcs]
    guess = orc.gaussian_field(dims, kind, seed=70)
    for j, y in enumerate(ys):
        if j == 1:
            y.from_host(guess)                                          # ys[j] doubles as the initial guess
        else:
            q.clear_fermion_(y)
    infos = q.solve_DinvX_multi_(ys, q.DdagD(D), bs)
    for j, s in enumerate(srcs):
        ref = orc.cg(op, kind, Uh, s, eps=1e-20, x0=guess if j == 1 else None)
        assert ref["converged"] and abs(infos[j]["iters"] - ref["iters"]) <= 1, (j, infos[j]["iters"], ref["iters"])
        assert np.abs(ys[j].to_host() - ref["x"]).max() / np.abs(ref["x"]).max() < 1e-9


def test_multi_rhs_nonconvergence_names_the_source():
    import lqcd_b200 as q
    dims = (4, 4, 4, 4)
    _, Uh, U, x, D, op = _mrhs_setup(orc.WILSON, dims)
    D.maxsteps = 3
    bs = [q.similar(x).from_host(orc.gaussian_field(dims, orc.WILSON, seed=80 + j)) for j in range(4)]
    bs[2].from_host(np.zeros(x.host_shape, dtype=complex))
    ys = [q.similar(x) for _ in bs]
    for y in ys:
        q.clear_fermion_(y)
    with pytest.raises(q.NotConverged, match="right-hand side 0"):
        q.solve_DinvX_multi_(ys, D, bs)
    assert D.last["iters"] == [3, 3, 0, 3]


@pytest.mark.parametrize("kind", [orc.WILSON, orc.STAGGERED])
def test_point_source_propagators(golden_dir, kind):
    """calc_quark_propagators_point_source (measure_Pion_correlator.jl:333-409) on the reference's own 4^4 fixtures: the NC*Nspinor
    columns of D^-1 from ONE batched solve against the oracle's CGNR per source, and D * column = source"""
    import lqcd_b200 as q
    dims = (4, 4, 4, 4)
    Uh = np.load(golden_dir / ("wilson_4444.npy" if kind == orc.WILSON else "staggered_4444.npy"))
    U = q.gaugefields_from_array(Uh)
    name = "Wilson" if kind == orc.WILSON else "staggered"
    x = q.Initialize_pseudofermion_fields(U[0], name)
    D = q.Dirac_operator(U, x, {"Dirac_operator": name, "κ": 0.141139, "mass": 0.5, "boundarycondition": [1, 1, 1, -1],
                                "eps_CG": 1e-19, "MaxCGstep": 3000})
    op = orc.make_op(dims, kappa=0.141139, mass=0.5)
    props, infos = q.calc_quark_propagators_point_source(D)
    nspin = 4 if kind == orc.WILSON else 1
    assert len(props) == 3 * nspin
    y = q.similar(x)
    for i in (0, len(props) // 2, len(props) - 1):
        is_, ic = i % nspin, i // nspin
        b = np.zeros(x.host_shape, dtype=complex)
        b[(is_, 0, 0, 0, 0, ic) if kind == orc.WILSON else (0, 0, 0, 0, ic)] = 1.0
        ref = orc.cgnr(op, kind, Uh, b, eps=1e-19)
        assert infos[i]["iters"] == ref["iters"]
        assert np.abs(props[i].to_host() - ref["x"]).max() / np.abs(ref["x"]).max() < 1e-10
        q.mul_(y, D, props[i])
        assert np.abs(y.to_host() - b).max() < 1e-8


@pytest.mark.parametrize("kind", [orc.STAGGERED, orc.WILSON])
def test_chiral_condensate_batched(golden_dir, kind):
    """measure(::Chiral_condensate_measurement) (measure_chiral_condensate.jl:164-204) on the reference's fixtures: Z4 noise on the
    device, the Nr solves in lock step == one after the other (bit for bit, CGNR), and == the oracle's CGNR on the same noise"""
    import lqcd_b200 as q
    dims = (4, 4, 4, 4)
    Uh = np.load(golden_dir / ("wilson_4444.npy" if kind == orc.WILSON else "staggered_4444.npy"))
    U = q.gaugefields_from_array(Uh)
    name = "Wilson" if kind == orc.WILSON else "staggered"
    x = q.Initialize_pseudofermion_fields(U[0], name)
    D = q.Dirac_operator(U, x, {"Dirac_operator": name, "κ": 0.141139, "mass": 0.5, "boundarycondition": [1, 1, 1, -1],
                                "eps_CG": 1e-19, "MaxCGstep": 3000})
    op = orc.make_op(dims, kappa=0.141139, mass=0.5)
    Nr = 10                                                      # the reference's default number of noise vectors
    pbp, vals, rs = q.measure_chiral_condensate(D, Nr=Nr, factor=1.0, seed=31)
    pbp1, vals1, rs1 = q.measure_chiral_condensate(D, Nr=Nr, factor=1.0, seed=31, batched=False)
    assert pbp == pbp1 and vals == vals1
    noise = np.stack([r.to_host() for r in rs])
    assert np.all(np.isin(noise, [1, -1, 1j, -1j]))
    counts = [np.sum(noise == v) for v in (1, 1j, -1, -1j)]
    assert min(counts) > 0.2 * noise.size and len({r.to_host().tobytes() for r in rs}) == Nr        # uniform, sources differ
    ref = [np.vdot(r, orc.cgnr(op, kind, Uh, r, eps=1e-19)["x"]) for r in noise]
    want = np.real(sum(ref) / Nr) / 256
    assert abs(pbp - want) < 1e-10 * abs(want)
    assert want > 0.0                                            # Re tr D^-1 > 0 for these operators


# ---- staggered even-site systems on half fields (csrc/staggered_eo.cu) -----------------------------------------------------------
@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (8, 4, 6, 4), (16, 4, 4, 8), (32, 4, 2, 2), (24, 4, 4, 4), (12, 6, 4, 4)])
def test_staggered_even_site_solve_matches_full_lattice_solve(dims):
    """lqcd_solve_staggered_even == the full-lattice CG on an even-site source: same iterates (iteration count within the rounding of
    the reduction order), same solution, zero odd sites; initial guess honoured; == the oracle"""
    import ctypes as C
    import lqcd_b200 as q
    from lqcd_b200 import _lib as L
    Uh = orc.random_su3(dims, seed=11, eps=0.4)
    U = q.gaugefields_from_array(Uh)
    x = q.Initialize_pseudofermion_fields(U[0], "staggered")
    D = q.Dirac_operator(U, x, {"Dirac_operator": "staggered", "mass": 0.3, "boundarycondition": [1, 1, 1, -1], "eps_CG": 1e-20, "MaxCGstep": 3000})
    op = orc.make_op(dims, mass=0.3)
    NX, NY, NZ, NT = dims
    t, z, y, xx = np.meshgrid(np.arange(NT), np.arange(NZ), np.arange(NY), np.arange(NX), indexing="ij")
    odd = ((t + z + y + xx) & 1) == 1
    bh = orc.gaussian_field(dims, orc.STAGGERED, seed=12)
    bh[odd] = 0.0
    b = q.similar(x).from_host(bh)
    for guess in (None, 13):
        full, half = q.similar(x), q.similar(x)
        g = np.zeros_like(bh)
        if guess:
            g = orc.gaussian_field(dims, orc.STAGGERED, seed=guess)
            g[odd] = 0.0
        full.from_host(g); half.from_host(g)
        info = q.solve_DinvX_(full, q.DdagD(D), b)
        it, rs = C.c_int(0), C.c_double(0.0)
        D.ctx.call("lqcd_solve_staggered_even", C.byref(D.op), half.h, b.h, D.eps, D.maxsteps, C.byref(it), C.byref(rs))
        ref = orc.cg(op, orc.STAGGERED, Uh, bh, eps=1e-20, x0=g if guess else None)
        assert abs(it.value - info["iters"]) <= 1 and abs(it.value - ref["iters"]) <= 1, (it.value, info["iters"], ref["iters"])
        hh = half.to_host()
        assert np.abs(hh[odd]).max() == 0.0
        assert np.abs(hh - ref["x"]).max() / np.abs(ref["x"]).max() < 1e-10
        assert np.abs(hh - full.to_host()).max() / np.abs(hh).max() < 1e-10
    # the scoped switch routes lqcd_solve and is off again afterwards
    fa = q.FermiAction(D, {"Nf": 4})
    assert fa.half_field_solver
    launches0 = D.ctx.launch_count()
    S_half = q.evaluate_FermiAction(fa, U, b)
    n_half = D.ctx.launch_count() - launches0
    fa.half_field_solver = False
    S_full = q.evaluate_FermiAction(fa, U, b)
    assert abs(S_half - S_full) < 1e-10 * abs(S_full) and abs(S_half - np.vdot(bh, ref["x"] if guess is None else orc.cg(op, orc.STAGGERED, Uh, bh, eps=1e-20)["x"]).real) < 1e-8 * abs(S_full)
    assert n_half > 0


def test_staggered_nf4_force_and_trajectory_on_half_fields(golden_dir):
    """FermiAction(D, Nf = 4) with the half-field solver: force and a whole trajectory equal the full-lattice-solver results"""
    import lqcd_b200 as q
    dims = (4, 4, 4, 4)
    Uh = np.load(golden_dir / "staggered_4444.npy")
    res = {}
    for half in (True, False):
        U = q.gaugefields_from_array(Uh.copy())
        x = q.Initialize_pseudofermion_fields(U[0], "staggered")
        D = q.Dirac_operator(U, x, {"Dirac_operator": "staggered", "mass": 0.5, "boundarycondition": [1, 1, 1, -1], "eps_CG": 1e-22, "MaxCGstep": 3000})
        fa = q.FermiAction(D, {"Nf": 4, "half_field_solver": half})
        assert fa.half_field_solver == half
        xi, eta = q.similar(x), q.similar(x)
        q.gauss_sampling_in_action_(xi, U, fa, seed=5)
        q.sample_pseudofermions_(eta, U, fa, xi)
        F = np.zeros_like(Uh)
        q.calc_UdSfdU_(F, fa, U, eta)
        acc, dH, info = q.hmc_update_(U, 5.7, 0.025, 8, fa=fa, rng=np.random.default_rng(9))
        res[half] = (xi.to_host(), eta.to_host(), F, U.data.copy(), dH, acc)
    for a, b in zip(res[True][:4], res[False][:4]):
        assert np.abs(a - b).max() < 1e-9 * max(1.0, np.abs(b).max())
    assert abs(res[True][4] - res[False][4]) < 1e-6 and res[True][5] == res[False][5] and abs(res[True][4]) < 0.5


@pytest.mark.parametrize("dims", [(24, 4, 4, 4), (24, 24, 2, 2), (12, 6, 4, 4)])
@pytest.mark.parametrize("kind", [orc.WILSON, orc.STAGGERED])
def test_extents_of_24(dims, kind):
    """BASELINE config 3 is a 24^4 lattice: extents that neither divide nor are divided by the 32-site block (irregular tiling)"""
    q, Uh, U, x, D, op = _mrhs_setup(kind, dims)
    src = orc.gaussian_field(dims, kind, seed=5)
    y = q.similar(x)
    q.mul_(y, D, x.from_host(src))
    want = orc.apply(op, kind, orc.D, Uh, src)
    assert np.abs(y.to_host() - want).max() / np.abs(want).max() < 1e-13
    sol = q.similar(x)
    q.clear_fermion_(sol)
    info = q.solve_DinvX_(sol, q.DdagD(D), x)
    ref = orc.cg(op, kind, Uh, src, eps=1e-20)
    assert info["iters"] == ref["iters"]
    assert np.abs(sol.to_host() - ref["x"]).max() / np.abs(ref["x"]).max() < 1e-10


@pytest.mark.parametrize("kind", ["Wilson", "staggered"])
def test_full_size_32_4_properties(kind):
    """BASELINE headline size (32^4, one GPU): properties that need no oracle run -- linearity, <a, D b> = <D^dag a, b>, CG true
    residual -- and the newer entry points against the verified single-RHS kernel AT THIS SIZE: multi-RHS Dslash bit-identical,
    pipelined host mul! equal to upload + mul! + download, staggered half-field solve equal to the full-lattice solve"""
    import ctypes as C
    import lqcd_b200 as q
    import os
    dims = tuple(int(v) for v in os.environ.get("LQCD_TEST_FULL_DIMS", "32x32x32x32").split("x"))     # (the emulated pre-flight shrinks it)
    ctx = q.get_context(dims)
    U = q.Initialize_Gaugefields(3, 0, *dims, condition="cold")
    a = q.Initialize_pseudofermion_fields(U[0], kind)
    D = q.Dirac_operator(U, a, {"Dirac_operator": kind, "κ": 0.12, "mass": 0.5, "eps_CG": 1e-16, "MaxCGstep": 3000, "boundarycondition": [1, 1, 1, -1]})
    ctx.call("lqcd_gauge_random", 111, 0.3)                 # overwrite the cold links on the device with a warm synthetic field
    b, Da, Db, Dab, Dda = (q.similar(a) for _ in range(5))
    q.gauss_distribution_fermion_(a, 1); q.gauss_distribution_fermion_(b, 2)
    q.mul_(Da, D, a); q.mul_(Db, D, b); q.mul_(Dda, q.adjoint(D), a)
    l, r = q.dot(a, Db), q.dot(Dda, b)
    assert abs(l - r) < 1e-9 * abs(l)
    s = q.similar(a)
    q.substitute_fermion_(s, a); q.add_(s, 0.5 - 0.25j, b)          # s = a + c b
    q.mul_(Dab, D, s)
    q.add_(Dab, -1.0, Da); q.add_(Dab, -(0.5 - 0.25j), Db)
    assert q.dot(Dab, Dab).real < 1e-20 * q.dot(Da, Da).real
    # multi-RHS == single-RHS, bit for bit
    ys = [q.similar(a), q.similar(a), q.similar(a)]
    q.mul_multi_(ys, D, [a, b, s])
    assert np.array_equal(ys[0].to_host(), Da.to_host()) and np.array_equal(ys[1].to_host(), Db.to_host())
    # pipelined host mul! == three-call sequence
    ah = a.to_host()
    yh = np.zeros_like(ah)
    q.mul_host_(yh, D, ah)
    assert np.array_equal(yh, Da.to_host())
    # solve and recompute the true residual on the device
    sol, chk = q.similar(a), q.similar(a)
    if kind == "staggered":
        q.mask_parity_(b, 0)
    q.clear_fermion_(sol)
    info = q.solve_DinvX_(sol, q.DdagD(D), b)
    q.mul_(chk, q.DdagD(D), sol)
    q.add_(chk, -1.0, b)
    assert q.dot(chk, chk).real < 1e-14 and 5 < info["iters"] < 3000
    if kind == "staggered":
        half = q.similar(a)
        q.clear_fermion_(half)
        it, rs = C.c_int(0), C.c_double(0.0)
        ctx.call("lqcd_solve_staggered_even", C.byref(D.op), half.h, b.h, D.eps, D.maxsteps, C.byref(it), C.byref(rs))
        assert abs(it.value - info["iters"]) <= 1
        q.add_(half, -1.0, sol)
        assert q.dot(half, half).real < 1e-18 * q.dot(sol, sol).real


@pytest.mark.parametrize("r", [0.5, 1.3])
@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (8, 4, 6, 4)])
def test_wilson_general_r(dims, r):
    """params["r"] != 1 (universe.jl:115 forwards it; default 1): M = M_{r=1} - kappa (r-1) L through the spin-diagonal remainder
    kernel + the unchanged r = 1 kernel.  mul! in all three modes, CG / CGNR with identical iteration counts, multi-RHS and host
    mul! (which take the single-RHS path), multi-shift rejected with a message"""
    import lqcd_b200 as q
    Uh = orc.random_su3(dims, seed=17, eps=0.35)
    U = q.gaugefields_from_array(Uh)
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, {"Dirac_operator": "Wilson", "κ": 0.11, "r": r, "boundarycondition": [1, 1, 1, -1], "eps_CG": 1e-20, "MaxCGstep": 3000})
    op = orc.make_op(dims, kappa=0.11, r=r)
    src = orc.gaussian_field(dims, orc.WILSON, seed=18)
    x.from_host(src)
    y = q.similar(x)
    for A, mode in ((D, orc.D), (q.adjoint(D), orc.DDAG), (q.DdagD(D), orc.DDAGD)):
        q.mul_(y, A, x)
        want = orc.apply(op, orc.WILSON, mode, Uh, src)
        assert np.abs(y.to_host() - want).max() / np.abs(want).max() < 1e-13, mode
    sol = q.similar(x)
    q.clear_fermion_(sol)
    info = q.solve_DinvX_(sol, q.DdagD(D), x)
    ref = orc.cg(op, orc.WILSON, Uh, src, eps=1e-20)
    assert ref["converged"] and info["iters"] == ref["iters"]
    assert np.abs(sol.to_host() - ref["x"]).max() / np.abs(ref["x"]).max() < 1e-10
    q.clear_fermion_(sol)
    info = q.solve_DinvX_(sol, D, x)
    ref = orc.cgnr(op, orc.WILSON, Uh, src, eps=1e-20)
    assert ref["converged"] and info["iters"] == ref["iters"]
    ys = [q.similar(x), q.similar(x)]
    q.mul_multi_(ys, D, [x, sol])
    q.mul_(y, D, x)
    assert np.array_equal(ys[0].to_host(), y.to_host())
    yh = np.zeros_like(src)
    q.mul_host_(yh, D, src)
    assert np.array_equal(yh, y.to_host())
    with pytest.raises(q.LqcdError, match="r = 1 only"):
        q.shiftedcg_([q.similar(x)], D, x, [0.1])


def _env_dims(name, default):
    import os
    return tuple(int(v) for v in os.environ.get(name, default).split("x"))


def test_config1_16_4_evenodd_cg():
    """BASELINE config 1 at full size (16^4 Wilson, even-odd preconditioned CG on one GPU): the Schur-preconditioned solve returns the
    solution of the full system (true residual recomputed with the verified operator), in fewer iterations than the plain solve"""
    import lqcd_b200 as q
    dims = _env_dims("LQCD_TEST_CONFIG1_DIMS", "16x16x16x16")            # (the emulated pre-flight shrinks it)
    ctx = q.get_context(dims)
    U = q.Initialize_Gaugefields(3, 0, *dims, condition="cold")
    b = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    params = {"Dirac_operator": "Wilson", "κ": 0.125, "eps_CG": 1e-16, "MaxCGstep": 3000, "boundarycondition": [1, 1, 1, -1]}
    D = q.Dirac_operator(U, b, params)
    Deo = q.Dirac_operator(U, b, dict(params, evenodd=True))
    ctx.call("lqcd_gauge_random", 111, 0.3)
    q.gauss_distribution_fermion_(b, 5)
    x_full, x_eo, chk = q.similar(b), q.similar(b), q.similar(b)
    q.clear_fermion_(x_full); q.clear_fermion_(x_eo)
    i_full = q.solve_DinvX_(x_full, D, b)
    i_eo = q.solve_DinvX_(x_eo, Deo, b)
    q.mul_(chk, D, x_eo)
    q.add_(chk, -1.0, b)
    assert q.dot(chk, chk).real < 1e-14
    q.add_(x_eo, -1.0, x_full)
    assert q.dot(x_eo, x_eo).real < 1e-12 * q.dot(x_full, x_full).real
    assert i_eo["iters"] < i_full["iters"]


def test_config2_24_4_staggered_nf4_trajectory():
    """BASELINE config 2 at full size (24^4 staggered Nf = 4 HMC on one GPU, test_staggered.toml integrator): device-resident
    trajectories with the even-site action on half fields conserve H at O(dtau^2), keep the links in SU(3), and the half-field
    and full-lattice solvers give the same Delta H"""
    import lqcd_b200 as q
    dims = _env_dims("LQCD_TEST_CONFIG2_DIMS", "24x24x24x24")
    ctx = q.get_context(dims)
    U0 = q.Initialize_Gaugefields(3, 0, *dims, condition="cold")
    x = q.Initialize_pseudofermion_fields(U0[0], "staggered")
    D = q.Dirac_operator(U0, x, {"Dirac_operator": "staggered", "mass": 0.5, "eps_CG": 1e-18, "MaxCGstep": 3000, "boundarycondition": [1, 1, 1, -1]})
    ctx.call("lqcd_gauge_random", 111, 0.2)
    warm = q.get_links(ctx)                                   # synthetic warm start (there is no thermalised 24^4 fixture)
    dH = {}
    for half in (True, False):
        for dtau, steps in ((0.02, 4), (0.01, 8)):
            U = q.gaugefields_from_array(warm.copy())
            fa = q.FermiAction(D, {"Nf": 4, "half_field_solver": half})
            acc, d, info = q.hmc_update_(U, 5.7, dtau, steps, fa=fa, rng=np.random.default_rng(4))
            dH[(half, dtau)] = d
            assert info["cg_iters"] > 0
            if acc:
                M = U.data.reshape(-1, 3, 3)[:: max(1, U.data.size // 9 // 4096)]
                assert np.abs(np.einsum("nij,nkj->nik", M, M.conj()) - np.eye(3)).max() < 1e-9
    assert 2.5 < dH[(True, 0.02)] / dH[(True, 0.01)] < 6.0, dH
    assert abs(dH[(True, 0.02)] - dH[(False, 0.02)]) < 1e-6 * max(1.0, abs(dH[(False, 0.02)])), dH


@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (8, 4, 6, 4)])
def test_clover_force_matches_oracle(dims):
    """calc_UdSfdU! for the Wilson-clover action: hopping part + clover-term part (csrc/clover_force.cu, gather form) against the
    oracle's scatter form, which is pinned by finite differences (tests/test_clover.py)"""
    q, Uh, U, x, D, op, clov = _clover_setup(dims, 0.11)
    fa = q.FermiAction(D, {})
    eta_h = orc.gaussian_field(dims, orc.WILSON, seed=29)
    eta = q.similar(x).from_host(eta_h)
    F = np.zeros_like(Uh)
    info = q.calc_UdSfdU_(F, fa, U, eta)
    ref = orc.cg(op, orc.WILSON, Uh, eta_h, eps=1e-20)
    assert info["iters"] == ref["iters"]
    Fr = orc.force(op, orc.WILSON, Uh, ref["x"], orc.apply(op, orc.WILSON, orc.D, Uh, ref["x"]))
    op0 = orc.make_op(dims, kappa=0.11)
    Fhop = orc.force(op0, orc.WILSON, Uh, ref["x"], orc.apply(op, orc.WILSON, orc.D, Uh, ref["x"]))
    assert np.abs(Fr - Fhop).max() > 1e-3 * np.abs(Fr).max()              # the clover part matters in this check
    assert np.abs(F - Fr).max() < 1e-9 * np.abs(Fr).max()


def test_wilson_clover_hmc_trajectory():
    """device-resident Sexton-Weingarten trajectories with the Wilson-clover action (the reference's disabled test_wilsonclover.jl
    setup: kappa = 0.141139, c_SW = 1.5612): Delta H = O(dtau^2) -- the relative weight of the clover-term force is right"""
    import lqcd_b200 as q
    dims = (4, 4, 4, 4)
    Uh = orc.random_su3(dims, seed=37, eps=0.3)
    dH = {}
    for dtau, steps in ((0.04, 4), (0.02, 8)):
        U = q.gaugefields_from_array(Uh.copy())
        x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
        D = q.Dirac_operator(U, x, {"Dirac_operator": "WilsonClover", "κ": 0.12, "Clover_coefficient": CSW, "eps_CG": 1e-20, "MaxCGstep": 3000,
                                    "boundarycondition": [1, 1, 1, -1]})
        fa = q.FermiAction(D, {})
        acc, d, info = q.hmc_update_(U, 5.7, dtau, steps, fa=fa, SextonWeingargten=True, Nsw=4, rng=np.random.default_rng(8))
        dH[dtau] = d
        assert info["cg_iters"] > 0
    assert abs(dH[0.04]) < 1.0 and 2.5 < dH[0.04] / dH[0.02] < 6.0, dH


@pytest.mark.parametrize("nrhs", [2, 5, 12])
def test_multi_rhs_wilson_clover(nrhs):
    """the multi-RHS Wilson kernel with the clover term in its epilogue: bit-identical to the single-RHS Wilson-clover kernel;
    batched CGNR with the same per-source iteration counts as single solves and as the oracle"""
    dims = (8, 4, 4, 4)
    q, Uh, U, x, D, op, clov = _clover_setup(dims, 0.12)
    srcs = [orc.gaussian_field(dims, orc.WILSON, seed=300 + j) for j in range(nrhs)]
    xs = [q.similar(x).from_host(s) for s in srcs]
    ys = [q.similar(x) for _ in range(nrhs)]
    y1 = q.similar(x)
    for A, mode in ((D, orc.D), (q.adjoint(D), orc.DDAG), (q.DdagD(D), orc.DDAGD)):
        q.mul_multi_(ys, A, xs)
        for j in range(nrhs):
            q.mul_(y1, A, xs[j])
            assert np.array_equal(ys[j].to_host(), y1.to_host()), (mode, j)
        want = orc.apply(op, orc.WILSON, mode, Uh, srcs[0])
        assert np.abs(ys[0].to_host() - want).max() / np.abs(want).max() < 1e-13
    for y in ys:
        q.clear_fermion_(y)
    infos = q.solve_DinvX_multi_(ys[:3], D, xs[:3]) if nrhs >= 3 else q.solve_DinvX_multi_(ys, D, xs)
    for j, info in enumerate(infos):
        ref = orc.cgnr(op, orc.WILSON, Uh, srcs[j], eps=1e-20)
        assert ref["converged"] and info["iters"] == ref["iters"]
        assert np.abs(ys[j].to_host() - ref["x"]).max() / np.abs(ref["x"]).max() < 1e-10

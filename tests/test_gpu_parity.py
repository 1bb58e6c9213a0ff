"""
GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI (ctypes
mirror of the Julia shim), against the CPU oracle on the same inputs.

Tolerances (fp64): one operator application <= 1e-13 relative to max|y| (different but equivalent
summation order: spin-projected vs the reference's un-projected form); solver iteration counts identical
to the oracle; solutions within 1e-10 relative (north_star: residuals within 1e-12 relative).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import lqcd_b200 as q
from oracle import oracle as orc

KAPPA = 0.141139
BC = [1, 1, 1, -1]


def wparams(kappa=KAPPA, **kw):
    p = {"Dirac_operator": "Wilson", "κ": kappa, "r": 1.0, "eps_CG": 1e-19, "MaxCGstep": 3000,
         "verbose_level": 1, "boundarycondition": BC}
    p.update(kw)
    return p


def sparams(mass=0.5, **kw):
    p = {"Dirac_operator": "staggered", "mass": mass, "eps_CG": 1e-19, "MaxCGstep": 3000,
         "verbose_level": 1, "boundarycondition": BC}
    p.update(kw)
    return p


def relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.fixture(scope="module")
def Uw(golden_dir):
    return np.load(golden_dir / "wilson_4444.npy")


@pytest.fixture(scope="module")
def Us(golden_dir):
    return np.load(golden_dir / "staggered_4444.npy")


def test_layout_roundtrip_and_plaquette(Uw):
    U = q.gaugefields_from_array(Uw)
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, wparams())
    import ctypes as C
    p = C.c_double()
    D.ctx.call("lqcd_gauge_plaquette", C.byref(p))
    assert abs(p.value - 0.565800226845) < 1e-11
    back = np.zeros_like(Uw)
    ptrs = (C.c_void_p * 4)(*[back[mu].ctypes.data for mu in range(4)])
    D.ctx.call("lqcd_gauge_download", ptrs, 3, 0)
    assert np.array_equal(back, Uw)
    src = orc.gaussian_field((4, 4, 4, 4), orc.WILSON, seed=3)
    assert np.array_equal(x.from_host(src).to_host(), src)
    # wing (Nwing = 1, test/test_wilson.toml:14) upload strips the halo
    Uwing = np.zeros((4, 6, 6, 6, 6, 3, 3), dtype=complex)
    Uwing[:, 1:-1, 1:-1, 1:-1, 1:-1] = Uw
    ptrs = (C.c_void_p * 4)(*[Uwing[mu].ctypes.data for mu in range(4)])
    D.ctx.call("lqcd_gauge_upload", ptrs, 3, 1)
    D.ctx.call("lqcd_gauge_plaquette", C.byref(p))
    assert abs(p.value - 0.565800226845) < 1e-11


@pytest.mark.parametrize("mode", ["D", "Ddag", "DdagD"])
def test_wilson_dslash_fixture(Uw, mode):
    dims = (4, 4, 4, 4)
    U = q.gaugefields_from_array(Uw)
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, wparams())
    src = orc.gaussian_field(dims, orc.WILSON, seed=11)
    x.from_host(src)
    y = q.similar(x)
    A = {"D": D, "Ddag": q.adjoint(D), "DdagD": q.DdagD(D)}[mode]
    q.mul_(y, A, x)
    op = orc.make_op(dims, kappa=KAPPA)
    want = orc.apply(op, orc.WILSON, {"D": orc.D, "Ddag": orc.DDAG, "DdagD": orc.DDAGD}[mode], Uw, src)
    assert relerr(y.to_host(), want) < 1e-13


@pytest.mark.parametrize("mode", ["D", "Ddag", "DdagD"])
def test_staggered_dslash_fixture(Us, mode):
    dims = (4, 4, 4, 4)
    U = q.gaugefields_from_array(Us)
    x = q.Initialize_pseudofermion_fields(U[0], "staggered")
    D = q.Dirac_operator(U, x, sparams())
    src = orc.gaussian_field(dims, orc.STAGGERED, seed=12)
    x.from_host(src)
    y = q.similar(x)
    A = {"D": D, "Ddag": q.adjoint(D), "DdagD": q.DdagD(D)}[mode]
    q.mul_(y, A, x)
    op = orc.make_op(dims, mass=0.5)
    want = orc.apply(op, orc.STAGGERED, {"D": orc.D, "Ddag": orc.DDAG, "DdagD": orc.DDAGD}[mode], Us, src)
    assert relerr(y.to_host(), want) < 1e-13


@pytest.mark.parametrize("dims", [(8, 4, 6, 4), (16, 8, 4, 4), (4, 4, 2, 2), (32, 4, 4, 4), (6, 8, 4, 4), (64, 2, 2, 4)])
@pytest.mark.parametrize("kind", ["Wilson", "staggered"])
def test_dslash_odd_shapes(dims, kind):
    """ragged / non-cubic lattices incl. the Domainwall fixture's 4*4*2*2 shape, X > 32 and the irregular
    block lattice (6*8*4*4): periodic x,y,z + antiperiodic t, synthetic hot links."""
    Uh = orc.random_su3(dims, seed=5)
    U = q.gaugefields_from_array(Uh)
    x = q.Initialize_pseudofermion_fields(U[0], kind)
    k = orc.WILSON if kind == "Wilson" else orc.STAGGERED
    D = q.Dirac_operator(U, x, wparams(0.12) if kind == "Wilson" else sparams(0.3))
    src = orc.gaussian_field(dims, k, seed=13)
    x.from_host(src)
    y = q.similar(x)
    op = orc.make_op(dims, kappa=0.12, mass=0.3)
    for A, m in ((D, orc.D), (q.adjoint(D), orc.DDAG)):
        q.mul_(y, A, x)
        assert relerr(y.to_host(), orc.apply(op, k, m, Uh, src)) < 1e-13


def test_boundary_conditions_periodic():
    dims = (4, 4, 4, 8)
    Uh = orc.random_su3(dims, seed=6)
    U = q.gaugefields_from_array(Uh)
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, wparams(0.1, boundarycondition=[1, -1, 1, 1]))
    src = orc.gaussian_field(dims, orc.WILSON, seed=14)
    x.from_host(src)
    y = q.similar(x)
    q.mul_(y, D, x)
    op = orc.make_op(dims, kappa=0.1, bc=(1, -1, 1, 1))
    assert relerr(y.to_host(), orc.apply(op, orc.WILSON, orc.D, Uh, src)) < 1e-13


def test_blas(Uw):
    dims = (4, 4, 4, 4)
    U = q.gaugefields_from_array(Uw)
    a = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    b = q.similar(a)
    ha, hb = orc.gaussian_field(dims, orc.WILSON, seed=1), orc.gaussian_field(dims, orc.WILSON, seed=2)
    a.from_host(ha); b.from_host(hb)
    d = q.dot(a, b)
    assert abs(d - np.vdot(ha, hb)) < 1e-11
    q.add_(b, 0.3 - 0.2j, a)
    assert relerr(b.to_host(), hb + (0.3 - 0.2j) * ha) < 1e-15
    q.add_xpby_(1.5 + 0.5j, b, a)
    assert relerr(b.to_host(), (1.5 + 0.5j) * (hb + (0.3 - 0.2j) * ha) + ha) < 1e-15
    q.clear_fermion_(b)
    assert np.abs(b.to_host()).max() == 0
    q.setindex_global_(b, 1, 2, 1, 1, 1, 1, 3)
    h = b.to_host()
    assert h[2, 0, 0, 0, 0, 1] == 1 and np.abs(h).sum() == 1
    assert b[2, 1, 1, 1, 1, 3] == 1


@pytest.mark.parametrize("source", ["point", "gauss"])
@pytest.mark.parametrize("kappa", [KAPPA, 0.12])
def test_cg_matches_oracle_iterations(Uw, source, kappa):
    """solve_DinvX!(y, DdagD, b): iteration count identical to the oracle, solution within 1e-10."""
    dims = (4, 4, 4, 4)
    U = q.gaugefields_from_array(Uw)
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, wparams(kappa))
    src = orc.point_source(dims, orc.WILSON, 0, 0) if source == "point" else orc.gaussian_field(dims, orc.WILSON, seed=112)
    x.from_host(src)
    sol = q.similar(x)
    q.clear_fermion_(sol)
    info = q.solve_DinvX_(sol, q.DdagD(D), x, history=True)
    op = orc.make_op(dims, kappa=kappa)
    ref = orc.cg(op, orc.WILSON, Uw, src, hist=True)
    assert info["iters"] == ref["iters"]
    assert relerr(sol.to_host(), ref["x"]) < 1e-10
    n = len(ref["hist"])
    assert np.allclose(info["hist"][: n - 5], ref["hist"][: n - 5], rtol=1e-6)
    true_r = src - orc.apply(op, orc.WILSON, orc.DDAGD, Uw, sol.to_host())
    assert np.vdot(true_r, true_r).real < 1e-18


@pytest.mark.parametrize("method", ["bicg", "bicgstab"])
def test_solve_D_matches_oracle(Uw, method):
    """solve_DinvX!(y, D, b) as the pion-correlator code calls it (point source, colour 1 spin 1)."""
    dims = (4, 4, 4, 4)
    U = q.gaugefields_from_array(Uw)
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, wparams(0.12, method_CG=method))
    q.clear_fermion_(x)
    if method == "bicg":
        q.setindex_global_(x, 1, 1, 1, 1, 1, 1, 1)
    else:       # BiCGStab with r0~ = r0 breaks down on a point source (rho_1 = 0 for r = 1)
        x.from_host(orc.gaussian_field(dims, orc.WILSON, seed=21))
    src = x.to_host()
    sol = q.similar(x)
    q.clear_fermion_(sol)
    info = q.solve_DinvX_(sol, D, x)
    op = orc.make_op(dims, kappa=0.12)
    ref = (orc.cgnr if method == "bicg" else orc.bicgstab)(op, orc.WILSON, Uw, src)
    assert abs(info["iters"] - ref["iters"]) <= (0 if method == "bicg" else 2)
    assert relerr(sol.to_host(), ref["x"]) < 1e-9
    # adjoint target
    q.clear_fermion_(sol)
    q.solve_DinvX_(sol, q.adjoint(D), x)
    true_r = src - orc.apply(op, orc.WILSON, orc.DDAG, Uw, sol.to_host())
    assert np.vdot(true_r, true_r).real < 1e-18


def test_staggered_cg(Us):
    dims = (4, 4, 4, 4)
    U = q.gaugefields_from_array(Us)
    x = q.Initialize_pseudofermion_fields(U[0], "staggered")
    D = q.Dirac_operator(U, x, sparams(0.5))
    src = orc.gaussian_field(dims, orc.STAGGERED, seed=112)
    x.from_host(src)
    sol = q.similar(x)
    q.clear_fermion_(sol)
    info = q.solve_DinvX_(sol, q.DdagD(D), x)
    op = orc.make_op(dims, mass=0.5)
    ref = orc.cg(op, orc.STAGGERED, Us, src)
    assert info["iters"] == ref["iters"]
    assert relerr(sol.to_host(), ref["x"]) < 1e-10


def test_multishift(Us):
    dims = (4, 4, 4, 4)
    U = q.gaugefields_from_array(Us)
    x = q.Initialize_pseudofermion_fields(U[0], "staggered")
    D = q.Dirac_operator(U, x, sparams(0.1, eps_CG=1e-22))
    src = orc.gaussian_field(dims, orc.STAGGERED, seed=31)
    x.from_host(src)
    shifts = [0.0, 0.05, 0.4, 2.0]
    ys = [q.similar(x) for _ in shifts]
    info = q.shiftedcg_(ys, D, x, shifts)
    op = orc.make_op(dims, mass=0.1)
    ref = orc.mscg(op, orc.STAGGERED, Us, src, shifts, eps=1e-22)
    assert info["iters"] == ref["iters"]
    for y, xr in zip(ys, ref["xs"]):
        assert relerr(y.to_host(), xr) < 1e-10


def test_bicgstab_breakdown_is_reported(Uw):
    """the oracle and the GPU agree that BiCGStab breaks down (NaN) on a point source; surfaced as an error."""
    U = q.gaugefields_from_array(Uw)
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, wparams(0.12, method_CG="bicgstab"))
    q.setindex_global_(x, 1, 1, 1, 1, 1, 1, 1)
    sol = q.similar(x)
    q.clear_fermion_(sol)
    with pytest.raises(q.NotConverged) as e:
        q.solve_DinvX_(sol, D, x)
    assert "breakdown" in str(e.value)
    ref = orc.bicgstab(orc.make_op((4, 4, 4, 4), kappa=0.12), orc.WILSON, Uw, x.to_host())
    assert not ref["converged"]


def test_nonconvergence_is_an_error(Uw):
    U = q.gaugefields_from_array(Uw)
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, wparams(MaxCGstep=5))
    q.gauss_distribution_fermion_(x, 5)
    sol = q.similar(x)
    with pytest.raises(q.NotConverged):
        q.solve_DinvX_(sol, q.DdagD(D), x)


def test_initial_guess_is_used(Uw):
    dims = (4, 4, 4, 4)
    U = q.gaugefields_from_array(Uw)
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, wparams(0.12))
    src = orc.gaussian_field(dims, orc.WILSON, seed=9)
    x.from_host(src)
    sol = q.similar(x)
    q.clear_fermion_(sol)
    q.solve_DinvX_(sol, q.DdagD(D), x)
    info = q.solve_DinvX_(sol, q.DdagD(D), x)      # second call starts from the solution
    assert info["iters"] <= 2


def test_16_4_size_independent_properties():
    """BASELINE config 2 size (16^4): gamma5-hermiticity, <a, D b> = <D^dag a, b>, CG true residual --
    properties that need no oracle run."""
    import ctypes as C
    dims = (16, 16, 16, 16)
    ctx = q.get_context(dims)
    ctx.call("lqcd_gauge_random", 111, -1.0)
    U = q.Initialize_Gaugefields(3, 0, *dims, condition="cold")
    a = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, a, wparams(0.12))
    ctx.call("lqcd_gauge_random", 111, -1.0)            # overwrite the cold links on the device
    b, Db, Dda = q.similar(a), q.similar(a), q.similar(a)
    q.gauss_distribution_fermion_(a, 1); q.gauss_distribution_fermion_(b, 2)
    q.mul_(Db, D, b); q.mul_(Dda, q.adjoint(D), a)
    l, r = q.dot(a, Db), q.dot(Dda, b)
    assert abs(l - r) < 1e-9 * abs(l)
    # solve on a warm field and recompute the true residual on the device
    ctx.call("lqcd_gauge_random", 111, 0.3)
    sol, chk = q.similar(a), q.similar(a)
    q.clear_fermion_(sol)
    info = q.solve_DinvX_(sol, q.DdagD(D), b)
    q.mul_(chk, q.DdagD(D), sol)
    q.add_(chk, -1.0, b)
    assert q.dot(chk, chk).real < 1e-17
    assert 5 < info["iters"] < 3000


@pytest.mark.parametrize("kind", ["Wilson", "staggered"])
def test_fermion_force_matches_oracle(Uw, Us, kind):
    """calc_UdSfdU!(UdSfdU, fermi_action, U, eta) (AbstractMD.jl:129): device CG + outer products vs the oracle
    (whose force is itself pinned by finite differences of S_f in tests/test_oracle.py)."""
    dims = (4, 4, 4, 4)
    Uh = Uw if kind == "Wilson" else Us
    k = orc.WILSON if kind == "Wilson" else orc.STAGGERED
    U = q.gaugefields_from_array(Uh)
    x = q.Initialize_pseudofermion_fields(U[0], kind)
    D = q.Dirac_operator(U, x, wparams(0.12, eps_CG=1e-24) if kind == "Wilson" else sparams(0.5, eps_CG=1e-24))
    fa = q.FermiAction(D, {"Nf": 2 if kind == "Wilson" else 8})       # staggered Nf = 2 would be RHMC (README.md:132)
    eta = q.similar(x)
    phi = orc.gaussian_field(dims, k, seed=41)
    eta.from_host(phi)
    F = np.zeros_like(Uh)
    info = q.calc_UdSfdU_(F, fa, U, eta)
    op = orc.make_op(dims, kappa=0.12, mass=0.5)
    ref = orc.cg(op, k, Uh, phi, eps=1e-24)
    Y = orc.apply(op, k, orc.D, Uh, ref["x"])
    Fref = orc.force(op, k, Uh, ref["x"], Y)
    assert info["iters"] == ref["iters"]
    assert relerr(F, Fref) < 1e-10
    assert abs(info["action"] - np.vdot(phi, ref["x"]).real) < 1e-9 * abs(info["action"])


def test_pseudofermion_heatbath_identity(Uw):
    """standardMD.jl:95-96 + standardHMC.jl:54,69: eta = D^dag xi  =>  S_f = eta^dag (D^dag D)^-1 eta = xi^dag xi."""
    U = q.gaugefields_from_array(Uw)
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, wparams(0.125))
    fa = q.FermiAction(D, {})
    xi, eta = q.similar(x), q.similar(x)
    q.gauss_sampling_in_action_(xi, U, fa, seed=7)
    q.sample_pseudofermions_(eta, U, fa, xi)
    Sold = q.dot(xi, xi).real
    Snew = q.evaluate_FermiAction(fa, U, eta)
    assert abs(Sold - Snew) < 1e-8 * Sold
    assert abs(Sold - 12 * 256) < 0.1 * 12 * 256          # <xi^dag xi> = number of complex components


def test_mask_parity(Us):
    """lqcd_fermion_mask_parity: even-site restriction (building block of the staggered Nf=4 fields); D^dag D keeps it."""
    dims = (4, 4, 4, 4)
    U = q.gaugefields_from_array(Us)
    x = q.Initialize_pseudofermion_fields(U[0], "staggered")
    D = q.Dirac_operator(U, x, sparams(0.5))
    src = orc.gaussian_field(dims, orc.STAGGERED, seed=5)
    x.from_host(src)
    q.mask_parity_(x, 0)
    t, z, y, xx = np.meshgrid(*[np.arange(4)] * 4, indexing="ij")
    even = ((xx + y + z + t) % 2 == 0)[..., None]
    assert np.array_equal(x.to_host(), src * even)
    out = q.similar(x)
    q.mul_(out, q.DdagD(D), x)
    assert np.abs(out.to_host() * (~even)).max() == 0.0
    sol = q.similar(x)
    q.clear_fermion_(sol)
    info = q.solve_DinvX_(sol, q.DdagD(D), x)
    ref = orc.cg(orc.make_op(dims, mass=0.5), orc.STAGGERED, Us, np.ascontiguousarray(src * even))
    assert info["iters"] == ref["iters"]
    assert np.abs(sol.to_host() * (~even)).max() == 0.0


def test_verbose_level_3_prints_the_residual_history(Uw, capsys):
    """universe.jl:133 passes verbose_level to the operator; at level 3 upstream prints '<i>-th eps: <r.r>' every CG step"""
    U = q.gaugefields_from_array(Uw)
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, wparams(0.12, verbose_level=3))
    x.from_host(orc.gaussian_field((4, 4, 4, 4), orc.WILSON, seed=9))
    sol = q.similar(x)
    q.clear_fermion_(sol)
    info = q.solve_DinvX_(sol, q.DdagD(D), x)
    lines = [ln for ln in capsys.readouterr().out.splitlines() if "-th eps:" in ln]
    assert len(lines) == info["iters"] + 1 and lines[0].startswith("0-th eps:")
    assert float(lines[-1].split(":")[1]) < 1e-19

"""Gauge-sector molecular dynamics: the steps of src/md/AbstractMD.jl:78-135 and the leapfrog
integrators of src/md/standardMD.jl:125-165.  CPU: the oracle's restatement is pinned by what an integrator must satisfy
whatever the generator normalisation -- energy conservation at O(dtau^2) (a wrong force factor breaks the scaling) and
reversibility.  GPU (hardware-verified since the round-1 driver run): device steps and a
whole Sexton-Weingarten trajectory against the oracle composed step by step."""
import numpy as np
import pytest

from oracle import oracle as orc

DIMS = (4, 4, 4, 4)
BETA = 5.7            # test/test_wilson.toml
KAPPA = 0.141139


@pytest.fixture(scope="module")
def Uw(golden_dir):
    return np.load(golden_dir / "wilson_4444.npy")


def _traj(U, P, dtau, n, nsw=0, fermion=None):
    """runMD_QPQ! / runMD_QPQ_sw! (standardMD.jl:125-165) composed from the oracle's steps; fermion = (op, kind, eta)"""
    U, P = U.copy(), P.copy()

    def pf(eps):
        op, kind, eta = fermion[:3]
        if kind == orc.WILSON and op.csw != 0.0:
            orc.clover_build(op, U)                 # the clover term follows the links
        if len(fermion) == 4:              # RHMC: fermion[3] = rational approximation of the action (alpha_j, beta_j)
            ra = fermion[3]
            xs = orc.mscg(op, kind, U, eta, list(ra.beta), eps=1e-22, maxsteps=5000)["xs"]
            F = sum(a * orc.force(op, kind, U, X, orc.apply(op, kind, orc.D, U, X)) for a, X in zip(ra.alpha, xs))
            orc.md_update_p_force(DIMS, P, F, eps)
            return
        X = orc.cg(op, kind, U, eta, eps=1e-22)["x"]
        orc.md_update_p_force(DIMS, P, orc.force(op, kind, U, X, orc.apply(op, kind, orc.D, U, X)), eps)

    for _ in range(n):
        if nsw == 0:
            orc.md_update_u(DIMS, U, P, 0.5 * dtau)
            orc.md_update_p_gauge(DIMS, P, U, dtau, BETA)
            if fermion:
                pf(dtau)
            orc.md_update_u(DIMS, U, P, 0.5 * dtau)
        else:
            for half in range(2):
                for _ in range(nsw // 2):
                    orc.md_update_u(DIMS, U, P, 0.5 * dtau / nsw)
                    orc.md_update_p_gauge(DIMS, P, U, dtau / nsw, BETA)
                    orc.md_update_u(DIMS, U, P, 0.5 * dtau / nsw)
                if half == 0 and fermion:
                    pf(dtau)
    return U, P


def _H(U, P, fermion=None):
    H = orc.md_kinetic(DIMS, P) + orc.md_gauge_action(DIMS, U, BETA)
    if fermion and len(fermion) == 4:
        op, kind, eta, ra = fermion
        xs = orc.mscg(op, kind, U, eta, list(ra.beta), eps=1e-22, maxsteps=5000)["xs"]
        H += np.vdot(eta, ra.alpha0 * eta + sum(a * X for a, X in zip(ra.alpha, xs))).real
    elif fermion:
        op, kind, eta = fermion
        if kind == orc.WILSON and op.csw != 0.0:
            orc.clover_build(op, U)
        H += np.vdot(eta, orc.cg(op, kind, U, eta, eps=1e-22)["x"]).real
    return H


def test_wilson_clover_leapfrog_energy_scaling(Uw):
    """Wilson-clover pseudofermions (the reference's disabled test_wilsonclover.jl physics: c_SW = 1.5612): with the hopping AND
    the clover-term force the leapfrog conserves H at O(dtau^2) on the oracle"""
    op = orc.make_op(DIMS, kappa=0.12, csw=1.5612)
    orc.clover_build(op, Uw)
    xi = orc.gaussian_field(DIMS, orc.WILSON, seed=9)
    eta = orc.apply(op, orc.WILSON, orc.DDAG, Uw, xi)
    P0 = orc.md_momenta(DIMS, seed=4)
    f = (op, orc.WILSON, eta)
    H0 = _H(Uw, P0, f)
    assert abs(H0 - (orc.md_kinetic(DIMS, P0) + orc.md_gauge_action(DIMS, Uw, BETA) + np.vdot(xi, xi).real)) < 1e-8 * abs(H0)
    dH = []
    for dtau, n in ((0.05, 6), (0.025, 12), (0.0125, 24)):
        U1, P1 = _traj(Uw, P0, dtau, n, nsw=0, fermion=f)
        dH.append(_H(U1, P1, f) - H0)
    print("dH", dH)
    assert abs(dH[0]) < 1.0 and 2.5 < dH[0] / dH[1] < 6.0 and 3.0 < dH[1] / dH[2] < 5.0


def _rational_nf2():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "latticeqcd.jl_b200"))
    from lqcd_b200 import rhmc
    return rhmc.rational_approx(-2 / 8.0, 12, 0.22, 17.0)


def test_rhmc_leapfrog_energy_scaling(golden_dir):
    """RHMC trajectory (test/test_Nf2.toml physics on its fixture): with the force sum_j alpha_j F(X_j, Y_j) the leapfrog conserves
    H = K + S_g + eta^dag r(D^dag D) eta at O(dtau^2) -- the relative weight of the rational force against the gauge force"""
    U0 = np.load(golden_dir / "staggered_nf2_4444.npy")
    op = orc.make_op(DIMS, mass=0.5)
    eta = orc.gaussian_field(DIMS, orc.STAGGERED, seed=19)
    f = (op, orc.STAGGERED, eta, _rational_nf2())
    P0 = orc.md_momenta(DIMS, seed=6)
    H0 = _H(U0, P0, f)
    dH = []
    for dtau, n in ((0.1, 4), (0.05, 8), (0.025, 16)):
        U1, P1 = _traj(U0, P0, dtau, n, nsw=0, fermion=f)
        dH.append(_H(U1, P1, f) - H0)
    print("dH", dH)
    assert abs(dH[0]) < 1.0 and 2.5 < dH[0] / dH[1] < 6.0 and 3.0 < dH[1] / dH[2] < 5.0


def test_momenta_and_kinetic_term():
    P = orc.md_momenta(DIMS, seed=3)
    M = np.swapaxes(P, -1, -2)
    assert np.abs(M + np.conj(np.swapaxes(M, -1, -2))).max() < 1e-15          # anti-Hermitian
    assert np.abs(np.trace(M, axis1=-2, axis2=-1)).max() < 1e-15              # traceless
    K = orc.md_kinetic(DIMS, P)                                               # = sum a^2 / 2, 8 real a per link
    assert abs(K / (4 * 256 * 8 / 2) - 1.0) < 0.05


def test_quenched_leapfrog_energy_scaling_and_reversibility(Uw):
    P0 = orc.md_momenta(DIMS, seed=3)
    H0 = _H(Uw, P0)
    dH = []
    for dtau, n in ((0.1, 10), (0.05, 20), (0.025, 40)):
        U1, P1 = _traj(Uw, P0, dtau, n)
        dH.append(_H(U1, P1) - H0)
    assert abs(dH[0]) < 2.0 and 3.0 < dH[0] / dH[1] < 5.0 and 3.0 < dH[1] / dH[2] < 5.0      # O(dtau^2)
    U1, P1 = _traj(Uw, P0, 0.05, 20)
    U2, P2 = _traj(U1, -P1, 0.05, 20)
    assert np.abs(U2 - Uw).max() < 1e-12 and np.abs(P2 + P0).max() < 1e-12
    assert np.abs(np.einsum("...ij,...kj->...ik", U1, U1.conj()) - np.eye(3)).max() < 1e-9     # exp keeps SU(3)


@pytest.mark.parametrize("kind", [orc.WILSON, orc.STAGGERED])
def test_dynamical_leapfrog_energy_scaling(Uw, kind):
    """with the pseudofermion force (P_update_fermion!): the relative weight of gauge and fermion force must be right"""
    op = orc.make_op(DIMS, kappa=0.12, mass=0.5)
    xi = orc.gaussian_field(DIMS, kind, seed=9)
    eta = orc.apply(op, kind, orc.DDAG, Uw, xi)
    P0 = orc.md_momenta(DIMS, seed=4)
    f = (op, kind, eta)
    H0 = _H(Uw, P0, f)
    assert abs(H0 - (orc.md_kinetic(DIMS, P0) + orc.md_gauge_action(DIMS, Uw, BETA) + np.vdot(xi, xi).real)) < 1e-8 * abs(H0)
    dH = []
    for dtau, n in ((0.05, 10), (0.025, 20), (0.0125, 40)):          # asymptotic regime of the fermion force
        U1, P1 = _traj(Uw, P0, dtau, n, nsw=0, fermion=f)
        dH.append(_H(U1, P1, f) - H0)
    print("dH", dH)
    assert abs(dH[0]) < 1.0 and 2.5 < dH[0] / dH[1] < 6.0 and 3.0 < dH[1] / dH[2] < 5.0
    # Sexton-Weingarten nesting integrates the same Hamiltonian
    U2, P2 = _traj(Uw, P0, 0.05, 10, nsw=4, fermion=f)
    assert abs(_H(U2, P2, f) - H0) < 1.0


# ---- device ----------------------------------------------------------------------------------------------------------------


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (8, 4, 6, 4)])
def test_md_steps_match_oracle(dims):
    import lqcd_b200 as q
    U0 = orc.random_su3(dims, seed=31, eps=0.4)
    P0 = orc.md_momenta(dims, seed=32)
    U = q.gaugefields_from_array(U0.copy())
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, {"Dirac_operator": "Wilson", "κ": 0.12, "eps_CG": 1e-22, "MaxCGstep": 3000, "boundarycondition": [1, 1, 1, -1]})
    ctx = D.ctx
    q.set_momenta_(ctx, P0)
    assert np.array_equal(q.get_momenta(ctx), P0)
    assert abs(q.kinetic_energy(ctx) - orc.md_kinetic(dims, P0)) < 1e-10 * orc.md_kinetic(dims, P0)
    assert abs(q.gauge_action(ctx, BETA) - orc.md_gauge_action(dims, U0, BETA)) < 1e-10 * abs(orc.md_gauge_action(dims, U0, BETA))
    Ur, Pr = U0.copy(), P0.copy()
    q.P_update_(ctx, 0.07, BETA); orc.md_update_p_gauge(dims, Pr, Ur, 0.07, BETA)
    assert np.abs(q.get_momenta(ctx) - Pr).max() < 1e-13
    q.U_update_(ctx, 0.11); orc.md_update_u(dims, Ur, Pr, 0.11)
    assert np.abs(q.get_links(ctx) - Ur).max() < 1e-13
    op = orc.make_op(dims, kappa=0.12)
    eta_h = orc.gaussian_field(dims, orc.WILSON, seed=33)
    eta = q.similar(x).from_host(eta_h)
    its = q.P_update_fermion_(D, eta, 0.05)
    ref = orc.cg(op, orc.WILSON, Ur, eta_h, eps=1e-22)
    orc.md_update_p_force(dims, Pr, orc.force(op, orc.WILSON, Ur, ref["x"], orc.apply(op, orc.WILSON, orc.D, Ur, ref["x"])), 0.05)
    assert its == ref["iters"]
    assert np.abs(q.get_momenta(ctx) - Pr).max() < 1e-10
    # device Gaussian momenta: anti-Hermitian, traceless, <a^2> = 1
    q.gauss_distribution_momenta_(ctx, 7)
    M = np.swapaxes(q.get_momenta(ctx), -1, -2)
    assert np.abs(M + np.conj(np.swapaxes(M, -1, -2))).max() < 1e-15 and np.abs(np.trace(M, axis1=-2, axis2=-1)).max() < 1e-15
    nl = 4 * int(np.prod(dims))
    assert abs(q.kinetic_energy(ctx) / (nl * 4) - 1.0) < 0.05



@pytest.mark.gpu
def test_sw_trajectory_matches_oracle(Uw):
    """runMD_QPQ_sw! on the device == the oracle's steps composed in the same order (test/test_wilson.toml integrator with a
    shorter trajectory): links and momenta after the trajectory, and Delta H"""
    import lqcd_b200 as q
    U = q.gaugefields_from_array(Uw.copy())
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, {"Dirac_operator": "Wilson", "κ": KAPPA, "eps_CG": 1e-22, "MaxCGstep": 3000, "boundarycondition": [1, 1, 1, -1]})
    ctx = D.ctx
    op = orc.make_op(DIMS, kappa=KAPPA)
    xi = orc.gaussian_field(DIMS, orc.WILSON, seed=41)
    eta_h = orc.apply(op, orc.WILSON, orc.DDAG, Uw, xi)
    eta = q.similar(x).from_host(eta_h)
    P0 = orc.md_momenta(DIMS, seed=42)
    q.set_momenta_(ctx, P0)
    H0 = q.kinetic_energy(ctx) + q.gauge_action(ctx, BETA) + np.vdot(xi, xi).real
    its = q.runMD_(ctx, BETA, 0.05, 3, D, eta, SextonWeingargten=True, Nsw=10)
    assert its > 0
    Ur, Pr = _traj(Uw, P0, 0.05, 3, nsw=10, fermion=(op, orc.WILSON, eta_h))
    assert np.abs(q.get_links(ctx) - Ur).max() < 1e-9 and np.abs(q.get_momenta(ctx) - Pr).max() < 1e-9
    X = q.similar(x)
    q.clear_fermion_(X)
    q.solve_DinvX_(X, q.DdagD(D), eta)
    H1 = q.kinetic_energy(ctx) + q.gauge_action(ctx, BETA) + q.dot(eta, X).real
    assert abs((H1 - H0) - (_H(Ur, Pr, (op, orc.WILSON, eta_h)) - _H(Uw, P0, (op, orc.WILSON, eta_h)))) < 1e-6
    assert abs(H1 - H0) < 1.0



@pytest.mark.gpu
def test_hmc_update_accepts_and_moves_links(Uw):
    """update!(StandardHMC, U) with the MD on the device (standardHMC.jl:41-91): a short trajectory is accepted with
    |Delta H| << 1, the links change, stay in SU(3), and the plaquette stays in the reference's 10 % window around the
    value the reference's own test expects after its trajectories (test/runtests.jl:89-99, debugplaqdata.txt:7)"""
    import lqcd_b200 as q
    U = q.gaugefields_from_array(Uw.copy())
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    D = q.Dirac_operator(U, x, {"Dirac_operator": "Wilson", "κ": KAPPA, "eps_CG": 1e-19, "MaxCGstep": 3000, "boundarycondition": [1, 1, 1, -1]})
    fa = q.FermiAction(D, {})
    rng = np.random.default_rng(5)
    acc, dH, info = q.hmc_update_(U, BETA, 0.05, 4, fa=fa, SextonWeingargten=True, Nsw=10, rng=rng)
    assert abs(dH) < 0.5 and info["cg_iters"] > 0
    if acc:
        assert np.abs(U.data - Uw).max() > 1e-3
        assert np.abs(np.einsum("...ij,...kj->...ik", U.data, U.data.conj()) - np.eye(3)).max() < 1e-9
    plaq = orc.plaquette(DIMS, U.data)
    assert abs(plaq - 0.5784043949012552) / 0.5784043949012552 < 0.1



@pytest.mark.gpu
def test_rhmc_trajectory_matches_oracle(golden_dir):
    """runMD! with the RHMC pseudofermion action on the device (lqcd_md_trajectory_rational: one multi-shift CG + accumulated outer
    products per fermion force) == the oracle's steps composed in the same order; then hmc_update_ with FermiAction(D, Nf = 2)"""
    import lqcd_b200 as q
    U0 = np.load(golden_dir / "staggered_nf2_4444.npy")
    U = q.gaugefields_from_array(U0.copy())
    x = q.Initialize_pseudofermion_fields(U[0], "staggered")
    D = q.Dirac_operator(U, x, {"Dirac_operator": "staggered", "mass": 0.5, "eps_CG": 1e-22, "MaxCGstep": 5000, "boundarycondition": [1, 1, 1, -1]})
    ctx = D.ctx
    op = orc.make_op(DIMS, mass=0.5)
    ra = _rational_nf2()
    eta_h = orc.gaussian_field(DIMS, orc.STAGGERED, seed=19)
    eta = q.similar(x).from_host(eta_h)
    P0 = orc.md_momenta(DIMS, seed=6)
    f = (op, orc.STAGGERED, eta_h, ra)
    # one fermion momentum update
    q.set_momenta_(ctx, P0)
    its = q.P_update_fermion_(D, eta, 0.05, rational=ra)
    Ur, Pr = U0.copy(), P0.copy()
    xs = orc.mscg(op, orc.STAGGERED, Ur, eta_h, list(ra.beta), eps=1e-22, maxsteps=5000)
    F = sum(a * orc.force(op, orc.STAGGERED, Ur, X, orc.apply(op, orc.STAGGERED, orc.D, Ur, X)) for a, X in zip(ra.alpha, xs["xs"]))
    orc.md_update_p_force(DIMS, Pr, F, 0.05)
    assert its == xs["iters"]
    assert np.abs(q.get_momenta(ctx) - Pr).max() < 1e-10
    # a Sexton-Weingarten trajectory
    q.set_momenta_(ctx, P0)
    its = q.runMD_(ctx, BETA, 0.05, 3, D, eta, SextonWeingargten=True, Nsw=4, rational=ra)
    assert its > 0
    Ur, Pr = _traj(U0, P0, 0.05, 3, nsw=4, fermion=f)
    assert np.abs(q.get_links(ctx) - Ur).max() < 1e-9 and np.abs(q.get_momenta(ctx) - Pr).max() < 1e-9
    # update!(StandardHMC, U) with the RHMC action built the way the wrapper does (universe.jl:106-110,138)
    U = q.gaugefields_from_array(U0.copy())
    D = q.Dirac_operator(U, x, {"Dirac_operator": "staggered", "mass": 0.5, "eps_CG": 1e-20, "MaxCGstep": 5000, "boundarycondition": [1, 1, 1, -1]})
    fa = q.FermiAction(D, {"Nf": 2, "rational_lambda_min": 0.22, "rational_lambda_max": 17.0})
    acc, dH, info = q.hmc_update_(U, BETA, 0.05, 4, fa=fa, rng=np.random.default_rng(3))
    assert abs(dH) < 0.5 and info["cg_iters"] > 0
    if acc:
        assert np.abs(U.data - U0).max() > 1e-3
        assert np.abs(np.einsum("...ij,...kj->...ik", U.data, U.data.conj()) - np.eye(3)).max() < 1e-9


class _AlwaysReject:
    """Metropolis draw of 2.0: exp(min(0, -dH)) <= 1 is never >= it -> every trajectory is rejected (seeds stay random)"""

    def __init__(self, seed):
        self._r = np.random.default_rng(seed)

    def integers(self, *a, **k):
        return self._r.integers(*a, **k)

    def random(self):
        return 2.0


@pytest.mark.gpu
def test_rejected_trajectory_and_second_operator_leave_each_operator_on_its_own_links(golden_dir):
    """All operators of a context share one device link buffer.  After a REJECTED trajectory the device holds the evolved links
    (hmc_update_ must restore U), and building a second operator with another U re-binds the buffer (the first operator must
    re-upload its own U on next use): mul! through D has to give the same bits before and after both events."""
    import lqcd_b200 as q
    Uh = np.load(golden_dir / "wilson_4444.npy")
    U = q.gaugefields_from_array(Uh.copy())
    x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
    params = {"Dirac_operator": "Wilson", "κ": 0.12, "eps_CG": 1e-20, "MaxCGstep": 3000, "boundarycondition": [1, 1, 1, -1]}
    D = q.Dirac_operator(U, x, params)
    fa = q.FermiAction(D, {})
    x.from_host(orc.gaussian_field((4, 4, 4, 4), orc.WILSON, seed=77))
    y = q.similar(x)
    q.mul_(y, D, x)
    before = y.to_host()
    acc, dH, info = q.hmc_update_(U, BETA, 0.05, 4, fa=fa, rng=_AlwaysReject(5))
    assert not acc and np.array_equal(U.data, Uh)
    assert np.array_equal(q.get_links(D.ctx), Uh)                 # device restored to the host configuration
    q.mul_(y, D, x)
    assert np.array_equal(y.to_host(), before)
    # a second operator on the same lattice with other links
    U2 = q.Initialize_Gaugefields(3, 0, 4, 4, 4, 4, condition="cold")
    D2 = q.Dirac_operator(U2, x, params)
    y2 = q.similar(x)
    q.mul_(y2, D2, x)
    assert not np.array_equal(y2.to_host(), before)
    q.mul_(y, D, x)                                               # D finds the buffer re-bound and uploads its own U again
    assert np.array_equal(y.to_host(), before)
    props, infos = q.calc_quark_propagators_point_source(D2, U=U)  # measurement mirrors re-bind like the reference: D = m.D(U)
    ref = orc.cgnr(orc.make_op((4, 4, 4, 4), kappa=0.12), orc.WILSON, Uh, orc.point_source((4, 4, 4, 4), orc.WILSON, 0, 0), eps=1e-20)
    assert infos[0]["iters"] == ref["iters"]

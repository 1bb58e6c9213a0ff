"""
Independent numpy restatement (np.roll based, whole-lattice array ops) of the Wilson and staggered
operators, used ONLY to cross-check oracle/lqcd_oracle.c at 4^4.  Deliberately shares no code with the
C oracle: different traversal (per-direction full-lattice passes like the reference's un-fused CPU
path, SURVEY.md 2.2 "Fused ops: None") and its own gamma tables.
"""
import numpy as np

G = np.zeros((4, 4, 4), dtype=complex)
G[0][0, 3] = -1j; G[0][1, 2] = -1j; G[0][2, 1] = 1j; G[0][3, 0] = 1j
G[1][0, 3] = -1;  G[1][1, 2] = 1;   G[1][2, 1] = 1;  G[1][3, 0] = -1
G[2][0, 2] = -1j; G[2][1, 3] = 1j;  G[2][2, 0] = 1j; G[2][3, 1] = -1j
G[3][0, 2] = -1;  G[3][1, 3] = -1;  G[3][2, 0] = -1; G[3][3, 1] = -1
G5 = np.diag([1, 1, -1, -1]).astype(complex)
AX = {0: 3, 1: 2, 2: 1, 3: 0}     # direction mu -> numpy axis of [t,z,y,x]


def links_mat(U):
    """host layout [mu,t,z,y,x,b,a] -> matrices [mu,t,z,y,x,a,b]"""
    return np.swapaxes(U, -1, -2)


def shift(f, mu, sign, bc, site_axes_offset=0):
    """f(n + sign*mu) with boundary phase bc[mu] on wrap.  site axes start at site_axes_offset."""
    ax = AX[mu] + site_axes_offset
    g = np.roll(f, -sign, axis=ax)
    if bc[mu] != 1:
        n = f.shape[ax]
        idx = [slice(None)] * f.ndim
        idx[ax] = n - 1 if sign > 0 else 0
        g[tuple(idx)] *= bc[mu]
    return g


def wilson(U, psi, kappa, r=1.0, bc=(1, 1, 1, -1), dagger=False):
    """psi [alpha,t,z,y,x,c];  y = x - kappa sum_mu [(r-g)U x(n+mu) + (r+g)U^dag(n-mu) x(n-mu)]"""
    M = links_mat(U)
    acc = np.zeros_like(psi)
    one = np.eye(4)
    for mu in range(4):
        gm = -G[mu] if dagger else G[mu]
        fwd = np.einsum("tzyxab,stzyxb->stzyxa", M[mu], shift(psi, mu, +1, bc, 1))
        Ub = np.roll(M[mu], 1, axis=AX[mu])                       # U_mu(n-mu)
        bwd = np.einsum("tzyxba,stzyxb->stzyxa", np.conj(Ub), shift(psi, mu, -1, bc, 1))
        acc += kappa * np.einsum("sr,rtzyxa->stzyxa", r * one - gm, fwd)
        acc += kappa * np.einsum("sr,rtzyxa->stzyxa", r * one + gm, bwd)
    return psi - acc


def eta(dims, mu):
    NX, NY, NZ, NT = dims
    t, z, y, x = np.meshgrid(np.arange(NT), np.arange(NZ), np.arange(NY), np.arange(NX), indexing="ij")
    e = [np.zeros_like(x), x, x + y, x + y + z][mu]
    return (1 - 2 * (e % 2)).astype(float)


def staggered(U, chi, mass, bc=(1, 1, 1, -1), dagger=False):
    M = links_mat(U)
    dims = chi.shape[3], chi.shape[2], chi.shape[1], chi.shape[0]
    acc = np.zeros_like(chi)
    for mu in range(4):
        fwd = np.einsum("tzyxab,tzyxb->tzyxa", M[mu], shift(chi, mu, +1, bc))
        Ub = np.roll(M[mu], 1, axis=AX[mu])
        bwd = np.einsum("tzyxba,tzyxb->tzyxa", np.conj(Ub), shift(chi, mu, -1, bc))
        acc += 0.5 * eta(dims, mu)[..., None] * (fwd - bwd)
    return mass * chi + (-acc if dagger else acc)


def gauge_transform(U, g):
    """U_mu(n) -> g(n) U_mu(n) g(n+mu)^dag ;  g [t,z,y,x,a,b]"""
    M = links_mat(U)
    out = np.empty_like(M)
    for mu in range(4):
        gf = np.roll(g, -1, axis=AX[mu])
        out[mu] = g @ M[mu] @ np.conj(np.swapaxes(gf, -1, -2))
    return np.ascontiguousarray(np.swapaxes(out, -1, -2))


# ---- Wilson-clover (new capability; convention in oracle/lqcd_oracle.h orc_op.csw) -----------------------------
PLANES = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]


def _at(M, *steps):
    """field M[t,z,y,x,...] evaluated at n + sum of steps (mu, s)"""
    for mu, s in steps:
        M = np.roll(M, -s, axis=AX[mu])
    return M


def _dag(M):
    return np.conj(np.swapaxes(M, -1, -2))


def fmunu(U, traceless=True):
    """F^_mu_nu = (Q - Q^dag)/8 per plane, Q = four leaves; returns [plane, t,z,y,x, a, b]"""
    M = links_mat(U)
    out = []
    for mu, nu in PLANES:
        Um, Un = M[mu], M[nu]
        Q = Um @ _at(Un, (mu, 1)) @ _dag(_at(Um, (nu, 1))) @ _dag(Un)
        Q = Q + Un @ _dag(_at(Um, (mu, -1), (nu, 1))) @ _dag(_at(Un, (mu, -1))) @ _at(Um, (mu, -1))
        Q = Q + _dag(_at(Um, (mu, -1))) @ _dag(_at(Un, (mu, -1), (nu, -1))) @ _at(Um, (mu, -1), (nu, -1)) @ _at(Un, (nu, -1))
        Q = Q + _dag(_at(Un, (nu, -1))) @ _at(Um, (nu, -1)) @ _at(Un, (mu, 1), (nu, -1)) @ _dag(Um)
        F = (Q - _dag(Q)) / 8
        if traceless:
            F = F - np.trace(F, axis1=-2, axis2=-1)[..., None, None] * np.eye(3) / 3
        out.append(F)
    return np.array(out)


def sigma(mu, nu):
    return 0.5j * (G[mu] @ G[nu] - G[nu] @ G[mu])


def clover_term(U, psi, kappa, csw):
    """(A - 1) psi = kappa csw sum_{mu<nu} sigma_mu_nu (x) (i F^_mu_nu) psi"""
    F = fmunu(U)
    out = np.zeros_like(psi)
    for p, (mu, nu) in enumerate(PLANES):
        out += kappa * csw * np.einsum("sr,tzyxab,rtzyxb->stzyxa", sigma(mu, nu), 1j * F[p], psi)
    return out


def wilson_clover(U, psi, kappa, csw, r=1.0, bc=(1, 1, 1, -1), dagger=False):
    return wilson(U, psi, kappa, r, bc, dagger) + clover_term(U, psi, kappa, csw)

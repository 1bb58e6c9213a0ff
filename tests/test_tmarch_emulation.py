"""
CPU emulation of the index / shared-memory-window logic of the experimental t-marching Wilson kernel
(latticeqcd.jl_b200/csrc/wilson_dslash3.cu, LQCD_WILSON_KERNEL=3, off by default).

The kernel cannot be run in the build container (no GPU), so its CONTROL structure is mirrored here one-to-one in
numpy -- CTA -> (patch, chunk), warp -> block of the slice, lane -> site, the 3-slot spinor window with its load /
reuse schedule, the in-patch vs global classification of the spatial neighbours, t wrap-around and boundary phases --
and the result is compared with the independent numpy operator (tests/np_ref.py).  Every formula below has the same
name in the CUDA source.
"""
import numpy as np
import pytest

import np_ref
from oracle import oracle as orc

S = np_ref  # gamma tables


def to_aosoa_spinor(psi):
    """host [alpha,t,z,y,x,c] -> f[blk, k=3*alpha+c, lane]"""
    al, T, Z, Y, X, _ = psi.shape
    V = T * Z * Y * X
    flat = psi.reshape(4, V, 3)                          # [alpha, site, c]
    f = flat.transpose(1, 0, 2).reshape(V // 32, 32, 12)  # [blk, lane, k]
    return np.ascontiguousarray(f.transpose(0, 2, 1))    # [blk, k, lane]


def to_aosoa_links(U):
    """host [mu,t,z,y,x,b,a] -> g[blk, mu, e=3a+b, lane]"""
    V = int(np.prod(U.shape[1:5]))
    M = np.swapaxes(U, -1, -2).reshape(4, V, 9)          # [mu, site, 3a+b]
    g = M.transpose(1, 0, 2).reshape(V // 32, 32, 4, 9)
    return np.ascontiguousarray(g.transpose(0, 2, 3, 1))


def from_aosoa_spinor(f, dims):
    X, Y, Z, T = dims
    V = X * Y * Z * T
    flat = f.transpose(0, 2, 1).reshape(V, 4, 3).transpose(1, 0, 2)
    return np.ascontiguousarray(flat.reshape(4, T, Z, Y, X, 3))


def hop_math(mu, fwd, dag, psi12, U9, phase):
    """(1 -+ gamma_mu) U psi for one site, un-projected (emulation checks indexing, not the projector trick)."""
    sgn = (-1 if fwd else +1) * (-1 if dag else +1)       # D: forward (1-g), backward (1+g)
    P = np.eye(4) + sgn * S.G[mu]
    Um = U9.reshape(3, 3)
    if not fwd:
        Um = Um.conj().T
    sp = psi12.reshape(4, 3)
    return phase * (P @ (sp @ Um.T))                       # [alpha, colour]


def emulate(dims, kappa, bc, U, psi, Lc, dag=False, pc=(1, 2, 2)):
    X, Y, Z, T = dims
    # block shape (make_tiling): s[i] = min(d[i], rem)
    rem, s = 32, []
    for d in (X, Y, Z, T):
        si = min(d, rem); assert rem % si == 0 and d % si == 0
        s.append(si); rem //= si
    assert s[3] == 1, "kernel 3 needs blocks inside one t-slice"
    nb = [X // s[0], Y // s[1], Z // s[2]]
    c = list(pc)
    assert all(nb[i] % c[i] == 0 for i in range(3))
    nt = [nb[i] // c[i] for i in range(3)]
    W = c[0] * c[1] * c[2]
    nsb = nb[0] * nb[1] * nb[2]
    npatch = nt[0] * nt[1] * nt[2]
    assert T % Lc == 0
    nchunk = T // Lc
    f_in = to_aosoa_spinor(psi)
    g = to_aosoa_links(U)
    f_out = np.zeros_like(f_in)
    loads_issued = 0
    for bid in range(npatch * nchunk):
        patch, chunk = bid % npatch, bid // npatch
        p0, p1, p2 = patch % nt[0], (patch // nt[0]) % nt[1], patch // (nt[0] * nt[1])
        t0 = chunk * Lc
        win = np.zeros((3, W, 12, 32), dtype=complex)
        loaded_rel = [None, None, None]

        def bslice_of(w):
            w0, w1, w2 = w % c[0], (w // c[0]) % c[1], w // (c[0] * c[1])
            b0, b1, b2 = p0 * c[0] + w0, p1 * c[1] + w1, p2 * c[2] + w2
            return b0 + nb[0] * (b1 + nb[1] * b2)

        def issue_load(rel):
            nonlocal loads_issued
            t = (t0 - 1 + rel + T) % T
            for w in range(W):
                win[rel % 3, w] = f_in[bslice_of(w) + t * nsb]
                loads_issued += 1
            loaded_rel[rel % 3] = rel

        for rel in range(3):
            issue_load(rel)
        for r in range(1, Lc + 1):
            t = t0 + r - 1
            assert loaded_rel[(r - 1) % 3] == r - 1 and loaded_rel[r % 3] == r and loaded_rel[(r + 1) % 3] == r + 1
            for w in range(W):
                bsl = bslice_of(w)
                blk = bsl + t * nsb
                for lane in range(32):
                    ssl = bsl * 32 + lane
                    x, y, z = ssl % X, (ssl // X) % Y, ssl // (X * Y)
                    acc = np.zeros((4, 3), dtype=complex)
                    for mu, (coord, dim, stride) in enumerate(((x, X, 1), (y, Y, X), (z, Z, X * Y))):
                        for fwd in (1, 0):
                            wrapd = (coord == dim - 1) if fwd else (coord == 0)
                            nssl = (ssl - (dim - 1) * stride if wrapd else ssl + stride) if fwd else \
                                   (ssl + (dim - 1) * stride if wrapd else ssl - stride)
                            nbl, nl = nssl >> 5, nssl & 31
                            q0, q1, q2 = nbl % nb[0], (nbl // nb[0]) % nb[1], nbl // (nb[0] * nb[1])
                            inp = all(pp * cc <= qq < (pp + 1) * cc for pp, cc, qq in ((p0, c[0], q0), (p1, c[1], q1), (p2, c[2], q2)))
                            if inp:
                                nw = (q0 - p0 * c[0]) + c[0] * ((q1 - p1 * c[1]) + c[1] * (q2 - p2 * c[2]))
                                sp = win[r % 3, nw, :, nl]
                            else:
                                sp = f_in[nbl + t * nsb, :, nl]
                            link = g[blk, mu, :, lane] if fwd else g[nbl + t * nsb, mu, :, nl]
                            acc += hop_math(mu, fwd, dag, sp, link, bc[mu] if wrapd else 1.0)
                    # t direction: same block position and lane, neighbouring window slots
                    tm = (t - 1 + T) % T
                    acc += hop_math(3, 1, dag, win[(r + 1) % 3, w, :, lane], g[blk, 3, :, lane], bc[3] if t == T - 1 else 1.0)
                    acc += hop_math(3, 0, dag, win[(r - 1) % 3, w, :, lane], g[bsl + tm * nsb, 3, :, lane], bc[3] if t == 0 else 1.0)
                    own = win[r % 3, w, :, lane].reshape(4, 3)
                    f_out[blk, :, lane] = (own - kappa * acc).reshape(12)
            if r + 2 <= Lc + 1:
                issue_load(r + 2)                      # into slot (r+2)%3 == (r-1)%3, free after this step
    return from_aosoa_spinor(f_out, dims), loads_issued


@pytest.mark.parametrize("dims,Lc,pc", [((8, 8, 4, 4), 2, (1, 2, 2)), ((16, 4, 2, 6), 3, (1, 2, 1)), ((32, 2, 2, 4), 4, (1, 2, 2)), ((8, 4, 4, 4), 4, (1, 1, 2))])
@pytest.mark.parametrize("dag", [False, True])
def test_tmarch_window_logic(dims, Lc, pc, dag):
    U = orc.random_su3(dims, seed=3)
    psi = orc.gaussian_field(dims, orc.WILSON, seed=4)
    got, nloads = emulate(dims, 0.13, (1, 1, 1, -1), U, psi, Lc, dag=dag, pc=pc)
    want = np_ref.wilson(U, psi, 0.13, dagger=dag)
    assert np.abs(got - want).max() < 1e-13
    # every chunk loads Lc + 2 slices of its patch
    X, Y, Z, T = dims
    assert nloads == (X * Y * Z * T // 32) * (Lc + 2) // Lc

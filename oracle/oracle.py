"""
ctypes front-end of the CPU ORACLE (oracle/lqcd_oracle.c).  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

*** PARITY UNPINNED *** -- see oracle/lqcd_oracle.h.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py may import this module.

numpy host layouts (C-order; identical bytes to the Julia column-major arrays):
    links   U   : complex128[4, NT, NZ, NY, NX, 3(b), 3(a)]   U[mu,t,z,y,x,b,a] = U_mu(x,y,z,t)[a,b]
    Wilson  psi : complex128[4(alpha), NT, NZ, NY, NX, 3(c)]
    stagg.  chi : complex128[NT, NZ, NY, NX, 3(c)]
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

WILSON, STAGGERED, WILSON_EO = 0, 1, 2
D, DDAG, DDAGD = 0, 1, 2


class OrcOp(C.Structure):
    _fields_ = [
        ("dims", C.c_int * 4),
        ("bc", C.c_double * 4),
        ("kappa", C.c_double),
        ("r", C.c_double),
        ("rplusg", C.c_double * (4 * 4 * 4 * 2)),
        ("rminusg", C.c_double * (4 * 4 * 4 * 2)),
        ("mass", C.c_double),
        ("csw", C.c_double),
        ("clov", C.c_void_p),
    ]


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc, OpenMP)."""
    so = _HERE / "liblqcd_oracle.so"
    src = [_HERE / "lqcd_oracle.c", _HERE / "lqcd_oracle.h"]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in src):
        subprocess.check_call(["make", "-C", str(_HERE), "-s"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = _HERE / "liblqcd_oracle.so"
        if not so.exists():
            build()
        # libgomp's default spin-waiting burns the container's CPU quota between parallel regions and gets
        # the process throttled (observed: 60 ms instead of 1 ms per 8^4 application); sleep instead.
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        L = C.CDLL(str(so))
        vp, i64, dbl, ci = C.c_void_p, C.c_int64, C.c_double, C.c_int
        op = C.POINTER(OrcOp)
        pp = C.POINTER(C.c_void_p)
        L.orc_set_gamma.argtypes = [op, dbl]
        L.orc_set_threads.argtypes = [ci]; L.orc_set_threads.restype = ci
        L.orc_apply.argtypes = [op, ci, ci, vp, pp, vp, vp]
        for name in ("orc_cg", "orc_cgnr", "orc_bicgstab"):
            f = getattr(L, name)
            f.argtypes = [op, ci, vp, pp, vp, dbl, ci, C.POINTER(dbl), vp]
            f.restype = ci
        L.orc_mscg.argtypes = [op, ci, pp, pp, vp, C.POINTER(dbl), ci, dbl, ci, C.POINTER(dbl)]
        L.orc_mscg.restype = ci
        L.orc_plaquette.argtypes = [C.POINTER(ci), pp]; L.orc_plaquette.restype = dbl
        L.orc_wilson_force.argtypes = [op, pp, pp, vp, vp]
        L.orc_staggered_force.argtypes = [op, pp, pp, vp, vp]
        L.orc_clover_force.argtypes = [op, pp, pp, vp, vp]
        L.orc_clover_build.argtypes = [op, vp, vp, pp]
        L.orc_eo_solve.argtypes = [op, ci, ci, vp, pp, vp, dbl, ci, C.POINTER(dbl), vp]
        L.orc_eo_solve.restype = ci
        L.orc_wilson_hop_parity.argtypes = [op, ci, ci, vp, pp, vp]
        L.orc_md_update_u.argtypes = [C.POINTER(ci), pp, pp, dbl]
        L.orc_md_update_p_gauge.argtypes = [C.POINTER(ci), pp, pp, dbl, dbl]
        L.orc_md_update_p_force.argtypes = [C.POINTER(ci), pp, pp, dbl]
        L.orc_md_kinetic.argtypes = [C.POINTER(ci), pp]; L.orc_md_kinetic.restype = dbl
        L.orc_md_gauge_action.argtypes = [C.POINTER(ci), pp, dbl]; L.orc_md_gauge_action.restype = dbl
        _LIB = L
    return _LIB


def set_threads(n: int) -> int:
    return lib().orc_set_threads(int(n))


def make_op(dims, kappa=0.141139, r=1.0, mass=0.5, bc=(1, 1, 1, -1), csw=0.0) -> OrcOp:
    """Defaults are the reference's: hop/r/mass/BoundaryCondition, parameter_structs.jl:126-133."""
    op = OrcOp()
    for i in range(4):
        op.dims[i] = int(dims[i])
        op.bc[i] = float(bc[i])
    op.kappa, op.mass, op.csw = float(kappa), float(mass), float(csw)
    lib().orc_set_gamma(C.byref(op), float(r))
    return op


def gamma_tables(op: OrcOp):
    """(rplusg, rminusg) as complex arrays [mu, row, col]."""
    rp = np.frombuffer(op.rplusg, dtype=np.complex128).reshape(4, 4, 4).copy()
    rm = np.frombuffer(op.rminusg, dtype=np.complex128).reshape(4, 4, 4).copy()
    return rp, rm


def _uptrs(U: np.ndarray):
    assert U.dtype == np.complex128 and U.flags.c_contiguous and U.shape[0] == 4
    arr = (C.c_void_p * 4)(*[U[mu].ctypes.data for mu in range(4)])
    return arr


def _chk(a: np.ndarray):
    assert a.dtype == np.complex128 and a.flags.c_contiguous
    return a.ctypes.data


def field_shape(dims, kind):
    NX, NY, NZ, NT = dims
    return (4, NT, NZ, NY, NX, 3) if kind == WILSON else (NT, NZ, NY, NX, 3)


def apply(op: OrcOp, kind: int, mode: int, U: np.ndarray, x: np.ndarray) -> np.ndarray:
    y = np.empty_like(x)
    lib().orc_apply(C.byref(op), kind, mode, _chk(y), _uptrs(U), _chk(x), None)
    return y


def dot(a: np.ndarray, b: np.ndarray) -> complex:
    return complex(np.vdot(a, b))


def _solve(fn, op, kind, U, b, x0, eps, maxsteps, want_hist):
    x = np.zeros_like(b) if x0 is None else np.ascontiguousarray(x0).copy()
    rs = C.c_double(0.0)
    hist = np.full(maxsteps + 1, np.nan) if want_hist else None
    it = fn(C.byref(op), kind, _chk(x), _uptrs(U), _chk(b), float(eps), int(maxsteps), C.byref(rs),
            hist.ctypes.data if want_hist else None)
    out = {"x": x, "iters": it, "resid_sq": rs.value, "converged": it >= 0}
    if want_hist:
        out["hist"] = hist[: (it if it >= 0 else maxsteps) + 1]
    return out


def cg(op, kind, U, b, x0=None, eps=1e-19, maxsteps=3000, hist=False):
    """solve_DinvX!(x, DdagD, b): CG; eps on |r|^2 absolute (parameter_structs.jl:174)."""
    return _solve(lib().orc_cg, op, kind, U, b, x0, eps, maxsteps, hist)


def cgnr(op, kind, U, b, x0=None, eps=1e-19, maxsteps=3000, hist=False):
    """solve_DinvX!(x, D, b): upstream 'bicg' == CGNR."""
    return _solve(lib().orc_cgnr, op, kind, U, b, x0, eps, maxsteps, hist)


def bicgstab(op, kind, U, b, x0=None, eps=1e-19, maxsteps=3000, hist=False):
    return _solve(lib().orc_bicgstab, op, kind, U, b, x0, eps, maxsteps, hist)


def mscg(op, kind, U, b, shifts, eps=1e-19, maxsteps=3000):
    shifts = np.ascontiguousarray(shifts, dtype=np.float64)
    xs = [np.zeros_like(b) for _ in shifts]
    xp = (C.c_void_p * len(xs))(*[x.ctypes.data for x in xs])
    rs = C.c_double(0.0)
    it = lib().orc_mscg(C.byref(op), kind, xp, _uptrs(U), _chk(b),
                        shifts.ctypes.data_as(C.POINTER(C.c_double)), len(xs), float(eps), int(maxsteps),
                        C.byref(rs))
    return {"xs": xs, "iters": it, "resid_sq": rs.value, "converged": it >= 0}


def eo_solve(op, U, b, method="bicg", dagger=False, x0=None, eps=1e-19, maxsteps=3000, hist=False):
    """even-odd (Schur) preconditioned solve of M x = b / M^dag x = b; method "bicg" (= CGNR) or "bicgstab" on Mhat."""
    m = {"bicg": 0, "cgnr": 0, "bicgstab": 1}[method]
    fn = lambda opp, kind, x, u, bb, e, ms, rs, h: lib().orc_eo_solve(opp, m, int(bool(dagger)), x, u, bb, e, ms, rs, h)
    return _solve(fn, op, WILSON, U, b, x0, eps, maxsteps, hist)


def hop_parity(op, U, x, parity, dagger=False):
    y = np.empty_like(x)
    lib().orc_wilson_hop_parity(C.byref(op), int(bool(dagger)), int(parity), _chk(y), _uptrs(U), _chk(x))
    return y


def clover_build(op: OrcOp, U: np.ndarray, want_f=False):
    """Builds the clover term A(n) for op (kappa, csw, gamma tables) and attaches it to op (op.clov points into the
    returned array, which is also kept alive on the op object).  Returns clov [V, 2, 6(j), 6(i)] (and F^ [V,6,3(b),3(a)])."""
    V = int(np.prod([op.dims[i] for i in range(4)]))
    clov = np.zeros((V, 2, 6, 6), dtype=np.complex128)
    f = np.zeros((V, 6, 3, 3), dtype=np.complex128) if want_f else None
    lib().orc_clover_build(C.byref(op), clov.ctypes.data, f.ctypes.data if want_f else None, _uptrs(U))
    op._clov_keepalive = clov
    op.clov = clov.ctypes.data
    return (clov, f) if want_f else clov


def plaquette(dims, U) -> float:
    d = (C.c_int * 4)(*[int(v) for v in dims])
    return lib().orc_plaquette(d, _uptrs(U))


# ---- gauge-sector MD steps (src/md/AbstractMD.jl:78-135); U, P: complex128[4,NT,NZ,NY,NX,3,3] link layout, updated in place
def _dims(dims):
    return (C.c_int * 4)(*dims)


def md_update_u(dims, U, P, eps):
    lib().orc_md_update_u(_dims(dims), _uptrs(U), _uptrs(P), float(eps))


def md_update_p_gauge(dims, P, U, eps, beta):
    lib().orc_md_update_p_gauge(_dims(dims), _uptrs(P), _uptrs(U), float(eps), float(beta))


def md_update_p_force(dims, P, F, eps):
    lib().orc_md_update_p_force(_dims(dims), _uptrs(P), _uptrs(F), float(eps))


def md_kinetic(dims, P) -> float:
    return lib().orc_md_kinetic(_dims(dims), _uptrs(P))


def md_gauge_action(dims, U, beta) -> float:
    return lib().orc_md_gauge_action(_dims(dims), _uptrs(U), float(beta))


def md_momenta(dims, seed=1) -> np.ndarray:
    """p = sum_a a_a i lambda_a / 2 with a_a ~ N(0,1) (numpy RNG; the device has its own counter-based generator)"""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((4,) + tuple(dims[::-1]) + (8,))
    lam = np.zeros((8, 3, 3), dtype=complex)
    lam[0][0, 1] = lam[0][1, 0] = 1; lam[1][0, 1] = -1j; lam[1][1, 0] = 1j; lam[2][0, 0] = 1; lam[2][1, 1] = -1
    lam[3][0, 2] = lam[3][2, 0] = 1; lam[4][0, 2] = -1j; lam[4][2, 0] = 1j; lam[5][1, 2] = lam[5][2, 1] = 1
    lam[6][1, 2] = -1j; lam[6][2, 1] = 1j; lam[7] = np.diag([1, 1, -2]) / np.sqrt(3)
    P = 0.5j * np.einsum("...a,aij->...ij", a, lam)
    return np.ascontiguousarray(np.swapaxes(P, -1, -2))          # host layout [.., b, a]


def force(op, kind, U, X, Y) -> np.ndarray:
    """UdSfdU for X = (D^dag D)^-1 phi, Y = D X.  Wilson-clover (op.csw != 0): hopping part + clover-term part."""
    out = np.zeros_like(U)
    fn = lib().orc_wilson_force if kind == WILSON else lib().orc_staggered_force
    fn(C.byref(op), _uptrs(out), _uptrs(U), _chk(X), _chk(Y))
    if kind == WILSON and op.csw != 0.0:
        lib().orc_clover_force(C.byref(op), _uptrs(out), _uptrs(U), _chk(X), _chk(Y))
    return out


# ----------------------------------------------------------------------------------------------
# fixture / synthetic input helpers (formats: SURVEY.md section 4)
# ----------------------------------------------------------------------------------------------
def load_bridgetext(path, dims, nc=3) -> np.ndarray:
    """Bridge++ text (universe.jl:66-68 load_BridgeText!): one float per line; site-major x fastest;
    inside a site mu, row a, col b, (re, im).  Returns the numpy host link layout."""
    NX, NY, NZ, NT = dims
    raw = np.loadtxt(path, dtype=np.float64)
    assert raw.size == NX * NY * NZ * NT * 4 * nc * nc * 2, raw.size
    f = raw.reshape(NT, NZ, NY, NX, 4, nc, nc, 2)
    z = f[..., 0] + 1j * f[..., 1]                      # [t,z,y,x,mu,a,b]
    return np.ascontiguousarray(z.transpose(4, 0, 1, 2, 3, 6, 5))   # [mu,t,z,y,x,b,a]


def load_ildg(path, dims, nc=3) -> np.ndarray:
    """ILDG/LIME (universe.jl:62-65): big-endian float64 payload of record 'ildg-binary-data',
    same ordering as the Bridge text."""
    NX, NY, NZ, NT = dims
    n = NX * NY * NZ * NT * 4 * nc * nc * 2
    data = Path(path).read_bytes()
    pos = 0
    payload = None
    while pos + 144 <= len(data):
        magic = int.from_bytes(data[pos:pos + 4], "big")
        if magic != 0x456789AB:
            break
        length = int.from_bytes(data[pos + 8:pos + 16], "big")
        rtype = data[pos + 16:pos + 144].split(b"\0")[0].decode()
        pos += 144
        if rtype == "ildg-binary-data":
            payload = data[pos:pos + length]
        pos += (length + 7) // 8 * 8
    assert payload is not None and len(payload) == n * 8
    raw = np.frombuffer(payload, dtype=">f8").astype(np.float64)
    f = raw.reshape(NT, NZ, NY, NX, 4, nc, nc, 2)
    z = f[..., 0] + 1j * f[..., 1]
    return np.ascontiguousarray(z.transpose(4, 0, 1, 2, 3, 6, 5))


def random_su3(dims, seed=111, eps=None) -> np.ndarray:
    """Seeded synthetic SU(3) links.  eps=None: 'hot' (Haar via QR of a complex Gaussian);
    eps=float: 'warm' field exp(i eps H) with H a random traceless Hermitian matrix."""
    NX, NY, NZ, NT = dims
    rng = np.random.default_rng(seed)
    shape = (4, NT, NZ, NY, NX)
    a = rng.standard_normal(shape + (3, 3)) + 1j * rng.standard_normal(shape + (3, 3))
    if eps is None:
        q, r = np.linalg.qr(a)
        dg = np.diagonal(r, axis1=-2, axis2=-1)
        q = q * (dg / np.abs(dg))[..., None, :]
        det = np.linalg.det(q)
        q = q / (det ** (1.0 / 3.0))[..., None, None]
        m = q
    else:
        h = (a + np.conj(np.swapaxes(a, -1, -2))) / 2
        h = h - np.trace(h, axis1=-2, axis2=-1)[..., None, None] * np.eye(3) / 3
        w, v = np.linalg.eigh(h)
        m = (v * np.exp(1j * eps * w)[..., None, :]) @ np.conj(np.swapaxes(v, -1, -2))
    # m[..., a, b] -> host layout [..., b, a]
    return np.ascontiguousarray(np.swapaxes(m, -1, -2))


def gaussian_field(dims, kind, seed=112) -> np.ndarray:
    """complex Gaussian, sigma^2 = 1/2 per real component (SURVEY.md 8d / App. C.6 heatbath normalisation)."""
    rng = np.random.default_rng(seed)
    shp = field_shape(dims, kind)
    return np.ascontiguousarray((rng.standard_normal(shp) + 1j * rng.standard_normal(shp)) * np.sqrt(0.5))


def point_source(dims, kind, color=0, spin=0) -> np.ndarray:
    """setindex_global!(b, 1, ic, 1,1,1,1, is) -- measure_Pion_correlator.jl:374-376."""
    b = np.zeros(field_shape(dims, kind), dtype=np.complex128)
    if kind == WILSON:
        b[spin, 0, 0, 0, 0, color] = 1.0
    else:
        b[0, 0, 0, 0, color] = 1.0
    return b

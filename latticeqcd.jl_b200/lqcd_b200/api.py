"""
Host-side mirror of the reference's operator interface for the Dirac-solve path, above the C ABI.

The reference reaches this path through Julia generic functions defined by LatticeDiracOperators.jl /
Gaugefields.jl.  No Julia runtime exists in the build image, so this module is the Python
twin of latticeqcd.jl_b200/julia/LQCDB200.jl: same function names (Julia's ``f!`` is spelled ``f_``), same
argument order and meaning, same error behaviour, so the parity tests read like the reference's call sites:

    U  = Initialize_Gaugefields(3, 0, NX, NY, NZ, NT, condition="cold")          # universe.jl:41-49
    x  = Initialize_pseudofermion_fields(U[0], "Wilson", nowing=True)            # universe.jl:112
    D  = Dirac_operator(U, x, {"Dirac_operator": "Wilson", "κ": 0.141139, ...})  # universe.jl:113-137
    mul_(y, D, x); mul_(y, adjoint(D), x); mul_(y, DdagD(D), x)                  # measure_Pion_correlator.jl:379
    solve_DinvX_(y, D, b)                                                        # measure_Pion_correlator.jl:399
    fa = FermiAction(D, {"Nf": 2}); calc_UdSfdU_(UdSfdU, fa, U, eta)              # universe.jl:138, AbstractMD.jl:129
    U  = load_gaugefield("conf_00000100.ildg", (4, 4, 4, 4), "ILDG"); save_textdata(U, "out.txt")   # universe.jl:62-68, lqcd.jl:236-242
    props, infos = calc_quark_propagators_point_source(D)                        # measure_Pion_correlator.jl:333-409 (one batched solve)
    pbp, _, _ = measure_chiral_condensate(D, Nr=10)                              # measure_chiral_condensate.jl:164-204
    accepted, dH, info = hmc_update_(U, beta, dtau, MDsteps, fa=fa)               # standardHMC.jl:41-91, trajectory on the device

Link fields stay host numpy arrays in the Julia memory layout (the gauge sector keeps using them on the
CPU, SURVEY.md 8b "Selection" option 1); they are mirrored to the device when an operator is built or
re-bound with ``D(U)``.  Pseudofermion fields are device-resident handles with explicit
``to_host()`` / ``from_host()``.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import weakref

import numpy as np

from . import _lib as L

_CTX = {}     # (dims, procgrid, rank, device) -> Context


class Context:
    """One per process and GPU (lqcd_ctx_create)."""

    def __init__(self, dims, procgrid=(1, 1, 1, 1), rank=0, device=0):
        self.lib = L.load()
        self.dims = tuple(int(d) for d in dims)
        self.procgrid = tuple(int(p) for p in procgrid)
        self.rank, self.device = int(rank), int(device)
        h = C.c_void_p()
        d = (C.c_int * 4)(*self.dims)
        p = (C.c_int * 4)(*self.procgrid)
        L.check(None, self.lib.lqcd_ctx_create(d, p, self.rank, self.device, C.byref(h)))
        self.h = h
        ld, og = (C.c_int * 4)(), (C.c_int * 4)()
        self.call("lqcd_local_dims", ld, og)
        self.local_dims, self.origin = tuple(ld), tuple(og)
        self.gauge_epoch = None
        self.dist = None              # torch.distributed module once connect_ranks ran (multi-rank host barriers)
        self._fin = weakref.finalize(self, self.lib.lqcd_ctx_destroy, h)

    def call(self, name, *args):
        L.check(self.h, getattr(self.lib, name)(self.h, *args))

    def launch_count(self) -> int:
        n = C.c_uint64()
        self.call("lqcd_launch_count", C.byref(n))
        return n.value

    def synchronize(self):
        self.call("lqcd_synchronize")

    def barrier(self):
        """host barrier over the ranks of the job (no-op on a single rank)"""
        if self.dist is not None:
            self.synchronize()
            self.dist.barrier()

    def stream(self) -> int:
        s = C.c_void_p()
        self.call("lqcd_stream", C.byref(s))
        return s.value


def get_context(dims, procgrid=(1, 1, 1, 1), rank=0, device=0) -> Context:
    key = (tuple(dims), tuple(procgrid), rank, device)
    if key not in _CTX:
        _CTX[key] = Context(dims, procgrid, rank, device)
    return _CTX[key]


def decompose(dims, procgrid, rank):
    """lqcd_decompose: (local_dims, origin, nbr_lo, nbr_hi) -- runs without a GPU."""
    lib = L.load()
    a = [(C.c_int * 4)() for _ in range(4)]
    st = lib.lqcd_decompose((C.c_int * 4)(*dims), (C.c_int * 4)(*procgrid), int(rank), *a)
    if st != 0:
        raise ValueError(f"bad decomposition dims={dims} procgrid={procgrid} rank={rank}")
    return tuple(tuple(v) for v in a)


def exchange_handles(blob: bytes, dist) -> bytes:
    """all-gather one fixed-size byte blob per rank, in rank order (torch.distributed: gloo on CPU tensors,
    nccl on CUDA tensors).  The Julia shim does the same with MPI.Allgather."""
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)


def connect_ranks(ctx: Context, dist):
    """Map every rank's comm buffer into every other rank (CUDA IPC over NVLink): lqcd_comm_export ->
    all-gather -> lqcd_comm_connect -> barrier."""
    buf = C.create_string_buffer(L.IPC_HANDLE_BYTES)
    ctx.call("lqcd_comm_export", buf)
    allh = exchange_handles(buf.raw, dist)
    ctx.call("lqcd_comm_connect", C.c_char_p(allh))
    ctx.dist = dist
    dist.barrier()


# ---------------------------------------------------------------------------------------------------
# link fields (Gaugefields.jl container; host-resident, Julia layout)
# ---------------------------------------------------------------------------------------------------
class GaugeField:
    """U[mu]: AbstractGaugefields{NC,4}.  ``.U`` is ComplexF64[NC,NC,NX,NY,NZ,NT] in Julia order, i.e. the
    C-order numpy array [NT,NZ,NY,NX,b,a].  Fields used by the reference: .NC .NV .NX .NY .NZ .NT, size()."""

    def __init__(self, arr, NC, dims, parent):
        self.U, self.NC = arr, NC
        self.NX, self.NY, self.NZ, self.NT = dims
        self.NV = int(np.prod(dims))
        self.parent = parent

    def size(self):
        return (self.NC, self.NC, self.NX, self.NY, self.NZ, self.NT)


class Gaugefields(list):
    """Vector{<:AbstractGaugefields} of length 4 sharing one [4,NT,NZ,NY,NX,3,3] array."""

    def __init__(self, data, dims, ctx_args=None):
        self.data = np.ascontiguousarray(data, dtype=np.complex128)
        self.dims = tuple(dims)
        self.ctx_args = ctx_args or {}
        super().__init__(GaugeField(self.data[mu], 3, self.dims, self) for mu in range(4))

    def context(self) -> Context:
        return get_context(self.dims, **self.ctx_args)


def Initialize_Gaugefields(NC, Nwing, NX, NY, NZ, NT, condition="cold", seed=111, **ctx_args) -> Gaugefields:
    """src/system/universe.jl:41-49.  cold = identity, hot = random SU(3) (numpy Haar, seeded)."""
    if NC != 3:
        raise ValueError("the B200 path implements NC = 3")
    dims = (NX, NY, NZ, NT)
    data = np.zeros((4, NT, NZ, NY, NX, 3, 3), dtype=np.complex128)
    if condition == "cold":
        for a in range(3):
            data[..., a, a] = 1.0
    elif condition == "hot":
        rng = np.random.default_rng(seed)
        a = rng.standard_normal(data.shape) + 1j * rng.standard_normal(data.shape)
        q, r = np.linalg.qr(a)
        dg = np.diagonal(r, axis1=-2, axis2=-1)
        q = q * (dg / np.abs(dg))[..., None, :]
        q = q / (np.linalg.det(q) ** (1.0 / 3.0))[..., None, None]
        data[:] = np.swapaxes(q, -1, -2)
    else:
        raise ValueError(f"condition {condition!r} not supported")
    return Gaugefields(data, dims, ctx_args)


def gaugefields_from_array(arr, global_dims=None, **ctx_args) -> Gaugefields:
    """Wrap links given as [4,NT,NZ,NY,NX,3(b),3(a)] (what load_BridgeText!/ILDG produce, universe.jl:62-68).
    Multi-rank: `arr` is this rank's LOCAL block, `global_dims` the full lattice and ctx_args carries
    procgrid / rank / device (upstream's *_mpi field types hold local arrays the same way)."""
    _, NT, NZ, NY, NX, _, _ = arr.shape
    return Gaugefields(arr, tuple(global_dims) if global_dims else (NX, NY, NZ, NT), ctx_args)


# ---------------------------------------------------------------------------------------------------
# gauge configurations in the reference's file formats (universe.jl:58-77, lqcd.jl:226-247; csrc/gauge_io.cu)
# ---------------------------------------------------------------------------------------------------
_FORMATS = {"ILDG": L.IO_ILDG, "BridgeText": L.IO_BRIDGETEXT}


def _io_call(name, path, fmt, dims, nc, arr):
    lib = L.load()
    ptrs = (C.c_void_p * 4)(*[arr[mu].ctypes.data for mu in range(4)])
    st = getattr(lib, name)(str(path).encode(), _FORMATS[fmt], (C.c_int * 4)(*dims), nc, ptrs, 0)
    L.check(None, st)


def load_gaugefield(path, dims, loadU_format="ILDG", NC=3, **ctx_args) -> Gaugefields:
    """`initial = path`, loadU_format in {"ILDG", "BridgeText"} (universe.jl:62-68: ILDG(filename) + load_gaugefield!, load_BridgeText!)
    into host link arrays (no GPU needed).  dims = (NX, NY, NZ, NT)."""
    NX, NY, NZ, NT = dims
    data = np.zeros((4, NT, NZ, NY, NX, NC, NC), dtype=np.complex128)
    _io_call("lqcd_io_read_gauge", path, loadU_format, dims, NC, data)
    if NC != 3:
        return data
    return Gaugefields(data, dims, ctx_args)


def load_BridgeText_(path, U: Gaugefields):
    """load_BridgeText!(filename, U, L, NC) (universe.jl:66-68): overwrite U with the configuration in the file"""
    _io_call("lqcd_io_read_gauge", path, "BridgeText", U.dims, 3, U.data)
    return U


def save_binarydata(U, path):
    """save_binarydata(U, filename) (lqcd.jl:239): ILDG.  U: Gaugefields or a raw [4,NT,NZ,NY,NX,NC,NC] array"""
    data = U.data if isinstance(U, Gaugefields) else np.ascontiguousarray(U, dtype=np.complex128)
    _, NT, NZ, NY, NX, NC, _ = data.shape
    _io_call("lqcd_io_write_gauge", path, "ILDG", (NX, NY, NZ, NT), NC, data)


def save_textdata(U, path):
    """save_textdata(U, filename) (lqcd.jl:242): Bridge++ text"""
    data = U.data if isinstance(U, Gaugefields) else np.ascontiguousarray(U, dtype=np.complex128)
    _, NT, NZ, NY, NX, NC, _ = data.shape
    _io_call("lqcd_io_write_gauge", path, "BridgeText", (NX, NY, NZ, NT), NC, data)


def plaquette(ctx: Context) -> float:
    """average plaquette Re tr U_p / NC of the device links (what the reference prints every trajectory, lqcd.jl:186-192)"""
    out = C.c_double()
    ctx.call("lqcd_gauge_plaquette", C.byref(out))
    return out.value


def load_gaugefield_device_(ctx: Context, path, loadU_format="ILDG"):
    """file -> device links of this rank's block (lqcd_gauge_load): nothing passes through host link arrays"""
    ctx.call("lqcd_gauge_load", str(path).encode(), _FORMATS[loadU_format])


def save_gaugefield_device(ctx: Context, path, saveU_format="ILDG"):
    """device links -> file (lqcd_gauge_save).  Across ranks (ILDG): the rank owning the lattice origin creates the file first"""
    fmt = _FORMATS[saveU_format]
    if ctx.dist is None:
        ctx.call("lqcd_gauge_save", str(path).encode(), fmt)
        return
    first = all(o == 0 for o in ctx.origin)
    if first:
        ctx.call("lqcd_gauge_save", str(path).encode(), fmt)
    ctx.barrier()
    if not first:
        ctx.call("lqcd_gauge_save", str(path).encode(), fmt)
    ctx.barrier()


# ---------------------------------------------------------------------------------------------------
# pseudofermion fields (device resident)
# ---------------------------------------------------------------------------------------------------
class FermionField:
    def __init__(self, ctx: Context, kind: int):
        self.ctx, self.kind = ctx, kind
        h = C.c_void_p()
        ctx.call("lqcd_fermion_alloc", kind, C.byref(h))
        self.h = h
        self._fin = weakref.finalize(self, ctx.lib.lqcd_fermion_free, ctx.h, h)

    @property
    def host_shape(self):
        NX, NY, NZ, NT = self.ctx.local_dims
        return (4, NT, NZ, NY, NX, 3) if self.kind == L.WILSON else (NT, NZ, NY, NX, 3)

    def from_host(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.complex128)
        assert a.shape == self.host_shape, (a.shape, self.host_shape)
        self.ctx.call("lqcd_fermion_upload", self.h, a.ctypes.data, 0)
        return self

    def to_host(self):
        out = np.empty(self.host_shape, dtype=np.complex128)
        self.ctx.call("lqcd_fermion_download", self.h, out.ctypes.data, 0)
        return out

    # Julia-style indexing psi[ic, ix, iy, iz, it, ialpha] (1-based), measure_Pion_correlator.jl:244
    def __getitem__(self, idx):
        ic, ix, iy, iz, it, ia = idx
        h = self.to_host()
        return h[ia - 1, it - 1, iz - 1, iy - 1, ix - 1, ic - 1] if self.kind == L.WILSON else h[it - 1, iz - 1, iy - 1, ix - 1, ic - 1]


def Initialize_pseudofermion_fields(U1: GaugeField, kind: str, nowing=True, L5=None) -> FermionField:
    """src/system/universe.jl:107,112."""
    k = {"Wilson": L.WILSON, "staggered": L.STAGGERED, "Staggered": L.STAGGERED}.get(kind)
    if k is None:
        raise ValueError(f"fermion kind {kind!r} is not on the B200 path (Wilson, staggered)")
    return FermionField(U1.parent.context(), k)


def similar(x: FermionField) -> FermionField:                       # standardMD.jl:50-51
    return FermionField(x.ctx, x.kind)


def clear_fermion_(x: FermionField):                                # measure_Pion_correlator.jl:370
    x.ctx.call("lqcd_fermion_zero", x.h)


def substitute_fermion_(dst: FermionField, src: FermionField):
    dst.ctx.call("lqcd_fermion_copy", dst.h, src.h)


def setindex_global_(b: FermionField, v, ic, ix, iy, iz, it, ialpha):   # measure_Pion_correlator.jl:376
    if v != 1:
        raise ValueError("only unit point sources are supported")
    site = (C.c_int * 4)(ix - 1, iy - 1, iz - 1, it - 1)
    b.ctx.call("lqcd_fermion_point_source", b.h, site, ic - 1, ialpha - 1)


def mask_parity_(x: FermionField, parity=0):
    """keep the sites with (x+y+z+t) % 2 == parity, zero the others (staggered Nf = 4 even-site pseudofermions)."""
    x.ctx.call("lqcd_fermion_mask_parity", x.h, int(parity))


def gauss_distribution_fermion_(x: FermionField, seed=112):
    x.ctx.call("lqcd_fermion_gaussian", x.h, int(seed))


def Z4_distribution_fermi_(x: FermionField, seed=114):              # measure_chiral_condensate.jl:181
    x.ctx.call("lqcd_fermion_z4", x.h, int(seed))


def dot(a: FermionField, b: FermionField) -> complex:               # standardHMC.jl:54
    out = (C.c_double * 2)()
    a.ctx.call("lqcd_blas_dot", a.h, b.h, out)
    return complex(out[0], out[1])


def add_(y: FermionField, a, x: FermionField):                      # add!(y, a, x): y += a x
    a = complex(a)
    y.ctx.call("lqcd_blas_axpy", a.real, a.imag, x.h, y.h)


def add_xpby_(b, y: FermionField, x: FermionField):                 # add!(b, y, 1, x): y = b y + x
    b = complex(b)
    y.ctx.call("lqcd_blas_xpby", x.h, b.real, b.imag, y.h)


# ---------------------------------------------------------------------------------------------------
# operators
# ---------------------------------------------------------------------------------------------------
class DiracOperator:
    """Dirac_operator(U, x, params) (universe.jl:137).  Callable: D(U) re-binds (and re-uploads) the links
    (measure_Pion_correlator.jl:338)."""

    mode = L.OP_D

    def __init__(self, U: Gaugefields, x: FermionField, params: dict):
        self.params = dict(params)
        name = params["Dirac_operator"]
        self.op = L.LqcdOp()
        if name in ("Wilson", "WilsonClover"):
            self.op.kind = L.WILSON
            self.op.kappa = float(params["κ"])
            self.op.r = float(params.get("r", 1.0))
            if name == "WilsonClover":      # new capability (the surveyed wrapper stops at parameter_structs.jl:125)
                self.op.csw = float(params.get("Clover_coefficient", 1.5612))
        elif name in ("staggered", "Staggered"):
            self.op.kind = L.STAGGERED
            self.op.mass = float(params["mass"])
        else:
            raise ValueError(f"Dirac_operator {name!r} not supported")      # universe.jl:130 error("not supported")
        bc = params.get("boundarycondition", [1, 1, 1, -1])                  # parameter_structs.jl:133
        for i in range(4):
            self.op.bc[i] = float(bc[i])
        self.eps = float(params.get("eps_CG", 1e-19))                        # parameter_structs.jl:174
        self.maxsteps = int(params.get("MaxCGstep", 3000))                   # parameter_structs.jl:175
        self.verbose = int(params.get("verbose_level", 1))
        self.method = params.get("method_CG", "bicg")
        self.ctx = x.ctx
        self.kind = self.op.kind
        self.last = {}
        self._bind(U)

    def _bind(self, U: Gaugefields):
        """D(U): upload the links.  All operators of a context share ONE device link buffer, so the context counts bindings
        (`binding_epoch`) and every use of an operator (`_ensure_bound`) re-uploads its own U when somebody else bound the
        buffer in between (a second operator with another U, a rejected HMC trajectory).  Editing U.data on the host is not
        noticed: call D(U) again, as the reference does before every measurement (measure_Pion_correlator.jl:338)."""
        self.U = U
        ptrs = (C.c_void_p * 4)(*[U.data[mu].ctypes.data for mu in range(4)])
        self.ctx.call("lqcd_gauge_upload", ptrs, 3, 0)
        self.ctx.binding_epoch = getattr(self.ctx, "binding_epoch", 0) + 1
        self._bound_epoch = self.ctx.binding_epoch
        if self.op.kind == L.WILSON and self.op.csw != 0.0:
            # the clover term depends on the links only: build it at the D(U) rebinding.  Its leaves reach one site into
            # the neighbouring ranks' links, so all ranks must have uploaded first and nobody may re-upload while a peer reads.
            self.ctx.barrier()
            self.ctx.call("lqcd_clover_term", C.byref(self.op), None)
            self.ctx.barrier()

    def _ensure_bound(self):
        if getattr(self, "_bound_epoch", None) != getattr(self.ctx, "binding_epoch", None):
            self._bind(self.U)

    def clover_term(self) -> np.ndarray:
        """dense clover blocks [V_local, 2, 6(j), 6(i)] (oracle layout), for inspection / tests"""
        V = int(np.prod(self.ctx.local_dims))
        out = np.zeros((V, 2, 6, 6), dtype=np.complex128)
        self.ctx.call("lqcd_clover_term", C.byref(self.op), out.ctypes.data_as(C.c_void_p))
        return out

    def __call__(self, U: Gaugefields):
        self._bind(U)
        return self


class _Derived:
    def __init__(self, D: DiracOperator, mode):
        self.D, self.mode = D, mode


def adjoint(D: DiracOperator):            # D'
    return _Derived(D, L.OP_DDAG)


def DdagD(D: DiracOperator):              # DdagD_operator
    return _Derived(D, L.OP_DDAGD)


def Dirac_operator(U, x, params) -> DiracOperator:
    return DiracOperator(U, x, params)


def _base(A):
    return (A, A.mode) if isinstance(A, DiracOperator) else (A.D, A.mode)


def mul_(y: FermionField, A, x: FermionField):
    """LinearAlgebra.mul!(y, A, x)."""
    D, mode = _base(A)
    D._ensure_bound()
    D.ctx.call("lqcd_dslash", C.byref(D.op), y.h, x.h, mode)


def mul_host_(y_host: np.ndarray, A, x_host: np.ndarray, y: FermionField = None, x: FermionField = None):
    """mul!(y, A, x) for HOST arrays in the Julia layout (what the Julia shim does for the reference's CPU pseudofermion
    types): one pipelined upload + Dslash + download (lqcd_dslash_host).  y / x: optional device scratch fields."""
    D, mode = _base(A)
    D._ensure_bound()
    x = x or FermionField(D.ctx, D.kind)
    y = y or FermionField(D.ctx, D.kind)
    assert x_host.dtype == np.complex128 and y_host.dtype == np.complex128 and x_host.flags.c_contiguous and y_host.flags.c_contiguous
    assert x_host.shape == x.host_shape and y_host.shape == y.host_shape, (x_host.shape, x.host_shape)
    D.ctx.call("lqcd_dslash_host", C.byref(D.op), y.h, x.h, y_host.ctypes.data, x_host.ctypes.data, mode, 0)
    return y_host


_METHODS = {"bicg": L.SOLVER_CGNR, "bicgstab": L.SOLVER_BICGSTAB, "preconditiond_bicgstab": L.SOLVER_BICGSTAB}


def solve_DinvX_(y: FermionField, A, x: FermionField, history=False):
    """solve_DinvX!(y, A, x): A y = x, y is the initial guess.  A::Dirac_operator (or its adjoint) -> the
    routine upstream calls 'bicg' (CGNR) unless params["method_CG"] says otherwise; A::DdagD -> CG."""
    D, mode = _base(A)
    D._ensure_bound()
    method = L.SOLVER_CG if mode == L.OP_DDAGD else _METHODS[D.method]
    it, rs = C.c_int(0), C.c_double(0.0)
    hist = np.full(D.maxsteps + 1, np.nan) if (history or D.verbose >= 3) else None
    hp = hist.ctypes.data_as(L.pdbl) if hist is not None else None
    try:
        # params["evenodd"] = True (new key): Schur-preconditioned solve for D / D' systems of the plain Wilson operator
        fn = "lqcd_solve_eo" if (D.params.get("evenodd") and mode != L.OP_DDAGD) else "lqcd_solve"
        D.ctx.call(fn, C.byref(D.op), y.h, x.h, method, mode, D.eps, D.maxsteps, C.byref(it), C.byref(rs), hp)
    finally:
        D.last = {"iters": it.value, "resid_sq": rs.value,
                  "hist": None if hist is None else hist[: it.value + 1]}
    if D.verbose >= 3 and hist is not None:
        for i, v in enumerate(D.last["hist"]):
            print(f"{i}-th eps: {v}")                                         # upstream println_verbose_level3
    return D.last


def _handles(fields):
    return (C.c_void_p * len(fields))(*[f.h.value for f in fields])


def mul_multi_(ys, A, xs):
    """mul!(ys[j], A, xs[j]) for all j in ONE pass over the links (lqcd_dslash_multi); bit-identical to the loop of mul_"""
    D, mode = _base(A)
    D._ensure_bound()
    assert len(ys) == len(xs)
    D.ctx.call("lqcd_dslash_multi", C.byref(D.op), _handles(ys), _handles(xs), len(ys), mode)
    return ys


def solve_DinvX_multi_(ys, A, xs):
    """solve_DinvX!(ys[j], A, xs[j]) for all j in lock step (lqcd_solve_multi): the loops over sources of
    calc_quark_propagators_point_source (measure_Pion_correlator.jl:333-349) and of the chiral condensate's noise vectors
    (measure_chiral_condensate.jl:176-182) as one batched solve.  ys[j] is the initial guess.  Returns a list of per-source infos;
    raises NotConverged like solve_DinvX_ if any source did not converge."""
    D, mode = _base(A)
    D._ensure_bound()
    assert len(ys) == len(xs)
    method = L.SOLVER_CG if mode == L.OP_DDAGD else _METHODS[D.method]
    n = len(ys)
    its, rs = (C.c_int * n)(), (C.c_double * n)()
    try:
        D.ctx.call("lqcd_solve_multi", C.byref(D.op), _handles(ys), _handles(xs), n, method, mode, D.eps, D.maxsteps, its, rs)
    finally:
        D.last = {"iters": list(its), "resid_sq": list(rs)}
    return [{"iters": its[j], "resid_sq": rs[j]} for j in range(n)]


def calc_quark_propagators_point_source(D: DiracOperator, origin=(0, 0, 0, 0), U: Gaugefields = None):
    """D^-1 on the NC*Nspinor point sources at `origin` (measure_Pion_correlator.jl:333-409: source i has spin is = (i-1) % Nspinor,
    colour ic = (i-is) / Nspinor, value 1 at the origin; clear_fermion!(p); solve_DinvX!(p, D, b)), all sources in one batched
    solve.  Returns (propagators, infos), propagators[i] a device field like the reference's deepcopy(p).  origin = (x, y, z, t).
    U: the configuration to measure on -- re-bound first, D = m.D(U) as measure_Pion_correlator.jl:338 does; None = D's own U."""
    if U is not None:
        D(U)
    nspin = 4 if D.kind == L.WILSON else 1
    n = 3 * nspin
    bs = [FermionField(D.ctx, D.kind) for _ in range(n)]
    ps = [FermionField(D.ctx, D.kind) for _ in range(n)]
    x0, y0, z0, t0 = origin
    for i in range(n):
        is_, ic = i % nspin, i // nspin
        h = np.zeros(bs[i].host_shape, dtype=np.complex128)
        if D.kind == L.WILSON:
            h[is_, t0, z0, y0, x0, ic] = 1.0
        else:
            h[t0, z0, y0, x0, ic] = 1.0
        bs[i].from_host(h)
        clear_fermion_(ps[i])
    infos = solve_DinvX_multi_(ps, D, bs)
    return ps, infos


def measure_chiral_condensate(D: DiracOperator, Nr=10, factor=1.0, seed=114, batched=True, U: Gaugefields = None):
    """measure(m::Chiral_condensate_measurement, itrj, U) (measure_chiral_condensate.jl:164-204): for ir = 1:Nr { clear_fermion!(p);
    Z4_distribution_fermi!(r); solve_DinvX!(p, D, r); pbp += dot(r, p) }, pbp_value = real(pbp / Nr) / NV * factor.  batched: the Nr
    solves advance in lock step (lqcd_solve_multi); False: one solve_DinvX_ after the other like the reference.  Returns
    (pbp_value, per-source values, noise fields).  U: configuration to measure on (D = m.D(U), measure_chiral_condensate.jl:170)."""
    if U is not None:
        D(U)
    rs = [FermionField(D.ctx, D.kind) for _ in range(Nr)]
    ps = [FermionField(D.ctx, D.kind) for _ in range(Nr)]
    for ir, (r, p) in enumerate(zip(rs, ps)):
        Z4_distribution_fermi_(r, seed + ir)
        clear_fermion_(p)
    if batched:
        for j in range(0, Nr, 16):
            solve_DinvX_multi_(ps[j:j + 16], D, rs[j:j + 16])
    else:
        for r, p in zip(rs, ps):
            solve_DinvX_(p, D, r)
    NV = int(np.prod(D.ctx.dims))                                   # U[1].NV: the global volume
    vals = [dot(r, p) for r, p in zip(rs, ps)]
    return float(np.real(sum(vals) / Nr) / NV * factor), vals, rs


def shiftedcg_(ys, D: DiracOperator, x: FermionField, shifts, eps=None, maxsteps=None):
    """upstream shiftedcg: (DdagD + shifts[j]) ys[j] = x."""
    D._ensure_bound()
    sh = np.ascontiguousarray(shifts, dtype=np.float64)
    hs = (C.c_void_p * len(ys))(*[y.h.value for y in ys])
    it, rs = C.c_int(0), C.c_double(0.0)
    D.ctx.call("lqcd_multishift_cg", C.byref(D.op), hs, x.h, sh.ctypes.data_as(L.pdbl), len(ys),
               D.eps if eps is None else eps, D.maxsteps if maxsteps is None else maxsteps, C.byref(it), C.byref(rs))
    return {"iters": it.value, "resid_sq": rs.value}


# ---------------------------------------------------------------------------------------------------
# fermion actions.  FermiAction(D, parameters_action) (universe.jl:138) picks, like upstream (README.md:132):
#   Wilson, staggered Nf = 8 : S_f = eta^dag (D^dag D)^-1 eta                                     -> FermiActionB200
#   staggered Nf = 4         : the same on even-site pseudofermions (D^dag D does not couple parities) -> FermiActionB200(even)
#   staggered other Nf       : rational HMC, det(D^dag D)^{Nf/8} via multi-shift CG                  -> RHMCFermiAction
# ---------------------------------------------------------------------------------------------------
class FermiActionB200:
    """HMC pseudofermion action S_f = eta^dag (D^dag D)^-1 eta."""

    def __init__(self, D: DiracOperator, parameters_action: dict):
        self.D = D
        self.Nf = parameters_action.get("Nf", None)
        # staggered Nf = 4 (UPSTREAM-RECALL, unverified): the pseudofermion eta lives on even sites only.
        # D^dag D = m^2 - Dh^2 does not couple the parities, so eta_e^dag (D^dag D)_ee^-1 eta_e describes half the tastes.
        # Heat bath: xi0 Gaussian on ALL sites, eta = P_even D^dag xi0  =>  <eta eta^dag> = m^2 + D_eo D_eo^dag = (D^dag D)_ee, the
        # distribution the action needs (zeroing the odd sites of xi0 as well would give <eta eta^dag> = m^2, a wrong heat bath).
        # But eta <- xi0 is a projection, so xi0^dag xi0 is NOT the initial action that update! takes as Sfold = dot(xi, xi)
        # (standardHMC.jl:54).  gauss_sampling_in_action_ therefore hands back xi = D X with X = (D^dag D)^-1 eta (one CG): then
        # P_even D^dag xi = eta and xi^dag xi = eta^dag (D^dag D)^-1 eta exactly, and the reference's update! sequence stays exact
        # (tests/test_reference_regressions.py: dH = O(dtau^2), not the -O(V) offset of the naive shortcut).
        # Kept behind this flag; with it off the action is the 8-taste one.
        self.even_only = bool(D.kind == L.STAGGERED and self.Nf == 4 and parameters_action.get("even_site_pseudofermions", True))
        self._temporary_fermionfields = [FermionField(D.ctx, D.kind) for _ in range(4)]   # standardMD.jl:50
        # even-site action: its CG solves run on half fields unless switched off or the lattice does not qualify
        ld = D.ctx.local_dims
        self.half_field_solver = bool(self.even_only and parameters_action.get("half_field_solver", True)
                                      and all(d % 2 == 0 for d in ld) and (int(np.prod(ld)) // 2) % 32 == 0)
        self.last = {}


class RHMCFermiAction:
    """Rational HMC for staggered Nf not in {4, 8} (test/test_Nf2.toml, test_Nf3.toml): rational approximations of
    (D^dag D)^{+Nf/16} (heat bath) and (D^dag D)^{-Nf/8} (action, force) evaluated with lqcd_multishift_cg; the force
    sum_j alpha_j force(X_j, Y_j) is accumulated on the device (lqcd_fermion_force_xy).  Spectral range: the staggered
    hop is anti-Hermitian with norm <= 4, so spec(D^dag D) lies in [m^2, m^2 + 16]."""

    def __init__(self, D: DiracOperator, parameters_action: dict):
        from . import rhmc
        if D.kind != L.STAGGERED:
            raise ValueError("RHMC is wired for the staggered operator")
        self.D, self.Nf = D, int(parameters_action["Nf"])
        m2 = float(D.op.mass) ** 2
        lo = parameters_action.get("rational_lambda_min", 0.9 * m2)
        hi = parameters_action.get("rational_lambda_max", 1.05 * (m2 + 16.0))
        self.rhmc = rhmc.RHMCAction(rhmc.B200Backend(D), self.Nf, lo, hi, order=int(parameters_action.get("rational_order", 12)),
                                    tolerance=float(parameters_action.get("rational_tolerance", 1e-6)))
        self.even_only = False
        self._temporary_fermionfields = [FermionField(D.ctx, D.kind) for _ in range(4)]
        self.last = {}

    # single-call device paths (lqcd_rational_apply / lqcd_fermion_force_rational / lqcd_md_trajectory_rational): the poles
    # never come back to the host; fa.rhmc (backend-generic, shared with the CPU oracle in tests) stays as the cross-check
    @staticmethod
    def _coeffs(ra):
        a = np.ascontiguousarray(ra.alpha, dtype=np.float64)
        b = np.ascontiguousarray(ra.beta, dtype=np.float64)
        return float(ra.alpha0), a, b

    def rational_apply_(self, y: FermionField, ra, x: FermionField, want_dot=False):
        """y = ra(D^dag D) x; returns (CG iterations, Re<x, y> or None)"""
        D = self.D
        a0, a, b = self._coeffs(ra)
        it, d = C.c_int(0), C.c_double(0.0)
        D.ctx.call("lqcd_rational_apply", C.byref(D.op), y.h, x.h, a0, a.ctypes.data_as(L.pdbl), b.ctypes.data_as(L.pdbl), len(a),
                   D.eps, D.maxsteps, C.byref(it), C.byref(d) if want_dot else None)
        return it.value, (d.value if want_dot else None)


def FermiAction(D, parameters_action):
    Nf = parameters_action.get("Nf", None)
    if D.kind == L.STAGGERED and Nf not in (None, 4, 8):
        return RHMCFermiAction(D, parameters_action)
    return FermiActionB200(D, parameters_action)


@contextlib.contextmanager
def _even_site_solves(fa):
    """while an even-site (staggered Nf = 4) action solves (D^dag D) X = eta: route the CG to checkerboarded half fields
    (lqcd_solve_staggered_even: half the sites, same iterates).  parameters_action["half_field_solver"] = False keeps the
    full-lattice CG."""
    on = bool(getattr(fa, "even_only", False) and getattr(fa, "half_field_solver", False) and fa.D.ctx.dist is None)
    if on:
        fa.D.ctx.call("lqcd_set_staggered_even_solve", 1)
    try:
        yield
    finally:
        if on:
            fa.D.ctx.call("lqcd_set_staggered_even_solve", 0)


def _even_site_xi_(fa, D, xi: FermionField):
    """xi <- D (D^dag D)^-1 P_even D^dag xi (see FermiActionB200): afterwards eta = P_even D^dag xi and xi^dag xi = S_f(eta)"""
    X, eta = fa._temporary_fermionfields[0], fa._temporary_fermionfields[3]
    mul_(eta, adjoint(D), xi)
    mask_parity_(eta, 0)
    clear_fermion_(X)
    with _even_site_solves(fa):
        fa.last = solve_DinvX_(X, DdagD(D), eta)
    mul_(xi, D, X)


def gauss_sampling_in_action_(xi: FermionField, U, fa, seed=112, _bound=False):     # standardMD.jl:95
    gauss_distribution_fermion_(xi, seed)
    if fa.even_only:
        _even_site_xi_(fa, fa.D if _bound else fa.D(U), xi)


def sample_pseudofermions_(eta: FermionField, U, fa, xi: FermionField):   # standardMD.jl:96
    """Wilson / staggered Nf = 8: eta = D^dag xi; staggered Nf = 4: odd sites of eta zeroed afterwards
    (App. C.7); RHMC: eta = (D^dag D)^{Nf/16} xi by the rational approximation."""
    fa.D(U)
    if isinstance(fa, RHMCFermiAction):
        it, _ = fa.rational_apply_(eta, fa.rhmc.r_heatbath, xi)
        fa.last = {"iters": it}
        return
    mul_(eta, adjoint(fa.D), xi)
    if fa.even_only:
        mask_parity_(eta, 0)


def evaluate_FermiAction(fa, U, eta: FermionField) -> float:           # standardHMC.jl:69-71
    """S_f = eta^dag (D^dag D)^-1 eta: one CG solve (RHMC: eta^dag (D^dag D)^{-Nf/8} eta, one multi-shift solve)."""
    D = fa.D(U)
    if isinstance(fa, RHMCFermiAction):
        it, S = fa.rational_apply_(fa._temporary_fermionfields[0], fa.rhmc.r_action, eta, want_dot=True)
        fa.last = {"iters": it, "action": S}
        return S
    X = fa._temporary_fermionfields[0]
    clear_fermion_(X)
    with _even_site_solves(fa):
        fa.last = solve_DinvX_(X, DdagD(D), eta)
    return dot(eta, X).real


def calc_UdSfdU_(UdSfdU: np.ndarray, fa, U, eta: FermionField):         # AbstractMD.jl:129
    """Fermion MD force: X = (D^dag D)^-1 eta (CG), Y = D X, per-mu colour outer products.
    UdSfdU: complex128[4,NT,NZ,NY,NX,3,3] in the link layout, overwritten.  RHMC: sum_j alpha_j force(X_j, Y_j) with the
    shifted solutions of ONE multi-shift CG, accumulated on the device."""
    D = fa.D(U)
    out = (C.c_void_p * 4)(*[UdSfdU[mu].ctypes.data for mu in range(4)])
    if isinstance(fa, RHMCFermiAction):
        _, a, b = fa._coeffs(fa.rhmc.r_action)
        it = C.c_int(0)
        D.ctx.call("lqcd_fermion_force_rational", C.byref(D.op), eta.h, a.ctypes.data_as(L.pdbl), b.ctypes.data_as(L.pdbl), len(a),
                   D.eps, D.maxsteps, out, 0, C.byref(it))
        fa.last = {"iters": it.value}
        return fa.last
    X = fa._temporary_fermionfields[0]
    clear_fermion_(X)
    it, act = C.c_int(0), C.c_double(0.0)
    with _even_site_solves(fa):
        D.ctx.call("lqcd_fermion_force", C.byref(D.op), eta.h, X.h, D.eps, D.maxsteps, out, 0, C.byref(it), C.byref(act))
    fa.last = {"iters": it.value, "action": act.value}
    return fa.last


# ---------------------------------------------------------------------------------------------------
# gauge-sector molecular dynamics on the device (src/md/AbstractMD.jl:78-135, src/md/standardMD.jl:103-165,
# src/updates/standardHMC.jl:41-91): links, momenta and pseudofermions stay in HBM for a whole trajectory
# ---------------------------------------------------------------------------------------------------
def _link_ptrs(a):
    return (C.c_void_p * 4)(*[a[mu].ctypes.data for mu in range(4)])


def gauss_distribution_momenta_(ctx: Context, seed=113):             # gauss_distribution!(md.p), standardMD.jl:86
    ctx.call("lqcd_md_momenta_gaussian", int(seed))


def set_momenta_(ctx: Context, P: np.ndarray):
    """P: complex128[4,NT,NZ,NY,NX,3,3] anti-Hermitian traceless matrices in the link layout"""
    P = np.ascontiguousarray(P, dtype=np.complex128)
    ctx.call("lqcd_md_momenta_upload", _link_ptrs(P), 0)


def get_momenta(ctx: Context) -> np.ndarray:
    NX, NY, NZ, NT = ctx.local_dims
    P = np.zeros((4, NT, NZ, NY, NX, 3, 3), dtype=np.complex128)
    ctx.call("lqcd_md_momenta_download", _link_ptrs(P), 0)
    return P


def get_links(ctx: Context) -> np.ndarray:
    NX, NY, NZ, NT = ctx.local_dims
    U = np.zeros((4, NT, NZ, NY, NX, 3, 3), dtype=np.complex128)
    ctx.call("lqcd_gauge_download", _link_ptrs(U), 3, 0)
    return U


def kinetic_energy(ctx: Context) -> float:                            # md.p * md.p / 2, standardHMC.jl:47
    out = C.c_double()
    ctx.call("lqcd_md_kinetic", C.byref(out))
    return out.value


def gauge_action(ctx: Context, beta: float) -> float:                 # -evaluate_GaugeAction(gauge_action, U) / NC, standardHMC.jl:49-50
    out = C.c_double()
    ctx.call("lqcd_md_gauge_action", float(beta), C.byref(out))
    return out.value


def U_update_(ctx: Context, eps_dtau: float):                         # U_update!(U, p, eps, md) with eps*md.dtau folded in
    ctx.call("lqcd_md_update_u", float(eps_dtau))


def P_update_(ctx: Context, eps_dtau: float, beta: float):            # P_update!
    ctx.call("lqcd_md_update_p", float(eps_dtau), float(beta))


def P_update_fermion_(D: DiracOperator, eta: FermionField, eps_dtau: float, rational=None) -> int:      # P_update_fermion!
    """rational: RationalApprox of the RHMC action (fa.rhmc.r_action) -> force of eta^dag r(D^dag D) eta"""
    it = C.c_int(0)
    if rational is not None:
        _, a, b = RHMCFermiAction._coeffs(rational)
        D.ctx.call("lqcd_md_update_p_fermion_rational", C.byref(D.op), eta.h, a.ctypes.data_as(L.pdbl), b.ctypes.data_as(L.pdbl), len(a),
                   float(eps_dtau), D.eps, D.maxsteps, C.byref(it))
        return it.value
    D.ctx.call("lqcd_md_update_p_fermion", C.byref(D.op), eta.h, float(eps_dtau), D.eps, D.maxsteps, C.byref(it))
    return it.value


def runMD_(ctx: Context, beta, dtau, MDsteps, D: DiracOperator = None, eta: FermionField = None, SextonWeingargten=False, Nsw=2,
           rational=None) -> int:
    """runMD!(U, md) on the device-resident links (standardMD.jl:103-165); returns the CG iterations spent in fermion forces.
    rational: RationalApprox of the RHMC action -> every fermion force is one multi-shift CG (lqcd_md_trajectory_rational)"""
    its = C.c_longlong(0)
    nsw = int(Nsw) if SextonWeingargten else 0
    if rational is not None:
        _, a, b = RHMCFermiAction._coeffs(rational)
        ctx.call("lqcd_md_trajectory_rational", C.byref(D.op), eta.h, a.ctypes.data_as(L.pdbl), b.ctypes.data_as(L.pdbl), len(a),
                 float(beta), float(dtau), int(MDsteps), nsw, D.eps, D.maxsteps, C.byref(its))
        return its.value
    ctx.call("lqcd_md_trajectory", C.byref(D.op) if D is not None else None, eta.h if eta is not None else None,
             float(beta), float(dtau), int(MDsteps), nsw, D.eps if D is not None else 0.0, D.maxsteps if D is not None else 1, C.byref(its))
    return its.value


def hmc_update_(U: Gaugefields, beta, dtau, MDsteps, fa=None, SextonWeingargten=False, Nsw=2, rng=None, seed=1):
    """update!(updatemethod::StandardHMC, U) (standardHMC.jl:41-91) with the molecular dynamics on the device: upload U once,
    sample p / xi / eta, S_old, runMD!, S_new, Metropolis test on the host, and U is overwritten only when accepted.
    Fermion actions: plain HMC (Wilson, staggered Nf = 4, 8) and RHMC (staggered, other Nf: heat bath, force and action through
    the rational approximations, one multi-shift CG each); returns (accepted, S_new - S_old, info)."""
    ctx = U.context()
    is_rhmc = fa is not None and isinstance(fa, RHMCFermiAction)
    rng = rng or np.random.default_rng(seed)
    if fa is not None:
        D = fa.D(U)                                                   # uploads the links
    else:
        D = None
        ctx.call("lqcd_gauge_upload", _link_ptrs(U.data), 3, 0)
    gauss_distribution_momenta_(ctx, int(rng.integers(1 << 62)))
    S_old = kinetic_energy(ctx) + gauge_action(ctx, beta)
    eta = None
    if fa is not None:
        xi, eta = fa._temporary_fermionfields[1], fa._temporary_fermionfields[2]
        gauss_sampling_in_action_(xi, U, fa, seed=int(rng.integers(1 << 62)), _bound=True)    # D is bound to U already
        if is_rhmc:
            fa.rational_apply_(eta, fa.rhmc.r_heatbath, xi)           # eta = (D^dag D)^{Nf/16} xi
        else:
            mul_(eta, adjoint(D), xi)                                 # sample_pseudofermions! without re-uploading U
        if fa.even_only:
            mask_parity_(eta, 0)
        if is_rhmc:
            # S_old with the SAME functional as S_new (eta^dag r_action(DdagD) eta): heat bath and action are two independent
            # fits (each ~1e-7), so dot(xi, xi) would leave a systematic dH of ~1e-7 * S_f that grows with the volume
            S_old += fa.rational_apply_(fa._temporary_fermionfields[0], fa.rhmc.r_action, eta, want_dot=True)[1]
        else:
            S_old += dot(xi, xi).real                                 # standardHMC.jl:54
    with _even_site_solves(fa):
        its = runMD_(ctx, beta, dtau, MDsteps, D, eta, SextonWeingargten, Nsw, rational=fa.rhmc.r_action if is_rhmc else None)
    S_new = kinetic_energy(ctx) + gauge_action(ctx, beta)
    if is_rhmc:
        S_new += fa.rational_apply_(fa._temporary_fermionfields[0], fa.rhmc.r_action, eta, want_dot=True)[1]
    elif fa is not None:
        X = fa._temporary_fermionfields[0]
        clear_fermion_(X)
        with _even_site_solves(fa):
            solve_DinvX_(X, DdagD(D), eta)
        S_new += dot(eta, X).real
    accept = bool(np.exp(min(0.0, S_old - S_new)) >= rng.random())    # exp(Sold - Snew) >= rand(), standardHMC.jl:79
    if accept:
        U.data[...] = get_links(ctx)
    elif D is not None:
        D(U)                                  # rejected: the device still holds the evolved links -- restore U (and the clover term)
    else:
        ctx.call("lqcd_gauge_upload", _link_ptrs(U.data), 3, 0)
    ctx.binding_epoch = getattr(ctx, "binding_epoch", 0) + 1          # other operators of this context re-bind their own U on next use
    if D is not None:
        D._bound_epoch = ctx.binding_epoch                            # ... D itself is in sync with U (accepted: downloaded; rejected: restored)
    return accept, S_new - S_old, {"cg_iters": its, "S_old": S_old, "S_new": S_new}

"""
ctypes binding of liblqcd_b200.so (include/lqcd_b200.h).  This is the Python twin of the Julia ``ccall``
shim (latticeqcd.jl_b200/julia/LQCDB200.jl): same symbols, same argument meaning, same error rule
(non-zero status -> exception carrying lqcd_last_error).  There is NO fallback: if the CUDA library is
missing or no B200 is visible, loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parents[1]
SO_PATH = Path(os.environ["LQCD_B200_LIB"]) if os.environ.get("LQCD_B200_LIB") else PKG / "liblqcd_b200.so"
HEADER = PKG.parent / "include" / "lqcd_b200.h"

LQCD_OK, ERR_ARG, ERR_CUDA, ERR_COMM, ERR_NOCONV, ERR_NOGPU, ERR_STATE = range(7)
WILSON, STAGGERED = 0, 1
OP_D, OP_DDAG, OP_DDAGD = 0, 1, 2
SOLVER_CG, SOLVER_CGNR, SOLVER_BICGSTAB = 0, 1, 2
IO_ILDG, IO_BRIDGETEXT = 0, 1
IPC_HANDLE_BYTES = 256


class LqcdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"liblqcd_b200 error {code}: {msg}")
        self.code = code


class NotConverged(LqcdError):
    """Mirror of upstream's error("The CG is not converged!")."""


class LqcdOp(C.Structure):
    _fields_ = [("kind", C.c_int), ("kappa", C.c_double), ("r", C.c_double), ("mass", C.c_double),
                ("csw", C.c_double), ("bc", C.c_double * 4)]


_lib = None

vp, i32, u64, dbl, sz = C.c_void_p, C.c_int, C.c_uint64, C.c_double, C.c_size_t
pi32, pdbl, pvp, pop = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_void_p), C.POINTER(LqcdOp)

# name -> (restype, argtypes).  Must list every symbol include/lqcd_b200.h declares (tests check this).
SIGNATURES = {
    "lqcd_ctx_create": (i32, [pi32, pi32, i32, i32, pvp]),
    "lqcd_ctx_destroy": (i32, [vp]),
    "lqcd_last_error": (C.c_char_p, [vp]),
    "lqcd_abi_version": (i32, []),
    "lqcd_local_dims": (i32, [vp, pi32, pi32]),
    "lqcd_synchronize": (i32, [vp]),
    "lqcd_host_register": (i32, [vp, vp, sz]),
    "lqcd_host_unregister": (i32, [vp, vp]),
    "lqcd_gauge_upload": (i32, [vp, pvp, i32, i32]),
    "lqcd_gauge_download": (i32, [vp, pvp, i32, i32]),
    "lqcd_gauge_random": (i32, [vp, u64, dbl]),
    "lqcd_gauge_plaquette": (i32, [vp, pdbl]),
    "lqcd_fermion_alloc": (i32, [vp, i32, pvp]),
    "lqcd_fermion_free": (i32, [vp, vp]),
    "lqcd_fermion_upload": (i32, [vp, vp, vp, i32]),
    "lqcd_fermion_download": (i32, [vp, vp, vp, i32]),
    "lqcd_fermion_zero": (i32, [vp, vp]),
    "lqcd_fermion_copy": (i32, [vp, vp, vp]),
    "lqcd_fermion_gaussian": (i32, [vp, vp, u64]),
    "lqcd_fermion_z4": (i32, [vp, vp, u64]),
    "lqcd_fermion_point_source": (i32, [vp, vp, pi32, i32, i32]),
    "lqcd_fermion_mask_parity": (i32, [vp, vp, i32]),
    "lqcd_blas_axpy": (i32, [vp, dbl, dbl, vp, vp]),
    "lqcd_blas_xpby": (i32, [vp, vp, dbl, dbl, vp]),
    "lqcd_blas_scale": (i32, [vp, dbl, dbl, vp]),
    "lqcd_blas_dot": (i32, [vp, vp, vp, pdbl]),
    "lqcd_blas_norm2": (i32, [vp, vp, pdbl]),
    "lqcd_dslash": (i32, [vp, pop, vp, vp, i32]),
    "lqcd_dslash_host": (i32, [vp, pop, vp, vp, vp, vp, i32, i32]),
    "lqcd_clover_term": (i32, [vp, pop, vp]),
    "lqcd_solve": (i32, [vp, pop, vp, vp, i32, i32, dbl, i32, pi32, pdbl, pdbl]),
    "lqcd_solve_eo": (i32, [vp, pop, vp, vp, i32, i32, dbl, i32, pi32, pdbl, pdbl]),
    "lqcd_solve_staggered_even": (i32, [vp, pop, vp, vp, dbl, i32, pi32, pdbl]),
    "lqcd_set_staggered_even_solve": (i32, [vp, i32]),
    "lqcd_multishift_cg": (i32, [vp, pop, pvp, vp, pdbl, i32, dbl, i32, pi32, pdbl]),
    "lqcd_io_read_gauge": (i32, [C.c_char_p, i32, pi32, i32, pvp, i32]),
    "lqcd_io_write_gauge": (i32, [C.c_char_p, i32, pi32, i32, pvp, i32]),
    "lqcd_gauge_load": (i32, [vp, C.c_char_p, i32]),
    "lqcd_gauge_save": (i32, [vp, C.c_char_p, i32]),
    "lqcd_dslash_multi": (i32, [vp, pop, pvp, pvp, i32, i32]),
    "lqcd_solve_multi": (i32, [vp, pop, pvp, pvp, i32, i32, i32, dbl, i32, pi32, pdbl]),
    "lqcd_fermion_force": (i32, [vp, pop, vp, vp, dbl, i32, pvp, i32, pi32, pdbl]),
    "lqcd_fermion_force_xy": (i32, [vp, pop, vp, vp, dbl, i32]),
    "lqcd_fermion_force_download": (i32, [vp, pvp, i32]),
    "lqcd_fermion_force_rational": (i32, [vp, pop, vp, pdbl, pdbl, i32, dbl, i32, pvp, i32, pi32]),
    "lqcd_rational_apply": (i32, [vp, pop, vp, vp, dbl, pdbl, pdbl, i32, dbl, i32, pi32, pdbl]),
    "lqcd_md_momenta_gaussian": (i32, [vp, u64]),
    "lqcd_md_momenta_upload": (i32, [vp, pvp, i32]),
    "lqcd_md_momenta_download": (i32, [vp, pvp, i32]),
    "lqcd_md_kinetic": (i32, [vp, pdbl]),
    "lqcd_md_gauge_action": (i32, [vp, dbl, pdbl]),
    "lqcd_md_update_u": (i32, [vp, dbl]),
    "lqcd_md_update_p": (i32, [vp, dbl, dbl]),
    "lqcd_md_update_p_fermion": (i32, [vp, pop, vp, dbl, dbl, i32, pi32]),
    "lqcd_md_trajectory": (i32, [vp, pop, vp, dbl, dbl, i32, i32, dbl, i32, C.POINTER(C.c_longlong)]),
    "lqcd_md_update_p_fermion_rational": (i32, [vp, pop, vp, pdbl, pdbl, i32, dbl, dbl, i32, pi32]),
    "lqcd_md_trajectory_rational": (i32, [vp, pop, vp, pdbl, pdbl, i32, dbl, dbl, i32, i32, dbl, i32, C.POINTER(C.c_longlong)]),
    "lqcd_comm_export": (i32, [vp, vp]),
    "lqcd_comm_connect": (i32, [vp, vp]),
    "lqcd_decompose": (i32, [pi32, pi32, i32, pi32, pi32, pi32, pi32]),
    "lqcd_launch_count": (i32, [vp, C.POINTER(u64)]),
    "lqcd_time_dslash": (i32, [vp, pop, vp, vp, i32, i32, i32, pdbl, pdbl]),
    "lqcd_stream": (i32, [vp, pvp]),
}


def load() -> C.CDLL:
    """Load the CUDA library.  Raises if it has not been built -- the product path never degrades."""
    global _lib
    if _lib is None:
        if not SO_PATH.exists():
            raise FileNotFoundError(
                f"{SO_PATH} is missing: build it with `python latticeqcd.jl_b200/build.py` "
                "(or __graft_entry__.build()).  There is no CPU fallback.")
        lib = C.CDLL(str(SO_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export it
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(ctx, status: int):
    if status == LQCD_OK:
        return
    msg = load().lqcd_last_error(ctx)
    msg = msg.decode(errors="replace") if msg else "?"
    raise (NotConverged if status == ERR_NOCONV else LqcdError)(status, msg)

"""Round-2 helper (needs a B200): correctness + timing of the experimental t-marching kernel (LQCD_WILSON_KERNEL=3)
against the CPU oracle and against kernel 1.  The library caches the kernel choice per process, so this script re-executes
itself once per kernel family.   python tools/check_kernel3.py"""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "latticeqcd.jl_b200")]


def child(family):
    import numpy as np
    import lqcd_b200 as q
    from lqcd_b200 import _lib as L
    from oracle import oracle as orc
    ok = True
    for dims in [(32, 4, 4, 4), (8, 8, 8, 8), (16, 8, 4, 8), (32, 8, 8, 16)]:
        Uh = orc.random_su3(dims, seed=5)
        U = q.gaugefields_from_array(Uh)
        x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
        D = q.Dirac_operator(U, x, {"Dirac_operator": "Wilson", "κ": 0.12, "r": 1.0, "boundarycondition": [1, 1, 1, -1]})
        src = orc.gaussian_field(dims, orc.WILSON, seed=13)
        x.from_host(src)
        y = q.similar(x)
        op = orc.make_op(dims, kappa=0.12)
        for A, m, nm in ((D, orc.D, "D"), (q.adjoint(D), orc.DDAG, "Ddag")):
            q.mul_(y, A, x)
            want = orc.apply(op, orc.WILSON, m, Uh, src)
            err = np.abs(y.to_host() - want).max() / np.abs(want).max()
            print(f"kernel {family} {dims} {nm}: rel err {err:.2e}", flush=True)
            ok &= err < 1e-13
        sol = q.similar(x)
        q.clear_fermion_(sol)
        D.eps, D.maxsteps = 1e-18, 3000
        info = q.solve_DinvX_(sol, q.DdagD(D), x)
        ref = orc.cg(op, orc.WILSON, Uh, src, eps=1e-18)
        print(f"kernel {family} {dims} CG iters {info['iters']} (oracle {ref['iters']})", flush=True)
        ok &= info["iters"] == ref["iters"]
    for dims in [(32, 32, 32, 32), (32, 32, 16, 8)]:
        ctx = q.get_context(dims)
        ctx.call("lqcd_gauge_random", 111, -1.0)
        a, b = q.FermionField(ctx, L.WILSON), q.FermionField(ctx, L.WILSON)
        q.gauss_distribution_fermion_(a, 1)
        op = L.LqcdOp(); op.kind = L.WILSON; op.kappa = 0.12; op.r = 1.0
        for i, v in enumerate([1, 1, 1, -1]):
            op.bc[i] = v
        mean, mn = C.c_double(), C.c_double()
        for flush in (0, 1):
            ctx.call("lqcd_time_dslash", C.byref(op), b.h, a.h, 0, 5, flush, C.byref(mean), C.byref(mn))
            ctx.call("lqcd_time_dslash", C.byref(op), b.h, a.h, 0, 30, flush, C.byref(mean), C.byref(mn))
            V = dims[0] * dims[1] * dims[2] * dims[3]
            print(f"kernel {family} {dims} flush={flush}: {mean.value * 1e3:.1f} us  {960 * V / mean.value / 1e6:.0f} GB/s", flush=True)
    print(f"kernel {family}: {'OK' if ok else 'MISMATCH'}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    else:
        for fam in ("1", "3"):
            env = dict(os.environ, LQCD_WILSON_KERNEL=fam)
            subprocess.run([sys.executable, __file__, fam], env=env)

"""Correctness worker for the Wilson kernel families (LQCD_WILSON_KERNEL=4: t-marching TMA kernel csrc/wilson_tmarch.cu; unset: the
register-resident default with two-row links, LQCD_LINKS12=0 full links):
operator applications and a CG solve (fused |Dp|^2 / residual-update epilogues) against the oracle on lattices whose tiling
qualifies for the kernel, with several chunkings (LQCD_TM_CHUNKS).  The library caches its knobs per process, hence a worker
(tests/test_emu_preflight.py under emulation with both bulk-copy completion schedules, tests/test_gpu_extended.py on the B200)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "latticeqcd.jl_b200")]
import lqcd_b200 as q                     # noqa: E402
from oracle import oracle as orc          # noqa: E402


def main():
    ok = True
    for dims in [(32, 4, 4, 4), (8, 8, 8, 8), (16, 8, 4, 8), (8, 8, 4, 6), (16, 4, 4, 2)]:
        Uh = orc.random_su3(dims, seed=5)
        U = q.gaugefields_from_array(Uh)
        x = q.Initialize_pseudofermion_fields(U[0], "Wilson")
        D = q.Dirac_operator(U, x, {"Dirac_operator": "Wilson", "κ": 0.12, "r": 1.0, "boundarycondition": [1, 1, 1, -1]})
        src = orc.gaussian_field(dims, orc.WILSON, seed=13)
        x.from_host(src)
        y = q.similar(x)
        op = orc.make_op(dims, kappa=0.12)
        for A, m, nm in ((D, orc.D, "D"), (q.adjoint(D), orc.DDAG, "Ddag"), (q.DdagD(D), orc.DDAGD, "DdagD")):
            q.mul_(y, A, x)
            want = orc.apply(op, orc.WILSON, m, Uh, src)
            err = np.abs(y.to_host() - want).max() / np.abs(want).max()
            print(f"tmarch {dims} {nm}: rel err {err:.2e}", flush=True)
            ok &= bool(err < 1e-13)
        sol = q.similar(x)
        q.clear_fermion_(sol)
        D.eps, D.maxsteps = 1e-18, 3000
        info = q.solve_DinvX_(sol, q.DdagD(D), x)
        ref = orc.cg(op, orc.WILSON, Uh, src, eps=1e-18)
        dev = np.abs(sol.to_host() - ref["x"]).max() / np.abs(ref["x"]).max()
        print(f"tmarch {dims} CG iters {info['iters']} (oracle {ref['iters']}), solution rel dev {dev:.2e}", flush=True)
        ok &= info["iters"] == ref["iters"] and bool(dev < 1e-10)
    print("TMARCH OK" if ok else "TMARCH MISMATCH", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

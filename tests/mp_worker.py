"""
Multi-rank parity worker (launched by tests/test_multirank.py through torch.distributed.run, gloo backend so
that several ranks may share one GPU on a single-GPU box: the CUDA-IPC / peer-memory path is identical).

Every rank builds the same global synthetic fields, uploads its LOCAL block, runs the operator and the
solvers through the C ABI, and rank 0 compares the gathered result with the CPU oracle on the global lattice.
Rank 0 computes ALL oracle references first (the other ranks wait at a host barrier): between two
collective library calls the ranks must not be seconds apart, the device-side waits have a timeout.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "latticeqcd.jl_b200")]
import lqcd_b200 as q                     # noqa: E402
from oracle import oracle as orc          # noqa: E402


def main():
    dims = tuple(int(v) for v in sys.argv[1].split("x"))
    pg = tuple(int(v) for v in sys.argv[2].split("x"))
    kind_name = sys.argv[3]
    # variant: "" = operator + CG + CGNR (the subset that has run on hardware); "full" = + multi-shift CG + fermion force;
    # "clover" = Wilson-clover operator (peer-mapped links for the clover leaves) + multi-shift CG
    variant = sys.argv[4] if len(sys.argv) > 4 else ""
    csw = 1.5612 if variant == "clover" else 0.0
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert world == int(np.prod(pg)), f"procgrid {pg} needs {int(np.prod(pg))} ranks, got {world}"
    ndev = torch.cuda.device_count()
    dev = int(os.environ.get("LOCAL_RANK", rank)) % max(ndev, 1)     # ndev = 0 only under tests/emu (no CUDA device)
    kind = orc.WILSON if kind_name == "Wilson" else orc.STAGGERED
    Ug = orc.random_su3(dims, seed=17, eps=0.4)
    src = orc.gaussian_field(dims, kind, seed=23)
    op = orc.make_op(dims, kappa=0.125, mass=0.2, csw=csw)
    shifts = [0.05, 0.4, 1.7]
    ref = {}
    if rank == 0:
        orc.set_threads(max(1, (os.cpu_count() or 8) // 2))
        if csw:
            keep_clover = orc.clover_build(op, Ug)      # attaches the term to op (the array must stay alive)
        for m, nm in ((orc.D, "D"), (orc.DDAG, "Ddag"), (orc.DDAGD, "DdagD")):
            ref[nm] = orc.apply(op, kind, m, Ug, src)
        ref["dot"] = np.vdot(src, ref["DdagD"])
        ref["cg"] = orc.cg(op, kind, Ug, src, eps=1e-18, maxsteps=2000, hist=True)
        ref["cgnr"] = orc.cgnr(op, kind, Ug, src, eps=1e-18, maxsteps=2000, hist=True)
        ref["ms"] = orc.mscg(op, kind, Ug, src, shifts, eps=1e-18, maxsteps=2000)
    dist.barrier()

    ctx = q.get_context(dims, procgrid=pg, rank=rank, device=dev)
    q.connect_ranks(ctx, dist)
    (lx, ly, lz, lt), (ox, oy, oz, ot) = ctx.local_dims, ctx.origin
    sl = (slice(ot, ot + lt), slice(oz, oz + lz), slice(oy, oy + ly), slice(ox, ox + lx))
    Ul = np.ascontiguousarray(Ug[(slice(None),) + sl])
    U = q.gaugefields_from_array(Ul, global_dims=dims, procgrid=pg, rank=rank, device=dev)
    x = q.Initialize_pseudofermion_fields(U[0], kind_name)
    params = {"Dirac_operator": ("WilsonClover" if csw else "Wilson") if kind == orc.WILSON else "staggered", "κ": 0.125, "mass": 0.2,
              "Clover_coefficient": csw,
              "eps_CG": 1e-18, "MaxCGstep": 2000, "boundarycondition": [1, 1, 1, -1]}
    D = q.Dirac_operator(U, x, params)
    loc = (lambda a: np.ascontiguousarray(a[(slice(None),) + sl])) if kind == orc.WILSON else (lambda a: np.ascontiguousarray(a[sl]))
    x.from_host(loc(src))
    y = q.similar(x)

    def gather(f):
        h = torch.from_numpy(f.to_host().view(np.float64))      # gloo has no complex dtypes
        out = [torch.empty_like(h) for _ in range(world)] if rank == 0 else None
        dist.gather(h, out, dst=0)
        if rank != 0:
            return None
        full = np.zeros(orc.field_shape(dims, kind), dtype=complex)
        for r in range(world):
            (ld, og, _, _) = q.decompose(dims, pg, r)
            s2 = (slice(og[3], og[3] + ld[3]), slice(og[2], og[2] + ld[2]), slice(og[1], og[1] + ld[1]), slice(og[0], og[0] + ld[0]))
            if kind == orc.WILSON:
                full[(slice(None),) + s2] = out[r].numpy().view(np.complex128)
            else:
                full[s2] = out[r].numpy().view(np.complex128)
        return full

    fails = []

    def same_iters(k, r, eps=1e-18):
        """identical iteration count, except when the oracle's own |r|^2 sits within 5 % of eps at the stopping step: the
        oracle's OpenMP reductions are not bit-reproducible run to run and such a borderline case flips by one."""
        if k == r["iters"]:
            return True
        h = r["hist"]
        near = [abs(h[i] / eps - 1.0) < 0.05 for i in (k, r["iters"]) if 0 <= i < len(h)]
        return abs(k - r["iters"]) == 1 and any(near)

    def check(name, got, want, tol):
        if rank == 0:
            err = np.abs(got - want).max() / np.abs(want).max()
            print(f"[mp {world} ranks {pg}] {kind_name} {name}: rel err {err:.2e}", flush=True)
            if not err < tol:
                fails.append(name)

    for A, nm in ((D, "D"), (q.adjoint(D), "Ddag"), (q.DdagD(D), "DdagD")):
        q.mul_(y, A, x)
        got = gather(y)
        if rank == 0:
            check(nm, got, ref[nm], 1e-13)
    d = q.dot(x, y)                                     # global dot product (in-kernel all-reduce)
    if rank == 0:
        e = abs(d - ref["dot"]) / abs(ref["dot"])
        print(f"[mp] dot rel err {e:.2e}", flush=True)
        if e > 1e-10:
            fails.append("dot")
    for name, A, key in (("CG", q.DdagD(D), "cg"), ("CGNR", D, "cgnr")):
        sol = q.similar(x)
        q.clear_fermion_(sol)
        info = q.solve_DinvX_(sol, A, x)
        got = gather(sol)
        if rank == 0:
            print(f"[mp] {name} iters {info['iters']} (oracle {ref[key]['iters']})", flush=True)
            if not same_iters(info["iters"], ref[key]):
                fails.append(f"{name} iters")
            check(f"{name} solution", got, ref[key]["x"], 1e-9)
    if not variant:                                     # the hardware-verified subset stops here (test_multirank_parity)
        shifts = []
    ys = [q.similar(x) for _ in shifts]                 # multi-shift CG across ranks (zeta recurrences from all-reduced scalars)
    info = q.shiftedcg_(ys, D, x, shifts) if shifts else {"iters": None}
    for j in range(len(shifts)):
        got = gather(ys[j])
        if rank == 0:
            check(f"multishift x[{j}]", got, ref["ms"]["xs"][j], 1e-9)
    if rank == 0 and shifts:
        print(f"[mp] multishift iters {info['iters']} (oracle {ref['ms']['iters']})", flush=True)
        if info["iters"] != ref["ms"]["iters"]:
            fails.append("multishift iters")
    if variant == "full":                               # fermion force across ranks (force halo slots, force.cu)
        fa = q.FermiAction(D, {"Nf": 8 if kind == orc.STAGGERED else 2})
        F = np.zeros((4, lt, lz, ly, lx, 3, 3), dtype=complex)
        finfo = q.calc_UdSfdU_(F, fa, U, x)
        h = torch.from_numpy(F.view(np.float64))
        outF = [torch.empty_like(h) for _ in range(world)] if rank == 0 else None
        dist.gather(h, outF, dst=0)
        if rank == 0:
            full = np.zeros((4,) + tuple(dims[::-1]) + (3, 3), dtype=complex)
            for r in range(world):
                (ld, og, _, _) = q.decompose(dims, pg, r)
                s2 = (slice(og[3], og[3] + ld[3]), slice(og[2], og[2] + ld[2]), slice(og[1], og[1] + ld[1]), slice(og[0], og[0] + ld[0]))
                full[(slice(None),) + s2] = outF[r].numpy().view(np.complex128)
            Xr = ref["cg"]["x"]
            Fr = orc.force(op, kind, Ug, Xr, orc.apply(op, kind, orc.D, Ug, Xr))
            check("fermion force", full, Fr, 1e-8)
            act = np.vdot(src, Xr).real
            if abs(finfo["action"] - act) > 1e-9 * abs(act):
                fails.append("action")
    flag = torch.tensor([len(fails)])
    dist.broadcast(flag, 0)
    dist.barrier()
    if rank == 0 and fails:
        print("FAILED:", fails, flush=True)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()

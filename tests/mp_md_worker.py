"""Multi-rank molecular-dynamics worker (tests/test_multirank.py, tests/test_emu_preflight.py): every rank holds its local block
of links and momenta; a Sexton-Weingarten trajectory with Wilson pseudofermions runs through lqcd_md_trajectory (staples and
plaquettes read the neighbour ranks' peer-mapped links, device-side barriers between sub-steps) and rank 0 compares the gathered
links / momenta / energies with the oracle's steps composed on the global lattice."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "latticeqcd.jl_b200"), str(ROOT / "tests")]
import lqcd_b200 as q                     # noqa: E402
from oracle import oracle as orc          # noqa: E402

BETA, KAPPA = 5.7, 0.12


def main():
    dims = tuple(int(v) for v in sys.argv[1].split("x"))
    pg = tuple(int(v) for v in sys.argv[2].split("x"))
    rhmc_mode = len(sys.argv) > 3 and sys.argv[3] == "rhmc"      # staggered Nf = 2 rational action instead of Wilson pseudofermions
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = int(os.environ.get("LOCAL_RANK", rank)) % max(torch.cuda.device_count(), 1)
    Ug = orc.random_su3(dims, seed=51, eps=0.4)
    Pg = orc.md_momenta(dims, seed=52)
    kind = orc.STAGGERED if rhmc_mode else orc.WILSON
    op = orc.make_op(dims, kappa=KAPPA, mass=0.5)
    xi = orc.gaussian_field(dims, kind, seed=53)
    eta_g = orc.apply(op, kind, orc.DDAG, Ug, xi)
    import test_md
    ra = test_md._rational_nf2() if rhmc_mode else None
    ref = {}
    if rank == 0:
        test_md.DIMS = dims
        f = (op, kind, eta_g, ra) if rhmc_mode else (op, kind, eta_g)
        ref["K0"], ref["Sg0"] = orc.md_kinetic(dims, Pg), orc.md_gauge_action(dims, Ug, BETA)
        ref["U"], ref["P"] = test_md._traj(Ug, Pg, 0.05, 2, nsw=4, fermion=f)
        ref["H0"], ref["H1"] = test_md._H(Ug, Pg, f), test_md._H(ref["U"], ref["P"], f)
    dist.barrier()
    ctx = q.get_context(dims, procgrid=pg, rank=rank, device=dev)
    q.connect_ranks(ctx, dist)
    (lx, ly, lz, lt), (ox, oy, oz, ot) = ctx.local_dims, ctx.origin
    sl = (slice(None), slice(ot, ot + lt), slice(oz, oz + lz), slice(oy, oy + ly), slice(ox, ox + lx))
    U = q.gaugefields_from_array(np.ascontiguousarray(Ug[sl]), global_dims=dims, procgrid=pg, rank=rank, device=dev)
    x = q.Initialize_pseudofermion_fields(U[0], "staggered" if rhmc_mode else "Wilson")
    D = q.Dirac_operator(U, x, {"Dirac_operator": "staggered" if rhmc_mode else "Wilson", "κ": KAPPA, "mass": 0.5, "eps_CG": 1e-22,
                                "MaxCGstep": 3000, "boundarycondition": [1, 1, 1, -1]})
    ctx.barrier()                                      # every rank has uploaded its links before anybody reads a neighbour's
    q.set_momenta_(ctx, np.ascontiguousarray(Pg[sl]))
    eta = q.similar(x).from_host(np.ascontiguousarray(eta_g[sl[1:] if rhmc_mode else sl]))    # staggered fields have no spin axis
    K0, Sg0 = q.kinetic_energy(ctx), q.gauge_action(ctx, BETA)
    its = q.runMD_(ctx, BETA, 0.05, 2, D, eta, SextonWeingargten=True, Nsw=4, rational=ra)
    X = q.similar(x)
    if rhmc_mode:
        fa = q.FermiAction(D, {"Nf": 2, "rational_lambda_min": 0.22, "rational_lambda_max": 17.0})
        Sf = fa.rational_apply_(X, ra, eta, want_dot=True)[1]
    else:
        q.clear_fermion_(X)
        q.solve_DinvX_(X, q.DdagD(D), eta)
        Sf = q.dot(eta, X).real
    H1 = q.kinetic_energy(ctx) + q.gauge_action(ctx, BETA) + Sf

    def gather(a):
        h = torch.from_numpy(np.ascontiguousarray(a).view(np.float64))
        out = [torch.empty_like(h) for _ in range(world)] if rank == 0 else None
        dist.gather(h, out, dst=0)
        if rank != 0:
            return None
        full = np.zeros((4,) + tuple(dims[::-1]) + (3, 3), dtype=complex)
        for r in range(world):
            (ld, og, _, _) = q.decompose(dims, pg, r)
            s2 = (slice(None), slice(og[3], og[3] + ld[3]), slice(og[2], og[2] + ld[2]), slice(og[1], og[1] + ld[1]), slice(og[0], og[0] + ld[0]))
            full[s2] = out[r].numpy().view(np.complex128)
        return full

    Ud, Pd = gather(q.get_links(ctx)), gather(q.get_momenta(ctx))
    fails = []
    if rank == 0:
        checks = {"kinetic": abs(K0 - ref["K0"]) / ref["K0"], "gauge action": abs(Sg0 - ref["Sg0"]) / abs(ref["Sg0"]),
                  "links": np.abs(Ud - ref["U"]).max(), "momenta": np.abs(Pd - ref["P"]).max(),
                  "H after": abs(H1 - ref["H1"]) / abs(ref["H1"])}
        for k, v in checks.items():
            print(f"[mp-md {world} ranks {pg}] {k}: {v:.2e}", flush=True)
            if not v < 1e-9:
                fails.append(k)
        print(f"[mp-md] dH = {ref['H1'] - ref['H0']:.4e}, CG iterations in forces {its}", flush=True)
    flag = torch.tensor([len(fails)])
    dist.broadcast(flag, 0)
    dist.barrier()
    if rank == 0 and fails:
        print("FAILED:", fails, flush=True)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()

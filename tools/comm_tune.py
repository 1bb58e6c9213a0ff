"""Development aid (multi-GPU): sweeps the halo-pack knobs of the Wilson / staggered Dslash inside ONE job per N and prints a table
with the in-kernel phase timeline (LQCD_COMM_TIMING=1).  Launch: torchrun --nproc-per-node N tools/comm_tune.py [lattice]"""
import ctypes as C, os, sys, io, contextlib
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "latticeqcd.jl_b200")]
os.environ["LQCD_COMM_TUNE"] = "1"
import numpy as np
import torch, torch.distributed as dist
import lqcd_b200 as q
from lqcd_b200 import _lib as L
import bench
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
dims = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "32x32x32x32").split("x"))
pg = bench.choose_procgrid(world)
if os.environ.get("LQCD_PROCGRID"):
    pg = tuple(int(v) for v in os.environ["LQCD_PROCGRID"].split(","))
ctx = q.get_context(dims, procgrid=pg, rank=rank, device=int(os.environ.get("LOCAL_RANK", rank)))
q.connect_ranks(ctx, dist)
ctx.call("lqcd_gauge_random", 111, -1.0)
mean, mn = C.c_double(), C.c_double()
V = int(np.prod(dims))
for kind, name, fps in ((L.WILSON, "wilson", 1368), (L.STAGGERED, "staggered", 582)):
    op = L.LqcdOp(); op.kind, op.kappa, op.r, op.mass = kind, 0.12, 1.0, 0.5
    for i, b in enumerate([1, 1, 1, -1]): op.bc[i] = b
    x, y = q.FermionField(ctx, kind), q.FermionField(ctx, kind)
    q.gauss_distribution_fermion_(x, 112)
    for mode in ("0", "1"):
        for spt in (["1", "4"] if mode == "0" else ["1", "2", "4", "8", "16", "32"]):
            os.environ["LQCD_SELF_PACK"], os.environ["LQCD_PACK_SPT"] = mode, spt
            os.environ["LQCD_COMM_TIMING"] = "0"
            ctx.call("lqcd_time_dslash", C.byref(op), y.h, x.h, L.OP_D, 20, 0, C.byref(mean), C.byref(mn))
            dist.barrier()
            ctx.call("lqcd_time_dslash", C.byref(op), y.h, x.h, L.OP_D, 300, 0, C.byref(mean), C.byref(mn))
            t = torch.tensor([mean.value], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
            nrm = q.dot(y, y).real
            os.environ["LQCD_COMM_TIMING"] = "1"          # same setting again with the phase stamps (printed by the library on stderr)
            dist.barrier()
            if rank == 0:
                print(f"N={world} {name} self_pack={mode} spt={spt:>2}: {float(t[0])*1e3:7.1f} us/apply  {fps*V/float(t[0])/1e6:9.0f} GFLOP/s  |Dx|^2={nrm:.6f}", flush=True)
            sys.stderr.flush()
            ctx.call("lqcd_time_dslash", C.byref(op), y.h, x.h, L.OP_D, 100, 0, C.byref(mean), C.byref(mn))
            sys.stderr.flush()
dist.barrier()
dist.destroy_process_group()

"""Multi-rank gauge file I/O worker (tests/test_multirank.py, tests/test_emu_preflight.py): every rank loads ONLY its block of a
global ILDG / BridgeText file straight into its device links (lqcd_gauge_load), the links equal the slice of the global field,
the plaquette (which reads neighbour ranks' links) equals the oracle's, and an ILDG file saved by all ranks together
(lqcd_gauge_save: each rank writes its own rows) is byte-identical to the input."""
import os
import sys
import tempfile
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
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    dev = int(os.environ.get("LOCAL_RANK", rank)) % max(torch.cuda.device_count(), 1)
    U = orc.random_su3(dims, seed=61, eps=0.4)
    tmp = Path(tempfile.gettempdir()) / f"lqcd_b200_mpio_{os.environ.get('MASTER_PORT', '0')}"
    if rank == 0:
        tmp.mkdir(exist_ok=True)
        q.save_binarydata(U, tmp / "in.ildg")
        q.save_textdata(U, tmp / "in.txt")
    dist.barrier()
    ctx = q.get_context(dims, procgrid=pg, rank=rank, device=dev)
    q.connect_ranks(ctx, dist)
    (lx, ly, lz, lt), (ox, oy, oz, ot) = ctx.local_dims, ctx.origin
    sl = (slice(None), slice(ot, ot + lt), slice(oz, oz + lz), slice(oy, oy + ly), slice(ox, ox + lx))
    fails = []
    for fmt, name in (("BridgeText", "in.txt"), ("ILDG", "in.ildg")):
        q.load_gaugefield_device_(ctx, tmp / name, fmt)
        if not np.array_equal(q.get_links(ctx), U[sl]):
            fails.append(f"{fmt} load")
    ctx.barrier()
    plaq = q.plaquette(ctx) if hasattr(q, "plaquette") else None
    if plaq is not None and abs(plaq - orc.plaquette(dims, U)) > 1e-12:
        fails.append("plaquette")
    q.save_gaugefield_device(ctx, tmp / "out.ildg", "ILDG")
    if rank == 0 and (tmp / "out.ildg").read_bytes() != (tmp / "in.ildg").read_bytes():
        fails.append("ILDG save")
    flag = torch.tensor([len(fails)])
    dist.all_reduce(flag)
    if fails:
        print(f"FAILED rank {rank}:", fails, flush=True)
    elif rank == 0:
        print(f"[mp-io {pg}] load (text, ILDG), plaquette {plaq}, collective ILDG save: ok", flush=True)
    dist.barrier()
    if rank == 0:
        for f in tmp.iterdir():
            f.unlink()
        tmp.rmdir()
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()

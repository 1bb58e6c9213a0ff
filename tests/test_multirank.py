"""Multi-rank tests.  CPU part: decomposition geometry and the gloo handle exchange (world_size 2).
GPU part: the real peer-memory halo exchange / in-kernel all-reduce against the oracle, with the ranks
sharing whatever GPUs exist (2-8 ranks also run on a single B200: CUDA IPC works within one device)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import lqcd_b200 as q

ROOT = Path(__file__).resolve().parents[1]


def run_ranks(n, script, *args, timeout=600):
    env = dict(os.environ)
    env.setdefault("OMP_NUM_THREADS", "4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + (os.getpid() % 2000)), str(script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


@pytest.mark.parametrize("dims,pg", [((8, 8, 8, 16), (1, 1, 1, 2)), ((32, 32, 32, 32), (1, 1, 2, 4)), ((8, 4, 6, 12), (2, 1, 3, 1))])
def test_decompose_tiles_the_lattice(dims, pg):
    n = int(np.prod(pg))
    seen = np.zeros(dims[::-1], dtype=int)
    for r in range(n):
        ld, og, lo, hi = q.decompose(dims, pg, r)
        assert all(ld[i] * pg[i] == dims[i] for i in range(4))
        seen[og[3]:og[3] + ld[3], og[2]:og[2] + ld[2], og[1]:og[1] + ld[1], og[0]:og[0] + ld[0]] += 1
        for mu in range(4):
            assert q.decompose(dims, pg, hi[mu])[2][mu] == r        # my upper neighbour's lower neighbour is me
            assert q.decompose(dims, pg, lo[mu])[3][mu] == r
            og_hi = q.decompose(dims, pg, hi[mu])[1]
            assert og_hi[mu] == (og[mu] + ld[mu]) % dims[mu]
    assert (seen == 1).all()
    with pytest.raises(ValueError):
        q.decompose((8, 8, 8, 9), (1, 1, 1, 2), 0)


def test_handle_exchange_gloo_world2(tmp_path):
    """host-side plumbing of lqcd_comm_export/connect: fixed-size blobs all-gathered in rank order (gloo, CPU)."""
    script = tmp_path / "w.py"
    script.write_text(f"""
import sys
sys.path[:0] = [{str(ROOT)!r}, {str(ROOT / 'latticeqcd.jl_b200')!r}]
import torch.distributed as dist
import lqcd_b200 as q
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
blob = bytes([r]) * 256
allb = q.exchange_handles(blob, dist)
assert len(allb) == 256 * w and all(allb[256 * i:256 * (i + 1)] == bytes([i]) * 256 for i in range(w)), allb[:8]
ld, og, lo, hi = q.decompose((8, 8, 8, 8), (1, 1, 1, w), r)
assert og[3] == r * 8 // w and lo[3] == (r - 1) % w and hi[3] == (r + 1) % w
dist.barrier(); dist.destroy_process_group()
""")
    res = run_ranks(2, script, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("dims,pg,kind", [
    ("8x8x8x16", "1x1x1x2", "Wilson"),
    ("8x8x8x16", "1x1x1x4", "Wilson"),
    ("8x8x8x8", "1x1x2x2", "Wilson"),
    ("8x8x8x8", "1x1x2x4", "staggered"),
    ("4x8x8x8", "1x2x1x2", "staggered"),
])
def test_multirank_parity(dims, pg, kind):
    n = int(np.prod([int(v) for v in pg.split("x")]))
    res = run_ranks(n, ROOT / "tests" / "mp_worker.py", dims, pg, kind, timeout=900)
    sys.stdout.write(res.stdout[-3000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("dims,pg,kind", [("8x8x8x16", "1x1x1x2", "Wilson"), ("8x8x8x8", "1x1x2x2", "staggered")])
def test_multirank_force_parity(dims, pg, kind):
    """same worker as test_multirank_parity; kept as a separate staged test because the worker now also checks the multi-rank
    fermion force and multi-shift CG, which have not run on hardware yet (tests/mp_worker.py "full")"""
    n = int(np.prod([int(v) for v in pg.split("x")]))
    res = run_ranks(n, ROOT / "tests" / "mp_worker.py", dims, pg, kind, "full", timeout=240)
    sys.stdout.write(res.stdout[-3000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("dims,pg", [("8x8x8x16", "1x1x1x2"), ("8x8x8x8", "1x1x2x2")])
def test_multirank_clover_parity(dims, pg):
    """Wilson-clover across ranks: the clover leaves read the neighbours' links through peer-mapped link arrays (clover.cu)"""
    n = int(np.prod([int(v) for v in pg.split("x")]))
    res = run_ranks(n, ROOT / "tests" / "mp_worker.py", dims, pg, "Wilson", "clover", timeout=240)
    sys.stdout.write(res.stdout[-3000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("dims,pg,action", [("8x8x8x16", "1x1x1x2", "wilson"), ("8x8x8x8", "1x1x2x2", "wilson"), ("8x8x8x8", "1x1x1x2", "rhmc")])
def test_multirank_md_trajectory(dims, pg, action):
    """Sexton-Weingarten trajectory across ranks (tests/mp_md_worker.py): Wilson pseudofermions, and the staggered Nf = 2 RHMC
    action (multi-shift CG + accumulated rational force per step; BASELINE config 5 in miniature)"""
    n = int(np.prod([int(v) for v in pg.split("x")]))
    res = run_ranks(n, ROOT / "tests" / "mp_md_worker.py", dims, pg, action, timeout=240)
    sys.stdout.write(res.stdout[-3000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.gpu
def test_multirank_gauge_io():
    """every rank loads only its block of a global ILDG / BridgeText file; collective ILDG save is byte-identical (tests/mp_io_worker.py)"""
    res = run_ranks(2, ROOT / "tests" / "mp_io_worker.py", "8x8x8x8", "1x1x1x2", timeout=240)
    sys.stdout.write(res.stdout[-2000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]

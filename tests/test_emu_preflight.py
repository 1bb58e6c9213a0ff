"""
Pre-flight of the device code on the GPU-less build container (CPU suite, NOT a parity claim).

tests/emu/ compiles the library's own .cu/.cuh sources with g++ against a SIMT emulator (fibers for the threads of a CTA,
CTAs in blockIdx order, fake CUDA runtime with red-zoned "device" memory, POSIX-shm CUDA-IPC).  This file runs the `-m gpu`
parity tests -- unchanged, through the same ctypes C-ABI binding -- against that emulated build in a subprocess with
LQCD_B200_LIB pointing at it.  What it buys: kernels written while no B200 was reachable (clover, even-odd, ...) are checked
against the oracle (indices, epilogues, reductions, finish ops, the host-side Krylov drivers, multi-process halo exchange and
in-kernel all-reduce protocol) before their first run on hardware.  What it does not show: anything about sm_100a code
generation, the GPU memory model, or performance -- the real gate stays `pytest -m gpu` on the B200.

The product never sees the emulated library: lqcd_b200/_lib.py loads liblqcd_b200.so (nvcc, sm_100a) unless the TEST sets
LQCD_B200_LIB, and nothing under latticeqcd.jl_b200/ references tests/emu.
"""
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests" / "emu"))


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    return build_emu.build()


def _env(emu_lib, **extra):
    env = dict(os.environ)
    env["LQCD_B200_LIB"] = str(emu_lib)
    env.setdefault("OMP_NUM_THREADS", "2")
    env.update(extra)
    return env


def test_translation_rejects_unknown_constructs():
    import build_emu
    with pytest.raises(build_emu.TranslateError):
        build_emu.translate('__device__ void f() { asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1;" :: "r"(0), "l"(0ull)); }')
    out = build_emu.translate("void g(int n) { k<1, 2><<<n, 128, 0, s>>>(a, f(b, c)); }")
    assert 'emu::launch(dim3(n), dim3(128), 0, s, [&]() { k<1, 2>(a, f(b, c)); }, "k")' in out


def test_gpu_suite_under_emulation(emu_lib):
    """every single-rank `-m gpu` test, xfail markers ignored (--runxfail): the not-yet-on-hardware kernels must pass here"""
    cmd = [sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "tests/test_rhmc.py", "tests/test_md.py", "tests/test_gauge_io.py", "tests/test_reference_regressions.py", "tests/test_gpu_extended.py", "-m", "gpu", "-q", "-x",
           "--runxfail", "-p", "no:cacheprovider", "-n", "4",
           "--deselect", "tests/test_gpu_parity.py::test_16_4_size_independent_properties"]      # 2 min under emulation
    r = subprocess.run(cmd, cwd=ROOT, env=_env(emu_lib, LQCD_TEST_NTRAJ="2", OMP_NUM_THREADS="1", LQCD_TEST_FULL_DIMS="8x8x4x4", LQCD_TEST_CONFIG1_DIMS="8x4x4x4", LQCD_TEST_CONFIG2_DIMS="12x6x4x4"), capture_output=True, text=True, timeout=1500)
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 148, tail


def test_gpu_suite_is_schedule_independent_under_emulation(emu_lib):
    """the single-rank `-m gpu` suite again with the CTAs of every launch executed in DESCENDING blockIdx order and the threads of
    a CTA resumed from the highest index down: a kernel whose CTAs exchange data within one launch (in-place stencil, missing
    double buffer) or that misses a barrier between a shared-memory write and another thread's read fails in one of the orders"""
    cmd = [sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "tests/test_md.py", "tests/test_gauge_io.py", "tests/test_gpu_extended.py", "-m", "gpu", "-q",
           "-x", "--runxfail", "-p", "no:cacheprovider", "-n", "4",
           "-k", "not size_independent and not cgnr_matches_single and not multi_rhs_cg_on and not pipelined_host"]
    r = subprocess.run(cmd, cwd=ROOT, env=_env(emu_lib, LQCD_EMU_CTA_ORDER="reverse", LQCD_EMU_THREAD_ORDER="reverse", OMP_NUM_THREADS="1", LQCD_TEST_FULL_DIMS="8x8x4x4", LQCD_TEST_CONFIG1_DIMS="8x4x4x4", LQCD_TEST_CONFIG2_DIMS="12x6x4x4"),
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 90, r.stdout[-2000:]


@pytest.mark.parametrize("dims,pg,kind", [("4x4x4x8", "1x1x1x2", "Wilson full"), ("4x4x4x4", "1x1x2x2", "staggered full"),
                                          ("4x4x8x4", "1x1x2x2", "Wilson clover")])
def test_multirank_under_emulation(emu_lib, dims, pg, kind):
    """tests/mp_worker.py as separate processes: peer-mapped halo slots, sequence flags, in-kernel all-reduce (POSIX shm IPC)"""
    n = 1
    for v in pg.split("x"):
        n *= int(v)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(31500 + (os.getpid() % 2000)), "tests/mp_worker.py", dims, pg, *kind.split()]
    r = subprocess.run(cmd, cwd=ROOT, env=_env(emu_lib, LQCD_EMU_SHM="1", LQCD_COMM_TIMEOUT_S="120"), capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "FAILED" not in r.stdout


@pytest.mark.parametrize("bulk", ["early", "late"])
@pytest.mark.parametrize("chunks", ["auto", "1", "2"])
@pytest.mark.parametrize("family", ["4", "5", "5-pipelined"])
def test_tmarch_kernel_under_emulation(emu_lib, bulk, chunks, family):
    """t-marching TMA kernel (cp.async.bulk window + link planes on mbarriers, modelled by tests/emu): bulk copies completing at
    issue (earliest) and only when somebody waits on their mbarrier (latest) -- a missing wait or a premature slot refill
    shows up as NaNs / mismatches in one of the two; chunks = tasks per patch (window re-priming, persistent task loop)"""
    # 4: one CTA per SM, 3-slot window; 5: two CTAs per SM, two-row planes, carried t- hop; pipelined: LQCD_TM_PIPE=1 (the next task's
    # copies requested during the last step).  LQCD_EMU_SMS=1: one emulated SM, so that every CTA loops over several tasks.
    extra = {"LQCD_WILSON_KERNEL": family[0], "LQCD_EMU_SMS": "1"}
    if family.endswith("pipelined"):
        extra["LQCD_TM_PIPE"] = "1"
    if chunks != "auto":
        extra["LQCD_TM_CHUNKS"] = chunks
    r = subprocess.run([sys.executable, "tests/tmarch_worker.py"], cwd=ROOT, capture_output=True, text=True, timeout=900,
                       env=_env(emu_lib, LQCD_EMU_BULK=bulk, LQCD_EMU_TRACE="1", **extra))
    assert r.returncode == 0 and "TMARCH OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    m = re.search(r"launches wilson_tmarch%s_kernel\s+(\d+)" % ("2" if family[0] == "5" else ""), r.stderr)
    assert m and int(m.group(1)) > 500 and "launches wilson_dslash_kernel" not in r.stderr, r.stderr[-2000:]


@pytest.mark.parametrize("action", ["wilson", "rhmc"])
def test_multirank_md_trajectory_under_emulation(emu_lib, action):
    """device-resident HMC trajectory on 4 ranks (2 partitioned directions: face and corner links from peer-mapped arrays,
    device-side barriers between the MD sub-steps); rhmc = staggered Nf = 2 rational action (lqcd_md_trajectory_rational)"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
           "--master-port", str(35500 + (os.getpid() % 2000)), "tests/mp_md_worker.py", "4x4x4x4", "1x1x2x2", action]
    r = subprocess.run(cmd, cwd=ROOT, env=_env(emu_lib, LQCD_EMU_SHM="1", LQCD_COMM_TIMEOUT_S="120"), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "FAILED" not in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_config5_shape_rhmc_trajectory_on_8_ranks_under_emulation(emu_lib):
    """BASELINE config 5 in miniature: staggered Nf = 2 RHMC trajectory (multi-shift CG + rational force per step) on 8 ranks in the
    1.1.2.4 process grid bench.py uses at N = 8"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=8", "--master-addr", "127.0.0.1",
           "--master-port", str(39500 + (os.getpid() % 1000)), "tests/mp_md_worker.py", "4x4x8x16", "1x1x2x4", "rhmc"]
    r = subprocess.run(cmd, cwd=ROOT, env=_env(emu_lib, LQCD_EMU_SHM="1", LQCD_COMM_TIMEOUT_S="300", OMP_NUM_THREADS="1"),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "FAILED" not in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_multirank_gauge_io_under_emulation(emu_lib):
    """file <-> device links across 4 ranks: block-wise load of ILDG / BridgeText, multi-rank plaquette, collective ILDG save"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
           "--master-port", str(37500 + (os.getpid() % 2000)), "tests/mp_io_worker.py", "4x4x8x4", "1x1x2x2"]
    r = subprocess.run(cmd, cwd=ROOT, env=_env(emu_lib, LQCD_EMU_SHM="1", LQCD_COMM_TIMEOUT_S="120"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FAILED" not in r.stdout and "ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_bench_experiments_leg_under_emulation(emu_lib, monkeypatch):
    """bench.py's "experiments" leg (isolated child processes timing the not-yet-on-hardware variants at round end): every child
    runs, checks itself against the verified default path and reports ok -- here on a tiny lattice against the emulated build"""
    sys.path.insert(0, str(ROOT))
    import bench
    for k, v in _env(emu_lib, LQCD_EXP_SMALL="4x4x4x4").items():
        monkeypatch.setenv(k, v)
    res = bench.run_experiments("8x4x4x4", 0, 600.0)
    assert set(res) == set(bench.EXPERIMENTS)
    bad = {k: v for k, v in res.items() if not v.get("ok")}
    assert not bad, bad
    assert res["staggered_mrhs"]["bit_identical_to_single_rhs"] and res["staggered_mrhs_r3"]["bit_identical_to_single_rhs"]
    assert res["tmarch_kernel"]["max_rel_dev_vs_default"] < 1e-13 and res["links_full"]["max_rel_dev_vs_default"] < 1e-13
    assert res["tmarch2_kernel"]["bit_identical_to_default"] and res["tmarch2_pipelined"]["bit_identical_to_default"]


def test_bench_multirank_experiments_leg_under_emulation(emu_lib):
    """bench.py's N > 1 experiments leg: every rank spawns one child per knob setting, the children form their own process group,
    connect over (emulated) CUDA IPC and time Dslash / CG; the knob must not change |D x|^2 or the CG residual"""
    import json
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(38500 + (os.getpid() % 1000)), "tests/mp_bench_exp_worker.py", "4x4x4x8"]
    r = subprocess.run(cmd, cwd=ROOT, env=_env(emu_lib, LQCD_EMU_SHM="1", LQCD_COMM_TIMEOUT_S="120", LQCD_EXP_REPS="3", LQCD_EXP_CG="10"),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULTS ")][-1]
    res = json.loads(line[len("RESULTS "):])
    sys.path.insert(0, str(ROOT))
    import bench
    assert set(bench.EXPERIMENTS_MULTI) <= set(res)
    for name in bench.EXPERIMENTS_MULTI:
        assert res[name].get("ok"), (name, res[name])
        # (bit-identical except where the arithmetic differs by construction: full vs two-row links, the t-marching kernel's tiling)
        assert abs(res[name]["resid_sq"] - res["default"]["resid_sq"]) <= 1e-12 * res["default"]["resid_sq"] and res[name]["cg_iters"] == 10
    c4 = res["default"]["config4_wilson_clover_cg"]
    assert c4["ok"] and c4["cg_iters_eps1e-10"] > 5, c4


def test_bench_headline_path_under_emulation(emu_lib):
    """bench.py's run_b200 end to end (tests/bench_emu_harness.py: torch.cuda / cudart events stubbed, emulated library, tiny
    lattice): one JSON line with every key of the contract, the pipelined host call selected through its isolated probe, all
    experiments ok"""
    import json
    r = subprocess.run([sys.executable, "tests/bench_emu_harness.py", "--lattice", "8x4x4x4", "--steps", "3", "--warmup", "1", "--cg-iters", "5"],
                       cwd=ROOT, env=_env(emu_lib, LQCD_EXP_SMALL="4x4x4x4"), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                                    # stdout carries exactly ONE JSON line
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    # --steps 3 is reported as given, the event bracket covers max(steps, 200) applications (insensitive to a small --steps)
    assert d["steps"] == 3 and d["gpu_launches"] == 200 == d["config"]["applications_timed"]
    assert d["roofline"]["bound"] == "hbm" and d["cpu_baseline"]["kind"] == "port"
    assert d["config"]["parity"]["oracle_ok"] and d["config"]["parity"]["max_rel_dev_vs_oracle"] < 1e-13
    assert d["config"]["cg_iters_per_s"] > 0 and d["roofline"]["staggered"]["ms"] > 0 and d["roofline"]["per"] == "GPU"
    assert "lqcd_dslash_host" in d["e2e"]["call"] and "note" not in d["e2e"]
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * 4 * 4 * 4 * 12 * 16
    assert all(v.get("ok") for v in d["experiments"].values()), d["experiments"]


def test_c_example_under_emulation(emu_lib, tmp_path, golden_dir):
    """examples/propagator.c (plain C client of the ABI) linked against the emulated build: plaquette of the reference's fixture,
    per-source CGNR iteration counts and the pion correlator equal the oracle's"""
    sys.path.insert(0, str(ROOT / "tests"))
    import test_c_example as t
    exe = t._compile(tmp_path, emu_lib.parent, emu_lib.name)
    t.check_against_oracle(exe, golden_dir, tmp_path)


def test_bench_watchdog_prints_the_headline_if_the_experiments_leg_overruns(emu_lib):
    import json
    r = subprocess.run([sys.executable, "tests/bench_emu_harness.py", "--lattice", "8x4x4x4", "--steps", "3", "--warmup", "1", "--cg-iters", "5"],
                       cwd=ROOT, env=_env(emu_lib, LQCD_EXP_SMALL="4x4x4x4", LQCD_BENCH_WATCHDOG_S="2"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and "watchdog" in d["experiments"]["error"]

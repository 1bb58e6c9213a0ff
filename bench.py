#!/usr/bin/env python
"""
bench.py -- headline benchmark of the B200-native Dirac-solve path (contract: see the task prompt).

Metric (BASELINE.json): Wilson Dslash GFLOP/s (+ CG iterations/s) at 32^4, fp64, 1/2/4/8 B200 vs CPU ref.
A "step" is ONE application y = D x of the Wilson operator (LinearAlgebra.mul!(y, D, x)) on the full
32^4 lattice (configs[3] volume; 16^4 of configs[1] is L2-resident and is a parity-test size, not a bench line),
synthetic hot SU(3) links generated on the device (seed 111) and a Gaussian source (seed 112).

  value      whole-job Dslash GFLOP/s = 1368 flop/site x V / t, t = (one CUDA-event bracket around the K timed
             back-to-back applications on the library stream) / K, max over ranks; inputs resident in HBM and
             larger than L2 at N=1 (806 MB vs 126 MB); the L2-flushed per-application time is reported next to it
             (config.ms_flushed).
  roofline   HBM: achieved = 960 B/site x V / t  against MEASURED_PEAKS.json:hbm_gbs; traffic = DRAM bytes per
             launch from the committed ncu capture (profiles/wilson_dslash_traffic.json).
  e2e        the same metric through the reference-facing call with HOST buffers: per step the source is
             copied host->device (pinned), mul_(y, D, x) runs, and the result is copied back.
  cg         CG iterations/s of solve_DinvX_(y, DdagD, b) (device resident and through host buffers).
  cpu_baseline / --impl reference: the CPU oracle (oracle/lqcd_oracle.c, a restatement of the reference's
             Julia CPU path; the reference itself is Julia and cannot run here) on all host threads.

  experiments  NOT part of the headline: after everything above is measured, code paths that were written while no B200 was
             reachable (t-marching / persistent / multi-RHS / shared-memory-link kernels, clover, even-odd and half-field
             solves, device-resident MD and RHMC trajectories; at N > 1 the multi-GPU knobs) are timed and self-checked in
             isolated child processes with a time budget (LQCD_BENCH_EXPERIMENTS=0 switches the leg off).

N > 1 (torchrun): strong scaling -- the 32^4 lattice is split along T (then Z) over the ranks.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path[:0] = [str(ROOT), str(ROOT / "latticeqcd.jl_b200")]

FLOP_PER_SITE = 1368          # 1320 hopping + 48 xpay
BYTES_PER_SITE = 960          # 576 links + 192 in + 192 out
CG_BYTES_PER_SITE = 4224      # un-fused algorithmic figure (2 mul! + 12 vector passes), SURVEY.md 8d
STAG_BYTES_PER_SITE = 672     # 576 links + 48 in + 48 out
STAG_FLOP_PER_SITE = 582
# N-independent fingerprints of the bench workload (hot links seed 111, Gaussian sources seed 112, kappa 0.12 / mass 0.5): measured
# at N = 1 where the same run compares y = D x with the oracle; every N must reproduce them to 1e-12 (deterministic reductions).
EXPECTED = {"32x32x32x32": {"norm_Dx_sq": 15490456.52178676, "staggered_norm_Dx_sq": 7083744.73603815, "cg_converged_iters_eps1e-10": 140}}
KAPPA = 0.12
BC = [1, 1, 1, -1]


def metric_name():
    """BASELINE.json's metric string (the Dslash GFLOP/s part is `value`, the CG part is reported under "cg")."""
    try:
        return json.loads((ROOT / "BASELINE.json").read_text())["metric"]
    except Exception:
        return "Wilson Dslash GFLOP/s & CG iters/s at 32^4, 1/2/4/8 B200 vs CPU ref"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lattice", default="32x32x32x32")
    ap.add_argument("--cg-iters", type=int, default=100)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--probe-pipe", action="store_true", help=argparse.SUPPRESS)   # child process of probe_pipe_isolated()
    ap.add_argument("--experiment", default=None, help=argparse.SUPPRESS)          # child process of run_experiments()
    ap.add_argument("--experiment-multi", default=None, help=argparse.SUPPRESS)    # child process of run_experiments_multi()
    return ap.parse_args()


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text())["hbm_gbs"], "measured (MEASURED_PEAKS.json:hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
# CPU arm (oracle).  Used for cpu_baseline (rank 0, N=1) and for --impl reference.
# ---------------------------------------------------------------------------------------------------
def cpu_dslash(dims, steps, warmup, threads, U=None, x=None):
    """times the oracle's Wilson D on the host cores.  U / x: host arrays to use (the links and source the GPU holds -- the
    result is then returned for the parity assertion); default: synthetic fields of the same shape."""
    import numpy as np
    from oracle import oracle as orc
    orc.build()
    threads = orc.set_threads(threads)
    op = orc.make_op(dims, kappa=KAPPA, bc=tuple(BC))
    NX, NY, NZ, NT = dims
    V = NX * NY * NZ * NT
    if U is None:
        # synthetic links: a 4^4 Haar block tiled over the lattice (content does not affect CPU timing)
        small = orc.random_su3((4, 4, 4, 4), seed=111)
        reps = (1, NT // 4, NZ // 4, NY // 4, NX // 4, 1, 1)
        U = np.ascontiguousarray(np.tile(small, reps))
    if x is None:
        rng = np.random.default_rng(112)
        x = np.ascontiguousarray(rng.standard_normal((4, NT, NZ, NY, NX, 3)) + 1j * rng.standard_normal((4, NT, NZ, NY, NX, 3)))
    y = None
    for _ in range(warmup):
        y = orc.apply(op, orc.WILSON, orc.D, U, x)
    t0 = time.perf_counter()
    for _ in range(steps):
        y = orc.apply(op, orc.WILSON, orc.D, U, x)
    dt = (time.perf_counter() - t0) / steps
    return {"gflops": FLOP_PER_SITE * V / dt / 1e9, "ms": dt * 1e3, "threads": threads, "y": y}


def workload_string(lattice):
    """config.workload -- identical in both arms (the driver compares the strings)"""
    return f"Wilson Dslash mul!(y,D,x) {lattice} SU(3) hot links, kappa={KAPPA}, r=1, bc={BC}"


def probe_live_reference():
    """Is the reference itself runnable on this box?  (it is Julia + un-vendored LatticeDiracOperators.jl /
    Gaugefields.jl.)  Recorded in the reference arm's line; the CPU arm falls back to the oracle port when it is not."""
    import shutil
    julia = shutil.which("julia")
    ref_dir = ROOT / "baseline" / "_ref"
    has_ref = ref_dir.is_dir() and any(ref_dir.iterdir())
    pkgs = False
    if julia:
        try:
            r = subprocess.run([julia, "-e", "using LatticeDiracOperators, Gaugefields; print(1)"], capture_output=True, text=True, timeout=120)
            pkgs = r.returncode == 0
        except Exception:
            pkgs = False
    return {"julia": julia, "baseline/_ref": bool(has_ref), "LatticeDiracOperators.jl importable": pkgs,
            "usable": bool(julia and pkgs)}


def run_reference(args, dims):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    thr = host_threads()
    probe = probe_live_reference()
    # one step = one full-lattice application (~50 ms on 16 cores): --steps / --warmup are honoured as given
    steps, warm = max(1, args.steps), max(0, args.warmup)
    r = cpu_dslash(dims, steps, warm, thr)
    line = {
        "impl": "reference", "metric": metric_name(), "value": r["gflops"], "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": r["ms"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(args.lattice),
                   "arm": "CPU oracle port of the reference's Julia path (oracle/lqcd_oracle.c, OpenMP); live reference probe: " + json.dumps(probe)},
        "cpu_baseline": {"value": r["gflops"], "unit": "GFLOP/s", "cores": r["threads"], "kind": "port",
                         "sample": f"{steps} full-lattice applications at {args.lattice}"},
        "e2e": {"value": r["gflops"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# small cudart wrapper: events on the library's own stream
# ---------------------------------------------------------------------------------------------------
class Cudart:
    def __init__(self):
        self.rt = None
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                self.rt = C.CDLL(name)
                break
            except OSError:
                continue
        if self.rt is None:
            raise RuntimeError("libcudart not found")
        self.rt.cudaEventCreate.argtypes = [C.POINTER(C.c_void_p)]
        self.rt.cudaEventRecord.argtypes = [C.c_void_p, C.c_void_p]
        self.rt.cudaEventSynchronize.argtypes = [C.c_void_p]
        self.rt.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]

    def event(self):
        e = C.c_void_p()
        assert self.rt.cudaEventCreate(C.byref(e)) == 0
        return e

    def record(self, e, stream):
        assert self.rt.cudaEventRecord(e, C.c_void_p(stream)) == 0

    def elapsed_ms(self, e0, e1):
        assert self.rt.cudaEventSynchronize(e1) == 0
        ms = C.c_float()
        assert self.rt.cudaEventElapsedTime(C.byref(ms), e0, e1) == 0
        return ms.value


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={device}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm = sorted(int(r[1]) for r in rows if len(r) >= 9 and r[1].isdigit())
        reasons = set()
        for r in rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        mx = [int(r[2]) for r in rows if len(r) >= 9 and r[2].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def probe_pipe_child(dims):
    """Child process: lqcd_dslash_host against the three-call sequence on a fresh context.  Prints PIPE_OK on bit-for-bit
    agreement.  Runs isolated so that a device fault in the pipelined path (not yet run on hardware when it was written)
    cannot leave a sticky CUDA error in the measuring process."""
    import numpy as np
    import lqcd_b200 as q
    from lqcd_b200 import _lib as L
    ctx = q.get_context(dims, procgrid=(1, 1, 1, 1), rank=0, device=int(os.environ.get("LOCAL_RANK", "0")))
    ctx.call("lqcd_gauge_random", 111, -1.0)
    op = L.LqcdOp()
    op.kind, op.kappa, op.r = L.WILSON, KAPPA, 1.0
    for i, b in enumerate(BC):
        op.bc[i] = b
    x, y = q.FermionField(ctx, L.WILSON), q.FermionField(ctx, L.WILSON)
    q.gauss_distribution_fermion_(x, 112)
    hx = np.ascontiguousarray(x.to_host())
    hy = np.zeros_like(hx)
    for rep in range(2):                      # second pass reuses the staging buffers / events
        ctx.call("lqcd_fermion_upload", x.h, hx.ctypes.data, 0)
        ctx.call("lqcd_dslash", C.byref(op), y.h, x.h, L.OP_D)
        ctx.call("lqcd_fermion_download", y.h, hy.ctypes.data, 0)
        ref = hy.copy()
        hy[...] = 0
        ctx.call("lqcd_dslash_host", C.byref(op), y.h, x.h, hy.ctypes.data, hx.ctypes.data, L.OP_D, 0)
        ctx.synchronize()
        if not np.array_equal(ref, hy):
            print("PIPE_MISMATCH", flush=True)
            return
    print("PIPE_OK", flush=True)


def probe_pipe_isolated(lattice, local_rank):
    """(ok, note): run probe_pipe_child in a subprocess with a timeout"""
    env = dict(os.environ, LOCAL_RANK=str(local_rank))
    for k in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--probe-pipe", "--lattice", lattice], env=env,
                           capture_output=True, text=True, timeout=300)
    except subprocess.TimeoutExpired:
        return False, "pipelined call timed out in the isolated probe: not used"
    if r.returncode == 0 and "PIPE_OK" in r.stdout:
        return True, None
    tail = (r.stdout + r.stderr).strip().splitlines()[-1:] or [""]
    return False, f"pipelined call failed the isolated probe (exit {r.returncode}: {tail[0][:160]}): not used"


# ---------------------------------------------------------------------------------------------------
# "experiments": first hardware timings of code paths that were written while no B200 was reachable (pre-flighted under
# tests/emu only).  NOT part of value / roofline / e2e: each runs in its own child process after the headline measurement is
# complete (a device fault or hang there cannot touch the measuring process), is checked for correctness against the verified
# default path, and its numbers are reported under the "experiments" key for the next round's tuning.
# ---------------------------------------------------------------------------------------------------
EXPERIMENTS = {
    # name: (environment of the child, what it measures)
    "default": ({}, "default Wilson kernel: register-resident, one thread per site, two-row links (reference for the rows below; writes the 16^4 comparison vector)"),
    "tmarch_kernel": ({"LQCD_WILSON_KERNEL": "4"}, "t-marching kernel with TMA-staged spinor window and link planes (wilson_tmarch.cu, experimental)"),
    "tmarch2_kernel": ({"LQCD_WILSON_KERNEL": "5"}, "second-generation t-marching TMA kernel: two CTAs per SM, two-row link planes, carried t-backward hop (wilson_tmarch.cu, experimental)"),
    # verified under tests/emu only: its child gets 40 s, so a hang costs that much of the leg's budget and nothing else
    "tmarch2_pipelined": ({"LQCD_WILSON_KERNEL": "5", "LQCD_TM_PIPE": "1"}, "second-generation t-marching kernel with pipelined tasks: the next task's copies are requested during the last step of the current one, x phase without a CTA barrier"),
    "staggered_mrhs": ({}, "staggered: single-RHS kernel vs 12 right-hand sides, default grouping (4 per thread)"),
    "staggered_mrhs_r2": ({"LQCD_MRHS_R_STAGGERED": "2"}, "staggered, 2 right-hand sides per thread"),
    "staggered_mrhs_r3": ({"LQCD_MRHS_R_STAGGERED": "3"}, "staggered, 3 right-hand sides per thread"),
    "staggered_mrhs_r6": ({"LQCD_MRHS_R_STAGGERED": "6"}, "staggered, 6 right-hand sides per thread"),
    "links_full": ({"LQCD_LINKS12": "0"}, "Dslash kernels reading the full 3x3 links instead of the two-row copy (links12.cu)"),
    "propagator": ({}, "12 point-source CGNR solves (measure_Pion_correlator.jl:333-409) through lqcd_solve_multi (Wilson: single-RHS kernels, one source after the other) vs 12 x lqcd_solve"),
    "clover": ({}, "Wilson-clover Dslash (csw = 1.5612)"),
    "evenodd": ({}, "even-odd preconditioned CGNR vs full CGNR, 16^4"),
    "staggered_even": ({}, "staggered CG on an even-site source: half-field solver vs full-lattice solver"),
    "md": ({}, "device-resident Sexton-Weingarten trajectory with Wilson pseudofermions, 16^4"),
    "force": ({}, "fermion-force outer-product kernels (Wilson, staggered) alone, 32^4"),
    "rhmc_md": ({}, "device-resident RHMC trajectory (staggered Nf = 2: multi-shift CG + rational force per step), 16^4"),
}


def _timed(fn, reps):
    """wall clock around calls that return with the library stream drained (every exported call is blocking-on-return)"""
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) * 1e3 / reps


def experiment_child(name, dims):
    """child process: prints ONE JSON dict on stdout (whatever was measured before an error is kept)"""
    out = {"name": name, "ok": False}
    try:
        _experiment_body(name, dims, out)
    except BaseException as exc:                     # a failed check / library error: report it, keep the partial numbers
        out["ok"] = False
        out["error"] = repr(exc)[:300]
    print("EXPERIMENT " + json.dumps(out), flush=True)


def _experiment_body(name, dims, out):
    import numpy as np
    import lqcd_b200 as q
    from lqcd_b200 import _lib as L
    ref_file = Path(tempfile.gettempdir()) / f"lqcd_b200_exp_ref_{os.environ.get('LQCD_EXP_TAG', '0')}.npy"
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    small = tuple(int(v) for v in os.environ.get("LQCD_EXP_SMALL", "16x16x16x16").split("x"))     # (tests shrink it under emulation)
    V = int(np.prod(dims))

    def setup(d, kind=L.WILSON, csw=0.0, eps=-1.0):
        ctx = q.get_context(d, procgrid=(1, 1, 1, 1), rank=0, device=dev)
        ctx.call("lqcd_gauge_random", 111, eps)
        op = L.LqcdOp()
        op.kind, op.kappa, op.r, op.mass, op.csw = kind, KAPPA, 1.0, 0.5, csw
        for i, b in enumerate(BC):
            op.bc[i] = b
        x, y = q.FermionField(ctx, kind), q.FermionField(ctx, kind)
        q.gauss_distribution_fermion_(x, 112)
        return ctx, op, x, y

    def dslash_ms(ctx, op, y, x, reps=20):
        mean, mn = C.c_double(), C.c_double()
        ctx.call("lqcd_time_dslash", C.byref(op), y.h, x.h, L.OP_D, 3, 0, C.byref(mean), C.byref(mn))
        ctx.call("lqcd_time_dslash", C.byref(op), y.h, x.h, L.OP_D, reps, 0, C.byref(mean), C.byref(mn))
        return mean.value

    if name in ("default", "register_kernel", "links_full", "tmarch_kernel", "tmarch2_kernel", "tmarch2_pipelined"):
        ctx, op, x, y = setup(small)
        ctx.call("lqcd_dslash", C.byref(op), y.h, x.h, L.OP_D)
        got = y.to_host()
        if name == "default":
            np.save(ref_file, got)
            out["max_rel_dev_vs_default"] = 0.0
        else:
            ref = np.load(ref_file)
            out["max_rel_dev_vs_default"] = float(np.abs(got - ref).max() / np.abs(ref).max())
            out["bit_identical_to_default"] = bool(np.array_equal(got, ref))
        out["ok"] = out["max_rel_dev_vs_default"] < 1e-13
        out["ms_16^4"] = dslash_ms(ctx, op, y, x, 50)
        ctx2, op2, x2, y2 = setup(dims)
        ms = dslash_ms(ctx2, op2, y2, x2)
        out.update({"ms_per_apply": ms, "GB/s": BYTES_PER_SITE * V / ms / 1e6, "frac_of_peak": BYTES_PER_SITE * V / ms / 1e6 / peaks()[0]})
        # CG iterations/s with this kernel variant (warm field, fixed 60 iterations)
        ctx2.call("lqcd_gauge_random", 111, 0.3)
        it, rs = C.c_int(0), C.c_double(0.0)
        for n_it in (10, 60):
            q.clear_fermion_(y2)
            t0 = time.perf_counter()
            st = ctx2.lib.lqcd_solve(ctx2.h, C.byref(op2), y2.h, x2.h, L.SOLVER_CG, L.OP_DDAGD, 0.0, n_it, C.byref(it), C.byref(rs), None)
            ctx2.synchronize()
            dt = time.perf_counter() - t0
        out.update({"cg_iters_per_s": it.value / dt, "cg_resid_sq_after_60": rs.value})
    elif name == "force":
        res = {}
        for kind, key, csw in ((L.WILSON, "wilson", 0.0), (L.STAGGERED, "staggered", 0.0), (L.WILSON, "wilson_clover (hopping + clover-term kernels)", 1.5612)):
            ctx, op, x, y = setup(dims, kind=kind, eps=0.3, csw=csw)
            ctx.call("lqcd_dslash", C.byref(op), y.h, x.h, L.OP_D)                    # some (X, Y) pair: the kernel's cost does not depend on it
            t = _timed(lambda: ctx.call("lqcd_fermion_force_xy", C.byref(op), x.h, y.h, 1.0, 0), 10)
            # compulsory bytes per site: links read + force written (576 each) + X and Y read once (their forward neighbours are cache hits)
            res[key] = {"ms_outer_product_kernel": t, "GB/s (compulsory bytes: links + force + X + Y)": (576 + 576 + 2 * (192 if kind == L.WILSON else 48)) * V / t / 1e6}
        out["ok"] = True
        out.update(res)
    elif name.startswith("mrhs_") or name.startswith("staggered_mrhs"):
        kind = L.STAGGERED if name.startswith("staggered_mrhs") else L.WILSON
        nrhs = 12
        ctx, op, x, y = setup(dims, kind)
        xs = [x] + [q.FermionField(ctx, kind) for _ in range(nrhs - 1)]
        for j, f in enumerate(xs[1:]):
            q.gauss_distribution_fermion_(f, 200 + j)
        ys = [q.FermionField(ctx, kind) for _ in range(nrhs)]
        hx = (C.c_void_p * nrhs)(*[f.h.value for f in xs])
        hy = (C.c_void_p * nrhs)(*[f.h.value for f in ys])
        ctx.call("lqcd_dslash_multi", C.byref(op), hy, hx, nrhs, L.OP_D)
        same = True
        for j in (0, 5, 11):                       # bit-for-bit against the verified single-RHS kernel
            ctx.call("lqcd_dslash", C.byref(op), y.h, xs[j].h, L.OP_D)
            same = same and bool(np.array_equal(y.to_host(), ys[j].to_host()))
        out["ok"] = out["bit_identical_to_single_rhs"] = same
        ms = _timed(lambda: ctx.call("lqcd_dslash_multi", C.byref(op), hy, hx, nrhs, L.OP_D), 10)
        single = dslash_ms(ctx, op, y, x)
        per_site = 672 if kind == L.STAGGERED else BYTES_PER_SITE
        out.update({"nrhs": nrhs, "ms_per_pass": ms, "ms_per_rhs": ms / nrhs, "ms_single_rhs_kernel": single, "speedup_per_rhs": single * nrhs / ms,
                    "equivalent_single_rhs_GB/s": per_site * V * nrhs / ms / 1e6})
    elif name == "propagator":
        ctx, op, x, y = setup(dims, eps=0.3)
        nsrc = 12
        bs, xs = [q.FermionField(ctx, L.WILSON) for _ in range(nsrc)], [q.FermionField(ctx, L.WILSON) for _ in range(nsrc)]
        for i, b in enumerate(bs):                     # spin-colour point sources at the origin
            q.clear_fermion_(b)
            q.setindex_global_(b, 1.0, i // 4 + 1, 1, 1, 1, 1, i % 4 + 1)
        hb = (C.c_void_p * nsrc)(*[f.h.value for f in bs])
        hx = (C.c_void_p * nsrc)(*[f.h.value for f in xs])
        its, rss = (C.c_int * nsrc)(), (C.c_double * nsrc)()
        for f in xs:
            q.clear_fermion_(f)
        t0 = time.perf_counter()
        ctx.call("lqcd_solve_multi", C.byref(op), hx, hb, nsrc, L.SOLVER_CGNR, L.OP_D, 1e-16, 3000, its, rss)
        t_multi = time.perf_counter() - t0
        it1, rs1 = C.c_int(0), C.c_double(0.0)
        seq_iters = []
        t0 = time.perf_counter()
        for i in range(nsrc):
            q.clear_fermion_(y)
            ctx.call("lqcd_solve", C.byref(op), y.h, bs[i].h, L.SOLVER_CGNR, L.OP_D, 1e-16, 3000, C.byref(it1), C.byref(rs1), None)
            seq_iters.append(it1.value)
        t_seq = time.perf_counter() - t0
        same = bool(np.array_equal(y.to_host(), xs[nsrc - 1].to_host()))
        out["ok"] = list(its) == seq_iters and same
        out.update({"iters": list(its), "same_iteration_counts": list(its) == seq_iters, "last_solution_bit_identical": same,
                    "wall_ms_lock_step": t_multi * 1e3, "wall_ms_sequential": t_seq * 1e3, "speedup": t_seq / t_multi})
    elif name == "clover":
        ctx, op, x, y = setup(dims, csw=1.5612)
        ms = dslash_ms(ctx, op, y, x)
        # gamma5-hermiticity <a, D b> = <D^dag a, b> as the size-independent check (the clover term is Hermitian)
        a = q.FermionField(ctx, L.WILSON)
        q.gauss_distribution_fermion_(a, 113)
        z = q.FermionField(ctx, L.WILSON)
        ctx.call("lqcd_dslash", C.byref(op), y.h, x.h, L.OP_D)
        ctx.call("lqcd_dslash", C.byref(op), z.h, a.h, L.OP_DDAG)
        l, r = q.dot(a, y), q.dot(z, x)
        out["adjoint_identity_rel_dev"] = abs(l - r) / abs(l)
        out["ok"] = out["adjoint_identity_rel_dev"] < 1e-12
        out.update({"ms_per_apply": ms, "GB/s": (BYTES_PER_SITE + 576) * V / ms / 1e6})
    elif name == "evenodd":
        ctx, op, x, y = setup(small, eps=0.3)
        it, rs = C.c_int(0), C.c_double(0.0)
        res = {}
        for fn in ("lqcd_solve", "lqcd_solve_eo"):
            q.clear_fermion_(y)
            t0 = time.perf_counter()
            ctx.call(fn, C.byref(op), y.h, x.h, L.SOLVER_CGNR, L.OP_D, 1e-16, 3000, C.byref(it), C.byref(rs), None)
            res[fn] = {"iters": it.value, "resid_sq": rs.value, "wall_ms": (time.perf_counter() - t0) * 1e3, "sol": y.to_host()}
        dev_rel = float(np.abs(res["lqcd_solve"]["sol"] - res["lqcd_solve_eo"]["sol"]).max() / np.abs(res["lqcd_solve"]["sol"]).max())
        out["ok"] = dev_rel < 1e-6
        out.update({"solution_rel_dev": dev_rel, "full": {k: v for k, v in res["lqcd_solve"].items() if k != "sol"},
                    "evenodd": {k: v for k, v in res["lqcd_solve_eo"].items() if k != "sol"}})
    elif name == "staggered_even":
        ctx, op, x, y = setup(dims, kind=L.STAGGERED, eps=0.3)
        ctx.call("lqcd_fermion_mask_parity", x.h, 0)
        it, rs = C.c_int(0), C.c_double(0.0)
        res = {}
        for key in ("full", "half"):
            q.clear_fermion_(y)
            t0 = time.perf_counter()
            if key == "full":
                ctx.call("lqcd_solve", C.byref(op), y.h, x.h, L.SOLVER_CG, L.OP_DDAGD, 1e-12, 3000, C.byref(it), C.byref(rs), None)
            else:
                ctx.call("lqcd_solve_staggered_even", C.byref(op), y.h, x.h, 1e-12, 3000, C.byref(it), C.byref(rs))
            res[key] = {"iters": it.value, "resid_sq": rs.value, "wall_ms": (time.perf_counter() - t0) * 1e3, "sol": y.to_host()}
        dev_rel = float(np.abs(res["full"]["sol"] - res["half"]["sol"]).max() / np.abs(res["full"]["sol"]).max())
        out["ok"] = dev_rel < 1e-8 and abs(res["full"]["iters"] - res["half"]["iters"]) <= 1
        out.update({"solution_rel_dev": dev_rel, "full": {k: v for k, v in res["full"].items() if k != "sol"},
                    "half": {k: v for k, v in res["half"].items() if k != "sol"}})
    elif name == "md":
        ctx, op, x, y = setup(small, eps=0.3)
        its = C.c_longlong(0)

        def quenched(dtau, steps):          # same start every time: links and momenta come from counter-based generators
            ctx.call("lqcd_gauge_random", 111, 0.3)
            ctx.call("lqcd_md_momenta_gaussian", 7)
            K0, S0, K1, S1 = C.c_double(), C.c_double(), C.c_double(), C.c_double()
            ctx.call("lqcd_md_kinetic", C.byref(K0)); ctx.call("lqcd_md_gauge_action", 5.7, C.byref(S0))
            t0 = time.perf_counter()
            ctx.call("lqcd_md_trajectory", None, None, 5.7, dtau, steps, 0, 0.0, 1, C.byref(its))
            wall = time.perf_counter() - t0
            ctx.call("lqcd_md_kinetic", C.byref(K1)); ctx.call("lqcd_md_gauge_action", 5.7, C.byref(S1))
            return (K1.value + S1.value) - (K0.value + S0.value), wall

        dH1, wall = quenched(0.02, 10)
        dH2, _ = quenched(0.01, 20)
        out["ok"] = bool(2.5 < dH1 / dH2 < 6.0)                  # leapfrog: Delta H = O(dtau^2)
        out.update({"quenched_dH_dtau0.02": dH1, "quenched_dH_dtau0.01": dH2, "wall_ms_10_steps": wall * 1e3})
        ctx.call("lqcd_gauge_random", 111, 0.3)
        ctx.call("lqcd_md_momenta_gaussian", 7)
        t0 = time.perf_counter()
        ctx.call("lqcd_md_trajectory", C.byref(op), x.h, 5.7, 0.02, 2, 4, 1e-16, 3000, C.byref(its))
        out.update({"dynamical_wall_ms_2_steps_nsw4": (time.perf_counter() - t0) * 1e3, "cg_iters": its.value})
    elif name == "rhmc_md":
        from lqcd_b200 import rhmc
        ra = rhmc.rational_approx(-2 / 8.0, 12, 0.22, 17.0)             # x^(-Nf/8), Nf = 2, mass 0.5: spec(DdagD) in [0.25, 16.25]
        al = np.ascontiguousarray(ra.alpha, dtype=np.float64)
        sh = np.ascontiguousarray(ra.beta, dtype=np.float64)
        pa, ps = al.ctypes.data_as(L.pdbl), sh.ctypes.data_as(L.pdbl)
        ctx, op, x, y = setup(small, kind=L.STAGGERED, eps=0.3)
        its, it1, S = C.c_longlong(0), C.c_int(0), C.c_double(0.0)

        def run(dtau, steps):
            ctx.call("lqcd_gauge_random", 111, 0.3)
            ctx.call("lqcd_md_momenta_gaussian", 7)
            K, G = C.c_double(), C.c_double()
            H = []
            for leg in range(2):
                ctx.call("lqcd_md_kinetic", C.byref(K)); ctx.call("lqcd_md_gauge_action", 5.7, C.byref(G))
                ctx.call("lqcd_rational_apply", C.byref(op), y.h, x.h, float(ra.alpha0), pa, ps, len(al), 1e-18, 3000, C.byref(it1), C.byref(S))
                H.append(K.value + G.value + S.value)
                if leg == 0:
                    t0 = time.perf_counter()
                    ctx.call("lqcd_md_trajectory_rational", C.byref(op), x.h, pa, ps, len(al), 5.7, dtau, steps, 0, 1e-18, 3000, C.byref(its))
                    wall = time.perf_counter() - t0
            return H[1] - H[0], wall, its.value

        dH1, wall, nit = run(0.02, 4)
        dH2, _, _ = run(0.01, 8)
        out["ok"] = bool(2.5 < dH1 / dH2 < 6.0)
        out.update({"dH_dtau0.02": dH1, "dH_dtau0.01": dH2, "wall_ms_4_steps": wall * 1e3, "multishift_cg_iters": nit, "poles": len(al)})
    else:
        out["error"] = "unknown experiment"


def run_experiments(lattice, local_rank, budget_s):
    """parent side: every experiment in its own child process, bounded in time; returns {name: result}"""
    results = {}
    t_start = time.perf_counter()
    tag = str(os.getpid())
    for name, (env_extra, what) in EXPERIMENTS.items():
        left = budget_s - (time.perf_counter() - t_start)
        if left < 20:
            results[name] = {"skipped": "time budget of the experiments leg used up"}
            continue
        env = dict(os.environ, LOCAL_RANK=str(local_rank), LQCD_EXP_TAG=tag, **env_extra)
        for k in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        try:
            r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--experiment", name, "--lattice", lattice], env=env,
                               capture_output=True, text=True, timeout=min(left, 40 if name == "tmarch2_pipelined" else 120))
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("EXPERIMENT ")]
            if line:
                results[name] = json.loads(line[-1][len("EXPERIMENT "):])
            else:
                tail = (r.stdout + r.stderr).strip().splitlines()[-1:] or [""]
                results[name] = {"ok": False, "error": f"exit {r.returncode}: {tail[0][:200]}"}
        except subprocess.TimeoutExpired:
            results[name] = {"ok": False, "error": "timed out"}
        except Exception as exc:
            results[name] = {"ok": False, "error": repr(exc)}
        results[name]["what"] = what
    try:
        (Path(tempfile.gettempdir()) / f"lqcd_b200_exp_ref_{tag}.npy").unlink()
    except OSError:
        pass
    return results


# N > 1: the multi-GPU knobs that were prepared for the strong-scaling gap but never timed.  Every rank
# of the job spawns ONE child with its own RANK / LOCAL_RANK; the children of one experiment form their own process group (gloo,
# host plumbing only) on a different port, create contexts, connect over CUDA IPC and time the Dslash and a CG with the knob set
# in their environment.  The parents only wait.  Same isolation argument as above.
EXPERIMENTS_MULTI = {       # most informative first: the leg stops starting new ones when its time budget is used up
    "default": ({}, "defaults: Wilson Dslash + CG (reference for the knob rows), BASELINE configs[3] (32^4 Wilson-clover CG) and, on 8 ranks, "
                    "configs[4] (32^3 x 64 staggered Nf = 2 RHMC trajectory, 12 poles)"),
    "timeline": ({"LQCD_COMM_TIMING": "1"}, "in-kernel phase stamps of the Wilson Dslash (pack / interior / face tiles / flag waits)"),
    "self_pack_spt4": ({"LQCD_SELF_PACK": "1", "LQCD_PACK_SPT": "4"}, "pack CTAs lead the Dslash kernel, 4 face sites per pack thread (4x fewer pack CTAs)"),
    "self_pack": ({"LQCD_SELF_PACK": "1"}, "pack CTAs lead the Dslash kernel"),
    "tmarch_kernel": ({"LQCD_WILSON_KERNEL": "4"}, "t-marching TMA Wilson kernel (cyclic march: the two halo slices come last)"),
    "tmarch2_kernel": ({"LQCD_WILSON_KERNEL": "5"}, "second-generation t-marching TMA Wilson kernel (two CTAs per SM)"),
    "separate_pack": ({"LQCD_SELF_PACK": "0"}, "pack kernel on the priority stream"),
    "links_full": ({"LQCD_LINKS12": "0"}, "full 3x3 links instead of the two-row copy"),
    "pack_fence_sys": ({"LQCD_PACK_FENCE": "sys"}, "system-scope fence per pack CTA (round-1 default)"),
    "tmarch2_pipelined": ({"LQCD_WILSON_KERNEL": "5", "LQCD_TM_PIPE": "1"}, "second-generation t-marching kernel with pipelined tasks (tests/emu only so far)"),
}


def experiment_multi_child(name, dims):
    import numpy as np
    import torch.distributed as dist
    import lqcd_b200 as q
    from lqcd_b200 import _lib as L
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = int(os.environ.get("LOCAL_RANK", rank))
    pg = choose_procgrid(world)
    if os.environ.get("LQCD_PROCGRID"):
        pg = tuple(int(v) for v in os.environ["LQCD_PROCGRID"].split(","))
    ctx = q.get_context(dims, procgrid=pg, rank=rank, device=dev)
    q.connect_ranks(ctx, dist)
    ctx.call("lqcd_gauge_random", 111, 0.3)
    op = L.LqcdOp()
    op.kind, op.kappa, op.r = L.WILSON, KAPPA, 1.0
    for i, b in enumerate(BC):
        op.bc[i] = b
    x, y, sol = q.FermionField(ctx, L.WILSON), q.FermionField(ctx, L.WILSON), q.FermionField(ctx, L.WILSON)
    q.gauss_distribution_fermion_(x, 112)
    mean, mn = C.c_double(), C.c_double()
    ctx.call("lqcd_time_dslash", C.byref(op), y.h, x.h, L.OP_D, 5, 0, C.byref(mean), C.byref(mn))
    dist.barrier()
    reps, cg_it = int(os.environ.get("LQCD_EXP_REPS", "100")), int(os.environ.get("LQCD_EXP_CG", "200"))      # (tests shrink them)
    ctx.call("lqcd_time_dslash", C.byref(op), y.h, x.h, L.OP_D, reps, 0, C.byref(mean), C.byref(mn))
    ms = mean.value
    norm_y = q.dot(y, y).real                          # global |D x|^2: must not depend on the knob
    it, rs = C.c_int(0), C.c_double(0.0)

    def cg(maxit):
        q.clear_fermion_(sol)
        st = ctx.lib.lqcd_solve(ctx.h, C.byref(op), sol.h, x.h, L.SOLVER_CG, L.OP_DDAGD, 0.0, maxit, C.byref(it), C.byref(rs), None)
        assert st in (L.LQCD_OK, L.ERR_NOCONV), ctx.lib.lqcd_last_error(ctx.h)
    cg(10)
    dist.barrier()
    t0 = time.perf_counter()
    cg(cg_it)
    ctx.synchronize()
    cg_s = time.perf_counter() - t0
    import torch
    t = torch.tensor([ms, cg_s], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    V = int(np.prod(dims))
    out = {"name": name, "ok": True, "n_gpus": world, "procgrid": list(pg), "ms_per_apply": float(t[0]), "GFLOP/s": FLOP_PER_SITE * V / float(t[0]) / 1e6,
           "cg_iters_per_s": it.value / float(t[1]), "norm_Dx_sq": norm_y, "cg_iters": it.value, "resid_sq": rs.value}
    if name == "default":
        try:
            out["config4_wilson_clover_cg"] = _config4_clover_cg(ctx, dist, q, L, x, sol, V)
        except Exception as exc:
            out["config4_wilson_clover_cg"] = {"ok": False, "error": repr(exc)[:300]}
        if world == 8 and os.environ.get("LQCD_EXP_CONFIG5", "1") != "0":
            try:
                out["config5_staggered_rhmc_trajectory"] = _config5_rhmc_trajectory(dist, q, L, pg, rank, dev)
            except Exception as exc:
                out["config5_staggered_rhmc_trajectory"] = {"ok": False, "error": repr(exc)[:300]}
    if rank == 0:
        print("EXPERIMENT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def _config4_clover_cg(ctx, dist, q, L, x, sol, V):
    """BASELINE configs[3]: Wilson-clover CG (csw = 1.5612, src/system/parameter_structs.jl:125) on the decomposed lattice: iteration
    count (must equal the single-GPU count, EXPECTED), iterations/s, true-residual check through the operator itself"""
    op = L.LqcdOp()
    op.kind, op.kappa, op.r, op.csw = L.WILSON, KAPPA, 1.0, 1.5612
    for i, b in enumerate(BC):
        op.bc[i] = b
    ctx.barrier()
    ctx.call("lqcd_clover_term", C.byref(op), None)         # leaves reach into the neighbour ranks' links (peer mapped)
    ctx.barrier()
    it, rs = C.c_int(0), C.c_double(0.0)
    q.clear_fermion_(sol)
    ctx.call("lqcd_solve", C.byref(op), sol.h, x.h, L.SOLVER_CG, L.OP_DDAGD, 1e-10, 3000, C.byref(it), C.byref(rs), None)     # warm-up + count
    n_conv = it.value
    q.clear_fermion_(sol)
    dist.barrier()
    t0 = time.perf_counter()
    ctx.call("lqcd_solve", C.byref(op), sol.h, x.h, L.SOLVER_CG, L.OP_DDAGD, 1e-10, 3000, C.byref(it), C.byref(rs), None)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    chk = q.FermionField(ctx, L.WILSON)
    ctx.call("lqcd_dslash", C.byref(op), chk.h, sol.h, L.OP_DDAGD)
    q.add_(chk, -1.0, x)
    true_rr = q.dot(chk, chk).real
    return {"ok": bool(true_rr < 1e-8 and it.value == n_conv), "csw": 1.5612, "cg_iters_eps1e-10": it.value, "resid_sq": rs.value, "true_resid_sq": true_rr,
            "cg_iters_per_s": it.value / dt, "ms_per_solve": dt * 1e3}


def _config5_rhmc_trajectory(dist, q, L, pg, rank, dev):
    """BASELINE configs[4]: one molecular-dynamics trajectory of staggered Nf = 2 RHMC (test/test_Nf2.toml scaled to 32^3 x 64): x^(-1/4)
    with 12 poles, every fermion force = ONE multi-shift CG + 12 accumulated outer products, leapfrog, all device resident
    (lqcd_md_trajectory_rational); reports dH, the multi-shift iterations and the wall time"""
    import numpy as np
    from lqcd_b200 import rhmc
    dims5 = tuple(int(v) for v in os.environ.get("LQCD_EXP_CONFIG5_DIMS", "32x32x32x64").split("x"))
    ctx = q.get_context(dims5, procgrid=pg, rank=rank, device=dev)
    q.connect_ranks(ctx, dist)
    ra = rhmc.rational_approx(-2 / 8.0, 12, 0.22, 17.0)              # mass 0.5: spec(DdagD) in [0.25, 16.25]
    al = np.ascontiguousarray(ra.alpha, dtype=np.float64)
    sh = np.ascontiguousarray(ra.beta, dtype=np.float64)
    pa, ps = al.ctypes.data_as(L.pdbl), sh.ctypes.data_as(L.pdbl)
    op = L.LqcdOp()
    op.kind, op.mass = L.STAGGERED, 0.5
    for i, b in enumerate(BC):
        op.bc[i] = b
    eta, tmp = q.FermionField(ctx, L.STAGGERED), q.FermionField(ctx, L.STAGGERED)
    q.gauss_distribution_fermion_(eta, 112)
    its, it1, S = C.c_longlong(0), C.c_int(0), C.c_double(0.0)
    steps, dtau = int(os.environ.get("LQCD_EXP_CONFIG5_STEPS", "5")), 0.02
    ctx.call("lqcd_gauge_random", 111, 0.3)
    ctx.call("lqcd_md_momenta_gaussian", 7)
    K, G = C.c_double(), C.c_double()
    H = []
    wall = 0.0
    for leg in range(2):
        ctx.call("lqcd_md_kinetic", C.byref(K)); ctx.call("lqcd_md_gauge_action", 5.7, C.byref(G))
        ctx.call("lqcd_rational_apply", C.byref(op), tmp.h, eta.h, float(ra.alpha0), pa, ps, len(al), 1e-18, 3000, C.byref(it1), C.byref(S))
        H.append(K.value + G.value + S.value)
        if leg == 0:
            dist.barrier()
            t0 = time.perf_counter()
            ctx.call("lqcd_md_trajectory_rational", C.byref(op), eta.h, pa, ps, len(al), 5.7, dtau, steps, 0, 1e-18, 3000, C.byref(its))
            ctx.synchronize()
            wall = time.perf_counter() - t0
    dH = H[1] - H[0]
    return {"ok": bool(abs(dH) < 0.05 * abs(S.value) and its.value > 0), "lattice": "x".join(map(str, dims5)), "poles": len(al), "md_steps": steps, "dtau": dtau,
            "dH": dH, "S_f": S.value, "multishift_cg_iters_total": its.value, "action_solve_iters": it1.value, "ms_per_trajectory": wall * 1e3,
            "rational_max_rel_err": ra.max_rel_err}


def run_experiments_multi(lattice, rank, local_rank, world, budget_s, barrier):
    """called by EVERY rank of the job (collective): returns {name: result} on rank 0, None elsewhere"""
    results = {}
    base_port = int(os.environ.get("MASTER_PORT", "29500"))
    t_start = time.perf_counter()
    for idx, (name, (env_extra, what)) in enumerate(EXPERIMENTS_MULTI.items()):
        barrier()                                       # all parents decide together (rank 0's clock is not shared: fixed schedule)
        if 60 + (idx + 1) * 34 > budget_s:
            results[name] = {"skipped": "time budget of the experiments leg", "what": what}
            continue
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(local_rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(base_port + 101 + idx), **env_extra)
        for k in ("TORCHELASTIC_RUN_ID", "GROUP_RANK", "ROLE_RANK", "LOCAL_WORLD_SIZE", "GROUP_WORLD_SIZE", "ROLE_WORLD_SIZE", "TORCHELASTIC_RESTART_COUNT",
                  "TORCHELASTIC_MAX_RESTARTS", "TORCHELASTIC_USE_AGENT_STORE", "TORCH_NCCL_ASYNC_ERROR_HANDLING"):
            env.pop(k, None)
        res = {"ok": False}
        try:
            r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--experiment-multi", name, "--lattice", lattice], env=env,
                               capture_output=True, text=True, timeout=150 if name == "default" else 75)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("EXPERIMENT ")]
            if line:
                res = json.loads(line[-1][len("EXPERIMENT "):])
                tl = [ln for ln in r.stderr.splitlines() if ln.startswith("[lqcd comm timeline")]
                if tl:
                    res["timeline_rank0"] = tl[-1]
            elif rank == 0:
                tail = (r.stdout + r.stderr).strip().splitlines()[-1:] or [""]
                res["error"] = f"exit {r.returncode}: {tail[0][:200]}"
        except subprocess.TimeoutExpired:
            res["error"] = "timed out"
        except Exception as exc:
            res["error"] = repr(exc)
        res["what"] = what
        results[name] = res
    barrier()
    ref = results.get("default", {}).get("norm_Dx_sq")
    for v in results.values():
        if ref and "norm_Dx_sq" in v:
            v["ok"] = bool(abs(v["norm_Dx_sq"] - ref) <= 1e-12 * abs(ref))
    results["wall_s"] = time.perf_counter() - t_start
    return results if rank == 0 else None


def choose_procgrid(n):
    # T first, then Z (north_star): keep T_local >= 8 where possible
    return {1: (1, 1, 1, 1), 2: (1, 1, 1, 2), 4: (1, 1, 1, 4), 8: (1, 1, 2, 4)}[n]


def run_b200(args, dims):
    # stdout must carry exactly ONE JSON line: park the real stdout, send everything else (library banners) to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import lqcd_b200 as q
    from lqcd_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    sys.excepthook = lambda t, v, tb: (sys.__excepthook__(t, v, tb), sys.stderr.flush(), os._exit(1))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # NCCL is host plumbing only (handle exchange, barrier, max-over-ranks).  Its NCCL_DEBUG banner goes wherever fd 1 points:
        # stderr, since the real stdout is parked above -- the JSON line stays alone on stdout and the log stays available.
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pg = choose_procgrid(world)
    if os.environ.get("LQCD_PROCGRID"):          # experiment knob, e.g. LQCD_PROCGRID=1,1,1,8
        pg = tuple(int(v) for v in os.environ["LQCD_PROCGRID"].split(","))
    ctx = q.get_context(dims, procgrid=pg, rank=rank, device=local_rank)
    if world > 1:
        q.connect_ranks(ctx, dist)
    ctx.call("lqcd_gauge_random", 111, -1.0)
    Vloc = int(np.prod(ctx.local_dims))
    V = int(np.prod(dims))
    op = L.LqcdOp()
    op.kind, op.kappa, op.r = L.WILSON, KAPPA, 1.0
    for i, b in enumerate(BC):
        op.bc[i] = b
    x, y = q.FermionField(ctx, L.WILSON), q.FermionField(ctx, L.WILSON)
    q.gauss_distribution_fermion_(x, 112)
    rt = Cudart()
    stream = ctx.stream()

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident Dslash: K timed applications, L2 flushed between ----------------
    # Timing rule: the K applications run back to back inside ONE CUDA-event bracket on the library stream
    # (no host sync inside, so ranks do not skew).  At N=1 the inputs (806 MB) exceed the 126 MB L2; the
    # L2-flushed per-application figure is reported next to it.  At N=8 the local working set (~100 MB) is
    # L2-resident by construction of the strong-scaling problem -- stated in config.
    # The bracket must not depend on how small --steps is: lqcd_time_dslash first runs two untimed applications (they align the
    # ranks through the halo flags, so no launch skew from the host barrier is measured), and at least MIN_TIMED applications are
    # timed; ms_per_step = bracket / applications, `steps` is reported as given.
    MIN_TIMED = 200 * world               # N = 8: 1600 applications of ~35 us, so the bracket is tens of ms at every N
    reps = max(args.steps, MIN_TIMED)
    mean, mn = C.c_double(), C.c_double()
    # the clock sampler (nvidia-smi polling GPU `local_rank` of rank 0 every 100 ms) starts BEFORE the warm-up: its start-up (fork,
    # NVML initialisation) must not fall into a bracket that is only a few ms long at N = 8
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        time.sleep(0.5)
    ctx.call("lqcd_time_dslash", C.byref(op), y.h, x.h, L.OP_D, max(args.warmup, 3), 0, C.byref(mean), C.byref(mn))
    barrier()
    l0 = ctx.launch_count()
    t_wall0 = time.perf_counter()
    ctx.call("lqcd_time_dslash", C.byref(op), y.h, x.h, L.OP_D, reps, 0, C.byref(mean), C.byref(mn))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = (ctx.launch_count() - l0) * reps // (reps + 2)          # kernels inside the event bracket (the two aligning applications excluded)
    ms = max_over_ranks(mean.value)
    norm_Dx_sq = q.dot(y, y).real                                      # global |D x|^2: independent of N and of every tuning knob
    ms_flushed = None
    if world == 1:
        ctx.call("lqcd_time_dslash", C.byref(op), y.h, x.h, L.OP_D, min(args.steps, 30), 1, C.byref(mean), C.byref(mn))
        ms_flushed = mean.value
    gflops = FLOP_PER_SITE * V / (ms * 1e-3) / 1e9
    gbs = BYTES_PER_SITE * V / (ms * 1e-3) / 1e9

    # ---------------- staggered Dslash (north_star names it next to Wilson): same lattice, same links ----------------
    sop = L.LqcdOp()
    sop.kind, sop.mass = L.STAGGERED, 0.5
    for i, b in enumerate(BC):
        sop.bc[i] = b
    sx, sy = q.FermionField(ctx, L.STAGGERED), q.FermionField(ctx, L.STAGGERED)
    q.gauss_distribution_fermion_(sx, 112)
    ctx.call("lqcd_time_dslash", C.byref(sop), sy.h, sx.h, L.OP_D, 3, 0, C.byref(mean), C.byref(mn))
    barrier()
    ctx.call("lqcd_time_dslash", C.byref(sop), sy.h, sx.h, L.OP_D, reps, 0, C.byref(mean), C.byref(mn))
    stag_ms = max_over_ranks(mean.value)
    stag_norm = q.dot(sy, sy).real
    stag_flushed = None
    if world == 1:
        ctx.call("lqcd_time_dslash", C.byref(sop), sy.h, sx.h, L.OP_D, 30, 1, C.byref(mean), C.byref(mn))
        stag_flushed = mean.value
    del sx, sy

    # ---------------- device-resident CG: iterations/s of solve_DinvX!(y, DdagD, b) ------------------
    ctx.call("lqcd_gauge_random", 111, 0.3)                         # warm field: realistic conditioning
    sol = q.FermionField(ctx, L.WILSON)
    it, rs = C.c_int(0), C.c_double(0.0)

    def cg_run(maxit):
        q.clear_fermion_(sol)
        st = ctx.lib.lqcd_solve(ctx.h, C.byref(op), sol.h, x.h, L.SOLVER_CG, L.OP_DDAGD, 0.0, maxit, C.byref(it), C.byref(rs), None)
        assert st in (L.LQCD_OK, L.ERR_NOCONV), ctx.lib.lqcd_last_error(ctx.h)
        return it.value

    cg_run(10)
    barrier()
    e0, e1 = rt.event(), rt.event()
    rt.record(e0, stream)
    n_it = cg_run(args.cg_iters)
    rt.record(e1, stream)
    cg_ms = max_over_ranks(rt.elapsed_ms(e0, e1))
    cg_ips = n_it / (cg_ms * 1e-3)
    # converged solve for the residual report (from a zero guess: the count must not depend on --cg-iters)
    it_conv, rs_conv = None, None
    q.clear_fermion_(sol)
    st = ctx.lib.lqcd_solve(ctx.h, C.byref(op), sol.h, x.h, L.SOLVER_CG, L.OP_DDAGD, 1e-10, 3000, C.byref(it), C.byref(rs), None)
    if st == L.LQCD_OK:
        it_conv, rs_conv = it.value, rs.value
    ctx.call("lqcd_gauge_random", 111, -1.0)
    # ---------------- e2e: host buffers through the public API ----------------------------------------
    e2e = None
    e2e_cg = None
    try:                              # every rank moves its LOCAL block (all N): collective calls, same count on every rank
        shape = x.host_shape
        hx = torch.empty(shape, dtype=torch.complex128).pin_memory()
        hy = torch.empty(shape, dtype=torch.complex128).pin_memory()
        hx.copy_(torch.from_numpy(x.to_host()))
        U = None
        D = q.DiracOperator.__new__(q.DiracOperator)                  # operator bound to the links already on the device
        D.op, D.ctx, D.kind, D.mode = op, ctx, L.WILSON, L.OP_D
        D.eps, D.maxsteps, D.verbose, D.method, D.last = 0.0, args.cg_iters, 1, "bicg", {}
        D._bound_epoch = getattr(ctx, "binding_epoch", None)           # bound to what the device holds (no host copy of these links)
        hxn, hyn = hx.numpy(), hy.numpy()

        def e2e_step_3call():
            ctx.call("lqcd_fermion_upload", x.h, hxn.ctypes.data, 0)
            q.mul_(y, D, x)
            ctx.call("lqcd_fermion_download", y.h, hyn.ctypes.data, 0)

        def e2e_step_pipe():          # the Julia shim's mul!(y, D, x) on host fields: one pipelined call (host_pipeline.cu)
            ctx.call("lqcd_dslash_host", C.byref(op), y.h, x.h, hyn.ctypes.data, hxn.ctypes.data, L.OP_D, 0)

        # the pipelined call must reproduce the three-call sequence bit for bit on this box, else it is not used
        e2e_step, e2e_call = e2e_step_3call, "x.from_host(h); mul_(y, D, x); y.to_host()  [lqcd_fermion_upload + lqcd_dslash + lqcd_fermion_download]"
        pipe_note, ref_y = None, None
        pipe_ok = os.environ.get("LQCD_E2E_PIPE", "1") != "0"
        if pipe_ok and world == 1:             # N > 1 takes the three-call sequence inside lqcd_dslash_host anyway
            pipe_ok, pipe_note = probe_pipe_isolated(args.lattice, local_rank)
        if pipe_ok:
            try:
                e2e_step_3call()
                ref_y = hyn.copy()
                hyn[...] = 0
                e2e_step_pipe()
                if np.array_equal(ref_y, hyn):
                    e2e_step, e2e_call = e2e_step_pipe, "mul_host_(y_h, D, x_h)  [lqcd_dslash_host: slab-pipelined H2D | convert + Dslash + convert | D2H]"
                else:
                    pipe_note = "pipelined call disagreed with the three-call sequence: not used"
            except Exception as exc2:
                pipe_note = f"pipelined call failed ({exc2!r}): not used"
            ref_y = None

        for _ in range(3):
            e2e_step()
        barrier()
        k = min(args.steps, 20)
        t0 = time.perf_counter()
        for _ in range(k):
            e2e_step()
        barrier()
        dt = max_over_ranks((time.perf_counter() - t0) / k)
        nbytes = hx.numel() * 16 * world          # all ranks together
        e2e = {"value": FLOP_PER_SITE * V / dt / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "ms_per_step": dt * 1e3, "call": e2e_call}
        if pipe_note:
            e2e["note"] = pipe_note
        if e2e_step is e2e_step_pipe:        # also time the unpipelined sequence, for the record
            barrier()
            t0 = time.perf_counter()
            for _ in range(min(k, 5)):
                e2e_step_3call()
            barrier()
            e2e["ms_per_step_three_calls"] = max_over_ranks((time.perf_counter() - t0) / min(k, 5)) * 1e3
        # CG through host buffers: upload source, solve, download solution
        ctx.call("lqcd_gauge_random", 111, 0.3)
        t0 = time.perf_counter()
        ctx.call("lqcd_fermion_upload", x.h, hxn.ctypes.data, 0)
        n2 = cg_run(args.cg_iters)
        ctx.call("lqcd_fermion_download", sol.h, hyn.ctypes.data, 0)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e_cg = {"value": n2 / dt, "unit": "CG iterations/s", "iters": n2, "h2d_bytes": nbytes, "d2h_bytes": nbytes}
    except Exception as exc:           # never lose the device-resident result because of the host-buffer leg
        e2e = e2e or {"error": repr(exc)}
        e2e_cg = e2e_cg or {"error": repr(exc)}

    clocks = sampler.stop() if sampler else None

    # ---------------- CPU baseline (oracle port, bounded sample) ---------------------------------------
    # At N = 1 the oracle runs on the links and the source the GPU actually holds (downloaded), and its result is compared with
    # the GPU's y = D x of the timed region: parity asserted at the bench size, in the bench run.
    cpu = None
    parity = {"norm_Dx_sq": norm_Dx_sq, "staggered_norm_Dx_sq": stag_norm, "cg_converged_iters_eps1e-10": it_conv}
    if rank == 0 and world == 1 and not args.no_cpu:
        thr = host_threads()
        ctx.call("lqcd_gauge_random", 111, -1.0)                      # the timed region's hot links (the CG legs used a warm field)
        ctx.call("lqcd_dslash", C.byref(op), y.h, x.h, L.OP_D)
        r = cpu_dslash(dims, 3, 1, thr, U=q.get_links(ctx), x=x.to_host())
        dev = float(np.abs(y.to_host() - r["y"]).max() / np.abs(r["y"]).max())
        parity["max_rel_dev_vs_oracle"] = dev
        parity["oracle_ok"] = bool(dev < 1e-13)
        cpu = {"value": r["gflops"], "unit": "GFLOP/s", "cores": r["threads"], "kind": "port",
               "sample": f"3 full-lattice Wilson applications at {args.lattice} ({r['ms']:.0f} ms each) on the device's own links and source, oracle/lqcd_oracle.c with OpenMP"}
    # N-independence: |D x|^2 and the converged CG iteration count must equal the single-GPU values (committed constants,
    # re-derived at N = 1 against the oracle above)
    exp = EXPECTED.get(args.lattice)
    if exp:
        parity["expected"] = exp
        def close(a, b):
            return b is None or abs(a - b) <= 1e-12 * abs(b)
        parity["n_independent_ok"] = bool(close(norm_Dx_sq, exp["norm_Dx_sq"]) and close(stag_norm, exp["staggered_norm_Dx_sq"])
                                          and (exp["cg_converged_iters_eps1e-10"] is None or it_conv == exp["cg_converged_iters_eps1e-10"]))
    parity["ok"] = bool(parity.get("oracle_ok", True) and parity.get("n_independent_ok", True))

    # the headline line is complete here; the experiments leg can only ADD a key to it
    line = None
    if rank == 0:
        peak, peak_src = peaks()
        traffic = None
        tp = ROOT / "profiles" / "wilson_dslash_traffic.json"
        if tp.exists():
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        line = {
            "metric": metric_name(), "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(args.lattice),
                       "procgrid": list(pg), "l2": f"not flushed: per-GPU inputs {806 // world} MB vs 126 MB L2 (N=8: L2-resident by strong scaling); ms_flushed = per-application time with a 512 MB memset between applications (N=1 only)",
                       "timing": f"one CUDA-event bracket around max(steps, {MIN_TIMED} = 200 x GPUs) back-to-back applications on the library stream after two untimed aligning applications, / applications, max over ranks",
                       "applications_timed": reps, "ms_flushed": ms_flushed, "wall_s_timed_region": t_wall,
                       # the CG half of the metric, where the driver keeps it
                       "cg_iters_per_s": cg_ips, "cg_iters_timed": n_it,
                       "e2e_cg_iters_per_s": (e2e_cg or {}).get("value"),
                       "staggered_dslash_ms": stag_ms, "staggered_dslash_gflops": STAG_FLOP_PER_SITE * V / (stag_ms * 1e-3) / 1e9,
                       "parity": parity},
            # per GPU: each GPU moves 1/N of the algorithmic bytes per application against ITS OWN HBM peak
            "roofline": {"bound": "hbm", "achieved": gbs / world, "peak": peak, "unit": "GB/s", "frac": gbs / world / peak, "traffic": traffic if world == 1 else None,
                         "per": "GPU", "achieved_aggregate": gbs,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_SITE * V // world, "kernel": "wilson_dslash_kernel",
                         "flushed_frac": (BYTES_PER_SITE * V / (ms_flushed * 1e-3) / 1e9 / peak) if ms_flushed else None,
                         "staggered": {"kernel": "staggered_dslash_kernel", "ms": stag_ms, "ms_flushed": stag_flushed,
                                       "achieved": STAG_BYTES_PER_SITE * V / (stag_ms * 1e-3) / 1e9 / world,
                                       "frac": STAG_BYTES_PER_SITE * V / (stag_ms * 1e-3) / 1e9 / world / peak,
                                       "algorithmic_bytes_per_launch": STAG_BYTES_PER_SITE * V // world}},
            "cg": {"iters_per_s": cg_ips, "iters": n_it, "ms": cg_ms, "roofline_frac_unfused": CG_BYTES_PER_SITE * V * cg_ips / 1e9 / peak,
                   "converged_iters_eps1e-10": it_conv, "resid_sq": rs_conv, "field": "warm eps=0.3"},
            "e2e": e2e, "e2e_cg": e2e_cg, "cpu_baseline": cpu, "gpu_launches": launches, "clocks": clocks,
            "experiments": None,
        }
    emitted = threading.Lock()

    def emit(experiments):
        if rank != 0 or not emitted.acquire(blocking=False):       # exactly one JSON line, whoever gets here first
            return
        line["experiments"] = experiments
        out = os.environ.get("LQCD_BENCH_EXPERIMENTS_OUT")          # builder-side runs keep the whole dict (profiles/r2_experiments_n{N}.json)
        if out and experiments:
            try:
                Path(out).write_text(json.dumps(experiments, indent=1))
            except OSError:
                pass
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    want = os.environ.get("LQCD_BENCH_EXPERIMENTS", "1") != "0"
    budget = float(os.environ.get("LQCD_BENCH_EXPERIMENTS_S", "150"))
    watchdog = None
    if want:
        # (every rank arms it: a rank that is stuck behind a lost peer must not keep the launcher waiting either)
        # if the experiments leg (child processes, at N > 1 also barriers between the parents) overruns badly, the measured line is
        # printed without it and the process ends: the headline can never be lost to the diagnostics
        def overrun():
            emit({"error": "experiments leg overran its time limit; headline printed by the watchdog"})
            os._exit(0)
        watchdog = threading.Timer(float(os.environ.get("LQCD_BENCH_WATCHDOG_S", 2.0 * budget + 120.0)), overrun)
        watchdog.daemon = True
        watchdog.start()
    experiments = None
    if want and world == 1 and rank == 0:
        try:
            experiments = run_experiments(args.lattice, local_rank, budget)
        except Exception as exc:
            experiments = {"error": repr(exc)}
    if want and world > 1:
        try:
            experiments = run_experiments_multi(args.lattice, rank, local_rank, world, budget, barrier)
        except Exception as exc:
            experiments = {"error": repr(exc)}
    if watchdog is not None:
        watchdog.cancel()
    emit(experiments)
    if rank == 0 and not parity["ok"]:          # a fast kernel whose results differ from the oracle's / from N = 1 is not done
        sys.stderr.write(f"bench.py: PARITY FAILED: {json.dumps(parity)}\n")
        sys.stderr.flush()
        os._exit(3)


def main():
    args = parse_args()
    dims = tuple(int(v) for v in args.lattice.split("x"))
    if args.experiment_multi:
        experiment_multi_child(args.experiment_multi, dims)
    elif args.experiment:
        experiment_child(args.experiment, dims)
    elif args.probe_pipe:
        probe_pipe_child(dims)
    elif args.impl == "reference":
        run_reference(args, dims)
    else:
        run_b200(args, dims)


if __name__ == "__main__":
    main()

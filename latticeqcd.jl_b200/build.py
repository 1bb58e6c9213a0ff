"""
Builds liblqcd_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache): the .so travels to the GPU
box with the repo snapshot.  Usage: python build.py [--force] [--verbose]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "liblqcd_b200.so"
SOURCES = ["context.cu", "wilson_dslash.cu", "wilson_tmarch.cu", "links12.cu", "staggered_dslash.cu", "blas.cu", "solvers.cu", "comm.cu", "force.cu", "clover.cu", "wilson_clover.cu", "wilson_eo.cu", "host_pipeline.cu", "gauge_md.cu", "mrhs.cu", "gauge_io.cu", "staggered_eo.cu", "wilson_general_r.cu", "clover_force.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-ccbin", HOSTCXX, "-Xcompiler", "-fPIC,-O2", "-fmad=true"]


def _deps():
    return list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "lqcd_b200.h"]


def build(force=False, verbose=False):
    # experiment builds: LQCD_BUILD_DEFS="-DLQCD_LINK_HINT=2" LQCD_BUILD_SUFFIX=_hint2 -> liblqcd_b200_hint2.so
    global OUT
    defs = os.environ.get("LQCD_BUILD_DEFS", "").split()
    suffix = os.environ.get("LQCD_BUILD_SUFFIX", "")
    if suffix:
        OUT = HERE / f"liblqcd_b200{suffix}.so"
    objdir = HERE / ("build" + suffix)
    objdir.mkdir(exist_ok=True)
    newest_dep = max(p.stat().st_mtime for p in _deps())
    jobs = []
    for src in SOURCES:
        s = CSRC / src
        o = objdir / (src + ".o")
        if force or not o.exists() or o.stat().st_mtime < max(s.stat().st_mtime, newest_dep):
            cmd = [NVCC, *FLAGS, *defs, "-c", str(s), "-o", str(o)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)
    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + cmd[-3])
    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, jobs))
    objs = [str(objdir / (s + ".o")) for s in SOURCES]
    if jobs or not OUT.exists():
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", HOSTCXX, "-o", str(OUT), *objs, "-lcudart"]
        run(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

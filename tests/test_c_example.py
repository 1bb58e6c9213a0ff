"""The boundary is a C ABI: include/lqcd_b200.h must be valid C (not only C++), and a plain-C client must be able to drive the
path.  examples/propagator.c computes the 12 point-source propagators and the pion correlator through lqcd_solve_multi."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "latticeqcd.jl_b200")]
from oracle import oracle as orc          # noqa: E402


def _compile(tmp_path, libdir, libname):
    exe = tmp_path / "propagator"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "propagator.c"),
           f"-L{libdir}", f"-l:{libname}", f"-Wl,-rpath,{libdir}", "-lm", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_valid_c99_and_cxx(tmp_path):
    src = tmp_path / "hdr.c"
    src.write_text('#include "lqcd_b200.h"\nint main(void) { return lqcd_abi_version() == LQCD_ABI_VERSION ? 0 : 1; }\n')
    for comp, std, extra in (("gcc", "-std=c99", []), ("gcc", "-std=c11", []), ("g++", "-std=c++11", ["-x", "c++"])):
        r = subprocess.run([comp, std, "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}", "-fsyntax-only", *extra, str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, (comp, std, r.stderr)


def test_c_example_compiles_against_the_library_and_fails_loudly_without_a_gpu(tmp_path):
    exe = _compile(tmp_path, ROOT / "latticeqcd.jl_b200", "liblqcd_b200.so")
    r = subprocess.run([str(exe), "4", "4", "4", "4"], capture_output=True, text=True)
    if r.returncode != 0:                         # the build container: no GPU -> status LQCD_ERR_NOGPU and a message, never a fallback
        assert "no CPU fallback" in r.stderr or "sm_100a" in r.stderr, r.stderr


def check_against_oracle(exe, golden_dir, tmp_path, env=None):
    import lqcd_b200 as q
    U = np.load(golden_dir / "wilson_4444.npy")
    q.save_binarydata(U, tmp_path / "w.ildg")
    r = subprocess.run([str(exe), "4", "4", "4", "4", str(tmp_path / "w.ildg")], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = r.stdout.splitlines()
    plaq = float([ln for ln in lines if ln.startswith("plaquette")][0].split()[1])
    assert abs(plaq - 0.565800226845) < 1e-11                        # SURVEY.md section 4 value for this fixture
    corr = np.array([float(ln.split("=")[1]) for ln in lines if ln.startswith("C(")])
    iters = [int(ln.split(":")[1].split()[0]) for ln in lines if ln.startswith("source")]
    op = orc.make_op((4, 4, 4, 4), kappa=0.141139)
    ref, ref_it = np.zeros(4), []
    for i in range(12):
        b = np.zeros((4, 4, 4, 4, 4, 3), dtype=complex)
        b[i % 4, 0, 0, 0, 0, i // 4] = 1.0
        s = orc.cgnr(op, orc.WILSON, U, b, eps=1e-19)
        ref += (np.abs(s["x"]) ** 2).sum(axis=(0, 2, 3, 4, 5))
        ref_it.append(s["iters"])
    assert iters == ref_it
    assert np.abs(corr - ref).max() / ref.max() < 1e-9


@pytest.mark.gpu
def test_c_example_on_the_device(tmp_path, golden_dir):
    exe = _compile(tmp_path, ROOT / "latticeqcd.jl_b200", "liblqcd_b200.so")
    check_against_oracle(exe, golden_dir, tmp_path)

import os
import sys
import time
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "latticeqcd.jl_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


@pytest.hookimpl(tryfirst=True)
def pytest_cmdline_main(config):
    """The CPU suite (`-m "not gpu"`) spends most of its time in independent subprocess-based pre-flight tests: run it on 4 xdist
    workers unless the caller chose a distribution.  GPU runs (`-m gpu`) stay serial and in collection order."""
    opt = config.option
    if (getattr(opt, "markexpr", "") == "not gpu" and hasattr(opt, "numprocesses") and not opt.numprocesses
            and not os.environ.get("PYTEST_XDIST_WORKER") and not getattr(opt, "collectonly", False)
            and getattr(opt, "dist", "no") == "no" and os.environ.get("LQCD_TEST_SERIAL") != "1"):
        os.environ.setdefault("OMP_NUM_THREADS", "2")          # the oracle's OpenMP teams would oversubscribe the workers' cores
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        opt.numprocesses = 4
        opt.dist = "load"
        opt.tx = ["popen"] * 4
        sys.stderr.write("[conftest] -m \"not gpu\": running on 4 xdist workers with OMP_NUM_THREADS=2 (LQCD_TEST_SERIAL=1 or an explicit -n / --dist keeps your settings)\n")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # Build the CUDA library (nvcc cross-compiles without a GPU; incremental, a no-op when up to date) and the oracle ONCE, in
    # the controlling process, before any xdist worker starts: workers must never race on a missing / half-written .so.
    # (`-m gpu` runs on the B200 box use the library that travelled with the snapshot: nothing is rebuilt there.)
    if (not os.environ.get("PYTEST_XDIST_WORKER") and os.environ.get("LQCD_TEST_NO_BUILD") != "1"
            and getattr(config.option, "markexpr", "") != "gpu"):
        try:
            import importlib.util
            import shutil
            if shutil.which("nvcc") or Path("/usr/local/cuda/bin/nvcc").exists():
                spec = importlib.util.spec_from_file_location("lqcd_b200_build", ROOT / "latticeqcd.jl_b200" / "build.py")
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                mod.build()
            from oracle import oracle as _orc
            _orc.build()
        except Exception as exc:          # a stale / missing library must not surface later as unrelated test errors
            raise pytest.UsageError(f"[conftest] building liblqcd_b200 / the oracle failed: {exc!r} (LQCD_TEST_NO_BUILD=1 skips the build)")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"

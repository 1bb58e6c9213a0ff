"""Runs bench.py's headline path (run_b200) end to end against the EMULATED library on a tiny lattice (tests/test_emu_preflight.py).
torch.cuda and the cudart event wrapper are stubbed (no GPU in the build container); everything else -- argument handling, the
probe of the pipelined host call, the e2e legs, the cpu_baseline leg, the experiments leg and the assembly of the JSON line -- is
the code that runs on the B200 box.  Purpose: a Python-level mistake in bench.py must not be discovered at round end."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench        # noqa: E402

torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self, *a, **k: self


class FakeCudart:
    def event(self):
        return [0.0]

    def record(self, e, stream):
        e[0] = time.perf_counter()

    def elapsed_ms(self, e0, e1):
        return max((e1[0] - e0[0]) * 1e3, 1e-3)


bench.Cudart = FakeCudart
sys.argv = ["bench.py"] + sys.argv[1:]
bench.main()

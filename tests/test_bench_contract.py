"""bench.py contract checks that need no GPU: the reference arm (CPU oracle) prints exactly one JSON line with the
keys the driver reads, and non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--lattice", "8x8x8x8", "--steps", "2", "--warmup", "1"],
                          capture_output=True, text=True, timeout=300, env=env)


def test_reference_arm_json_line():
    r = run()
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert d["impl"] == "reference" and d["metric"] == base["metric"] and d["unit"] == "GFLOP/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["value"] > 0


def test_reference_arm_other_ranks_are_silent():
    r = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_honours_steps_and_shares_the_workload_string():
    """the driver compares steps / warmup / config.workload of the two arms: the CPU arm runs exactly the requested steps and prints
    the same workload string as the B200 arm; the live-reference probe (julia, baseline/_ref) is recorded in the line"""
    sys.path.insert(0, str(ROOT))
    import bench
    r = run()
    d = json.loads(r.stdout.strip())
    assert d["steps"] == 2 and d["warmup"] == 1
    assert d["config"]["workload"] == bench.workload_string("8x8x8x8")
    assert "live reference probe" in d["config"]["arm"] and '"julia"' in d["config"]["arm"]
    assert d["cpu_baseline"]["sample"].startswith("2 full-lattice applications")

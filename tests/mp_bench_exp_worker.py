"""parents of bench.py's multi-rank experiments leg (tests/test_emu_preflight.py): every rank calls run_experiments_multi, which
spawns one child per rank and experiment; rank 0 prints the collected results as JSON"""
import json
import os
import sys
from pathlib import Path

import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench      # noqa: E402

dist.init_process_group("gloo")
res = bench.run_experiments_multi(sys.argv[1], dist.get_rank(), int(os.environ.get("LOCAL_RANK", "0")), dist.get_world_size(), 600.0, dist.barrier)
if dist.get_rank() == 0:
    print("RESULTS " + json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()

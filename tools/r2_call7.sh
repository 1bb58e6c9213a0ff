#!/bin/bash
# call 7 (8 GPUs): N=8 bench with the multi-rank experiments (timeline, pack modes, tmarch, configs 4 and 5)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
LQCD_BENCH_EXPERIMENTS_OUT=gpurun_out/c7_experiments_n8.json LQCD_BENCH_EXPERIMENTS_S=240 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c7_bench_n8.json 2> gpurun_out/c7_bench_n8.err
python - <<'P'
import json
try:
    d=json.load(open('gpurun_out/c7_bench_n8.json')); e=d.pop('experiments'); print(json.dumps(d)[:2600])
    for k,v in (e or {}).items(): print(k, json.dumps({a:b for a,b in v.items() if a!='what'})[:1400] if isinstance(v,dict) else v)
except Exception as ex: print("bench parse failed", ex)
P
tail -5 gpurun_out/c7_bench_n8.err

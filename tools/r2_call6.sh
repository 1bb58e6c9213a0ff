#!/bin/bash
# call 6 (2 GPUs): occupancy variant of the two-row-link kernel; N=2 bench with the multi-rank experiments (timeline, pack modes, tmarch, config 4); multi-rank parity on 2 real GPUs
mkdir -p gpurun_out
nvidia-smi -L
for lat in 32x32x32x32 32x32x32x16 32x32x16x8; do
  for lib in liblqcd_b200.so liblqcd_b200_mb4.so; do
    echo -n "$lat $lib: "; LQCD_B200_LIB=latticeqcd.jl_b200/$lib timeout 60 python tools/quick_bench.py $lat 2>&1 | grep -E "wilson" | sed 's/^\([a-z]*\) .*flush=\([01]\) mean *\([0-9.]*\) us.*/\1 flush=\2 \3 us;/' | tr '\n' ' '; echo
  done
done > gpurun_out/c6_mb4.txt 2>&1
cat gpurun_out/c6_mb4.txt
LQCD_BENCH_EXPERIMENTS_OUT=gpurun_out/c6_experiments_n2.json LQCD_BENCH_EXPERIMENTS_S=420 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c6_bench_n2.json 2> gpurun_out/c6_bench_n2.err
python - <<'P'
import json
try:
    d=json.load(open('gpurun_out/c6_bench_n2.json')); e=d.pop('experiments'); print(json.dumps(d)[:2500])
    for k,v in (e or {}).items(): print(k, json.dumps({a:b for a,b in v.items() if a!='what'})[:900] if isinstance(v,dict) else v)
except Exception as ex: print("bench parse failed", ex)
P
tail -5 gpurun_out/c6_bench_n2.err
timeout 600 python -m pytest tests/test_multirank.py -m gpu -q -x -p no:cacheprovider > gpurun_out/c6_pytest_multirank.txt 2>&1; tail -4 gpurun_out/c6_pytest_multirank.txt

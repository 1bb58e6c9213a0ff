#!/bin/bash
# call 11 (8-GPU box): bench lines at N = 8 (with the default experiment child: configs 4 and 5), 4, 2; multi-rank parity with one GPU per rank
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
LQCD_BENCH_EXPERIMENTS_OUT=gpurun_out/c11_experiments_n8.json LQCD_BENCH_EXPERIMENTS_S=100 timeout 400 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c11_bench_n8.json 2> gpurun_out/c11_bench_n8.err
LQCD_BENCH_EXPERIMENTS=0 timeout 300 $TR --nproc-per-node 4 --master-port 29543 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/c11_bench_n4.json 2> gpurun_out/c11_bench_n4.err
LQCD_BENCH_EXPERIMENTS=0 timeout 300 $TR --nproc-per-node 2 --master-port 29545 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c11_bench_n2.json 2> gpurun_out/c11_bench_n2.err
python - <<'P'
import json
for n in (8,4,2):
    try:
        d=json.load(open(f'gpurun_out/c11_bench_n{n}.json')); e=d.pop('experiments',None)
        print(n, d['ms_per_step'], d['value'], d['config']['cg_iters_per_s'], d['config']['staggered_dslash_ms'], d['config']['parity']['ok'], d['roofline']['frac'], d['e2e']['value'], d['config']['e2e_cg_iters_per_s'])
        if e: print(json.dumps(e)[:2500])
    except Exception as ex: print(n, "parse failed", ex)
P
tail -2 gpurun_out/c11_bench_n8.err
timeout 600 python -m pytest tests/test_multirank.py -m gpu -q -p no:cacheprovider > gpurun_out/c11_pytest_multirank.txt 2>&1; tail -4 gpurun_out/c11_pytest_multirank.txt

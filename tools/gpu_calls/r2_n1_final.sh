#!/bin/bash
# call 10 (1 GPU): full strict GPU suite on the final code, N=1 bench line (experiments off), ncu launch list of bench.py
mkdir -p gpurun_out
timeout 60 python tools/debug/tm_probe.py 32x8x8x16 2>&1 | tail -1 | tee gpurun_out/c10_probe.txt
grep -q "rel err [0-9.]*e-1[5-9]" gpurun_out/c10_probe.txt || { echo "probe failed: stop"; exit 1; }
LQCD_BENCH_EXPERIMENTS=0 timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/c10_bench_n1.json 2> gpurun_out/c10_bench_n1.err; cut -c1-3000 gpurun_out/c10_bench_n1.json; tail -3 gpurun_out/c10_bench_n1.err
LQCD_BENCH_EXPERIMENTS=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/c10_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --cg-iters 20 > gpurun_out/c10_launches_bench.log 2>&1
tail -c 400 gpurun_out/c10_launches_bench.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c10_pytest.txt 2>&1; tail -6 gpurun_out/c10_pytest.txt

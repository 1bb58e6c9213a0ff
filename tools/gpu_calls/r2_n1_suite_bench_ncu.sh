#!/bin/bash
# call 9 (1 GPU): strict GPU suite, N=1 bench + experiments, reference arm, ncu (full capture of the two Dslash kernels, launch list of bench.py)
mkdir -p gpurun_out
timeout 60 python tools/debug/tm_probe.py 32x8x8x16 2>&1 | tail -1 | tee gpurun_out/c9_probe.txt
grep -q "rel err [0-9.]*e-1[5-9]" gpurun_out/c9_probe.txt || { echo "probe failed: stop"; exit 1; }
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c9_pytest.txt 2>&1; tail -6 gpurun_out/c9_pytest.txt
LQCD_BENCH_EXPERIMENTS_OUT=gpurun_out/c9_experiments_n1.json LQCD_BENCH_EXPERIMENTS_S=300 timeout 800 python bench.py --steps 20 --warmup 5 > gpurun_out/c9_bench_n1.json 2> gpurun_out/c9_bench_n1.err
python - <<'P'
import json
try:
    d=json.load(open('gpurun_out/c9_bench_n1.json')); e=d.pop('experiments'); print(json.dumps(d)[:3500])
    for k,v in (e or {}).items(): print(k, json.dumps({a:b for a,b in v.items() if a!='what'})[:700] if isinstance(v,dict) else v)
except Exception as ex: print("bench parse failed", ex)
P
tail -3 gpurun_out/c9_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c9_bench_reference.json 2>/dev/null; cut -c1-900 gpurun_out/c9_bench_reference.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"wilson_dslash_kernel|staggered_dslash_kernel" --launch-skip 6 -c 2 -f -o gpurun_out/c9_ncu_wilson python tools/quick_bench.py 32x32x32x32 > gpurun_out/c9_ncu_wilson.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:staggered_dslash_kernel --launch-skip 6 -c 2 -f -o gpurun_out/c9_ncu_stag python tools/quick_bench.py 32x32x32x32 > gpurun_out/c9_ncu_stag.log 2>&1
LQCD_BENCH_EXPERIMENTS=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/c9_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --cg-iters 20 > gpurun_out/c9_launches_bench.log 2>&1
tail -2 gpurun_out/c9_launches_bench.log | cut -c1-300

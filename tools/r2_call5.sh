#!/bin/bash
# call 5 (1 GPU): two-row links -- parity subset, timing (full vs two-row links, Wilson + staggered, four local volumes), multi-RHS groupings, bench
mkdir -p gpurun_out
timeout 60 python tools/debug/tm_probe.py 32x8x8x16 2>&1 | tail -1 | tee gpurun_out/c5_probe.txt
grep -q "rel err [0-9.]*e-1[5-9]" gpurun_out/c5_probe.txt || { echo "probe failed: stop"; exit 1; }
for lat in 32x32x32x32 32x32x32x16 32x32x32x8 32x32x16x8 16x16x16x16; do
  for cfg in "A=1" "LQCD_LINKS12=0"; do
    echo -n "$lat $cfg: "; env $cfg timeout 60 python tools/quick_bench.py $lat 2>&1 | grep -E "wilson|stag" | sed 's/^\([a-z]*\) .*flush=\([01]\) mean *\([0-9.]*\) us.*/\1 flush=\2 \3 us;/' | tr '\n' ' '; echo
  done
done > gpurun_out/c5_sweep.txt 2>&1
cat gpurun_out/c5_sweep.txt
timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py tests/test_gpu_extended.py tests/test_md.py tests/test_rhmc.py -m gpu -q -x -p no:cacheprovider > gpurun_out/c5_pytest.txt 2>&1; tail -5 gpurun_out/c5_pytest.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wilson_dslash_kernel --launch-skip 6 -c 2 -f -o gpurun_out/c5_ncu_k1g12 python tools/quick_bench.py 32x32x32x32 > gpurun_out/c5_ncu.log 2>&1
LQCD_BENCH_EXPERIMENTS_OUT=gpurun_out/c5_experiments_n1.json LQCD_BENCH_EXPERIMENTS_S=240 timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c5_bench.json')); e=d.pop('experiments'); print(json.dumps(d)[:3000]);
for k,v in (e or {}).items(): print(k, {a:b for a,b in v.items() if a!='what'})
"; tail -3 gpurun_out/c5_bench.err

#!/bin/bash
# call 4: (1) which mbarrier wait loop hangs?  (2) if the default build is healthy: timing sweep, ncu, parity tests, bench
mkdir -p gpurun_out
probe() { echo "== $1 $2"; env LQCD_B200_LIB=latticeqcd.jl_b200/$1 timeout 30 python tools/debug/tm_probe.py $2 2>&1 | tail -2; echo "rc=$?"; }
{ probe liblqcd_b200.so 32x4x4x4; probe liblqcd_b200_asmloop.so 32x4x4x4; probe liblqcd_b200.so 8x8x8x8; probe liblqcd_b200.so 32x8x8x16; } > gpurun_out/c4_probe.txt 2>&1
cat gpurun_out/c4_probe.txt
if ! grep -q "probe (32, 8, 8, 16): rel err [0-9.]*e-1[5-9]" gpurun_out/c4_probe.txt; then echo "default build unhealthy: stop"; exit 1; fi
timeout 300 python tests/tmarch_worker.py > gpurun_out/c4_tmarch_worker.txt 2>&1; tail -2 gpurun_out/c4_tmarch_worker.txt
grep -q "TMARCH OK" gpurun_out/c4_tmarch_worker.txt || { echo "worker failed: stop"; exit 1; }
for lat in 32x32x32x32 32x32x32x16 32x32x16x8 16x16x16x16; do
  for cfg in "A=1" "LQCD_TM_CHUNKS=1" "LQCD_TM_CHUNKS=2" "LQCD_TM_CHUNKS=4" "LQCD_TM_CHUNKS=8" "LQCD_WILSON_KERNEL=1"; do
    echo -n "$lat $cfg: "; env $cfg timeout 60 python tools/quick_bench.py $lat 2>&1 | grep wilson | sed 's/.*flush=\([01]\) mean *\([0-9.]*\) us.*/flush=\1 \2 us/' | tr '\n' ' '; echo
  done
done > gpurun_out/c4_sweep.txt 2>&1
cat gpurun_out/c4_sweep.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wilson_tmarch --launch-skip 6 -c 2 -f -o gpurun_out/c4_ncu_tm python tools/quick_bench.py 32x32x32x32 > gpurun_out/c4_ncu_tm.log 2>&1
timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider > gpurun_out/c4_pytest.txt 2>&1; tail -5 gpurun_out/c4_pytest.txt
LQCD_BENCH_EXPERIMENTS=0 timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; tail -c 2500 gpurun_out/c4_bench.json; tail -5 gpurun_out/c4_bench.err

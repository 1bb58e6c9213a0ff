#!/bin/bash
# Round-2 call 2 (1 GPU): t-marching TMA kernel -- correctness on hardware, timing per local volume / chunking, ncu; BASELINE-size parity tests; new bench line
set -x
mkdir -p gpurun_out
timeout 600 python tests/tmarch_worker.py > gpurun_out/c2_tmarch_worker.txt 2>&1; tail -3 gpurun_out/c2_tmarch_worker.txt
for lat in 32x32x32x32 32x32x32x16 32x32x32x8 32x32x16x8 16x16x16x16; do
  for cfg in "A=1" "LQCD_TM_CHUNKS=1" "LQCD_TM_CHUNKS=2" "LQCD_TM_CHUNKS=4" "LQCD_TM_CHUNKS=8" "LQCD_WILSON_KERNEL=1"; do
    echo -n "$lat $cfg: "; env $cfg timeout 120 python tools/quick_bench.py $lat 2>&1 | grep wilson | sed 's/.*flush=\([01]\) mean *\([0-9.]*\) us.*/flush=\1 \2 us/' | tr '\n' ' '; echo
  done
done > gpurun_out/c2_sweep.txt 2>&1
cat gpurun_out/c2_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wilson_tmarch --launch-skip 6 -c 2 -f -o gpurun_out/c2_ncu_tm python tools/quick_bench.py 32x32x32x32 > gpurun_out/c2_ncu_tm.log 2>&1
timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider > gpurun_out/c2_pytest.txt 2>&1; tail -5 gpurun_out/c2_pytest.txt
LQCD_BENCH_EXPERIMENTS=0 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; tail -c 1500 gpurun_out/c2_bench.json; tail -5 gpurun_out/c2_bench.err

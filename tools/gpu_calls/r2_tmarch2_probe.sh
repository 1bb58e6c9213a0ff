#!/bin/bash
# 1 GPU, short: second-generation t-marching kernel (LQCD_WILSON_KERNEL=5: two CTAs per SM, two-row link planes, carried t-backward
# hop).  Bounded probe first (a hang costs 60 s, not the call), then correctness worker, then timing against the default kernel.
mkdir -p gpurun_out
LQCD_WILSON_KERNEL=5 timeout 60 python tools/quick_bench.py 16x16x16x16 2>&1 | grep -E "wilson" | tee gpurun_out/c17_probe.txt
grep -q "wilson" gpurun_out/c17_probe.txt || { echo "probe failed or hung: stop"; exit 1; }
LQCD_WILSON_KERNEL=5 timeout 150 python tests/tmarch_worker.py 2>&1 | tail -6 | tee gpurun_out/c17_worker.txt
grep -q "TMARCH OK" gpurun_out/c17_worker.txt || { echo "worker failed: stop"; exit 1; }
for cfg in "A=1" "LQCD_WILSON_KERNEL=5" "LQCD_WILSON_KERNEL=5 LQCD_TM_CHUNKS=1" "LQCD_WILSON_KERNEL=5 LQCD_TM_CHUNKS=2" "LQCD_WILSON_KERNEL=5 LQCD_TM_CHUNKS=4" "LQCD_WILSON_KERNEL=5 LQCD_TM_CHUNKS=8" "LQCD_WILSON_KERNEL=5 LQCD_TM_CHUNKS=16"; do
  echo -n "32x32x32x32 $cfg: "; env $cfg timeout 40 python tools/quick_bench.py 32x32x32x32 2>&1 | grep -E "wilson" | sed 's/^\([a-z]*\) .*flush=\([01]\) mean *\([0-9.]*\) us.*/\1 flush=\2 \3 us;/' | tr '\n' ' '; echo
done 2>&1 | tee gpurun_out/c17_tm2_sweep.txt

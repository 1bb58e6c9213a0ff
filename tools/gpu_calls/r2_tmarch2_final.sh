#!/bin/bash
# 1 GPU, <= 40 s: second-generation t-marching kernel with next-task L2 prefetch: correctness worker (auto chunking and 2 chunks),
# then 32^4 timing with the prefetch on / off.
mkdir -p gpurun_out
LQCD_WILSON_KERNEL=5 timeout 40 python tests/tmarch_worker.py 2>&1 | tail -2 | tee gpurun_out/c18_worker.txt
grep -q "TMARCH OK" gpurun_out/c18_worker.txt || { echo "worker failed or hung: stop"; exit 1; }
LQCD_WILSON_KERNEL=5 LQCD_TM_CHUNKS=2 timeout 40 python tests/tmarch_worker.py 2>&1 | tail -2 | tee gpurun_out/c18_worker_2chunks.txt
for cfg in "LQCD_WILSON_KERNEL=5" "LQCD_WILSON_KERNEL=5 LQCD_TM_PREFETCH=0" "LQCD_WILSON_KERNEL=5 LQCD_TM_CHUNKS=4"; do
  echo -n "32x32x32x32 $cfg: "; env $cfg timeout 30 python tools/quick_bench.py 32x32x32x32 2>&1 | grep -E "wilson" | sed 's/^\([a-z]*\) .*flush=\([01]\) mean *\([0-9.]*\) us.*/\1 flush=\2 \3 us;/' | tr '\n' ' '; echo
done 2>&1 | tee gpurun_out/c18_tm2_prefetch.txt

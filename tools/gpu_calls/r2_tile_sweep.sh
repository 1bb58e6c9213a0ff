#!/bin/bash
# 1 GPU: warps-per-CTA / CTA-tile sweep of the final default Wilson kernel (two-row links, 16 warps per SM) + final sanity (smoke, parity subset)
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/c16_smoke.txt
grep -q "smoke ok" gpurun_out/c16_smoke.txt || { echo "smoke failed: stop"; exit 1; }
for cfg in "A=1" "LQCD_WPC=2" "LQCD_WPC=1" "LQCD_TILE=1,4,1,1" "LQCD_TILE=1,1,4,1" "LQCD_TILE=1,1,2,2" "LQCD_TILE=1,2,1,2" "LQCD_TILE=1,1,1,4" "LQCD_WPC=2 LQCD_TILE=1,2,1,1" "LQCD_WPC=2 LQCD_TILE=1,1,1,2"; do
  echo -n "32x32x32x32 $cfg: "; env $cfg timeout 60 python tools/quick_bench.py 32x32x32x32 2>&1 | grep -E "wilson|stag" | sed 's/^\([a-z]*\) .*flush=\([01]\) mean *\([0-9.]*\) us.*/\1 flush=\2 \3 us;/' | tr '\n' ' '; echo
done 2>&1 | tee gpurun_out/c16_tile_sweep.txt
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2

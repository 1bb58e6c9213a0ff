#!/bin/bash
# debug call: where does the t-marching kernel stop on hardware?  (bounded waits, <= 40 s per probe)
mkdir -p gpurun_out
for lat in 32x4x4x4 8x8x8x8 32x8x8x16; do
  for ch in auto 1; do
    extra=""; [ "$ch" != auto ] && extra="LQCD_TM_CHUNKS=$ch"
    echo "== $lat chunks=$ch"; env $extra LQCD_B200_LIB=latticeqcd.jl_b200/liblqcd_b200_dbg.so timeout 40 python tools/debug/tm_probe.py $lat 2>&1 | tail -4
  done
done 2>&1 | tee gpurun_out/c3_probe.txt

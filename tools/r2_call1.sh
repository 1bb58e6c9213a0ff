#!/bin/bash
# Round-2 call 1 (1 GPU): live-reference probe, strict GPU suite, lost round-1 experiment timings, kernel-3 / occupancy sweeps, ncu.
set -x
mkdir -p gpurun_out
{ echo "julia: $(which julia 2>&1)"; ls -la /root/repo/baseline/_ref 2>&1 | head; ls /root/reference 2>&1 | head -3; nproc; nvidia-smi -L; } > gpurun_out/c1_probe.txt 2>&1
# sweeps first (cheap, most informative)
for lat in 32x32x32x32 32x32x32x16 32x32x32x8 32x32x16x8; do
  for cfg in "A=1" "LQCD_WPC=2 LQCD_LB=64,7" "LQCD_WPC=1 LQCD_LB=64,7" "LQCD_LB=128,4" "LQCD_WILSON_KERNEL=3" "LQCD_WILSON_KERNEL=3 LQCD_K3_CHUNKS=1" "LQCD_WILSON_KERNEL=3 LQCD_K3_CHUNKS=2" "LQCD_WILSON_KERNEL=3 LQCD_K3_CHUNKS=8"; do
    echo -n "$lat $cfg: "; env $cfg timeout 120 python tools/quick_bench.py $lat 2>&1 | grep wilson | tr '\n' '|'; echo
  done
done > gpurun_out/c1_sweep.txt 2>&1
timeout 300 python tools/check_kernel3.py > gpurun_out/c1_kernel3.txt 2>&1
# ncu: default kernel and kernel 3 at 32^4 (2 launches each after warm-up)
for fam in 1 3; do
  LQCD_WILSON_KERNEL=$fam timeout 600 ncu --set full --clock-control none --import-source on -k regex:wilson_dslash --launch-skip 6 -c 2 -f -o gpurun_out/c1_ncu_k$fam python tools/quick_bench.py 32x32x32x32 > gpurun_out/c1_ncu_k$fam.log 2>&1
done
# bench with the experiments leg -> full JSON kept
LQCD_BENCH_EXPERIMENTS_S=300 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
# strict GPU suite (xfail markers ignored)
LQCD_STAGED_BUDGET_S=100000 timeout 1200 python -m pytest tests -m gpu -q -x --runxfail -p no:cacheprovider > gpurun_out/c1_pytest.txt 2>&1
tail -5 gpurun_out/c1_pytest.txt
cat gpurun_out/c1_sweep.txt

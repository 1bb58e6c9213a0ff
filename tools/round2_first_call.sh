#!/bin/bash
# Round-2 opening experiment battery (run under gpurun; 1 GPU unless noted).  Each block is independent.
set -x
mkdir -p gpurun_out
# 1. every GPU test incl. the staged ones as plain tests (a device fault in one must not hide the rest: one process per file)
for f in tests/test_gpu_parity.py tests/test_rhmc.py tests/test_md.py tests/test_gauge_io.py tests/test_reference_regressions.py tests/test_c_example.py tests/test_zz_gpu_unverified.py tests/test_multirank.py; do
  LQCD_STAGED_BUDGET_S=100000 timeout 900 python -m pytest $f -m gpu -q -rA --runxfail -p no:cacheprovider 2>&1 | tail -400 > gpurun_out/tests_$(basename $f .py).txt
done
python tools/staged_report.py gpurun_out/tests_*.txt > gpurun_out/staged_report.txt 2>&1
# 2. headline bench with the experiments leg (first timings of kernel 3, persistent CTAs, multi-RHS R = 2 / 3 / 4, clover, even-odd, MD)
LQCD_BENCH_EXPERIMENTS_S=400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_with_experiments.json 2> gpurun_out/bench_with_experiments.err
python tools/check_kernel3.py 2>&1 | tee gpurun_out/kernel3.txt                       # experimental t-march kernel
for c in 1 2 4 8; do LQCD_WILSON_KERNEL=3 LQCD_K3_CHUNKS=$c python tools/quick_bench.py 32x32x32x32 2>&1 | grep wilson | sed "s/^/k3 chunks=$c /"; done | tee gpurun_out/kernel3_chunks.txt
for lat in 32x32x32x32 32x32x16x8; do for cfg in "A=1" "LQCD_PERSIST=1"; do echo -n "$lat $cfg: "; env $cfg python tools/quick_bench.py $lat 2>&1 | grep wilson; done; done | tee gpurun_out/persist_n1.txt
# on a >= 2 GPU box:
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  ./tools/commprobe | tee gpurun_out/commprobe.txt
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
  for cfg in "A=1" "LQCD_PERSIST=1" "LQCD_WPC=1" "LQCD_WPC=2" "LQCD_PERSIST=1 LQCD_HALO_POLL=relaxed" "LQCD_HALO_POLL=relaxed" "LQCD_PACK_FENCE=g" "LQCD_HALO_POLL=relaxed LQCD_PACK_FENCE=g" "LQCD_SELF_PACK=1" "LQCD_SELF_PACK=0"; do
    for lat in 32x32x32x32 32x32x16x16; do
      echo -n "$lat $cfg: "; env $cfg $TR --master-port 295$((40+RANDOM%50)) bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu --lattice $lat 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1),'us',round(d['cg']['iters_per_s']),'it/s')"
    done
  done | tee gpurun_out/comm_knobs.txt
fi

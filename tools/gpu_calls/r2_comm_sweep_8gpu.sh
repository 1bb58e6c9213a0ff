#!/bin/bash
# call 8 (8-GPU box): halo-pack knob sweep with in-kernel timelines at N = 8, 4, 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29531 tools/comm_tune.py 32x32x32x32 > gpurun_out/c8_tune_n8.txt 2>&1
timeout 200 $TR --nproc-per-node 4 --master-port 29533 tools/comm_tune.py 32x32x32x32 > gpurun_out/c8_tune_n4.txt 2>&1
LQCD_PROCGRID=1,1,2,2 timeout 200 $TR --nproc-per-node 4 --master-port 29535 tools/comm_tune.py 32x32x32x32 > gpurun_out/c8_tune_n4_22.txt 2>&1
timeout 200 $TR --nproc-per-node 2 --master-port 29537 tools/comm_tune.py 32x32x32x32 > gpurun_out/c8_tune_n2.txt 2>&1
for f in n8 n4 n4_22 n2; do echo "== $f"; grep -E "^N=|timeline rank 0" gpurun_out/c8_tune_$f.txt | sed 's/\[lqcd comm timeline rank 0, \(.*\) Dslash, mean of .* applications, us after the first CTA started\]/   tl:/' ; done

#!/bin/bash
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$R --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/sp_k3.log 2>&1; echo "exit $?"; tail -5 gpurun_out/sp_k3.log | cut -c1-400
$R --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/sp_k10.log 2>&1; echo "exit $?"; grep '^{' gpurun_out/sp_k10.log | cut -c1-250
AA_BENCH_NO_GATHER=1 $R --master-port 29520 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/sp_ng.log 2>&1; echo "exit $?"; grep '^{' gpurun_out/sp_ng.log | cut -c1-250

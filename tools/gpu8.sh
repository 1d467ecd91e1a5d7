#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$R --master-port 29531 tools/run_multigpu_cases.py both > gpurun_out/multigpu_cases_$N.log 2>&1; echo "cases exit $?"
grep '^{' gpurun_out/multigpu_cases_$N.log | cut -c1-700
$R --master-port 29532 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1; echo "bench exit $?"
grep '^{' gpurun_out/bench_${N}gpu.log | cut -c1-400
grep -i "error\|Traceback" gpurun_out/multigpu_cases_$N.log gpurun_out/bench_${N}gpu.log | head -5

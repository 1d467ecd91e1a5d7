#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:cond_cluster -c 1 -f -o gpurun_out/cond_cluster_r02 python tools/bench_cond.py --clips 256 --seconds 4 --reps 1 > gpurun_out/ncu_cond.log 2>&1
echo "ncu exit $?"

#!/bin/bash
# ncu --set full of the analysis kernel on the default bench (n = 4096) and on the n = 2048 geometry
mkdir -p gpurun_out
B="--no-e2e --no-cpu --steps 1 --warmup 3"
python bench.py $B > gpurun_out/plain_v17_4096.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:analyze_kernel -s 3 -c 1 -f -o gpurun_out/r02_v17_4096 python bench.py $B > gpurun_out/ncu_v17_4096.log 2>&1; echo "ncu 4096 exit $?"
B2="$B --n 2048 --sr 44100 --seconds 10 --clips 4096"
python bench.py $B2 > gpurun_out/plain_v17_2048.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:analyze_kernel -s 3 -c 1 -f -o gpurun_out/r02_v17_2048 python bench.py $B2 > gpurun_out/ncu_v17_2048.log 2>&1; echo "ncu 2048 exit $?"

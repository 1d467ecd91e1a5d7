#!/bin/bash
# ncu passes (one gpurun call): launch list of a short default bench, then a full capture of the analysis
# kernel on the default bench launch itself (bytes per launch comparable with roofline.achieved).
mkdir -p gpurun_out
A="--no-e2e --no-cpu --steps 2 --warmup 3"
python bench.py $A > gpurun_out/plain_launches.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py $A > gpurun_out/ncu_launches.log 2>&1
echo "launches exit $?"
B="--no-e2e --no-cpu --steps 1 --warmup 3"
python bench.py $B > gpurun_out/plain_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:analyze_kernel -s 3 -c 1 -f -o gpurun_out/prof_final \
    python bench.py $B > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"
tail -2 gpurun_out/ncu_full.log | cut -c1-300

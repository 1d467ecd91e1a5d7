#!/bin/bash
# one ncu --set full capture of the analysis kernel on the default bench launch; usage: gpu_prof_one.sh NAME [bench args]
NAME=${1:-prof}; shift
mkdir -p gpurun_out
B="--no-e2e --no-cpu --steps 1 --warmup 3 $@"
python bench.py $B > gpurun_out/plain_$NAME.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$NAME.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:analyze_kernel -s 3 -c 1 -f -o gpurun_out/$NAME \
    python bench.py $B > gpurun_out/ncu_$NAME.log 2>&1
echo "ncu exit $?"
tail -1 gpurun_out/plain_$NAME.log | cut -c1-400

#!/bin/bash
# round 2, final measurements of the shipped build: ncu (launch list + full capture, both geometries), extra benches
mkdir -p gpurun_out
B="--no-e2e --no-cpu --steps 1 --warmup 3"
python bench.py $B > gpurun_out/plain_v18_4096.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:analyze_kernel -s 3 -c 1 -f -o gpurun_out/r02_v18_4096 python bench.py $B > gpurun_out/ncu_v18_4096.log 2>&1; echo "ncu 4096 exit $?"
B2="$B --n 2048 --sr 44100 --seconds 10 --clips 4096"
python bench.py $B2 > gpurun_out/plain_v18_2048.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:analyze_kernel -s 3 -c 1 -f -o gpurun_out/r02_v18_2048 python bench.py $B2 > gpurun_out/ncu_v18_2048.log 2>&1; echo "ncu 2048 exit $?"
# launch list of the default bench command (device leg + e2e leg), per-launch durations
python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/plain_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_v18.csv python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
python tools/bench_extra.py > gpurun_out/extra_single_gpu_v18.json 2> gpurun_out/extra.err; echo "extra exit $?"
python tools/bench_cond.py > gpurun_out/bench_cond_v18.json 2> gpurun_out/cond.err; echo "cond exit $?"

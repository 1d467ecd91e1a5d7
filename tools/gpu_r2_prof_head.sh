#!/bin/bash
# ncu captures of the HEAD build (after the streaming completion-word changes moved the kernel's layout): full captures
# of the analysis kernel at N = 4096 / 2048 and the launch list of the default bench.  Plain runs first.
mkdir -p gpurun_out
B="--no-e2e --no-cpu --steps 1 --warmup 3"
B2="$B --n 2048 --sr 44100 --seconds 10 --clips 4096"
timeout 300 python bench.py $B > gpurun_out/head_plain_4096.json 2>/dev/null; echo "plain 4096 exit $?"
timeout 300 python bench.py $B2 > gpurun_out/head_plain_2048.json 2>/dev/null; echo "plain 2048 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:analyze_kernel -s 3 -c 1 -f -o gpurun_out/r02_v20_4096 python bench.py $B > gpurun_out/ncu_v20_4096.log 2>&1; echo "ncu 4096 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:analyze_kernel -s 3 -c 1 -f -o gpurun_out/r02_v20_2048 python bench.py $B2 > gpurun_out/ncu_v20_2048.log 2>&1; echo "ncu 2048 exit $?"
timeout 600 python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/head_bench_nocpu.json 2>/dev/null; echo "bench exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_v20.csv python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/ncu_launches_v20.log 2>&1; echo "ncu launches exit $?"
ls -la gpurun_out | tail -12

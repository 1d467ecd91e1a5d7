#!/bin/bash
# round 2, run A: strict parity suite + soak + baseline bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt; lscpu | grep -i "numa\|model name\|socket" >> gpurun_out/gpu.txt
nvidia-smi topo -m >> gpurun_out/gpu.txt 2>&1
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout -s KILL 600 python tools/parity_soak.py --clips 24 --seed 5 > gpurun_out/soak.json 2> gpurun_out/soak.err
echo "soak exit $?"
timeout -s KILL 300 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 > gpurun_out/bench_dev.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; grep -c "parity" gpurun_out/pytest_gpu.log; tail -c 600 gpurun_out/soak.json; tail -1 gpurun_out/bench_dev.log | cut -c1-300

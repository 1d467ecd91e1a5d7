#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --workload cfg3 > gpurun_out/bench_cfg3_new.json 2>/dev/null; python -c "
import sys,json
d=json.loads(open('gpurun_out/bench_cfg3_new.json').read().strip().splitlines()[-1]); print(d['latency_us'], d['latency_us_python_loop'], d['cpu_baseline']['value'])"
bash tools/gpu_r2_ab.sh

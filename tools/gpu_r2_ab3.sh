#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream.py tests/test_gpu_analyze.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --workload cfg3 > gpurun_out/bench_cfg3_new.json 2>/dev/null; python -c "
import sys,json
d=json.loads(open('gpurun_out/bench_cfg3_new.json').read().strip().splitlines()[-1]); print(d['latency_us'], d['latency_us_python_loop'], d['cpu_baseline']['value'])"
B="--no-e2e --no-cpu --steps 5 --warmup 3"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']/1e6,2), 'Mframes/s', round(d['roofline']['kernel_ms'],2), 'ms')"; }
timeout -s KILL 300 python bench.py $B 2>/dev/null | show "default n4096"; timeout -s KILL 300 python bench.py $B --n 2048 --sr 44100 --seconds 10 --clips 4096 2>/dev/null | show "default n2048"
bash tools/gpu_r2_streamprof.sh

#!/bin/bash
bash tools/gpu_r2_ab.sh
timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --workload cfg3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['latency_us'])"

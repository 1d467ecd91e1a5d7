#!/bin/bash
# streaming path: tests, then the cfg3 latency bench (push -> poll per frame)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -x -q > gpurun_out/stream_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/stream_pytest.log
timeout 300 python bench.py --workload cfg3 > gpurun_out/bench_cfg3_new.json 2> gpurun_out/bench_cfg3_new.err; echo "cfg3 exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_cfg3_new.json').read().strip().splitlines()[-1])
print(d['latency_us'], d.get('cpu_baseline',{}).get('value'))
"

#!/bin/bash
# full GPU suite + cfg3 latency + default bench of the current build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/check_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/check_pytest.log
timeout 300 python bench.py --workload cfg3 > gpurun_out/bench_cfg3_new.json 2> gpurun_out/bench_cfg3_new.err; echo "cfg3 exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_cfg3_new.json').read().strip().splitlines()[-1])
print(d['latency_us'], d['latency_us_python_loop'], d.get('cpu_baseline',{}).get('value'))
"
timeout 600 python bench.py --no-cpu > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_check.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['frac_of_h2d_ceiling'])
"

#!/bin/bash
# re-entry check of the restored checkout: full GPU suite (timed), smoke, default bench, reference arm
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/verify_pytest.log 2>&1; echo "pytest exit $? ($(( $(date +%s) - S )) s)"; tail -3 gpurun_out/verify_pytest.log
S=$(date +%s)
timeout 300 python __graft_entry__.py smoke > gpurun_out/verify_smoke.log 2>&1; echo "smoke exit $? ($(( $(date +%s) - S )) s)"; tail -2 gpurun_out/verify_smoke.log
S=$(date +%s)
timeout 600 python bench.py > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err; echo "bench exit $? ($(( $(date +%s) - S )) s)"
python -c "
import json
d=json.loads(open('gpurun_out/verify_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['frac_of_h2d_ceiling'],'frac',d['roofline']['frac'],d['clocks'])
"

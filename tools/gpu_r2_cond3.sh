#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cond.py -m gpu -x -q > gpurun_out/cond_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/cond_pytest.log
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 filters_gate %.2f ms  agc %.2f ms' % (d['filters_gate']['ms'], d['agc']['ms']))"; }
timeout 600 python tools/bench_cond.py --reps 3 > gpurun_out/bench_cond_cluster.json 2> gpurun_out/bench_cond_cluster.err; echo "cluster exit $?"; show cluster < gpurun_out/bench_cond_cluster.json
timeout 600 python tools/bench_cond.py --reps 3 --clips 2048 --seconds 15 2>/dev/null | show cluster_2048x15
bash tools/gpu_r2_condprof.sh
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/cond_launches.csv python tools/bench_cond.py --clips 1024 --seconds 30 --reps 1 > gpurun_out/ncu_cond_launches.log 2>&1; echo "ncu launches exit $?"

#!/bin/bash
# round 2, run C: full GPU suite, cfg4 state modes (1 GPU), the other BASELINE workloads through bench.py
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 900 python tools/cfg4_state_modes.py 3600 4096 > gpurun_out/cfg4_1gpu.json 2> gpurun_out/cfg4_1gpu.err; echo "cfg4 exit $?"; tail -c 400 gpurun_out/cfg4_1gpu.err
for W in cfg1 cfg3 cfg5; do
  timeout -s KILL 900 python bench.py --workload $W > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "$W exit $?"; tail -c 300 gpurun_out/bench_$W.err
done
timeout -s KILL 600 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 exit $?"
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_cfg2_ref.json 2> gpurun_out/bench_cfg2_ref.err; echo "ref exit $?"
timeout -s KILL 600 python tools/parity_soak.py --clips 24 --seed 5 > gpurun_out/soak.json 2> gpurun_out/soak.err; echo "soak exit $?"
for f in gpurun_out/bench_cfg*.json; do echo "== $f"; cut -c1-700 $f; done

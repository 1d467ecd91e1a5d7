#!/bin/bash
# quick perf/correctness iteration: GPU tests, then device-resident bench for build variants.
# usage: gpu_quick.sh ["AA_E_BIG=16 AA_THREADS_PER_SM=512" ...]   (each arg = env for one variant build)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
B="--no-e2e --no-cpu --steps 3 --warmup 3"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']/1e6,2), 'Mframes/s', round(d['roofline']['kernel_ms'],2), 'ms frac', round(d['roofline']['frac'],4))"; }
timeout -s KILL 300 python bench.py $B 2>&1 | tee gpurun_out/bench_quick.log | show "default n4096"
timeout -s KILL 300 python bench.py $B --n 2048 --sr 44100 --seconds 10 --clips 4096 2>&1 | tee gpurun_out/bench_quick_2048.log | show "default n2048"
for V in "$@"; do
  env $V python audio-analyzer-rs_b200/build.py > /dev/null 2>&1
  timeout -s KILL 300 python bench.py $B 2>&1 | show "[$V] n4096"
  timeout -s KILL 300 python bench.py $B --n 2048 --sr 44100 --seconds 10 --clips 4096 2>&1 | show "[$V] n2048"
done

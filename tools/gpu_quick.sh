#!/bin/bash
# quick perf/correctness iteration: GPU tests, then device-resident bench for the given variants
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
B="--no-e2e --no-cpu --steps 3 --warmup 3"
python bench.py $B 2>&1 | tee gpurun_out/bench_quick.log | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default', d['value']/1e6, 'Mframes/s', d['roofline']['kernel_ms'], 'ms frac', d['roofline']['frac'])"
python bench.py $B --n 2048 --sr 44100 --seconds 10 --clips 4096 2>&1 | tee gpurun_out/bench_quick_2048.log | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n2048', d['value']/1e6, 'Mframes/s', d['roofline']['kernel_ms'], 'ms frac', d['roofline']['frac'])"
for V in "$@"; do
  AA_MINB_SCALE=$V python audio-analyzer-rs_b200/build.py > /dev/null 2>&1
  python bench.py $B 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MINB_SCALE=$V', d['value']/1e6, 'Mframes/s', d['roofline']['kernel_ms'], 'ms frac', d['roofline']['frac'])"
done

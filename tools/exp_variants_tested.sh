#!/bin/bash
# like exp_variants.sh, but every variant first runs the analysis / streaming GPU tests (a variant that fails is not benched)
B="--no-e2e --no-cpu --steps 3 --warmup 3"
show() { grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,2), 'Mframes/s')"; }
T="tests/test_gpu_analyze.py tests/test_gpu_stream.py"
timeout -s KILL 300 python -m pytest $T -m gpu -q -x --tb=line -p no:cacheprovider 2>&1 | tail -1
python bench.py $B 2>&1 | show "default n4096"
python bench.py $B --n 2048 --sr 44100 --seconds 10 --clips 4096 2>&1 | show "default n2048"
cp audio-analyzer-rs_b200/libaa_gpu.so /tmp/keep.so
for V in variants/libaa_gpu_*.so; do
  [ -f "$V" ] || continue
  cp $V audio-analyzer-rs_b200/libaa_gpu.so
  timeout -s KILL 300 python -m pytest $T -m gpu -q -x --tb=line -p no:cacheprovider 2>&1 | tail -1
  python bench.py $B 2>&1 | show "$(basename $V) n4096"
  python bench.py $B --n 2048 --sr 44100 --seconds 10 --clips 4096 2>&1 | show "$(basename $V) n2048"
done
cp /tmp/keep.so audio-analyzer-rs_b200/libaa_gpu.so

#!/bin/bash
# quick iteration: GPU analysis tests (stop at first failure) + device-resident bench at N=4096 and N=2048
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_analyze.py tests/test_gpu_fft.py tests/test_gpu_stream.py tests/test_build_properties.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
B="--no-e2e --no-cpu --steps 5 --warmup 3"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']/1e6,2), 'Mframes/s', round(d['roofline']['kernel_ms'],2), 'ms frac', round(d['roofline']['frac'],4))"; }
timeout -s KILL 300 python bench.py $B 2>&1 | tee gpurun_out/bench_quick.log | show "default n4096"
timeout -s KILL 300 python bench.py $B --n 2048 --sr 44100 --seconds 10 --clips 4096 2>&1 | tee gpurun_out/bench_quick_2048.log | show "default n2048"
[ -n "$SMALL_GEOS" ] && { timeout -s KILL 300 python bench.py $B --n 1024 --sr 48000 --seconds 10 --clips 4096 2>&1 | show "default n1024"; timeout -s KILL 300 python bench.py $B --n 256 --sr 48000 --seconds 10 --clips 4096 --features 2 --no-mags 2>&1 | show "default n256-onset"; }
for V in variants/libaa_gpu_*.so; do
  [ -f "$V" ] || continue
  cp audio-analyzer-rs_b200/libaa_gpu.so /tmp/keep.so
  cp $V audio-analyzer-rs_b200/libaa_gpu.so
  [ -n "$TEST_VARIANTS" ] && { timeout -s KILL 600 python -m pytest tests/test_gpu_analyze.py tests/test_gpu_stream.py -m gpu -q --tb=line -x -p no:cacheprovider 2>&1 | tail -2; }
  timeout -s KILL 300 python bench.py $B 2>&1 | show "$(basename $V) n4096"
  timeout -s KILL 300 python bench.py $B --n 2048 --sr 44100 --seconds 10 --clips 4096 2>&1 | show "$(basename $V) n2048"
  [ -n "$SMALL_GEOS" ] && { timeout -s KILL 300 python bench.py $B --n 1024 --sr 48000 --seconds 10 --clips 4096 2>&1 | show "$(basename $V) n1024"; timeout -s KILL 300 python bench.py $B --n 256 --sr 48000 --seconds 10 --clips 4096 --features 2 --no-mags 2>&1 | show "$(basename $V) n256-onset"; }
  cp /tmp/keep.so audio-analyzer-rs_b200/libaa_gpu.so
done

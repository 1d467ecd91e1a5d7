#!/bin/bash
# A/B of the built variants against the in-tree library: device-resident bench at N=4096 and N=2048 (no tests)
mkdir -p gpurun_out
B="--no-e2e --no-cpu --steps 5 --warmup 3"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']/1e6,2), 'Mframes/s', round(d['roofline']['kernel_ms'],2), 'ms')"; }
run() { timeout -s KILL 300 python bench.py $B 2>/dev/null | show "$1 n4096"; timeout -s KILL 300 python bench.py $B --n 2048 --sr 44100 --seconds 10 --clips 4096 2>/dev/null | show "$1 n2048"; }
run default
cp audio-analyzer-rs_b200/libaa_gpu.so /tmp/keep.so
for V in variants/libaa_gpu_*.so; do
  [ -f "$V" ] || continue
  cp $V audio-analyzer-rs_b200/libaa_gpu.so
  run $(basename $V .so)
done
cp /tmp/keep.so audio-analyzer-rs_b200/libaa_gpu.so
run default_again

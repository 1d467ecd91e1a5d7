#!/bin/bash
mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 filters_gate %.2f ms  agc %.2f ms' % (d['filters_gate']['ms'], d['agc']['ms']))"; }
python tools/bench_cond.py --reps 2 2>&1 | show default
cp audio-analyzer-rs_b200/libaa_gpu.so /tmp/keep.so
for V in variants/libaa_gpu_*.so; do
  [ -f "$V" ] || continue
  cp $V audio-analyzer-rs_b200/libaa_gpu.so
  python tools/bench_cond.py --reps 2 2>&1 | show $(basename $V .so)
done
cp /tmp/keep.so audio-analyzer-rs_b200/libaa_gpu.so

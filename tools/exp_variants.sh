#!/bin/bash
# bench the default build, then every prebuilt variants/libaa_gpu_*.so (built here with AA_DEF_* knobs: see build.py,
# AA_SO_OUT); the full log goes to gpurun_out/variants.log
mkdir -p gpurun_out
B="--no-e2e --no-cpu --steps 3 --warmup 3"
show() { grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,2), 'Mframes/s')" 2>&1 | tail -1; }
{
python bench.py $B 2>&1 | show "default n4096"
python bench.py $B --n 2048 --sr 44100 --seconds 10 --clips 4096 2>&1 | show "default n2048"
cp audio-analyzer-rs_b200/libaa_gpu.so /tmp/keep.so
for V in variants/libaa_gpu_*.so; do
  [ -f "$V" ] || continue
  cp $V audio-analyzer-rs_b200/libaa_gpu.so
  python bench.py $B 2>&1 | show "$(basename $V) n4096"
  [ -n "$N2048" ] && python bench.py $B --n 2048 --sr 44100 --seconds 10 --clips 4096 2>&1 | show "$(basename $V) n2048"
done
cp /tmp/keep.so audio-analyzer-rs_b200/libaa_gpu.so
python bench.py $B 2>&1 | show "default n4096 (again)"
} 2>&1 | tee gpurun_out/variants.log

#!/bin/bash
# aa_fft_forward roofline for the default build and every variants/libaa_gpu_*.so
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fft.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
echo "default:"; python tools/bench_fft.py 2>&1 >gpurun_out/fft_default.json | tail -1
cp audio-analyzer-rs_b200/libaa_gpu.so /tmp/keep.so
for V in variants/libaa_gpu_*.so; do
  [ -f "$V" ] || continue
  cp $V audio-analyzer-rs_b200/libaa_gpu.so
  echo "$(basename $V):"; python tools/bench_fft.py 2>&1 >gpurun_out/fft_$(basename $V .so).json | tail -1
done
cp /tmp/keep.so audio-analyzer-rs_b200/libaa_gpu.so

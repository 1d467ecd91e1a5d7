#!/bin/bash
mkdir -p gpurun_out
cp audio-analyzer-rs_b200/libaa_gpu.so /tmp/keep.so
cp variants/libaa_gpu_streamprof.so audio-analyzer-rs_b200/libaa_gpu.so
timeout 300 python tools/stream_timeline.py 1024 > gpurun_out/stream_timeline.log 2>&1; echo "exit $?"
cp /tmp/keep.so audio-analyzer-rs_b200/libaa_gpu.so
cat gpurun_out/stream_timeline.log | tail -14

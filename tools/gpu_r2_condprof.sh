#!/bin/bash
# experiment: per-stage wait / busy cycles of the cluster conditioning kernel (variant built with AA_DEF_COND_PROF=1)
mkdir -p gpurun_out
cp audio-analyzer-rs_b200/libaa_gpu.so /tmp/keep.so
cp variants/libaa_gpu_condprof.so audio-analyzer-rs_b200/libaa_gpu.so
timeout 300 python tools/bench_cond.py --clips 64 --seconds 10 --reps 1 > gpurun_out/condprof.log 2>&1; echo "exit $?"
cp /tmp/keep.so audio-analyzer-rs_b200/libaa_gpu.so
grep "rank" gpurun_out/condprof.log | head -40

#!/bin/bash
# Build experiment variants of libaa_gpu.so next to the in-tree library, here (nvcc cross-compiles without a GPU):
#   tools/build_variants.sh "winpf:AA_DEF_WIN_PREFETCH=1" "noredux:AA_DEF_NO_REDUX=1 AA_DEF_NO_KFB=1"   (AA_DEF_X=v -> -DAA_X=v)
# Each argument is NAME:ENV...; the result is variants/libaa_gpu_NAME.so, which tools/exp_variants.sh benches on
# the GPU box after the default build (gpurun ships variants/ with the snapshot; *.so is git-ignored).
mkdir -p variants
for V in "$@"; do
  NAME="${V%%:*}"; ENVS="${V#*:}"
  env AA_SO_OUT="$PWD/variants/libaa_gpu_$NAME.so" $ENVS python audio-analyzer-rs_b200/build.py --ptxas 2>&1 \
    | grep -A2 "analyze_kernelILi4096ELb1ELb1ELb0ELi2E" | grep -E "spill|error" | sed "s/^/[$NAME] /"
done
ls -la variants

#!/bin/bash
# packed sub-blocks (analyze_kernel SUBS; build the in-tree library with -DAA_SUBS_4096=3 -DAA_SUBS_2048=2 first):
# segment / full-size / state-carry tests, then A/B against AA_NO_PACK=1 in the same build
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_analyze.py -m gpu -q --tb=short -x -p no:cacheprovider -k "packed or segments or full_size or many_clips or state_carry" > gpurun_out/pack_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pack_pytest.log
B="--no-e2e --no-cpu --steps 5 --warmup 3"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']/1e6,2), 'Mframes/s', round(d['roofline']['kernel_ms'],2), 'ms frac', round(d['roofline']['frac'],4))"; }
for rep in 1 2; do
timeout -s KILL 300 python bench.py $B 2>>gpurun_out/pack_bench.err | tee -a gpurun_out/pack_bench.log | show "packed n4096"
AA_NO_PACK=1 timeout -s KILL 300 python bench.py $B 2>>gpurun_out/pack_bench.err | tee -a gpurun_out/pack_bench.log | show "plain  n4096"
timeout -s KILL 300 python bench.py $B --n 2048 --sr 44100 --seconds 10 --clips 4096 2>>gpurun_out/pack_bench.err | tee -a gpurun_out/pack_bench.log | show "packed n2048"
AA_NO_PACK=1 timeout -s KILL 300 python bench.py $B --n 2048 --sr 44100 --seconds 10 --clips 4096 2>>gpurun_out/pack_bench.err | tee -a gpurun_out/pack_bench.log | show "plain  n2048"
done

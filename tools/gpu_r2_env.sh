#!/bin/bash
# speculative envelope chain (AA_ENV_SPECULATE): conditioning tests + bench of the default build, the same tests on the
# build that redoes EVERY chunk through the exact path (variants/libaa_gpu_envredo.so), bench of the exact-only build
mkdir -p gpurun_out
SO=audio-analyzer-rs_b200/libaa_gpu.so
run_tests() { timeout -s KILL 600 python -m pytest tests/test_gpu_cond.py -m gpu -q --tb=short -x -p no:cacheprovider 2>&1 | tail -2; }
bench() { timeout -s KILL 600 python tools/bench_cond.py --reps 3 2>/dev/null | tail -1 | tee -a gpurun_out/env_bench.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', {k: (round(v,2) if isinstance(v,float) else v) for k,v in d.items() if 'ms' in k})"; }
echo "== default (speculative)"; run_tests; bench spec
cp $SO /tmp/keep.so
for V in envredo envexact; do
  [ -f variants/libaa_gpu_$V.so ] || continue
  cp variants/libaa_gpu_$V.so $SO
  echo "== $V"; run_tests; bench $V
done
cp /tmp/keep.so $SO
echo "== default again"; bench spec2

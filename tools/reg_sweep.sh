#!/bin/bash
B="--no-e2e --no-cpu --steps 3 --warmup 3"
show() { grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,2), 'Mframes/s')"; }
for V in "AA_THREADS_PER_SM=960" "AA_THREADS_PER_SM=640" "AA_THREADS_PER_SM=320" "AA_NTAIL=1 AA_THREADS_PER_SM=1152" ; do
  env $V python audio-analyzer-rs_b200/build.py --ptxas 2>&1 | grep -A3 "analyze_kernelILi4096ELb\(0ELb0\|1ELb1\)ELb0" | grep -E "Used" | tr '\n' ' '; echo
  python bench.py $B --features 0 2>&1 | show "[$V] features=0"
  python bench.py $B --features 15 2>&1 | show "[$V] features=15"
done

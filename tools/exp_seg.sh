#!/bin/bash
# segment-length sweep of the batch path (AA_SEG_MIN = shortest time segment in frames, 0 = whole clips)
B="--no-e2e --no-cpu --steps 3 --warmup 3"
show() { grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,2), 'Mframes/s')"; }
for M in ${SEGS:-0 64 32 128}; do
  for C in ${CLIPS:-1024}; do AA_SEG_MIN=$M python bench.py $B --clips $C 2>&1 | show "AA_SEG_MIN=$M clips=$C"; done
done

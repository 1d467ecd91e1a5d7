#!/bin/bash
B="--no-e2e --no-cpu --steps 3 --warmup 3"
for F in 0 4 2 6 1 9 15; do
  python bench.py $B --features $F 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('features=$F n4096 spectra', round(d['value']/1e6,2), 'Mframes/s frac', round(d['roofline']['frac'],4))"
done
python bench.py $B --features 15 --no-mags 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('features=15 n4096 no-mags', round(d['value']/1e6,2), 'Mframes/s frac', round(d['roofline']['frac'],4))"
python bench.py $B --features 0 --n 2048 --sr 44100 --seconds 10 --clips 4096 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('features=0 n2048 spectra', round(d['value']/1e6,2), 'Mframes/s frac', round(d['roofline']['frac'],4))"

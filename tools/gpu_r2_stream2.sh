#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/stream_launches.csv python tools/stream_probe.py 1024 > gpurun_out/stream_probe.log 2>&1; echo "exit $?"
python - <<'PY'
import csv, statistics
rows=[r for r in csv.reader(open('gpurun_out/stream_launches.csv')) if len(r)>5]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
v=[float(r[iv]) for r in rows[1:] if 'analyze' in r[ik]]
print(len(v), "launches; duration ns median", statistics.median(v), "min", min(v), "max", max(v))
PY

#!/bin/bash
# round 2: time-sliced host pipeline -- full GPU suite on the new build, then the headline bench with the slice count varied
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/slice_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/slice_pytest.log
for S in 0 4 8 12; do
  AA_HOST_SLICES=$S timeout 600 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/slice_bench_$S.json 2> gpurun_out/slice_bench_$S.err; echo "slices $S exit $?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/slice_bench_$S.json").read().strip().splitlines()[-1])
print("slices $S value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("frac_of_h2d_ceiling"), "spectra", d.get("e2e_spectra",{}).get("value"))
PY
done
timeout 600 python bench.py > gpurun_out/slice_bench_default.json 2> gpurun_out/slice_bench_default.err; echo "default exit $?"; cat gpurun_out/slice_bench_default.json

#!/bin/bash
T="tests/test_gpu_analyze.py::test_production_kernel_variants_equal_the_tap_build"
python -m pytest "$T" -q -x --tb=line -p no:cacheprovider 2>&1 | tail -4 | cut -c1-200
python - <<'P'
import sys, importlib, numpy as np
sys.path.insert(0, "tests")
import signals
aa = importlib.import_module("audio-analyzer-rs_b200")
n, sr = 2048, 44100.0
clips = np.stack([signals.multitone(300 + i, sr, 24 * n) for i in range(5)]); clips[4] *= 0
cfg = aa.Config(n=n, sample_rate=sr, max_freq=900.0)
tap = aa.Analyzer(cfg).analyze_host(clips, want_dbg=True)
prod = aa.Analyzer(cfg).analyze_host(clips, want_dbg=False)
prod2 = aa.Analyzer(cfg).analyze_host(clips, want_dbg=False)
for k in ("n_pitches", "flux", "energy", "centroid", "burst_count", "max_excess", "flags", "energy_ema"):
    a, b, c = tap["features"][k], prod["features"][k], prod2["features"][k]
    print(k, "tap!=prod:", int((a != b).sum()), "prod!=prod2:", int((b != c).sum()), "of", a.size)
d = tap["features"]["centroid"] != prod["features"]["centroid"]
print("where:", np.argwhere(d)[:10].tolist())
print("tap", tap["features"]["centroid"][d][:5], "prod", prod["features"]["centroid"][d][:5])
print("mags equal:", np.array_equal(tap["mags"], prod["mags"]))
P
cp audio-analyzer-rs_b200/libaa_gpu.so /tmp/keep.so; cp variants/libaa_gpu_nokfb.so audio-analyzer-rs_b200/libaa_gpu.so
python -m pytest "$T" -q -x --tb=line -p no:cacheprovider 2>&1 | tail -3 | cut -c1-200
python bench.py --no-e2e --no-cpu --steps 3 --warmup 3 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nokfb n4096', round(d['value']/1e6,2))"
cp /tmp/keep.so audio-analyzer-rs_b200/libaa_gpu.so

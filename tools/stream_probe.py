#!/usr/bin/env python
"""A short live stream (one hop per push) for profiling the streaming launch: ncu --metrics gpu__time_duration.sum"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import signals
aa = importlib.import_module("audio-analyzer-rs_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
hop = n // 4
x = signals.note_sequence(1, 48000.0, n + hop * 300)
st = aa.Stream(aa.Config(n=n, sample_rate=48000.0))
st.push(x[: n - hop])
pos = n - hop
lat = []
for i in range(300):
    t0 = time.perf_counter()
    st.push(x[pos:pos + hop]); fr = st.poll(4)
    lat.append(time.perf_counter() - t0)
    pos += hop
print("p50 us", 1e6 * float(np.percentile(lat[50:], 50)))

#!/usr/bin/env python
"""Timeline of one streaming launch (library built with AA_DEF_STREAM_PROF=1): %globaltimer stamps the kernel leaves behind
the completion word, averaged over pushes, next to the host-side push / poll times."""
import ctypes as C, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import signals
aa = importlib.import_module("audio-analyzer-rs_b200")
ffi = importlib.import_module("audio-analyzer-rs_b200._ffi")
lib = ffi.lib()
lib.aa_stream_debug_stamps.restype = C.POINTER(C.c_uint64)
lib.aa_stream_debug_stamps.argtypes = [C.c_void_p]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
hop = n // 4
x = signals.note_sequence(1, 48000.0, n + hop * 400)
st = aa.Stream(aa.Config(n=n, sample_rate=48000.0))
st.push(x[: n - hop]); pos = n - hop
rows = []
for i in range(400):
    st.push(x[pos:pos + hop]); fr = st.poll(4); pos += hop
    p = lib.aa_stream_debug_stamps(st._h)
    rows.append([p[4 + k] for k in range(10)])
r = np.array(rows[100:], dtype=np.int64)
names = ["entry", "after init sync", "item fetched", "state loaded", "first window landed", "main frame done", "tail got frame", "tail records written", "before fence", "after fence"]
base = r[:, 0]
for k, nm in enumerate(names):
    d = (r[:, k] - base) / 1e3
    print(f"{nm:24s} +{np.median(d):7.2f} us (p90 {np.percentile(d, 90):.2f})")

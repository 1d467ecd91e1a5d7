#!/usr/bin/env python
"""aa_fft_forward_device roofline per length (CUDA events, batches larger than L2).  Prints one JSON line."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_extra as be
out = {"hbm_peak_GBps": be.PEAK, "fft_forward": []}
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for n, batch in ((4096, 400000), (2048, 800000), (1024, 1600000), (512, 3200000), (256, 6400000)):
    out["fft_forward"].append(be.fft_roofline(n, batch))
print(json.dumps(out))
print(" ".join(f"n{r['n']}:{r['frac_of_measured_hbm']:.3f}" for r in out["fft_forward"]), file=sys.stderr)

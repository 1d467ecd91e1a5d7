#!/usr/bin/env python
"""Extra single-GPU measurements beside bench.py (CUDA events, inputs larger than L2):
  * aa_fft_forward_device (the FftProcessor replacement) roofline for several lengths
  * the fused analysis kernel on the other BASELINE geometries / output modes
Prints one JSON object."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
aa = importlib.import_module("audio-analyzer-rs_b200")
PEAK = 6462.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, iters=5, warm=3):
    s = torch.cuda.current_stream()
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters):
        fn()
    e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def fft_roofline(n, batch):
    x = torch.randn(batch, n, device="cuda")
    out = torch.empty(batch, n // 2 + 1, 2, device="cuda")
    f = aa.FftProcessor(n)
    ms = timed(lambda: f.forward_device(x.data_ptr(), batch, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    nbytes = batch * (4 * n + 8 * (n // 2 + 1))
    return {"n": n, "batch": batch, "ms": ms, "frames_per_s": batch / ms * 1e3, "GBps": nbytes / ms / 1e6,
            "frac_of_measured_hbm": nbytes / ms / 1e6 / PEAK}


def analyze(n, sr, seconds, clips, features, mags):
    clip_len = int(seconds * sr) // 4 * 4
    an = aa.Analyzer(aa.Config(n=n, sample_rate=sr, features=features))
    T = an.num_frames(clip_len)
    half = n // 2 + 1
    x = torch.empty(clips, clip_len, device="cuda")
    aa.synth_clips_device(x.data_ptr(), clips, clip_len, clip_len, sr, 0xA0D15)
    mg = torch.empty(clips * T, half, device="cuda") if mags else None
    ft = torch.empty(clips * T, 96, device="cuda", dtype=torch.uint8)
    stb = torch.empty(clips * T, 136, device="cuda", dtype=torch.uint8) if features & 8 else None
    st = torch.cuda.current_stream().cuda_stream
    ms = timed(lambda: an.analyze_device(x.data_ptr(), clips, clip_len, clip_len, mags=mg.data_ptr() if mags else 0,
                                         features=ft.data_ptr(), stable=stb.data_ptr() if stb is not None else 0,
                                         stream=st))
    bpf = 4 * (n // 4) + (4 * half if mags else 0) + 96 + (136 if features & 8 else 0)
    fps = clips * T / ms * 1e3
    return {"n": n, "sr": sr, "clips": clips, "seconds": seconds, "features": features, "mags": mags, "ms": ms,
            "frames_per_s": fps, "audio_s_per_s": fps * (n // 4) / sr, "bytes_per_frame": bpf,
            "GBps": fps * bpf / 1e9, "frac_of_measured_hbm": fps * bpf / 1e9 / PEAK}


def main():
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    out = {"hbm_peak_GBps": PEAK, "fft_forward": [], "analyze": []}
    for n, batch in ((4096, 400000), (2048, 800000), (1024, 1600000), (256, 6400000)):
        out["fft_forward"].append(fft_roofline(n, batch))
    out["analyze"].append(analyze(2048, 44100.0, 10.0, 4096, 15, True))     # cfg5 geometry, per-GPU share
    out["analyze"].append(analyze(2048, 44100.0, 10.0, 4096, 9, False))     # reference STFT worker: pitch + tracker
    out["analyze"].append(analyze(256, 48000.0, 10.0, 2048, 2, False))      # reference onset geometry 256 / 64
    out["analyze"].append(analyze(1024, 48000.0, 10.0, 4096, 15, True))     # cfg3 geometry, batched
    out["analyze"].append(analyze(4096, 48000.0, 30.0, 1024, 0, True))      # STFT only
    print(json.dumps(out))


if __name__ == "__main__":
    main()

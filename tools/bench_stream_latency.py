#!/usr/bin/env python
"""BASELINE configs[2]: real-time streaming, 256-sample hop, 1024-pt window, one mono stream.
Per-frame latency p50/p99 of push(256 samples) -> poll (H2D + kernel + D2H, wall clock), next to the
CPU oracle's per-frame time on one core.  Prints one JSON line."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import signals  # noqa: E402


def main():
    aa = importlib.import_module("audio-analyzer-rs_b200")
    from oracle import aa_oracle_py as O

    n, hop, sr = 1024, 256, 48000.0
    warm, iters = 1000, 10000
    x = signals.multitone(5, sr, (warm + iters + 8) * hop)
    st = aa.Stream(aa.Config(n=n, sample_rate=sr))
    st.push(x[: n - hop])                       # pre-fill so every later push completes exactly one frame
    lat = np.zeros(warm + iters)
    pos = n - hop
    for i in range(warm + iters):
        chunk = x[pos:pos + hop]
        t0 = time.perf_counter()
        st.push(chunk)
        fr = st.poll(4)
        lat[i] = time.perf_counter() - t0
        assert len(fr) == 1
        pos += hop
    lat = lat[warm:] * 1e6
    # CPU oracle, one core, same stream: mean per-frame time (frame loop of aao_analyze_clip)
    cfg = O.make_config(n, hop, sr)
    t0 = time.perf_counter()
    r = O.analyze_clip(cfg, x, want_mags=False)
    cpu_us = (time.perf_counter() - t0) / r["T"] * 1e6
    print(json.dumps({
        "workload": "single mono stream, 1024-pt window, 256-sample hop @ 48 kHz (5.33 ms of audio per frame)",
        "gpu_push_poll_us": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)),
                             "mean": float(lat.mean()), "max": float(lat.max())},
        "cpu_oracle_us_per_frame_mean": cpu_us, "frames": iters,
        "real_time_budget_us": 1e6 * hop / sr,
    }))


if __name__ == "__main__":
    main()

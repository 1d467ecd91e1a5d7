#!/usr/bin/env python
"""Throughput of the conditioning chain at the cfg2 batch shape, next to the CPU oracle.
Prints one JSON line.  Usage: bench_cond.py [--clips 1024] [--seconds 30] [--sr 48000]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1024)
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--sr", type=float, default=48000.0)
    ap.add_argument("--slot", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch

    aa = importlib.import_module("audio-analyzer-rs_b200")
    n = int(a.seconds * a.sr)
    n -= n % 4
    buf = torch.empty(a.clips * n, dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    res = {}

    def sm_clock():
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
            return {"sm_mhz": pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                    "sm_max_mhz": pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)}
        except Exception:
            return None

    for agc in (False, True):
        cond = aa.Conditioner(a.sr, a.slot, agc=agc)
        n_slots = cond.num_slots(n)
        dyn = torch.empty(a.clips * n_slots * 8, dtype=torch.int32, device="cuda")
        times = []
        for r in range(a.reps + 1):
            aa.synth_clips_device(buf.data_ptr(), a.clips, n, n, a.sr, 0xA0D10, stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            cond.process_device(buf.data_ptr(), a.clips, n, n, dyn.data_ptr() if agc else 0, stream)
            e1.record()
            torch.cuda.synchronize()
            if r:
                times.append(e0.elapsed_time(e1))
            clk = sm_clock()          # right after the timed region, the GPU still at its load clocks
        ms = float(np.median(times))
        res["agc" if agc else "filters_gate"] = {"ms": ms, "samples_per_s": a.clips * n / (ms * 1e-3),
                                                 "audio_s_per_s": a.clips * a.seconds / (ms * 1e-3), "all_ms": times,
                                                 "cycles_per_sample": (ms * 1e-3 * clk["sm_mhz"] * 1e6 / n) if clk else None,
                                                 "clocks": clk}
    # CPU oracle on a bounded sample (one core)
    from oracle import aa_oracle_py as O

    x = buf[: 4 * n].cpu().numpy().reshape(4, n)
    aa.synth_clips_device(buf.data_ptr(), 4, n, n, a.sr, 0xA0D10, stream)
    torch.cuda.synchronize()
    x = buf[: 4 * n].cpu().numpy().reshape(4, n)
    t0 = time.perf_counter()
    for c in x:
        O.condition_clip(c, a.sr, a.slot, agc=True)
    dt = time.perf_counter() - t0
    res["cpu_oracle_1core"] = {"samples_per_s": 4 * n / dt, "audio_s_per_s": 4 * a.seconds / dt, "sample": "4 clips"}
    res["config"] = {"clips": a.clips, "seconds": a.seconds, "sr": a.sr, "slot_len": a.slot}
    print(json.dumps(res))


if __name__ == "__main__":
    main()

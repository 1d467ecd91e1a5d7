#!/usr/bin/env python
"""bench.py -- STFT + pitch/spectral-feature throughput of the frame-analysis hot path.

Default workload (BASELINE.json configs[1], "cfg2"): a batch of 1024 synthetic 30 s clips at 48 kHz per
GPU, 4096-point Hann STFT, hop 1024, all features of the path (adaptive floor + extract_pitches +
PitchTracker, flux / burst / FluxTracker, centroid), magnitudes written ("spectra" mode).
One step = one pass of the hot path over the whole batch.

  python bench.py [--gpus N --steps K --warmup W]          our arm (1 process per GPU under torchrun)
  python bench.py --impl reference ...                      the CPU restatement of the reference
                                                            (oracle/, kind "port": the Rust crate
                                                            cannot be built here), all host threads
  python bench.py --workload {cfg1,cfg2,cfg3,cfg5}          the other BASELINE configs:
      cfg1  one 440 Hz sine, 44.1 kHz mono 10 s, 2048-pt / hop 512 + pitch (stft.rs:169-170): latency-bound on a GPU
      cfg3  one live stream, 1024-pt window, 256-sample hop: push -> poll latency p50 / p99 through aa_stream_*
      cfg5  65 536 clips x 10 s @ 44.1 kHz, 2048-pt / hop 512, features + per-clip summaries, clips sharded over
            the GPUs, summaries all-gathered (NCCL)

Rank 0 prints ONE JSON line.  `value` is device-resident throughput (inputs in HBM), `e2e` is the same
metric through aa_analyze_host with pinned HOST buffers (H2D + kernels + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PKG = "audio-analyzer-rs_b200"
METRIC = "STFT+pitch frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--clips", type=int, default=None, help="clips per GPU (default: the workload's)")
    ap.add_argument("--seconds", type=float, default=None)
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--sr", type=float, default=None)
    ap.add_argument("--features", type=int, default=None)
    ap.add_argument("--no-mags", action="store_true", default=None, help="features-only output mode")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-clips", type=int, default=0, help="clips in the bounded CPU sample (0 = auto)")
    a = ap.parse_args()
    w = WORKLOADS[a.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.clips is None:
        a.clips = w["clips"] if not w.get("total_clips") else max(1, w["total_clips"] // world)
    for k in ("seconds", "n", "sr", "features"):
        if getattr(a, k) is None:
            setattr(a, k, w[k])
    if a.no_mags is None:
        a.no_mags = not w["mags"]
    return a


# BASELINE.json configs; "clips" is per GPU (weak scaling) unless "total_clips" is given (sharded, cfg5)
WORKLOADS = {
    "cfg1": dict(clips=1, seconds=10.0, n=2048, sr=44100.0, features=1 | 8, mags=True, signal="sine440"),
    "cfg2": dict(clips=1024, seconds=30.0, n=4096, sr=48000.0, features=15, mags=True),
    "cfg3": dict(clips=1, seconds=0.0, n=1024, sr=48000.0, features=15, mags=False, stream=True),
    "cfg5": dict(clips=None, total_clips=65536, seconds=10.0, n=2048, sr=44100.0, features=15, mags=False),
}


def workload_name(a):
    return (f"{a.workload}: {a.clips} clips/GPU x {a.seconds:g} s @ {a.sr / 1000:g} kHz, {a.n}-pt Hann STFT hop "
            f"{a.n // 4}, features=0x{a.features:x}" + ("" if a.no_mags else " + magnitudes"))


def config_dict(a, world, frames):
    """The same dict in both arms (the driver compares them)."""
    return {
        "workload": workload_name(a), "frames_per_gpu_per_step": frames,
        "l2": "inputs larger than L2 (126 MB)" if frames * (a.n // 4) * 4 > 126e6 else "inputs fit L2 (latency-bound case)",
        "sharding": f"{world} x {a.clips} independent clips, summaries all-gathered" if world > 1 else "single GPU",
    }


def flops_per_frame(n):
    """Algorithmic FP32 flops per frame: N/2-point complex FFT + split post-pass + magnitudes +
    the per-bin recurrences."""
    half = n // 2 + 1
    return 2.5 * n * np.log2(n / 2) + 10 * (n / 2) + n + 4 * half


def recorded_counts(n, mags, features):
    """Executed warp-instructions per frame of the analysis kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "kernel_counts.json")
    try:
        d = json.load(open(p))
    except Exception:
        return None
    exact = d.get(f"n{n}_{'spectra' if mags else 'features'}_f{features}")
    if exact:
        return exact
    near = d.get(f"n{n}_spectra_f15")      # same window size, other output mode / feature set: an upper estimate
    if near:
        return dict(near, source=near.get("source", "") + " -- captured in spectra mode with every feature on: an upper "
                                                           "estimate for this output mode", approx=True)
    return None


def bytes_per_frame(n, features, mags):
    """Algorithmic HBM bytes per frame: every new input sample once + the outputs once."""
    b = 4 * (n // 4)
    if mags:
        b += 4 * (n // 2 + 1)
    b += 96                       # aa_frame_features
    if features & 8:
        b += 136                  # aa_stable_pitches
    return b


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(n, mags):
    """DRAM bytes per launch of the analysis kernel from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
        return d.get(f"n{n}_{'spectra' if mags else 'features'}")
    except Exception:
        return None


def cpu_baseline(a, sample_clips, threads):
    """Time the oracle (CPU restatement of the reference path) on a bounded sample of the workload."""
    from oracle import aa_oracle_py as O

    cfg = O.make_config(a.n, a.n // 4, a.sr, features=a.features)
    T = O.num_frames(sample_clips.shape[1], a.n, a.n // 4)
    t0 = time.perf_counter()
    O.analyze_batch(cfg, sample_clips, threads, want_mags=not a.no_mags)
    dt = time.perf_counter() - t0
    return sample_clips.shape[0] * T / dt, dt


def clip_len_of(a):
    clip_len = int(a.seconds * a.sr)
    return clip_len - clip_len % 4


def sine440(a, n_clips):
    """cfg1: x[n] = 0.5 sin(2 pi 440 n / sr) evaluated in f64, rounded to f32 (the reference has
    no plain sine generator)."""
    t = np.arange(clip_len_of(a), dtype=np.float64)
    x = (0.5 * np.sin(2.0 * np.pi * 440.0 * t / a.sr)).astype(np.float32)
    return np.ascontiguousarray(np.broadcast_to(x, (n_clips, x.shape[0])))


def host_clips(a, n_clips, first_clip=0):
    """The first n_clips clips of the workload as a host array -- the SAME BITS the device arm analyses: the
    counter-based device generator (SURVEY.md 8d recipe) is run once, untimed, and read back.  Without a GPU
    (never the case on the bench box) a numpy recipe of the same shape stands in and says so."""
    if WORKLOADS[a.workload].get("signal") == "sine440":
        return sine440(a, n_clips), "0.5 sin(2 pi 440 t), f64 -> f32"
    clip_len = clip_len_of(a)
    try:
        import torch

        if torch.cuda.is_available():
            aa = importlib.import_module(PKG)
            out = np.empty((n_clips, clip_len), np.float32)
            step = max(1, min(n_clips, int(2e9 // (4 * clip_len))))
            for c0 in range(0, n_clips, step):
                nc = min(step, n_clips - c0)
                d = torch.empty(nc, clip_len, device="cuda", dtype=torch.float32)
                aa.synth_clips_device(d.data_ptr(), nc, clip_len, clip_len, a.sr, 0xA0D10 + first_clip + c0)
                torch.cuda.synchronize()
                out[c0:c0 + nc] = d.cpu().numpy()
                del d
            return out, "device generator bits (seed 0xA0D10), read back once, untimed"
    except Exception:
        pass
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import signals

    distinct = [signals.multitone(0xA0D10 + c, a.sr, clip_len) for c in range(min(n_clips, 8))]
    return np.stack([distinct[c % len(distinct)] for c in range(n_clips)]), "numpy recipe (no GPU visible), 8 distinct clips tiled"


def stream_latency_cpu(a, x, warm, iters):
    """cfg3 on the CPU: the oracle's frame loop, one core, timed per frame the way the stream is fed (one hop in,
    one frame out)."""
    from oracle import aa_oracle_py as O

    n, hop = a.n, a.n // 4
    cfg = O.make_config(n, hop, a.sr, features=a.features)
    # whole-stream run for the mean (the recurrent state lives inside aao_analyze_clip) ...
    t0 = time.perf_counter()
    r = O.analyze_clip(cfg, x, want_mags=False)
    mean_us = (time.perf_counter() - t0) / r["T"] * 1e6
    # ... and frame-at-a-time calls of the stage functions for the distribution (window + FFT + |X| + floors +
    # extract_pitches + tracker + onset frame), state carried by the oracle's objects
    half = n // 2 + 1
    win = O.hann_window(n)
    pf, trk, on = O.PitchFloor(half), O.Tracker(), O.Onset(half)
    gf = O.global_floor(-96.0, half)
    bw = np.float32(a.sr) / np.float32(n)
    lat = np.zeros(warm + iters)
    for i in range(warm + iters):
        fr = x[i * hop:i * hop + n]
        t0 = time.perf_counter()
        m = O.magnitudes(O.rfft_f32(fr * win))
        eff = pf.update(m, gf)
        pairs, _, _ = O.extract_pitches(m, eff, float(bw))
        trk.process(pairs)
        on.frame(m, gf)
        lat[i] = time.perf_counter() - t0
    lat = lat[warm:] * 1e6
    return mean_us, float(np.percentile(lat, 50)), float(np.percentile(lat, 99))


def stream_signal(a, frames):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import signals

    return signals.multitone(5, a.sr, (frames + 8) * (a.n // 4) + a.n)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    if WORKLOADS[a.workload].get("stream"):
        warm, iters = 200, 2000
        x = stream_signal(a, warm + iters)
        mean_us, p50, p99 = stream_latency_cpu(a, x, warm, iters)
        line = {
            "impl": "reference", "metric": "stream push->poll latency p50", "value": mean_us, "unit": "us",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": mean_us / 1e3,
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a)},
            "latency_us": {"mean_in_frame_loop": mean_us, "p50_ctypes_per_stage": p50, "p99_ctypes_per_stage": p99},
            "cpu_baseline": {"value": mean_us, "unit": "us", "cores": 1, "kind": "port",
                             "sample": f"{iters} frames of one stream through the oracle's frame loop"},
            "e2e": {"value": mean_us, "unit": "us", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return
    clip_len = clip_len_of(a)
    T = (clip_len - a.n) // (a.n // 4) + 1
    # size of the per-step sample: the whole per-GPU batch when a step stays under ~12 s of wall time on this
    # host, else a prefix of it (rate metric); calibrated on one clip per thread, which is also the warm-up
    n_cal = min(a.clips, threads)
    cal, cal_src = host_clips(a, n_cal)
    rate, _ = cpu_baseline(a, cal, threads)
    n_sample = a.cpu_clips or min(a.clips, max(n_cal, int(rate * 12.0 / T) // threads * threads))
    clips, src = (cal, cal_src) if n_sample == n_cal else host_clips(a, n_sample)
    vals = []
    for i in range(a.warmup + a.steps):
        v, dt = cpu_baseline(a, clips, threads)
        if i >= a.warmup:
            vals.append((v, dt))
    total_dt = sum(d for _, d in vals)
    value = a.steps * n_sample * T / total_dt
    whole = n_sample == a.clips
    sample = (f"{'the whole per-GPU batch: ' if whole else 'a prefix of the per-GPU batch: '}{n_sample} clips x "
              f"{a.seconds:g} s per step ({n_sample * T} frames), {threads} threads; {src}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total_dt / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (device generator, seed 0xA0D10)",
        "config": config_dict(a, 1, a.clips * T), "audio_s_per_s": value * (a.n // 4) / a.sr,
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of the reference path (oracle/aa_oracle.c); the Rust crate cannot be built here",
    }
    print(json.dumps(line), flush=True)


def run_stream(a):
    """cfg3: one live stream through aa_stream_push / aa_stream_poll, one hop per push: wall-clock latency of
    push -> poll (memcpy into the mapped sample ring, one kernel launch reading and writing host memory, spin on the
    completion word) per frame, p50 / p99 -- timed from Python and from a C loop inside the library."""
    import torch

    aa = importlib.import_module(PKG)
    n, hop = a.n, a.n // 4
    warm, iters = 1000, 10000
    x = stream_signal(a, warm + iters)
    st = aa.Stream(aa.Config(n=n, sample_rate=a.sr, features=a.features))
    st.push(x[: n - hop])                       # pre-fill so every later push completes exactly one frame
    lat = np.zeros(warm + iters)
    pos = n - hop
    sampler = ClockSampler(0)
    sampler.start()
    for i in range(warm + iters):
        chunk = x[pos:pos + hop]
        t0 = time.perf_counter()
        st.push(chunk)
        fr = st.poll(4)
        lat[i] = time.perf_counter() - t0
        assert len(fr) == 1
        pos += hop
    lat_py = lat[warm:] * 1e6
    # the same loop inside the library (aa_stream_probe_latency): what a native caller of the C ABI -- the reference's
    # Rust worker thread -- pays per frame; the Python loop above adds two ctypes calls and the array handling per frame
    x2 = stream_signal(a, warm + iters)
    st2 = aa.Stream(aa.Config(n=n, sample_rate=a.sr, features=a.features))
    st2.push(x2[: n - hop])
    lat_c, frames_c = st2.probe_latency(x2[n - hop: n - hop + (warm + iters) * hop], hop)
    assert frames_c == warm + iters
    clocks = sampler.stop()
    lat = lat_c[warm:]
    p50, p99 = float(np.percentile(lat, 50)), float(np.percentile(lat, 99))
    line = {
        "metric": "stream push->poll latency p50", "value": p50, "unit": "us", "n_gpus": 1, "steps": iters,
        "warmup": warm, "ms_per_step": float(lat.mean()) / 1e3, "higher_is_better": False, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "real_time_budget_us": 1e6 * hop / a.sr},
        "latency_us": {"p50": p50, "p99": p99, "mean": float(lat.mean()), "max": float(lat.max()),
                       "timed": "inside the library (aa_stream_probe_latency: push + poll per hop in a C loop)"},
        "latency_us_python_loop": {"p50": float(np.percentile(lat_py, 50)), "p99": float(np.percentile(lat_py, 99)),
                                   "mean": float(lat_py.mean())},
        "e2e": {"value": p50, "unit": "us", "h2d_bytes_per_step": 4 * hop, "d2h_bytes_per_step": 96 + 136,
                "note": "the metric IS end to end: host samples in, host records out, per frame"},
        "roofline": {"bound": "latency", "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                     "note": "one frame per launch: launch + PCIe round trip, no bandwidth roofline applies"},
        "clocks": clocks, "gpu_launches": iters,
    }
    if not a.no_cpu:
        mean_us, c50, c99 = stream_latency_cpu(a, x, 200, 2000)
        line["cpu_baseline"] = {"value": mean_us, "unit": "us", "cores": 1, "kind": "port",
                                "sample": "the same stream through the oracle's frame loop on one core (mean per frame); "
                                          f"per-stage ctypes calls p50 {c50:.1f} / p99 {c99:.1f} us",
                                "p50_ctypes_per_stage": c50, "p99_ctypes_per_stage": c99}
    print(json.dumps(line), flush=True)


def run_ours(a):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local)
    if WORKLOADS[a.workload].get("stream"):
        if rank == 0:
            importlib.import_module(PKG).set_device(local)
            run_stream(a)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    aa = importlib.import_module(PKG)
    sh = importlib.import_module(PKG + ".sharding")
    aa.set_device(local)

    n, hop, half = a.n, a.n // 4, a.n // 2 + 1
    clip_len = clip_len_of(a)
    n_clips = a.clips                      # per rank: independent clips
    first_clip, _ = sh.clip_range(n_clips * world, rank, world)
    an = aa.Analyzer(aa.Config(n=n, sample_rate=a.sr, features=a.features))
    T = an.num_frames(clip_len)
    frames = n_clips * T
    mags_on = not a.no_mags
    dev = torch.device("cuda", local)
    clips = torch.empty(n_clips, clip_len, device=dev, dtype=torch.float32)
    if WORKLOADS[a.workload].get("signal") == "sine440":
        clips.copy_(torch.from_numpy(sine440(a, n_clips)))
    else:
        aa.synth_clips_device(clips.data_ptr(), n_clips, clip_len, clip_len, a.sr, 0xA0D10 + first_clip)
    mags = torch.empty(frames, half, device=dev, dtype=torch.float32) if mags_on else None
    feat = torch.empty(frames, 96, device=dev, dtype=torch.uint8)
    stab = torch.empty(frames, 136, device=dev, dtype=torch.uint8) if a.features & 8 else None
    summ = torch.empty(n_clips, 32, device=dev, dtype=torch.uint8)
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)     # the launching stream; the events below are recorded on it
    torch.cuda.set_stream(stream)

    def step(with_summaries=True):
        an.analyze_device(clips.data_ptr(), n_clips, clip_len, clip_len,
                          mags=mags.data_ptr() if mags_on else 0, features=feat.data_ptr(),
                          stable=stab.data_ptr() if stab is not None else 0,
                          summaries=summ.data_ptr() if with_summaries else 0, stream=stream.cuda_stream)
        if world > 1 and with_summaries and not os.environ.get("AA_BENCH_NO_GATHER"):
            return sh.gather_summaries(summ)     # the only collective: 32 B per clip over NCCL
        return summ

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record(stream)
    for _ in range(a.steps):
        step()
        launches += an.last_launches
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)

    # dominant kernel alone (the fused analysis kernel), same stream, same events
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    for _ in range(a.steps):
        step(with_summaries=False)
    k1.record(stream)
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / a.steps
    clocks = sampler.stop()

    # ---- end to end through the host-buffer C ABI call ------------------------------------------
    e2e = None
    if not a.no_e2e:
        # (cfg5 on one GPU holds 115 GB of clips in HBM; the host-buffer leg runs on a bounded prefix)
        ne = min(n_clips, 8192)
        fe = ne * T
        h_clips = aa.pinned_empty((ne, clip_len), np.float32)
        aa.lib().aa_memcpy_d2h(h_clips.ctypes.data, clips.data_ptr(), h_clips.nbytes)
        h_feat = aa.pinned_empty((fe,), aa.FEATURES_DTYPE)
        h_stab = aa.pinned_empty((fe,), aa.STABLE_DTYPE) if a.features & 8 else None
        h_summ = aa.pinned_empty((ne,), aa.SUMMARY_DTYPE)

        def e2e_step(h_mags=None):
            an.analyze_host_into(h_clips, ne, clip_len, clip_len, mags=h_mags, features=h_feat, stable=h_stab,
                                 summaries=h_summ)

        def timed(fn):
            fn()                                 # warm-up: staging allocation
            barrier()
            t0 = time.perf_counter()
            n_launch = 0
            for _ in range(a.steps):
                fn()
                n_launch += an.last_launches
            torch.cuda.synchronize()
            return time.perf_counter() - t0, n_launch

        e2e_s, e2e_launches = timed(e2e_step)
        # supply ceiling of this box at this GPU count: the same pinned buffer through a plain cudaMemcpy on every
        # rank at once (tools/h2d_probe.py is the stand-alone version): e2e cannot be faster than this copy
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            aa.lib().aa_memcpy_h2d(clips.data_ptr(), h_clips.ctypes.data, h_clips.nbytes)
        torch.cuda.synchronize()
        h2d_probe_s = (time.perf_counter() - t0) / 2
        # device result == host-path result
        same = bool((np.frombuffer(feat[:fe].cpu().numpy().tobytes(), np.uint8)
                     == np.frombuffer(h_feat.tobytes(), np.uint8)).all())
        e2e = {"seconds": e2e_s, "launches": e2e_launches, "matches_device_path": same, "clips": ne, "frames": fe,
               "h2d": int(h_clips.nbytes),
               "d2h": int(h_feat.nbytes + (h_stab.nbytes if h_stab is not None else 0) + h_summ.nbytes),
               "spectra_seconds": 0.0, "spectra_d2h": 0, "h2d_probe_s": h2d_probe_s}
        # The same call with the magnitudes copied back as well (what the CPU arm produces): needs a pinned
        # host buffer of frames x (n/2+1) floats per rank, so only when the host has the memory for it
        if mags_on:
            need = fe * half * 4
            try:
                import psutil

                room = psutil.virtual_memory().available
            except Exception:
                room = 0
            if room > 2.0 * world * need + (8 << 30):
                h_mags = aa.pinned_empty((fe, half), np.float32)
                sp_s, _ = timed(lambda: e2e_step(h_mags))
                e2e["spectra_seconds"] = sp_s
                e2e["spectra_d2h"] = e2e["d2h"] + int(h_mags.nbytes)
                del h_mags
        # The same call on 16-bit mono PCM (the reference's input callback takes i16 devices too, mod.rs:691):
        # the samples are the batch quantised to i16, converted on the device (aa_analyze_host_pcm), so the
        # PCIe-bound path moves half the bytes.  Reported beside e2e, not instead of it.
        np.multiply(h_clips, 32767.0, out=h_clips)
        np.rint(h_clips, out=h_clips)
        tmp16 = h_clips.astype(np.int16)         # pageable; the pinned f32 buffer is released before the
        del h_clips                              # pinned i16 buffer is allocated (8 ranks share one host)
        h_pcm = aa.pinned_empty((ne, clip_len), np.int16)
        h_pcm[...] = tmp16
        del tmp16

        def pcm_step():
            an.analyze_host_pcm_into(h_pcm, aa.PCM_I16, 1, ne, clip_len, clip_len, features=h_feat,
                                     stable=h_stab, summaries=h_summ)

        e2e["pcm16_seconds"], _ = timed(pcm_step)
        e2e["pcm16_h2d"] = int(h_pcm.nbytes)

    # ---- reduce over ranks (max time) -----------------------------------------------------------
    times = torch.tensor([ms, kernel_ms, e2e["seconds"] if e2e else 0.0, e2e["pcm16_seconds"] if e2e else 0.0,
                          e2e["spectra_seconds"] if e2e else 0.0, e2e["h2d_probe_s"] if e2e else 0.0],
                         device=dev, dtype=torch.float64)
    ok = torch.tensor([1.0 if (e2e and e2e["spectra_seconds"] > 0) else 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    ms, kernel_ms, e2e_s, pcm_s, sp_s, probe_s = (float(v) for v in times.tolist())

    if rank == 0:
        total_frames = world * frames * a.steps
        value = total_frames / (ms / 1e3)
        bpf = bytes_per_frame(n, a.features, mags_on)
        peak, peak_src = peak_hbm()
        kernel_fps = frames / (kernel_ms / 1e3)
        achieved = kernel_fps * bpf / 1e9
        # the other two roofs of this kernel (report %HBM and %FP32 / issue side by side): FP32 from the
        # algorithmic flops per frame, issue slots from the executed warp-instructions per frame of the committed
        # ncu capture (profiles/kernel_counts.json) at the SM clock sampled during the timed region
        props = torch.cuda.get_device_properties(dev)
        sms = props.multi_processor_count
        clk = (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6
        fp32_frac = kernel_fps * flops_per_frame(n) / (sms * 128 * 2 * clk)
        counts = recorded_counts(n, mags_on, a.features)
        issue_frac = kernel_fps * counts["warp_instr_per_frame"] / (sms * 4 * clk) if counts else None
        fracs = {"hbm": achieved / peak, "fp32": fp32_frac}
        if issue_frac is not None:
            fracs["issue"] = issue_frac
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "strong" if WORKLOADS[a.workload].get("total_clips") else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (device generator, seed 0xA0D10)",
            "config": config_dict(a, world, frames), "audio_s_per_s": value * hop / a.sr,
            "roofline": {
                # achieved / peak / frac are the contract's HBM pair (algorithmic bytes / kernel time against the
                # measured copy bandwidth); `bound` names the roof that is actually closest
                # (fewer clips than SMs: a clip is one CTA walking its frames in time order -- nothing saturates)
                "bound": "latency" if n_clips * world < sms else max(fracs, key=fracs.get),
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "issue_frac": issue_frac, "fp32_frac": fp32_frac,
                "warp_instr_per_frame": counts["warp_instr_per_frame"] if counts else None,
                "counts_source": counts.get("source") if counts else None,
                "flops_per_frame": flops_per_frame(n),
                "traffic": recorded_traffic(n, mags_on), "kernel": "aa::analyze_kernel",
                "kernel_ms": kernel_ms, "bytes_per_frame": bpf, "peak_source": peak_src,
            },
            "clocks": clocks,
            "gpu_launches": launches,
        }
        if e2e:
            ef = world * e2e["frames"] * a.steps
            line["e2e"] = {
                "value": ef / e2e_s, "unit": "frames/s",
                "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                "ms_per_step": 1e3 * e2e_s / a.steps, "launches": e2e["launches"],
                "matches_device_path": e2e["matches_device_path"], "clips_per_gpu": e2e["clips"],
                # the host-to-device supply ceiling measured in this run: every rank copying the same pinned input
                # with a plain cudaMemcpy at the same time (max over ranks); e2e can at best equal it
                "h2d_ceiling_gbs": world * e2e["h2d"] / probe_s / 1e9,
                "h2d_ceiling_frames_per_s": world * e2e["frames"] / probe_s,
                "frac_of_h2d_ceiling": (ef / e2e_s) / (world * e2e["frames"] / probe_s),
                "outputs": "feature records + stable pitches + summaries (magnitudes stay on device; e2e_spectra copies them too)",
            }
            if ok.item() > 0:
                line["e2e_spectra"] = {
                    "value": ef / sp_s, "unit": "frames/s", "h2d_bytes_per_step": e2e["h2d"],
                    "d2h_bytes_per_step": e2e["spectra_d2h"], "ms_per_step": 1e3 * sp_s / a.steps,
                    "outputs": "magnitudes + feature records + stable pitches + summaries (everything the CPU arm produces)",
                }
            line["e2e_pcm16"] = {
                "value": ef / pcm_s, "unit": "frames/s", "h2d_bytes_per_step": e2e["pcm16_h2d"],
                "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": 1e3 * pcm_s / a.steps,
                "input": "the same batch as 16-bit mono PCM, converted on the device (aa_analyze_host_pcm)",
            }
        if world == 1 and not a.no_cpu:
            threads = os.cpu_count() or 1
            n_sample = a.cpu_clips or max(threads, min(4 * threads, 128))
            n_sample = min(n_sample, n_clips)
            sample = clips[:n_sample].cpu().numpy()
            cpu_baseline(a, sample[: min(n_sample, threads)], threads)        # warm-up (page-in, thread start)
            v, dt = cpu_baseline(a, sample, threads)
            line["cpu_baseline"] = {
                "value": v, "unit": "frames/s", "cores": threads, "kind": "port",
                "sample": f"first {n_sample} clips of the batch ({n_sample * T} frames, {dt:.1f} s wall, after one "
                          f"warm-up pass), oracle/aa_oracle.c clip-parallel over {threads} threads",
            }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- STFT + pitch/spectral-feature throughput of the frame-analysis hot path.

Workload (BASELINE.json configs[1]): a batch of 1024 synthetic 30 s clips at 48 kHz per GPU,
4096-point Hann STFT, hop 1024, all features of the path (adaptive floor + extract_pitches +
PitchTracker, flux / burst / FluxTracker, centroid), magnitudes written ("spectra" mode).
One step = one pass of the hot path over the whole batch.

  python bench.py [--gpus N --steps K --warmup W]          our arm (1 process per GPU under torchrun)
  python bench.py --impl reference ...                      the CPU restatement of the reference
                                                            (oracle/, kind "port": the Rust crate
                                                            cannot be built here), all host threads

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
    ap.add_argument("--clips", type=int, default=1024, help="clips per GPU")
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--sr", type=float, default=48000.0)
    ap.add_argument("--features", type=int, default=15)
    ap.add_argument("--no-mags", action="store_true", help="features-only output mode")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-clips", type=int, default=0, help="clips in the bounded CPU sample (0 = auto)")
    return ap.parse_args()


def workload_name(a):
    return (f"{a.clips} clips/GPU x {a.seconds:g} s @ {a.sr / 1000:g} kHz, {a.n}-pt Hann STFT hop {a.n // 4}, "
            f"features=0x{a.features:x}" + ("" if a.no_mags else " + magnitudes"))


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


def numpy_clips(a, n_clips, seed):
    """CPU-side synthetic clips with the same recipe as the device generator (numpy; used only by the
    reference arm, which must not need a GPU)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import signals

    clip_len = int(a.seconds * a.sr)
    distinct = [signals.multitone(seed + c, a.sr, clip_len) for c in range(min(n_clips, 8))]
    return np.stack([distinct[c % len(distinct)] for c in range(n_clips)])


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample = a.cpu_clips or max(threads, min(4 * threads, 128))
    clips = numpy_clips(a, n_sample, 0xA0D10)
    vals = []
    for i in range(a.warmup + a.steps):
        v, dt = cpu_baseline(a, clips, threads)
        if i >= a.warmup:
            vals.append((v, dt))
    T = (clips.shape[1] - a.n) // (a.n // 4) + 1
    total_dt = sum(d for _, d in vals)
    value = a.steps * n_sample * T / total_dt
    sample = f"{n_sample} clips x {a.seconds:g} s per step ({n_sample * T} frames), {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total_dt / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "audio_s_per_s": value * (a.n // 4) / a.sr},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of the reference path (oracle/aa_oracle.c); the Rust crate cannot be built here",
    }
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    aa = importlib.import_module(PKG)
    sh = importlib.import_module(PKG + ".sharding")
    aa.set_device(local)

    n, hop, half = a.n, a.n // 4, a.n // 2 + 1
    clip_len = int(a.seconds * a.sr)
    clip_len -= clip_len % 4
    n_clips = a.clips                      # per rank: weak scaling over independent clips
    first_clip, _ = sh.clip_range(n_clips * world, rank, world)
    an = aa.Analyzer(aa.Config(n=n, sample_rate=a.sr, features=a.features))
    T = an.num_frames(clip_len)
    frames = n_clips * T
    mags_on = not a.no_mags
    dev = torch.device("cuda", local)
    clips = torch.empty(n_clips, clip_len, device=dev, dtype=torch.float32)
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
        h_clips = aa.pinned_empty((n_clips, clip_len), np.float32)
        aa.lib().aa_memcpy_d2h(h_clips.ctypes.data, clips.data_ptr(), h_clips.nbytes)
        h_feat = aa.pinned_empty((frames,), aa.FEATURES_DTYPE)
        h_stab = aa.pinned_empty((frames,), aa.STABLE_DTYPE) if a.features & 8 else None
        h_summ = aa.pinned_empty((n_clips,), aa.SUMMARY_DTYPE)

        def e2e_step():
            an.analyze_host_into(h_clips, n_clips, clip_len, clip_len, features=h_feat, stable=h_stab,
                                 summaries=h_summ)

        e2e_step()                               # warm-up: staging allocation
        barrier()
        t0 = time.perf_counter()
        e2e_launches = 0
        for _ in range(a.steps):
            e2e_step()
            e2e_launches += an.last_launches
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        # device result == host-path result
        same = bool((np.frombuffer(feat.cpu().numpy().tobytes(), np.uint8)
                     == np.frombuffer(h_feat.tobytes(), np.uint8)).all())
        e2e = {"seconds": e2e_s, "launches": e2e_launches, "matches_device_path": same,
               "h2d": int(h_clips.nbytes),
               "d2h": int(h_feat.nbytes + (h_stab.nbytes if h_stab is not None else 0) + h_summ.nbytes)}
        # The same call on 16-bit mono PCM (the reference's input callback takes i16 devices too, mod.rs:691):
        # the samples are the batch quantised to i16, converted on the device (aa_analyze_host_pcm), so the
        # PCIe-bound path moves half the bytes.  Reported beside e2e, not instead of it.
        np.multiply(h_clips, 32767.0, out=h_clips)
        np.rint(h_clips, out=h_clips)
        tmp16 = h_clips.astype(np.int16)         # pageable; the pinned f32 buffer is released before the
        del h_clips                              # pinned i16 buffer is allocated (8 ranks share one host)
        h_pcm = aa.pinned_empty((n_clips, clip_len), np.int16)
        h_pcm[...] = tmp16
        del tmp16

        def pcm_step():
            an.analyze_host_pcm_into(h_pcm, aa.PCM_I16, 1, n_clips, clip_len, clip_len, features=h_feat,
                                     stable=h_stab, summaries=h_summ)

        pcm_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            pcm_step()
        torch.cuda.synchronize()
        e2e["pcm16_seconds"] = time.perf_counter() - t0
        e2e["pcm16_h2d"] = int(h_pcm.nbytes)

    # ---- reduce over ranks (max time) -----------------------------------------------------------
    times = torch.tensor([ms, kernel_ms, e2e["seconds"] if e2e else 0.0, e2e["pcm16_seconds"] if e2e else 0.0],
                         device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, kernel_ms, e2e_s, pcm_s = (float(v) for v in times.tolist())

    if rank == 0:
        total_frames = world * frames * a.steps
        value = total_frames / (ms / 1e3)
        bpf = bytes_per_frame(n, a.features, mags_on)
        peak, peak_src = peak_hbm()
        achieved = frames * bpf / (kernel_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (device generator, seed 0xA0D10)",
            "config": {
                "workload": workload_name(a), "frames_per_gpu_per_step": frames,
                "audio_s_per_s": value * hop / a.sr, "l2": "inputs (5.9 GB/GPU) larger than L2",
                "sharding": f"{world} x {n_clips} independent clips, summaries all-gathered" if world > 1
                else "single GPU",
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": recorded_traffic(n, mags_on), "kernel": "aa::analyze_kernel",
                "kernel_ms": kernel_ms, "bytes_per_frame": bpf, "peak_source": peak_src,
            },
            "clocks": clocks,
            "gpu_launches": launches,
        }
        if e2e:
            line["e2e"] = {
                "value": world * frames * a.steps / e2e_s, "unit": "frames/s",
                "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                "ms_per_step": 1e3 * e2e_s / a.steps, "launches": e2e["launches"],
                "matches_device_path": e2e["matches_device_path"],
                "outputs": "feature records + stable pitches + summaries (magnitudes stay on device)",
            }
            line["e2e_pcm16"] = {
                "value": world * frames * a.steps / pcm_s, "unit": "frames/s", "h2d_bytes_per_step": e2e["pcm16_h2d"],
                "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": 1e3 * pcm_s / a.steps,
                "input": "the same batch as 16-bit mono PCM, converted on the device (aa_analyze_host_pcm)",
            }
        if world == 1 and not a.no_cpu:
            threads = os.cpu_count() or 1
            n_sample = a.cpu_clips or max(threads, min(4 * threads, 128))
            n_sample = min(n_sample, n_clips)
            sample = clips[:n_sample].cpu().numpy()
            v, dt = cpu_baseline(a, sample, threads)
            line["cpu_baseline"] = {
                "value": v, "unit": "frames/s", "cores": threads, "kind": "port",
                "sample": f"first {n_sample} clips of the batch ({n_sample * T} frames, {dt:.1f} s wall), "
                          f"oracle/aa_oracle.c clip-parallel over {threads} threads",
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

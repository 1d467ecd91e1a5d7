"""Deterministic synthetic test signals (numpy only; shared by golden script and tests).

The reference has no plain sine generator, so cfg1's tone is
generated here: x[n] = 0.5*sin(2*pi*440*n/sr) evaluated in float64, rounded to f32.
"""
from __future__ import annotations

import numpy as np


def sine(freq=440.0, sr=44100.0, n=441000, amp=0.5, phase=0.0):
    t = np.arange(n, dtype=np.float64)
    return (amp * np.sin(2.0 * np.pi * freq * t / sr + phase)).astype(np.float32)


def multitone(seed, sr=48000.0, n=96000, noise_db=-60.0):
    """K = 1..4 harmonic tones (6 partials, 1/h amplitudes) + white noise."""
    rng = np.random.default_rng(seed)
    k = int(rng.integers(1, 5))
    t = np.arange(n, dtype=np.float64) / sr
    x = np.zeros(n, np.float64)
    for _ in range(k):
        f0 = float(np.exp(rng.uniform(np.log(55.0), np.log(1760.0))))
        a = float(rng.uniform(0.05, 0.5)) / k
        ph = rng.uniform(0, 2 * np.pi, 6)
        for h in range(1, 7):
            if f0 * h < sr / 2:
                x += a / h * np.sin(2 * np.pi * f0 * h * t + ph[h - 1])
    x += 10.0 ** (noise_db / 20.0) * rng.standard_normal(n)
    return x.astype(np.float32)


def note_sequence(seed, sr=48000.0, n=48000, n_notes=6, noise_db=-70.0):
    """Plucked-note sequence with sharp attacks and exponential decays (onset tests)."""
    rng = np.random.default_rng(seed)
    x = 10.0 ** (noise_db / 20.0) * rng.standard_normal(n)
    starts = np.sort(rng.integers(int(0.05 * n), int(0.9 * n), n_notes))
    t = np.arange(n, dtype=np.float64) / sr
    for s in starts:
        f0 = float(np.exp(rng.uniform(np.log(110.0), np.log(1320.0))))
        a = float(rng.uniform(0.1, 0.6))
        env = np.zeros(n)
        tt = t[s:] - t[s]
        env[s:] = np.minimum(tt / 0.002, 1.0) * np.exp(-tt / 0.15)
        for h in range(1, 5):
            x += a / h * env * np.sin(2 * np.pi * f0 * h * t)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def chord_vibrato(seed, sr=48000.0, n=480000):
    """cfg4-style stream: 5-tone chord with +-20 cent vibrato at 5 Hz."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / sr
    x = np.zeros(n)
    for f0 in (220.0, 277.18, 329.63, 440.0, 659.26):
        dev = f0 * (2 ** (20 / 1200) - 1)
        ph0 = rng.uniform(0, 2 * np.pi)
        # phase = 2*pi*int(f0 + dev*sin(2*pi*5t)) dt
        ph = 2 * np.pi * f0 * t - dev / 5.0 * np.cos(2 * np.pi * 5.0 * t) + ph0
        x += 0.12 * np.sin(ph) + 0.04 * np.sin(2 * ph) + 0.02 * np.sin(3 * ph)
    x += 1e-4 * rng.standard_normal(n)
    return x.astype(np.float32)


def frames_of(x, n, hop):
    T = (len(x) - n) // hop + 1 if len(x) >= n else 0
    idx = np.arange(n)[None, :] + hop * np.arange(T)[:, None]
    return x[idx]

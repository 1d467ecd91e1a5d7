#!/usr/bin/env python
"""Inputs of the golden-vector kit: the synthetic signals of tests/signals.py as float32 .npy files."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import signals  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref")
os.makedirs(OUT, exist_ok=True)
CASES = {
    # name: (sample rate, samples); whole 1024-sample slots only (the input callback never pushes a partial slot,
    # mod.rs:799-803)
    "sine440_44k": (44100.0, signals.sine(440.0, 44100.0, 430 * 1024)),           # BASELINE cfg1 (10 s)
    "multitone_48k": (48000.0, signals.multitone(100, 48000.0, 200 * 1024)),
    "chord_48k": (48000.0, signals.chord_vibrato(0xA0D14, 48000.0, 200 * 1024)),
    "notes_48k": (48000.0, signals.note_sequence(7, 48000.0, 200 * 1024, n_notes=12)),
}
for name, (sr, x) in CASES.items():
    np.save(os.path.join(OUT, f"in_{name}.npy"), x.astype(np.float32))
    with open(os.path.join(OUT, f"in_{name}.sr"), "w") as f:
        f.write(f"{sr}\n")
print("wrote", len(CASES), "inputs to", OUT)

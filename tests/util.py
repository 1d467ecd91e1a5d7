"""Shared comparison helpers for the parity tests."""
from __future__ import annotations

import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAG_TOL = 1e-4          # north_star: max relative error on magnitude, normalised by the frame maximum
NEAR_TIE = 1e-5         # oracle min_margin below this = documented near-tie (aa_oracle.h)


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(path):
    g = np.load(path)
    d = {k: g[k] for k in g.files}
    d["n"], d["hop"] = int(d["n"]), int(d["hop"])
    d["sr"], d["db"] = float(d["sr"]), float(d["db"])
    d["half"] = d["n"] // 2 + 1
    d["peaks"] = np.unpackbits(d["peak_bits"], axis=1)[:, : d["half"]]
    return d


def mag_err(a, ref):
    """per-frame max |a - ref| / max(ref)"""
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    den = np.maximum(ref.max(axis=-1), 1e-30)
    return np.abs(a - ref).max(axis=-1) / den


def ulp_close(a, b, rel=2e-6, abs_=0.0):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) <= rel * np.maximum(np.abs(a), np.abs(b)) + abs_


def compare_pitch_records(feat_a, feat_b, diag=None, rel=2e-5):
    """Frames whose pitch lists differ (count, or freq/score beyond `rel`).
    Returns (bad_frames, near_tie_frames): bad = differing frames that are NOT near-ties."""
    na, nb = feat_a["n_pitches"], feat_b["n_pitches"]
    same_n = na == nb
    fa, fb = feat_a["pitch"]["freq"], feat_b["pitch"]["freq"]
    sa, sb = feat_a["pitch"]["score"], feat_b["pitch"]["score"]
    close = ulp_close(fa, fb, rel).all(axis=-1) & ulp_close(sa, sb, rel).all(axis=-1)
    differ = ~(same_n & close)
    if diag is None:
        return np.nonzero(differ)[0], np.array([], int)
    tie = diag["min_margin"] < NEAR_TIE
    return np.nonzero(differ & ~tie)[0], np.nonzero(differ & tie)[0]


def compare_stable(st_a, st_b, rel=2e-5):
    same_n = st_a["n"] == st_b["n"]
    close = ulp_close(st_a["pitch"]["freq"], st_b["pitch"]["freq"], rel).all(axis=-1) & ulp_close(
        st_a["pitch"]["score"], st_b["pitch"]["score"], rel
    ).all(axis=-1)
    return np.nonzero(~(same_n & close))[0]

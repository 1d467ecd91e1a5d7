#!/usr/bin/env python
"""Parity soak: many random clips through the C ABI, every window size, compared stage-isolated with the CPU
oracle (the same checks as tests/test_gpu_analyze.py: bit-exact floors / peak masks / burst counts / max_excess,
pitch lists equal except oracle-flagged near-ties), END TO END with every difference explained as a
magnitude-level near-tie or downstream of one (tests/parity.py: e2e_* columns, e2e_unexplained must be 0), plus
the conditioning chain.  Prints one JSON line.
Usage: parity_soak.py [--clips 48] [--seed 1]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
import signals  # noqa: E402
import util  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=48)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    aa = importlib.import_module("audio-analyzer-rs_b200")
    from oracle import aa_oracle_py as O

    rng = np.random.default_rng(a.seed)
    out = {"cases": [], "seed": a.seed}
    t0 = time.time()
    for n, sr in [(256, 48000.0), (512, 22050.0), (1024, 48000.0), (2048, 44100.0), (4096, 48000.0), (4096, 96000.0)]:
        length = 40 * n
        clips = []
        for c in range(a.clips):
            s = int(rng.integers(1 << 30))
            x = signals.multitone(s, sr, length, noise_db=float(rng.choice([-80.0, -60.0, -40.0]))) if c % 3 else \
                signals.note_sequence(s, sr, length, n_notes=6)
            clips.append(x * np.float32(rng.uniform(0.05, 1.0)))
        clips = np.stack(clips).astype(np.float32)
        db = float(rng.choice([-96.0, -70.0, -50.0]))
        an = aa.Analyzer(aa.Config(n=n, sample_rate=sr, noise_floor_db=db))
        tap = an.analyze_host(clips, want_dbg=True)
        prod = an.analyze_host(clips, want_dbg=False)
        case = {"n": n, "sr": sr, "db": db, "clips": a.clips, "frames": int(a.clips * tap["T"]), "floor_mismatch": 0,
                "peak_mismatch": 0, "burst_mismatch": 0, "maxex_mismatch": 0, "pitch_hard_mismatch": 0, "near_tie_frames": 0,
                "flag_mismatch": 0, "prod_vs_tap_mismatch": 0, "max_mag_err": 0.0,
                # end to end (signal -> oracle vs signal -> GPU), every difference classified by tests/parity.py
                "e2e_peak_bits_differ": 0, "e2e_burst_bits_differ": 0, "e2e_pitch_lists_differ": 0, "e2e_unexplained": 0,
                "e2e_floor_taint_starts": 0, "e2e_peak_tie_frac_max": 0.0}
        for k in ("features", "stable", "mags"):
            case["prod_vs_tap_mismatch"] += int(tap[k].tobytes() != prod[k].tobytes())
        cfg = O.make_config(n, n // 4, sr, noise_floor_db=db)
        for c in range(a.clips):
            kw = dict(want_floor=True, want_peaks=True, want_diag=True, want_state=True)
            iso = O.analyze_clip(cfg, mags_in=tap["mags"][c], **kw)
            e2e = O.analyze_clip(cfg, clips[c], want_mags=True, **kw)
            rep = parity.explain(O, cfg, e2e, iso)
            case["e2e_peak_bits_differ"] += rep["peak_bits_differ"]
            case["e2e_burst_bits_differ"] += rep["burst_bits_differ"]
            case["e2e_pitch_lists_differ"] += rep["pitch_lists_differ"]
            case["e2e_unexplained"] += rep["unexplained"]
            case["e2e_floor_taint_starts"] += rep["pitch_floor_taint_starts"] + rep["onset_floor_taint_starts"]
            case["e2e_peak_tie_frac_max"] = max(case["e2e_peak_tie_frac_max"], rep["peak_tie_frac"])
            case["max_mag_err"] = max(case["max_mag_err"], float(util.mag_err(tap["mags"][c], e2e["mags"]).max()))
            g, o = tap["features"][c], iso["features"]
            case["floor_mismatch"] += int(not np.array_equal(tap["dbg_floor"][c], iso["floor"]))
            case["peak_mismatch"] += int(not np.array_equal(tap["dbg_peaks"][c], iso["peaks"]))
            case["burst_mismatch"] += int((g["burst_count"] != o["burst_count"]).sum())
            case["maxex_mismatch"] += int((g["max_excess"] != o["max_excess"]).sum())
            bad, ties = util.compare_pitch_records(g, o, iso["diag"])
            case["pitch_hard_mismatch"] += len(bad)
            case["near_tie_frames"] += len(ties)
            case["flag_mismatch"] += int((g["flags"] != o["flags"]).sum())
        out["cases"].append(case)
    # conditioning chain
    sr, L = 48000.0, 1024
    clips = np.stack([signals.note_sequence(int(rng.integers(1 << 30)), sr, 200 * L, n_notes=8, noise_db=-90.0)
                      for _ in range(a.clips)]).astype(np.float32)
    got, dyn = aa.Conditioner(sr, L, agc=False).process_host(clips)
    ref = np.stack([O.condition_clip(x, sr, L, agc=False)[0] for x in clips])
    full, fdyn = aa.Conditioner(sr, L, agc=True).process_host(clips)
    rd = np.stack([O.condition_clip(x, sr, L, agc=True)[1] for x in clips])
    out["conditioning"] = {"clips": a.clips, "slots": int(fdyn.size),
                           "filter_gate_bit_exact": bool(np.array_equal(got.view(np.uint32), ref.view(np.uint32))),
                           "dynamics_flag_mismatch": int((fdyn["flags"] != rd["flags"]).sum()),
                           "level_mismatch": int((fdyn["level"] != rd["level"]).sum()),
                           "max_gain_rel_err": float(np.max(np.abs(fdyn["effective_gain"] - rd["effective_gain"])
                                                            / rd["effective_gain"]))}
    out["seconds"] = time.time() - t0
    out["hard_failures"] = int(sum(c["floor_mismatch"] + c["peak_mismatch"] + c["burst_mismatch"] + c["maxex_mismatch"]
                                   + c["pitch_hard_mismatch"] + c["prod_vs_tap_mismatch"] + c["e2e_unexplained"] for c in out["cases"])
                               + (0 if out["conditioning"]["filter_gate_bit_exact"] else 1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()

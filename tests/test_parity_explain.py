"""CPU test of the end-to-end near-tie classifier (tests/parity.py).  The "other implementation" here is a
float64 pocketfft spectrum rounded to f32 (it differs from the oracle's f32 Stockham by ~2e-7 of the frame
maximum, like the GPU does): every discrete difference between the two runs must be explained, and a
difference that is NOT a near-tie must be caught."""
import numpy as np
import pytest

import parity
import signals

KW = dict(want_floor=True, want_peaks=True, want_diag=True, want_state=True)


def two_runs(O, x, n, sr):
    cfg = O.make_config(n, n // 4, sr)
    A = O.analyze_clip(cfg, x, **KW)
    B = O.analyze_clip(cfg, mags_in=parity.f64_magnitudes(x, n, n // 4, O.hann_window(n)), **KW)
    return cfg, A, B


@pytest.mark.parametrize("n,sr", [(256, 48000.0), (1024, 48000.0), (2048, 44100.0), (4096, 48000.0)])
def test_every_end_to_end_difference_is_a_near_tie(O, n, sr):
    total_diff = 0
    for x in (signals.multitone(100, sr, 60 * n), signals.multitone(102, sr, 60 * n), signals.note_sequence(7, sr, 60 * n)):
        cfg, A, B = two_runs(O, x, n, sr)
        rep = parity.explain(O, cfg, A, B, f"n={n}")
        print(parity.summarize(rep))
        assert rep["unexplained"] == 0, rep["detail"]
        assert rep["d_rel_max"] < 1e-6
        # the classifier is not vacuous: near-ties are a small minority of the decisions on ordinary signals
        assert rep["peak_tie_frac"] < 0.02 and rep["burst_tie_frac"] < 0.01
        total_diff += rep["peak_bits_differ"] + rep["burst_bits_differ"] + rep["pitch_lists_differ"]
    print("differences explained:", total_diff)


def test_pure_sine_regime_is_flagged_not_hidden(O):
    """cfg1: on a pure digital sine the leakage bins decay to the FFT's own rounding noise;
    there most peak decisions ARE near-ties, thousands of peak bits differ between two correct FFTs, every one of
    them is explained per bin -- and the pitch bin lists still agree on every frame."""
    cfg, A, B = two_runs(O, signals.sine(440.0, 44100.0, 441000), 2048, 44100.0)
    rep = parity.explain(O, cfg, A, B, "cfg1")
    print(parity.summarize(rep))
    assert rep["unexplained"] == 0
    assert rep["peak_bits_differ"] > 1000 and rep["peak_tie_frac"] > 0.2
    assert rep["pitch_lists_differ"] == 0


def test_a_real_mismatch_is_not_explained(O):
    """Negative control: decisions that differ although their operands are far apart must be reported."""
    n, sr = 2048, 44100.0
    cfg, A, B = two_runs(O, signals.multitone(101, sr, 60 * n), n, sr)
    assert parity.explain(O, cfg, A, B)["unexplained"] == 0
    # (1) a peak bit flipped on a strong, isolated peak
    t = 30
    k = int(np.argmax(A["mags"][t][30:900])) + 30
    assert A["peaks"][t, k] == 1
    B2 = dict(B, peaks=B["peaks"].copy())
    B2["peaks"][t, k] = 0
    rep = parity.explain(O, cfg, A, B2)
    assert rep["peak_bits_unexplained"] == 1 and rep["detail"][0] == ("peak", t, k)
    # (2) a different pitch-bin list on a frame without any near-tie
    d = B["diag"].copy()
    dmax = np.abs(A["mags"] - B["mags"]).max()
    quiet = np.nonzero((np.minimum(A["diag"]["cand_eps"], d["cand_eps"]) > 5 * parity.SLACK * dmax) & (d["n_out"] > 0)
                       & ~(A["peaks"] != B["peaks"]).any(axis=1))[0]
    assert len(quiet) > 10
    d["out_bins"][quiet[3], 0] += 1
    rep = parity.explain(O, cfg, A, dict(B, diag=d))
    assert rep["pitch_lists_unexplained"] == 1
    # (3) a burst decision taken with a wrong threshold (2.0 instead of 2.5) shifts the onset floor of that bin
    x = signals.note_sequence(9, 48000.0, 60 * 1024)
    cfg, A, B = two_runs(O, x, 1024, 48000.0)
    on = B["onset_nf"].copy()
    t = int(np.argmax(A["features"]["burst_count"]))
    k = int(np.argmax(A["mags"][t]))
    on[t:t + 20, k] *= 0.5
    rep = parity.explain(O, cfg, A, dict(B, onset_nf=on))
    assert rep["unexplained"] > 0

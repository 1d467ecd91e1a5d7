"""End-to-end index parity: prove that EVERY discrete difference between two runs of the frame-analysis
path that consumed slightly different magnitudes is a magnitude-level near-tie, or lies downstream of one
through the reference's recurrent per-bin state.

north_star: "pitch lag / bin indices bit-exact except for documented near-ties".  Two correct FFT
implementations (rustfft's SIMD kernels, the oracle's Stockham, the GPU's packed Stockham) differ by
~1e-7 of the frame maximum, so signal -> GPU and signal -> oracle cannot agree on a `>` whose operands are
that close.  Instead of bounding the mismatch *rate*, this module explains each mismatch:

  run A   oracle(signal)                 -- magnitudes mA
  run B   oracle(magnitudes of the other implementation, mB)   (for the GPU: B == the GPU's own records,
          which the stage-isolated tests prove bit-exact)

  d_t     = max_k |mA - mB| at frame t          (measured, not assumed)
  D_t     = max_{t' <= t} d_t'                  (what the recurrent state may have accumulated)
  a decision  lhs > rhs  taken at (t, k) is a NEAR-TIE iff  |lhs - rhs| <= SLACK * D_t * sens,
  where sens = L1 norm of d(lhs - rhs)/d(magnitudes, state values) -- the smallest uniform perturbation
  that flips it, to first order, is |lhs - rhs| / sens.

Decisions, per bin (reference lines):
  peak pick                stft.rs:465      m > floor && m >= m[k-1] && m >= m[k+1]
  pitch-floor branch       stft.rs:351      above_ratio > 1.5 && vol_norm < 0.15   (frozen vs updated floor)
  onset burst              onset.rs:318     m / max(floor, eps) > 2.5              (count + floor := 1.3 m)
  (stft.rs:355 `mag > floor` and onset.rs:323 `mag > floor` choose between two updates that both vanish at
   the tie -- alpha * (m - floor) -> 0 -- so a flip there cannot move the state by more than the gap itself.)
and per frame the candidate-level decisions of extract_pitches (oracle: aao_pitch_diag.cand_eps, classes
104-124: stft.rs:479, 492, 506, 509-510, 516, 536, 547, 556, 575-580, 592, 600, 613).

A difference is EXPLAINED iff
  * peak bit (t, k): the peak decision at (t, k) is a near-tie, or the pitch floor of bin k is tainted;
  * burst bit (t, k): the burst decision at (t, k) is a near-tie, or the onset floor of bin k is tainted;
  * a bin's floor becomes TAINTED (state differs by more than SLACK_STATE * D_t) only at a frame where the
    branch that feeds it is a near-tie at that bin (or, for the onset floor, where the burst bit of that
    bin already differs for an explained reason); it stays tainted until the two runs reconverge;
  * pitch list (t): cand_eps[t] <= SLACK * D_t, or the frame's peak masks differ (each bit explained
    above), or a peak / candidate bin of the frame has a tainted pitch floor.
Everything else is UNEXPLAINED and the tests assert there is none.
"""
from __future__ import annotations

import numpy as np

SLACK = 2.0          # linearisation head-room on the measured magnitude difference
SLACK_STATE = 8.0    # a floor value is "materially different" beyond this many D_t (EMAs of magnitudes carry
                     # at most ~1.3 D_t when no branch flipped; the volatility-dependent alpha adds a few)
ULPS = 8.0 * 2.0 ** -24   # the f32 recurrences round their own state: a few ulp of the operands are never material


def _f32(x):
    return np.asarray(x, np.float32)


def explain(O, cfg, A, B, label=""):
    """A, B: dicts from O.analyze_clip(..., want_floor, want_peaks, want_diag, want_state) on the same clip.
    Returns a dict of counts; `unexplained` must be 0."""
    n, half = cfg.n, cfg.n // 2 + 1
    T = A["T"]
    assert B["T"] == T
    feats = int(cfg.features)
    bw = np.float32(cfg.sample_rate) / np.float32(n)
    gf = np.float32(O.global_floor(cfg.noise_floor_db, half))
    mA, mB = A["mags"].astype(np.float64), B["mags"].astype(np.float64)
    d_t = np.abs(mA - mB).max(axis=1)
    D_t = np.maximum.accumulate(d_t)
    thr = (SLACK * D_t)[:, None]                      # [T,1]
    thr_state = (SLACK_STATE * D_t)[:, None]
    out = {"T": T, "d_max": float(D_t[-1]), "d_rel_max": float((d_t / np.maximum(mA.max(axis=1), 1e-30)).max()),
           "unexplained": 0, "detail": []}

    # ------------------------------------------------------------------ pitch side
    if feats & 1:
        min_bin = max(int(np.ceil(np.float32(cfg.min_freq) / bw)), 1)
        max_bin = min(int(np.floor(np.float32(cfg.max_freq) / bw)), half - 2)
        ks = np.arange(min_bin + 1, max_bin)          # bins the peak pick looks at (stft.rs:463)
        in_range = np.zeros(half, bool)
        in_range[ks] = True
        eff = A["floor"].astype(np.float64)
        nfA, nfB = A["pitch_nf"].astype(np.float64), B["pitch_nf"].astype(np.float64)
        fs = (nfA <= 2.5 * float(gf)).astype(np.float64)      # clamped floor = constant
        m = mA
        mL = np.concatenate([m[:, :1], m[:, :-1]], axis=1)
        mR = np.concatenate([m[:, 1:], m[:, -1:]], axis=1)
        g1 = (m - eff) / (1.0 + fs)
        g2 = (m - mL) / 2.0
        g3 = (m - mR) / 2.0
        thr1 = thr + ULPS * np.maximum(m, eff)
        rob_true = (g1 > thr1) & (g2 > thr) & (g3 > thr)
        rob_false = (g1 < -thr1) | (g2 < -thr) | (g3 < -thr)
        peak_tie = ~(rob_true | rob_false) & in_range[None, :]

        # pitch-floor branch (stft.rs:351): fl = floor before the update, vol after it
        fl = np.concatenate([nfA[:1], nfA[:-1]], axis=0)
        vol = A["pitch_vol"].astype(np.float64)
        a_gap = (m - 1.5 * np.maximum(fl, 0.01)) / (1.0 + 1.5 * (fl > 0.01))
        b_gap = (vol - 0.15 * np.maximum(m, 0.05)) / (2.0 + 0.15 * (m > 0.05))
        thra = thr + ULPS * np.maximum(m, fl)
        thrb = thr + ULPS * np.maximum(vol, m)
        sus_true = (a_gap > thra) & (b_gap < -thrb)
        sus_false = (a_gap < -thra) | (b_gap > thrb)
        sus_tie = ~(sus_true | sus_false)
        sus_tie[0] = False                                # first frame: plain initialisation (:326-331)

        taintP = np.abs(nfA - nfB) > thr_state + ULPS * np.maximum(nfA, nfB)
        startP = taintP.copy()
        startP[1:] &= ~taintP[:-1]
        bad_start = startP & ~sus_tie
        bad_start[:, ~in_range] = False                   # floors outside the peak range are read nowhere
        peak_diff = (A["peaks"] != B["peaks"])
        bad_peak = peak_diff & ~(peak_tie | taintP)
        # pitch lists, by integer bins (the index-parity target)
        da, db = A["diag"], B["diag"]
        list_diff = (da["n_out"] != db["n_out"]) | (da["out_bins"] != db["out_bins"]).any(axis=1)
        cand_tie = np.minimum(da["cand_eps"], db["cand_eps"]).astype(np.float64) <= SLACK * D_t
        peaks_differ = peak_diff.any(axis=1)
        taint_at_peak = (taintP & ((A["peaks"] != 0) | (B["peaks"] != 0))).any(axis=1)
        bad_list = list_diff & ~(cand_tie | peaks_differ | taint_at_peak)
        out.update(peak_bits_differ=int(peak_diff.sum()), peak_bits_unexplained=int(bad_peak.sum()),
                   peak_tie_frac=float(peak_tie[:, ks].mean()) if len(ks) else 0.0,
                   pitch_lists_differ=int(list_diff.sum()), pitch_lists_unexplained=int(bad_list.sum()),
                   cand_tie_frames=int(cand_tie.sum()),
                   pitch_floor_taint_starts=int((startP & in_range[None, :]).sum()),
                   pitch_floor_taint_unexplained=int(bad_start.sum()),
                   pitch_floor_tainted_frac=float(taintP[:, ks].mean()) if len(ks) else 0.0)
        out["unexplained"] += int(bad_peak.sum()) + int(bad_list.sum()) + int(bad_start.sum())
        for name, arr in (("peak", bad_peak), ("floor-taint", bad_start)):
            for t, k in zip(*np.nonzero(arr)):
                out["detail"].append((name, int(t), int(k)))
        for t in np.nonzero(bad_list)[0]:
            out["detail"].append(("pitch-list", int(t), int(da["cand_src"][t])))

    # ------------------------------------------------------------------ onset side
    if feats & 2:
        eps_floor = max(float(gf), 0.01)                                       # onset.rs:302
        onA, onB = A["onset_nf"].astype(np.float64), B["onset_nf"].astype(np.float64)

        def pre_floor(on, mags):
            first = np.maximum(mags[:1], float(gf))                            # onset.rs:304-309
            return np.concatenate([first, on[:-1]], axis=0)

        fA, fB = pre_floor(onA, mA), pre_floor(onB, mB)

        def burst_bits(mags32, pre):
            fk = np.maximum(_f32(pre), np.float32(eps_floor))
            return (_f32(mags32) / fk) > np.float32(2.5)                       # onset.rs:316-318, f32 like the oracle

        bA, bB = burst_bits(A["mags"], fA), burst_bits(B["mags"], fB)
        # the burst bits recomputed from each run's own floor taps must reproduce that run's burst counts: a run
        # whose counts do not follow from its state is wrong, whatever the margins say
        for run, bits in ((A, bA), (B, bB)):
            off = np.nonzero(bits.sum(axis=1) != run["features"]["burst_count"])[0]
            out["unexplained"] += len(off)
            out["detail"] += [("burst-count-vs-state", int(t), -1) for t in off[:5]]
        gap = (mA - 2.5 * np.maximum(fA, eps_floor)) / (1.0 + 2.5 * 1.3 * (fA > eps_floor))
        burst_tie = np.abs(gap) <= thr + ULPS * np.maximum(mA, 2.5 * fA)
        taintO_post = np.abs(onA - onB) > thr_state * 1.3 + ULPS * np.maximum(onA, onB)    # state after frame t
        taintO_pre = np.concatenate([np.zeros_like(taintO_post[:1]), taintO_post[:-1]], axis=0)
        burst_diff = bA != bB
        bad_burst = burst_diff & ~(burst_tie | taintO_pre)
        startO = taintO_post & ~taintO_pre
        bad_startO = startO & ~(burst_tie | burst_diff)
        out.update(burst_bits_differ=int(burst_diff.sum()), burst_bits_unexplained=int(bad_burst.sum()),
                   burst_count_frames_differ=int((A["features"]["burst_count"] != B["features"]["burst_count"]).sum()),
                   burst_tie_frac=float(burst_tie.mean()),
                   onset_floor_taint_starts=int(startO.sum()), onset_floor_taint_unexplained=int(bad_startO.sum()),
                   onset_floor_tainted_frac=float(taintO_post.mean()))
        out["unexplained"] += int(bad_burst.sum()) + int(bad_startO.sum())
        for name, arr in (("burst", bad_burst), ("onset-floor-taint", bad_startO)):
            for t, k in zip(*np.nonzero(arr)):
                out["detail"].append((name, int(t), int(k)))
    out["detail"] = out["detail"][:20]
    out["label"] = label
    return out


def summarize(rep):
    keys = ("T", "d_rel_max", "peak_bits_differ", "peak_bits_unexplained", "pitch_lists_differ", "pitch_lists_unexplained",
            "cand_tie_frames", "peak_tie_frac", "pitch_floor_taint_starts", "pitch_floor_taint_unexplained",
            "burst_bits_differ", "burst_bits_unexplained", "burst_count_frames_differ", "burst_tie_frac",
            "onset_floor_taint_starts", "onset_floor_taint_unexplained", "unexplained")
    return f"[parity {rep.get('label', '')}] " + " ".join(
        f"{k}={rep[k]:.3g}" if isinstance(rep.get(k), float) else f"{k}={rep[k]}" for k in keys if k in rep)


def f64_magnitudes(x, n, hop, window):
    """|rfft| of the windowed frames in float64, rounded to f32: an independent 'other implementation'."""
    T = (len(x) - n) // hop + 1
    idx = np.arange(n)[None, :] + hop * np.arange(T)[:, None]
    fr = (np.asarray(x, np.float32)[idx] * np.asarray(window, np.float32)[None, :]).astype(np.float64)
    return np.abs(np.fft.rfft(fr, axis=1)).astype(np.float32)

"""CPU tests of the conditioning-chain oracle (oracle/aa_oracle_cond.c).

The reference has no test on this code (parity unpinned).  The C restatement is checked against an
independent restatement written here with numpy float32 scalars (every operation rounded to f32, in
the order of src/audio_io/mod.rs:433-487 and src/audio_io/dynamics.rs:194-360) and against
known-answer properties of the chain."""
import math

import numpy as np
import pytest

f32 = np.float32


def np_params(sr, slot_len):
    """mod.rs:357-418 and dynamics.rs:164-189 in float32."""
    sr = f32(sr)
    pi = f32(math.pi)

    def biquad(freq, lpf):
        w0 = f32(2.0) * pi * f32(freq) / sr
        c, s = f32(math.cos(float(w0))), f32(math.sin(float(w0)))   # libm double -> f32 (<= 1 ulp from cosf)
        alpha = s / (f32(2.0) * f32(0.707))
        if lpf:
            b0, b1, b2 = (f32(1) - c) / f32(2), f32(1) - c, (f32(1) - c) / f32(2)
        else:
            b0, b1, b2 = (f32(1) + c) / f32(2), -(f32(1) + c), (f32(1) + c) / f32(2)
        a0, a1, a2 = f32(1) + alpha, f32(-2) * c, f32(1) - alpha
        return [b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0]

    return dict(hp=biquad(40.0, False), lp=biquad(14000.0, True))


def np_filter_gate(x, p):
    """mod.rs:433-487 with the coefficient values of the C oracle (so that the arithmetic, not libm, is compared)."""
    hp, lp = [f32(v) for v in p.hp], [f32(v) for v in p.lp]
    thr, rc, hold_n = f32(p.gate_threshold_linear), f32(p.release_coeff), int(p.gate_hold_samples)
    hx1 = hx2 = hy1 = hy2 = lx1 = lx2 = ly1 = ly2 = env = f32(0)
    hold = 0
    out = np.empty_like(x)
    for i, v in enumerate(x):
        v = f32(v)
        h = hp[0] * v + hp[1] * hx1 + hp[2] * hx2 - hp[3] * hy1 - hp[4] * hy2
        hx2, hx1, hy2, hy1 = hx1, v, hy1, h
        lo = lp[0] * h + lp[1] * lx1 + lp[2] * lx2 - lp[3] * ly1 - lp[4] * ly2
        lx2, lx1, ly2, ly1 = lx1, h, ly1, lo
        a = abs(lo)
        if a > env:
            env, hold = a, hold_n
        else:
            env = rc * env + (f32(1) - rc) * a
        if env >= thr:
            g = f32(1)
        elif hold > 0:
            hold -= 1
            g = f32(1)
        else:
            r = env / thr
            g = r * r * r * r
        out[i] = lo * g
    return out


def test_coefficients_follow_the_reference_formulas(O):
    for sr in (44100.0, 48000.0):
        p = O.cond_params(sr, 1024)
        q = np_params(sr, 1024)
        assert np.allclose(list(p.hp), q["hp"], rtol=3e-7, atol=0)
        assert np.allclose(list(p.lp), q["lp"], rtol=3e-7, atol=0)
        assert p.gate_threshold_linear == pytest.approx(1e-3, rel=1e-6)
        assert p.gate_hold_samples == int(f32(0.020) * f32(sr))
        assert p.release_coeff == pytest.approx(math.exp(-1.0 / (0.040 * sr)), rel=1e-6)
        slot_rate = sr / 1024
        assert p.smooth_alpha == pytest.approx(1 - math.exp(-1 / (240 * slot_rate)), rel=2e-3)
        assert p.silence_decay_alpha == pytest.approx(1 - math.exp(-1 / (10 * slot_rate)), rel=1e-4)
        # HPF: zero DC gain (b0 + b1 + b2 == 0 up to rounding); LPF: unit DC gain
        assert abs(sum(p.hp[:3])) < 1e-6
        assert sum(p.lp[:3]) / (1 + p.lp[3] + p.lp[4]) == pytest.approx(1.0, rel=1e-5)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_filter_gate_is_bit_identical_to_the_numpy_restatement(O, seed):
    rng = np.random.default_rng(seed)
    sr, L = 48000.0, 256
    n = 48 * L
    t = np.arange(n) / sr
    # short burst, then a level that sinks through the -60 dBFS gate threshold (hold, then ratio^4), then near silence
    x = 0.02 * np.sin(2 * np.pi * 330 * t) * (t < 0.01) + 3e-3 * rng.standard_normal(n) * np.exp(-t * 200) \
        + 2e-5 * rng.standard_normal(n)
    x = x.astype(np.float32)
    y, _ = O.condition_clip(x, sr, slot_len=L, agc=False)
    ref = np_filter_gate(x, O.cond_params(sr, L))
    assert np.array_equal(y.view(np.uint32), ref.view(np.uint32))
    assert np.abs(y[-L:]).max() < 1e-3 * np.abs(x[-L:]).max()       # the gate did close


def test_gate_and_filters_known_answers(O):
    sr, L = 48000.0, 1024
    n = 40 * L
    # DC is removed by the 40 Hz high-pass
    y, _ = O.condition_clip(np.full(n, 0.25, np.float32), sr, L, agc=False)
    assert np.abs(y[-L:]).max() < 1e-4
    # a 1 kHz tone in the pass band goes through the filters and an open gate essentially unchanged
    t = np.arange(n) / sr
    x = (0.1 * np.sin(2 * np.pi * 1000 * t)).astype(np.float32)
    y, _ = O.condition_clip(x, sr, L, agc=False)
    assert np.abs(y[-4 * L:]).max() == pytest.approx(0.1, rel=0.02)
    # a STEADY tone 20 dB below the gate threshold still passes: every cycle's peak exceeds the decayed envelope
    # and re-arms the 20 ms hold (mod.rs:461-464), so the ratio^4 branch is never reached
    x = (1e-4 * np.sin(2 * np.pi * 1000 * t)).astype(np.float32)
    y, _ = O.condition_clip(x, sr, L, agc=False)
    assert np.abs(y[-4 * L:]).max() == pytest.approx(1e-4, rel=0.02)


def np_agc(slots, p):
    """dynamics.rs:194-360 on already gated slots (float32 scalars, numpy sorts)."""
    L = slots.shape[1]
    long_h, play_h = np.zeros(256, np.float32), np.zeros(5000, np.float32)
    lpos = lfill = ppos = pfill = 0
    gain = f32(1)
    out = []

    def db(v):
        return f32(20) * f32(math.log10(float(max(f32(v), f32(1e-9)))))

    for s in slots:
        ss = f32(0)
        for v in s:
            ss = ss + v * v
        rms = f32(math.sqrt(float(ss / f32(L))))     # sqrt of an f32 in double, rounded back = IEEE sqrtf
        rms_db = db(rms)
        ln = 256 if lfill else max(lpos, 1)
        srt = np.sort(long_h[:ln])
        nf_db = db(max(srt[int(f32(ln - 1) * f32(0.10))], f32(1e-9)))
        floor_db = nf_db if ln >= 32 else f32(p.bootstrap_floor_db)
        active = rms_db > floor_db + f32(p.active_snr_db)
        broadband = False
        if active:
            msq = rms * rms
            q = f32(0)
            for v in s:
                v2 = v * v
                q = q + v2 * v2
            mq = q / f32(L)
            k = mq / (msq * msq) if msq > f32(1e-18) else f32(3)
            broadband = bool(k >= f32(2.75) and k <= f32(3.8) and rms_db < f32(-45))
        playing = active and not broadband
        if (not active) or broadband:
            long_h[lpos] = rms
            lpos = (lpos + 1) % 256
            lfill = lfill or lpos == 0
        if playing:
            play_h[ppos] = rms
            ppos = (ppos + 1) % 5000
            pfill = pfill or ppos == 0
        pn = 5000 if pfill else ppos
        if pn > 0:
            srt = np.sort(play_h[:pn])
            med_db = db(max(srt[(pn - 1) // 2], f32(1e-9)))
            p95_db = db(max(srt[int(f32(pn - 1) * f32(0.95))], f32(1e-9)))
            raw = min(max(f32(p.target_db) - p95_db, f32(0)), f32(p.max_boost_db))
        else:
            raw, med_db = f32(0), rms_db
        if playing:
            tgt = f32(math.pow(10.0, float(raw / f32(20))))
            gain = gain + f32(p.smooth_alpha) * (tgt - gain)
        else:
            gain = gain + f32(p.silence_decay_alpha) * (f32(1) - gain)
        peak = max(np.abs(s).max(), f32(1e-9))
        eff = min(gain, f32(0.97) / peak)
        if not playing:
            lvl = 0
        else:
            r = rms_db - med_db
            lvl = 1 + sum(r >= f32(b) for b in (-15.0, -9.0, -4.5, -1.5, 1.5, 4.5, 9.0))
        out.append((lvl, rms_db, db(eff), med_db, nf_db, eff, int(active) | 2 * int(broadband) | 4 * int(playing)))
    return out


def test_agc_matches_the_numpy_restatement(O):
    rng = np.random.default_rng(5)
    sr, L, n_slots = 48000.0, 64, 700
    n = L * n_slots
    t = np.arange(n) / sr
    env = np.where((t * 7).astype(int) % 3 == 0, 0.0, 1.0) * (0.05 + 0.3 * np.abs(np.sin(2 * np.pi * 0.9 * t)))
    x = (env * np.sin(2 * np.pi * 523.25 * t) + 3e-4 * rng.standard_normal(n)).astype(np.float32)
    gated, _ = O.condition_clip(x, sr, L, agc=False)
    y, dyn = O.condition_clip(x, sr, L, agc=True)
    ref = np_agc(gated.reshape(n_slots, L), O.cond_params(sr, L))
    lvl = np.array([r[0] for r in ref])
    eff = np.array([r[5] for r in ref], np.float64)
    assert (lvl == dyn["level"]).mean() > 0.99          # log10 differs by an ulp between libm paths: near-ties
    assert np.allclose(dyn["effective_gain"], eff, rtol=2e-5)
    assert np.allclose(dyn["rms_db"], [r[1] for r in ref], atol=2e-4)
    assert np.allclose(dyn["noise_floor_db"], [r[4] for r in ref], atol=2e-4)
    assert np.allclose(dyn["session_median_db"], [r[3] for r in ref], atol=2e-4)
    assert (dyn["flags"] == np.array([r[6] for r in ref])).mean() > 0.99
    # the gain was applied to the gated signal
    assert np.allclose(y.reshape(n_slots, L), gated.reshape(n_slots, L) * dyn["effective_gain"][:, None], rtol=1e-6, atol=0)
    assert set(np.unique(dyn["level"])) >= {0, 5}


def test_agc_known_answers(O):
    sr, L = 48000.0, 1024
    # silence: never active, level Silence, gain stays at 1 (0 dB), noise floor reads the 1e-9 clamp (-180 dB)
    y, dyn = O.condition_clip(np.zeros(50 * L, np.float32), sr, L)
    assert np.all(dyn["level"] == 0) and np.all(dyn["effective_gain"] == 1.0)
    assert dyn["noise_floor_db"][0] == pytest.approx(-180.0, abs=1e-3)
    # a steady -20 dBFS tone: active from the first slot, classified mf against its own median, and the gain
    # creeps towards target(-18 dBFS) - p95 with the 240 s time constant
    t = np.arange(200 * L) / sr
    x = (0.1 * math.sqrt(2) * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    y, dyn = O.condition_clip(x, sr, L)
    assert np.all(dyn["level"][2:] == 5)
    assert dyn["rms_db"][-1] == pytest.approx(-20.0, abs=0.05)
    g = dyn["effective_gain"]
    assert np.all(np.diff(g[2:]) > 0) and 1.0 < g[-1] < 10 ** (2.0 / 20)
    # peak headroom: a full-scale slot can never be boosted above 0.97 / peak
    x[:] = 0.999 * np.sign(np.sin(2 * np.pi * 440 * t))
    y, dyn = O.condition_clip(x, sr, L)
    assert np.all(dyn["effective_gain"] <= 0.97 / 0.9 + 1e-6)


def test_only_full_slots_are_processed(O):
    sr, L = 44100.0, 1024
    x = np.random.default_rng(3).standard_normal(3 * L + 100).astype(np.float32) * 0.1
    y, dyn = O.condition_clip(x, sr, L)
    assert len(dyn) == 3
    assert np.array_equal(y[3 * L:], x[3 * L:])         # mod.rs:799-803: a partial slot is never delivered
    assert not np.array_equal(y[:3 * L], x[:3 * L])


# ---- tuner post-stage: PINNED by the reference's own known-answer tests -------------

def test_interval_known_answers_of_the_reference(O):
    """theory.rs:545-583: the reference's own Interval tests, replayed on the oracle."""
    c4 = np.float32(261.63)
    for semis, name in [(7, "Per5"), (12, "Per8"), (4, "Maj3"), (3, "Min3"), (5, "Per4")]:
        hi = c4 * np.float32(2.0) if semis == 12 else c4 * np.float32(2.0) ** np.float32(semis / 12.0)
        got, acc = O.interval(float(c4), float(hi), 0)
        assert got == name and abs(acc) < 1.0
    # theory.rs:307-312: a zero first frequency returns the Per8 default instead of dividing by zero
    assert O.interval(0.0, 440.0, 0) == ("Per8", 0.0)
    # ratios above an octave fold down (:314-316); the three tuning systems use their own tables
    assert O.interval(220.0, 660.0, 0)[0] == "Per5" and O.interval(220.0, 660.0, 1) == ("Per5", pytest.approx(0.0, abs=1e-3))
    assert O.interval(243.0, 256.0, 2)[0] == "Min2" and abs(O.interval(243.0, 256.0, 2)[1]) < 1e-3


def test_tuner_frame_branches(O):
    """tuner.rs:152-193."""
    assert O.tuner_frame(np.zeros((0, 2)))["kind"] == 0
    assert O.tuner_frame([[440.0, 0.4]]) == dict(kind=1, best=0, lo=0, hi=0, interval=0, accuracy=0.0)
    r = O.tuner_frame([[660.0, 0.4], [440.0, 0.9]])
    assert (r["kind"], r["lo"], r["hi"], O.INT_TYPES[r["interval"]]) == (2, 1, 0, "Per5")
    assert O.tuner_frame([[440.0, 0.4], [550.0, 0.9], [660.0, 0.9]])["kind"] == 3
    # SinglePitch mode: the highest score wins, the LAST one on ties (Iterator::max_by)
    assert O.tuner_frame([[440.0, 0.4], [550.0, 0.9], [660.0, 0.9]], 0, True)["best"] == 2


# ---- input callback: device sample format -> mono f32 (mod.rs:765-792, dasp_sample 0.11.0) --------------

def test_ingest_follows_the_published_conversions(O):
    rng = np.random.default_rng(9)
    i16 = rng.integers(-32768, 32768, 4096, dtype=np.int16)
    i16[:4] = [-32768, 32767, 0, -1]
    # dasp_sample: i16 -> f32 is s / 32768 (exact), u16 goes through i16
    assert np.array_equal(O.ingest(i16, O.PCM_I16, 1), i16.astype(np.float32) / np.float32(32768.0))
    u16 = (i16.astype(np.int32) + 32768).astype(np.uint16)
    assert np.array_equal(O.ingest(u16, O.PCM_U16, 1), O.ingest(i16, O.PCM_I16, 1))
    # stereo: (0 + l + r) / 2; more than two channels: only the first two are mixed (mod.rs:777, 786-791)
    st = i16.reshape(-1, 2)
    want = (st[:, 0].astype(np.float32) / 32768 + st[:, 1].astype(np.float32) / 32768) / np.float32(2)
    assert np.array_equal(O.ingest(st, O.PCM_I16, 2), want)
    quad = i16.reshape(-1, 4)
    assert np.array_equal(O.ingest(quad, O.PCM_I16, 4), O.ingest(np.ascontiguousarray(quad[:, :2]), O.PCM_I16, 2))
    f = rng.standard_normal(1024).astype(np.float32)
    assert np.array_equal(O.ingest(f, O.PCM_F32, 1), f)
    fs = f.reshape(-1, 2)
    assert np.array_equal(O.ingest(fs, O.PCM_F32, 2), ((np.float32(0) + fs[:, 0]) + fs[:, 1]) / np.float32(2))

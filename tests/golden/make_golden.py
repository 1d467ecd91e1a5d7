#!/usr/bin/env python
"""Mint the golden fixtures under tests/golden/.

The reference (a Rust crate) cannot be built or imported here and holds no test
vector for this path, so these fixtures are NOT outputs of the
reference binary: they come from an independent numpy restatement of the
reference source written separately from oracle/aa_oracle.c -- array-at-a-time
float32 numpy for the per-bin recurrences, float64 pocketfft for the spectrum.
Two independent readings agreeing is the strongest pin available ("parity
unpinned" in DESIGN.md).  Citations: /root/reference/src/audio_io/stft.rs and
src/analysis/onset.rs.

Run:  python tests/golden/make_golden.py     (rewrites tests/golden/*.npz)
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import signals  # noqa: E402

F = np.float32
MAX_HARMONICS, MAX_NOTES = 14, 8


def hann(n):
    # stft.rs:641-648 -- x = i/n in f32; 0.5 - 0.5*cos(2*pi_f32*x)
    i = np.arange(n, dtype=F)
    x = i / F(n)
    return (F(0.5) - F(0.5) * np.cos(F(2.0) * F(np.pi) * x)).astype(F)


def mags_f64(x, n, hop, window):
    fr = signals.frames_of(x, n, hop).astype(F) * window[None, :]   # f32 product, stft.rs:298
    spec = np.fft.rfft(fr.astype(np.float64), axis=1)
    return np.abs(spec).astype(F)


def global_floor(db, half):
    # stft.rs:323-324
    return F(F(10.0) ** (F(db) / F(20.0))) * F(half) / F(2.0)


class PitchFloorNp:
    """stft.rs:209-224, 326-367, vectorised over bins."""

    def __init__(self, half):
        self.nf = np.zeros(half, F)
        self.prev = np.zeros(half, F)
        self.vol = np.zeros(half, F)
        self.init = False

    def update(self, mag, gf):
        if not self.init:
            self.nf = np.maximum(mag, gf * F(5.0)).astype(F)
            self.prev = mag.copy()
            self.init = True
        else:
            floor = self.nf
            delta = np.abs(mag - self.prev)
            self.vol = (self.vol * F(0.75) + delta * (F(1.0) - F(0.75))).astype(F)
            self.prev = mag.copy()
            above = mag / np.maximum(floor, F(0.01))
            vol_norm = np.clip(self.vol / np.maximum(mag, F(0.05)), F(0.0), F(1.0)).astype(F)
            sustained = (above > F(1.5)) & (vol_norm < F(0.15))
            alpha = np.where(mag > floor, F(0.04) + (F(0.35) - F(0.04)) * vol_norm, F(0.02)).astype(F)
            upd = (floor + alpha * (mag - floor)).astype(F)
            self.nf = np.where(sustained, floor, upd).astype(F)
        return np.minimum(self.nf, gf * F(2.5)).astype(F)


def extract_pitches(m, half, bw, fmin, fmax, nf):
    """stft.rs:443-620.  Returns (pairs list, peak mask, out bins)."""
    min_bin = max(int(np.ceil(F(fmin) / F(bw))), 1)
    max_bin = min(int(np.floor(F(fmax) / F(bw))), max(half - 2, 0))
    mask = np.zeros(half, np.uint8)
    if min_bin >= max_bin:
        return [], mask, []
    k = np.arange(min_bin + 1, max_bin)
    pk = (m[k] > nf[k]) & (m[k] >= m[k - 1]) & (m[k] >= m[k + 1])
    peaks = k[pk]
    mask[peaks] = 1
    if len(peaks) == 0:
        return [], mask, []
    scores = {}
    frac = {}
    for kk in peaks:
        kk = int(kk)
        fund = m[kk]
        if fund < nf[kk] * F(5.0):
            scores[kk] = F(0.0)
            frac[kk] = F(0.0)
            continue
        yl, yc, yr = np.log(m[kk - 1]), np.log(m[kk]), np.log(m[kk + 1])
        denom = F(yl - F(2.0) * yc + yr)
        if abs(denom) < F(1e-30):
            delta = F(0.0)
        else:
            delta = F(np.clip(F(0.5) * F(yl - yr) / denom, F(-1.0), F(1.0)))
        fb = F(F(kk) + delta)
        frac[kk] = fb
        score = F(fund)
        last = kk
        longest = cur = total = 0
        for n in range(2, MAX_HARMONICS + 1):
            ef = F(fb * F(n))
            if ef >= F(half):
                break
            s0 = max(int(np.floor(F(ef - F(1.0)))), 0)
            s0 = max(s0, last + 1)
            s1 = min(int(np.ceil(F(ef + F(1.0)))), half - 1)
            best, bm = 0, F(0.0)
            for h in range(s0, s1 + 1):
                if mask[h] and m[h] > bm:
                    bm, best = m[h], h
            if best != 0:
                score = F(score + bm)
                last = best
                cur += 1
                total += 1
            else:
                longest = max(longest, cur)
                cur = 0
        longest = max(longest, cur)
        if longest < 3 and fund < F(15.0) * nf[kk]:
            scores[kk] = F(0.0)
        else:
            ls = F(np.log2(F(F(0.5) + score)))
            sm = F(F(F(1.0) + F(longest) + F(total) / F(2.0)) / F(F(1.0) + F(MAX_HARMONICS)))
            scores[kk] = F(ls * sm)
    max_score = F(0.0)
    for kk in peaks:
        max_score = max(max_score, scores[int(kk)])
    if max_score == 0.0:
        return [], mask, []
    cutoff = F(max_score * F(0.5))
    cand = [(int(kk), scores[int(kk)]) for kk in peaks if scores[int(kk)] >= cutoff]
    keep = []
    for i, (bi, si) in enumerate(cand):
        fi = F(frac[bi] * F(bw))
        sup = False
        for j, (bj, sj) in enumerate(cand):
            if i == j:
                continue
            fj = F(frac[bj] * F(bw))
            ratio = F(fi / fj)
            nearest = F(np.floor(ratio + F(0.5)))     # round half away (ratio > 0)
            if 2.0 <= nearest <= 5.0 and abs(F(ratio / nearest - F(1.0))) < F(0.03) \
                    and si < F(sj * F(1.05)):
                sup = True
                break
        if not sup:
            keep.append((bi, si))
    keep.sort(key=lambda c: -float(c[1]))           # stable: ties keep ascending bin
    ded = []
    for bi, si in keep:
        if not any(abs(F(frac[bi] - frac[bj])) < F(2.0) for bj, _ in ded):
            ded.append((bi, si))
    ded = ded[:MAX_NOTES]
    out, bins = [], []
    for bi, si in ded:
        fq = F(frac[bi] * F(bw))
        if fmin <= fq <= fmax:
            out.append((fq, si))
            bins.append(bi)
    return out, mask, bins


class TrackerNp:
    """stft.rs:19-117"""

    def __init__(self):
        self.tr = []          # [freq, score, life]

    def process(self, raw, onset):
        matched = [False] * len(self.tr)
        for rf, rs in raw:
            found = False
            for i, t in enumerate(self.tr):
                if matched[i]:
                    continue
                if F(abs(F(t[0] - rf)) / t[0]) < F(0.03):
                    t[0] = F(rf) if onset else F(F(t[0] * F(0.6)) + F(rf * F(0.4)))
                    t[1] = F(rs)
                    t[2] = min(t[2] + 1, 3)
                    matched[i] = True
                    found = True
                    break
            if not found:
                self.tr.append([F(rf), F(rs), 1])
                matched.append(True)
        out = []
        i = 0
        while i < len(self.tr):
            if not matched[i]:
                self.tr[i][2] = 0 if onset else self.tr[i][2] - 1
            if self.tr[i][2] <= 0:
                del self.tr[i]
                del matched[i]
            else:
                if self.tr[i][2] >= 2:
                    out.append((self.tr[i][0], self.tr[i][1]))
                i += 1
        return out


class OnsetNp:
    """onset.rs:149-186, 261-357, 47-84"""

    def __init__(self, half):
        self.half = half
        self.prev = np.zeros(half, F)
        self.nf = np.zeros(half, F)
        self.init = False
        self.ema = F(0.0)
        self.thr = F(0.0)
        self.since = 4            # frames_since_onset, onset.rs:200

    def frame(self, m, gf):
        half = self.half
        sm = m.copy()
        sm[1:-1] = ((m[:-2] + m[1:-1]).astype(F) + m[2:]).astype(F) / F(3.0)
        w = (F(1.0) - np.arange(half, dtype=F) / F(half)).astype(F)
        diff = (sm - self.prev).astype(F)
        contrib = np.where(diff > 0, diff * w, F(0.0)).astype(F)
        # sequential f32 accumulation, skipping non-positive terms exactly as the loop does
        flux = F(0.0)
        pos = contrib[diff > 0]
        if len(pos):
            flux = np.cumsum(pos, dtype=F)[-1]
        energy = np.cumsum(m, dtype=F)[-1]
        self.prev = m.copy()
        eps = max(gf, F(0.01))
        if not self.init:
            self.nf = np.maximum(m, gf).astype(F)
            self.init = True
        fk = np.maximum(self.nf, eps)
        r = (m / fk).astype(F)
        burst = r > F(2.5)
        rise = (~burst) & (m > self.nf)
        nf = np.where(burst, m * F(1.3),
                      np.where(rise, self.nf + F(0.1) * (m - self.nf), self.nf + F(0.04) * (m - self.nf)))
        self.nf = nf.astype(F)
        count = int(burst.sum())
        max_ex = F(max(F(0.0), r.max()))
        if count < 2:
            flux = F(0.0)
        mem = F(0.84) if energy > self.ema else F(0.95)
        self.ema = F(F(self.ema * mem) + F(energy * F(F(1.0) - mem)))
        memory = F(0.84) if flux > self.thr else F(0.89)
        is_on = flux > self.thr
        self.thr = F(F(self.thr * memory) + F(flux * F(F(1.0) - memory)))
        if self.thr < F(0.9):
            self.thr = F(0.9)
        flux_onset = bool(is_on and flux > F(self.thr * F(1.5)))
        burst_onset = bool(max_ex > F(3.0) and count >= 3)
        rising = bool(energy > F(self.ema * F(1.5)))
        detected = flux_onset and burst_onset
        fired = detected and rising and self.since >= 3          # onset.rs:403 (no ticks, calibrated)
        if fired or (detected and self.since < 3):               # onset.rs:535-539
            self.since = 0
        else:
            self.since += 1
        flags = (1 if flux_onset else 0) | (2 if burst_onset else 0) | \
                (4 if detected else 0) | (8 if rising else 0) | (16 if fired else 0)
        return flux, energy, count, max_ex, flags, self.ema


def centroid(m, bw):
    k = np.arange(len(m), dtype=np.float64)
    den = m.astype(np.float64).sum()
    return F(0.0) if den <= 0 else F((k * m).sum() / den * float(bw))


def run_case(name, x, sr, n, hop, db=-96.0, onset_pattern=None, max_frames=None):
    half = n // 2 + 1
    w = hann(n)
    mags = mags_f64(x, n, hop, w)
    if max_frames:
        mags = mags[:max_frames]
        x = x[: n + (max_frames - 1) * hop]
    T = mags.shape[0]
    bw = F(F(sr) / F(n))
    gf = global_floor(db, half)
    pf, on, trk = PitchFloorNp(half), OnsetNp(half), TrackerNp()
    n_p = np.zeros(T, np.uint32)
    pit = np.zeros((T, MAX_NOTES, 2), F)
    bins = -np.ones((T, MAX_NOTES), np.int32)
    n_s = np.zeros(T, np.uint32)
    stab = np.zeros((T, 16, 2), F)
    masks = np.zeros((T, half), np.uint8)
    sc = np.zeros((T, 5), F)      # flux, energy, centroid, max_excess, ema
    ic = np.zeros((T, 2), np.uint32)   # burst_count, flags
    floors = np.zeros((T, half), F)
    onset_in = np.zeros(T, np.uint8)
    if onset_pattern is not None:
        onset_in[onset_pattern[onset_pattern < T]] = 1
    for t in range(T):
        eff = pf.update(mags[t], gf)
        floors[t] = eff
        out, mask, ob = extract_pitches(mags[t], half, bw, 24.0, 10000.0, eff)
        masks[t] = mask
        n_p[t] = len(out)
        for i, (fq, s) in enumerate(out):
            pit[t, i] = (fq, s)
            bins[t, i] = ob[i]
        st = trk.process(out, bool(onset_in[t]))
        n_s[t] = len(st)
        for i, (fq, s) in enumerate(st[:16]):
            stab[t, i] = (fq, s)
        flux, energy, count, max_ex, flags, ema = on.frame(mags[t], gf)
        sc[t] = (flux, energy, centroid(mags[t], bw), max_ex, ema)
        ic[t] = (count, flags)
    # keep the floor only as a digest + a few frames (size)
    keep = sorted(set([0, 1, 2, T // 2, T - 1]))
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        samples=x.astype(F), sr=np.float64(sr), n=np.int32(n), hop=np.int32(hop), db=np.float64(db),
        window=w, mags=mags, n_pitches=n_p, pitches=pit, out_bins=bins, n_stable=n_s, stable=stab,
        peak_bits=np.packbits(masks, axis=1), scalars=sc, ints=ic, onset_in=onset_in,
        floor_frames=np.array(keep, np.int32), floors=floors[keep],
        floor_sum=floors.astype(np.float64).sum(axis=1),
    )
    print(f"{name}: T={T} half={half} pitched_frames={(n_p > 0).sum()} "
          f"onsets={(ic[:, 1] & 4 > 0).sum()} max_mag={mags.max():.4f}")


def main():
    run_case("cfg1_sine440_2048", signals.sine(440.0, 44100.0, 441000), 44100.0, 2048, 512, max_frames=64)
    run_case("sine440_48k_4096", signals.sine(440.0, 48000.0, 96000), 48000.0, 4096, 1024, max_frames=40)
    run_case("multitone_48k_4096", signals.multitone(7, 48000.0, 96000), 48000.0, 4096, 1024, max_frames=40)
    run_case("multitone_44k_2048", signals.multitone(11, 44100.0, 66150), 44100.0, 2048, 512, max_frames=64,
             onset_pattern=np.array([10, 33]))
    run_case("notes_48k_256", signals.note_sequence(3, 48000.0, 24000), 48000.0, 256, 64, max_frames=360)
    run_case("notes_48k_1024", signals.note_sequence(5, 48000.0, 48000), 48000.0, 1024, 256, max_frames=160)


if __name__ == "__main__":
    main()

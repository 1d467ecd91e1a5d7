"""ctypes binding of the CPU oracle (oracle/aa_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/aa_oracle.h.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  PARITY UNPINNED: the reference cannot be built here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libaa_oracle.so")

FEAT_PITCH, FEAT_ONSET, FEAT_CENTROID, FEAT_TRACKER = 1, 2, 4, 8
FLAG_FLUX_ONSET, FLAG_BURST_ONSET, FLAG_ONSET_DETECTED, FLAG_ENERGY_RISING, FLAG_ONSET_FIRED = 1, 2, 4, 8, 16
MAX_NOTES, MAX_STABLE = 8, 16

FEATURES_DTYPE = np.dtype(
    [
        ("n_pitches", "<u4"),
        ("pitch", [("freq", "<f4"), ("score", "<f4")], (MAX_NOTES,)),
        ("flux", "<f4"),
        ("energy", "<f4"),
        ("centroid", "<f4"),
        ("burst_count", "<u4"),
        ("max_excess", "<f4"),
        ("flags", "<u4"),
        ("energy_ema", "<f4"),
    ]
)
assert FEATURES_DTYPE.itemsize == 96

STABLE_DTYPE = np.dtype(
    [
        ("n", "<u4"),
        ("reserved", "<u4"),
        ("pitch", [("freq", "<f4"), ("score", "<f4")], (MAX_STABLE,)),
    ]
)
assert STABLE_DTYPE.itemsize == 136

DIAG_DTYPE = np.dtype(
    [
        ("n_peaks", "<i4"),
        ("n_scored", "<i4"),
        ("n_candidates", "<i4"),
        ("n_out", "<i4"),
        ("out_bins", "<i4", (MAX_NOTES,)),
        ("min_margin", "<f4"),
        ("margin_src", "<i4"),
        ("cand_eps", "<f4"),
        ("cand_src", "<i4"),
    ]
)
assert DIAG_DTYPE.itemsize == 64


class Taps(C.Structure):
    """aao_taps (aa_oracle.h)."""
    _fields_ = [(k, C.c_void_p) for k in ("mags", "floor", "peak_mask", "features", "stable", "diag",
                                          "pitch_nf", "pitch_vol", "onset_nf")]


class Config(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("hop", C.c_int32),
        ("sample_rate", C.c_float),
        ("min_freq", C.c_float),
        ("max_freq", C.c_float),
        ("noise_floor_db", C.c_float),
        ("features", C.c_uint32),
    ]


class CondParams(C.Structure):
    """aao_cond_params (aa_oracle.h): coefficients of the conditioning chain."""
    _fields_ = [
        ("hp", C.c_float * 5),
        ("lp", C.c_float * 5),
        ("gate_threshold_linear", C.c_float),
        ("release_coeff", C.c_float),
        ("gate_hold_samples", C.c_int32),
        ("target_db", C.c_float),
        ("max_boost_db", C.c_float),
        ("smooth_alpha", C.c_float),
        ("silence_decay_alpha", C.c_float),
        ("active_snr_db", C.c_float),
        ("bootstrap_floor_db", C.c_float),
        ("slot_len", C.c_int32),
    ]


DYNAMICS_DTYPE = np.dtype(
    [
        ("level", "<i4"),
        ("rms_db", "<f4"),
        ("gain_db", "<f4"),
        ("session_median_db", "<f4"),
        ("noise_floor_db", "<f4"),
        ("effective_gain", "<f4"),
        ("flags", "<u4"),
        ("reserved", "<u4"),
    ]
)
assert DYNAMICS_DTYPE.itemsize == 32


def make_config(n=2048, hop=512, sample_rate=44100.0, min_freq=24.0, max_freq=10000.0,
                noise_floor_db=-96.0, features=FEAT_PITCH | FEAT_ONSET | FEAT_CENTROID | FEAT_TRACKER):
    return Config(n, hop, sample_rate, min_freq, max_freq, noise_floor_db, features)


def build(force: bool = False) -> str:
    """Compile oracle/aa_oracle.c -> oracle/libaa_oracle.so (gcc)."""
    srcs = [os.path.join(_HERE, f) for f in ("aa_oracle.c", "aa_oracle_cond.c", "aa_oracle.h")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(p) > os.path.getmtime(_SO) for p in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libaa_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    fp = C.POINTER(C.c_float)
    dp = C.POINTER(C.c_double)
    u8p = C.POINTER(C.c_uint8)
    vp = C.c_void_p
    L.aao_hann_window.argtypes = [C.c_int, fp]
    L.aao_fft_create.restype = vp
    L.aao_fft_create.argtypes = [C.c_int]
    L.aao_fft_destroy.argtypes = [vp]
    L.aao_fft_forward.argtypes = [vp, fp, fp]
    L.aao_rdft_f64.argtypes = [dp, C.c_int, dp]
    L.aao_magnitudes.argtypes = [fp, C.c_int, fp]
    L.aao_global_floor.restype = C.c_float
    L.aao_global_floor.argtypes = [C.c_float, C.c_int]
    L.aao_pitch_floor_create.restype = vp
    L.aao_pitch_floor_create.argtypes = [C.c_int]
    L.aao_pitch_floor_destroy.argtypes = [vp]
    L.aao_pitch_floor_reset.argtypes = [vp]
    L.aao_pitch_floor_update.argtypes = [vp, fp, C.c_float, fp]
    L.aao_extract_pitches.restype = C.c_int
    L.aao_extract_pitches.argtypes = [fp, C.c_int, C.c_float, C.c_float, C.c_float, fp, fp, u8p, vp]
    L.aao_tracker_create.restype = vp
    L.aao_tracker_destroy.argtypes = [vp]
    L.aao_tracker_reset.argtypes = [vp]
    L.aao_tracker_process.restype = C.c_int
    L.aao_tracker_process.argtypes = [vp, fp, C.c_int, C.c_int, fp, C.c_int]
    L.aao_onset_create.restype = vp
    L.aao_onset_create.argtypes = [C.c_int]
    L.aao_onset_destroy.argtypes = [vp]
    L.aao_onset_reset.argtypes = [vp]
    L.aao_onset_frame.argtypes = [vp, fp, C.c_float, vp]
    L.aao_centroid.restype = C.c_float
    L.aao_centroid.argtypes = [fp, C.c_int, C.c_float]
    L.aao_note_from_freq.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int), fp]
    L.aao_yin_lag.restype = C.c_int
    L.aao_yin_lag.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_float, dp]
    L.aao_num_frames.restype = C.c_int64
    L.aao_num_frames.argtypes = [C.c_int64, C.c_int, C.c_int]
    L.aao_analyze_clip.restype = C.c_int64
    L.aao_analyze_clip.argtypes = [C.POINTER(Config), fp, C.c_int64, fp, u8p, fp, fp, u8p, vp, vp, vp]
    L.aao_analyze_clip_ex.restype = C.c_int64
    L.aao_analyze_clip_ex.argtypes = [C.POINTER(Config), fp, C.c_int64, fp, u8p, C.POINTER(Taps)]
    L.aao_analyze_batch.restype = C.c_int64
    L.aao_analyze_batch.argtypes = [C.POINTER(Config), fp, C.c_int64, C.c_int64, C.c_int, fp, vp, vp]
    L.aao_cond_params_init.argtypes = [C.POINTER(CondParams), C.c_float, C.c_int]
    L.aao_interval.restype = C.c_int
    L.aao_interval.argtypes = [C.c_float, C.c_float, C.c_int, fp]
    ip = C.POINTER(C.c_int)
    L.aao_tuner_frame.argtypes = [fp, C.c_int, C.c_int, C.c_int, ip, ip, ip, ip, ip, fp]
    L.aao_ingest.argtypes = [vp, C.c_int, C.c_int, C.c_int64, fp]
    L.aao_onset_gates.restype = None
    L.aao_onset_gates.argtypes = [C.POINTER(C.c_float)] * 3 + [C.POINTER(C.c_uint32)] * 2
    L.aao_stamp_onset.restype = None
    L.aao_stamp_onset.argtypes = [C.c_double, C.c_int64, C.c_float, C.c_float, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                  C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.aao_onset_events.restype = C.c_int64
    L.aao_onset_events.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int64, vp]
    L.aao_cond_clip.restype = C.c_int64
    L.aao_cond_clip.argtypes = [C.POINTER(CondParams), fp, C.c_int64, vp, C.c_int]
    _lib = L
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def hann_window(n: int) -> np.ndarray:
    w = np.empty(n, np.float32)
    lib().aao_hann_window(n, _fp(w))
    return w


def rfft_f32(x: np.ndarray) -> np.ndarray:
    """FftProcessor::process_forward restatement; returns complex64[n/2+1]."""
    x = np.ascontiguousarray(x, np.float32).copy()
    n = x.shape[0]
    p = lib().aao_fft_create(n)
    if not p:
        raise ValueError("n must be a power of two >= 4")
    spec = np.empty(2 * (n // 2 + 1), np.float32)
    lib().aao_fft_forward(p, _fp(x), _fp(spec))
    lib().aao_fft_destroy(p)
    return spec.view(np.complex64)


def rdft_f64(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float64)
    n = x.shape[0]
    spec = np.empty(2 * (n // 2 + 1), np.float64)
    lib().aao_rdft_f64(x.ctypes.data_as(C.POINTER(C.c_double)), n,
                       spec.ctypes.data_as(C.POINTER(C.c_double)))
    return spec.view(np.complex128)


def magnitudes(spec: np.ndarray) -> np.ndarray:
    s = np.ascontiguousarray(spec, np.complex64).view(np.float32)
    half = s.shape[0] // 2
    m = np.empty(half, np.float32)
    lib().aao_magnitudes(_fp(s), half, _fp(m))
    return m


def global_floor(noise_floor_db: float, half: int) -> float:
    return float(lib().aao_global_floor(noise_floor_db, half))


def extract_pitches(mags, floor, bin_width, min_freq=24.0, max_freq=10000.0):
    """Returns (pairs[n,2] float32, peak_mask uint8[half], diag record)."""
    mags = np.ascontiguousarray(mags, np.float32)
    floor = np.ascontiguousarray(floor, np.float32)
    half = mags.shape[0]
    pairs = np.zeros((MAX_NOTES, 2), np.float32)
    mask = np.zeros(half, np.uint8)
    diag = np.zeros(1, DIAG_DTYPE)
    n = lib().aao_extract_pitches(_fp(mags), half, bin_width, min_freq, max_freq, _fp(floor),
                                  _fp(pairs), _u8p(mask), _vp(diag))
    return pairs[:n].copy(), mask, diag[0]


class PitchFloor:
    def __init__(self, half):
        self.half = half
        self._p = lib().aao_pitch_floor_create(half)

    def update(self, mags, gfloor):
        mags = np.ascontiguousarray(mags, np.float32)
        eff = np.empty(self.half, np.float32)
        lib().aao_pitch_floor_update(self._p, _fp(mags), gfloor, _fp(eff))
        return eff

    def __del__(self):
        if getattr(self, "_p", None):
            lib().aao_pitch_floor_destroy(self._p)
            self._p = None


class Tracker:
    def __init__(self):
        self._p = lib().aao_tracker_create()

    def process(self, pairs, onset=False):
        pairs = np.ascontiguousarray(pairs, np.float32).reshape(-1, 2)
        out = np.zeros((64, 2), np.float32)
        n = lib().aao_tracker_process(self._p, _fp(pairs), pairs.shape[0], int(onset), _fp(out), 64)
        return out[:n].copy()

    def __del__(self):
        if getattr(self, "_p", None):
            lib().aao_tracker_destroy(self._p)
            self._p = None


class Onset:
    def __init__(self, half):
        self.half = half
        self._p = lib().aao_onset_create(half)

    def frame(self, mags, gfloor):
        mags = np.ascontiguousarray(mags, np.float32)
        out = np.zeros(1, FEATURES_DTYPE)
        lib().aao_onset_frame(self._p, _fp(mags), gfloor, _vp(out))
        return out[0]

    def __del__(self):
        if getattr(self, "_p", None):
            lib().aao_onset_destroy(self._p)
            self._p = None


def centroid(mags, bin_width):
    mags = np.ascontiguousarray(mags, np.float32)
    return float(lib().aao_centroid(_fp(mags), mags.shape[0], bin_width))


NOTE_NAMES = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]


def note_from_freq(freq, base_freq=440.0):
    """Note::from_freq (theory.rs:195-209) -> (name like 'A4', octave, semis, cents)."""
    o, s_, c = C.c_int(0), C.c_int(0), C.c_float(0)
    lib().aao_note_from_freq(freq, base_freq, C.byref(o), C.byref(s_), C.byref(c))
    return f"{NOTE_NAMES[s_.value]}{o.value}", o.value, s_.value, c.value


def yin_lag(frame, min_lag, max_lag, threshold=0.1):
    frame = np.ascontiguousarray(frame, np.float32)
    cm = np.zeros(max_lag + 1, np.float64)
    lag = lib().aao_yin_lag(_fp(frame), frame.shape[0], min_lag, max_lag, threshold,
                            cm.ctypes.data_as(C.POINTER(C.c_double)))
    return lag, cm


def num_frames(length, n, hop):
    return int(lib().aao_num_frames(length, n, hop))


def analyze_clip(cfg: Config, samples=None, mags_in=None, onset_in=None, want_mags=True,
                 want_floor=False, want_peaks=False, want_diag=False, want_state=False):
    """Run the frame loop on one clip.  Returns a dict of numpy arrays.
    want_state: also the raw per-bin recurrent state after every frame (pitch_nf, pitch_vol, onset_nf)."""
    half = cfg.n // 2 + 1
    if mags_in is not None:
        mags_in = np.ascontiguousarray(mags_in, np.float32)
        T = mags_in.shape[0]
        length = cfg.n + (T - 1) * cfg.hop
        samples_arr = None
    else:
        samples_arr = np.ascontiguousarray(samples, np.float32)
        length = samples_arr.shape[0]
        T = num_frames(length, cfg.n, cfg.hop)
    out = {"T": T}
    if T <= 0:
        return out
    if onset_in is not None:
        onset_in = np.ascontiguousarray(onset_in, np.uint8)
    mags = np.empty((T, half), np.float32) if want_mags else None
    floor = np.empty((T, half), np.float32) if want_floor else None
    peaks = np.empty((T, half), np.uint8) if want_peaks else None
    feat = np.zeros(T, FEATURES_DTYPE)
    stable = np.zeros(T, STABLE_DTYPE)
    diag = np.zeros(T, DIAG_DTYPE) if want_diag else None
    state = {k: (np.zeros((T, half), np.float32) if want_state else None) for k in ("pitch_nf", "pitch_vol", "onset_nf")}
    taps = Taps()
    for k, a in dict(mags=mags, floor=floor, peak_mask=peaks, features=feat, stable=stable, diag=diag, **state).items():
        setattr(taps, k, a.ctypes.data if a is not None else None)
    got = lib().aao_analyze_clip_ex(C.byref(cfg), _fp(samples_arr), length, _fp(mags_in), _u8p(onset_in),
                                    C.byref(taps))
    assert got == T
    out.update(mags=mags, floor=floor, peaks=peaks, features=feat, stable=stable, diag=diag, **state)
    return out


def analyze_batch(cfg: Config, clips: np.ndarray, n_threads: int, want_mags=False):
    clips = np.ascontiguousarray(clips, np.float32)
    n_clips, clip_len = clips.shape
    T = num_frames(clip_len, cfg.n, cfg.hop)
    half = cfg.n // 2 + 1
    mags = np.empty((n_clips, T, half), np.float32) if want_mags else None
    feat = np.zeros((n_clips, T), FEATURES_DTYPE)
    stable = np.zeros((n_clips, T), STABLE_DTYPE)
    total = lib().aao_analyze_batch(C.byref(cfg), _fp(clips), n_clips, clip_len, n_threads,
                                    _fp(mags), _vp(feat), _vp(stable))
    assert total == T * n_clips
    return {"T": T, "mags": mags, "features": feat, "stable": stable}


def cond_params(sample_rate: float, slot_len: int = 1024) -> CondParams:
    p = CondParams()
    lib().aao_cond_params_init(C.byref(p), float(sample_rate), int(slot_len))
    return p


def condition_clip(samples: np.ndarray, sample_rate: float, slot_len: int = 1024, agc: bool = True):
    """Reference conditioning chain (mod.rs:351-487 + dynamics.rs:194-360) over one clip.
    Returns (conditioned samples, per-slot dynamics records)."""
    x = np.ascontiguousarray(samples, dtype=np.float32).copy()
    p = cond_params(sample_rate, slot_len)
    n_slots = len(x) // slot_len
    dyn = np.zeros(n_slots, dtype=DYNAMICS_DTYPE)
    got = lib().aao_cond_clip(C.byref(p), _fp(x), len(x), _vp(dyn), 1 if agc else 0)
    assert got == n_slots
    return x, dyn


INT_TYPES = ["Min2", "Maj2", "Min3", "Maj3", "Per4", "Aug4", "Per5", "Min6", "Maj6", "Min7", "Maj7", "Per8"]


def interval(f_lo: float, f_hi: float, system: int = 0):
    """Interval::new (theory.rs:306-382): returns (IntType name, accuracy)."""
    acc = C.c_float()
    idx = lib().aao_interval(float(f_lo), float(f_hi), int(system), C.byref(acc))
    return INT_TYPES[idx], acc.value


def tuner_frame(pairs, system: int = 0, single_pitch_mode: bool = False):
    """Tuner::run's per-frame branch (tuner.rs:148-193) -> dict(kind, best, lo, hi, interval, accuracy)."""
    a = np.ascontiguousarray(pairs, np.float32).reshape(-1, 2)
    k, b, lo, hi, iv = (C.c_int() for _ in range(5))
    acc = C.c_float()
    lib().aao_tuner_frame(_fp(a), len(a), int(system), 1 if single_pitch_mode else 0, C.byref(k), C.byref(b),
                          C.byref(lo), C.byref(hi), C.byref(iv), C.byref(acc))
    return dict(kind=k.value, best=b.value, lo=lo.value, hi=hi.value, interval=iv.value, accuracy=acc.value)


PCM_F32, PCM_I16, PCM_U16 = 0, 1, 2
_PCM_DTYPES = {0: np.float32, 1: np.int16, 2: np.uint16}


def ingest(pcm: np.ndarray, fmt: int, channels: int) -> np.ndarray:
    """Input-callback conversion + downmix (mod.rs:765-792, dasp_sample 0.11.0): interleaved PCM -> mono f32."""
    a = np.ascontiguousarray(pcm, _PCM_DTYPES[fmt]).reshape(-1)
    n_frames = a.size // channels
    out = np.zeros(n_frames, np.float32)
    lib().aao_ingest(_vp(a), int(fmt), int(channels), n_frames, _fp(out))
    return out


ONSET_EVENT_DTYPE = np.dtype([("beat_position", "<f8"), ("sample_position", "<i8"), ("frame", "<i8"),
                              ("velocity", "<f4"), ("reserved", "<u4")])
assert ONSET_EVENT_DTYPE.itemsize == 32


def onset_gates():
    """The gates of onset.rs:153 / :79 / :356 / :403 as the oracle applies them."""
    a, b, c = C.c_float(), C.c_float(), C.c_float()
    d, e = C.c_uint32(), C.c_uint32()
    lib().aao_onset_gates(C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e))
    return {"flux_multiplier": a.value, "flux_threshold_floor": b.value, "excess_gate": c.value,
            "count_gate": d.value, "refire_frames": e.value}


def stamp_onset(current_beats, output_frames, bpm, sample_rate, input_lat, output_lat, calibration, sample_offset):
    """MusicalTransport::stamp_onset (timing.rs:311-337): (beat_position f64, output_samples i64)."""
    bp, os_ = C.c_double(0.0), C.c_int64(0)
    lib().aao_stamp_onset(float(current_beats), int(output_frames), float(bpm), float(sample_rate), int(input_lat),
                          int(output_lat), int(calibration), int(sample_offset), C.byref(bp), C.byref(os_))
    return bp.value, os_.value


def onset_events(features: np.ndarray, n: int, hop: int, sample_rate: float, bpm: float = 120.0, max_events: int = 256):
    """Offline onset event list of one clip (onset.rs:383-456 + timing.rs:311-337): (events, total count)."""
    f = np.ascontiguousarray(features, FEATURES_DTYPE)
    out = np.zeros(max_events, ONSET_EVENT_DTYPE)
    cnt = lib().aao_onset_events(_vp(f), len(f), int(n), int(hop), float(sample_rate), float(bpm), int(max_events), _vp(out))
    return out[: min(cnt, max_events)], int(cnt)

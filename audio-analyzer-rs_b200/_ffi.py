"""ctypes binding of libaa_gpu.so (include/aa_gpu.h).

Host-side mirror of the reference's entry points for this path:
  FftProcessor::new / process_forward / process_inverse   src/dsp/fft.rs:14,33,39
  STFT::detect_pitches frame body                         src/audio_io/stft.rs:273-438
  OnsetDetector::detect_onsets frame body                 src/analysis/onset.rs:244-357
No torch types cross the ABI: device buffers are passed as integer addresses
(e.g. torch.Tensor.data_ptr()), streams as the integer cudaStream_t.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libaa_gpu.so")
_HEADER = os.path.join(os.path.dirname(_HERE), "include", "aa_gpu.h")

FEAT_PITCH, FEAT_ONSET, FEAT_CENTROID, FEAT_TRACKER, FEAT_ALL = 1, 2, 4, 8, 15
FLAG_FLUX_ONSET, FLAG_BURST_ONSET, FLAG_ONSET_DETECTED, FLAG_ENERGY_RISING, FLAG_ONSET_FIRED = 1, 2, 4, 8, 16
MAX_NOTES, MAX_STABLE = 8, 16

_PITCH = [("freq", "<f4"), ("score", "<f4")]
FEATURES_DTYPE = np.dtype(
    [
        ("n_pitches", "<u4"),
        ("pitch", _PITCH, (MAX_NOTES,)),
        ("flux", "<f4"),
        ("energy", "<f4"),
        ("centroid", "<f4"),
        ("burst_count", "<u4"),
        ("max_excess", "<f4"),
        ("flags", "<u4"),
        ("energy_ema", "<f4"),
    ]
)
STABLE_DTYPE = np.dtype([("n", "<u4"), ("reserved", "<u4"), ("pitch", _PITCH, (MAX_STABLE,))])
SUMMARY_DTYPE = np.dtype(
    [
        ("n_frames", "<u4"),
        ("n_pitched", "<u4"),
        ("n_onsets", "<u4"),
        ("mean_top_freq", "<f4"),
        ("mean_centroid", "<f4"),
        ("mean_flux", "<f4"),
        ("mean_energy", "<f4"),
        ("max_energy", "<f4"),
    ]
)
NOTE_RECORD_DTYPE = np.dtype(
    [("n", "<u4"), ("reserved", "<u4"),
     ("note", [("semis", "u1"), ("octave", "u1"), ("reserved", "<u2"), ("cents", "<f4")], (MAX_STABLE,))]
)
assert NOTE_RECORD_DTYPE.itemsize == 136
NOTE_NAMES = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]
STREAM_FRAME_DTYPE = np.dtype(
    [("frame_index", "<i8"), ("features", FEATURES_DTYPE), ("stable", STABLE_DTYPE)]
)
assert FEATURES_DTYPE.itemsize == 96 and STABLE_DTYPE.itemsize == 136
assert SUMMARY_DTYPE.itemsize == 32 and STREAM_FRAME_DTYPE.itemsize == 240


class AAError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"aa_gpu error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    """aa_config.  Defaults are the reference's constants (stft.rs:169-174)."""

    _fields_ = [
        ("n", C.c_int32),
        ("hop", C.c_int32),
        ("sample_rate", C.c_float),
        ("min_freq", C.c_float),
        ("max_freq", C.c_float),
        ("noise_floor_db", C.c_float),
        ("features", C.c_uint32),
    ]

    def __init__(self, n=2048, hop=None, sample_rate=44100.0, min_freq=24.0, max_freq=10000.0,
                 noise_floor_db=-96.0, features=FEAT_ALL):
        super().__init__(n, n // 4 if hop is None else hop, sample_rate, min_freq, max_freq,
                         noise_floor_db, features)


COND_AGC, COND_CARRY = 1, 2
PCM_F32, PCM_I16, PCM_U16 = 0, 1, 2
PCM_DTYPES = {0: np.float32, 1: np.int16, 2: np.uint16}


class CondConfig(C.Structure):
    """aa_cond_config (include/aa_gpu.h): the conditioning chain of mod.rs:351-487 + dynamics.rs."""
    _fields_ = [("sample_rate", C.c_float), ("slot_len", C.c_int32), ("flags", C.c_uint32)]


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


class _Outputs(C.Structure):
    _fields_ = [
        ("mags", C.c_void_p),
        ("features", C.c_void_p),
        ("stable", C.c_void_p),
        ("summaries", C.c_void_p),
        ("dbg_floor", C.c_void_p),
        ("dbg_peaks", C.c_void_p),
    ]


_lib = None


def lib_path() -> str:
    return _SO


def lib():
    """dlopen libaa_gpu.so; fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise ImportError(
            f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    L = C.CDLL(_SO)
    vp, i32, i64, f32, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64
    pvp = C.POINTER(C.c_void_p)
    sig = {
        "aa_last_error": (C.c_char_p, []),
        "aa_version": (i32, []),
        "aa_device_count": (i32, [C.POINTER(i32)]),
        "aa_set_device": (i32, [i32]),
        "aa_host_alloc": (i32, [C.c_size_t, pvp]),
        "aa_host_free": (i32, [vp]),
        "aa_device_alloc": (i32, [C.c_size_t, pvp]),
        "aa_device_free": (i32, [vp]),
        "aa_memcpy_h2d": (i32, [vp, vp, C.c_size_t]),
        "aa_memcpy_d2h": (i32, [vp, vp, C.c_size_t]),
        "aa_device_synchronize": (i32, []),
        "aa_fft_create": (i32, [i32, pvp]),
        "aa_fft_destroy": (i32, [vp]),
        "aa_fft_len": (i32, [vp]),
        "aa_fft_forward": (i32, [vp, vp, i64, vp]),
        "aa_fft_forward_device": (i32, [vp, vp, i64, vp, vp]),
        "aa_fft_inverse": (i32, [vp, vp, i64, vp]),
        "aa_fft_inverse_device": (i32, [vp, vp, i64, vp, vp]),
        "aa_config_default_pitch": (None, [C.POINTER(Config), f32]),
        "aa_config_default_onset": (None, [C.POINTER(Config), f32]),
        "aa_analyzer_create": (i32, [C.POINTER(Config), pvp]),
        "aa_analyzer_destroy": (i32, [vp]),
        "aa_num_frames": (i64, [C.POINTER(Config), i64]),
        "aa_plan_segments": (i32, [i64, i64, i32, vp]),
        "aa_analyze_device": (i32, [vp, vp, i64, i64, i64, vp, C.POINTER(_Outputs), vp]),
        "aa_analyze_host": (i32, [vp, vp, i64, i64, i64, vp, C.POINTER(_Outputs)]),
        "aa_state_floats": (i64, [C.POINTER(Config)]),
        "aa_analyze_device_carry": (i32, [vp, vp, i64, i64, i64, vp, C.POINTER(_Outputs), vp, vp]),
        "aa_analyzer_last_launches": (i64, [vp]),
        "aa_stream_create": (i32, [C.POINTER(Config), pvp]),
        "aa_stream_destroy": (i32, [vp]),
        "aa_stream_push": (i32, [vp, vp, i32]),
        "aa_stream_set_noise_floor_db": (i32, [vp, f32]),
        "aa_stream_signal_onset": (i32, [vp]),
        "aa_stream_poll": (i32, [vp, vp, i32, C.POINTER(i32)]),
        "aa_stream_reset": (i32, [vp]),
        "aa_stream_probe_latency": (i32, [vp, vp, i32, i32, vp, C.POINTER(i64)]),
        "aa_synth_clips_device": (i32, [vp, i64, i64, i64, f32, u64, vp]),
        "aa_yin_device": (i32, [vp, vp, i64, i64, i64, vp, vp, vp]),
        "aa_yin_host": (i32, [vp, vp, i64, i64, i64, vp, vp]),
        "aa_notes_from_stable_device": (i32, [vp, i64, f32, vp, vp]),
        "aa_notes_from_stable_host": (i32, [vp, i64, f32, vp]),
        "aa_ingest_device": (i32, [vp, i32, i32, i64, i64, i64, i64, vp, vp]),
        "aa_analyze_host_pcm": (i32, [vp, vp, i32, i32, i64, i64, i64, vp, C.POINTER(_Outputs)]),
        "aa_onset_events_device": (i32, [C.POINTER(Config), vp, i64, i64, f32, i32, vp, vp, vp]),
        "aa_onset_events_host": (i32, [C.POINTER(Config), vp, i64, i64, f32, i32, vp, vp]),
        "aa_tuner_from_stable_device": (i32, [vp, i64, f32, i32, i32, vp, vp]),
        "aa_tuner_from_stable_host": (i32, [vp, i64, f32, i32, i32, vp]),
        "aa_conditioner_create": (i32, [C.POINTER(CondConfig), pvp]),
        "aa_conditioner_destroy": (i32, [vp]),
        "aa_conditioner_reset": (i32, [vp]),
        "aa_cond_num_slots": (i64, [C.POINTER(CondConfig), i64]),
        "aa_condition_device": (i32, [vp, vp, i64, i64, i64, vp, vp]),
        "aa_condition_host": (i32, [vp, vp, i64, i64, i64, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def header_symbols() -> list[str]:
    """Every function include/aa_gpu.h declares with AA_API."""
    text = open(_HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"AA_API\s+[^;(]*?\b(aa_[a-z0-9_]+)\s*\(", text)))


def exported_symbols() -> list[str]:
    L = lib()
    return [s for s in header_symbols() if hasattr(L, s)]


def _check(code):
    if code != 0:
        raise AAError(code, lib().aa_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_int32(0)
    _check(lib().aa_device_count(C.byref(n)))
    return n.value


def set_device(i: int):
    _check(lib().aa_set_device(i))


def num_frames(cfg: Config, clip_len: int) -> int:
    return int(lib().aa_num_frames(C.byref(cfg), clip_len))


def plan_segments(T: int, n_clips: int, resident_ctas: int) -> list[int]:
    """Frame boundaries of the time segments the batch path cuts every clip into (aa_plan_segments)."""
    starts = (C.c_int32 * 9)()
    n = int(lib().aa_plan_segments(T, n_clips, resident_ctas, starts))
    return [int(starts[i]) for i in range(n + 1)]


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over pinned host memory from aa_host_alloc (freed with the array)."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    _check(lib().aa_host_alloc(max(nbytes, 1), C.byref(p)))
    addr = p.value
    buf = (C.c_char * max(nbytes, 1)).from_address(addr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, lambda a=addr: lib().aa_host_free(C.c_void_p(a)))
    return arr


class FftProcessor:
    """FftProcessor (reference src/dsp/fft.rs:6-41) on the GPU."""

    def __init__(self, n: int):
        self.n = n
        h = C.c_void_p()
        _check(lib().aa_fft_create(n, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().aa_fft_destroy(self._h)
            self._h = None

    __del__ = close

    def process_forward(self, windowed: np.ndarray) -> np.ndarray:
        """fft.rs:33.  [n] or [batch, n] float32 -> complex64 [.., n/2+1].  A wrong
        length raises (the reference panics through unwrap(), fft.rs:69)."""
        x = np.ascontiguousarray(windowed, np.float32)
        if x.shape[-1] != self.n:
            raise AAError(-1, f"process_forward: expected length {self.n}, got {x.shape[-1]}")
        batch = int(np.prod(x.shape[:-1])) if x.ndim > 1 else 1
        out = np.empty(x.shape[:-1] + (self.n // 2 + 1,), np.complex64)
        _check(lib().aa_fft_forward(self._h, _ptr(x), batch, _ptr(out)))
        return out

    def process_inverse(self, spectrum: np.ndarray) -> np.ndarray:
        """fft.rs:39.  complex64 [.., n/2+1] -> float32 [.., n], unnormalised."""
        s = np.ascontiguousarray(spectrum, np.complex64)
        if s.shape[-1] != self.n // 2 + 1:
            raise AAError(-1, f"process_inverse: expected {self.n // 2 + 1} bins, got {s.shape[-1]}")
        batch = int(np.prod(s.shape[:-1])) if s.ndim > 1 else 1
        out = np.empty(s.shape[:-1] + (self.n,), np.float32)
        _check(lib().aa_fft_inverse(self._h, _ptr(s), batch, _ptr(out)))
        return out

    def forward_device(self, in_ptr: int, batch: int, out_ptr: int, stream: int = 0):
        _check(lib().aa_fft_forward_device(self._h, C.c_void_p(in_ptr), batch, C.c_void_p(out_ptr),
                                           C.c_void_p(stream)))

    def inverse_device(self, spec_ptr: int, batch: int, out_ptr: int, stream: int = 0):
        _check(lib().aa_fft_inverse_device(self._h, C.c_void_p(spec_ptr), batch, C.c_void_p(out_ptr),
                                           C.c_void_p(stream)))


class Analyzer:
    """Batch frame analysis of offline clips (the STFT / OnsetDetector frame loops)."""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        h = C.c_void_p()
        _check(lib().aa_analyzer_create(C.byref(cfg), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().aa_analyzer_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def half(self):
        return self.cfg.n // 2 + 1

    def num_frames(self, clip_len: int) -> int:
        return num_frames(self.cfg, clip_len)

    @property
    def last_launches(self) -> int:
        return int(lib().aa_analyzer_last_launches(self._h))

    def analyze_device(self, clips_ptr: int, n_clips: int, clip_len: int, clip_stride: int | None = None,
                       mags: int = 0, features: int = 0, stable: int = 0, summaries: int = 0,
                       dbg_floor: int = 0, dbg_peaks: int = 0, onset_in: int = 0, stream: int = 0):
        """Asynchronous on `stream`; all pointers are device addresses (0 = not produced)."""
        o = _Outputs(mags or None, features or None, stable or None, summaries or None,
                     dbg_floor or None, dbg_peaks or None)
        _check(lib().aa_analyze_device(self._h, C.c_void_p(clips_ptr), n_clips, clip_len,
                                       clip_len if clip_stride is None else clip_stride,
                                       C.c_void_p(onset_in) if onset_in else None, C.byref(o),
                                       C.c_void_p(stream) if stream else None))

    @property
    def state_floats(self) -> int:
        """Floats in one analyzer state block (aa_state_floats)."""
        return int(lib().aa_state_floats(C.byref(self.cfg)))

    def analyze_device_carry(self, clips_ptr: int, n_clips: int, clip_len: int, clip_stride: int, state_ptr: int,
                             mags: int = 0, features: int = 0, stable: int = 0, onset_in: int = 0, stream: int = 0):
        """aa_analyze_device_carry: clip c starts from, and leaves its final analyzer state in, the state block
        state_ptr + c * state_floats * 4 (device memory; zeros = a fresh analyzer)."""
        o = _Outputs(mags or None, features or None, stable or None, None, None, None)
        _check(lib().aa_analyze_device_carry(self._h, C.c_void_p(clips_ptr), n_clips, clip_len, clip_stride,
                                             C.c_void_p(onset_in) if onset_in else None, C.byref(o),
                                             C.c_void_p(state_ptr), C.c_void_p(stream) if stream else None))

    def analyze_host_into(self, clips: np.ndarray, n_clips: int, clip_len: int, clip_stride: int,
                          mags=None, features=None, stable=None, summaries=None, dbg_floor=None,
                          dbg_peaks=None, onset_in=None):
        """aa_analyze_host on caller-provided (ideally pinned) host arrays."""
        o = _Outputs(_ptr(mags), _ptr(features), _ptr(stable), _ptr(summaries), _ptr(dbg_floor),
                     _ptr(dbg_peaks))
        _check(lib().aa_analyze_host(self._h, _ptr(clips), n_clips, clip_len, clip_stride, _ptr(onset_in),
                                     C.byref(o)))

    def analyze_host_pcm(self, pcm: np.ndarray, fmt: int, channels: int, want_mags=True, want_stable=True,
                         want_summaries=True):
        """aa_analyze_host_pcm: pcm [n_clips, clip_len * channels] interleaved (float32 / int16 / uint16 per fmt);
        clip_len (frames) must be a multiple of 4 here because the rows are contiguous."""
        pcm = np.ascontiguousarray(pcm, PCM_DTYPES.get(fmt, np.int16))     # an unknown format is rejected by the ABI
        n_clips, row = pcm.shape
        clen = row // channels
        if row % channels or clen % 4:
            raise AAError(-1, "row length must be channels * clip_len with clip_len % 4 == 0")
        T = self.num_frames(clen)
        res = {"T": T}
        res["features"] = np.zeros((n_clips, T), FEATURES_DTYPE)
        res["mags"] = np.zeros((n_clips, T, self.half), np.float32) if want_mags else None
        res["stable"] = np.zeros((n_clips, T), STABLE_DTYPE) if want_stable else None
        res["summaries"] = np.zeros(n_clips, SUMMARY_DTYPE) if want_summaries else None
        o = _Outputs(_ptr(res["mags"]), _ptr(res["features"]), _ptr(res["stable"]), _ptr(res["summaries"]), None, None)
        _check(lib().aa_analyze_host_pcm(self._h, _ptr(pcm), int(fmt), int(channels), n_clips, clen, clen, None,
                                         C.byref(o)))
        return res

    def analyze_host_pcm_into(self, pcm: np.ndarray, fmt: int, channels: int, n_clips: int, clip_len: int,
                              clip_stride: int, features=None, stable=None, summaries=None):
        """aa_analyze_host_pcm on caller-provided (ideally pinned) arrays."""
        o = _Outputs(None, _ptr(features), _ptr(stable), _ptr(summaries), None, None)
        _check(lib().aa_analyze_host_pcm(self._h, _ptr(pcm), int(fmt), int(channels), n_clips, clip_len, clip_stride,
                                         None, C.byref(o)))

    def analyze_host(self, clips: np.ndarray, want_mags=True, want_stable=True, want_summaries=True,
                     want_dbg=False, onset_in=None, clip_stride=None, clip_len=None):
        """clips: float32 [n_clips, clip_len] (or a 1-D stream with clip_stride / clip_len given,
        which is how hop-aligned chunks with a window halo are expressed)."""
        clips = np.ascontiguousarray(clips, np.float32)
        if clips.ndim == 2:
            n_clips, clen = clips.shape
            stride = clen
        else:
            clen, stride = int(clip_len), int(clip_stride)
            n_clips = (clips.shape[0] - clen) // stride + 1
        if stride % 4:
            if clips.ndim != 2:
                raise AAError(-1, "clip_stride must be a multiple of 4 samples")
            # the ABI wants a stride that is a multiple of 4 samples: re-pack with padding
            pad = (-stride) % 4
            packed = np.zeros((n_clips, stride + pad), np.float32)
            packed[:, :clen] = clips
            clips, stride = packed, stride + pad
        T = self.num_frames(clen)
        half = self.half
        res = {"T": T}
        res["features"] = np.zeros((n_clips, T), FEATURES_DTYPE)
        res["mags"] = np.zeros((n_clips, T, half), np.float32) if want_mags else None
        res["stable"] = np.zeros((n_clips, T), STABLE_DTYPE) if want_stable else None
        res["summaries"] = np.zeros(n_clips, SUMMARY_DTYPE) if want_summaries else None
        res["dbg_floor"] = np.zeros((n_clips, T, half), np.float32) if want_dbg else None
        res["dbg_peaks"] = np.zeros((n_clips, T, half), np.uint8) if want_dbg else None
        if onset_in is not None:
            onset_in = np.ascontiguousarray(onset_in, np.uint8)
        self.analyze_host_into(clips, n_clips, clen, stride, res["mags"], res["features"], res["stable"],
                               res["summaries"], res["dbg_floor"], res["dbg_peaks"], onset_in)
        return res


class Stream:
    """Streaming analyzer: the body of the STFT / OnsetDetector worker threads
    (stft.rs:240-438, onset.rs:216-543) as push / poll."""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        h = C.c_void_p()
        _check(lib().aa_stream_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._buf = np.zeros(256, STREAM_FRAME_DTYPE)

    def close(self):
        if getattr(self, "_h", None):
            lib().aa_stream_destroy(self._h)
            self._h = None

    __del__ = close

    def push(self, samples: np.ndarray):
        s = np.ascontiguousarray(samples, np.float32)
        _check(lib().aa_stream_push(self._h, _ptr(s), s.shape[0]))

    def poll(self, max_frames: int = 256) -> np.ndarray:
        if max_frames > self._buf.shape[0]:
            self._buf = np.zeros(max_frames, STREAM_FRAME_DTYPE)
        n = C.c_int32(0)
        _check(lib().aa_stream_poll(self._h, _ptr(self._buf), max_frames, C.byref(n)))
        return self._buf[: n.value].copy()

    def probe_latency(self, samples: np.ndarray, count: int):
        """aa_stream_probe_latency: push `count` samples at a time + poll, timed inside the library (no Python in
        the loop); returns (latency_us per push, frames polled)."""
        x = np.ascontiguousarray(samples, np.float32)
        n_pushes = x.shape[0] // count
        lat = np.zeros(n_pushes, np.float64)
        frames = C.c_int64(0)
        _check(lib().aa_stream_probe_latency(self._h, _ptr(x), int(count), int(n_pushes), _ptr(lat), C.byref(frames)))
        return lat, int(frames.value)

    def set_noise_floor_db(self, db: float):
        _check(lib().aa_stream_set_noise_floor_db(self._h, db))

    def signal_onset(self):
        _check(lib().aa_stream_signal_onset(self._h))

    def reset(self):
        _check(lib().aa_stream_reset(self._h))


def synth_clips_device(ptr: int, n_clips: int, clip_len: int, clip_stride: int, sample_rate: float,
                       seed: int, stream: int = 0):
    _check(lib().aa_synth_clips_device(C.c_void_p(ptr), n_clips, clip_len, clip_stride, sample_rate, seed,
                                       C.c_void_p(stream) if stream else None))


def notes_from_stable(stable: np.ndarray, base_freq: float = 440.0) -> np.ndarray:
    """Note::from_freq (theory.rs:195-209) for every stable pitch: records parallel to `stable`."""
    st = np.ascontiguousarray(stable, STABLE_DTYPE)
    out = np.zeros(st.shape, NOTE_RECORD_DTYPE)
    _check(lib().aa_notes_from_stable_host(_ptr(st), st.size, base_freq, _ptr(out)))
    return out


def notes_from_stable_device(stable_ptr: int, n_frames: int, base_freq: float, out_ptr: int, stream: int = 0):
    _check(lib().aa_notes_from_stable_device(C.c_void_p(stable_ptr), n_frames, base_freq, C.c_void_p(out_ptr),
                                             C.c_void_p(stream) if stream else None))


class YinConfig(C.Structure):
    _fields_ = [("n", C.c_int32), ("hop", C.c_int32), ("min_lag", C.c_int32), ("max_lag", C.c_int32),
                ("threshold", C.c_float)]


def yin_host(clips: np.ndarray, n: int, hop: int, min_lag: int, max_lag: int, threshold: float = 0.1):
    """YIN-style lag search per frame (a14, NEW): returns (lag int32 [n_clips, T], d'(lag) float32)."""
    clips = np.ascontiguousarray(np.atleast_2d(clips), np.float32)
    n_clips, clip_len = clips.shape
    T = 0 if clip_len < n else (clip_len - n) // hop + 1
    lag = np.zeros((n_clips, T), np.int32)
    cm = np.zeros((n_clips, T), np.float32)
    cfg = YinConfig(n, hop, min_lag, max_lag, threshold)
    _check(lib().aa_yin_host(C.byref(cfg), _ptr(clips), n_clips, clip_len, clip_len, _ptr(lag), _ptr(cm)))
    return lag, cm


def yin_device(cfg: YinConfig, clips_ptr: int, n_clips: int, clip_len: int, clip_stride: int, lag_ptr: int,
               cmnd_ptr: int = 0, stream: int = 0):
    _check(lib().aa_yin_device(C.byref(cfg), C.c_void_p(clips_ptr), n_clips, clip_len, clip_stride,
                               C.c_void_p(lag_ptr), C.c_void_p(cmnd_ptr) if cmnd_ptr else None,
                               C.c_void_p(stream) if stream else None))


class Conditioner:
    """Mirror of the reducer thread's per-slot work (mod.rs:431-490): filters, gate, DynamicsTracker."""

    def __init__(self, sample_rate: float, slot_len: int = 1024, agc: bool = True, carry: bool = False):
        self.cfg = CondConfig(float(sample_rate), int(slot_len), (COND_AGC if agc else 0) | (COND_CARRY if carry else 0))
        h = C.c_void_p()
        _check(lib().aa_conditioner_create(C.byref(self.cfg), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().aa_conditioner_destroy(self._h)
            self._h = None

    __del__ = close

    def num_slots(self, clip_len: int) -> int:
        return int(lib().aa_cond_num_slots(C.byref(self.cfg), int(clip_len)))

    def reset(self):
        _check(lib().aa_conditioner_reset(self._h))

    def process_host(self, clips: np.ndarray):
        """Condition [n_clips, clip_len] f32 clips; returns (conditioned copy, dynamics [n_clips, n_slots])."""
        x = np.ascontiguousarray(np.atleast_2d(clips), np.float32).copy()
        n_clips, clip_len = x.shape
        if clip_len % 4:
            raise ValueError("clip_len must be a multiple of 4 for contiguous clips")
        dyn = np.zeros((n_clips, self.num_slots(clip_len)), DYNAMICS_DTYPE)
        _check(lib().aa_condition_host(self._h, _ptr(x), n_clips, clip_len, clip_len,
                                       _ptr(dyn) if (self.cfg.flags & COND_AGC) else None))
        return x, dyn

    def process_device(self, clips_ptr: int, n_clips: int, clip_len: int, clip_stride: int, dyn_ptr: int = 0,
                       stream: int = 0):
        _check(lib().aa_condition_device(self._h, C.c_void_p(clips_ptr), n_clips, clip_len, clip_stride,
                                         C.c_void_p(dyn_ptr) if dyn_ptr else None,
                                         C.c_void_p(stream) if stream else None))


TUNER_RECORD_DTYPE = np.dtype(
    [("kind", "u1"), ("best", "u1"), ("lo", "u1"), ("hi", "u1"), ("interval", "<u4"), ("accuracy", "<f4"), ("cents", "<f4")]
)
assert TUNER_RECORD_DTYPE.itemsize == 16
INT_TYPES = ["Min2", "Maj2", "Min3", "Maj3", "Per4", "Aug4", "Per5", "Min6", "Maj6", "Min7", "Maj7", "Per8"]
SYSTEM_EQUAL, SYSTEM_JUST, SYSTEM_PYTHAGOREAN = 0, 1, 2


def tuner_from_stable(stable: np.ndarray, base_freq: float = 440.0, system: int = 0,
                      single_pitch_mode: bool = False) -> np.ndarray:
    """Tuner::run's per-frame branch + Interval::new (tuner.rs:148-193, theory.rs:306-382) on stable pitches."""
    st = np.ascontiguousarray(stable, STABLE_DTYPE).reshape(-1)
    out = np.zeros(st.shape[0], TUNER_RECORD_DTYPE)
    _check(lib().aa_tuner_from_stable_host(_ptr(st), st.shape[0], float(base_freq), int(system),
                                           1 if single_pitch_mode else 0, _ptr(out)))
    return out.reshape(np.shape(stable))


def ingest_device(pcm_ptr: int, fmt: int, channels: int, n_clips: int, clip_len: int, in_stride: int, out_stride: int,
                  out_ptr: int, stream: int = 0):
    """aa_ingest_device: interleaved PCM (f32 / i16 / u16) -> mono f32 clips (mod.rs:765-792)."""
    _check(lib().aa_ingest_device(C.c_void_p(pcm_ptr), int(fmt), int(channels), n_clips, clip_len, in_stride, out_stride,
                                  C.c_void_p(out_ptr), C.c_void_p(stream) if stream else None))


ONSET_EVENT_DTYPE = np.dtype([("beat_position", "<f8"), ("sample_position", "<i8"), ("frame", "<i8"),
                              ("velocity", "<f4"), ("reserved", "<u4")])
assert ONSET_EVENT_DTYPE.itemsize == 32


def onset_events(cfg: Config, features: np.ndarray, clip_len: int, bpm: float = 120.0, max_events: int = 256):
    """Offline OnsetEvent lists (onset.rs:383-456 + timing.rs:311-337) from feature records [n_clips, T]:
    returns (events [n_clips, max_events], counts [n_clips])."""
    f = np.ascontiguousarray(np.atleast_2d(features), FEATURES_DTYPE)
    n_clips = f.shape[0]
    ev = np.zeros((n_clips, max_events), ONSET_EVENT_DTYPE)
    cnt = np.zeros(n_clips, np.int32)
    _check(lib().aa_onset_events_host(C.byref(cfg), _ptr(f), n_clips, int(clip_len), float(bpm), int(max_events),
                                      _ptr(ev), _ptr(cnt)))
    return ev, cnt

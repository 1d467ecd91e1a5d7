"""audio-analyzer-rs_b200 -- B200 (sm_100a) frame-analysis path of audio-analyzer-rs.

The product is libaa_gpu.so (CUDA kernels + C ABI, include/aa_gpu.h).  This package
is the thin Python host binding used by the tests and bench.py; the C++ mirror of the
reference's Rust types lives in host/audio_engine_gpu.hpp.

The directory name contains '-', so import it with
    importlib.import_module("audio-analyzer-rs_b200")
There is no CPU fallback: loading fails loudly if libaa_gpu.so is missing, and every
create call fails with AA_ERR_NO_DEVICE when no sm_100 GPU is visible.
"""
from ._ffi import (  # noqa: F401
    AAError,
    Analyzer,
    COND_AGC,
    COND_CARRY,
    CondConfig,
    Conditioner,
    Config,
    DYNAMICS_DTYPE,
    FEAT_ALL,
    FEAT_CENTROID,
    FEAT_ONSET,
    FEAT_PITCH,
    FEAT_TRACKER,
    FEATURES_DTYPE,
    FLAG_BURST_ONSET,
    FLAG_ENERGY_RISING,
    FLAG_FLUX_ONSET,
    FLAG_ONSET_FIRED,
    NOTE_NAMES,
    ONSET_EVENT_DTYPE,
    onset_events,
    PCM_F32,
    PCM_I16,
    PCM_U16,
    ingest_device,
    NOTE_RECORD_DTYPE,
    FLAG_ONSET_DETECTED,
    FftProcessor,
    STABLE_DTYPE,
    STREAM_FRAME_DTYPE,
    SUMMARY_DTYPE,
    Stream,
    TUNER_RECORD_DTYPE,
    INT_TYPES,
    tuner_from_stable,
    YinConfig,
    device_count,
    exported_symbols,
    header_symbols,
    lib,
    lib_path,
    notes_from_stable,
    notes_from_stable_device,
    num_frames,
    plan_segments,
    pinned_empty,
    set_device,
    synth_clips_device,
    yin_device,
    yin_host,
)
from .build import build as build_native  # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]

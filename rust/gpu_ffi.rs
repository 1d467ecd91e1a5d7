//! Raw binding of `include/aa_gpu.h` (libaa_gpu.so, sm_100a).  UNCOMPILED SOURCE -- see rust/README.md.
//!
//! Every struct is a `#[repr(C)]` mirror of the header's; sizes are asserted at compile time so a header
//! change cannot go unnoticed.  All calls return `aa_status` (0 = AA_OK, negative = error, message in
//! `aa_last_error()`); nothing here panics or aborts on the C side.
#![allow(non_camel_case_types, dead_code)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_void};

pub const AA_OK: i32 = 0;
pub const AA_FEAT_PITCH: u32 = 1;
pub const AA_FEAT_ONSET: u32 = 2;
pub const AA_FEAT_CENTROID: u32 = 4;
pub const AA_FEAT_TRACKER: u32 = 8;

pub const AA_FLAG_FLUX_ONSET: u32 = 1; // FluxTracker::update() returned true   onset.rs:355
pub const AA_FLAG_BURST_ONSET: u32 = 2; // max_excess > 3 && count >= 3          onset.rs:356
pub const AA_FLAG_ONSET_DETECTED: u32 = 4; // both                               onset.rs:357
pub const AA_FLAG_ENERGY_RISING: u32 = 8; // energy > ema * 1.5                  onset.rs:373
pub const AA_FLAG_ONSET_FIRED: u32 = 16; // offline gating only (no transport)

pub const AA_MAX_NOTES: usize = 8; // stft.rs:452
pub const AA_MAX_STABLE: usize = 16;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct aa_config {
    pub n: i32,
    pub hop: i32,
    pub sample_rate: f32,
    pub min_freq: f32,
    pub max_freq: f32,
    pub noise_floor_db: f32,
    pub features: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct aa_pitch {
    pub freq: f32,
    pub score: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct aa_frame_features {
    pub n_pitches: u32,
    pub pitch: [aa_pitch; AA_MAX_NOTES],
    pub flux: f32,
    pub energy: f32,
    pub centroid: f32,
    pub burst_count: u32,
    pub max_excess: f32,
    pub flags: u32,
    pub energy_ema: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct aa_stable_pitches {
    pub n: u32,
    pub reserved: u32,
    pub pitch: [aa_pitch; AA_MAX_STABLE],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct aa_stream_frame {
    pub frame_index: i64,
    pub features: aa_frame_features,
    pub stable: aa_stable_pitches,
}

const _: () = assert!(std::mem::size_of::<aa_config>() == 28);
const _: () = assert!(std::mem::size_of::<aa_frame_features>() == 96);
const _: () = assert!(std::mem::size_of::<aa_stable_pitches>() == 136);
const _: () = assert!(std::mem::size_of::<aa_stream_frame>() == 240);

#[repr(C)]
pub struct aa_fft {
    _private: [u8; 0],
}
#[repr(C)]
pub struct aa_stream {
    _private: [u8; 0],
}

extern "C" {
    pub fn aa_last_error() -> *const c_char;
    pub fn aa_version() -> i32;
    pub fn aa_device_count(count: *mut i32) -> i32;
    pub fn aa_set_device(device: i32) -> i32;

    // FftProcessor (src/dsp/fft.rs:14, 33, 39)
    pub fn aa_fft_create(n: i32, out: *mut *mut aa_fft) -> i32;
    pub fn aa_fft_destroy(h: *mut aa_fft) -> i32;
    pub fn aa_fft_len(h: *const aa_fft) -> i32;
    pub fn aa_fft_forward(h: *mut aa_fft, in_host: *const f32, batch: i64, out_host: *mut f32) -> i32;
    pub fn aa_fft_inverse(h: *mut aa_fft, spec_host: *const f32, batch: i64, out_host: *mut f32) -> i32;

    // analyzer worker bodies (stft.rs:176-440, onset.rs:138-545) as push / poll
    pub fn aa_config_default_pitch(cfg: *mut aa_config, sample_rate: f32);
    pub fn aa_config_default_onset(cfg: *mut aa_config, sample_rate: f32);
    pub fn aa_stream_create(cfg: *const aa_config, out: *mut *mut aa_stream) -> i32;
    pub fn aa_stream_destroy(h: *mut aa_stream) -> i32;
    pub fn aa_stream_reset(h: *mut aa_stream) -> i32;
    pub fn aa_stream_push(h: *mut aa_stream, samples: *const f32, count: i32) -> i32;
    pub fn aa_stream_set_noise_floor_db(h: *mut aa_stream, db: f32) -> i32;
    pub fn aa_stream_signal_onset(h: *mut aa_stream) -> i32;
    pub fn aa_stream_poll(h: *mut aa_stream, out: *mut aa_stream_frame, max: i32, n_out: *mut i32) -> i32;
}

/// `aa_last_error()` as an owned string.
pub fn last_error() -> String {
    unsafe {
        let p = aa_last_error();
        if p.is_null() {
            String::new()
        } else {
            CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    }
}

/// Status code -> `Result`, carrying the library's message.
pub fn check(status: i32) -> Result<(), String> {
    if status == AA_OK {
        Ok(())
    } else {
        Err(format!("libaa_gpu status {status}: {}", last_error()))
    }
}

/// Keeps `*mut c_void`-style handles out of signatures above this module.
pub type RawHandle = *mut c_void;

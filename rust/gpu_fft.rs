//! Drop-in for `src/dsp/fft.rs`: the same type name and the same three methods, forwarding to the B200 through
//! `aa_fft_*`.  UNCOMPILED SOURCE -- see rust/README.md.
//!
//! Differences a caller can observe (all documented in include/aa_gpu.h):
//!   * lengths: 256 / 512 / 1024 / 2048 / 4096 only (the reference plans any length); `new` panics otherwise,
//!     like the reference's planner would on an internal error -- use `try_new` to get the message instead;
//!   * `process_forward` does NOT clobber its input (realfft uses it as scratch, fft.rs:66-71);
//!   * one call = one H2D + kernel + D2H round trip (~tens of microseconds): use the stream / batch API for
//!     throughput, this type exists so that code written against `FftProcessor` keeps compiling.
use rustfft::num_complex::Complex;

use super::gpu_ffi as ffi;

pub struct FftProcessor {
    h: *mut ffi::aa_fft,
    len: usize,
    spectrum: Vec<Complex<f32>>, // borrowed out by process_forward, like FftForward::spectrum (fft.rs:48)
    output: Vec<f32>,            // borrowed out by process_inverse, like FftInverse::output (fft.rs:78)
}

// one instance per worker thread, as in the reference (stft.rs:180, onset.rs:141); the handle owns a CUDA stream
unsafe impl Send for FftProcessor {}

impl FftProcessor {
    /// `FftProcessor::new(len)` (fft.rs:14).
    pub fn new(len: usize) -> Self {
        Self::try_new(len).expect("aa_fft_create failed")
    }

    pub fn try_new(len: usize) -> Result<Self, String> {
        let mut h = std::ptr::null_mut();
        ffi::check(unsafe { ffi::aa_fft_create(len as i32, &mut h) })?;
        Ok(Self { h, len, spectrum: vec![Complex::new(0.0, 0.0); len / 2 + 1], output: vec![0.0; len] })
    }

    /// `process_forward(&mut self, windowed) -> &[Complex<f32>]` (fft.rs:33): unnormalised, `len/2 + 1` bins,
    /// `X[0]` and `X[len/2]` purely real.  Panics on a wrong length like the reference's `.unwrap()` (fft.rs:69).
    pub fn process_forward(&mut self, windowed: &mut [f32]) -> &[Complex<f32>] {
        assert_eq!(windowed.len(), self.len, "FftProcessor::process_forward: wrong input length");
        let out = self.spectrum.as_mut_ptr() as *mut f32; // Complex<f32> is repr(C) { re, im }
        ffi::check(unsafe { ffi::aa_fft_forward(self.h, windowed.as_ptr(), 1, out) }).unwrap();
        &self.spectrum
    }

    /// `process_inverse(&mut self, spectrum) -> &[f32]` (fft.rs:39; never called by the reference): unnormalised.
    pub fn process_inverse(&mut self, windowed: &mut [Complex<f32>]) -> &[f32] {
        assert_eq!(windowed.len(), self.len / 2 + 1, "FftProcessor::process_inverse: wrong input length");
        let inp = windowed.as_ptr() as *const f32;
        ffi::check(unsafe { ffi::aa_fft_inverse(self.h, inp, 1, self.output.as_mut_ptr()) }).unwrap();
        &self.output
    }
}

impl Drop for FftProcessor {
    fn drop(&mut self) {
        unsafe {
            ffi::aa_fft_destroy(self.h);
        }
    }
}

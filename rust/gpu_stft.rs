//! Drop-in for `src/audio_io/stft.rs`: `STFT` with the reference's constructor, `detect_pitches` signature,
//! `stop / pause / resume` and `Drop`, whose worker thread forwards the frame loop (stft.rs:273-438: window, FFT,
//! magnitudes, adaptive per-bin floor, extract_pitches, PitchTracker) to the B200 through `aa_stream_*`.
//! UNCOMPILED SOURCE -- see rust/README.md.
//!
//! What stays in Rust, unchanged in behaviour: the AtomicI8 state machine (-1 stop / 0 pause / 1 run,
//! stft.rs:127-135), draining and releasing slots while paused (stft.rs:227-236), the slot release / reclaim
//! protocol (stft.rs:256-258), reading `noise_floor_db` from the shared DynamicsOutput (stft.rs:322), consuming
//! `onset_pending` (stft.rs:387) and pushing `(Vec<(f32, f32)>, beat)` for non-empty frames (stft.rs:431-434).
//! What cannot happen any more: the ring overrun (stft.rs:262-266) -- every slot is analysed before the next pop.
use std::{
    sync::{
        Arc,
        atomic::{AtomicBool, AtomicI8, Ordering},
    },
    thread,
    time::Duration,
};

use crossbeam_channel::Sender;
use rtrb::Producer;

use crate::{
    audio_io::{SlotPool, dynamics::DynamicsOutput, timing::MusicalTransport},
    dsp::gpu_ffi as ffi,
};

/// Short-time Fourier transform worker for pitch detection, GPU edition.
pub struct STFT {
    state: Arc<AtomicI8>,
    handle: u8,
    reducer_remove_tx: Sender<u8>,
}

impl STFT {
    pub fn stop(&mut self) {
        self.state.store(-1, Ordering::Relaxed);
    }
    pub fn pause(&mut self) {
        self.state.store(0, Ordering::Relaxed);
    }
    pub fn resume(&mut self) {
        self.state.store(1, Ordering::Relaxed);
    }
}

impl Drop for STFT {
    fn drop(&mut self) {
        let _ = self.reducer_remove_tx.send(self.handle); // stft.rs:141
        self.stop();
    }
}

/// Owns the `aa_stream` inside the worker thread (the handle is single-threaded, like `&mut FftProcessor`).
struct GpuStream(*mut ffi::aa_stream);
unsafe impl Send for GpuStream {}
impl Drop for GpuStream {
    fn drop(&mut self) {
        unsafe {
            ffi::aa_stream_destroy(self.0);
        }
    }
}

impl STFT {
    pub fn new(handle: u8, reducer_remove_tx: Sender<u8>) -> Self {
        STFT { handle, reducer_remove_tx, state: Arc::new(AtomicI8::new(0)) }
    }

    /// Same arguments as the reference (stft.rs:155-165).  GPU initialisation happens on the calling thread so a
    /// failure can be reported: `Err(msg)` maps to `AudioEngineError::SpawnFailed { component: "tuner", msg }`.
    pub fn try_detect_pitches(
        &mut self,
        slots: Arc<SlotPool>,
        mut cons: rtrb::Consumer<usize>,
        reclaim: Sender<usize>,
        sr: u32,
        note_tx: Producer<(Vec<(f32, f32)>, f64)>,
        dynamics_output: Arc<parking_lot::RwLock<DynamicsOutput>>,
        transport: Arc<MusicalTransport>,
        onset_pending: Arc<AtomicBool>,
    ) -> Result<(), String> {
        let mut cfg = unsafe { std::mem::zeroed::<ffi::aa_config>() };
        unsafe { ffi::aa_config_default_pitch(&mut cfg, sr as f32) }; // 2048 / 512, 24 Hz .. 10 kHz (stft.rs:169-174)
        let mut raw = std::ptr::null_mut();
        ffi::check(unsafe { ffi::aa_stream_create(&cfg, &mut raw) })?;
        let gpu = GpuStream(raw);

        self.state.store(1, Ordering::Relaxed);
        let state = self.state.clone();

        thread::spawn(move || {
            let gpu = gpu;
            let mut note_tx = note_tx;
            let mut frames: Vec<ffi::aa_stream_frame> = Vec::with_capacity(32);
            unsafe { frames.set_len(32) }; // plain-old-data records, filled by aa_stream_poll before they are read
            let mut last_db = f32::NAN;

            while state.load(Ordering::Relaxed) != -1 || !cons.is_empty() {
                let st = state.load(Ordering::Relaxed);
                if st == 0 || st == -1 {
                    // stft.rs:227-236: paused / stopping -> drain and release
                    while let Ok(idx) = cons.pop() {
                        if idx < slots.slots.len() {
                            slots.release(idx);
                        }
                        let _ = reclaim.send(idx);
                    }
                    thread::sleep(Duration::from_millis(10));
                    continue;
                }

                let mut new_data = false;
                while let Ok(idx) = cons.pop() {
                    if idx >= slots.slots.len() {
                        let _ = reclaim.send(idx); // stft.rs:242-245
                        continue;
                    }
                    // stft.rs:322: the global floor follows the dynamics tracker
                    let db = dynamics_output.read().noise_floor_db;
                    if db != last_db {
                        unsafe { ffi::aa_stream_set_noise_floor_db(gpu.0, db) };
                        last_db = db;
                    }
                    // stft.rs:387: an onset reported by the onset detector snaps the tracker on the next frame
                    if onset_pending.swap(false, Ordering::Relaxed) {
                        unsafe { ffi::aa_stream_signal_onset(gpu.0) };
                    }
                    let status = unsafe {
                        let slot_slice = &*slots.slots[idx].get();
                        ffi::aa_stream_push(gpu.0, slot_slice.as_ptr(), slot_slice.len() as i32)
                    };
                    if slots.release(idx) {
                        let _ = reclaim.send(idx); // stft.rs:256-258
                    }
                    if status != ffi::AA_OK {
                        log::error!("aa_stream_push: {}", ffi::last_error());
                        continue;
                    }
                    new_data = true;

                    // every frame this slot completed (two per 1024-sample slot at hop 512)
                    loop {
                        let mut n: i32 = 0;
                        let st = unsafe { ffi::aa_stream_poll(gpu.0, frames.as_mut_ptr(), frames.len() as i32, &mut n) };
                        if st != ffi::AA_OK || n <= 0 {
                            break;
                        }
                        for f in &frames[..n as usize] {
                            let k = (f.stable.n as usize).min(ffi::AA_MAX_STABLE);
                            if k > 0 {
                                // stft.rs:431-434
                                let pitches: Vec<(f32, f32)> = f.stable.pitch[..k].iter().map(|p| (p.freq, p.score)).collect();
                                let _ = note_tx.push((pitches, transport.get_accumulated_beats()));
                            }
                        }
                    }
                }
                if !new_data {
                    thread::sleep(Duration::from_millis(1)); // stft.rs:268-271
                }
            }
        });
        Ok(())
    }

    /// The reference's infallible signature (stft.rs:155): logs instead of returning the error.
    pub fn detect_pitches(
        &mut self,
        slots: Arc<SlotPool>,
        cons: rtrb::Consumer<usize>,
        reclaim: Sender<usize>,
        sr: u32,
        note_tx: Producer<(Vec<(f32, f32)>, f64)>,
        dynamics_output: Arc<parking_lot::RwLock<DynamicsOutput>>,
        transport: Arc<MusicalTransport>,
        onset_pending: Arc<AtomicBool>,
    ) {
        if let Err(e) = self.try_detect_pitches(slots, cons, reclaim, sr, note_tx, dynamics_output, transport, onset_pending) {
            log::error!("GPU STFT not started: {e}");
        }
    }
}

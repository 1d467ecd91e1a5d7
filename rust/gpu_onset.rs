//! Drop-in for `src/analysis/onset.rs`: `OnsetDetector` with the reference's constructor, `detect_onsets`
//! signature, `stop / pause / resume` and `Drop`.  The frame loop up to the decision inputs (onset.rs:254-357:
//! window, FFT, magnitudes, weighted smoothed flux, per-bin burst floor, energy EMA, FluxTracker) runs on the B200
//! through `aa_stream_*`; everything that needs the live transport stays here, line for line in behaviour:
//! calibration timeout (onset.rs:359-371), `stamp_onset` + metronome tick guard (onset.rs:383-401), the
//! calibration residual logic (onset.rs:404-440), `onset_tx.push` / `onset_pending` (onset.rs:451-453) and the
//! `frames_since_onset` gate (onset.rs:403, 535-539).  UNCOMPILED SOURCE -- see rust/README.md.
use std::{
    sync::{
        Arc,
        atomic::{AtomicBool, AtomicI8, AtomicI64, Ordering},
    },
    thread,
    time::Duration,
};

use crossbeam_channel::Sender;
use rtrb::{Consumer, Producer};

use crate::{
    audio_io::SlotPool,
    audio_io::dynamics::DynamicsOutput,
    audio_io::timing::{MusicalTransport, OnsetEvent},
    dsp::gpu_ffi as ffi,
};

pub struct OnsetDetector {
    state: Arc<AtomicI8>,
    handle: u8,
    reducer_remove_tx: Sender<u8>,
}

impl OnsetDetector {
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

impl Drop for OnsetDetector {
    fn drop(&mut self) {
        let _ = self.reducer_remove_tx.send(self.handle);
        self.stop();
    }
}

struct GpuStream(*mut ffi::aa_stream);
unsafe impl Send for GpuStream {}
impl Drop for GpuStream {
    fn drop(&mut self) {
        unsafe {
            ffi::aa_stream_destroy(self.0);
        }
    }
}

impl OnsetDetector {
    pub fn new(handle: u8, reducer_remove_tx: Sender<u8>) -> Self {
        OnsetDetector { state: Arc::new(AtomicI8::new(0)), handle, reducer_remove_tx }
    }

    /// Same arguments as the reference (onset.rs:104-114); `Err(msg)` maps to
    /// `AudioEngineError::SpawnFailed { component: "onset", msg }`.
    pub fn try_detect_onsets(
        &mut self,
        transport: Arc<MusicalTransport>,
        slots: Arc<SlotPool>,
        mut cons: Consumer<usize>,
        reclaim: Sender<usize>,
        mut onset_tx: Producer<OnsetEvent>,
        onset_pending: Arc<AtomicBool>,
        dynamics_output: Arc<parking_lot::RwLock<DynamicsOutput>>,
        calibration_target: Arc<AtomicI64>,
    ) -> Result<(), String> {
        let window_size: i64 = 256; // onset.rs:122
        let hop_size: i64 = 64; // onset.rs:123
        let mut cfg = unsafe { std::mem::zeroed::<ffi::aa_config>() };
        unsafe { ffi::aa_config_default_onset(&mut cfg, transport.get_sample_rate()) };
        let mut raw = std::ptr::null_mut();
        ffi::check(unsafe { ffi::aa_stream_create(&cfg, &mut raw) })?;
        let gpu = GpuStream(raw);

        self.state.store(1, Ordering::Relaxed);
        let state = self.state.clone();

        // onset.rs:126-136: calibration bookkeeping taken on the calling thread
        let calibration_initially_done = transport.is_calibrated();
        let calibration_start_frame = transport.get_output_frames();
        let calibration_timeout_samples = transport.get_sample_rate() as i64 * 2;

        thread::spawn(move || {
            let gpu = gpu;
            let mut calibration_done = calibration_initially_done;
            const TICK_GUARD_S: f64 = 0.015; // onset.rs:186
            let mut frames_since_onset: usize = 4; // onset.rs:200
            let mut frames: Vec<ffi::aa_stream_frame> = Vec::with_capacity(64);
            unsafe { frames.set_len(64) };
            let mut last_db = f32::NAN;

            while state.load(Ordering::Relaxed) != -1 || !cons.is_empty() {
                let st = state.load(Ordering::Relaxed);
                if st == 0 || st == -1 {
                    // onset.rs:203-213
                    match cons.pop() {
                        Ok(idx) => {
                            if slots.release(idx) {
                                let _ = reclaim.send(idx);
                            }
                        }
                        Err(_) => thread::sleep(Duration::from_millis(5)),
                    }
                    continue;
                }

                let mut new_data = false;
                while let Ok(idx) = cons.pop() {
                    debug_assert!(idx < slots.slots.len());
                    let db = dynamics_output.read().noise_floor_db; // onset.rs:300
                    if db != last_db {
                        unsafe { ffi::aa_stream_set_noise_floor_db(gpu.0, db) };
                        last_db = db;
                    }
                    let status = unsafe {
                        let slot_slice = &*slots.slots[idx].get();
                        ffi::aa_stream_push(gpu.0, slot_slice.as_ptr(), slot_slice.len() as i32)
                    };
                    if slots.release(idx) {
                        let _ = reclaim.send(idx);
                    }
                    if status != ffi::AA_OK {
                        log::error!("aa_stream_push: {}", ffi::last_error());
                        continue;
                    }
                    new_data = true;

                    loop {
                        let mut n: i32 = 0;
                        let st = unsafe { ffi::aa_stream_poll(gpu.0, frames.as_mut_ptr(), frames.len() as i32, &mut n) };
                        if st != ffi::AA_OK || n <= 0 {
                            break;
                        }
                        for (i, fr) in frames[..n as usize].iter().enumerate() {
                            let f = &fr.features;
                            let onset_detected = f.flags & ffi::AA_FLAG_ONSET_DETECTED != 0; // onset.rs:357
                            let energy_rising = f.flags & ffi::AA_FLAG_ENERGY_RISING != 0; // onset.rs:373
                            let mut onset_fired = false;

                            // onset.rs:359-371: calibration timeout
                            if !calibration_done {
                                let elapsed = transport.get_output_frames() - calibration_start_frame;
                                if elapsed > calibration_timeout_samples {
                                    log::warn!("onset calibration timed out after {} samples — using offset 0", elapsed);
                                    transport.set_calibration_offset(0);
                                    calibration_done = true;
                                }
                            }

                            if onset_detected {
                                // samples the reference would still hold behind this frame when it processed it
                                // (`available_samples`, onset.rs:386): the window plus the frames of this poll that follow
                                let available = window_size + (n as i64 - 1 - i as i64) * hop_size;
                                let window_centre_offset = -(available - window_size / 2);
                                let velocity = (f.flux.max(f.max_excess * 5.0) / 50.0).clamp(0.0, 1.0); // onset.rs:388-390
                                let event = transport.stamp_onset(window_centre_offset, velocity);

                                let bpm = transport.get_bpm() as f64;
                                let tick_guard_beats = TICK_GUARD_S * bpm / 60.0;
                                let tick_dist = transport.nearest_tick_distance_beats(event.beat_position);
                                let suppressed_by_tick = tick_dist < tick_guard_beats;

                                if !suppressed_by_tick && energy_rising && frames_since_onset >= 3 {
                                    if !calibration_done {
                                        // onset.rs:404-440
                                        let target = calibration_target.load(Ordering::Relaxed);
                                        if target != 0 {
                                            let sr_f64 = transport.get_sample_rate() as f64;
                                            let calibration_samples = event.output_samples - target;
                                            let max_cal = (sr_f64 * 0.5) as i64;
                                            if calibration_samples < 0 || calibration_samples > max_cal {
                                                log::warn!("onset calibration: rejected implausible residual — retrying");
                                            } else {
                                                transport.set_calibration_offset(calibration_samples);
                                                calibration_done = true;
                                                onset_pending.store(false, Ordering::Relaxed);
                                                onset_fired = true;
                                            }
                                        }
                                    } else {
                                        let _ = onset_tx.push(event); // onset.rs:451
                                        onset_pending.store(true, Ordering::Relaxed); // onset.rs:452
                                        onset_fired = true;
                                    }
                                }
                            }

                            // onset.rs:535-539
                            if onset_fired || onset_detected && frames_since_onset < 3 {
                                frames_since_onset = 0;
                            } else {
                                frames_since_onset = frames_since_onset.saturating_add(1);
                            }
                        }
                    }
                }
                if !new_data {
                    thread::sleep(Duration::from_millis(1));
                }
            }
        });
        Ok(())
    }

    /// The reference's infallible signature (onset.rs:104).
    pub fn detect_onsets(
        &mut self,
        transport: Arc<MusicalTransport>,
        slots: Arc<SlotPool>,
        cons: Consumer<usize>,
        reclaim: Sender<usize>,
        onset_tx: Producer<OnsetEvent>,
        onset_pending: Arc<AtomicBool>,
        dynamics_output: Arc<parking_lot::RwLock<DynamicsOutput>>,
        calibration_target: Arc<AtomicI64>,
    ) {
        if let Err(e) = self.try_detect_onsets(
            transport, slots, cons, reclaim, onset_tx, onset_pending, dynamics_output, calibration_target,
        ) {
            log::error!("GPU onset detector not started: {e}");
        }
    }
}

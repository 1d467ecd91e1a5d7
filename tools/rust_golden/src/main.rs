//! Golden-vector kit (see ../README.md).  NEVER COMPILED in the build image: no Rust toolchain there.
//!
//! The reference's source files are pulled in with include!() from $AA_REF_SRC under the module paths they
//! expect (`crate::dsp::fft`, `crate::audio_io::{SlotPool, dynamics, timing, stft}`, `crate::analysis::onset`),
//! so the code that runs is the reference's, byte for byte.
#![allow(dead_code, unused_imports, unused_variables)]

use std::{
    fs,
    io::{Read, Write},
    path::Path,
    sync::{
        Arc,
        atomic::{AtomicBool, AtomicI64, Ordering},
    },
    thread,
    time::Duration,
};

pub mod dsp {
    pub mod fft {
        include!(concat!(env!("AA_REF_SRC"), "/dsp/fft.rs"));
    }
}

pub mod audio_io {
    use std::{
        cell::UnsafeCell,
        sync::{
            Arc,
            atomic::{AtomicUsize, Ordering},
        },
    };

    /// Stand-in for the reference's SlotPool (audio_io/mod.rs:32-79: private constructor, lives beside the cpal
    /// code): the same public field and the same release() contract -- true when the last consumer let go.
    pub struct SlotPool {
        pub slots: Vec<UnsafeCell<Box<[f32]>>>,
        users: Vec<AtomicUsize>,
    }
    unsafe impl Send for SlotPool {}
    unsafe impl Sync for SlotPool {}
    impl SlotPool {
        pub fn with_slots(count: usize, len: usize) -> Arc<Self> {
            Arc::new(Self {
                slots: (0..count).map(|_| UnsafeCell::new(vec![0.0f32; len].into_boxed_slice())).collect(),
                users: (0..count).map(|_| AtomicUsize::new(0)).collect(),
            })
        }
        pub fn fill(&self, idx: usize, samples: &[f32], consumers: usize) {
            unsafe { (&mut *self.slots[idx].get()).copy_from_slice(samples) };
            self.users[idx].store(consumers, Ordering::SeqCst);
        }
        pub fn release(&self, idx: usize) -> bool {
            let mut cur = self.users[idx].load(Ordering::SeqCst);
            loop {
                if cur == 0 {
                    return false;
                }
                match self.users[idx].compare_exchange(cur, cur - 1, Ordering::SeqCst, Ordering::SeqCst) {
                    Ok(_) => return cur == 1,
                    Err(now) => cur = now,
                }
            }
        }
    }

    pub mod dynamics {
        include!(concat!(env!("AA_REF_SRC"), "/audio_io/dynamics.rs"));
    }
    pub mod timing {
        include!(concat!(env!("AA_REF_SRC"), "/audio_io/timing.rs"));
    }
    pub mod stft {
        include!(concat!(env!("AA_REF_SRC"), "/audio_io/stft.rs"));
    }
}

pub mod analysis {
    pub mod onset {
        include!(concat!(env!("AA_REF_SRC"), "/analysis/onset.rs"));
    }
}

use audio_io::{SlotPool, dynamics::DynamicsOutput, timing::MusicalTransport};

const SLOT: usize = 1024; // mod.rs:126-128

// ---- minimal .npy (v1.0, little endian, C order) ---------------------------------------------------
fn npy_write(path: &Path, descr: &str, shape: &[usize], bytes: &[u8]) {
    let dims = match shape.len() {
        1 => format!("({},)", shape[0]),
        _ => format!("({})", shape.iter().map(|d| d.to_string()).collect::<Vec<_>>().join(", ")),
    };
    let mut header = format!("{{'descr': '{descr}', 'fortran_order': False, 'shape': {dims}, }}");
    while (10 + header.len() + 1) % 64 != 0 {
        header.push(' ');
    }
    header.push('\n');
    let mut f = fs::File::create(path).expect("create npy");
    f.write_all(b"\x93NUMPY\x01\x00").unwrap();
    f.write_all(&(header.len() as u16).to_le_bytes()).unwrap();
    f.write_all(header.as_bytes()).unwrap();
    f.write_all(bytes).unwrap();
}

fn npy_read_f32(path: &Path) -> Vec<f32> {
    let mut raw = Vec::new();
    fs::File::open(path).expect("open npy").read_to_end(&mut raw).unwrap();
    assert!(&raw[..6] == b"\x93NUMPY" && raw[6] == 1, "expected an .npy v1 file");
    let hlen = u16::from_le_bytes([raw[8], raw[9]]) as usize;
    let header = std::str::from_utf8(&raw[10..10 + hlen]).unwrap();
    assert!(header.contains("<f4"), "expected float32 data");
    raw[10 + hlen..].chunks_exact(4).map(|b| f32::from_le_bytes([b[0], b[1], b[2], b[3]])).collect()
}

fn f32_bytes(v: &[f32]) -> Vec<u8> {
    v.iter().flat_map(|x| x.to_le_bytes()).collect()
}
fn f64_bytes(v: &[f64]) -> Vec<u8> {
    v.iter().flat_map(|x| x.to_le_bytes()).collect()
}

// ---- the three runs -------------------------------------------------------------------------------
/// FftProcessor::process_forward on the first `frames` raw (unwindowed) frames of length n, hop n/4.
fn run_fft(x: &[f32], n: usize, frames: usize) -> Vec<f32> {
    let mut p = dsp::fft::FftProcessor::new(n);
    let mut out = Vec::with_capacity(frames * (n + 2));
    for t in 0..frames {
        let mut buf = x[t * n / 4..t * n / 4 + n].to_vec();
        for c in p.process_forward(&mut buf) {
            out.push(c.re);
            out.push(c.im);
        }
    }
    out
}

/// Feeds `x` slot by slot to a worker through a real rtrb queue, the way the reducer thread does
/// (mod.rs:418-496); `after_slot` runs once the worker has had time to consume the slot.
fn feed_slots(x: &[f32], pool: &Arc<SlotPool>, prod: &mut rtrb::Producer<usize>, transport: &Arc<MusicalTransport>,
              mut after_slot: impl FnMut(usize)) {
    for (k, chunk) in x.chunks_exact(SLOT).enumerate() {
        let idx = k % pool.slots.len();
        transport.tick_output(SLOT as i64, k as f64 * SLOT as f64 / transport.get_sample_rate() as f64);
        pool.fill(idx, chunk, 1);
        while prod.push(idx).is_err() {
            thread::sleep(Duration::from_micros(200));
        }
        thread::sleep(Duration::from_millis(4)); // one slot is ~50 us of analysis: the worker is idle again
        after_slot(k);
    }
}

/// The whole STFT::detect_pitches worker.  Row layout: [slot index, beat, n, f0, s0, f1, s1, ...] (f64), one row
/// per pushed (non-empty) frame, up to 16 pitches.
fn run_stft(x: &[f32], sr: f32) -> Vec<f64> {
    let pool = SlotPool::with_slots(64, SLOT);
    let (mut prod, cons) = rtrb::RingBuffer::<usize>::new(1024);
    let (reclaim_tx, _reclaim_rx) = crossbeam_channel::unbounded::<usize>();
    let (remove_tx, _remove_rx) = crossbeam_channel::unbounded::<u8>();
    let (note_tx, mut note_rx) = rtrb::RingBuffer::<(Vec<(f32, f32)>, f64)>::new(64);
    let transport = MusicalTransport::new(120.0, sr);
    let dynamics = Arc::new(parking_lot::RwLock::new(DynamicsOutput::default())); // noise_floor_db = -96
    let onset_pending = Arc::new(AtomicBool::new(false));
    let mut stft = audio_io::stft::STFT::new(1, remove_tx);
    stft.detect_pitches(pool.clone(), cons, reclaim_tx, sr as u32, note_tx, dynamics, transport.clone(), onset_pending);
    let mut rows = Vec::new();
    feed_slots(x, &pool, &mut prod, &transport, |k| {
        while let Ok((pitches, beat)) = note_rx.pop() {
            rows.push(k as f64);
            rows.push(beat);
            rows.push(pitches.len() as f64);
            for i in 0..16 {
                let (f, s) = pitches.get(i).copied().unwrap_or((0.0, 0.0));
                rows.push(f as f64);
                rows.push(s as f64);
            }
        }
    });
    stft.stop();
    rows
}

/// The whole OnsetDetector::detect_onsets worker with a calibrated transport and no metronome ticks.
/// Row layout: [slot index, window-centre sample position, velocity, beat_position, raw_sample_offset] (f64).
fn run_onset(x: &[f32], sr: f32) -> Vec<f64> {
    let pool = SlotPool::with_slots(64, SLOT);
    let (mut prod, cons) = rtrb::RingBuffer::<usize>::new(1024);
    let (reclaim_tx, _reclaim_rx) = crossbeam_channel::unbounded::<usize>();
    let (remove_tx, _remove_rx) = crossbeam_channel::unbounded::<u8>();
    let (onset_tx, mut onset_rx) = rtrb::RingBuffer::<audio_io::timing::OnsetEvent>::new(1024);
    let transport = MusicalTransport::new(120.0, sr);
    transport.set_calibration_offset(0); // calibration done: the detector pushes events (onset.rs:441-453)
    let dynamics = Arc::new(parking_lot::RwLock::new(DynamicsOutput::default()));
    let onset_pending = Arc::new(AtomicBool::new(false));
    let target = Arc::new(AtomicI64::new(0));
    let mut det = analysis::onset::OnsetDetector::new(2, remove_tx);
    det.detect_onsets(transport.clone(), pool.clone(), cons, reclaim_tx, onset_tx, onset_pending, dynamics, target);
    let mut rows = Vec::new();
    feed_slots(x, &pool, &mut prod, &transport, |k| {
        while let Ok(ev) = onset_rx.pop() {
            rows.push(k as f64);
            rows.push((ev.output_samples) as f64); // = samples fed so far + raw offset = the window centre
            rows.push(ev.velocity as f64);
            rows.push(ev.beat_position);
            rows.push(ev.raw_sample_offset as f64);
        }
    });
    det.stop();
    rows
}

fn main() {
    let dir = std::env::args().nth(1).expect("usage: rust_golden <tests/golden/ref>");
    let dir = Path::new(&dir);
    let mut names: Vec<String> = fs::read_dir(dir)
        .expect("read dir")
        .filter_map(|e| e.ok())
        .filter_map(|e| e.file_name().into_string().ok())
        .filter(|n| n.starts_with("in_") && n.ends_with(".npy"))
        .map(|n| n[3..n.len() - 4].to_string())
        .collect();
    names.sort();
    for name in names {
        let x = npy_read_f32(&dir.join(format!("in_{name}.npy")));
        let sr: f32 = fs::read_to_string(dir.join(format!("in_{name}.sr"))).unwrap().trim().parse().unwrap();
        for n in [256usize, 2048, 4096] {
            let frames = 16;
            let spec = run_fft(&x, n, frames);
            npy_write(&dir.join(format!("{name}.spectra{n}.npy")), "<f4", &[frames, n / 2 + 1, 2], &f32_bytes(&spec));
        }
        let st = run_stft(&x, sr);
        npy_write(&dir.join(format!("{name}.stable.npy")), "<f8", &[st.len() / 35, 35], &f64_bytes(&st));
        let on = run_onset(&x, sr);
        npy_write(&dir.join(format!("{name}.onsets.npy")), "<f8", &[on.len() / 5, 5], &f64_bytes(&on));
        println!("{name}: {} pitch frames, {} onset events", st.len() / 35, on.len() / 5);
    }
}

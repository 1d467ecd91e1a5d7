// audio_engine_gpu.hpp -- C++ host-side mirror of the reference's Rust types for the
// frame-analysis path, layered on the C ABI of libaa_gpu.so (include/aa_gpu.h).
//
// The reference's host side is Rust; this image has no Rust toolchain, so the mirror is
// written in C++ (the reference is compiled code) with the same names, argument meaning and
// error behaviour, so that a maintainer can read it next to the Rust source:
//
//   FftProcessor          src/dsp/fft.rs:6-41        new / process_forward / process_inverse
//   STFT                  src/audio_io/stft.rs:120-441  new / detect_pitches / stop / pause / resume
//   OnsetDetector         src/analysis/onset.rs:21-546  new / detect_onsets / stop / pause / resume
//   PitchFrame            the (Vec<(f32,f32)>, f64) pushed on note_tx (stft.rs:431-434)
//   OnsetEvent            src/audio_io/timing.rs:77-87
//   Reducer               the reducer thread's per-slot work, src/audio_io/mod.rs:431-490
//                         (HPF / LPF / gate + DynamicsTracker::process_slot, dynamics.rs:194-360)
//   DynamicsOutput        src/audio_io/dynamics.rs:78-104
//   Tuner / TunerOutput   src/analysis/tuner.rs:27-208 (per-frame branch; labels formatted here)
//
// What stays on the host, exactly as in the reference: the worker thread and its
// stop/pause/resume state machine (AtomicI8 -1/0/1, stft.rs:127-135), the slot
// queue draining while paused (stft.rs:227-236), the "emit only non-empty frames" rule
// (stft.rs:431), and the transport-dependent onset gating (onset.rs:383-456: stamp_onset,
// tick guard, energy_rising, frames_since_onset).  What moves to the GPU: everything between
// "slot arrives" and "per-frame result" (window, FFT, magnitudes, floors, extract_pitches,
// PitchTracker, flux / burst / FluxTracker).
//
// Header-only; link with -laa_gpu.  Errors: the reference panics (unwrap) or returns
// anyhow::Error; here every failure throws aa::Error carrying aa_last_error().
#pragma once

#include <atomic>
#include <chrono>
#include <cstdint>
#include <deque>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/aa_gpu.h"

namespace audio_engine_gpu {

struct Error : std::runtime_error {
    aa_status code;
    Error(aa_status c, const std::string &what) : std::runtime_error(what), code(c) {}
};
inline void check(aa_status st)
{
    if (st != AA_OK) throw Error(st, aa_last_error());
}

struct Complex32 {
    float re, im;
};

// ---------------------------------------------------------------------------
// FftProcessor (src/dsp/fft.rs:6-41)
// ---------------------------------------------------------------------------
class FftProcessor {
public:
    // FftProcessor::new(len) (fft.rs:14)
    explicit FftProcessor(std::size_t len) : len_(len), spectrum_(len / 2 + 1), output_(len)
    {
        check(aa_fft_create(static_cast<int32_t>(len), &h_));
    }
    ~FftProcessor() { aa_fft_destroy(h_); }
    FftProcessor(const FftProcessor &) = delete;
    FftProcessor &operator=(const FftProcessor &) = delete;

    // process_forward(&mut self, windowed: &mut [f32]) -> &[Complex<f32>] (fft.rs:33).
    // The returned pointer addresses len/2+1 bins owned by the processor and stays valid until
    // the next call, as the borrowed slice does in the reference.  A wrong length throws where
    // the reference panics (unwrap, fft.rs:69).  The reference clobbers `windowed`; here it is
    // left intact, which no caller can depend on.
    const Complex32 *process_forward(const float *windowed, std::size_t n)
    {
        if (n != len_) throw Error(AA_ERR_INVALID, "process_forward: wrong input length");
        check(aa_fft_forward(h_, windowed, 1, reinterpret_cast<float *>(spectrum_.data())));
        return spectrum_.data();
    }
    // process_inverse(&mut self, &mut [Complex<f32>]) -> &[f32] (fft.rs:39), unnormalised.
    const float *process_inverse(const Complex32 *spectrum, std::size_t bins)
    {
        if (bins != len_ / 2 + 1) throw Error(AA_ERR_INVALID, "process_inverse: wrong spectrum length");
        check(aa_fft_inverse(h_, reinterpret_cast<const float *>(spectrum), 1, output_.data()));
        return output_.data();
    }
    std::size_t len() const { return len_; }
    std::size_t bins() const { return len_ / 2 + 1; }

private:
    std::size_t len_;
    aa_fft *h_ = nullptr;
    std::vector<Complex32> spectrum_;
    std::vector<float> output_;
};

// ---------------------------------------------------------------------------
// Slot hand-off.  In the reference the producer side is SlotPool + rtrb SPSC of slot
// indices (audio_io/mod.rs:32-79, 945-949); consumers copy each 1024-sample slot into a
// private ring and release it.  Here the queue carries the samples themselves and the
// "private ring" is the pinned-host + device ring inside aa_stream.
// ---------------------------------------------------------------------------
class SlotQueue {
public:
    void push(std::vector<float> slot)
    {
        std::lock_guard<std::mutex> g(m_);
        q_.push_back(std::move(slot));
    }
    bool pop(std::vector<float> &out)
    {
        std::lock_guard<std::mutex> g(m_);
        if (q_.empty()) return false;
        out = std::move(q_.front());
        q_.pop_front();
        return true;
    }
    bool is_empty()
    {
        std::lock_guard<std::mutex> g(m_);
        return q_.empty();
    }

private:
    std::mutex m_;
    std::deque<std::vector<float>> q_;
};

// DynamicsOutput.noise_floor_db (dynamics.rs:80-103) and the onset_pending flag shared by the
// two analyzers (mod.rs, stft.rs:387, onset.rs:452).
struct SharedAnalysisState {
    std::atomic<float> noise_floor_db{-96.0f};
    std::atomic<bool> onset_pending{false};
};

using PitchList = std::vector<std::pair<float, float>>;   // Vec<(freq, score)>
struct PitchFrame {
    PitchList pitches;   // stable pitches of PitchTracker (stft.rs:390)
    double beat;         // transport.get_accumulated_beats() at emission (stft.rs:432)
};

// ---------------------------------------------------------------------------
// STFT (src/audio_io/stft.rs:120-441)
// ---------------------------------------------------------------------------
class STFT {
public:
    // STFT::new(handle, reducer_remove_tx) (stft.rs:147): `on_drop` plays the role of
    // reducer_remove_tx.send(handle) in Drop (stft.rs:138-144).
    explicit STFT(uint8_t handle, std::function<void(uint8_t)> on_drop = {})
        : handle_(handle), on_drop_(std::move(on_drop))
    {
    }
    ~STFT()
    {
        if (on_drop_) on_drop_(handle_);
        stop();
        if (worker_.joinable()) worker_.join();
    }
    void stop() { state_.store(-1, std::memory_order_relaxed); }     // stft.rs:127
    void pause() { state_.store(0, std::memory_order_relaxed); }     // stft.rs:130
    void resume() { state_.store(1, std::memory_order_relaxed); }    // stft.rs:133

    // detect_pitches(slots, cons, reclaim, sr, note_tx, dynamics_output, transport, onset_pending)
    // (stft.rs:155-165).  `cons` is the slot queue, `note_tx` receives every non-empty
    // stable-pitch frame, `beats` stands for transport.get_accumulated_beats().
    void detect_pitches(SlotQueue &cons, uint32_t sr, std::function<void(PitchFrame)> note_tx,
                        SharedAnalysisState &shared, std::function<double()> beats)
    {
        aa_config cfg;
        aa_config_default_pitch(&cfg, static_cast<float>(sr));      // 2048 / 512, 24..10000 Hz (stft.rs:169-174)
        aa_stream *stream = nullptr;
        check(aa_stream_create(&cfg, &stream));                     // a GPU failure maps to SpawnFailed upstream
        state_.store(1, std::memory_order_relaxed);                 // stft.rs:166
        worker_ = std::thread([this, &cons, note_tx = std::move(note_tx), &shared, beats = std::move(beats), stream] {
            std::vector<float> slot;
            std::vector<aa_stream_frame> frames(64);
            float last_db = -96.0f;
            while (state_.load(std::memory_order_relaxed) != -1 || !cons.is_empty()) {      // stft.rs:226
                const int st = state_.load(std::memory_order_relaxed);
                if (st == 0 || st == -1) {                                                  // stft.rs:227-236
                    while (cons.pop(slot)) {}
                    std::this_thread::sleep_for(std::chrono::milliseconds(10));
                    continue;
                }
                bool new_data = false;
                while (cons.pop(slot)) {                                                    // stft.rs:240-260
                    const float db = shared.noise_floor_db.load(std::memory_order_relaxed); // stft.rs:322
                    if (db != last_db) { aa_stream_set_noise_floor_db(stream, db); last_db = db; }
                    if (shared.onset_pending.exchange(false, std::memory_order_relaxed))     // stft.rs:387
                        aa_stream_signal_onset(stream);
                    if (aa_stream_push(stream, slot.data(), static_cast<int32_t>(slot.size())) != AA_OK) break;
                    new_data = true;
                    int32_t n = 0;
                    while (aa_stream_poll(stream, frames.data(), static_cast<int32_t>(frames.size()), &n) == AA_OK && n > 0) {
                        for (int32_t i = 0; i < n; ++i) {
                            const aa_stable_pitches &sp = frames[i].stable;
                            if (sp.n == 0) continue;                                         // stft.rs:431
                            PitchFrame pf;
                            pf.pitches.reserve(sp.n);
                            for (uint32_t j = 0; j < sp.n; ++j) pf.pitches.emplace_back(sp.pitch[j].freq, sp.pitch[j].score);
                            pf.beat = beats ? beats() : 0.0;
                            note_tx(std::move(pf));
                        }
                    }
                }
                if (!new_data) std::this_thread::sleep_for(std::chrono::milliseconds(1));    // stft.rs:268-271
            }
            aa_stream_destroy(stream);
        });
    }

private:
    std::atomic<int8_t> state_{0};
    uint8_t handle_;
    std::function<void(uint8_t)> on_drop_;
    std::thread worker_;
};

// ---------------------------------------------------------------------------
// OnsetDetector (src/analysis/onset.rs:21-546)
// ---------------------------------------------------------------------------
struct OnsetEvent {            // src/audio_io/timing.rs:77-87
    double beat_position;
    int64_t raw_sample_offset;
    int64_t output_samples;
    float velocity;
};

// window_centre_offset of frame i of a poll that returned n frames (onset.rs:386-387:
// -(available_samples - window_size / 2)): when the reference processed that frame, the samples still buffered behind
// its start were its own window plus one hop for every frame that follows it in the burst.  With 1024-sample slots
// this gives exactly the offsets -128 ... -1088 the real crate logs (tests/golden/ref_log_onsets.json).
inline int64_t window_centre_offset(int32_t n_in_poll, int32_t i, int window_size, int hop)
{
    const int64_t available = window_size + static_cast<int64_t>(n_in_poll - 1 - i) * hop;
    return -(available - window_size / 2);
}

// MusicalTransport::stamp_onset (timing.rs:311-337) over plain values, for hosts that keep the transport's atomics
// themselves: the f64 operations in the reference's order.
inline OnsetEvent stamp_onset_at(double current_beats, int64_t output_frames, double bpm, double sample_rate,
                                 int64_t input_latency, int64_t output_latency, int64_t calibration,
                                 int64_t sample_offset, float velocity)
{
    const double beats_per_sample = bpm / (60.0 * sample_rate);
    const double latency_beats = static_cast<double>(input_latency + output_latency) * beats_per_sample;
    const double offset_beats = static_cast<double>(sample_offset) * beats_per_sample;
    const double calibration_beats = static_cast<double>(calibration) * beats_per_sample;
    OnsetEvent ev;
    ev.beat_position = current_beats - latency_beats + offset_beats - calibration_beats;
    ev.raw_sample_offset = sample_offset;
    ev.output_samples = output_frames - input_latency - output_latency + sample_offset - calibration;
    ev.velocity = velocity;
    return ev;
}

// The slice of MusicalTransport the onset gating needs (timing.rs): stamp_onset,
// nearest_tick_distance_beats, get_bpm.
struct TransportHooks {
    std::function<OnsetEvent(int64_t window_centre_offset, float velocity)> stamp_onset;
    std::function<double(double beat)> nearest_tick_distance_beats;
    std::function<float()> get_bpm;
};

class OnsetDetector {
public:
    explicit OnsetDetector(uint8_t handle, std::function<void(uint8_t)> on_drop = {})
        : handle_(handle), on_drop_(std::move(on_drop))
    {
    }
    ~OnsetDetector()
    {
        state_.store(-1, std::memory_order_relaxed);
        if (on_drop_) on_drop_(handle_);
        if (worker_.joinable()) worker_.join();
    }
    void stop() { state_.store(-1, std::memory_order_relaxed); }
    void pause() { state_.store(0, std::memory_order_relaxed); }
    void resume() { state_.store(1, std::memory_order_relaxed); }

    // detect_onsets(transport, slots, cons, reclaim, onset_tx, onset_pending, dynamics_output,
    // calibration_target) (onset.rs:104-114).  The round-trip latency self-calibration
    // (onset.rs:359-371, 404-440) needs the live output device and is left to the caller.
    void detect_onsets(TransportHooks transport, SlotQueue &cons, uint32_t sr,
                       std::function<void(OnsetEvent)> onset_tx, SharedAnalysisState &shared)
    {
        aa_config cfg;
        aa_config_default_onset(&cfg, static_cast<float>(sr));      // 256 / 64 (onset.rs:122-123)
        aa_stream *stream = nullptr;
        check(aa_stream_create(&cfg, &stream));
        state_.store(1, std::memory_order_relaxed);
        worker_ = std::thread([this, transport = std::move(transport), &cons, onset_tx = std::move(onset_tx), &shared,
                               stream] {
            constexpr double TICK_GUARD_S = 0.015;                  // onset.rs:186
            const int window_size = 256, hop = 64;
            std::vector<float> slot;
            std::vector<aa_stream_frame> frames(64);
            std::size_t frames_since_onset = 4;                     // onset.rs:200
            float last_db = -96.0f;
            while (state_.load(std::memory_order_relaxed) != -1 || !cons.is_empty()) {
                const int st = state_.load(std::memory_order_relaxed);
                if (st == 0 || st == -1) {                          // onset.rs:203-213
                    if (!cons.pop(slot)) std::this_thread::sleep_for(std::chrono::milliseconds(5));
                    continue;
                }
                bool new_data = false;
                while (cons.pop(slot)) {
                    const float db = shared.noise_floor_db.load(std::memory_order_relaxed);  // onset.rs:300
                    if (db != last_db) { aa_stream_set_noise_floor_db(stream, db); last_db = db; }
                    if (aa_stream_push(stream, slot.data(), static_cast<int32_t>(slot.size())) != AA_OK) break;
                    new_data = true;
                    int32_t n = 0;
                    while (aa_stream_poll(stream, frames.data(), static_cast<int32_t>(frames.size()), &n) == AA_OK && n > 0) {
                        for (int32_t i = 0; i < n; ++i) {
                            const aa_frame_features &f = frames[i].features;
                            const bool onset_detected = (f.flags & AA_FLAG_ONSET_DETECTED) != 0;   // onset.rs:357
                            const bool energy_rising = (f.flags & AA_FLAG_ENERGY_RISING) != 0;     // onset.rs:373
                            bool onset_fired = false;
                            if (onset_detected) {                                                  // onset.rs:383-456
                                // samples still buffered behind this frame when the reference would have
                                // processed it: the frames of this poll that follow it, plus the window
                                const int64_t window_centre_offset =
                                    audio_engine_gpu::window_centre_offset(n, i, window_size, hop);
                                float velocity = f.flux > f.max_excess * 5.0f ? f.flux : f.max_excess * 5.0f;
                                velocity = velocity / 50.0f;
                                velocity = velocity < 0.0f ? 0.0f : (velocity > 1.0f ? 1.0f : velocity);
                                const OnsetEvent ev = transport.stamp_onset(window_centre_offset, velocity);
                                const double bpm = transport.get_bpm();
                                const double tick_guard_beats = TICK_GUARD_S * bpm / 60.0;
                                const bool suppressed_by_tick =
                                    transport.nearest_tick_distance_beats(ev.beat_position) < tick_guard_beats;
                                if (!suppressed_by_tick && energy_rising && frames_since_onset >= 3) {
                                    onset_tx(ev);                                                  // onset.rs:451
                                    shared.onset_pending.store(true, std::memory_order_relaxed);  // onset.rs:452
                                    onset_fired = true;
                                }
                            }
                            if (onset_fired || (onset_detected && frames_since_onset < 3)) frames_since_onset = 0;  // onset.rs:535
                            else if (frames_since_onset != SIZE_MAX) ++frames_since_onset;
                        }
                    }
                }
                if (!new_data) std::this_thread::sleep_for(std::chrono::milliseconds(1));
            }
            aa_stream_destroy(stream);
        });
    }

private:
    std::atomic<int8_t> state_{0};
    uint8_t handle_;
    std::function<void(uint8_t)> on_drop_;
    std::thread worker_;
};

// ---------------------------------------------------------------------------
// Offline batch (NEW; no reference equivalent): clips -> per-frame records.
// ---------------------------------------------------------------------------
class BatchAnalyzer {
public:
    explicit BatchAnalyzer(const aa_config &cfg) : cfg_(cfg) { check(aa_analyzer_create(&cfg_, &h_)); }
    ~BatchAnalyzer() { aa_analyzer_destroy(h_); }
    BatchAnalyzer(const BatchAnalyzer &) = delete;
    BatchAnalyzer &operator=(const BatchAnalyzer &) = delete;
    int64_t num_frames(int64_t clip_len) const { return aa_num_frames(&cfg_, clip_len); }
    void analyze_host(const float *clips, int64_t n_clips, int64_t clip_len, int64_t clip_stride,
                      const aa_outputs &out, const uint8_t *onset_in = nullptr)
    {
        check(aa_analyze_host(h_, clips, n_clips, clip_len, clip_stride, onset_in, &out));
    }
    void analyze_device(const float *clips_dev, int64_t n_clips, int64_t clip_len, int64_t clip_stride,
                        const aa_outputs &out_dev, void *stream, const uint8_t *onset_in_dev = nullptr)
    {
        check(aa_analyze_device(h_, clips_dev, n_clips, clip_len, clip_stride, onset_in_dev, &out_dev, stream));
    }

private:
    aa_config cfg_;
    aa_analyzer *h_ = nullptr;
};

// ---------------------------------------------------------------------------
// Reducer: what the reducer thread does to every slot before it fans the slot out to the analyzers
// (src/audio_io/mod.rs:431-490).  The filters, the gate and DynamicsTracker::process_slot run behind
// aa_condition_host with the state carried in the handle; the consumer bookkeeping stays on the host.
// ---------------------------------------------------------------------------
enum class DynamicLevel { Silence, Ppp, Pp, P, Mp, Mf, F, Ff, Fff };     // dynamics.rs:49-60
inline const char *to_string(DynamicLevel l)                             // dynamics.rs:62-77
{
    static const char *n[] = {"silence", "ppp", "pp", "p", "mp", "mf", "f", "ff", "fff"};
    return n[static_cast<int>(l)];
}
struct DynamicsOutput {                                                  // dynamics.rs:78-104
    DynamicLevel level = DynamicLevel::Silence;
    float rms_db = -96.0f, gain_db = 0.0f, session_median_db = -96.0f, noise_floor_db = -96.0f;
};

class Reducer {
public:
    // sample_rate / slot_len as in AudioPipeline (mod.rs:330-355); AGC target -18 dBFS, 240 s smoothing
    Reducer(float sample_rate, int32_t slot_len = 1024)
    {
        cfg_.sample_rate = sample_rate;
        cfg_.slot_len = slot_len;
        cfg_.flags = AA_COND_AGC | AA_COND_CARRY;
        check(aa_conditioner_create(&cfg_, &h_));
    }
    ~Reducer() { aa_conditioner_destroy(h_); }
    Reducer(const Reducer &) = delete;
    Reducer &operator=(const Reducer &) = delete;
    // one slot, in place (mod.rs:433-490); the returned value is what the reference publishes into
    // Arc<RwLock<DynamicsOutput>> (dynamics.rs:355-362)
    DynamicsOutput process_slot(float *slot, size_t len)
    {
        if (len != static_cast<size_t>(cfg_.slot_len)) throw Error(AA_ERR_INVALID, "Reducer::process_slot: wrong slot length");
        aa_dynamics d{};
        check(aa_condition_host(h_, slot, 1, cfg_.slot_len, cfg_.slot_len, &d));
        DynamicsOutput o;
        o.level = static_cast<DynamicLevel>(d.level);
        o.rms_db = d.rms_db;
        o.gain_db = d.gain_db;
        o.session_median_db = d.session_median_db;
        o.noise_floor_db = d.noise_floor_db;
        return o;
    }
    void reset() { check(aa_conditioner_reset(h_)); }

private:
    aa_cond_config cfg_{};
    aa_conditioner *h_ = nullptr;
};

// ---------------------------------------------------------------------------
// Tuner (src/analysis/tuner.rs): the per-frame branch of Tuner::run with Note::from_freq and
// Interval::new evaluated on the device (aa_notes_from_stable_host, aa_tuner_from_stable_host); the
// label strings of TunerOutput (tuner.rs:27-52) are formatted here from the numeric records.
// ---------------------------------------------------------------------------
enum class TuningSystem { EqualTemperament = 0, JustIntonation = 1, Pythagorean = 2 };   // tuner.rs:13-18
enum class TunerMode { MultiPitch = 0, SinglePitch = 1 };                                // tuner.rs:21-25
struct TunerOutput {
    std::string label;
    float cents = 0.0f;
    std::vector<std::string> notes;
    std::vector<float> accuracies;
};

class Tuner {
public:
    TuningSystem system = TuningSystem::EqualTemperament;
    TunerMode mode = TunerMode::MultiPitch;
    float base = 440.0f;

    static std::string note_name(const aa_note &n)                                       // theory.rs:234-246
    {
        static const char *names[] = {"C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"};
        return std::string(names[n.semis % 12]) + std::to_string(static_cast<int>(n.octave));
    }
    static const char *interval_name(uint32_t t)                                         // theory.rs:285-298, 384-386
    {
        static const char *names[] = {"Min2", "Maj2", "Min3", "Maj3", "Per4", "Aug4", "Per5", "Min6", "Maj6", "Min7",
                                      "Maj7", "Per8"};
        return names[t < 12 ? t : 11];
    }
    // One batch of frames (as popped from note_rx, tuner.rs:141-150); empty frames give an empty label.
    std::vector<TunerOutput> run(const aa_stable_pitches *frames, int64_t n_frames) const
    {
        std::vector<aa_note_record> notes(static_cast<size_t>(n_frames));
        std::vector<aa_tuner_record> rec(static_cast<size_t>(n_frames));
        std::vector<TunerOutput> out(static_cast<size_t>(n_frames));
        if (n_frames == 0) return out;
        check(aa_notes_from_stable_host(frames, n_frames, base, notes.data()));
        check(aa_tuner_from_stable_host(frames, n_frames, base, static_cast<int32_t>(system),
                                        mode == TunerMode::SinglePitch ? 1 : 0, rec.data()));
        for (int64_t f = 0; f < n_frames; ++f) {
            const aa_tuner_record &r = rec[static_cast<size_t>(f)];
            const aa_note_record &nr = notes[static_cast<size_t>(f)];
            TunerOutput &o = out[static_cast<size_t>(f)];
            o.cents = r.cents;
            if (r.kind == 1) {                                                           // tuner.rs:158-166
                o.label = note_name(nr.note[r.best]);
                o.notes.push_back(o.label);
                o.accuracies.push_back(nr.note[r.best].cents);
            } else if (r.kind == 2) {                                                    // tuner.rs:167-181
                for (uint8_t i : {r.lo, r.hi}) {
                    o.notes.push_back(note_name(nr.note[i]));
                    o.accuracies.push_back(nr.note[i].cents);
                }
                o.label = interval_name(r.interval);
            } else if (r.kind == 3) {                                                    // tuner.rs:182-189
                for (uint32_t i = 0; i < nr.n && i < AA_MAX_STABLE; ++i) {
                    o.notes.push_back(note_name(nr.note[i]));
                    o.accuracies.push_back(nr.note[i].cents);
                    o.label += (i ? " " : "") + o.notes.back();
                }
            }
        }
        return out;
    }
};

}  // namespace audio_engine_gpu

// Exercises the C++ mirror of the reference's Rust types (host/audio_engine_gpu.hpp) over libaa_gpu.so.
// argv[1] = "nodevice": expect every constructor to fail with AA_ERR_NO_DEVICE (CPU-only container).
// argv[1] = "gpu": FftProcessor round trip + the STFT / OnsetDetector worker threads on synthetic slots.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../audio-analyzer-rs_b200/host/audio_engine_gpu.hpp"

using namespace audio_engine_gpu;

static int fail(const char *msg)
{
    std::fprintf(stderr, "FAIL: %s\n", msg);
    return 1;
}

int main(int argc, char **argv)
{
    const bool gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;
    if (!gpu) {
        try {
            FftProcessor f(2048);
            return fail("FftProcessor constructed without a device");
        } catch (const Error &e) {
            if (e.code != AA_ERR_NO_DEVICE) return fail("wrong error code");
            std::printf("ok: %s\n", e.what());
        }
        try {
            FftProcessor f(1000);
            return fail("invalid length accepted");
        } catch (const Error &e) {
            if (e.code != AA_ERR_UNSUPPORTED) return fail("wrong error code for bad length");
        }
        // ---- host-side onset stamping against what the real crate logged (tests/golden/ref_log_onsets.json, from the
        // reference's output.log): a 1024-sample slot behind a 192-sample leftover is a poll of 16 frames whose
        // window-centre offsets run from -1088 to -128; the first slot of a stream gives 13 frames from -896 ----
        if (window_centre_offset(16, 0, 256, 64) != -1088 || window_centre_offset(16, 15, 256, 64) != -128 ||
            window_centre_offset(13, 0, 256, 64) != -896 || window_centre_offset(16, 7, 256, 64) != -640)
            return fail("window_centre_offset lattice");
        // `beat_pos: 0.6653333333333337, transport beat: 0.7440000000000003, ... event_samples: 15968` (onset.rs:413-419)
        // at 120 BPM / 48 kHz: output frame 17 856, latency - offset = 1 888 samples
        {
            const OnsetEvent ev = stamp_onset_at(0.7440000000000003, 17856, 120.0, 48000.0, 400, 400, 0, -1088, 0.5f);
            if (ev.output_samples != 15968 || ev.beat_position != 0.6653333333333337 || ev.raw_sample_offset != -1088)
                return fail("stamp_onset_at against the reference's log line");
            std::printf("ok: stamp_onset_at reproduces the logged beat %.16f / sample %lld\n", ev.beat_position,
                        static_cast<long long>(ev.output_samples));
        }
        return 0;
    }

    // ---- FftProcessor: DC input -> X[0] = n, rest ~0; wrong length throws (reference: panic) ----
    {
        FftProcessor f(2048);
        std::vector<float> x(2048, 1.0f);
        const Complex32 *s = f.process_forward(x.data(), x.size());
        if (std::fabs(s[0].re - 2048.0f) > 1e-2f || std::fabs(s[1].re) > 1e-2f) return fail("fft DC");
        const float *back = f.process_inverse(s, f.bins());
        if (std::fabs(back[5] / 2048.0f - 1.0f) > 1e-5f) return fail("fft inverse");
        bool threw = false;
        try { f.process_forward(x.data(), 2047); } catch (const Error &) { threw = true; }
        if (!threw) return fail("wrong length accepted");
    }
    // ---- STFT worker: 440 Hz in 1024-sample slots -> A4 pitch frames ----
    {
        SlotQueue q;
        SharedAnalysisState shared;
        std::mutex m;
        std::vector<PitchFrame> got;
        const uint32_t sr = 44100;
        for (int s = 0; s < 40; ++s) {
            std::vector<float> slot(1024);
            for (int i = 0; i < 1024; ++i)
                slot[i] = (float)(0.5 * std::sin(2.0 * M_PI * 440.0 * (double)(s * 1024 + i) / sr));
            q.push(std::move(slot));
        }
        {
            STFT stft(3);
            stft.detect_pitches(q, sr, [&](PitchFrame pf) { std::lock_guard<std::mutex> g(m); got.push_back(std::move(pf)); },
                                shared, [] { return 1.5; });
            for (int i = 0; i < 400 && !q.is_empty(); ++i) std::this_thread::sleep_for(std::chrono::milliseconds(5));
            std::this_thread::sleep_for(std::chrono::milliseconds(50));
            stft.stop();
        }
        // 40 slots -> (40960 - 2048)/512 + 1 = 77 frames; PitchTracker shows the note from frame 2 on
        if (got.size() != 76) { std::fprintf(stderr, "frames %zu\n", got.size()); return fail("STFT frame count"); }
        for (auto &pf : got) {
            if (pf.pitches.size() != 1 || std::fabs(pf.pitches[0].first - 440.196f) > 0.01f || pf.beat != 1.5)
                return fail("STFT pitch");
        }
        std::printf("ok: STFT emitted %zu frames, %.3f Hz\n", got.size(), got[0].pitches[0].first);
    }
    // ---- OnsetDetector worker: clicks every 4800 samples ----
    {
        SlotQueue q;
        SharedAnalysisState shared;
        std::mutex m;
        std::vector<OnsetEvent> ev;
        const uint32_t sr = 48000;
        for (int s = 0; s < 30; ++s) {
            std::vector<float> slot(1024, 0.0f);
            for (int i = 0; i < 1024; ++i) {
                const int n = s * 1024 + i;
                const int ph = n % 4800;
                if (n > 2000 && ph < 400) slot[i] = (float)(0.6 * std::exp(-ph / 80.0) * std::sin(2.0 * M_PI * 1500.0 * ph / sr));
            }
            q.push(std::move(slot));
        }
        TransportHooks tr;
        tr.stamp_onset = [](int64_t off, float vel) { return OnsetEvent{0.0, off, 0, vel}; };
        tr.nearest_tick_distance_beats = [](double) { return 1e9; };
        tr.get_bpm = [] { return 120.0f; };
        {
            OnsetDetector od(4);
            od.detect_onsets(tr, q, sr, [&](OnsetEvent e) { std::lock_guard<std::mutex> g(m); ev.push_back(e); }, shared);
            for (int i = 0; i < 400 && !q.is_empty(); ++i) std::this_thread::sleep_for(std::chrono::milliseconds(5));
            std::this_thread::sleep_for(std::chrono::milliseconds(50));
            od.stop();
        }
        if (ev.size() < 4 || ev.size() > 8) { std::fprintf(stderr, "onsets %zu\n", ev.size()); return fail("onset count"); }
        if (!shared.onset_pending.load()) return fail("onset_pending not set");
        std::printf("ok: OnsetDetector emitted %zu events, velocity %.2f\n", ev.size(), ev[0].velocity);
    }
    // ---- Reducer: a -20 dBFS 440 Hz tone slot by slot -> "mf", DC removed, gain creeping up ----
    {
        const float sr = 48000.0f;
        Reducer red(sr, 1024);
        DynamicsOutput last;
        float peak = 0.0f;
        for (int s = 0; s < 60; ++s) {
            std::vector<float> slot(1024);
            for (int i = 0; i < 1024; ++i)
                slot[i] = 0.05f + (float)(0.1414 * std::sin(2.0 * M_PI * 440.0 * (double)(s * 1024 + i) / sr));
            last = red.process_slot(slot.data(), slot.size());
            if (s == 59) for (float v : slot) peak = std::fmax(peak, std::fabs(v));
        }
        if (last.level != DynamicLevel::Mf || std::fabs(last.rms_db + 20.0f) > 0.1f || !(last.gain_db > 0.0f))
            return fail("Reducer dynamics");
        if (std::fabs(peak - 0.1414f) > 0.004f) return fail("Reducer did not remove the DC offset");
        bool threw = false;
        std::vector<float> bad(1000);
        try { red.process_slot(bad.data(), bad.size()); } catch (const Error &) { threw = true; }
        if (!threw) return fail("Reducer accepted a wrong slot length");
        std::printf("ok: Reducer level %s rms %.2f dB gain %.4f dB\n", to_string(last.level), last.rms_db, last.gain_db);
    }
    // ---- Tuner: single note, a perfect fifth, a triad ----
    {
        std::vector<aa_stable_pitches> fr(4);
        std::memset(fr.data(), 0, sizeof(aa_stable_pitches) * fr.size());
        fr[1].n = 1; fr[1].pitch[0] = {440.0f, 0.6f};
        fr[2].n = 2; fr[2].pitch[0] = {392.0f, 0.5f}; fr[2].pitch[1] = {261.63f, 0.7f};
        fr[3].n = 3; fr[3].pitch[0] = {261.63f, 0.5f}; fr[3].pitch[1] = {329.63f, 0.5f}; fr[3].pitch[2] = {392.0f, 0.5f};
        Tuner tuner;
        const std::vector<TunerOutput> out = tuner.run(fr.data(), (int64_t)fr.size());
        if (!out[0].label.empty()) return fail("Tuner: empty frame");
        if (out[1].label != "A4" || std::fabs(out[1].cents) > 2.0f) return fail("Tuner: A4");
        if (out[2].label != "Per5" || out[2].notes.size() != 2 || out[2].notes[0] != "C4" || out[2].notes[1] != "G4")
            return fail("Tuner: fifth");
        if (out[3].label != "C4 E4 G4") return fail("Tuner: triad");
        std::printf("ok: Tuner labels %s / %s (%.2f cents) / %s\n", out[1].label.c_str(), out[2].label.c_str(), out[2].cents,
                    out[3].label.c_str());
    }
    return 0;
}

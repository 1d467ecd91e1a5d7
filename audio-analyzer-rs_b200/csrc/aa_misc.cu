// aa_misc.cu -- small support kernels: per-clip summaries (the multi-GPU gather payload)
// and the synthetic-clip generator used by bench.py and the tests.
#include "aa_internal.h"

namespace aa {

// ---------------------------------------------------------------------------
// Per-clip summary: one warp per clip reduces that clip's feature records.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) summary_kernel(const aa_frame_features *__restrict__ feat,
                                                      int64_t n_clips, int64_t T,
                                                      aa_clip_summary *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t clip = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (clip >= n_clips) return;
    const aa_frame_features *f = feat + clip * T;
    unsigned n_pitched = 0, n_onsets = 0;
    double s_top = 0.0, s_cent = 0.0, s_flux = 0.0, s_energy = 0.0;
    float max_energy = 0.0f;
    for (int64_t t = lane; t < T; t += 32) {
        const aa_frame_features r = f[t];
        if (r.n_pitches > 0) { ++n_pitched; s_top += (double)r.pitch[0].freq; }
        if (r.flags & AA_FLAG_ONSET_DETECTED) ++n_onsets;
        s_cent += (double)r.centroid;
        s_flux += (double)r.flux;
        s_energy += (double)r.energy;
        max_energy = fmaxf(max_energy, r.energy);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_pitched += __shfl_xor_sync(0xffffffffu, n_pitched, o);
        n_onsets += __shfl_xor_sync(0xffffffffu, n_onsets, o);
        s_top += __shfl_xor_sync(0xffffffffu, s_top, o);
        s_cent += __shfl_xor_sync(0xffffffffu, s_cent, o);
        s_flux += __shfl_xor_sync(0xffffffffu, s_flux, o);
        s_energy += __shfl_xor_sync(0xffffffffu, s_energy, o);
        max_energy = fmaxf(max_energy, __shfl_xor_sync(0xffffffffu, max_energy, o));
    }
    if (lane == 0) {
        aa_clip_summary s;
        s.n_frames = (uint32_t)T;
        s.n_pitched = n_pitched;
        s.n_onsets = n_onsets;
        s.mean_top_freq = n_pitched ? (float)(s_top / (double)n_pitched) : 0.0f;
        s.mean_centroid = T ? (float)(s_cent / (double)T) : 0.0f;
        s.mean_flux = T ? (float)(s_flux / (double)T) : 0.0f;
        s.mean_energy = T ? (float)(s_energy / (double)T) : 0.0f;
        s.max_energy = max_energy;
        out[clip] = s;
    }
}

cudaError_t launch_summaries(const aa_frame_features *feat, int64_t n_clips, int64_t T,
                             aa_clip_summary *out, cudaStream_t s)
{
    if (n_clips <= 0) return cudaSuccess;
    const int wpb = 4;
    const unsigned grid = (unsigned)((n_clips + wpb - 1) / wpb);
    summary_kernel<<<grid, wpb * 32, 0, s>>>(feat, n_clips, T, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Note::from_freq (reference src/analysis/theory.rs:195-209) for every stable pitch.
// One thread per (frame, slot); base_c0 = base_freq * 2^-4.75 is computed on the host.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) notes_kernel(const aa_stable_pitches *__restrict__ stable, int64_t n_frames,
                                                    float base_c0, aa_note_record *__restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t f = idx / AA_MAX_STABLE;
    const int slot = (int)(idx % AA_MAX_STABLE);
    if (f >= n_frames) return;
    const uint32_t n = stable[f].n;
    if (slot == 0) { out[f].n = n; out[f].reserved = 0u; }
    aa_note r;
    r.semis = 0; r.octave = 0; r.reserved = 0; r.cents = 0.0f;
    if ((uint32_t)slot < n) {
        const float freq = stable[f].pitch[slot].freq;
        const float lg = __fmul_rn(log2f(__fdiv_rn(freq, base_c0)), 1200.0f);          // :198
        const float o = __fdiv_rn(__fadd_rn(lg, 50.0f), 1200.0f);                      // :199 `as u8` saturates
        r.octave = (uint8_t)(!(o > 0.0f) ? 0 : (o >= 255.0f ? 255 : (int)o));
        const float sm = fmodf(roundf(__fdiv_rn(lg, 100.0f)), 12.0f);                  // :200 `as usize` saturates
        r.semis = (uint8_t)(!(sm > 0.0f) ? 0 : (int)sm);
        float c = fmodf(lg, 100.0f);                                                   // :201
        c = c < 50.0f ? c : -__fsub_rn(100.0f, c);                                     // :202-206
        r.cents = c;
    }
    out[f].note[slot] = r;
}

cudaError_t launch_notes(const aa_stable_pitches *stable, int64_t n_frames, float base_c0, aa_note_record *out,
                         cudaStream_t s)
{
    if (n_frames <= 0) return cudaSuccess;
    const int64_t total = n_frames * AA_MAX_STABLE;
    const unsigned grid = (unsigned)((total + 255) / 256);
    notes_kernel<<<grid, 256, 0, s>>>(stable, n_frames, base_c0, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Tuner::run's per-frame branch (reference src/analysis/tuner.rs:148-193) and Interval::new
// (src/analysis/theory.rs:306-382) on the stable pitches of every frame: which note is displayed, or which
// interval lies between the two notes and how far off it is.  One thread per frame.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float note_cents(float freq, float base_c0)
{
    const float lg = __fmul_rn(log2f(__fdiv_rn(freq, base_c0)), 1200.0f);              // theory.rs:198
    const float c = fmodf(lg, 100.0f);                                                  // :201
    return c < 50.0f ? c : -__fsub_rn(100.0f, c);                                       // :202-206
}

__global__ void __launch_bounds__(256) tuner_kernel(const aa_stable_pitches *__restrict__ stable, int64_t n_frames,
                                                    float base_c0, int system, int single_pitch_mode,
                                                    aa_tuner_record *__restrict__ out)
{
    // theory.rs:317-352 (f32 constant expressions fold with the same rounding as in Rust)
    const float JUST[13] = {1.0f, 16.0f / 15.0f, 9.0f / 8.0f, 6.0f / 5.0f, 5.0f / 4.0f, 4.0f / 3.0f, 45.0f / 32.0f,
                            3.0f / 2.0f, 8.0f / 5.0f, 5.0f / 3.0f, 9.0f / 5.0f, 15.0f / 8.0f, 2.0f};
    const float PYTH[13] = {1.0f, 256.0f / 243.0f, 9.0f / 8.0f, 32.0f / 27.0f, 81.0f / 64.0f, 4.0f / 3.0f, 729.0f / 512.0f,
                            3.0f / 2.0f, 128.0f / 81.0f, 27.0f / 16.0f, 32.0f / 9.0f, 243.0f / 128.0f, 2.0f};
    const float ET[13] = {1.0f, 1.0595f, 1.1225f, 1.1892f, 1.2599f, 1.3348f, 1.4142f, 1.4983f, 1.5874f, 1.6818f,
                          1.7818f, 1.8877f, 2.0f};
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const aa_stable_pitches &sp = stable[f];
    const int n = (int)min(sp.n, (uint32_t)AA_MAX_STABLE);
    aa_tuner_record r;
    r.kind = 0; r.best = 0; r.lo = 0; r.hi = 0; r.interval = 0u; r.accuracy = 0.0f; r.cents = 0.0f;
    if (n == 1 || (n > 0 && single_pitch_mode)) {
        int b = 0;                                             // Iterator::max_by: the last maximum
        for (int i = 1; i < n; ++i)
            if (!(sp.pitch[i].score < sp.pitch[b].score)) b = i;
        r.kind = 1; r.best = (uint8_t)b;
        r.cents = note_cents(sp.pitch[b].freq, base_c0);       // tuner.rs:163-165
    } else if (n == 2) {
        const int l = sp.pitch[1].freq < sp.pitch[0].freq ? 1 : 0;   // tuner.rs:169-170
        const float f_lo = sp.pitch[l].freq, f_hi = sp.pitch[1 - l].freq;
        r.kind = 2; r.lo = (uint8_t)l; r.hi = (uint8_t)(1 - l);
        if (f_lo == 0.0f) {                                    // theory.rs:307-312
            r.interval = 11u;
        } else {
            float ratio = __fdiv_rn(f_hi, f_lo);
            while (ratio > 2.0f) ratio = __fdiv_rn(ratio, 2.0f);
            const float *tab = system == 1 ? JUST : system == 2 ? PYTH : ET;
            int idx = 0;
            float best = fabsf(__fsub_rn(ratio, tab[0]));
            for (int i = 1; i < 13; ++i) {                     // min_by: the first minimum
                const float d = fabsf(__fsub_rn(ratio, tab[i]));
                if (d < best) { best = d; idx = i; }
            }
            r.interval = idx == 0 ? 11u : (uint32_t)(idx - 1);
            r.accuracy = __fmul_rn(-logf(__fdiv_rn(tab[idx], ratio)), 1732.5f);     // theory.rs:381
        }
        r.cents = r.accuracy;                                  // tuner.rs:181
    } else if (n >= 3) {
        r.kind = 3;                                            // tuner.rs:183-189: names only
    }
    out[f] = r;
}

cudaError_t launch_tuner(const aa_stable_pitches *stable, int64_t n_frames, float base_c0, int system,
                         int single_pitch_mode, aa_tuner_record *out, cudaStream_t s)
{
    if (n_frames <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n_frames + 255) / 256);
    tuner_kernel<<<grid, 256, 0, s>>>(stable, n_frames, base_c0, system, single_pitch_mode, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Offline onset events: the frames whose gating passed (AA_FLAG_ONSET_FIRED, onset.rs:383-456
// without a live transport) compacted per clip into the OnsetEvent list the reference pushes on onset_tx, stamped
// like MusicalTransport::stamp_onset (timing.rs:311-337) with zero latencies and the clip start as time zero.
// One warp per clip: ballot + prefix over 32 frames at a time keeps the events in time order.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) onset_events_kernel(const aa_frame_features *__restrict__ feat, int64_t n_clips,
                                                           int64_t T, int n, int hop, double beats_per_sample,
                                                           int max_events, aa_onset_event *__restrict__ events,
                                                           int32_t *__restrict__ counts)
{
    const int lane = threadIdx.x & 31;
    const int64_t clip = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (clip >= n_clips) return;
    const aa_frame_features *f = feat + clip * T;
    aa_onset_event *out = events + clip * (int64_t)max_events;
    int cnt = 0;
    for (int64_t base = 0; base < T; base += 32) {
        const int64_t fr = base + lane;
        bool fired = false;
        float flux = 0.f, maxex = 0.f;
        if (fr < T) {
            fired = (f[fr].flags & AA_FLAG_ONSET_FIRED) != 0u;
            flux = f[fr].flux;
            maxex = f[fr].max_excess;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, fired);
        const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
        if (fired && pos < max_events) {
            float v = __fdiv_rn(fmaxf(flux, __fmul_rn(maxex, 5.0f)), 50.0f);       // onset.rs:388-390
            v = fminf(fmaxf(v, 0.0f), 1.0f);
            aa_onset_event e;
            e.sample_position = fr * (int64_t)hop + n / 2;                         // window centre, onset.rs:386-387
            e.beat_position = (double)e.sample_position * beats_per_sample;        // timing.rs:313-326
            e.frame = fr;
            e.velocity = v;
            e.reserved = 0u;
            out[pos] = e;
        }
        cnt += __popc(bal);
    }
    if (lane == 0) counts[clip] = cnt;
}

cudaError_t launch_onset_events(const aa_frame_features *feat, int64_t n_clips, int64_t T, int n, int hop,
                                double beats_per_sample, int max_events, aa_onset_event *events, int32_t *counts,
                                cudaStream_t s)
{
    if (n_clips <= 0) return cudaSuccess;
    const int wpb = 4;
    const unsigned grid = (unsigned)((n_clips + wpb - 1) / wpb);
    onset_events_kernel<<<grid, wpb * 32, 0, s>>>(feat, n_clips, T, n, hop, beats_per_sample, max_events, events, counts);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Synthetic clips.
// ---------------------------------------------------------------------------
__host__ __device__ inline uint64_t splitmix64(uint64_t &x)
{
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ inline double u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

struct ClipVoice {
    int k;
    double f_over_sr[4];   // f0 / sample_rate
    float amp[4];
    float phase[4][6];     // in turns
};

__device__ inline ClipVoice make_voice(uint64_t seed, int64_t clip, float sample_rate)
{
    uint64_t st = seed + (uint64_t)clip;
    ClipVoice v;
    v.k = 1 + (int)(splitmix64(st) & 3ull);
    const double a_total = 0.05 + 0.45 * u01(splitmix64(st));
    for (int i = 0; i < 4; ++i) {
        const double lf = log(55.0) + (log(1760.0) - log(55.0)) * u01(splitmix64(st));
        v.f_over_sr[i] = exp(lf) / (double)sample_rate;
        // a_total bounds the peak: K voices, partial amplitudes 1/h, sum_{h<=6} 1/h = 2.45
        v.amp[i] = (float)(a_total / ((double)v.k * 2.45));
        for (int h = 0; h < 6; ++h) v.phase[i][h] = (float)u01(splitmix64(st));
    }
    return v;
}

__global__ void __launch_bounds__(256) synth_kernel(float *__restrict__ clips, int64_t n_clips, int64_t clip_len,
                                                    int64_t clip_stride, float sample_rate, uint64_t seed)
{
    const int64_t clip = blockIdx.y;
    if (clip >= n_clips) return;
    const ClipVoice v = make_voice(seed, clip, sample_rate);
    float *dst = clips + clip * clip_stride;
    const float noise_amp = 1.0e-3f * 1.7320508f;   // uniform with RMS 1e-3 (-60 dBFS)
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < clip_len;
         n += (int64_t)gridDim.x * blockDim.x) {
        float acc = 0.0f;
        for (int i = 0; i < v.k; ++i) {
            const double base = v.f_over_sr[i] * (double)n;
#pragma unroll
            for (int h = 1; h <= 6; ++h) {
                if (v.f_over_sr[i] * h < 0.5) {
                    double ph = base * (double)h;
                    ph -= floor(ph);
                    acc += (v.amp[i] / (float)h) * sinpif(2.0f * ((float)ph + v.phase[i][h - 1]));
                }
            }
        }
        uint64_t st = seed * 0x2545F4914F6CDD1Dull + (uint64_t)clip * 0x9E3779B97F4A7C15ull + (uint64_t)n;
        const float u = (float)u01(splitmix64(st)) * 2.0f - 1.0f;
        dst[n] = acc + noise_amp * u;
    }
}

cudaError_t launch_synth(float *clips, int64_t n_clips, int64_t clip_len, int64_t clip_stride,
                         float sample_rate, uint64_t seed, cudaStream_t s)
{
    if (n_clips <= 0 || clip_len <= 0) return cudaSuccess;
    int64_t bx = (clip_len + 255) / 256;
    if (bx > 64) bx = 64;
    // grid.y is limited to 65535: loop in slabs
    for (int64_t c0 = 0; c0 < n_clips; c0 += 65535) {
        const int64_t nc = (n_clips - c0) < 65535 ? (n_clips - c0) : 65535;
        dim3 grid((unsigned)bx, (unsigned)nc);
        synth_kernel<<<grid, 256, 0, s>>>(clips + c0 * clip_stride, nc, clip_len, clip_stride, sample_rate,
                                          seed + (uint64_t)c0);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace aa

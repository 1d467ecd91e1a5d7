// aa_internal.h -- shared between the kernel translation units and the C ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aa_gpu.h"

namespace aa {

// Device tables built once per handle (host computes in f64 / with the reference's
// own f32 formulas and uploads; see aa_api.cu::build_tables).
struct Tables {
    const float2 *tw;      // [n/2]   exp(-2 pi i k / (n/2))                complex-FFT twiddles
    const float2 *pt;      // [n/4]   0.5 * exp(-2 pi i k / n)              realfft post-pass twiddles
    const float2 *win2;    // [n/2]   (w[2m], w[2m+1]) periodic Hann, f32   stft.rs:641-648
    const float  *flux_w;  // [n/2+1] 1 - k/half in f32                     onset.rs:280
};

// Persistent analyzer state of one clip / stream (used by the streaming API and by
// serial chunk hand-off; batch mode passes nullptr and starts fresh).
//   float bins[4][half]  : nfP (stft.rs:209), vol (:211), prev (:210 == onset.rs:149), nfO (onset.rs:175)
//   float scalars[8]     : [0] FluxTracker.threshold, [1] energy_ema, [2] frames seen (as float),
//                          [3] tracker track count, [4] frames_since_onset, [5] tail state valid
//   float tracks[3][32]  : freq, score, life (as float)
__host__ __device__ inline size_t state_floats(int half) { return (size_t)4 * half + 8 + 96; }

constexpr int AA_MAX_SEG = 8;

struct AnalyzeParams {
    const float *clips;
    int64_t n_clips, clip_len, clip_stride, T;
    // per-frame arrays (onset_in, mags, features, stable, dbg_*) are indexed clip * out_T + out_f0 + f: a launch that
    // covers the frames [out_f0, out_f0 + T) of longer clips (a time slice of the host pipeline) writes its records
    // where a whole-clip launch would have put them.  Whole clips: out_T = T, out_f0 = 0.
    int64_t out_T, out_f0;
    const uint8_t *onset_in;
    float *mags;
    aa_frame_features *features;
    aa_stable_pitches *stable;
    float *dbg_floor;
    uint8_t *dbg_peaks;
    float *state;            // [n_clips][state_floats(half)] or nullptr
    unsigned char *scratch;  // [grid][analyze_scratch_bytes(n)] overflow space for frames with > 256 candidates
    int grid;                // persistent CTAs: min(n_clips, num_sms * analyze_ctas_per_sm(n))
    int packed;              // the grid fills the GPU: the resident sub-blocks of an SM may be launched as one CTA
                             // (analyze_kernel's SUBS; scratch and the static item order stay per sub-block)
    unsigned long long *work_counter;   // device-wide work queue (zeroed before the launch) or nullptr = static
    // Time segments (batch mode with more clips than resident CTAs): a work item is (clip, segment), items are
    // dealt segment-major from the queue, and a segment hands the analyzer state to the next one through
    // seg_state[clip] -- the same state block the streaming API carries -- guarded by seg_flags[clip][0 / 1]
    // ("segments whose main-warp / tail-warp state has been published", zeroed before the launch).  Segments
    // get shorter towards the end of the clip, so the ragged end of the batch is one short segment, not one clip.
    int n_seg;                          // 1 = whole clips
    int seg_start[AA_MAX_SEG + 1];      // frame range of segment s: [seg_start[s], seg_start[s + 1])
    float *seg_state;                   // [n_clips][state_floats(half)]
    unsigned *seg_flags;                // [n_clips][2]
    Tables tab;
    int n, hop, half;
    float bin_width, min_freq, max_freq;
    float global_floor;      // stft.rs:323-324
    int min_bin, max_bin;    // stft.rs:454-455
    uint32_t features_mask;
    // streaming (one clip, one CTA): tail warp w stores done_value to done_flag[w] (mapped host memory, system-scope
    // fence first) after the records of its last frame, so the host can spin on the words instead of synchronising
    // the stream (analyze_tail_warps() words)
    unsigned long long *done_flag;
    unsigned long long done_value;
};

cudaError_t launch_analyze(const AnalyzeParams &p, cudaStream_t s);
cudaError_t analyze_check_word(unsigned *word, cudaStream_t s);   // -DAA_CHECKED builds: violated index checks
size_t      analyze_smem_bytes(int n);
int         analyze_threads(int n);
int         analyze_ctas_per_sm(int n);
size_t      analyze_scratch_bytes(int n);

cudaError_t launch_fft_forward(int n, const Tables &tab, const float *in, int64_t batch, float *out,
                               int num_sms, cudaStream_t s);
cudaError_t launch_fft_inverse(int n, const Tables &tab, const float *spec, int64_t batch, float *out,
                               int num_sms, cudaStream_t s);

cudaError_t launch_summaries(const aa_frame_features *feat, int64_t n_clips, int64_t T,
                             aa_clip_summary *out, cudaStream_t s);
cudaError_t launch_yin(const float *clips, int64_t n_clips, int64_t clip_stride, int64_t T, int n, int hop,
                       int min_lag, int max_lag, float threshold, int32_t *lag_out, float *cmnd_out,
                       int num_sms, cudaStream_t s);
cudaError_t launch_notes(const aa_stable_pitches *stable, int64_t n_frames, float base_c0,
                         aa_note_record *out, cudaStream_t s);
// Conditioning chain (aa_cond.cu).  Coefficients are computed on the host with the reference's own f32 formulas
// (mod.rs:357-418, dynamics.rs:164-189) so that the device arithmetic starts from identical bits.
struct CondParams {
    float hp[5], lp[5];              // b0 b1 b2 a1 a2, normalised by a0
    float gate_threshold_linear, release_coeff;
    int32_t gate_hold_samples;
    float target_db, max_boost_db, smooth_alpha, silence_decay_alpha, active_snr_db, bootstrap_floor_db;
    int32_t slot_len;
};
int         analyze_tail_warps();
size_t      cond_agc_state_floats();
cudaError_t launch_cond_filter_gate(float *clips, int64_t n_clips, int64_t clip_stride, int64_t n_slots,
                                    const CondParams &p, float *stats, float *carry, cudaStream_t s);
cudaError_t launch_cond_agc(const float *stats, int64_t n_clips, int64_t n_slots, const CondParams &p, float *state,
                            int carry, float *gains, aa_dynamics *dyn, cudaStream_t s);
cudaError_t launch_cond_apply_gain(float *clips, int64_t n_clips, int64_t clip_stride, int64_t n_slots, int slot_len,
                                   const float *gains, int num_sms, cudaStream_t s);

cudaError_t launch_onset_events(const aa_frame_features *feat, int64_t n_clips, int64_t T, int n, int hop,
                                double beats_per_sample, int max_events, aa_onset_event *events, int32_t *counts,
                                cudaStream_t s);
cudaError_t launch_ingest(const void *pcm, int format, int channels, int64_t n_clips, int64_t clip_len,
                          int64_t in_stride, int64_t out_stride, float *out, int num_sms, cudaStream_t s);
cudaError_t launch_tuner(const aa_stable_pitches *stable, int64_t n_frames, float base_c0, int system,
                         int single_pitch_mode, aa_tuner_record *out, cudaStream_t s);
cudaError_t launch_synth(float *clips, int64_t n_clips, int64_t clip_len, int64_t clip_stride,
                         float sample_rate, uint64_t seed, cudaStream_t s);

}  // namespace aa

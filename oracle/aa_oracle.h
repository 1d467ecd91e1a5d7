/*
 * aa_oracle.h -- CPU restatement of the audio-analyzer-rs frame-analysis path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker or the timed CPU baseline.
 *
 * PARITY UNPINNED.  The reference is a Rust crate whose FFT arithmetic lives in
 * the un-vendored crates realfft 3.5.0 / rustfft 6.4.1 (Cargo.lock:8799-8802,
 * 9169-9172); no Rust toolchain exists in the build image, and the reference has
 * no test, fixture or golden vector on this path.  This
 * file therefore restates the reference source line by line (citations below are
 * relative to /root/reference) and is cross-checked against an independent
 * numpy restatement (tests/golden/make_golden.py) and a float64 DFT -- not
 * against outputs of the reference binary.  What the reference tree does hold is
 * replayed on it: the known answers of its own unit tests (theory.rs:405-448,
 * 545-604; timing.rs:728-771) and the one log of a live onset-detector run it
 * ships (output.log -> tests/golden/ref_log_onsets.json: stamp_onset to the last
 * printed digit, the window-centre offset lattice, the onset gates, the re-fire
 * guard).  None of that reaches the spectra or the pitch lists: PARITY UNPINNED
 * stands for rows a1-a12 until tools/rust_golden has been run.
 *
 * Arithmetic rules: every scalar the reference holds in f32 is held in `float`
 * here, operations are performed in the reference's order, and the file must be
 * compiled with -ffp-contract=off (no FMA contraction; Rust never contracts).
 * ln/log2/cos go through libm (logf/log2f/cosf), as Rust's f32 methods do on
 * Linux.
 */
#ifndef AA_ORACLE_H
#define AA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* feature-enable bits (same numeric values as include/aa_gpu.h) */
#define AAO_FEAT_PITCH    1u  /* stft.rs:320-381  adaptive floor + extract_pitches  */
#define AAO_FEAT_ONSET    2u  /* onset.rs:261-357 flux / burst / EMA / FluxTracker   */
#define AAO_FEAT_CENTROID 4u  /* NEW (no reference): spectral centroid               */
#define AAO_FEAT_TRACKER  8u  /* stft.rs:45-116   PitchTracker hysteresis            */

/* flag bits in aao_features.flags */
#define AAO_FLAG_FLUX_ONSET     1u  /* FluxTracker::update() returned true  onset.rs:355 */
#define AAO_FLAG_BURST_ONSET    2u  /* max_excess > 3 && count >= 3         onset.rs:356 */
#define AAO_FLAG_ONSET_DETECTED 4u  /* both                                 onset.rs:357 */
#define AAO_FLAG_ENERGY_RISING  8u  /* energy > ema * 1.5 (post-update ema) onset.rs:373 */
#define AAO_FLAG_ONSET_FIRED    16u  /* offline gating: detected && rising && frames_since_onset >= 3
                                        (onset.rs:403,535-539; tick guard and calibration need a live
                                        transport and are treated as passed) */

#define AAO_MAX_NOTES   8   /* stft.rs:452 */
#define AAO_MAX_STABLE 16   /* bound on displayed PitchTracker tracks, see aa_oracle.c */

typedef struct aao_config {
    int32_t  n;               /* window size (stft.rs:170 = 2048, onset.rs:122 = 256)   */
    int32_t  hop;             /* hop size    (stft.rs:169 = 512,  onset.rs:123 = 64)    */
    float    sample_rate;     /* sr          (stft.rs:160)                              */
    float    min_freq;        /* 24.0        (stft.rs:173)                              */
    float    max_freq;        /* 10000.0     (stft.rs:174)                              */
    float    noise_floor_db;  /* DynamicsOutput.noise_floor_db, default -96 (dynamics.rs:100) */
    uint32_t features;        /* AAO_FEAT_* */
} aao_config;

/* One record per frame; 96 bytes; identical layout to aa_frame_features. */
typedef struct aao_features {
    uint32_t n_pitches;                     /* raw pitches from extract_pitches, <= 8      */
    struct { float freq, score; } pitch[AAO_MAX_NOTES];
    float    flux;                          /* onset.rs:261-291, zeroed if burst<2 (:337)  */
    float    energy;                        /* onset.rs:276                                */
    float    centroid;                      /* NEW                                         */
    uint32_t burst_count;                   /* onset.rs:312-319                            */
    float    max_excess;                    /* onset.rs:329-331                            */
    uint32_t flags;                         /* AAO_FLAG_*                                  */
    float    energy_ema;                    /* onset.rs:350 (value after this frame)       */
} aao_features;

/* PitchTracker output for one frame; 136 bytes; identical to aa_stable_pitches. */
typedef struct aao_stable {
    uint32_t n;
    uint32_t reserved;
    struct { float freq, score; } pitch[AAO_MAX_STABLE];
} aao_stable;

/* Per-frame diagnostics of extract_pitches used by the parity tests. */
typedef struct aao_pitch_diag {
    int32_t  n_peaks;            /* len(peak_bins)                      stft.rs:462-469 */
    int32_t  n_scored;           /* peaks that passed the 5x floor gate stft.rs:479     */
    int32_t  n_candidates;       /* after the 0.5*max cutoff            stft.rs:553-562 */
    int32_t  n_out;              /* returned pitches                                    */
    int32_t  out_bins[AAO_MAX_NOTES];   /* integer bin of each returned pitch           */
    /* Smallest perturbation that would flip any discrete decision taken on a
     * transcendental-derived float: in units of frac_bin for the comb-window integer
     * boundaries (classes 1, 2, 11), relative for scores / ratios / frequencies
     * (cutoff 3, ghost test 4-6, sort order 7, dedup 8, range 9-10).  libm logf/log2f
     * and CUDA logf/log2f may differ by 1 ulp, so frames with min_margin below ~1e-5
     * are the documented near-ties.  Decisions on raw magnitudes are exact given
     * identical magnitudes and are not included.  */
    float    min_margin;
    int32_t  margin_src;         /* which decision class produced min_margin (1..10, aa_oracle.c) */
    /* Magnitude-perturbation margin of the CANDIDATE-level decisions of extract_pitches (everything after
     * the peak pick): the smallest uniform perturbation |dm| of the frame's magnitudes (absolute, first
     * order, worst-case signs; adaptive floor values count as magnitudes, clamped ones as constants) that
     * could flip one of them.  Two FFT implementations differ by some measured |dm| per frame; an
     * end-to-end pitch-list difference on identical peak masks is a documented near-tie iff cand_eps is
     * below that.  Classes (cand_src): 104 the 5x-floor gate stft.rs:479, 111 delta clamp :492, 112 comb
     * stop :506, 113 / 114 comb window low / high boundary :509-510, 115 best_mag :516, 116 the 15x-floor
     * test :536, 117 score sign :547, 118 cutoff :556, 119-121 ghost test :575-580, 122 sort :592, 123
     * dedup :600, 124 range :613.  The per-BIN decisions (peak pick :465, pitch-floor branches :351-355,
     * onset burst / floor onset.rs:318,323) are analysed per bin from the state taps (tests/parity.py).  */
    float    cand_eps;
    int32_t  cand_src;
} aao_pitch_diag;

/* ---- a1: window (stft.rs:641-648 == onset.rs:549-556) ------------------------- */
void aao_hann_window(int n, float *w);

/* ---- a3: FftProcessor (dsp/fft.rs:14-35,66-71) --------------------------------- */
typedef struct aao_fft aao_fft;
aao_fft *aao_fft_create(int n);                 /* n: power of two >= 4 */
void     aao_fft_destroy(aao_fft *p);
/* process_forward: `time` (n floats) is clobbered, `spec` receives n/2+1 interleaved
 * (re,im) pairs, unnormalised; spec[0].im == spec[n/2].im == 0.  */
void     aao_fft_forward(aao_fft *p, float *time, float *spec);
/* float64 truth: same transform evaluated in double. */
void     aao_rdft_f64(const double *x, int n, double *spec);

/* ---- a4: magnitudes (stft.rs:314-318, num_complex norm == hypot) -------------- */
void  aao_magnitudes(const float *spec, int half, float *mags);
/* ---- a5: global floor (stft.rs:322-324, onset.rs:300-302) --------------------- */
float aao_global_floor(float noise_floor_db, int half);

/* ---- a6: adaptive per-bin floor (stft.rs:209-224, 326-367) -------------------- */
typedef struct aao_pitch_floor aao_pitch_floor;
aao_pitch_floor *aao_pitch_floor_create(int half);
void aao_pitch_floor_destroy(aao_pitch_floor *s);
void aao_pitch_floor_reset(aao_pitch_floor *s);
void aao_pitch_floor_update(aao_pitch_floor *s, const float *mags, float global_floor,
                            float *effective_floor);

/* ---- a7: extract_pitches (stft.rs:443-620) ------------------------------------ */
/* out_pairs: up to 8 (freq, score) pairs; peak_mask (optional): half bytes, 1 at
 * each peak bin; diag optional.  Returns number of pitches. */
int aao_extract_pitches(const float *mags, int half, float bin_width, float min_freq,
                        float max_freq, const float *noise_floor, float *out_pairs,
                        uint8_t *peak_mask, aao_pitch_diag *diag);

/* ---- a8: PitchTracker (stft.rs:19-117) ---------------------------------------- */
typedef struct aao_tracker aao_tracker;
aao_tracker *aao_tracker_create(void);
void aao_tracker_destroy(aao_tracker *t);
void aao_tracker_reset(aao_tracker *t);
/* returns number of stable pitches written to out_pairs (<= max_out pairs);
 * the true count is returned even if it exceeds max_out. */
int  aao_tracker_process(aao_tracker *t, const float *raw_pairs, int n_raw, int onset,
                         float *out_pairs, int max_out);

/* ---- a10-a12: onset frame body (onset.rs:149-186, 261-357, 47-84) ------------- */
typedef struct aao_onset aao_onset;
aao_onset *aao_onset_create(int half);
void aao_onset_destroy(aao_onset *s);
void aao_onset_reset(aao_onset *s);
/* fills flux, energy, burst_count, max_excess, flags, energy_ema of *out */
void aao_onset_frame(aao_onset *s, const float *mags, float global_floor, aao_features *out);

/* ---- f2 (next row): Note::from_freq, analysis/theory.rs:195-209.  PINNED by the reference's own
 * known-answer tests (theory.rs:405-448): 440 -> A4 with |cents| < 2, 261.626 -> C4, C#4, cents in
 * [-50, 50].  semis: 0 = C .. 11 = B. */
void aao_note_from_freq(float freq, float base_freq, int *octave, int *semis, float *cents);

/* ---- a13: spectral centroid (NEW, self-defined): sum(k*bw*m)/sum(m), f64 acc --- */
float aao_centroid(const float *mags, int half, float bin_width);

/* ---- a14: YIN-style lag search (NEW, self-defined; see aa_oracle.c) ------------ */
/* Cumulative-mean-normalised difference over lags [min_lag, max_lag] on the raw
 * (unwindowed) frame, evaluated in float64; returns the chosen integer lag (0 if
 * none) and writes d'(lag) for lag in [0, max_lag] if cmnd != NULL. */
int aao_yin_lag(const float *frame, int n, int min_lag, int max_lag, float threshold,
                double *cmnd);

/* ---- frame loop (stft.rs:273-438, onset.rs:244-543) on one offline clip ------- */
/* T = (len - n)/hop + 1 frames (0 if len < n); frame t covers [t*hop, t*hop+n).
 * Optional outputs may be NULL.  onset_in (optional, T bytes) feeds the
 * PitchTracker's `onset` argument (stft.rs:387-390); NULL means never.
 * If mags_in != NULL the FFT stage is skipped and the given magnitudes
 * ([T][half]) are used instead (stage-isolated parity tests).
 * Returns T.  */
int64_t aao_analyze_clip(const aao_config *cfg, const float *samples, int64_t len,
                         const float *mags_in, const uint8_t *onset_in,
                         float *mags_out,            /* [T][half]            */
                         float *floor_out,           /* [T][half] eff. floor */
                         uint8_t *peak_mask_out,     /* [T][half]            */
                         aao_features *feat_out,     /* [T]                  */
                         aao_stable *stable_out,     /* [T]                  */
                         aao_pitch_diag *diag_out);  /* [T]                  */

int64_t aao_num_frames(int64_t len, int n, int hop);

/* The same frame loop with every tap of the parity tests; all pointers optional.  pitch_nf / pitch_vol /
 * onset_nf: the raw recurrent per-bin state AFTER the frame's update (noise_floor_per_bin stft.rs:209,
 * bin_volatility :211, the onset detector's noise_floor_per_bin onset.rs:175), [T][half] each. */
typedef struct aao_taps {
    float          *mags;        /* [T][half]            */
    float          *floor;       /* [T][half] eff. floor */
    uint8_t        *peak_mask;   /* [T][half]            */
    aao_features   *features;    /* [T]                  */
    aao_stable     *stable;      /* [T]                  */
    aao_pitch_diag *diag;        /* [T]                  */
    float          *pitch_nf;    /* [T][half]            */
    float          *pitch_vol;   /* [T][half]            */
    float          *onset_nf;    /* [T][half]            */
} aao_taps;
int64_t aao_analyze_clip_ex(const aao_config *cfg, const float *samples, int64_t len,
                            const float *mags_in, const uint8_t *onset_in, const aao_taps *taps);

/* Clip-parallel driver for the timed CPU baseline: clips are contiguous,
 * clip_len samples each; n_threads pthreads each take whole clips.  Only
 * feature records (and optionally magnitudes) are produced. */
int64_t aao_analyze_batch(const aao_config *cfg, const float *clips, int64_t n_clips,
                          int64_t clip_len, int n_threads, float *mags_out,
                          aao_features *feat_out, aao_stable *stable_out);

/* ---- input conditioning chain: aa_oracle_cond.c ------------------------
 * Reducer-thread HPF / LPF biquads + envelope gate (mod.rs:351-487) and
 * DynamicsTracker::process_slot (dynamics.rs:194-360). */
typedef struct aao_cond_params {
    float   hp[5], lp[5];            /* b0 b1 b2 a1 a2, normalised by a0   mod.rs:357-388 */
    float   gate_threshold_linear;   /* 10^(-60/20)                         mod.rs:401 */
    float   release_coeff;           /* exp(-1/(0.040 sr))                  mod.rs:408 */
    int32_t gate_hold_samples;       /* (0.020 sr) as usize                 mod.rs:413 */
    float   target_db, max_boost_db, smooth_alpha, silence_decay_alpha;   /* dynamics.rs:156-186 */
    float   active_snr_db, bootstrap_floor_db;                            /* dynamics.rs:188-189 */
    int32_t slot_len;                /* samples per slot (mod.rs: 1024) */
} aao_cond_params;

/* DynamicsOutput (dynamics.rs:78-91) of one slot + the gain that was applied; 32 bytes, identical
 * layout to aa_dynamics.  level: 0 Silence, 1 ppp, 2 pp, 3 p, 4 mp, 5 mf, 6 f, 7 ff, 8 fff. */
typedef struct aao_dynamics {
    int32_t  level;
    float    rms_db, gain_db, session_median_db, noise_floor_db;
    float    effective_gain;         /* min(current_gain_linear, 0.97 / peak)  dynamics.rs:327-328 */
    uint32_t flags;                  /* 1 is_active, 2 is_broadband, 4 is_playing */
    uint32_t reserved;
} aao_dynamics;

typedef struct aao_cond aao_cond;
void      aao_cond_params_init(aao_cond_params *p, float sample_rate, int slot_len);
aao_cond *aao_cond_create(const aao_cond_params *p);
void      aao_cond_destroy(aao_cond *c);
void      aao_cond_reset(aao_cond *c);
void      aao_cond_filter_gate(aao_cond *c, float *slot, int len);
void      aao_cond_agc(aao_cond *c, float *slot, int len, aao_dynamics *out, int apply);
/* one onset event of an offline clip; 32 bytes, identical layout to aa_onset_event */
typedef struct aao_onset_event {
    double   beat_position;
    int64_t  sample_position;
    int64_t  frame;
    float    velocity;
    uint32_t reserved;
} aao_onset_event;
int64_t   aao_onset_events(const aao_features *feat, int64_t T, int n, int hop, float sample_rate, float bpm,
                           int64_t max_events, aao_onset_event *out);
void      aao_onset_gates(float *flux_multiplier, float *flux_threshold_floor, float *excess_gate, uint32_t *count_gate,
                          uint32_t *refire_frames);
void      aao_stamp_onset(double current_beats, int64_t output_frames, float bpm, float sample_rate, int64_t input_lat,
                          int64_t output_lat, int64_t calibration, int64_t sample_offset, double *beat_position,
                          int64_t *output_samples);
int       aao_interval(float f_lo, float f_hi, int system, float *accuracy);
void      aao_tuner_frame(const float *pairs, int n, int system, int single_pitch_mode, int *kind, int *best, int *lo,
                          int *hi, int *interval, float *accuracy);
void      aao_ingest(const void *pcm, int format, int channels, int64_t n_frames, float *out);
int64_t   aao_cond_clip(const aao_cond_params *p, float *samples, int64_t len, aao_dynamics *dyn, int agc);

#ifdef __cplusplus
}
#endif
#endif /* AA_ORACLE_H */

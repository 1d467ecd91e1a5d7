/*
 * aa_gpu.h -- C ABI of libaa_gpu.so: the B200 (sm_100a) implementation of
 * audio-analyzer-rs's frame-analysis hot path.
 *
 * This is the drop-in boundary.  The reference (a Rust crate, paths below are
 * relative to its root) has no FFI of its own on this path; each entry point
 * here names the Rust item it replaces, and INTEGRATION.md shows the Rust
 * `extern "C"` binding a maintainer would add.
 *
 * Conventions
 *   - Plain pointers and sizes only; no C++ or torch types.
 *   - Every call returns aa_status: 0 = AA_OK, < 0 = error.  The library never
 *     aborts; aa_last_error() returns a thread-local message for the last
 *     failing call (the reference panics via unwrap() instead, fft.rs:69,99).
 *   - A handle is single-threaded (the reference's `&mut self` contract);
 *     distinct handles may be used concurrently from distinct threads.
 *   - `stream` arguments are a cudaStream_t passed as void* (NULL = the CUDA
 *     default stream, as everywhere in CUDA).  *_device entry points are
 *     asynchronous on that stream and ordered with the caller's other work on
 *     it; *_host entry points return when the results are in host memory.
 *   - There is no CPU fallback: with no usable sm_100 (compute capability 10.0)
 *     device every create call fails with AA_ERR_NO_DEVICE.
 *   - Supported geometry, narrower than the reference's: window sizes 256, 512,
 *     1024, 2048, 4096 and hop == n / 4 (the reference's two geometries are
 *     2048 / 512, stft.rs:169-170, and 256 / 64, onset.rs:122-123; its
 *     FftProcessor::new plans ANY length).  Anything else is AA_ERR_UNSUPPORTED
 *     at create time.
 *   - aa_set_device selects the GPU for handles created afterwards BY THE CALLING
 *     THREAD (thread-local).
 */
#ifndef AA_GPU_H
#define AA_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define AA_API __declspec(dllexport)
#else
#define AA_API __attribute__((visibility("default")))
#endif

typedef int32_t aa_status;
#define AA_OK               0
#define AA_ERR_INVALID     -1   /* bad argument (length mismatch, null pointer, ...)      */
#define AA_ERR_UNSUPPORTED -2   /* geometry the kernels are not built for                  */
#define AA_ERR_NO_DEVICE   -3   /* no CUDA device / not sm_100                             */
#define AA_ERR_CUDA        -4   /* CUDA runtime error, text in aa_last_error()             */
#define AA_ERR_OVERFLOW    -5   /* streaming ring overrun / output array too small         */

AA_API const char *aa_last_error(void);
AA_API int32_t     aa_version(void);                     /* 10000*major + 100*minor + patch */
AA_API aa_status   aa_device_count(int32_t *count);
AA_API aa_status   aa_set_device(int32_t device);        /* device used by handles the CALLING THREAD creates afterwards (thread-local) */

/* pinned host memory for the *_host entry points (pageable memory also works, slower) */
AA_API aa_status aa_host_alloc(size_t bytes, void **out);
AA_API aa_status aa_host_free(void *p);
AA_API aa_status aa_device_alloc(size_t bytes, void **out);
AA_API aa_status aa_device_free(void *p);
AA_API aa_status aa_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes);
AA_API aa_status aa_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes);
AA_API aa_status aa_device_synchronize(void);

/* ------------------------------------------------------------------------- *
 * FftProcessor  (src/dsp/fft.rs:6-41)
 *   FftProcessor::new(len)                      -> aa_fft_create
 *   process_forward(&mut self, &mut [f32])      -> aa_fft_forward      (fft.rs:33)
 *   process_inverse(&mut self, &mut [Complex])  -> aa_fft_inverse      (fft.rs:39)
 * Spectra are n/2+1 interleaved (re,im) f32 pairs, unnormalised, as realfft
 * returns them.  `batch` frames are transformed per call (the reference does
 * one); frame b starts at in + b*n and its spectrum at out + b*2*(n/2+1).
 * Unlike the reference the input is not clobbered.
 * ------------------------------------------------------------------------- */
typedef struct aa_fft aa_fft;
AA_API aa_status aa_fft_create(int32_t n, aa_fft **out);            /* n in {256,512,1024,2048,4096} */
AA_API aa_status aa_fft_destroy(aa_fft *h);
AA_API int32_t   aa_fft_len(const aa_fft *h);
AA_API aa_status aa_fft_forward(aa_fft *h, const float *in_host, int64_t batch, float *out_host);
AA_API aa_status aa_fft_forward_device(aa_fft *h, const float *in_dev, int64_t batch,
                                       float *out_dev, void *stream);
AA_API aa_status aa_fft_inverse(aa_fft *h, const float *spec_host, int64_t batch, float *out_host);
AA_API aa_status aa_fft_inverse_device(aa_fft *h, const float *spec_dev, int64_t batch,
                                       float *out_dev, void *stream);

/* ------------------------------------------------------------------------- *
 * Frame analysis: the bodies of STFT::detect_pitches (src/audio_io/stft.rs:
 * 273-438) and OnsetDetector::detect_onsets (src/analysis/onset.rs:244-357).
 * ------------------------------------------------------------------------- */
#define AA_FEAT_PITCH    1u  /* stft.rs:320-381  adaptive floor + extract_pitches (:443-620) */
#define AA_FEAT_ONSET    2u  /* onset.rs:261-357 flux / burst floor / EMA / FluxTracker      */
#define AA_FEAT_CENTROID 4u  /* NEW (no reference): spectral centroid in Hz                  */
#define AA_FEAT_TRACKER  8u  /* stft.rs:45-116   PitchTracker hysteresis (needs PITCH)       */
#define AA_FEAT_ALL      15u

#define AA_FLAG_FLUX_ONSET     1u   /* FluxTracker::update() == true       onset.rs:355 */
#define AA_FLAG_BURST_ONSET    2u   /* max_excess > 3 && burst_count >= 3  onset.rs:356 */
#define AA_FLAG_ONSET_DETECTED 4u   /* both                                onset.rs:357 */
#define AA_FLAG_ENERGY_RISING  8u   /* energy > 1.5 * energy_ema           onset.rs:373 */
#define AA_FLAG_ONSET_FIRED    16u   /* offline gating of onset.rs:403,535-539: detected && rising &&
                                       frames_since_onset >= 3 (no metronome ticks, calibration done).
                                       Velocity (onset.rs:388-389) = clamp(max(flux, 5*max_excess)/50, 0, 1). */

#define AA_MAX_NOTES   8    /* stft.rs:452 MAX_NOTES */
#define AA_MAX_STABLE 16    /* bound on displayed PitchTracker tracks */

typedef struct aa_config {
    int32_t  n;               /* window: 256,512,1024,2048,4096 (stft.rs:170 = 2048, onset.rs:122 = 256) */
    int32_t  hop;             /* hop: must be n/4, the reference geometry (stft.rs:169, onset.rs:123)      */
    float    sample_rate;     /* `sr` argument of detect_pitches (stft.rs:160)                           */
    float    min_freq;        /* MIN_FREQ = 24.0     (stft.rs:173) */
    float    max_freq;        /* MAX_FREQ = 10000.0  (stft.rs:174) */
    float    noise_floor_db;  /* DynamicsOutput.noise_floor_db, default -96 (dynamics.rs:93-103)         */
    uint32_t features;        /* AA_FEAT_* */
} aa_config;

AA_API void aa_config_default_pitch(aa_config *cfg, float sample_rate);  /* 2048/512, PITCH|TRACKER      */
AA_API void aa_config_default_onset(aa_config *cfg, float sample_rate);  /* 256/64,   ONSET              */

/* One record per frame, 96 bytes. */
typedef struct aa_pitch { float freq, score; } aa_pitch;
typedef struct aa_frame_features {
    uint32_t n_pitches;                 /* raw pitches of extract_pitches, <= 8               */
    aa_pitch pitch[AA_MAX_NOTES];       /* (Hz, score), descending score                      */
    float    flux;                      /* onset.rs:261-291, zeroed when burst_count < 2      */
    float    energy;                    /* onset.rs:276 sum of magnitudes                     */
    float    centroid;                  /* Hz                                                 */
    uint32_t burst_count;               /* onset.rs:312-319                                   */
    float    max_excess;                /* onset.rs:329-331                                   */
    uint32_t flags;                     /* AA_FLAG_*                                          */
    float    energy_ema;                /* onset.rs:350 after this frame                      */
} aa_frame_features;

/* PitchTracker::process output (stft.rs:45-116), 136 bytes; what `note_tx` carries (stft.rs:431-434). */
typedef struct aa_stable_pitches {
    uint32_t n;
    uint32_t reserved;
    aa_pitch pitch[AA_MAX_STABLE];
} aa_stable_pitches;

/* Per-clip summary (NEW; the payload of the multi-GPU gather), 32 bytes. */
typedef struct aa_clip_summary {
    uint32_t n_frames;
    uint32_t n_pitched;        /* frames with n_pitches > 0                      */
    uint32_t n_onsets;         /* frames with AA_FLAG_ONSET_DETECTED             */
    float    mean_top_freq;    /* mean of pitch[0].freq over pitched frames, Hz  */
    float    mean_centroid;
    float    mean_flux;
    float    mean_energy;
    float    max_energy;
} aa_clip_summary;

/* Output arrays; any pointer may be NULL (= not produced).  Device pointers for
 * aa_analyze_device, host pointers for aa_analyze_host.  Frames of clip c are at
 * index c*T + t with T = aa_num_frames(clip_len). */
typedef struct aa_outputs {
    float             *mags;       /* [n_clips*T][n/2+1] magnitude spectra (stft.rs:314-318) */
    aa_frame_features *features;   /* [n_clips*T]                                           */
    aa_stable_pitches *stable;     /* [n_clips*T]                                           */
    aa_clip_summary   *summaries;  /* [n_clips]  (needs `features`)                         */
    /* parity-test taps, NULL in production */
    float             *dbg_floor;  /* [n_clips*T][n/2+1] effective pitch floor (stft.rs:365-367) */
    uint8_t           *dbg_peaks;  /* [n_clips*T][n/2+1] is_peak (stft.rs:461-469)               */
} aa_outputs;

typedef struct aa_analyzer aa_analyzer;
AA_API aa_status aa_analyzer_create(const aa_config *cfg, aa_analyzer **out);
AA_API aa_status aa_analyzer_destroy(aa_analyzer *h);
/* T = (clip_len - n)/hop + 1, 0 if clip_len < n: frame t covers samples [t*hop, t*hop+n),
 * the offline reading of `while available_samples >= window_size` (stft.rs:273,436). */
AA_API int64_t   aa_num_frames(const aa_config *cfg, int64_t clip_len);
/* How aa_analyze_* cuts the T frames of every clip into time segments when n_clips clips are dealt to
 * `resident_ctas` persistent CTAs (NEW, no reference: the reference analyses one stream).  A segment hands the
 * analyzer state (what stft.rs:209-211, onset.rs:149-200 and PitchTracker keep between frames) to the next one
 * through HBM, so results do not depend on the plan.  Writes starts[0..n] (starts[0] = 0, starts[n] = T, at
 * most AA_MAX_SEGMENTS + 1 entries) and returns n >= 1; n == 1 means whole clips.  Pure host arithmetic. */
#define AA_MAX_SEGMENTS 8
AA_API int       aa_plan_segments(int64_t T, int64_t n_clips, int resident_ctas, int32_t *starts);

/* clips_dev: n_clips clips, clip c starting at clips_dev + c*clip_stride (samples,
 * multiple of 4; clips may overlap, which is how hop-aligned chunks of a long stream
 * with a window halo are expressed).  Analyzer state (floors, trackers) starts fresh
 * for every clip.  onset_in_dev (optional, [n_clips*T] bytes) is PitchTracker's
 * `onset` argument per frame (the reference's onset_pending flag, stft.rs:387).
 * A handle owns the work queue and the segment hand-off buffers of its launches: calls on one handle
 * must not overlap on the device (one stream per handle, or several handles). */
AA_API aa_status aa_analyze_device(aa_analyzer *h, const float *clips_dev, int64_t n_clips,
                                   int64_t clip_len, int64_t clip_stride,
                                   const uint8_t *onset_in_dev, const aa_outputs *out_dev,
                                   void *stream);
/* Stateful chunks of long streams (BASELINE cfg4; the reference's analyzers carry their state forever:
 * stft.rs:209-212, 338-363, onset.rs:149-200, PitchTracker stft.rs:19-117).  aa_state_floats: size of one analyzer
 * state block in floats (per-bin floors, volatility, previous magnitudes, FluxTracker / EMA scalars, tracks;
 * 4*(n/2+1) + 104).  aa_analyze_device_carry is aa_analyze_device where clip c STARTS from the state block
 * state_dev + c*aa_state_floats (an all-zero block is a fresh analyzer) and leaves its final state there, so
 * hop-aligned chunks of one stream (chunk c+1 = the samples from frame f1 on, no warm-up) chained through one
 * state block -- on one GPU or handed from rank to rank as a ~33 KB message -- reproduce the unchunked run bit
 * for bit.  Any number of clips (streams) per call; the launch is asynchronous on `stream`. */
AA_API int64_t   aa_state_floats(const aa_config *cfg);
AA_API aa_status aa_analyze_device_carry(aa_analyzer *h, const float *clips_dev, int64_t n_clips,
                                         int64_t clip_len, int64_t clip_stride,
                                         const uint8_t *onset_in_dev, const aa_outputs *out_dev,
                                         float *state_dev, void *stream);
/* Same, host buffers: H2D, kernels and D2H are pipelined -- over time slices of every clip for large batches (the
 * analyzer state is carried from slice to slice in device memory, outputs are byte-identical to a whole-clip launch),
 * over clip groups for small ones. */
AA_API aa_status aa_analyze_host(aa_analyzer *h, const float *clips_host, int64_t n_clips,
                                 int64_t clip_len, int64_t clip_stride,
                                 const uint8_t *onset_in_host, const aa_outputs *out_host);
/* Number of kernel launches issued by the last aa_analyze_* call on this handle. */
AA_API int64_t   aa_analyzer_last_launches(const aa_analyzer *h);

/* Device sample formats.  The reference's input callback (src/audio_io/mod.rs:657-716, 765-792) accepts f32, i16
 * and u16 devices with any channel count, converts every sample with cpal / dasp_sample's `to_sample::<f32>()`
 * (i16: s / 32768, u16: (s - 32768) / 32768) and mixes the first min(channels, 2) channels of a frame:
 * mixed = (0.0 + s0 [+ s1]) / channels_to_use.  aa_ingest_device does exactly that on the device;
 * aa_analyze_host_pcm is aa_analyze_host on interleaved PCM (clip_len / clip_stride count frames): 16-bit mono
 * input halves the host-to-device bytes of the end-to-end path, which is PCIe-bound. */
#define AA_PCM_F32 0
#define AA_PCM_I16 1
#define AA_PCM_U16 2
AA_API aa_status aa_ingest_device(const void *pcm_dev, int32_t format, int32_t channels, int64_t n_clips,
                                  int64_t clip_len, int64_t in_stride, int64_t out_stride, float *mono_dev,
                                  void *stream);
AA_API aa_status aa_analyze_host_pcm(aa_analyzer *h, const void *pcm_host, int32_t format, int32_t channels,
                                     int64_t n_clips, int64_t clip_len, int64_t clip_stride,
                                     const uint8_t *onset_in_host, const aa_outputs *out_host);

/* ------------------------------------------------------------------------- *
 * Note identification of the tuner stage (NEXT row f2): Note::from_freq
 * (src/analysis/theory.rs:195-209) applied to every stable pitch, as numbers
 * (name strings stay on the host: semis 0 = C, 1 = C#, ... 11 = B).
 * ------------------------------------------------------------------------- */
typedef struct aa_note { uint8_t semis, octave; uint16_t reserved; float cents; } aa_note;
typedef struct aa_note_record {           /* 136 bytes, one per frame, parallel to aa_stable_pitches */
    uint32_t n;
    uint32_t reserved;
    aa_note  note[AA_MAX_STABLE];
} aa_note_record;
/* base_freq: Tuner.base, 440 Hz by default (tuner.rs, theory.rs:196) */
AA_API aa_status aa_notes_from_stable_device(const aa_stable_pitches *stable_dev, int64_t n_frames,
                                             float base_freq, aa_note_record *notes_dev, void *stream);
AA_API aa_status aa_notes_from_stable_host(const aa_stable_pitches *stable_host, int64_t n_frames,
                                           float base_freq, aa_note_record *notes_host);

/* Offline onset events.  For every frame with AA_FLAG_ONSET_FIRED -- the offline reading of the
 * gating in OnsetDetector::detect_onsets (src/analysis/onset.rs:383-456: no metronome ticks to guard against,
 * calibration done) -- the OnsetEvent (src/audio_io/timing.rs:77-87) the reference would push on onset_tx, stamped
 * like MusicalTransport::stamp_onset (timing.rs:311-337) with zero input / output latency and calibration and the
 * clip start as time zero: velocity = clamp(max(flux, 5 max_excess) / 50, 0, 1) (onset.rs:388-390),
 * sample_position = frame * hop + n / 2 (the window centre), beat_position = sample_position * bpm / (60 sr).
 * events: [n_clips][max_events] in time order; counts[c] is the number of fired frames of clip c (it may exceed
 * max_events, in which case only the first max_events were written). */
typedef struct aa_onset_event {           /* 32 bytes */
    double   beat_position;
    int64_t  sample_position;
    int64_t  frame;
    float    velocity;
    uint32_t reserved;
} aa_onset_event;
AA_API aa_status aa_onset_events_device(const aa_config *cfg, const aa_frame_features *features_dev, int64_t n_clips,
                                        int64_t clip_len, float bpm, int32_t max_events, aa_onset_event *events_dev,
                                        int32_t *counts_dev, void *stream);
AA_API aa_status aa_onset_events_host(const aa_config *cfg, const aa_frame_features *features_host, int64_t n_clips,
                                      int64_t clip_len, float bpm, int32_t max_events, aa_onset_event *events_host,
                                      int32_t *counts_host);

/* Tuner::run's per-frame branch (src/analysis/tuner.rs:148-193) with Interval::new
 * (src/analysis/theory.rs:306-382) on the stable pitches: numeric form of TunerOutput.label / .cents
 * (the strings stay on the host).  system: 0 EqualTemperament, 1 JustIntonation, 2 Pythagorean
 * (tuner.rs:13-18); single_pitch_mode: TunerMode::SinglePitch (tuner.rs:21-25). */
typedef struct aa_tuner_record {          /* 16 bytes, one per frame, parallel to aa_stable_pitches */
    uint8_t  kind;      /* 0 nothing emitted (no stable pitch), 1 single note, 2 interval, 3 three or more notes */
    uint8_t  best;      /* kind 1: index of the displayed pitch (highest score; the last one on ties) */
    uint8_t  lo, hi;    /* kind 2: indices of the lower / higher pitch */
    uint32_t interval;  /* kind 2: IntType 0 Min2, 1 Maj2, 2 Min3, 3 Maj3, 4 Per4, 5 Aug4, 6 Per5, 7 Min6, 8 Maj6,
                           9 Min7, 10 Maj7, 11 Per8 (theory.rs:285-298) */
    float    accuracy;  /* kind 2: Interval::get_accuracy() */
    float    cents;     /* TunerOutput.cents: note cents (kind 1), interval accuracy (kind 2), 0 otherwise */
} aa_tuner_record;
AA_API aa_status aa_tuner_from_stable_device(const aa_stable_pitches *stable_dev, int64_t n_frames, float base_freq,
                                             int32_t system, int32_t single_pitch_mode,
                                             aa_tuner_record *out_dev, void *stream);
AA_API aa_status aa_tuner_from_stable_host(const aa_stable_pitches *stable_host, int64_t n_frames, float base_freq,
                                           int32_t system, int32_t single_pitch_mode, aa_tuner_record *out_host);

/* ------------------------------------------------------------------------- *
 * YIN-style lag search per frame (NEW, no reference code; the north_star asks
 * for "autocorrelation or YIN-style lag search").  On the raw (unwindowed) frame
 *   d(tau) = sum_{j<W} (x[j]-x[j+tau])^2, W = n - max_lag;  d'(tau) = d(tau)*tau / sum_{j<=tau} d(j)
 *   lag = first tau >= min_lag with d'(tau) < threshold, advanced to the local minimum; otherwise the
 *         arg-min of d' over [min_lag, max_lag].           frequency estimate = sample_rate / lag.
 * Frames as in aa_num_frames (n, hop); lag_out / cmnd_out: [n_clips*T] (cmnd_out may be NULL).
 * ------------------------------------------------------------------------- */
typedef struct aa_yin_config {
    int32_t n, hop;            /* frame length (<= 8192) and hop */
    int32_t min_lag, max_lag;  /* 1 <= min_lag <= max_lag < n */
    float   threshold;         /* 0.1 in the YIN paper */
} aa_yin_config;
AA_API aa_status aa_yin_device(const aa_yin_config *cfg, const float *clips_dev, int64_t n_clips,
                               int64_t clip_len, int64_t clip_stride, int32_t *lag_dev, float *cmnd_dev,
                               void *stream);
AA_API aa_status aa_yin_host(const aa_yin_config *cfg, const float *clips_host, int64_t n_clips,
                             int64_t clip_len, int64_t clip_stride, int32_t *lag_host, float *cmnd_host);

/* ------------------------------------------------------------------------- *
 * Streaming: the thread body of STFT::detect_pitches / OnsetDetector::
 * detect_onsets with the SlotPool -> private ring hand-off (stft.rs:240-266,
 * onset.rs:216-237, audio_io/mod.rs:32-79) replaced by rings in pinned host
 * memory that is mapped into the device address space: a push is a memcpy into
 * the sample ring and one kernel launch (the kernel reads the hops and writes the
 * records straight from / to host memory, then stores a completion word there);
 * a poll spins on that word.  The Rust worker thread stays; its body becomes
 * push / poll.
 * ------------------------------------------------------------------------- */
typedef struct aa_stream aa_stream;
typedef struct aa_stream_frame {
    int64_t           frame_index;     /* frames since stream creation                 */
    aa_frame_features features;
    aa_stable_pitches stable;
} aa_stream_frame;

AA_API aa_status aa_stream_create(const aa_config *cfg, aa_stream **out);
AA_API aa_status aa_stream_destroy(aa_stream *h);
/* Append `count` mono f32 samples (one 1024-sample slot in the reference, mod.rs:126-128;
 * any count <= ring capacity here) and run every frame that became complete. */
AA_API aa_status aa_stream_push(aa_stream *h, const float *samples, int32_t count);
AA_API aa_status aa_stream_set_noise_floor_db(aa_stream *h, float db);   /* stft.rs:322 */
AA_API aa_status aa_stream_signal_onset(aa_stream *h);                   /* onset_pending, stft.rs:387 */
/* Copy up to `max` completed frames (oldest first); *n_out receives the count. */
AA_API aa_status aa_stream_poll(aa_stream *h, aa_stream_frame *out, int32_t max, int32_t *n_out);
AA_API aa_status aa_stream_reset(aa_stream *h);
/* Measurement helper: `n_pushes` pushes of `count` samples each (taken from samples[n_pushes * count]) through
 * aa_stream_push + aa_stream_poll in a C loop -- what a native caller such as the Rust worker thread pays per push,
 * without a scripting language's call overhead.  latency_us[i] = wall-clock time of push i + the poll that follows it;
 * *frames_out = frames polled in total. */
AA_API aa_status aa_stream_probe_latency(aa_stream *h, const float *samples, int32_t count, int32_t n_pushes,
                                         double *latency_us, int64_t *frames_out);

/* ------------------------------------------------------------------------- *
 * Input conditioning chain (the step right before the analysis path).
 * Replaces the body of the reducer thread in AudioPipeline (src/audio_io/mod.rs:351-487:
 * 40 Hz high-pass and 14 kHz low-pass biquads, envelope follower, -60 dBFS gate with 20 ms hold)
 * and DynamicsTracker::process_slot (src/audio_io/dynamics.rs:194-360: slot RMS, noise-floor /
 * session percentiles, AGC gain smoothing with peak headroom, dynamics classification), slot by
 * slot (slot_len samples, 1024 in the reference: mod.rs:126-128).  Clips are conditioned in place;
 * only full slots exist in the reference (mod.rs:799-803), so samples after the last full slot of a
 * clip are left untouched.  aa_dynamics mirrors DynamicsOutput (dynamics.rs:78-91); its
 * noise_floor_db is what STFT / OnsetDetector read every frame (stft.rs:322, onset.rs:300).
 * ------------------------------------------------------------------------- */
#define AA_COND_AGC   1u   /* run DynamicsTracker (AGC + classification); otherwise filters + gate only */
#define AA_COND_CARRY 2u   /* keep filter / gate / tracker state in the handle between calls (streams) */
typedef struct aa_cond_config {
    float    sample_rate;
    int32_t  slot_len;     /* samples per slot, multiple of 4 (1024 in the reference) */
    uint32_t flags;        /* AA_COND_* */
} aa_cond_config;
typedef struct aa_dynamics {             /* 32 bytes, one per slot */
    int32_t  level;                      /* DynamicLevel: 0 Silence, 1 ppp, 2 pp, 3 p, 4 mp, 5 mf, 6 f, 7 ff, 8 fff */
    float    rms_db, gain_db, session_median_db, noise_floor_db;   /* dynamics.rs:83-90 */
    float    effective_gain;             /* linear gain applied to the slot (dynamics.rs:327-328) */
    uint32_t flags;                      /* 1 is_active, 2 is_broadband, 4 is_playing */
    uint32_t reserved;
} aa_dynamics;
typedef struct aa_conditioner aa_conditioner;
AA_API aa_status aa_conditioner_create(const aa_cond_config *cfg, aa_conditioner **out);
AA_API aa_status aa_conditioner_destroy(aa_conditioner *h);
AA_API aa_status aa_conditioner_reset(aa_conditioner *h);           /* forget carried state */
AA_API int64_t   aa_cond_num_slots(const aa_cond_config *cfg, int64_t clip_len);
/* clip c at clips + c*clip_stride (clip_stride % 4 == 0, base 16-byte aligned); dyn: [n_clips][num_slots] or NULL */
AA_API aa_status aa_condition_device(aa_conditioner *h, float *clips_dev, int64_t n_clips, int64_t clip_len,
                                     int64_t clip_stride, aa_dynamics *dyn_dev, void *stream);
AA_API aa_status aa_condition_host(aa_conditioner *h, float *clips_host, int64_t n_clips, int64_t clip_len,
                                   int64_t clip_stride, aa_dynamics *dyn_host);

/* ------------------------------------------------------------------------- *
 * Synthetic clips (bench/test support): clip c = K = 1..4 harmonic
 * tones (6 partials, 1/h amplitudes, f0 log-uniform in [55,1760] Hz) plus white
 * noise at -60 dBFS, all derived from splitmix64(seed + c).
 * ------------------------------------------------------------------------- */
AA_API aa_status aa_synth_clips_device(float *clips_dev, int64_t n_clips, int64_t clip_len,
                                       int64_t clip_stride, float sample_rate, uint64_t seed,
                                       void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AA_GPU_H */

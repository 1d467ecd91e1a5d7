/*
 * aa_oracle_cond.c -- CPU restatement of the reference's input conditioning chain
 * (SURVEY.md section 8f rank 1): the reducer thread's HPF / LPF biquads, envelope gate
 * (src/audio_io/mod.rs:351-472) and DynamicsTracker::process_slot (src/audio_io/dynamics.rs:194-360).
 *
 * TEST INFRASTRUCTURE ONLY (see aa_oracle.h).  PARITY UNPINNED: the reference has no test on this
 * code and cannot be built here; this is a line-by-line restatement (citations relative to
 * /root/reference), f32 everywhere the reference uses f32, operations in the reference's order,
 * compiled with -ffp-contract=off.  cos/sin/exp/powf/log10/sqrt go through libm as Rust's f32
 * methods do on Linux.
 */
#include "aa_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- parameters: mod.rs:340-418, dynamics.rs:156-192 ---------------------------------- */
static void calc_biquad(float freq, int is_lpf, float sample_rate, float out[5])   /* mod.rs:357-385 */
{
    const float PI_F = 3.14159265358979323846f;
    float w0 = 2.0f * PI_F * freq / sample_rate;
    float cos_w0 = cosf(w0);
    float sin_w0 = sinf(w0);
    float alpha = sin_w0 / (2.0f * 0.707f);
    float b0, b1, b2, a0, a1, a2;
    if (is_lpf) {
        b0 = (1.0f - cos_w0) / 2.0f;
        b1 = 1.0f - cos_w0;
        b2 = (1.0f - cos_w0) / 2.0f;
    } else {
        b0 = (1.0f + cos_w0) / 2.0f;
        b1 = -(1.0f + cos_w0);
        b2 = (1.0f + cos_w0) / 2.0f;
    }
    a0 = 1.0f + alpha;
    a1 = -2.0f * cos_w0;
    a2 = 1.0f - alpha;
    out[0] = b0 / a0; out[1] = b1 / a0; out[2] = b2 / a0; out[3] = a1 / a0; out[4] = a2 / a0;
}

void aao_cond_params_init(aao_cond_params *p, float sample_rate, int slot_len)
{
    memset(p, 0, sizeof(*p));
    calc_biquad(40.0f, 0, sample_rate, p->hp);                                /* mod.rs:387 */
    calc_biquad(14000.0f, 1, sample_rate, p->lp);                             /* mod.rs:388 */
    const float gate_threshold_db = -60.0f;                                   /* mod.rs:400 */
    p->gate_threshold_linear = powf(10.0f, gate_threshold_db / 20.0f);        /* mod.rs:401 */
    p->release_coeff = expf(-1.0f / (0.040f * sample_rate));                  /* mod.rs:408 */
    p->gate_hold_samples = (int32_t)(0.020f * sample_rate);                   /* mod.rs:413 */
    /* DynamicsTracker::new(sample_rate, slot_len, -18.0, 100.0, 240.0, ..)   mod.rs:347-355 */
    const float slot_rate = sample_rate / (float)slot_len;                    /* dynamics.rs:164 */
    p->target_db = -18.0f;
    p->max_boost_db = 100.0f;
    p->smooth_alpha = 1.0f - expf(-1.0f / (240.0f * slot_rate));              /* dynamics.rs:174 */
    p->silence_decay_alpha = 1.0f - expf(-1.0f / (10.0f * slot_rate));        /* dynamics.rs:175 */
    p->active_snr_db = 20.0f;                                                 /* dynamics.rs:188 */
    p->bootstrap_floor_db = -55.0f;                                           /* dynamics.rs:189 */
    p->slot_len = slot_len;
}

/* ---- state ------------------------------------------------------------------------------- */
#define LONG_LEN 256      /* dynamics.rs:168 */
#define PLAY_LEN 5000     /* dynamics.rs:172 */

struct aao_cond {
    aao_cond_params p;
    float hp_x1, hp_x2, hp_y1, hp_y2, lp_x1, lp_x2, lp_y1, lp_y2;   /* mod.rs:390-398 */
    float envelope;                                                 /* mod.rs:402 */
    uint32_t gate_hold_remaining;                                   /* mod.rs:414 */
    float long_history[LONG_LEN];
    int long_pos, long_filled;
    float play_history[PLAY_LEN];
    int play_pos, play_filled;
    float current_gain_linear;
    float sort_buf[PLAY_LEN];
};

aao_cond *aao_cond_create(const aao_cond_params *p)
{
    aao_cond *c = (aao_cond *)calloc(1, sizeof(aao_cond));
    if (!c) return NULL;
    c->p = *p;
    c->current_gain_linear = 1.0f;                                  /* dynamics.rs:183 */
    return c;
}

void aao_cond_destroy(aao_cond *c) { free(c); }

void aao_cond_reset(aao_cond *c)
{
    aao_cond_params p = c->p;
    memset(c, 0, sizeof(*c));
    c->p = p;
    c->current_gain_linear = 1.0f;
}

static inline float linear_to_db(float linear) { return 20.0f * log10f(fmaxf(linear, 1e-9f)); }  /* dynamics.rs:364-366 */
static inline float db_to_linear(float db) { return powf(10.0f, db / 20.0f); }                   /* dynamics.rs:369-371 */

static int cmp_f32(const void *a, const void *b)
{
    const float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* mod.rs:433-487: HPF, LPF, envelope follower, gate -- one slot, in place */
void aao_cond_filter_gate(aao_cond *c, float *slot, int len)
{
    const float *hp = c->p.hp, *lp = c->p.lp;
    const float thr = c->p.gate_threshold_linear, rc = c->p.release_coeff;
    for (int i = 0; i < len; ++i) {
        float x = slot[i];
        float hp_out = hp[0] * x + hp[1] * c->hp_x1 + hp[2] * c->hp_x2 - hp[3] * c->hp_y1 - hp[4] * c->hp_y2;
        c->hp_x2 = c->hp_x1; c->hp_x1 = x; c->hp_y2 = c->hp_y1; c->hp_y1 = hp_out;
        x = hp_out;
        float lp_out = lp[0] * x + lp[1] * c->lp_x1 + lp[2] * c->lp_x2 - lp[3] * c->lp_y1 - lp[4] * c->lp_y2;
        c->lp_x2 = c->lp_x1; c->lp_x1 = x; c->lp_y2 = c->lp_y1; c->lp_y1 = lp_out;
        x = lp_out;
        float abs_in = fabsf(x);
        if (abs_in > c->envelope) {                                            /* mod.rs:461-466 */
            c->envelope = abs_in;
            c->gate_hold_remaining = (uint32_t)c->p.gate_hold_samples;
        } else {
            c->envelope = rc * c->envelope + (1.0f - rc) * abs_in;
        }
        float gain;                                                            /* mod.rs:474-482 */
        if (c->envelope >= thr) {
            gain = 1.0f;
        } else if (c->gate_hold_remaining > 0) {
            c->gate_hold_remaining -= 1;
            gain = 1.0f;
        } else {
            float ratio = c->envelope / thr;
            gain = ratio * ratio * ratio * ratio;
        }
        slot[i] = x * gain;
    }
}

/* dynamics.rs:194-360: DynamicsTracker::process_slot.  apply != 0 scales the slot in place. */
void aao_cond_agc(aao_cond *c, float *slot, int len, aao_dynamics *out, int apply)
{
    /* 1. pre-gain slot RMS (:196-200) */
    float sum_sq = 0.0f;
    for (int i = 0; i < len; ++i) sum_sq += slot[i] * slot[i];
    const float rms_linear = sqrtf(sum_sq / (float)len);
    const float rms_db = linear_to_db(rms_linear);

    /* 2. long history -> noise floor (:203-221) */
    const int long_n = c->long_filled ? LONG_LEN : (c->long_pos > 1 ? c->long_pos : 1);
    float noise_floor_db;
    {
        memcpy(c->sort_buf, c->long_history, sizeof(float) * (size_t)long_n);
        qsort(c->sort_buf, (size_t)long_n, sizeof(float), cmp_f32);
        const int p10_idx = (int)((float)(long_n - 1) * 0.10f);
        noise_floor_db = linear_to_db(fmaxf(c->sort_buf[p10_idx], 1e-9f));
    }

    /* 3. active-frame gate (:224-229) */
    const float floor_db = long_n >= 32 ? noise_floor_db : c->p.bootstrap_floor_db;
    const int is_active = rms_db > floor_db + c->p.active_snr_db;

    /* 3b. broadband detection (:232-259) */
    int is_broadband = 0;
    if (is_active) {
        const float n = (float)len;
        const float mean_sq = rms_linear * rms_linear;
        float quad = 0.0f;
        for (int i = 0; i < len; ++i) {
            const float s2 = slot[i] * slot[i];
            quad += s2 * s2;
        }
        const float mean_quad = quad / n;
        const float kurtosis = mean_sq > 1e-18f ? mean_quad / (mean_sq * mean_sq) : 3.0f;
        is_broadband = kurtosis >= 2.75f && kurtosis <= 3.8f && rms_db < -45.0f;
    }
    const int is_playing = is_active && !is_broadband;                       /* :262 */

    if (!is_active || is_broadband) {                                         /* :266-272 */
        c->long_history[c->long_pos] = rms_linear;
        c->long_pos = (c->long_pos + 1) % LONG_LEN;
        if (c->long_pos == 0) c->long_filled = 1;
    }
    if (is_playing) {                                                         /* :275-282 */
        c->play_history[c->play_pos] = rms_linear;
        c->play_pos = (c->play_pos + 1) % PLAY_LEN;
        if (c->play_pos == 0) c->play_filled = 1;
    }

    /* 5. session statistics (:285-309) */
    const int play_n = c->play_filled ? PLAY_LEN : c->play_pos;
    float raw_gain_db, session_median_db;
    if (play_n > 0) {
        memcpy(c->sort_buf, c->play_history, sizeof(float) * (size_t)play_n);
        qsort(c->sort_buf, (size_t)play_n, sizeof(float), cmp_f32);
        const int p50_idx = (play_n - 1) / 2;
        const int p95_idx = (int)((float)(play_n - 1) * 0.95f);
        const float median_db = linear_to_db(fmaxf(c->sort_buf[p50_idx], 1e-9f));
        const float p95_db = linear_to_db(fmaxf(c->sort_buf[p95_idx], 1e-9f));
        float g = c->p.target_db - p95_db;
        if (g < 0.0f) g = 0.0f;
        if (g > c->p.max_boost_db) g = c->p.max_boost_db;
        raw_gain_db = g;
        session_median_db = median_db;
    } else {
        raw_gain_db = 0.0f;
        session_median_db = rms_db;
    }

    /* 6. smooth gain (:312-318) */
    if (is_playing) {
        const float target_linear = db_to_linear(raw_gain_db);
        c->current_gain_linear += c->p.smooth_alpha * (target_linear - c->current_gain_linear);
    } else {
        c->current_gain_linear += c->p.silence_decay_alpha * (1.0f - c->current_gain_linear);
    }

    /* 7. apply gain + peak-headroom clamp (:321-334) */
    float peak = 0.0f;
    for (int i = 0; i < len; ++i) peak = fmaxf(peak, fabsf(slot[i]));
    peak = fmaxf(peak, 1e-9f);
    const float headroom_limit = 0.97f / peak;
    const float effective_gain = fminf(c->current_gain_linear, headroom_limit);
    if (apply)
        for (int i = 0; i < len; ++i) slot[i] *= effective_gain;
    const float applied_gain_db = linear_to_db(effective_gain);

    /* 8. classification (:337-352): 0 Silence, 1 ppp ... 8 fff */
    int level;
    if (!is_playing) {
        level = 0;
    } else {
        const float r = rms_db - session_median_db;
        level = r < -15.0f ? 1 : r < -9.0f ? 2 : r < -4.5f ? 3 : r < -1.5f ? 4 : r < 1.5f ? 5 : r < 4.5f ? 6 : r < 9.0f ? 7 : 8;
    }
    if (out) {
        out->level = level;
        out->rms_db = rms_db;
        out->gain_db = applied_gain_db;
        out->session_median_db = session_median_db;
        out->noise_floor_db = noise_floor_db;
        out->effective_gain = effective_gain;
        out->flags = (uint32_t)((is_active ? 1 : 0) | (is_broadband ? 2 : 0) | (is_playing ? 4 : 0));
        out->reserved = 0;
    }
}

/* A whole clip, slot by slot (only full slots exist in the reference: mod.rs:799-803).  Returns the
 * number of slots processed; samples past the last full slot are left untouched.  agc: 0 = filters
 * and gate only, 1 = full chain.  dyn (optional): one record per slot. */
int64_t aao_cond_clip(const aao_cond_params *p, float *samples, int64_t len, aao_dynamics *dyn, int agc)
{
    aao_cond *c = aao_cond_create(p);
    if (!c) return -1;
    const int64_t n_slots = len / p->slot_len;
    for (int64_t s = 0; s < n_slots; ++s) {
        float *slot = samples + s * p->slot_len;
        aao_cond_filter_gate(c, slot, p->slot_len);
        if (agc) aao_cond_agc(c, slot, p->slot_len, dyn ? dyn + s : NULL, 1);
    }
    aao_cond_destroy(c);
    return n_slots;
}

/* ------------------------------------------------------------------------------------------
 * Tuner post-stage: Interval::new (src/analysis/theory.rs:306-382) and the
 * mode selection of Tuner::run (src/analysis/tuner.rs:152-193).  PINNED by the reference's own
 * known-answer tests (theory.rs:545-604), which tests/test_oracle_cond.py replays.
 * system: 0 EqualTemperament (default), 1 JustIntonation, 2 Pythagorean (tuner.rs:13-18).
 * Returns the IntType index 0 Min2 .. 10 Maj7, 11 Per8 (theory.rs:285-298). */
int aao_interval(float f_lo, float f_hi, int system, float *accuracy)
{
    static const float JUST[13] = {1.0f, 16.0f / 15.0f, 9.0f / 8.0f, 6.0f / 5.0f, 5.0f / 4.0f, 4.0f / 3.0f, 45.0f / 32.0f,
                                   3.0f / 2.0f, 8.0f / 5.0f, 5.0f / 3.0f, 9.0f / 5.0f, 15.0f / 8.0f, 2.0f};
    static const float PYTH[13] = {1.0f, 256.0f / 243.0f, 9.0f / 8.0f, 32.0f / 27.0f, 81.0f / 64.0f, 4.0f / 3.0f,
                                   729.0f / 512.0f, 3.0f / 2.0f, 128.0f / 81.0f, 27.0f / 16.0f, 32.0f / 9.0f,
                                   243.0f / 128.0f, 2.0f};
    static const float ET[13] = {1.0f, 1.0595f, 1.1225f, 1.1892f, 1.2599f, 1.3348f, 1.4142f, 1.4983f, 1.5874f, 1.6818f,
                                 1.7818f, 1.8877f, 2.0f};
    if (f_lo == 0.0f) { *accuracy = 0.0f; return 11; }                /* :307-312 */
    float ratio = f_hi / f_lo;                                        /* :313 */
    while (ratio > 2.0f) ratio /= 2.0f;                               /* :314-316 */
    const float *r = system == 1 ? JUST : system == 2 ? PYTH : ET;
    int idx = 0;
    float best = fabsf(ratio - r[0]);
    for (int i = 1; i < 13; ++i) {                                    /* :356-364 min_by: first minimum */
        const float d = fabsf(ratio - r[i]);
        if (d < best) { best = d; idx = i; }
    }
    *accuracy = -logf(r[idx] / ratio) * 1732.5f;                      /* :381 */
    return idx == 0 ? 11 : idx - 1;                                   /* :365-379 */
}

/* Tuner::run's branch for one frame of stable pitches (tuner.rs:148-193).  kind: 0 nothing emitted
 * (empty list), 1 single note (one pitch, or SinglePitch mode): best = index of the LAST pitch with the
 * maximum score (Iterator::max_by), 2 interval between the two pitches sorted by frequency, 3 three or
 * more notes (names only). */
void aao_tuner_frame(const float *pairs, int n, int system, int single_pitch_mode, int *kind, int *best, int *lo,
                     int *hi, int *interval, float *accuracy)
{
    *kind = 0; *best = 0; *lo = 0; *hi = 0; *interval = 0; *accuracy = 0.0f;
    if (n <= 0) return;
    if (n == 1 || single_pitch_mode) {
        int b = 0;
        for (int i = 1; i < n; ++i)
            if (!(pairs[2 * i + 1] < pairs[2 * b + 1])) b = i;     /* total_cmp on finite scores; ties -> last */
        *kind = 1; *best = b;
    } else if (n == 2) {
        const int l = pairs[2] < pairs[0] ? 1 : 0;                   /* sort_by total_cmp, stable */
        *kind = 2; *lo = l; *hi = 1 - l;
        *interval = aao_interval(pairs[2 * l], pairs[2 * (1 - l)], system, accuracy);
    } else {
        *kind = 3;
    }
}

/* ------------------------------------------------------------------------------------------
 * Input callback: device sample format -> mono f32 slot (src/audio_io/mod.rs:765-792).
 *   channels_to_use = min(channels, 2);  mixed = fold(0.0f32, acc + s.to_sample::<f32>()) / channels_to_use
 * `to_sample` is cpal 0.16.0's re-export of dasp_sample 0.11.0 (Cargo.lock:1768-1770, 2003-2005; crate not
 * vendored): i16 -> f32 is `s as f32 / 32_768.0`, u16 -> f32 goes through i16 (`s - 32768`), f32 is the identity.
 * format: 0 f32, 1 i16, 2 u16.  Every operation is exact for the integer formats.
 * ------------------------------------------------------------------------------------------ */
void aao_ingest(const void *pcm, int format, int channels, int64_t n_frames, float *out)
{
    const int use = channels < 2 ? channels : 2;
    for (int64_t f = 0; f < n_frames; ++f) {
        float acc = 0.0f;
        for (int c = 0; c < use; ++c) {
            const int64_t i = f * channels + c;
            float v;
            if (format == 1) v = (float)((const int16_t *)pcm)[i] / 32768.0f;
            else if (format == 2) v = (float)(int16_t)((int32_t)((const uint16_t *)pcm)[i] - 32768) / 32768.0f;
            else v = ((const float *)pcm)[i];
            acc = acc + v;
        }
        out[f] = acc / (float)use;
    }
}

/* ------------------------------------------------------------------------------------------
 * MusicalTransport::stamp_onset, audio_io/timing.rs:311-337, with the transport's atomics passed in:
 *   beats_per_sample = bpm / (60 sr)                                              :313-315 (f64)
 *   beat_position    = current_beats - (input_lat + output_lat) * bps + sample_offset * bps - calibration * bps
 *                      (in this order, :326-330)
 *   output_samples   = output_frames - input_lat - output_lat + sample_offset - calibration        :335-336
 * sample_offset is what the detector passes: -(available_samples - window_size / 2), onset.rs:386-387.
 * Pinned by a value the real crate logged (tests/golden/ref_log_onsets.json, from the reference's output.log).
 * ------------------------------------------------------------------------------------------ */
void aao_stamp_onset(double current_beats, int64_t output_frames, float bpm, float sample_rate, int64_t input_lat,
                     int64_t output_lat, int64_t calibration, int64_t sample_offset, double *beat_position,
                     int64_t *output_samples)
{
    const double sr = (double)sample_rate;
    const double b = (double)bpm;
    const double beats_per_sample = b / (60.0 * sr);
    const double latency_beats = (double)(input_lat + output_lat) * beats_per_sample;
    const double offset_beats = (double)sample_offset * beats_per_sample;
    const double calibration_beats = (double)calibration * beats_per_sample;
    *beat_position = current_beats - latency_beats + offset_beats - calibration_beats;
    *output_samples = output_frames - input_lat - output_lat + sample_offset - calibration;
}

/* ------------------------------------------------------------------------------------------
 * Offline onset events: what OnsetDetector pushes on onset_tx for every frame whose
 * gating passed (onset.rs:383-456 with no metronome ticks and calibration done = AAO_FLAG_ONSET_FIRED),
 * stamped like MusicalTransport::stamp_onset (timing.rs:311-337) with zero latencies / calibration and the
 * clip start as time zero:
 *   velocity        = clamp(max(flux, max_excess * 5) / 50, 0, 1)            onset.rs:388-390 (f32)
 *   sample_position = frame * hop + n / 2   (the window centre, onset.rs:386-387)
 *   beat_position   = sample_position * bpm / (60 * sr)                       timing.rs:313-326 (f64)
 * Returns the number of events of the clip (all of them are counted, at most max_events are written).
 * ------------------------------------------------------------------------------------------ */
int64_t aao_onset_events(const aao_features *feat, int64_t T, int n, int hop, float sample_rate, float bpm,
                         int64_t max_events, aao_onset_event *out)
{
    const double beats_per_sample = (double)bpm / (60.0 * (double)sample_rate);
    int64_t cnt = 0;
    for (int64_t f = 0; f < T; ++f) {
        if (!(feat[f].flags & AAO_FLAG_ONSET_FIRED)) continue;
        if (cnt < max_events) {
            float v = fmaxf(feat[f].flux, feat[f].max_excess * 5.0f) / 50.0f;
            if (v < 0.0f) v = 0.0f;
            if (v > 1.0f) v = 1.0f;
            aao_onset_event e;
            e.sample_position = f * (int64_t)hop + n / 2;
            e.beat_position = (double)e.sample_position * beats_per_sample;
            e.frame = f;
            e.velocity = v;
            e.reserved = 0;
            out[cnt] = e;
        }
        ++cnt;
    }
    return cnt;
}

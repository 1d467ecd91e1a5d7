/*
 * aa_oracle.c -- CPU restatement of the audio-analyzer-rs frame-analysis path.
 *
 * TEST INFRASTRUCTURE ONLY / PARITY UNPINNED -- see aa_oracle.h.
 * Compile with:  gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#include "aa_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* a1: periodic Hann window, stft.rs:641-648 (identical copy at onset.rs:549) */
/* ------------------------------------------------------------------------- */
void aao_hann_window(int n, float *w)
{
    /* std::f32::consts::PI is the f32 nearest to pi */
    const float pi_f32 = 3.14159274101257324219f;
    for (int i = 0; i < n; ++i) {
        float x = (float)i / (float)n;                 /* :644 */
        w[i] = 0.5f - 0.5f * cosf(2.0f * pi_f32 * x);  /* :645 */
    }
}

/* ------------------------------------------------------------------------- */
/* a3: FftProcessor::new / process_forward, dsp/fft.rs:14-35, 66-71.          */
/*                                                                            */
/* The arithmetic is in realfft 3.5.0 (RealToComplexEven) over rustfft 6.4.1, */
/* neither present under /root/reference.  Published algorithm restated:      */
/*   1. view the n real inputs as n/2 complex values z[m] = x[2m] + i x[2m+1] */
/*   2. unnormalised forward complex FFT of length n/2 (rustfft; here a       */
/*      Stockham radix-4(+2) with twiddles computed in f64 and rounded to f32,*/
/*      which is how rustfft builds its twiddle tables)                       */
/*   3. realfft's split post-pass with twiddles 0.5*exp(-2*pi*i*k/n), k=1..   */
/*      n/4-1, the DC/Nyquist pair from z[0] alone and the centre element     */
/*      conjugated.                                                           */
/* rustfft picks SIMD kernels at run time, so the reference's own low bits    */
/* are host dependent; parity on spectra is therefore a tolerance (1e-4 of    */
/* the frame maximum), never bit equality.                                    */
/* ------------------------------------------------------------------------- */
struct aao_fft {
    int n, n2;
    float *tw_re, *tw_im;   /* exp(-2 pi i k / n2), k < n2 (complex FFT twiddles) */
    float *pt_re, *pt_im;   /* 0.5*exp(-2 pi i k / n), k = 1..n/4-1 (post-pass)   */
    float *a_re, *a_im, *b_re, *b_im;   /* ping-pong work buffers */
};

static const double AAO_PI = 3.14159265358979323846;

aao_fft *aao_fft_create(int n)
{
    if (n < 4 || (n & (n - 1)) != 0) return NULL;
    aao_fft *p = (aao_fft *)calloc(1, sizeof(*p));
    p->n = n;
    p->n2 = n / 2;
    int n2 = p->n2;
    p->tw_re = (float *)malloc(sizeof(float) * n2);
    p->tw_im = (float *)malloc(sizeof(float) * n2);
    for (int k = 0; k < n2; ++k) {
        double a = -2.0 * AAO_PI * (double)k / (double)n2;
        p->tw_re[k] = (float)cos(a);
        p->tw_im[k] = (float)sin(a);
    }
    int q = n / 4;
    p->pt_re = (float *)malloc(sizeof(float) * (q > 0 ? q : 1));
    p->pt_im = (float *)malloc(sizeof(float) * (q > 0 ? q : 1));
    for (int k = 0; k < q; ++k) {
        double a = -2.0 * AAO_PI * (double)k / (double)n;
        p->pt_re[k] = (float)cos(a) * 0.5f;
        p->pt_im[k] = (float)sin(a) * 0.5f;
    }
    p->a_re = (float *)malloc(sizeof(float) * n2);
    p->a_im = (float *)malloc(sizeof(float) * n2);
    p->b_re = (float *)malloc(sizeof(float) * n2);
    p->b_im = (float *)malloc(sizeof(float) * n2);
    return p;
}

void aao_fft_destroy(aao_fft *p)
{
    if (!p) return;
    free(p->tw_re); free(p->tw_im); free(p->pt_re); free(p->pt_im);
    free(p->a_re); free(p->a_im); free(p->b_re); free(p->b_im);
    free(p);
}

/* One Stockham pass of radix 4 (or 2) from (xr,xi) to (yr,yi); ns = product of
 * the radices already applied. */
static void stockham4_f32(const aao_fft *p, int ns, const float *xr, const float *xi,
                          float *yr, float *yi)
{
    const int n2 = p->n2, q = n2 / 4, tstep = n2 / (4 * ns);
    for (int j = 0; j < q; ++j) {
        int k = j % ns;
        float v0r = xr[j], v0i = xi[j];
        float v1r = xr[j + q], v1i = xi[j + q];
        float v2r = xr[j + 2 * q], v2i = xi[j + 2 * q];
        float v3r = xr[j + 3 * q], v3i = xi[j + 3 * q];
        if (k) {
            float w1r = p->tw_re[k * tstep], w1i = p->tw_im[k * tstep];
            float w2r = p->tw_re[2 * k * tstep], w2i = p->tw_im[2 * k * tstep];
            float w3r = p->tw_re[3 * k * tstep], w3i = p->tw_im[3 * k * tstep];
            float t;
            t = v1r * w1r - v1i * w1i; v1i = v1r * w1i + v1i * w1r; v1r = t;
            t = v2r * w2r - v2i * w2i; v2i = v2r * w2i + v2i * w2r; v2r = t;
            t = v3r * w3r - v3i * w3i; v3i = v3r * w3i + v3i * w3r; v3r = t;
        }
        float a0r = v0r + v2r, a0i = v0i + v2i;
        float a1r = v0r - v2r, a1i = v0i - v2i;
        float a2r = v1r + v3r, a2i = v1i + v3i;
        float a3r = v1i - v3i, a3i = v3r - v1r;      /* (v1 - v3) * (-i) */
        int j0 = (j / ns) * ns * 4 + k;
        yr[j0] = a0r + a2r;          yi[j0] = a0i + a2i;
        yr[j0 + ns] = a1r + a3r;     yi[j0 + ns] = a1i + a3i;
        yr[j0 + 2 * ns] = a0r - a2r; yi[j0 + 2 * ns] = a0i - a2i;
        yr[j0 + 3 * ns] = a1r - a3r; yi[j0 + 3 * ns] = a1i - a3i;
    }
}

static void stockham2_f32(const aao_fft *p, int ns, const float *xr, const float *xi,
                          float *yr, float *yi)
{
    const int n2 = p->n2, h = n2 / 2, tstep = n2 / (2 * ns);
    for (int j = 0; j < h; ++j) {
        int k = j % ns;
        float v0r = xr[j], v0i = xi[j];
        float v1r = xr[j + h], v1i = xi[j + h];
        if (k) {
            float wr = p->tw_re[k * tstep], wi = p->tw_im[k * tstep];
            float t = v1r * wr - v1i * wi; v1i = v1r * wi + v1i * wr; v1r = t;
        }
        int j0 = (j / ns) * ns * 2 + k;
        yr[j0] = v0r + v1r;      yi[j0] = v0i + v1i;
        yr[j0 + ns] = v0r - v1r; yi[j0 + ns] = v0i - v1i;
    }
}

void aao_fft_forward(aao_fft *p, float *time, float *spec)
{
    const int n = p->n, n2 = p->n2;
    float *xr = p->a_re, *xi = p->a_im, *yr = p->b_re, *yi = p->b_im;
    for (int m = 0; m < n2; ++m) { xr[m] = time[2 * m]; xi[m] = time[2 * m + 1]; }
    /* realfft uses the input as scratch; mirror "clobbered" semantics (fft.rs:33) */
    memset(time, 0, sizeof(float) * (size_t)n);

    int ns = 1;
    while (ns < n2) {
        if (n2 / ns >= 4 && ((n2 / ns) & 0x55555555) != 0) {
            /* remaining length is a power of 4 */
            stockham4_f32(p, ns, xr, xi, yr, yi);
            ns *= 4;
        } else {
            stockham2_f32(p, ns, xr, xi, yr, yi);
            ns *= 2;
        }
        float *t;
        t = xr; xr = yr; yr = t;
        t = xi; xi = yi; yi = t;
    }
    /* xr/xi now hold Z[0..n2) in natural order */

    /* realfft RealToComplexEven post-pass */
    {
        float fr = xr[0], fi = xi[0];
        spec[0] = fr + fi;            spec[1] = 0.0f;          /* DC      */
        spec[2 * n2] = fr - fi;       spec[2 * n2 + 1] = 0.0f; /* Nyquist */
    }
    for (int k = 1; k < n / 4; ++k) {
        float ar = xr[k], ai = xi[k];              /* out     = Z[k]      */
        float br = xr[n2 - k], bi = xi[n2 - k];    /* out_rev = Z[n2 - k] */
        float tr = p->pt_re[k], ti = p->pt_im[k];
        float sum_re = ar + br, sum_im = ai + bi;
        float diff_re = ar - br, diff_im = ai - bi;
        float twiddled_re_sum_im = sum_im * tr;
        float twiddled_im_sum_im = sum_im * ti;
        float twiddled_re_diff_re = diff_re * tr;
        float twiddled_im_diff_re = diff_re * ti;
        float half_sum_re = 0.5f * sum_re;
        float half_diff_im = 0.5f * diff_im;
        float out_tw_re = twiddled_re_sum_im + twiddled_im_diff_re;
        float out_tw_im = twiddled_im_sum_im - twiddled_re_diff_re;
        spec[2 * k] = half_sum_re + out_tw_re;
        spec[2 * k + 1] = half_diff_im + out_tw_im;
        spec[2 * (n2 - k)] = half_sum_re - out_tw_re;
        spec[2 * (n2 - k) + 1] = out_tw_im - half_diff_im;
    }
    if (n >= 4) {                                  /* centre element: conj */
        int c = n / 4;
        spec[2 * c] = xr[c];
        spec[2 * c + 1] = -xi[c];
    }
}

/* float64 truth: recursive radix-2 DIT on the full real sequence (as complex). */
static void fft_f64_rec(double *re, double *im, int n, int stride, double *or_, double *oi)
{
    if (n == 1) { or_[0] = re[0]; oi[0] = im[0]; return; }
    int h = n / 2;
    fft_f64_rec(re, im, h, stride * 2, or_, oi);
    fft_f64_rec(re + stride, im + stride, h, stride * 2, or_ + h, oi + h);
    for (int k = 0; k < h; ++k) {
        double a = -2.0 * AAO_PI * (double)k / (double)n;
        double wr = cos(a), wi = sin(a);
        double tr = or_[k + h] * wr - oi[k + h] * wi;
        double ti = or_[k + h] * wi + oi[k + h] * wr;
        double er = or_[k], ei = oi[k];
        or_[k] = er + tr;      oi[k] = ei + ti;
        or_[k + h] = er - tr;  oi[k + h] = ei - ti;
    }
}

void aao_rdft_f64(const double *x, int n, double *spec)
{
    double *re = (double *)malloc(sizeof(double) * n * 4);
    double *im = re + n, *or_ = re + 2 * n, *oi = re + 3 * n;
    for (int i = 0; i < n; ++i) { re[i] = x[i]; im[i] = 0.0; }
    fft_f64_rec(re, im, n, 1, or_, oi);
    for (int k = 0; k <= n / 2; ++k) { spec[2 * k] = or_[k]; spec[2 * k + 1] = oi[k]; }
    free(re);
}

/* ------------------------------------------------------------------------- */
/* a4: magnitudes, stft.rs:314-318 (Complex::norm == hypot), onset.rs:271-272 */
/* ------------------------------------------------------------------------- */
void aao_magnitudes(const float *spec, int half, float *mags)
{
    for (int k = 0; k < half; ++k) mags[k] = hypotf(spec[2 * k], spec[2 * k + 1]);
}

/* a5: stft.rs:322-324 == onset.rs:300-301 */
float aao_global_floor(float noise_floor_db, int half)
{
    return powf(10.0f, noise_floor_db / 20.0f) * (float)half / 2.0f;
}

/* ------------------------------------------------------------------------- */
/* a6: adaptive per-bin floor, stft.rs:209-224 (state, constants), 326-367     */
/* ------------------------------------------------------------------------- */
struct aao_pitch_floor {
    int half;
    int initialized;          /* floor_initialized          :212 */
    float *nf;                /* noise_floor_per_bin        :209 */
    float *prev;              /* prev_mag_for_vol           :210 */
    float *vol;               /* bin_volatility             :211 */
};

aao_pitch_floor *aao_pitch_floor_create(int half)
{
    aao_pitch_floor *s = (aao_pitch_floor *)calloc(1, sizeof(*s));
    s->half = half;
    s->nf = (float *)calloc((size_t)half, sizeof(float));
    s->prev = (float *)calloc((size_t)half, sizeof(float));
    s->vol = (float *)calloc((size_t)half, sizeof(float));
    return s;
}

void aao_pitch_floor_destroy(aao_pitch_floor *s)
{
    if (!s) return;
    free(s->nf); free(s->prev); free(s->vol); free(s);
}

void aao_pitch_floor_reset(aao_pitch_floor *s)
{
    s->initialized = 0;
    memset(s->nf, 0, sizeof(float) * (size_t)s->half);
    memset(s->prev, 0, sizeof(float) * (size_t)s->half);
    memset(s->vol, 0, sizeof(float) * (size_t)s->half);
}

static inline float clampf(float x, float lo, float hi)
{
    /* f32::clamp: NaN stays NaN; not reachable here */
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}

void aao_pitch_floor_update(aao_pitch_floor *s, const float *mags, float global_floor,
                            float *effective_floor)
{
    const float FLOOR_BASE_ALPHA = 0.04f;   /* :219 */
    const float FLOOR_FAST_ALPHA = 0.35f;   /* :220 */
    const float FLOOR_RELEASE = 0.02f;      /* :221 */
    const float VOL_MEMORY = 0.75f;         /* :222 */
    const float NOTE_RATIO = 1.5f;          /* :223 */
    const float NOTE_VOL_MAX = 0.15f;       /* :224 */
    const int half = s->half;

    if (!s->initialized) {                                        /* :326-331 */
        for (int k = 0; k < half; ++k) {
            s->nf[k] = fmaxf(mags[k], global_floor * 5.0f);
            s->prev[k] = mags[k];
        }
        s->initialized = 1;
    } else {                                                      /* :338-363 */
        for (int k = 0; k < half; ++k) {
            float mag = mags[k];
            float floor_ = s->nf[k];
            float delta = fabsf(mag - s->prev[k]);
            s->vol[k] = s->vol[k] * VOL_MEMORY + delta * (1.0f - VOL_MEMORY);
            s->prev[k] = mag;

            float above_ratio = mag / fmaxf(floor_, 0.01f);
            float vol_norm = clampf(s->vol[k] / fmaxf(mag, 0.05f), 0.0f, 1.0f);
            int is_sustained_note = above_ratio > NOTE_RATIO && vol_norm < NOTE_VOL_MAX;

            if (!is_sustained_note) {
                float alpha;
                if (mag > floor_)
                    alpha = FLOOR_BASE_ALPHA + (FLOOR_FAST_ALPHA - FLOOR_BASE_ALPHA) * vol_norm;
                else
                    alpha = FLOOR_RELEASE;
                s->nf[k] += alpha * (mag - floor_);
            }
        }
    }
    for (int k = 0; k < half; ++k)                                /* :365-367 */
        effective_floor[k] = fminf(s->nf[k], global_floor * 2.5f);
}

/* parity-test taps: the raw recurrent state after the update of the current frame */
const float *aao_pitch_floor_nf(const aao_pitch_floor *s) { return s->nf; }
const float *aao_pitch_floor_vol(const aao_pitch_floor *s) { return s->vol; }

/* ------------------------------------------------------------------------- */
/* a7: STFT::extract_pitches, stft.rs:443-620                                  */
/* ------------------------------------------------------------------------- */
static inline size_t f32_as_usize(float x)
{
    /* Rust `as usize`: saturating, NaN -> 0 */
    if (!(x > 0.0f)) return 0;
    if (x >= 1.8446744e19f) return (size_t)-1;
    return (size_t)x;
}

static int g_margin_src_dummy;
static inline void margin_update2(float *m, int *src, float dist, int tag)
{
    if (dist < 0.0f) dist = -dist;
    if (dist < *m) { *m = dist; *src = tag; }
}
#define margin_update(m, dist, tag) margin_update2((m), &margin_src, (dist), (tag))

typedef struct { int bin; float score; } aao_cand;

/* Magnitude-perturbation margin (aao_pitch_diag.cand_eps): for a decision lhs > rhs taken on quantities
 * derived from the frame's magnitudes, the smallest uniform |dm| (every magnitude, and every floor value
 * that is not clamped to its constant, moved by at most dm in the worst direction) that could flip it is
 * |lhs - rhs| / (L1 norm of the gradient of lhs - rhs with respect to those values), first order. */
static inline void cand_update(float *m, int *src, float gap, float sens, int tag)
{
    if (!(sens > 0.0f)) return;
    if (gap < 0.0f) gap = -gap;
    float e = gap / sens;
    if (e < *m) { *m = e; *src = tag; }
}

static int extract_pitches_impl(const float *magnitudes, int half_size, float bin_width,
                                float min_freq, float max_freq, const float *noise_floor,
                                const float *floor_sens, float *out_pairs, uint8_t *peak_mask,
                                aao_pitch_diag *diag);

int aao_extract_pitches(const float *magnitudes, int half_size, float bin_width,
                        float min_freq, float max_freq, const float *noise_floor,
                        float *out_pairs, uint8_t *peak_mask, aao_pitch_diag *diag)
{
    return extract_pitches_impl(magnitudes, half_size, bin_width, min_freq, max_freq, noise_floor, NULL,
                                out_pairs, peak_mask, diag);
}

/* floor_sens[k] (optional): d noise_floor[k] / d magnitude bound -- 0 where the effective floor sits at its
 * clamp 2.5 * global_floor (stft.rs:366), 1 where it is the adaptive floor (an EMA of past magnitudes);
 * NULL = 1 everywhere. */
static int extract_pitches_impl(const float *magnitudes, int half_size, float bin_width,
                                float min_freq, float max_freq, const float *noise_floor,
                                const float *floor_sens, float *out_pairs, uint8_t *peak_mask,
                                aao_pitch_diag *diag)
{
    enum { MAX_HARMONICS = 14, MAX_NOTES = 8 };                     /* :451-452 */
    float margin = 1.0f;
    int margin_src = 0;
    float ceps = 1e30f;           /* cand_eps, see cand_update */
    int ceps_src = 0;
    (void)g_margin_src_dummy;
    if (diag) {
        memset(diag, 0, sizeof(*diag));
        diag->min_margin = 1.0f;
        diag->cand_eps = 1e30f;
        for (int i = 0; i < MAX_NOTES; ++i) diag->out_bins[i] = -1;
    }
#define FSENS(k) (floor_sens ? floor_sens[(k)] : 1.0f)
#define CEPS(gap, sens, tag) cand_update(&ceps, &ceps_src, (gap), (sens), (tag))
    if (peak_mask) memset(peak_mask, 0, (size_t)half_size);

    size_t min_bin = f32_as_usize(ceilf(min_freq / bin_width));      /* :454 */
    if (min_bin < 1) min_bin = 1;
    size_t max_bin = f32_as_usize(floorf(max_freq / bin_width));     /* :455 */
    size_t hs2 = half_size >= 2 ? (size_t)half_size - 2 : 0;
    if (max_bin > hs2) max_bin = hs2;
    if (min_bin >= max_bin) return 0;                                /* :457-459 */

    uint8_t *is_peak = (uint8_t *)calloc((size_t)half_size, 1);      /* :461 */
    int *peak_bins = (int *)malloc(sizeof(int) * (size_t)half_size); /* :462 */
    int n_peaks = 0;
    for (size_t k = min_bin + 1; k < max_bin; ++k) {                 /* :463-469 */
        float m = magnitudes[k];
        if (m > noise_floor[k] && m >= magnitudes[k - 1] && m >= magnitudes[k + 1]) {
            is_peak[k] = 1;
            peak_bins[n_peaks++] = (int)k;
        }
    }
    if (peak_mask) memcpy(peak_mask, is_peak, (size_t)half_size);
    if (diag) diag->n_peaks = n_peaks;
    if (n_peaks == 0) { free(is_peak); free(peak_bins); return 0; }  /* :471-473 */

    float *scores = (float *)calloc((size_t)half_size, sizeof(float));    /* :475 */
    float *frac_bins = (float *)calloc((size_t)half_size, sizeof(float)); /* :476 */
    float *sens_frac = (float *)calloc((size_t)half_size, sizeof(float));  /* |d frac_bin / d mags|_1 */
    float *sens_score = (float *)calloc((size_t)half_size, sizeof(float)); /* |d score / d mags|_1    */
    int n_scored = 0;
    for (int pi = 0; pi < n_peaks; ++pi) {                           /* :477 */
        int k = peak_bins[pi];
        float fund_mag = magnitudes[k];
        CEPS(fund_mag - noise_floor[k] * 5.0f, 1.0f + 5.0f * FSENS(k), 104);
        if (fund_mag < noise_floor[k] * 5.0f) {                      /* :479-482 */
            scores[k] = 0.0f;
            continue;
        }
        ++n_scored;
        float frac_bin;
        int frac_is_exact = 0;
        float sfrac = 0.0f;
        if (k >= 1 && k + 1 < half_size) {                           /* :484-494 */
            float y_l = logf(magnitudes[k - 1]);
            float y_c = logf(magnitudes[k]);
            float y_r = logf(magnitudes[k + 1]);
            float denom = y_l - 2.0f * y_c + y_r;
            float delta;
            if (fabsf(denom) < 1e-30f) delta = 0.0f;
            else {
                float raw = 0.5f * (y_l - y_r) / denom;
                delta = clampf(raw, -1.0f, 1.0f);
                /* a robustly clamped delta is an exact +-1: frac_bin is then an integer in
                 * every implementation and the comb-window boundaries carry no rounding risk */
                frac_is_exact = fabsf(raw) > 1.001f;
                if (!frac_is_exact) margin_update(&margin, fabsf(raw) - 1.0f, 11);
                {   /* d raw / d (m_l, m_c, m_r) through y = ln m */
                    float a = y_l - y_r, d2 = denom * denom;
                    float gl = fabsf(0.5f * (1.0f / denom - a / d2)) / magnitudes[k - 1];
                    float gc = fabsf(a / d2) / magnitudes[k];
                    float gr = fabsf(0.5f * (-1.0f / denom - a / d2)) / magnitudes[k + 1];
                    float sraw = gl + gc + gr;
                    CEPS(fabsf(raw) - 1.0f, sraw, 111);       /* the clamp itself (:492) */
                    sfrac = fabsf(raw) > 1.0f ? 0.0f : sraw;  /* a clamped delta is an exact +-1 */
                }
            }
            frac_bin = (float)k + delta;
        } else {
            frac_bin = (float)k;
        }
        frac_bins[k] = frac_bin;                                     /* :498 */
        sens_frac[k] = sfrac;
        float score = fund_mag;
        size_t last = (size_t)k;
        int longest_run = 0, current_run = 0, total_harms = 0;
        for (int n = 2; n <= MAX_HARMONICS; ++n) {                   /* :504 */
            float expected_f = frac_bin * (float)n;                  /* :505 */
            if (!frac_is_exact) margin_update(&margin, (expected_f - (float)half_size) / (float)n, 1);
            CEPS(expected_f - (float)half_size, (float)n * sfrac, 112);
            if (expected_f >= (float)half_size) break;               /* :506-508 */
            if (!frac_is_exact) {   /* distance of expected_f from the nearest integer, relative */
                float r = expected_f - floorf(expected_f);
                float d = r < 0.5f ? r : 1.0f - r;
                margin_update(&margin, d / (float)n, 2);   /* frac_bin perturbation that flips it */
            }
            size_t search_start = f32_as_usize(floorf(expected_f - 1.0f));   /* :509 */
            if (search_start < last + 1) search_start = last + 1;
            size_t search_end = f32_as_usize(ceilf(expected_f + 1.0f));      /* :510 */
            if (search_end > (size_t)half_size - 1) search_end = (size_t)half_size - 1;

            if (sfrac > 0.0f) {
                /* window boundaries: floor(e - 1) moves when e - 1 crosses an integer (one bin enters or
                 * leaves at the low end), ceil(e + 1) likewise at the high end; that matters only if the bin
                 * concerned is a peak the search could take */
                float lo = expected_f - 1.0f, hi = expected_f + 1.0f;
                float fl = floorf(lo), ce = ceilf(hi);
                long ifl = (long)fl, ice = (long)ce;
                float ns = (float)n * sfrac;
                if (ifl - 1 > (long)last && ifl - 1 >= 0 && ifl - 1 < half_size && is_peak[ifl - 1])
                    CEPS(lo - fl, ns, 113);                      /* bin fl-1 would enter */
                if (ifl > (long)last && ifl >= 0 && ifl < half_size && is_peak[ifl])
                    CEPS(fl + 1.0f - lo, ns, 113);               /* bin fl would leave   */
                if (ice > (long)last && ice >= 0 && ice <= half_size - 1 && is_peak[ice])
                    CEPS(hi - (ce - 1.0f), ns, 114);             /* bin ce would leave   */
                if (ice + 1 > (long)last && ice + 1 <= half_size - 1 && is_peak[ice + 1])
                    CEPS(ce - hi, ns, 114);                      /* bin ce+1 would enter */
            }
            size_t best_hbin = 0;                                    /* :512-520 */
            float best_mag = 0.0f, second_mag = -1.0f;
            for (size_t h = search_start; h <= search_end; ++h) {
                if (is_peak[h]) {
                    if (magnitudes[h] > best_mag) {
                        second_mag = best_mag;
                        best_mag = magnitudes[h];
                        best_hbin = h;
                    } else if (magnitudes[h] > second_mag) {
                        second_mag = magnitudes[h];
                    }
                }
            }
            if (best_hbin != 0 && second_mag > 0.0f) CEPS(best_mag - second_mag, 2.0f, 115);   /* :516 */
            if (best_hbin != 0) {                                    /* :521-531 */
                score += best_mag;
                last = best_hbin;
                current_run += 1;
                total_harms += 1;
            } else {
                if (current_run > longest_run) longest_run = current_run;
                current_run = 0;
            }
        }
        if (current_run > longest_run) longest_run = current_run;    /* :533-535 */
        if (longest_run < 3) CEPS(fund_mag - 15.0f * noise_floor[k], 1.0f + 15.0f * FSENS(k), 116);
        if (longest_run < 3 && fund_mag < 15.0f * noise_floor[k]) {  /* :536-537 */
            scores[k] = 0.0f;
        } else {                                                     /* :539-543 */
            const float STRUCT_BASE = 1.0f;
            float log_score = log2f(0.5f + score);
            float struct_mult = (STRUCT_BASE + (float)longest_run + (float)total_harms / 2.0f)
                                / (STRUCT_BASE + (float)MAX_HARMONICS);
            scores[k] = log_score * struct_mult;
            /* score = log2(0.5 + sum of 1 + total_harms magnitudes) * struct_mult */
            sens_score[k] = (float)(1 + total_harms) * fabsf(struct_mult) / ((0.5f + score) * 0.69314718f);
        }
    }
    if (diag) diag->n_scored = n_scored;

    float max_score = 0.0f;                                          /* :547 */
    for (int pi = 0; pi < n_peaks; ++pi) {
        float s = scores[peak_bins[pi]];
        if (s > max_score) max_score = s;          /* f32::max fold from 0.0 */
    }
    int n_out = 0;
    float sens_max = 0.0f;
    for (int pi = 0; pi < n_peaks; ++pi) {
        int k = peak_bins[pi];
        if (scores[k] != 0.0f) CEPS(scores[k], sens_score[k], 117);  /* sign of a score against the fold's 0.0 */
        if (scores[k] == max_score && sens_score[k] > sens_max) sens_max = sens_score[k];
    }
    if (max_score == 0.0f) goto done;                                /* :548-550 */
    {
        float cutoff = max_score * 0.5f;                             /* :551 */
        aao_cand *cand = (aao_cand *)malloc(sizeof(aao_cand) * (size_t)n_peaks);
        int nc = 0;
        for (int pi = 0; pi < n_peaks; ++pi) {                       /* :553-562 */
            int k = peak_bins[pi];
            if (scores[k] != 0.0f) margin_update(&margin, (scores[k] - cutoff) / cutoff, 3);
            if (scores[k] != 0.0f && scores[k] != max_score)
                CEPS(scores[k] - cutoff, sens_score[k] + 0.5f * sens_max, 118);
            if (scores[k] >= cutoff) { cand[nc].bin = k; cand[nc].score = scores[k]; ++nc; }
        }
        if (diag) diag->n_candidates = nc;

        uint8_t *suppressed = (uint8_t *)calloc((size_t)nc + 1, 1);  /* :566-583 */
        for (int i = 0; i < nc; ++i) {
            float freq_i = frac_bins[cand[i].bin] * bin_width;
            float score_i = cand[i].score;
            for (int j = 0; j < nc; ++j) {
                if (i == j) continue;
                float freq_j = frac_bins[cand[j].bin] * bin_width;
                float score_j = cand[j].score;
                float ratio = freq_i / freq_j;
                float nearest = roundf(ratio);     /* f32::round: half away from zero */
                {   /* margins: ratio near x.5, |ratio/nearest-1| near 0.03, score test */
                    float fr = ratio - floorf(ratio);
                    if (ratio > 1.4f && ratio < 5.6f) {
                        float fbi = frac_bins[cand[i].bin], fbj = frac_bins[cand[j].bin];
                        float sr_ = sens_frac[cand[i].bin] / fbj + fbi * sens_frac[cand[j].bin] / (fbj * fbj);
                        margin_update(&margin, fr - 0.5f, 4);
                        CEPS(fr - 0.5f, sr_, 119);
                        if (nearest >= 2.0f && nearest <= 5.0f) {
                            float dev = fabsf(ratio / nearest - 1.0f);
                            margin_update(&margin, (dev - 0.03f) / 0.03f, 5);
                            CEPS(dev - 0.03f, sr_ / nearest, 120);
                            if (dev < 0.03f) {
                                margin_update(&margin, (score_i - score_j * 1.05f) / score_i, 6);
                                CEPS(score_i - score_j * 1.05f,
                                     sens_score[cand[i].bin] + 1.05f * sens_score[cand[j].bin], 121);
                            }
                        }
                    }
                }
                if (nearest >= 2.0f && nearest <= 5.0f
                    && fabsf(ratio / nearest - 1.0f) < 0.03f
                    && score_i < score_j * 1.05f) {
                    suppressed[i] = 1;             /* `any` short-circuits; the loop goes on for the margins only */
                }
            }
        }
        int nk = 0;                                                  /* :584-589 */
        for (int i = 0; i < nc; ++i)
            if (!suppressed[i]) cand[nk++] = cand[i];
        free(suppressed);

        /* :591-592 sort_unstable_by descending score.  Order of exactly equal
         * scores is unspecified in the reference; this restatement keeps the
         * ascending-bin order for ties (stable insertion sort). */
        for (int i = 1; i < nk; ++i) {
            aao_cand c = cand[i];
            int j = i - 1;
            while (j >= 0 && cand[j].score < c.score) { cand[j + 1] = cand[j]; --j; }
            cand[j + 1] = c;
        }
        for (int i = 1; i < nk; ++i) {
            margin_update(&margin, (cand[i - 1].score - cand[i].score) / cand[i - 1].score, 7);
            CEPS(cand[i - 1].score - cand[i].score, sens_score[cand[i - 1].bin] + sens_score[cand[i].bin], 122);
        }

        const float MIN_BIN_SEPARATION = 2.0f;                       /* :594-605 */
        int nd = 0;
        for (int i = 0; i < nk; ++i) {
            float frac_i = frac_bins[cand[i].bin];
            int conflict = 0;
            for (int j = 0; j < nd; ++j) {
                float d = fabsf(frac_i - frac_bins[cand[j].bin]);
                margin_update(&margin, (d - MIN_BIN_SEPARATION) / MIN_BIN_SEPARATION, 8);
                CEPS(d - MIN_BIN_SEPARATION, sens_frac[cand[i].bin] + sens_frac[cand[j].bin], 123);
                if (d < MIN_BIN_SEPARATION) { conflict = 1; break; }
            }
            if (!conflict) cand[nd++] = cand[i];
        }
        if (nd > MAX_NOTES) nd = MAX_NOTES;                          /* :606 */

        for (int i = 0; i < nd; ++i) {                               /* :608-619 */
            float freq = frac_bins[cand[i].bin] * bin_width;
            margin_update(&margin, (freq - min_freq) / min_freq, 9);
            margin_update(&margin, (freq - max_freq) / max_freq, 10);
            CEPS(freq - min_freq, bin_width * sens_frac[cand[i].bin], 124);
            CEPS(freq - max_freq, bin_width * sens_frac[cand[i].bin], 124);
            if (freq >= min_freq && freq <= max_freq) {
                out_pairs[2 * n_out] = freq;
                out_pairs[2 * n_out + 1] = cand[i].score;
                if (diag) diag->out_bins[n_out] = cand[i].bin;
                ++n_out;
            }
        }
        free(cand);
    }
done:
    if (diag) {
        diag->n_out = n_out; diag->min_margin = margin; diag->margin_src = margin_src;
        diag->cand_eps = ceps; diag->cand_src = ceps_src;
    }
    free(is_peak); free(peak_bins); free(scores); free(frac_bins); free(sens_frac); free(sens_score);
    return n_out;
#undef FSENS
#undef CEPS
}

/* ------------------------------------------------------------------------- */
/* a8: PitchTracker, stft.rs:19-117                                            */
/*                                                                            */
/* Capacity: a track survives only while it was matched within the last three */
/* frames (max_life = 3) and each frame matches or creates at most 8 distinct */
/* tracks, so at most 24 tracks are alive between frames (32 during a frame)  */
/* and at most 16 have life >= display_threshold.                             */
/* ------------------------------------------------------------------------- */
#define AAO_TRACK_CAP 64
struct aao_tracker {
    int n;
    float freq[AAO_TRACK_CAP], score[AAO_TRACK_CAP];
    int life[AAO_TRACK_CAP];
};

aao_tracker *aao_tracker_create(void) { return (aao_tracker *)calloc(1, sizeof(aao_tracker)); }
void aao_tracker_destroy(aao_tracker *t) { free(t); }
void aao_tracker_reset(aao_tracker *t) { t->n = 0; }

int aao_tracker_process(aao_tracker *t, const float *raw_pairs, int n_raw, int onset,
                        float *out_pairs, int max_out)
{
    const int display_threshold = 2;   /* :39 */
    const int max_life = 3;            /* :40 */
    const float tolerance = 0.03f;     /* :41 */
    uint8_t matched[AAO_TRACK_CAP];
    int n_matched = t->n;                                           /* :46 */
    memset(matched, 0, sizeof(matched));

    for (int r = 0; r < n_raw; ++r) {                               /* :50-84 */
        float raw_freq = raw_pairs[2 * r], raw_score = raw_pairs[2 * r + 1];
        int found = 0;
        for (int i = 0; i < t->n; ++i) {
            if (matched[i]) continue;
            if (fabsf(t->freq[i] - raw_freq) / t->freq[i] < tolerance) {   /* :57 */
                if (onset) t->freq[i] = raw_freq;                           /* :61-65 */
                else t->freq[i] = t->freq[i] * 0.6f + raw_freq * 0.4f;
                t->score[i] = raw_score;
                t->life[i] = t->life[i] + 1 < max_life ? t->life[i] + 1 : max_life;
                matched[i] = 1;
                found = 1;
                break;
            }
        }
        if (!found && t->n < AAO_TRACK_CAP) {                        /* :76-83 */
            t->freq[t->n] = raw_freq;
            t->score[t->n] = raw_score;
            t->life[t->n] = 1;
            matched[t->n] = 1;
            t->n++;
            n_matched++;
        }
    }
    (void)n_matched;

    int n_out = 0;
    int i = 0;
    while (i < t->n) {                                              /* :90-113 */
        if (!matched[i]) {
            if (onset) t->life[i] = 0;
            else t->life[i] -= 1;
        }
        if (t->life[i] <= 0) {
            for (int j = i; j + 1 < t->n; ++j) {                    /* Vec::remove */
                t->freq[j] = t->freq[j + 1];
                t->score[j] = t->score[j + 1];
                t->life[j] = t->life[j + 1];
                matched[j] = matched[j + 1];
            }
            t->n--;
        } else {
            if (t->life[i] >= display_threshold) {
                if (n_out < max_out) {
                    out_pairs[2 * n_out] = t->freq[i];
                    out_pairs[2 * n_out + 1] = t->score[i];
                }
                ++n_out;
            }
            ++i;
        }
    }
    return n_out;
}

/* ------------------------------------------------------------------------- */
/* a10-a12: onset frame body, onset.rs:149-186 (state/consts), 261-357, 47-84  */
/* ------------------------------------------------------------------------- */
struct aao_onset {
    int half;
    float *prev_magnitude;        /* :149 */
    float *noise_floor_per_bin;   /* :175 */
    int floor_initialized;        /* :176 */
    float energy_ema;             /* :160 */
    float threshold;              /* FluxTracker.threshold :49,59 */
    uint32_t frames_since_onset;  /* :200, starts at 4 */
};

aao_onset *aao_onset_create(int half)
{
    aao_onset *s = (aao_onset *)calloc(1, sizeof(*s));
    s->half = half;
    s->prev_magnitude = (float *)calloc((size_t)half, sizeof(float));
    s->noise_floor_per_bin = (float *)calloc((size_t)half, sizeof(float));
    s->frames_since_onset = 4;
    return s;
}

void aao_onset_destroy(aao_onset *s)
{
    if (!s) return;
    free(s->prev_magnitude); free(s->noise_floor_per_bin); free(s);
}

const float *aao_onset_floor(const aao_onset *s) { return s->noise_floor_per_bin; }

/* the gates an onset has to pass before it is pushed (onset.rs:153 multiplier, :79 threshold floor, :356 burst gates,
 * :403 / :535 re-fire guard); exported so that tests can hold them against what the real crate logged
 * (tests/golden/ref_log_onsets.json) */
#define AAO_FLUX_MULTIPLIER      1.5f
#define AAO_FLUX_THRESHOLD_FLOOR 0.9f
#define AAO_EXCESS_GATE          3.0f
#define AAO_COUNT_GATE           3u
#define AAO_REFIRE_FRAMES        3u
void aao_onset_gates(float *flux_multiplier, float *flux_threshold_floor, float *excess_gate, uint32_t *count_gate,
                     uint32_t *refire_frames)
{
    *flux_multiplier = AAO_FLUX_MULTIPLIER;
    *flux_threshold_floor = AAO_FLUX_THRESHOLD_FLOOR;
    *excess_gate = AAO_EXCESS_GATE;
    *count_gate = AAO_COUNT_GATE;
    *refire_frames = AAO_REFIRE_FRAMES;
}

void aao_onset_reset(aao_onset *s)
{
    memset(s->prev_magnitude, 0, sizeof(float) * (size_t)s->half);
    memset(s->noise_floor_per_bin, 0, sizeof(float) * (size_t)s->half);
    s->floor_initialized = 0;
    s->energy_ema = 0.0f;
    s->threshold = 0.0f;
    s->frames_since_onset = 4;
}

void aao_onset_frame(aao_onset *s, const float *current_mags, float global_floor,
                     aao_features *out)
{
    const int half_size = s->half;
    const float ENERGY_EMA_RISE = 0.84f, ENERGY_EMA_DECAY = 0.95f;       /* :161-162 */
    const float BIN_BURST_RATIO = 2.5f, FLOOR_OVERCOMPENSATE = 1.3f;     /* :177-178 */
    const float FLOOR_RISE = 0.1f, FLOOR_DECAY = 0.04f;                  /* :179-180 */
    const float multiplier = AAO_FLUX_MULTIPLIER, rise_memory = 0.84f, decay_memory = 0.89f; /* :153 */

    float current_flux = 0.0f, frame_energy = 0.0f;                       /* :261-262 */
    for (int i = 0; i < half_size; ++i) {                                 /* :274-291 */
        float mag = current_mags[i];
        frame_energy += mag;
        float smoothed_mag;                                               /* :264-269 */
        if (i == 0 || i >= half_size - 1) smoothed_mag = current_mags[i];
        else smoothed_mag = (current_mags[i - 1] + current_mags[i] + current_mags[i + 1]) / 3.0f;
        float weight = 1.0f - ((float)i / (float)half_size);              /* :280 */
        float diff = smoothed_mag - s->prev_magnitude[i];
        if (diff > 0.0f) current_flux += diff * weight;
        s->prev_magnitude[i] = mag;                                       /* :290 */
    }

    float floor_eps = fmaxf(global_floor, 0.01f);                         /* :302 */
    if (!s->floor_initialized) {                                          /* :304-309 */
        for (int k = 0; k < half_size; ++k)
            s->noise_floor_per_bin[k] = fmaxf(current_mags[k], global_floor);
        s->floor_initialized = 1;
    }
    float max_bin_excess = 0.0f;                                          /* :311-332 */
    uint32_t bin_burst_count = 0;
    for (int k = 0; k < half_size; ++k) {
        float mag = current_mags[k];
        float floor_k = fmaxf(s->noise_floor_per_bin[k], floor_eps);
        float r = mag / floor_k;
        if (r > BIN_BURST_RATIO) {
            bin_burst_count += 1;
            s->noise_floor_per_bin[k] = mag * FLOOR_OVERCOMPENSATE;
        } else if (mag > s->noise_floor_per_bin[k]) {
            s->noise_floor_per_bin[k] += FLOOR_RISE * (mag - s->noise_floor_per_bin[k]);
        } else {
            s->noise_floor_per_bin[k] += FLOOR_DECAY * (mag - s->noise_floor_per_bin[k]);
        }
        if (r > max_bin_excess) max_bin_excess = r;
    }
    if (bin_burst_count < 2) current_flux = 0.0f;                         /* :337-339 */

    float ema_memory = frame_energy > s->energy_ema ? ENERGY_EMA_RISE : ENERGY_EMA_DECAY;
    s->energy_ema = s->energy_ema * ema_memory + frame_energy * (1.0f - ema_memory); /* :350 */

    /* FluxTracker::update, :67-83 */
    float memory = current_flux > s->threshold ? rise_memory : decay_memory;
    int is_onset = current_flux > s->threshold;
    s->threshold = s->threshold * memory + current_flux * (1.0f - memory);
    if (s->threshold < AAO_FLUX_THRESHOLD_FLOOR) s->threshold = AAO_FLUX_THRESHOLD_FLOOR;
    int flux_onset = is_onset && current_flux > (s->threshold * multiplier);

    int bin_burst_onset = max_bin_excess > AAO_EXCESS_GATE && bin_burst_count >= AAO_COUNT_GATE;  /* :356 */
    int onset_detected = flux_onset && bin_burst_onset;                   /* :357 */
    int energy_rising = frame_energy > s->energy_ema * 1.5f;              /* :373 */
    /* offline reading of :383-456 and :535-539: no metronome ticks to guard against
     * (suppressed_by_tick = false) and calibration already done */
    int onset_fired = onset_detected && energy_rising && s->frames_since_onset >= AAO_REFIRE_FRAMES;   /* :403 */
    if (onset_fired || (onset_detected && s->frames_since_onset < AAO_REFIRE_FRAMES)) s->frames_since_onset = 0;  /* :535 */
    else if (s->frames_since_onset != 0xffffffffu) s->frames_since_onset += 1;          /* saturating_add */

    out->flux = current_flux;
    out->energy = frame_energy;
    out->burst_count = bin_burst_count;
    out->max_excess = max_bin_excess;
    out->energy_ema = s->energy_ema;
    out->flags = (flux_onset ? AAO_FLAG_FLUX_ONSET : 0u) | (bin_burst_onset ? AAO_FLAG_BURST_ONSET : 0u)
               | (onset_detected ? AAO_FLAG_ONSET_DETECTED : 0u)
               | (energy_rising ? AAO_FLAG_ENERGY_RISING : 0u)
               | (onset_fired ? AAO_FLAG_ONSET_FIRED : 0u);
}

/* f2: Note::from_freq, analysis/theory.rs:195-209 */
void aao_note_from_freq(float freq, float base_freq, int *octave, int *semis, float *cents)
{
    float base = base_freq * powf(2.0f, -4.75f);                  /* :197 */
    float lg = log2f(freq / base) * 1200.0f;                      /* :198 */
    float o = (lg + 50.0f) / 1200.0f;                             /* :199 `as u8` saturates */
    int oct = !(o > 0.0f) ? 0 : (o >= 255.0f ? 255 : (int)o);
    float sm = fmodf(roundf(lg / 100.0f), 12.0f);                 /* :200 `as usize` saturates */
    int se = !(sm > 0.0f) ? 0 : (int)sm;
    float c = fmodf(lg, 100.0f);                                  /* :201 */
    c = c < 50.0f ? c : -(100.0f - c);                            /* :202-206 */
    *octave = oct;
    *semis = se;
    *cents = c;
}

/* a13 (NEW, no reference): spectral centroid in Hz */
float aao_centroid(const float *mags, int half, float bin_width)
{
    double num = 0.0, den = 0.0;
    for (int k = 0; k < half; ++k) {
        num += (double)k * (double)mags[k];
        den += (double)mags[k];
    }
    if (!(den > 0.0)) return 0.0f;
    return (float)(num / den * (double)bin_width);
}

/* a14 (NEW, no reference): YIN-style lag search, float64.
 * d(tau) = sum_{j<W} (x[j]-x[j+tau])^2 with W = n - max_lag;
 * d'(0) = 1, d'(tau) = d(tau) * tau / sum_{j=1..tau} d(j).
 * Pick the first tau >= min_lag with d'(tau) < threshold, then descend to the
 * local minimum; otherwise the global argmin over [min_lag, max_lag]. */
int aao_yin_lag(const float *frame, int n, int min_lag, int max_lag, float threshold,
                double *cmnd_out)
{
    if (max_lag >= n) max_lag = n - 1;
    if (min_lag < 1) min_lag = 1;
    if (min_lag > max_lag) return 0;
    int W = n - max_lag;
    double *dp = (double *)malloc(sizeof(double) * (size_t)(max_lag + 1));
    double run = 0.0;
    dp[0] = 1.0;
    for (int tau = 1; tau <= max_lag; ++tau) {
        double d = 0.0;
        for (int j = 0; j < W; ++j) {
            double e = (double)frame[j] - (double)frame[j + tau];
            d += e * e;
        }
        run += d;
        dp[tau] = run > 0.0 ? d * (double)tau / run : 1.0;
    }
    int best = 0;
    for (int tau = min_lag; tau <= max_lag; ++tau) {
        if (dp[tau] < (double)threshold) {
            while (tau + 1 <= max_lag && dp[tau + 1] < dp[tau]) ++tau;
            best = tau;
            break;
        }
    }
    if (best == 0) {
        best = min_lag;
        for (int tau = min_lag; tau <= max_lag; ++tau)
            if (dp[tau] < dp[best]) best = tau;
    }
    if (cmnd_out) memcpy(cmnd_out, dp, sizeof(double) * (size_t)(max_lag + 1));
    free(dp);
    return best;
}

/* ------------------------------------------------------------------------- */
/* Frame loop on an offline clip: stft.rs:273-438 / onset.rs:244-543 with the  */
/* ring buffer replaced by direct indexing (frame t = samples [t*hop,t*hop+n)).*/
/* ------------------------------------------------------------------------- */
int64_t aao_num_frames(int64_t len, int n, int hop)
{
    if (len < n || hop <= 0) return 0;
    return (len - n) / hop + 1;
}

int64_t aao_analyze_clip(const aao_config *cfg, const float *samples, int64_t len,
                         const float *mags_in, const uint8_t *onset_in,
                         float *mags_out, float *floor_out, uint8_t *peak_mask_out,
                         aao_features *feat_out, aao_stable *stable_out,
                         aao_pitch_diag *diag_out)
{
    aao_taps taps;
    memset(&taps, 0, sizeof(taps));
    taps.mags = mags_out;
    taps.floor = floor_out;
    taps.peak_mask = peak_mask_out;
    taps.features = feat_out;
    taps.stable = stable_out;
    taps.diag = diag_out;
    return aao_analyze_clip_ex(cfg, samples, len, mags_in, onset_in, &taps);
}

int64_t aao_analyze_clip_ex(const aao_config *cfg, const float *samples, int64_t len,
                            const float *mags_in, const uint8_t *onset_in, const aao_taps *taps)
{
    const int n = cfg->n, hop = cfg->hop, half = n / 2 + 1;
    const int64_t T = aao_num_frames(len, n, hop);
    if (T <= 0) return 0;
    float *mags_out = taps->mags, *floor_out = taps->floor;
    uint8_t *peak_mask_out = taps->peak_mask;
    aao_features *feat_out = taps->features;
    aao_stable *stable_out = taps->stable;
    aao_pitch_diag *diag_out = taps->diag;

    float *window = (float *)malloc(sizeof(float) * (size_t)n);
    float *td = (float *)malloc(sizeof(float) * (size_t)n);
    float *spec = (float *)malloc(sizeof(float) * 2 * (size_t)half);
    float *mags = (float *)malloc(sizeof(float) * (size_t)half);
    float *eff = (float *)malloc(sizeof(float) * (size_t)half);
    float *fsens = (float *)malloc(sizeof(float) * (size_t)half);
    aao_hann_window(n, window);                              /* stft.rs:179 */
    aao_fft *fft = mags_in ? NULL : aao_fft_create(n);       /* stft.rs:180 */
    aao_pitch_floor *pf = aao_pitch_floor_create(half);
    aao_onset *on = aao_onset_create(half);
    aao_tracker *trk = aao_tracker_create();
    const float bin_width = cfg->sample_rate / (float)n;     /* stft.rs:320 */
    const float global_floor = aao_global_floor(cfg->noise_floor_db, half);

    for (int64_t t = 0; t < T; ++t) {
        if (mags_in) {
            memcpy(mags, mags_in + t * half, sizeof(float) * (size_t)half);
        } else {
            const float *x = samples + t * hop;
            for (int i = 0; i < n; ++i) td[i] = x[i] * window[i];   /* stft.rs:296-299 */
            aao_fft_forward(fft, td, spec);                         /* stft.rs:306 */
            aao_magnitudes(spec, half, mags);                       /* stft.rs:314-318 */
        }
        if (mags_out) memcpy(mags_out + t * half, mags, sizeof(float) * (size_t)half);

        aao_features f;
        memset(&f, 0, sizeof(f));
        if (cfg->features & AAO_FEAT_PITCH) {
            aao_pitch_floor_update(pf, mags, global_floor, eff);
            if (floor_out) memcpy(floor_out + t * half, eff, sizeof(float) * (size_t)half);
            if (taps->pitch_nf) memcpy(taps->pitch_nf + t * half, aao_pitch_floor_nf(pf), sizeof(float) * (size_t)half);
            if (taps->pitch_vol) memcpy(taps->pitch_vol + t * half, aao_pitch_floor_vol(pf), sizeof(float) * (size_t)half);
            /* where the effective floor sits at its clamp it does not move with the magnitudes */
            for (int k = 0; k < half; ++k) fsens[k] = aao_pitch_floor_nf(pf)[k] > global_floor * 2.5f ? 0.0f : 1.0f;
            float pairs[2 * AAO_MAX_NOTES];
            int np = extract_pitches_impl(mags, half, bin_width, cfg->min_freq, cfg->max_freq, eff, fsens,
                                          pairs, peak_mask_out ? peak_mask_out + t * half : NULL,
                                          diag_out ? diag_out + t : NULL);
            f.n_pitches = (uint32_t)np;
            for (int i = 0; i < np; ++i) {
                f.pitch[i].freq = pairs[2 * i];
                f.pitch[i].score = pairs[2 * i + 1];
            }
            if ((cfg->features & AAO_FEAT_TRACKER) && stable_out) {
                float sp[2 * AAO_MAX_STABLE];
                int ns = aao_tracker_process(trk, pairs, np, onset_in ? onset_in[t] : 0, sp,
                                             AAO_MAX_STABLE);
                aao_stable *so = stable_out + t;
                memset(so, 0, sizeof(*so));
                so->n = (uint32_t)(ns < AAO_MAX_STABLE ? ns : AAO_MAX_STABLE);
                for (uint32_t i = 0; i < so->n; ++i) {
                    so->pitch[i].freq = sp[2 * i];
                    so->pitch[i].score = sp[2 * i + 1];
                }
            }
        }
        if (cfg->features & AAO_FEAT_ONSET) {
            aao_onset_frame(on, mags, global_floor, &f);
            if (taps->onset_nf) memcpy(taps->onset_nf + t * half, aao_onset_floor(on), sizeof(float) * (size_t)half);
        }
        if (cfg->features & AAO_FEAT_CENTROID) f.centroid = aao_centroid(mags, half, bin_width);
        if (feat_out) feat_out[t] = f;
    }

    aao_tracker_destroy(trk);
    aao_onset_destroy(on);
    aao_pitch_floor_destroy(pf);
    if (fft) aao_fft_destroy(fft);
    free(fsens); free(eff); free(mags); free(spec); free(td); free(window);
    return T;
}

/* ------------------------------------------------------------------------- */
/* Clip-parallel driver (timed CPU baseline)                                   */
/* ------------------------------------------------------------------------- */
typedef struct {
    const aao_config *cfg;
    const float *clips;
    int64_t n_clips, clip_len, T;
    int tid, n_threads;
    float *mags_out;
    aao_features *feat_out;
    aao_stable *stable_out;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    const int half = j->cfg->n / 2 + 1;
    for (int64_t c = j->tid; c < j->n_clips; c += j->n_threads) {
        aao_analyze_clip(j->cfg, j->clips + c * j->clip_len, j->clip_len, NULL, NULL,
                         j->mags_out ? j->mags_out + c * j->T * half : NULL, NULL, NULL,
                         j->feat_out ? j->feat_out + c * j->T : NULL,
                         j->stable_out ? j->stable_out + c * j->T : NULL, NULL);
    }
    return NULL;
}

int64_t aao_analyze_batch(const aao_config *cfg, const float *clips, int64_t n_clips,
                          int64_t clip_len, int n_threads, float *mags_out,
                          aao_features *feat_out, aao_stable *stable_out)
{
    const int64_t T = aao_num_frames(clip_len, cfg->n, cfg->hop);
    if (T <= 0 || n_clips <= 0) return 0;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    batch_job jobs[256];
    for (int i = 0; i < n_threads; ++i) {
        jobs[i] = (batch_job){cfg, clips, n_clips, clip_len, T, i, n_threads,
                              mags_out, feat_out, stable_out};
        pthread_create(&th[i], NULL, batch_worker, &jobs[i]);
    }
    for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
    return T * n_clips;
}

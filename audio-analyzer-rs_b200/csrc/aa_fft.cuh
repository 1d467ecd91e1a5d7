// aa_fft.cuh -- register-resident Stockham FFT building blocks (sm_100a).
//
// Replaces realfft::RealToComplex::process_with_scratch as called from
// FftProcessor::process_forward (reference src/dsp/fft.rs:33, 66-71): an n-point
// real transform is computed as an n/2-point complex FFT of z[m] = x[2m] + i x[2m+1]
// followed by a split post-pass.
//
// Thread mapping (verified by tools/stockham_model.py): a CTA of NT = N2/E threads
// owns one frame; in every pass thread t holds v[m] <-> element t + m*NT of that
// pass's input.  A radix-R pass does E/R butterflies per thread; butterfly b uses
// v[b + r*E/R], r < R.  Results go through a padded shared-memory exchange buffer
// except after the last pass, whose outputs land in place (v[m] = Z[t + m*NT]).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace aa {

// Complex arithmetic on the packed f32x2 pipe (sm_100: FADD2 / FMUL2 / FFMA2 do two IEEE f32 operations per
// issue slot; operand swaps, per-half negation and scalar broadcast are free operand modifiers).  The kernels
// built from these are issue-bound, so a complex add is ONE instruction and a complex multiply TWO.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    // a.x * (b.x, b.y) + a.y * (-b.y, b.x).  The half-negation must sit on the FFMA2 ADDEND (a free `.NP` operand
    // modifier): FMUL2 has no per-half negation, so negating b.y inside the multiply costs a MOV + FADD to build the
    // pair (-b.y, b.x) in registers -- four instructions per complex multiply instead of two.  RN(-x) == -RN(x), so
    // the result is bit-identical either way.
    const float2 t = __fmul2_rn(make_float2(a.y, a.y), make_float2(b.y, b.x));     // (a.y b.y, a.y b.x)
    return __ffma2_rn(make_float2(a.x, a.x), b, make_float2(-t.x, t.y));
}
// a * (cx + i cy) with a compile-time constant factor: the constants ride as scalar-broadcast immediates of FMUL2 /
// FFMA2 (two instructions; cmul() with a constant pair makes ptxas materialise the pair in registers first)
__device__ __forceinline__ float2 cmul_const(float2 a, float cx, float cy)
{
    const float2 t = __fmul2_rn(make_float2(a.y, a.x), make_float2(cy, cy));       // (a.y cy, a.x cy)
    return __ffma2_rn(a, make_float2(cx, cx), make_float2(-t.x, t.y));
}
__device__ __forceinline__ float2 cmul_negi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
__device__ __forceinline__ float2 cmul_posi(float2 a) { return make_float2(-a.y, a.x); }  // a * (+i)
__device__ __forceinline__ float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }

// Table loads (window, twiddles): read-only path.  With AA_TAB_EVICT_LAST the lines are marked evict-last in L1:
// the fused analysis kernel leaves the SM only ~23 KB of L1, which the tables (~24 KB hot) share with the output
// stores and whatever else passes through.
__device__ __forceinline__ float2 ld_table(const float2 *p)
{
#ifdef AA_TAB_EVICT_LAST
    float2 v;
    asm("ld.global.nc.L1::evict_last.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
// streaming output store: with AA_ST_NO_ALLOCATE it does not allocate a line in L1
__device__ __forceinline__ void st_stream(float *p, float v)
{
#ifdef AA_ST_NO_ALLOCATE
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
#else
    *p = v;
#endif
}

// exchange-buffer padding: one float2 every 16 keeps both the strided writes of the
// first pass and the unit-stride reads at the ideal wavefront count.
__host__ __device__ __forceinline__ constexpr int padidx(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int padded_len(int n) { return n + (n >> 4) + 1; }

template <int R>
struct Bfly;

template <>
struct Bfly<2> {
    static __device__ __forceinline__ void run(float2 (&x)[2])
    {
        float2 a = x[0];
        x[0] = cadd(a, x[1]);
        x[1] = csub(a, x[1]);
    }
};

template <>
struct Bfly<4> {
    static __device__ __forceinline__ void run(float2 (&x)[4])
    {
        float2 a0 = cadd(x[0], x[2]);
        float2 a1 = csub(x[0], x[2]);
        float2 a2 = cadd(x[1], x[3]);
        float2 a3 = cmul_negi(csub(x[1], x[3]));
        x[0] = cadd(a0, a2);
        x[1] = cadd(a1, a3);
        x[2] = csub(a0, a2);
        x[3] = csub(a1, a3);
    }
};

template <>
struct Bfly<8> {
    static __device__ __forceinline__ void run(float2 (&x)[8])
    {
        constexpr float C = 0.70710678118654752440f;
        float2 e[4] = {x[0], x[2], x[4], x[6]};
        float2 o[4] = {x[1], x[3], x[5], x[7]};
        Bfly<4>::run(e);
        Bfly<4>::run(o);
        // o[k] *= exp(-2 pi i k / 8)
        float2 o1 = cscale(cadd(o[1], cmul_negi(o[1])), C);      //  C * (x + y, y - x)
        float2 o2 = cmul_negi(o[2]);
        float2 o3 = cscale(cadd(o[3], cmul_posi(o[3])), -C);     // -C * (x - y, x + y)
        x[0] = cadd(e[0], o[0]);
        x[4] = csub(e[0], o[0]);
        x[1] = cadd(e[1], o1);
        x[5] = csub(e[1], o1);
        x[2] = cadd(e[2], o2);
        x[6] = csub(e[2], o2);
        x[3] = cadd(e[3], o3);
        x[7] = csub(e[3], o3);
    }
};

template <>
struct Bfly<16> {
    static __device__ __forceinline__ void run(float2 (&x)[16])
    {
        constexpr float C = 0.70710678118654752440f;
        constexpr float C1 = 0.92387953251128675613f;  // cos(pi/8)
        constexpr float S1 = 0.38268343236508977173f;  // sin(pi/8)
        float2 e[8] = {x[0], x[2], x[4], x[6], x[8], x[10], x[12], x[14]};
        float2 o[8] = {x[1], x[3], x[5], x[7], x[9], x[11], x[13], x[15]};
        Bfly<8>::run(e);
        Bfly<8>::run(o);
        // o[k] *= exp(-2 pi i k / 16) = (cos(k pi/8), -sin(k pi/8))
        float2 w[8];
        w[0] = o[0];
        w[1] = cmul_const(o[1], C1, -S1);
        w[2] = cscale(cadd(o[2], cmul_negi(o[2])), C);
        w[3] = cmul_const(o[3], S1, -C1);
        w[4] = cmul_negi(o[4]);
        w[5] = cmul_const(o[5], -S1, -C1);
        w[6] = cscale(cadd(o[6], cmul_posi(o[6])), -C);
        w[7] = cmul_const(o[7], -C1, -S1);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x[k] = cadd(e[k], w[k]);
            x[k + 8] = csub(e[k], w[k]);
        }
    }
};

// padidx(j0 + r*NS) == padidx(j0) + r*NS + ((r*NS) >> 4) for every exchange of the plans used here
// (j0 & 15 never carries into bit 4 when r*NS is added: see tools/stockham_model.py --check-pad), so
// one padded base address per butterfly is enough and the per-element offsets are immediates.
__host__ __device__ constexpr int padoff(int d) { return d + (d >> 4); }

// One Stockham pass.  NS = product of the radices already applied.  tw[i] =
// exp(-2 pi i * i / N2).  The butterfly's R-1 twiddles w^r (w = tw[k*TSTEP]) are built from ONE
// table load by a product tree of depth <= 4 (<= 4 extra roundings, far inside the magnitude
// tolerance) instead of R-1 dependent L2-latency loads.  If !LAST the outputs are scattered into
// `exch` (padded indexing); the caller synchronises and reloads with fft_reload().
// TWS: the twiddle table belongs to a transform TWS times longer (sub-transforms of fft_run_split).
// PRE: the butterflies' first twiddles w^1 were loaded by the caller (w1pre[b]) -- fft_run issues them before
// the barrier in front of the pass, so a load that misses L1 overlaps the barrier wait and the reload.
template <int N2, int E, int R, int NS, bool LAST, int TWS = 1, bool PRE = false>
__device__ __forceinline__ void fft_pass(float2 (&v)[E], int t, float2 *exch,
                                         const float2 *__restrict__ tw, const float2 *w1pre = nullptr)
{
    constexpr int NT = N2 / E;
    constexpr int BPT = E / R;
    static_assert(E % R == 0, "radix must divide the per-thread element count");
#pragma unroll
    for (int b = 0; b < BPT; ++b) {
        const int j = t + b * NT;
        float2 x[R];
#pragma unroll
        for (int r = 0; r < R; ++r) x[r] = v[b + r * BPT];
        const int k = j & (NS - 1);
        if (NS > 1) {
            constexpr int TSTEP = TWS * (N2 / (NS * R));
            float2 w[R];
            w[1] = PRE ? w1pre[b] : ld_table(&tw[k * TSTEP]);
#pragma unroll
            for (int r = 2; r < R; ++r) {
                const int hi = (r >= 8) ? 8 : (r >= 4 ? 4 : 2);   // largest power of two <= r
                const int lo = r - hi;
                w[r] = lo ? cmul(w[hi], w[lo]) : cmul(w[hi / 2], w[hi / 2]);
            }
#pragma unroll
            for (int r = 1; r < R; ++r) x[r] = cmul(x[r], w[r]);
        }
        Bfly<R>::run(x);
        if (LAST) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[b + r * BPT] = x[r];
        } else {
            const int j0 = (j / NS) * (NS * R) + k;
            float2 *dst = exch + padidx(j0);
#pragma unroll
            for (int r = 0; r < R; ++r) dst[padoff(r * NS)] = x[r];
        }
    }
}

template <int N2, int E>
__device__ __forceinline__ void fft_reload(float2 (&v)[E], int t, const float2 *exch)
{
    constexpr int NT = N2 / E;
    static_assert(NT % 16 == 0, "reload offsets assume a thread count that is a multiple of 16");
    const float2 *src = exch + padidx(t);
#pragma unroll
    for (int m = 0; m < E; ++m) v[m] = src[padoff(m * NT)];
}

// Per-window-size geometry: E complex elements per thread, NT = N/2/E threads.
// E trades registers / unrolled code size (instruction-cache footprint) against the number of
// shared-memory exchanges: 2048 = 16*16*8 (3 passes) or 8*8*8*4 (4 passes).
#ifndef AA_E_BIG
#define AA_E_BIG 8
#endif
template <int N>
struct Geo;
template <>
struct Geo<4096> { static constexpr int E = AA_E_BIG; };
template <>
struct Geo<2048> { static constexpr int E = AA_E_BIG; };
template <>
struct Geo<1024> { static constexpr int E = 8; };
template <>
struct Geo<512> { static constexpr int E = 8; };
template <>
struct Geo<256> { static constexpr int E = 4; };

// number of Stockham passes for an N2-point transform with at most E points per thread
__host__ __device__ constexpr int fft_num_passes(int n2, int e)
{
    int p = 0;
    for (int ns = 1; ns < n2; ns *= (n2 / ns < e ? n2 / ns : e)) ++p;
    return p;
}

// Generic pass sequence: radix min(E, remaining) per pass, exchange buffers alternate
// (pass p writes bufs[p & 1]); `sync` is the barrier among the participating threads and
// `after_first` runs once right after the first barrier (used to refill the hop ring).
// On return v[m] = Z[t + m*NT]; the exchange buffer NOT read last is bufs[(passes - 1) & 1].
// PF: the next pass's twiddle load is issued before the barrier in front of it (two registers across the
// barrier; the register-capped batch kernels do without)
template <int N2, int E, int NS, int PASS, bool PF = false, class Sync, class Hook>
__device__ __forceinline__ void fft_run(float2 (&v)[E], int t, float2 *buf0, float2 *buf1,
                                        const float2 *__restrict__ tw, Sync sync, Hook after_first,
                                        const float2 *w1cur = nullptr)
{
    constexpr int NT = N2 / E;
    constexpr int REM = N2 / NS;
    constexpr int R = REM < E ? REM : E;
    constexpr bool LAST = (NS * R == N2);
    constexpr bool PRE = PF && PASS > 0;
    float2 *ex = (PASS & 1) ? buf1 : buf0;
    fft_pass<N2, E, R, NS, LAST, 1, PRE>(v, t, ex, tw, w1cur);
    if constexpr (!LAST) {
        constexpr int NSn = NS * R;                       // geometry of the next pass
        constexpr int REMn = N2 / NSn;
        constexpr int Rn = REMn < E ? REMn : E;
        constexpr int BPTn = E / Rn;
        constexpr int TSTEPn = N2 / (NSn * Rn);
        float2 w1n[BPTn];
        if constexpr (PF) {
#pragma unroll
            for (int b = 0; b < BPTn; ++b) w1n[b] = ld_table(&tw[((t + b * NT) & (NSn - 1)) * TSTEPn]);
        }
        sync();
        if constexpr (PASS == 0) after_first();
        fft_reload<N2, E>(v, t, ex);
        fft_run<N2, E, NSn, PASS + 1, PF>(v, t, buf0, buf1, tw, sync, after_first, w1n);
    }
}

// Full N/2-point complex FFT of v (in: v[m] = z[t + m*NT]; out: v[m] = Z[t + m*NT]) for a whole CTA.
// exA / exB: two padded exchange buffers of padded_len(N/2) float2 each.  Contains __syncthreads().
template <int N>
__device__ __forceinline__ void fft_half_complex(float2 (&v)[Geo<N>::E], int t, float2 *exA,
                                                 float2 *exB, const float2 *__restrict__ tw)
{
    fft_run<N / 2, Geo<N>::E, 1, 0>(v, t, exA, exB, tw, [] { __syncthreads(); }, [] {});
}

// ---------------------------------------------------------------------------------------------------
// Split plan for the fused analysis kernel (N2 = 8 * M, E = 8, NT = M threads): two block-wide exchanges
// instead of one per Stockham pass.
//   Z[k1 + 8 k2] = sum_t W_M^(t k2) * [ W_N2^(t k1) * sum_m z[t + M m] W_8^(m k1) ]
//   1. thread t: radix-8 butterfly over its own v[m] = z[t + M m], then times W_N2^(t k1)      (registers)
//   2. block exchange: sub-transform k1 (M points) goes to the G = M/8 consecutive threads k1*G .. k1*G + G - 1
//      (one warp at N2 = 2048, half a warp at N2 = 1024)
//   3. each group runs the M-point Stockham passes on its own region of exA; only __syncwarp() inside
//   4. the results go to exB in natural order (block exchange), where the real-FFT post-pass finds both
//      Z[k] and its partner Z[N2 - k]
// `sync` is the block barrier, `after_first` runs right after the first one.  On return (after the second
// barrier) exB[padidx(k)] = Z[k] for every k.
template <int M>
struct SplitGeo {
    static constexpr int G = M / 8;                              // threads per sub-transform
    static constexpr int RS = (padidx(M - 1) + 1 + 7) & ~7;      // float2 units per region (272 / 136)
};

template <int M, int NS, int TWS>
__device__ __forceinline__ void fft_run_local(float2 (&v)[8], int tl, float2 *reg, const float2 *__restrict__ tw)
{
    constexpr int REM = M / NS;
    constexpr int R = REM < 8 ? REM : 8;
    constexpr bool LAST = (NS * R == M);
    fft_pass<M, 8, R, NS, LAST, TWS>(v, tl, reg, tw);
    if constexpr (!LAST) {
        __syncwarp();
        fft_reload<M, 8>(v, tl, reg);
        __syncwarp();                        // the next pass writes the same region
        fft_run_local<M, NS * R, TWS>(v, tl, reg, tw);
    }
}

template <int N2, class Sync, class Hook>
__device__ __forceinline__ void fft_run_split(float2 (&v)[8], int t, float2 *exA, float2 *exB,
                                              const float2 *__restrict__ tw, Sync sync, Hook after_first)
{
    constexpr int M = N2 / 8, NT = M, G = SplitGeo<M>::G, RS = SplitGeo<M>::RS;
    static_assert(G == 32 || G == 16, "a sub-transform must live inside one warp");
    // 1. radix-8 over m (decimation in frequency), twiddles W_N2^(t k1) from one table load
    Bfly<8>::run(v);
    {
        float2 w[8];
        w[1] = __ldg(&tw[t]);
#pragma unroll
        for (int r = 2; r < 8; ++r) {
            const int hi = r >= 4 ? 4 : 2;
            const int lo = r - hi;
            w[r] = lo ? cmul(w[hi], w[lo]) : cmul(w[hi / 2], w[hi / 2]);
        }
#pragma unroll
        for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], w[r]);
    }
    // 2. to the groups
#pragma unroll
    for (int r = 0; r < 8; ++r) exA[r * RS + t] = v[r];
    sync();
    after_first();
    const int k1 = t / G, tl = t % G;
    float2 *reg = exA + k1 * RS;
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = reg[tl + m * G];
    __syncwarp();                            // the passes below scatter into the same region
    // 3. M-point transform inside the group
    fft_run_local<M, 1, 8>(v, tl, reg, tw);
    // 4. v[m] = Z[k1 + 8 (tl + G m)] = Z[(k1 + 8 tl) + NT m], natural order in exB
    float2 *dst = exB + padidx(k1 + 8 * tl);
#pragma unroll
    for (int m = 0; m < 8; ++m) dst[padoff(m * NT)] = v[m];
    sync();
}

// realfft's split post-pass for one pair: a = Z[k], b = Z[N/2 - k], tw = 0.5*exp(-2 pi i k/N).
// Produces X[k] (lo) and X[N/2 - k] (hi).  For k = 0 (b = Z[0], tw = (0.5, 0)) this yields
// X[0] = (re+im, 0) and X[N/2] = (re-im, 0).
__device__ __forceinline__ void rfft_postpass(float2 a, float2 b, float2 tw, float2 &lo, float2 &hi)
{
    const float2 bc = make_float2(b.x, -b.y);
    const float2 P = cadd(a, bc);                 // (sum_re, diff_im)
    const float2 Q = csub(a, bc);                 // (diff_re, sum_im)
    // ot = sum_im * tw + diff_re * (tw.y, -tw.x)
    // (negation on the FFMA2 addend, not inside the FMUL2: see cmul)
    const float2 tq = __fmul2_rn(make_float2(Q.x, Q.x), make_float2(tw.y, tw.x));
    const float2 ot = __ffma2_rn(make_float2(Q.y, Q.y), tw, make_float2(tq.x, -tq.y));
    const float2 h = make_float2(0.5f, 0.5f);
    lo = __ffma2_rn(P, h, ot);
    const float2 hc = __ffma2_rn(P, h, make_float2(-ot.x, -ot.y));   // conj(X[N/2 - k])
    hi = make_float2(hc.x, -hc.y);
}

}  // namespace aa

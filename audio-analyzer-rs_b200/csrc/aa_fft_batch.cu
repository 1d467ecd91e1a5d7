// aa_fft_batch.cu -- batched real FFT kernels behind aa_fft_forward / aa_fft_inverse:
// the FftProcessor::process_forward / process_inverse replacement
// (reference src/dsp/fft.rs:33-41, 66-71, 96-101).
//
// One CTA transforms one frame at a time (grid-stride over the batch): coalesced float2
// loads straight into the register-resident Stockham FFT of aa_fft.cuh, the realfft
// split post-pass, coalesced float2 stores of the n/2+1 bins.
#include <atomic>
#include "aa_fft.cuh"
#include "aa_internal.h"
#include "aa_tma.cuh"

namespace aa {

// These two kernels are HBM-bound: what matters is how many frames an SM has in flight, not the instruction
// count.  The packed-arithmetic FFT core wants more registers (aligned pairs); capping the kernels at 64
// registers keeps 1024 threads per SM resident (measured at n = 4096 / 2048 / 1024: 71 / 69 / 76 % of the HBM
// peak without the cap, 76 / 70 / 95 % with it).  One-warp CTAs (n <= 512) are limited by the 32 CTAs an SM
// can hold, not by registers, and lose with the cap.
// Geometry of the batched kernels (independent of the fused analysis kernel's Geo<N>): E complex points per
// thread, NT = N/2/E threads per frame, FPB frames per CTA.  The short lengths pack several frames into a CTA:
// a 256-point frame is 16 threads x 8 points (two frames per warp, radix 8 * 8 * 2: three passes instead of the
// four of a one-warp radix-4 plan, and half the warp-instructions per frame -- at n = 256 the kernel is bound by
// issue slots, not by HBM), and four-warp CTAs lift the resident warps per SM past what 32 one-warp CTAs give.
template <int N>
struct BatchGeo {
    static constexpr int E = Geo<N>::E;
    static constexpr int FPB = 1;
};
#ifndef AA_FFT_E_BIG
#define AA_FFT_E_BIG 16            // points per thread at n = 4096 / 2048: radix 16 * 16 * 8 -- three passes / two exchanges
                                   // instead of four / three.  With E = 8 these two lengths sit at 0.78 / 0.74 of the HBM
                                   // peak with `mio_throttle` + `short_scoreboard` on top of the stall list (ncu,
                                   // profiles/r02): 68 KB of shared-memory traffic per 16 KB of HBM traffic puts the
                                   // shared-memory pipe at ~2/3 of ITS peak; one exchange less, at 128 registers and
                                   // 512 threads per SM, gives 0.90 / 0.94
#endif
template <>
struct BatchGeo<4096> {
    static constexpr int E = AA_FFT_E_BIG;
    static constexpr int FPB = 1;
};
template <>
struct BatchGeo<2048> {
    static constexpr int E = AA_FFT_E_BIG;
    static constexpr int FPB = 1;
};
#ifndef AA_FFT_PACK_SMALL
#define AA_FFT_PACK_SMALL 1
#endif
#if AA_FFT_PACK_SMALL
template <>
struct BatchGeo<256> {
    static constexpr int E = 8;      // NT = 16: two frames per warp
    static constexpr int FPB = 8;    // 128 threads
};
template <>
struct BatchGeo<512> {
    static constexpr int E = 8;      // NT = 32: one frame per warp
    static constexpr int FPB = 4;    // 128 threads
};
#endif
template <int N>
constexpr int fft_min_blocks()
{
#ifndef AA_FFT_THREADS_PER_SM
#define AA_FFT_THREADS_PER_SM 1024
#endif
    constexpr int nt = N / 2 / BatchGeo<N>::E * BatchGeo<N>::FPB;
    constexpr int per_sm = BatchGeo<N>::E >= 16 ? AA_FFT_THREADS_PER_SM / 2 : AA_FFT_THREADS_PER_SM;   // 128 / 64 registers
    return nt >= 64 ? per_sm / nt : 1;
}
// barrier among the NT threads of one frame group: the whole CTA, a named barrier per group, or the warp
template <int NT, int FPB>
__device__ __forceinline__ void group_sync(int grp)
{
    if constexpr (FPB == 1) {
        __syncthreads();
    } else if constexpr (NT > 32) {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(NT) : "memory");
    } else {
        __syncwarp();
    }
}

// Forward transform.  Optionally the next frame's samples are fetched while the current frame is transformed:
// one thread issues a TMA bulk copy (cp.async.bulk + mbarrier) into a staging buffer as soon as every thread has
// pulled the current frame into registers (right after the first barrier of the FFT), so the HBM read latency
// is hidden behind the arithmetic instead of being paid once per frame and CTA.  Measured per length (fraction
// of the HBM peak, plain / staged): n = 4096 0.76 / 0.72, n = 2048 0.70 / 0.74, n = 1024 0.95 / 0.91 -- the
// staging pass costs shared-memory bandwidth, so it is kept only where it wins.
template <int N>
struct FwdLayout {
    static constexpr int N2 = N / 2;
    static constexpr int E = BatchGeo<N>::E;
    static constexpr int FPB = BatchGeo<N>::FPB;
    static constexpr int NT = N2 / E;
#ifndef AA_FFT_STAGE_MIN
#define AA_FFT_STAGE_MIN 8192      // smallest / largest length whose next frame is staged through shared memory by TMA:
#endif                             // off by default -- the staging read costs shared-memory bandwidth, which is what
#ifndef AA_FFT_STAGE_MAX           // binds these kernels next to HBM (it won 4 % at n = 2048 only with the E = 8 plan)
#define AA_FFT_STAGE_MAX 0
#endif
    static constexpr bool STAGE = FPB == 1 && N >= AA_FFT_STAGE_MIN && N <= AA_FFT_STAGE_MAX;
    static constexpr int EXLEN = (padded_len(N2) + 1) & ~1;
    static constexpr size_t ex_bytes = sizeof(float2) * EXLEN;
    static constexpr size_t stage_off = 2 * ex_bytes * FPB;                 // float[N] (16-byte aligned: EXLEN is even)
    static constexpr size_t total = stage_off + (STAGE ? sizeof(float) * N : 0);
};

template <int N>
__global__ void __launch_bounds__(FwdLayout<N>::NT * FwdLayout<N>::FPB, fft_min_blocks<N>())
    fft_forward_kernel(const float *__restrict__ in, int64_t batch, float *__restrict__ out, Tables tab)
{
    using FL = FwdLayout<N>;
    constexpr int N2 = N / 2, E = FL::E, NT = FL::NT, FPB = FL::FPB, EH = E / 2, HALF = N2 + 1, CBIN = N2 / 2;
    constexpr bool STAGE = FL::STAGE;
    extern __shared__ __align__(16) unsigned char fsm[];
    const int grp = threadIdx.x / NT;          // frame group inside the CTA
    const int t = threadIdx.x % NT;
    float2 *exA = reinterpret_cast<float2 *>(fsm + (size_t)grp * 2 * FL::ex_bytes);
    float2 *exB = reinterpret_cast<float2 *>(fsm + (size_t)grp * 2 * FL::ex_bytes + FL::ex_bytes);
    float *stage = reinterpret_cast<float *>(fsm + FL::stage_off);
    __shared__ __align__(8) uint64_t bar;
    uint32_t phase = 0;
    if (STAGE) {
        if (t == 0) {
            mbar_init(&bar, 1);
            fence_proxy_async();
        }
        __syncthreads();
        if (t == 0 && (int64_t)blockIdx.x < batch) {
            mbar_expect_tx(&bar, N * 4);
            bulk_g2s(stage, in + (int64_t)blockIdx.x * N, N * 4, &bar);
        }
    }

#ifndef AA_FFT_REGPF_MAX
#define AA_FFT_REGPF_MAX 0         // lengths up to this one keep the NEXT frame's loads in flight in registers
#endif
    constexpr bool REGPF = !STAGE && FPB == 1 && N <= AA_FFT_REGPF_MAX;
    float2 nx[E];
    if (REGPF && (int64_t)blockIdx.x < batch) {
        const float2 *src = reinterpret_cast<const float2 *>(in + (int64_t)blockIdx.x * N);
#pragma unroll
        for (int m = 0; m < E; ++m) nx[m] = __ldg(&src[t + m * NT]);
    }
    // every group of a CTA runs the same number of iterations (a warp may hold two groups); a group past the end
    // of the batch transforms zeros and stores nothing
    for (int64_t fr0 = (int64_t)blockIdx.x * FPB; fr0 < batch; fr0 += (int64_t)gridDim.x * FPB) {
        const int64_t fr = fr0 + grp;
        const bool live = fr < batch;
        float2 v[E];
        if (STAGE) {
            mbar_wait(&bar, phase);
            phase ^= 1u;
            const float2 *src = reinterpret_cast<const float2 *>(stage);
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = src[t + m * NT];
        } else if (REGPF) {
            // the loads of the next frame are issued before this frame is transformed: twice the bytes in flight
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = nx[m];
            const int64_t nxt = fr + gridDim.x;
            if (nxt < batch) {
                const float2 *src = reinterpret_cast<const float2 *>(in + nxt * N);
#pragma unroll
                for (int m = 0; m < E; ++m) nx[m] = __ldg(&src[t + m * NT]);
            }
        } else {
            const float2 *src = reinterpret_cast<const float2 *>(in + (live ? fr : 0) * N);
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = live ? __ldg(&src[t + m * NT]) : make_float2(0.f, 0.f);
        }
        fft_run<N2, E, 1, 0>(
            v, t, exA, exB, tab.tw, [&] { group_sync<NT, FPB>(grp); },
            [&] {
                // every thread holds its samples in registers: the staging buffer can take the next frame
                const int64_t nxt = fr + gridDim.x;
                if (STAGE && t == 0 && nxt < batch) {
                    mbar_expect_tx(&bar, N * 4);
                    bulk_g2s(stage, in + nxt * N, N * 4, &bar);
                }
            });

        float2 *pbuf = ((fft_num_passes(N2, E) - 1) & 1) ? exB : exA;
#pragma unroll
        for (int m = EH; m < E; ++m) pbuf[padidx(t + m * NT - CBIN)] = v[m];
        if (t == 0) pbuf[padidx(CBIN)] = v[0];
        group_sync<NT, FPB>(grp);

        float2 *dst = reinterpret_cast<float2 *>(out + (live ? fr : 0) * 2 * (int64_t)HALF);
#pragma unroll
        for (int m = 0; m < EH; ++m) {
            const int k = t + m * NT;
            const float2 b = pbuf[padidx(CBIN - k)];
            float2 lo, hi;
            rfft_postpass(v[m], b, __ldg(&tab.pt[k]), lo, hi);
            if (k == 0) { lo.y = 0.0f; hi.y = 0.0f; }   // DC / Nyquist are purely real
            if (live) {
                dst[k] = lo;
                dst[N2 - k] = hi;
            }
        }
        if (t == 0 && live) dst[CBIN] = make_float2(v[EH].x, -v[EH].y);
        group_sync<NT, FPB>(grp);   // pbuf / exchange buffers are reused by the next frame
    }
}

// Inverse (complex -> real), realfft ComplexToReal semantics: unnormalised, i.e.
// inverse(forward(x)) == n * x; the imaginary parts of bins 0 and n/2 are ignored.
// Computed as conj(FFT(conj(Zin))) with Zin rebuilt from the half spectrum.
template <int N>
__global__ void __launch_bounds__(N / 2 / Geo<N>::E, fft_min_blocks<N>()) fft_inverse_kernel(const float *__restrict__ spec,
                                                                        int64_t batch,
                                                                        float *__restrict__ out, Tables tab)
{
    constexpr int N2 = N / 2, E = Geo<N>::E, NT = N2 / E, HALF = N2 + 1;
    constexpr int EXLEN = (padded_len(N2) + 1) & ~1;
    __shared__ __align__(16) float2 exA[EXLEN];
    __shared__ __align__(16) float2 exB[EXLEN];
    const int t = threadIdx.x;

    for (int64_t fr = blockIdx.x; fr < batch; fr += gridDim.x) {
        const float2 *X = reinterpret_cast<const float2 *>(spec + fr * 2 * (int64_t)HALF);
        float2 v[E];
#pragma unroll
        for (int m = 0; m < E; ++m) {
            // z[k] = (X[k] + conj(X[N2-k])) + i * W^{-k} * (X[k] - conj(X[N2-k])),  W = exp(-2 pi i/N)
            const int k = t + m * NT;
            float2 a = __ldg(&X[k]);
            float2 b = __ldg(&X[N2 - k]);
            if (k == 0) { a.y = 0.0f; b.y = 0.0f; }
            const float2 ev = make_float2(a.x + b.x, a.y - b.y);
            const float2 od = make_float2(a.x - b.x, a.y + b.y);
            // exp(+2 pi i k / N): from the post-pass table (0.5*exp(-2 pi i k/N)) for k < N/4, by symmetry above
            float2 w;
            if (k < N2 / 2) {
                const float2 h = __ldg(&tab.pt[k]);
                w = make_float2(2.0f * h.x, -2.0f * h.y);
            } else if (k == N2 / 2) {
                w = make_float2(0.0f, 1.0f);
            } else {
                const float2 h = __ldg(&tab.pt[N2 - k]);
                w = make_float2(-2.0f * h.x, -2.0f * h.y);
            }
            const float2 iw = make_float2(-w.y, w.x);   // i * w
            const float2 z = cadd(ev, cmul(iw, od));
            v[m] = make_float2(z.x, -z.y);              // conj for the forward-FFT trick
        }
        fft_half_complex<N>(v, t, exA, exB, tab.tw);
        float2 *dst = reinterpret_cast<float2 *>(out + fr * N);
#pragma unroll
        for (int m = 0; m < E; ++m) dst[t + m * NT] = make_float2(v[m].x, -v[m].y);
        __syncthreads();
    }
}

template <int N>
static cudaError_t launch_fwd(const Tables &tab, const float *in, int64_t batch, float *out, int num_sms,
                              cudaStream_t s)
{
    using FL = FwdLayout<N>;
    constexpr int NTHREADS = FL::NT * FL::FPB;
    int per_sm = 2048 / NTHREADS;
    if (per_sm > 16) per_sm = 16;
    int64_t grid = (int64_t)num_sms * per_sm;
    const int64_t need = (batch + FL::FPB - 1) / FL::FPB;
    if (grid > need) grid = need;
    static std::atomic<unsigned long long> configured{0ull};     // the opt-in shared-memory size is a per-device attribute
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (FL::total > 48 * 1024 && (dev >= 64 || !((configured.load(std::memory_order_acquire) >> dev) & 1ull))) {
        e = cudaFuncSetAttribute(fft_forward_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FL::total);
        if (e != cudaSuccess) return e;
        if (dev < 64) configured.fetch_or(1ull << dev, std::memory_order_release);
    }
    fft_forward_kernel<N><<<(unsigned)grid, NTHREADS, FL::total, s>>>(in, batch, out, tab);
    return cudaGetLastError();
}

template <int N>
static cudaError_t launch_inv(const Tables &tab, const float *spec, int64_t batch, float *out, int num_sms,
                              cudaStream_t s)
{
    constexpr int NT = N / 2 / Geo<N>::E;
    int per_sm = 2048 / NT;
    if (per_sm > 16) per_sm = 16;
    int64_t grid = (int64_t)num_sms * per_sm;
    if (grid > batch) grid = batch;
    fft_inverse_kernel<N><<<(unsigned)grid, NT, 0, s>>>(spec, batch, out, tab);
    return cudaGetLastError();
}

cudaError_t launch_fft_forward(int n, const Tables &tab, const float *in, int64_t batch, float *out,
                               int num_sms, cudaStream_t s)
{
    if (batch <= 0) return cudaSuccess;
    switch (n) {
        case 4096: return launch_fwd<4096>(tab, in, batch, out, num_sms, s);
        case 2048: return launch_fwd<2048>(tab, in, batch, out, num_sms, s);
        case 1024: return launch_fwd<1024>(tab, in, batch, out, num_sms, s);
        case 512: return launch_fwd<512>(tab, in, batch, out, num_sms, s);
        case 256: return launch_fwd<256>(tab, in, batch, out, num_sms, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_fft_inverse(int n, const Tables &tab, const float *spec, int64_t batch, float *out,
                               int num_sms, cudaStream_t s)
{
    if (batch <= 0) return cudaSuccess;
    switch (n) {
        case 4096: return launch_inv<4096>(tab, spec, batch, out, num_sms, s);
        case 2048: return launch_inv<2048>(tab, spec, batch, out, num_sms, s);
        case 1024: return launch_inv<1024>(tab, spec, batch, out, num_sms, s);
        case 512: return launch_inv<512>(tab, spec, batch, out, num_sms, s);
        case 256: return launch_inv<256>(tab, spec, batch, out, num_sms, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace aa

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
template <int N>
constexpr int fft_min_blocks()
{
    constexpr int nt = N / 2 / Geo<N>::E;
    return nt >= 64 ? 1024 / nt : 1;
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
    static constexpr int NT = N2 / Geo<N>::E;
#ifndef AA_FFT_STAGE_MIN
#define AA_FFT_STAGE_MIN 2048      // smallest / largest length whose next frame is staged through shared memory by TMA
#endif
#ifndef AA_FFT_STAGE_MAX
#define AA_FFT_STAGE_MAX 2048
#endif
    static constexpr bool STAGE = N >= AA_FFT_STAGE_MIN && N <= AA_FFT_STAGE_MAX;
    static constexpr int EXLEN = (padded_len(N2) + 1) & ~1;
    static constexpr size_t ex_bytes = sizeof(float2) * EXLEN;
    static constexpr size_t stage_off = 2 * ex_bytes;                       // float[N] (16-byte aligned: EXLEN is even)
    static constexpr size_t total = stage_off + (STAGE ? sizeof(float) * N : 0);
};

template <int N>
__global__ void __launch_bounds__(N / 2 / Geo<N>::E, fft_min_blocks<N>()) fft_forward_kernel(const float *__restrict__ in,
                                                                        int64_t batch,
                                                                        float *__restrict__ out, Tables tab)
{
    using FL = FwdLayout<N>;
    constexpr int N2 = N / 2, E = Geo<N>::E, NT = N2 / E, EH = E / 2, HALF = N2 + 1, CBIN = N2 / 2;
    constexpr bool STAGE = FL::STAGE;
    extern __shared__ __align__(16) unsigned char fsm[];
    float2 *exA = reinterpret_cast<float2 *>(fsm);
    float2 *exB = reinterpret_cast<float2 *>(fsm + FL::ex_bytes);
    float *stage = reinterpret_cast<float *>(fsm + FL::stage_off);
    __shared__ __align__(8) uint64_t bar;
    const int t = threadIdx.x;
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
    constexpr bool REGPF = !STAGE && N <= AA_FFT_REGPF_MAX;
    float2 nx[E];
    if (REGPF && (int64_t)blockIdx.x < batch) {
        const float2 *src = reinterpret_cast<const float2 *>(in + (int64_t)blockIdx.x * N);
#pragma unroll
        for (int m = 0; m < E; ++m) nx[m] = __ldg(&src[t + m * NT]);
    }
    for (int64_t fr = blockIdx.x; fr < batch; fr += gridDim.x) {
        float2 v[E];
        if (STAGE) {
            mbar_wait(&bar, phase);
            phase ^= 1u;
            const float2 *src = reinterpret_cast<const float2 *>(stage);
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = src[t + m * NT];
        } else if (REGPF) {
            // one-warp / two-warp CTAs are limited by the 32 CTAs an SM holds, not by registers: the loads of the
            // next frame are issued before this frame is transformed, which doubles the bytes in flight per CTA
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = nx[m];
            const int64_t nxt = fr + gridDim.x;
            if (nxt < batch) {
                const float2 *src = reinterpret_cast<const float2 *>(in + nxt * N);
#pragma unroll
                for (int m = 0; m < E; ++m) nx[m] = __ldg(&src[t + m * NT]);
            }
        } else {
            const float2 *src = reinterpret_cast<const float2 *>(in + fr * N);
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = __ldg(&src[t + m * NT]);
        }
        fft_run<N2, E, 1, 0>(
            v, t, exA, exB, tab.tw, [] { __syncthreads(); },
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
        __syncthreads();

        float2 *dst = reinterpret_cast<float2 *>(out + fr * 2 * (int64_t)HALF);
#pragma unroll
        for (int m = 0; m < EH; ++m) {
            const int k = t + m * NT;
            const float2 b = pbuf[padidx(CBIN - k)];
            float2 lo, hi;
            rfft_postpass(v[m], b, __ldg(&tab.pt[k]), lo, hi);
            if (k == 0) { lo.y = 0.0f; hi.y = 0.0f; }   // DC / Nyquist are purely real
            dst[k] = lo;
            dst[N2 - k] = hi;
        }
        if (t == 0) dst[CBIN] = make_float2(v[EH].x, -v[EH].y);
        __syncthreads();   // pbuf / exchange buffers are reused by the next frame
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
    constexpr int NT = N / 2 / Geo<N>::E;
    int per_sm = 2048 / NT;
    if (per_sm > 16) per_sm = 16;
    int64_t grid = (int64_t)num_sms * per_sm;
    if (grid > batch) grid = batch;
    using FL = FwdLayout<N>;
    static std::atomic<unsigned long long> configured{0ull};     // the opt-in shared-memory size is a per-device attribute
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (FL::total > 48 * 1024 && (dev >= 64 || !((configured.load(std::memory_order_acquire) >> dev) & 1ull))) {
        e = cudaFuncSetAttribute(fft_forward_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FL::total);
        if (e != cudaSuccess) return e;
        if (dev < 64) configured.fetch_or(1ull << dev, std::memory_order_release);
    }
    fft_forward_kernel<N><<<(unsigned)grid, NT, FL::total, s>>>(in, batch, out, tab);
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

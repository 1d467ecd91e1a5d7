// aa_yin.cu -- YIN-style lag search per frame (NEW, no reference code; the
// north_star lists "autocorrelation or YIN-style lag search").  Definition == oracle aao_yin_lag:
//   d(tau)  = sum_{j<W} (x[j] - x[j+tau])^2,  W = n - max_lag        (raw, unwindowed frame)
//   d'(tau) = d(tau) * tau / sum_{j=1..tau} d(j)                      (1 where the sum is 0)
//   lag     = first tau >= min_lag with d'(tau) < threshold, advanced while d'(tau+1) < d'(tau);
//             otherwise the first arg-min of d' over [min_lag, max_lag]
// One CTA per frame.  The frame sits in shared memory; each thread owns groups of four consecutive
// lags and slides a 4-sample register window over x[j+tau..], so a (j, 4 lags) step costs one
// broadcast load, one new load and eight FP instructions.  d(tau) is accumulated in f32 (direct form,
// no cancellation); the oracle uses f64, the lag index is compared exactly except near-ties.
#include "aa_internal.h"

namespace aa {

constexpr int YIN_THREADS = 256;

__global__ void __launch_bounds__(YIN_THREADS) yin_kernel(const float *__restrict__ clips, int64_t n_clips,
                                                          int64_t clip_stride, int64_t T, int n, int hop,
                                                          int min_lag, int max_lag, float threshold,
                                                          int32_t *__restrict__ lag_out,
                                                          float *__restrict__ cmnd_out)
{
    extern __shared__ __align__(16) float ysm[];
    float *x = ysm;                   // [n + 4]
    float *d = ysm + n + 4;           // [max_lag + 1] difference function, then d'
    __shared__ float s_part[YIN_THREADS / 32];
    __shared__ int s_first;
    __shared__ int s_args[YIN_THREADS / 32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int W = n - max_lag;

    for (int64_t fr = blockIdx.x; fr < n_clips * T; fr += gridDim.x) {
        const int64_t clip = fr / T, f = fr % T;
        const float *src = clips + clip * clip_stride + f * hop;
        for (int i = t; i < n; i += YIN_THREADS) x[i] = src[i];
        if (t < 4) x[n + t] = 0.0f;
        if (t == 0) { d[0] = 0.0f; s_first = 0x7fffffff; }
        __syncthreads();

        // ---- difference function, four lags per thread per group ----
        for (int tau0 = 1 + 4 * t; tau0 <= max_lag; tau0 += 4 * YIN_THREADS) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            float w0 = x[tau0], w1 = x[tau0 + 1], w2 = x[tau0 + 2];
            for (int j = 0; j < W; ++j) {
                const float xj = x[j];
                const float w3 = x[j + tau0 + 3];        // j + tau0 + 3 <= W - 1 + max_lag + 3 = n + 2
                float e;
                e = xj - w0; a0 = fmaf(e, e, a0);
                e = xj - w1; a1 = fmaf(e, e, a1);
                e = xj - w2; a2 = fmaf(e, e, a2);
                e = xj - w3; a3 = fmaf(e, e, a3);
                w0 = w1; w1 = w2; w2 = w3;
            }
            d[tau0] = a0;
            if (tau0 + 1 <= max_lag) d[tau0 + 1] = a1;
            if (tau0 + 2 <= max_lag) d[tau0 + 2] = a2;
            if (tau0 + 3 <= max_lag) d[tau0 + 3] = a3;
        }
        __syncthreads();

        // ---- cumulative-mean normalisation: block scan over tau (chunk per thread) ----
        const int chunk = (max_lag + YIN_THREADS) / YIN_THREADS;      // lags 1..max_lag, ceil
        const int lo = 1 + t * chunk;
        const int hi = min(lo + chunk - 1, max_lag);
        float local = 0.f;
        for (int tau = lo; tau <= hi; ++tau) local += d[tau];
        float incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) s_part[warp] = incl;
        __syncthreads();
        float base = 0.f;
        for (int w = 0; w < warp; ++w) base += s_part[w];
        float run = base + incl - local;
        for (int tau = lo; tau <= hi; ++tau) {
            const float dv = d[tau];
            run += dv;
            d[tau] = run > 0.0f ? dv * (float)tau / run : 1.0f;
        }
        __syncthreads();

        // ---- selection ----
        int myfirst = 0x7fffffff;
        float mymin = 3.0e38f;
        int myarg = 0x7fffffff;
        for (int tau = min_lag + t; tau <= max_lag; tau += YIN_THREADS) {
            const float v = d[tau];
            if (v < threshold && tau < myfirst) myfirst = tau;
            if (v < mymin) { mymin = v; myarg = tau; }      // ascending tau per thread: keeps the first minimum
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            myfirst = min(myfirst, __shfl_xor_sync(0xffffffffu, myfirst, o));
            const float ov = __shfl_xor_sync(0xffffffffu, mymin, o);
            const int oa = __shfl_xor_sync(0xffffffffu, myarg, o);
            if (ov < mymin || (ov == mymin && oa < myarg)) { mymin = ov; myarg = oa; }
        }
        if (lane == 0) {
            atomicMin(&s_first, myfirst);
            s_part[warp] = mymin;
            s_args[warp] = myarg;
        }
        __syncthreads();
        if (t == 0) {
            int best;
            if (s_first != 0x7fffffff) {
                int tau = s_first;
                while (tau + 1 <= max_lag && d[tau + 1] < d[tau]) ++tau;
                best = tau;
            } else {
                float bm = s_part[0];
                best = s_args[0];
                for (int w = 1; w < YIN_THREADS / 32; ++w)
                    if (s_part[w] < bm || (s_part[w] == bm && s_args[w] < best)) { bm = s_part[w]; best = s_args[w]; }
            }
            lag_out[fr] = best;
            if (cmnd_out) cmnd_out[fr] = d[best];
        }
        __syncthreads();
    }
}

cudaError_t launch_yin(const float *clips, int64_t n_clips, int64_t clip_stride, int64_t T, int n, int hop,
                       int min_lag, int max_lag, float threshold, int32_t *lag_out, float *cmnd_out,
                       int num_sms, cudaStream_t s)
{
    const int64_t frames = n_clips * T;
    if (frames <= 0) return cudaSuccess;
    const size_t smem = sizeof(float) * (size_t)(n + 4 + max_lag + 1);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(yin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int64_t grid = (int64_t)num_sms * 8;
    if (grid > frames) grid = frames;
    yin_kernel<<<(unsigned)grid, YIN_THREADS, smem, s>>>(clips, n_clips, clip_stride, T, n, hop, min_lag, max_lag,
                                                         threshold, lag_out, cmnd_out);
    return cudaGetLastError();
}

}  // namespace aa

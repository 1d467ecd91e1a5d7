// Microbenchmarks that informed the round-1 kernel design (run on a B200 through gpurun):
//  (1) issue rate of FFMA / FADD / FMUL vs their packed f32x2 forms (FFMA2 / FADD2 / FMUL2), alone and
//      mixed with ALU-pipe integer work;
//  (2) instruction-fetch behaviour: IPC of a straight-line loop body as a function of its size, with all
//      resident CTAs in phase or staggered.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) pipe_kernel(float *out, int iters, float s)
{
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 m = make_float2(s, 1.0f / s), c = make_float2(1e-3f, -1e-3f);
    unsigned q = threadIdx.x, r = blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { a[i].x = __fmaf_rn(a[i].x, m.x, c.x); a[i].y = __fmaf_rn(a[i].y, m.y, c.y); }        // 2 FFMA
                if (MODE == 1) { a[i] = __ffma2_rn(a[i], m, c); }                                                      // 1 FFMA2
                if (MODE == 2) { a[i].x = __fadd_rn(a[i].x, c.x); a[i].y = __fadd_rn(a[i].y, c.y); }                  // 2 FADD
                if (MODE == 3) { a[i] = __fadd2_rn(a[i], c); }                                                         // 1 FADD2
                if (MODE == 4) { a[i].x = __fmul_rn(a[i].x, m.x); a[i].y = __fmul_rn(a[i].y, m.y); }                  // 2 FMUL
                if (MODE == 5) { a[i] = __fmul2_rn(a[i], m); }                                                         // 1 FMUL2
                if (MODE == 6) { a[i].x = __fmaf_rn(a[i].x, m.x, c.x); a[i].y = __fmaf_rn(a[i].y, m.y, c.y);          // 2 FFMA + 2 LOP3/IADD
                                 q = (q ^ r) + 0x9e37u; r = (r & q) + it; }
                if (MODE == 7) { a[i] = __ffma2_rn(a[i], m, c); q = (q ^ r) + 0x9e37u; r = (r & q) + it; }             // 1 FFMA2 + 2 int
                if (MODE == 8) { a[i].x = __fmaf_rn(a[i].x, m.x, c.x); a[i].y = __fadd_rn(a[i].y, c.y); }             // FFMA + FADD
                if (MODE == 9) { q = (q ^ r) + 0x9e37u; r = (r & q) + it; }                                            // int only
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y;
    if (acc == 123.456f || q + r == 0x12345u) out[threadIdx.x] = acc;
}

// straight-line body of KI FFMA-pairs... use FFMA + LOP mixes so no single pipe limits; body size ~ 16*KI*? bytes
template <int KI>
__global__ void __launch_bounds__(320) icache_kernel(float *out, int iters, float s, int stagger)
{
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    unsigned q = threadIdx.x;
    if (stagger) {
        // desynchronise the CTAs of an SM: spin for a CTA-dependent number of cycles
        const long long t0 = clock64();
        const long long wait = (long long)(blockIdx.x % 7) * 1237 + (threadIdx.x >> 5) * 311;
        while (clock64() - t0 < wait) { }
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < KI; ++u) {
            // 3 instructions with distinct immediates so that the body cannot be rolled: FFMA, FADD, LOP3/IADD
            a[u & 7] = __fmaf_rn(a[u & 7], s, (float)(u + 1) * 1e-4f);
            a[(u + 3) & 7] = __fadd_rn(a[(u + 3) & 7], (float)(u + 7) * 1e-5f);
            q = (q ^ (0x1357u + u)) + (q >> 3);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += a[i];
    if (acc == 123.456f || q == 0x12345u) out[threadIdx.x] = acc;
}

template <typename F>
static float time_ms(F f)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f();   // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
}

template <int MODE>
static void run_pipe(const char *name, int instr_per_inner, float *out, int sms, double ghz)
{
    const int iters = 4000, ctas = sms * 8;
    float ms = time_ms([&] { pipe_kernel<MODE><<<ctas, 256>>>(out, iters, 1.0001f); });
    CK(cudaGetLastError());
    const double winstr = (double)ctas * 8 /*warps*/ * iters * 64.0 * instr_per_inner;
    printf("{\"bench\":\"pipe\",\"mode\":\"%s\",\"ms\":%.3f,\"warp_instr_per_clk_per_sm\":%.3f}\n", name, ms,
           winstr / (ms * 1e-3 * ghz * 1e9 * sms));
}

template <int KI>
static void run_icache(float *out, int sms, double ghz, int ctas_per_sm)
{
    for (int stagger = 0; stagger < 2; ++stagger) {
        const int iters = 400000 / KI + 1, ctas = sms * ctas_per_sm;
        float ms = time_ms([&] { icache_kernel<KI><<<ctas, 320>>>(out, iters, 1.0001f, stagger); });
        CK(cudaGetLastError());
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, icache_kernel<KI>));
        const double winstr = (double)ctas * 10 * (double)iters * KI * 4.0;   // 4 SASS instr per body step (FFMA, FADD, LOP3, LEA.HI)
        printf("{\"bench\":\"icache\",\"body_steps\":%d,\"approx_body_kb\":%.1f,\"ctas_per_sm\":%d,\"stagger\":%d,\"ms\":%.3f,"
               "\"approx_ipc_per_sm\":%.3f,\"regs\":%d}\n", KI, KI * 4 * 16 / 1024.0, ctas_per_sm, stagger, ms,
               winstr / (ms * 1e-3 * ghz * 1e9 * sms), fa.numRegs);
    }
}

int main()
{
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double ghz = khz * 1e-6;
    const int sms = pr.multiProcessorCount;
    printf("{\"gpu\":\"%s\",\"sms\":%d,\"clock_ghz\":%.3f}\n", pr.name, sms, ghz);
    float *out; CK(cudaMalloc(&out, 4096));
    run_pipe<0>("2xFFMA", 2, out, sms, ghz);
    run_pipe<1>("1xFFMA2", 1, out, sms, ghz);
    run_pipe<2>("2xFADD", 2, out, sms, ghz);
    run_pipe<3>("1xFADD2", 1, out, sms, ghz);
    run_pipe<4>("2xFMUL", 2, out, sms, ghz);
    run_pipe<5>("1xFMUL2", 1, out, sms, ghz);
    run_pipe<6>("2xFFMA+4int", 6, out, sms, ghz);
    run_pipe<7>("1xFFMA2+4int", 5, out, sms, ghz);
    run_pipe<8>("FFMA+FADD", 2, out, sms, ghz);
    run_pipe<9>("4int", 4, out, sms, ghz);
    for (int c = 1; c <= 3; c += 2) {
        run_icache<64>(out, sms, ghz, c);
        run_icache<256>(out, sms, ghz, c);
        run_icache<384>(out, sms, ghz, c);
        run_icache<512>(out, sms, ghz, c);
        run_icache<768>(out, sms, ghz, c);
        run_icache<1024>(out, sms, ghz, c);
        run_icache<1536>(out, sms, ghz, c);
        run_icache<2048>(out, sms, ghz, c);
    }
    return 0;
}

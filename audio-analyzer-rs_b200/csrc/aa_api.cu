// aa_api.cu -- the C ABI of libaa_gpu.so (include/aa_gpu.h): handles, tables, host
// pipelines (H2D / kernel / D2H over clip groups) and the streaming ring.
//
// There is deliberately no CPU fallback anywhere in this file: if no sm_100 device is
// usable every create call fails with AA_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sched.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <atomic>
#include <chrono>
#include "aa_internal.h"

using namespace aa;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local int g_device = 0;   // aa_set_device is per calling thread (distinct handles on distinct threads)

static aa_status fail(aa_status code, const std::string &msg)
{
    g_err = msg;
    return code;
}
static aa_status fail_cuda(cudaError_t e, const char *what)
{
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return AA_ERR_CUDA;
}
#define CU(call)                                           \
    do {                                                   \
        cudaError_t _e = (call);                           \
        if (_e != cudaSuccess) return fail_cuda(_e, #call); \
    } while (0)

extern "C" AA_API const char *aa_last_error(void) { return g_err.c_str(); }
extern "C" AA_API int32_t aa_version(void) { return 100; }

static aa_status check_device(int *num_sms)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return fail(AA_ERR_NO_DEVICE, "no CUDA device visible (libaa_gpu has no CPU fallback)");
    }
    if (g_device >= count) return fail(AA_ERR_NO_DEVICE, "selected device index out of range");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, g_device));
    if (prop.major != 10 || prop.minor != 0)   // sm_100a cubins only: no PTX, nothing runs on 10.3
        return fail(AA_ERR_NO_DEVICE, std::string("device '") + prop.name +
                                          "' is not sm_100; libaa_gpu is built for sm_100a only");
    CU(cudaSetDevice(g_device));
    if (num_sms) *num_sms = prop.multiProcessorCount;
    return AA_OK;
}

extern "C" AA_API aa_status aa_device_count(int32_t *count)
{
    if (!count) return fail(AA_ERR_INVALID, "count is null");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        cudaGetLastError();
        c = 0;
    }
    *count = c;
    return AA_OK;
}

extern "C" AA_API aa_status aa_set_device(int32_t device)
{
    if (device < 0) return fail(AA_ERR_INVALID, "negative device index");
    g_device = device;
    return AA_OK;
}

// CPUs that NVML reports as local to CUDA device `device` (its NUMA node); false if unknown.  NVML is looked up
// at run time (dlopen), so the library has no link-time dependency on it.
static bool gpu_local_cpus(int device, cpu_set_t *set)
{
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), device) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    void *lib = dlopen("libnvidia-ml.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!lib) return false;
    typedef int (*init_t)(void);
    typedef int (*byid_t)(const char *, void **);
    typedef int (*aff_t)(void *, unsigned, unsigned long *);
    init_t init = (init_t)dlsym(lib, "nvmlInit_v2");
    byid_t byid = (byid_t)dlsym(lib, "nvmlDeviceGetHandleByPciBusId_v2");
    aff_t aff = (aff_t)dlsym(lib, "nvmlDeviceGetCpuAffinity");
    bool ok = false;
    void *dev = nullptr;
    unsigned long words[16] = {0};
    if (init && byid && aff && init() == 0 && byid(bus, &dev) == 0 && aff(dev, 16, words) == 0) {
        CPU_ZERO(set);
        for (int w = 0; w < 16; ++w)
            for (int b = 0; b < 64; ++b)
                if ((words[w] >> b) & 1ul) {
                    CPU_SET(w * 64 + b, set);
                    ok = true;
                }
    }
    return ok;   // (NVML stays initialised and loaded for the life of the process)
}

// Pinned host memory for the host-buffer calls.  The allocation runs with the calling thread restricted to the CPUs
// that are local to the selected GPU, so the first-touch policy puts the pinned pages on that GPU's NUMA node: on a
// two-socket box with eight GPUs, copies from the remote socket cross the inter-socket link and the aggregate
// host-to-device rate of eight unbound processes collapses (tools/h2d_probe.py).  AA_NO_NUMA_BIND=1 in
// the environment switches the binding off.
extern "C" AA_API aa_status aa_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(AA_ERR_INVALID, "out is null");
    CU(cudaSetDevice(g_device));
    cpu_set_t old_set, gpu_set, use;
    bool bound = false;
    if (!getenv("AA_NO_NUMA_BIND") && sched_getaffinity(0, sizeof(old_set), &old_set) == 0 &&
        gpu_local_cpus(g_device, &gpu_set)) {
        CPU_AND(&use, &old_set, &gpu_set);
        if (CPU_COUNT(&use) > 0 && sched_setaffinity(0, sizeof(use), &use) == 0) bound = true;
    }
    const cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (bound) sched_setaffinity(0, sizeof(old_set), &old_set);
    if (e != cudaSuccess) return fail_cuda(e, "cudaHostAlloc");
    return AA_OK;
}
extern "C" AA_API aa_status aa_host_free(void *p)
{
    if (p) CU(cudaFreeHost(p));
    return AA_OK;
}
extern "C" AA_API aa_status aa_device_alloc(size_t bytes, void **out)
{
    if (!out) return fail(AA_ERR_INVALID, "out is null");
    CU(cudaSetDevice(g_device));
    CU(cudaMalloc(out, bytes ? bytes : 1));
    return AA_OK;
}
extern "C" AA_API aa_status aa_device_free(void *p)
{
    if (p) CU(cudaFree(p));
    return AA_OK;
}
extern "C" AA_API aa_status aa_memcpy_h2d(void *dst, const void *src, size_t bytes)
{
    CU(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return AA_OK;
}
extern "C" AA_API aa_status aa_memcpy_d2h(void *dst, const void *src, size_t bytes)
{
    CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return AA_OK;
}
extern "C" AA_API aa_status aa_device_synchronize(void)
{
    CU(cudaDeviceSynchronize());
    return AA_OK;
}

// ---------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------
static bool valid_n(int n) { return n == 256 || n == 512 || n == 1024 || n == 2048 || n == 4096; }

struct DeviceTables {
    void *mem = nullptr;
    Tables tab{};
};

static aa_status build_tables(int n, DeviceTables *dt)
{
    const int n2 = n / 2, n4 = n / 4, half = n2 + 1;
    const size_t nf2 = (size_t)n2 + n4 + n2;   // float2 entries: tw, pt, win2
    const size_t bytes = nf2 * sizeof(float2) + (size_t)((half + 3) & ~3) * sizeof(float);
    std::vector<unsigned char> host(bytes);
    float2 *tw = reinterpret_cast<float2 *>(host.data());
    float2 *pt = tw + n2;
    float2 *win2 = pt + n4;
    float *flux_w = reinterpret_cast<float *>(win2 + n2);
    const double PI = 3.14159265358979323846;
    for (int k = 0; k < n2; ++k) {
        const double a = -2.0 * PI * (double)k / (double)n2;
        tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    for (int k = 0; k < n4; ++k) {
        // realfft: twiddle computed in f64, rounded to f32, then halved
        const double a = -2.0 * PI * (double)k / (double)n;
        pt[k] = make_float2((float)std::cos(a) * 0.5f, (float)std::sin(a) * 0.5f);
    }
    {
        // periodic Hann exactly as stft.rs:641-648 (f32 throughout, libm cosf)
        const float pi_f32 = 3.14159274101257324219f;
        std::vector<float> w((size_t)n);
        for (int i = 0; i < n; ++i) {
            const float x = (float)i / (float)n;
            w[(size_t)i] = 0.5f - 0.5f * cosf(2.0f * pi_f32 * x);
        }
        for (int m = 0; m < n2; ++m) win2[m] = make_float2(w[2 * (size_t)m], w[2 * (size_t)m + 1]);
    }
    for (int k = 0; k < half; ++k) flux_w[k] = 1.0f - ((float)k / (float)half);   // onset.rs:280

    CU(cudaMalloc(&dt->mem, bytes));
    CU(cudaMemcpy(dt->mem, host.data(), bytes, cudaMemcpyHostToDevice));
    float2 *d = reinterpret_cast<float2 *>(dt->mem);
    dt->tab.tw = d;
    dt->tab.pt = d + n2;
    dt->tab.win2 = d + n2 + n4;
    dt->tab.flux_w = reinterpret_cast<const float *>(d + n2 + n4 + n2);
    return AA_OK;
}

// ---------------------------------------------------------------------------
// FftProcessor
// ---------------------------------------------------------------------------
struct aa_fft {
    int n = 0, num_sms = 0, device = 0;
    cudaStream_t stream = nullptr;
    DeviceTables dt;
    float *d_in = nullptr, *d_out = nullptr;
    int64_t cap = 0;
};

extern "C" AA_API aa_status aa_fft_create(int32_t n, aa_fft **out)
{
    if (!out) return fail(AA_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!valid_n(n)) return fail(AA_ERR_UNSUPPORTED, "FFT length must be one of 256, 512, 1024, 2048, 4096");
    int sms = 0;
    aa_status st = check_device(&sms);
    if (st != AA_OK) return st;
    aa_fft *h = new aa_fft();
    h->n = n;
    h->num_sms = sms;
    h->device = g_device;
    st = build_tables(n, &h->dt);
    if (st != AA_OK) { delete h; return st; }
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { cudaFree(h->dt.mem); delete h; return fail_cuda(e, "cudaStreamCreate"); }
    *out = h;
    return AA_OK;
}

extern "C" AA_API aa_status aa_fft_destroy(aa_fft *h)
{
    if (!h) return AA_OK;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    cudaFree(h->d_in);
    cudaFree(h->d_out);
    cudaFree(h->dt.mem);
    delete h;
    return AA_OK;
}

extern "C" AA_API int32_t aa_fft_len(const aa_fft *h) { return h ? h->n : 0; }

static aa_status fft_reserve(aa_fft *h, int64_t batch)
{
    if (batch <= h->cap) return AA_OK;
    cudaFree(h->d_in);
    cudaFree(h->d_out);
    h->d_in = h->d_out = nullptr;
    h->cap = 0;
    const size_t spec = 2 * (size_t)(h->n / 2 + 1);
    const size_t per = std::max((size_t)h->n, spec);
    CU(cudaMalloc(&h->d_in, sizeof(float) * per * (size_t)batch));
    CU(cudaMalloc(&h->d_out, sizeof(float) * per * (size_t)batch));
    h->cap = batch;
    return AA_OK;
}

extern "C" AA_API aa_status aa_fft_forward_device(aa_fft *h, const float *in_dev, int64_t batch,
                                                  float *out_dev, void *stream)
{
    if (!h || !in_dev || !out_dev || batch < 0) return fail(AA_ERR_INVALID, "aa_fft_forward_device: bad argument");
    if (((uintptr_t)in_dev & 15u) || ((uintptr_t)out_dev & 7u))
        return fail(AA_ERR_INVALID, "aa_fft_forward_device: input must be 16-byte aligned (TMA bulk copy), output 8-byte");
    CU(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;   // NULL = the CUDA default stream
    CU(launch_fft_forward(h->n, h->dt.tab, in_dev, batch, out_dev, h->num_sms, s));
    return AA_OK;
}

extern "C" AA_API aa_status aa_fft_forward(aa_fft *h, const float *in_host, int64_t batch, float *out_host)
{
    if (!h || !in_host || !out_host || batch < 0) return fail(AA_ERR_INVALID, "aa_fft_forward: bad argument");
    if (batch == 0) return AA_OK;
    CU(cudaSetDevice(h->device));
    aa_status st = fft_reserve(h, batch);
    if (st != AA_OK) return st;
    const size_t spec = 2 * (size_t)(h->n / 2 + 1);
    CU(cudaMemcpyAsync(h->d_in, in_host, sizeof(float) * (size_t)h->n * (size_t)batch, cudaMemcpyHostToDevice,
                       h->stream));
    CU(launch_fft_forward(h->n, h->dt.tab, h->d_in, batch, h->d_out, h->num_sms, h->stream));
    CU(cudaMemcpyAsync(out_host, h->d_out, sizeof(float) * spec * (size_t)batch, cudaMemcpyDeviceToHost,
                       h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return AA_OK;
}

extern "C" AA_API aa_status aa_fft_inverse_device(aa_fft *h, const float *spec_dev, int64_t batch,
                                                  float *out_dev, void *stream)
{
    if (!h || !spec_dev || !out_dev || batch < 0) return fail(AA_ERR_INVALID, "aa_fft_inverse_device: bad argument");
    if (((uintptr_t)spec_dev & 7u) || ((uintptr_t)out_dev & 7u))
        return fail(AA_ERR_INVALID, "aa_fft_inverse_device: buffers must be 8-byte aligned");
    CU(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;   // NULL = the CUDA default stream
    CU(launch_fft_inverse(h->n, h->dt.tab, spec_dev, batch, out_dev, h->num_sms, s));
    return AA_OK;
}

extern "C" AA_API aa_status aa_fft_inverse(aa_fft *h, const float *spec_host, int64_t batch, float *out_host)
{
    if (!h || !spec_host || !out_host || batch < 0) return fail(AA_ERR_INVALID, "aa_fft_inverse: bad argument");
    if (batch == 0) return AA_OK;
    CU(cudaSetDevice(h->device));
    aa_status st = fft_reserve(h, batch);
    if (st != AA_OK) return st;
    const size_t spec = 2 * (size_t)(h->n / 2 + 1);
    CU(cudaMemcpyAsync(h->d_in, spec_host, sizeof(float) * spec * (size_t)batch, cudaMemcpyHostToDevice, h->stream));
    CU(launch_fft_inverse(h->n, h->dt.tab, h->d_in, batch, h->d_out, h->num_sms, h->stream));
    CU(cudaMemcpyAsync(out_host, h->d_out, sizeof(float) * (size_t)h->n * (size_t)batch, cudaMemcpyDeviceToHost,
                       h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return AA_OK;
}

// ---------------------------------------------------------------------------
// analyzer
// ---------------------------------------------------------------------------
extern "C" AA_API void aa_config_default_pitch(aa_config *cfg, float sample_rate)
{
    if (!cfg) return;
    cfg->n = 2048;               // stft.rs:170
    cfg->hop = 512;              // stft.rs:169
    cfg->sample_rate = sample_rate;
    cfg->min_freq = 24.0f;       // stft.rs:173
    cfg->max_freq = 10000.0f;    // stft.rs:174
    cfg->noise_floor_db = -96.0f;  // dynamics.rs:100
    cfg->features = AA_FEAT_PITCH | AA_FEAT_TRACKER;
}

extern "C" AA_API void aa_config_default_onset(aa_config *cfg, float sample_rate)
{
    if (!cfg) return;
    cfg->n = 256;                // onset.rs:122
    cfg->hop = 64;               // onset.rs:123
    cfg->sample_rate = sample_rate;
    cfg->min_freq = 24.0f;
    cfg->max_freq = 10000.0f;
    cfg->noise_floor_db = -96.0f;
    cfg->features = AA_FEAT_ONSET;
}

static aa_status validate_config(const aa_config *cfg)
{
    if (!cfg) return fail(AA_ERR_INVALID, "config is null");
    if (!valid_n(cfg->n)) return fail(AA_ERR_UNSUPPORTED, "window size must be one of 256, 512, 1024, 2048, 4096");
    if (cfg->hop != cfg->n / 4)
        return fail(AA_ERR_UNSUPPORTED, "hop must be n/4 (the reference geometry, stft.rs:169 / onset.rs:123)");
    if (!(cfg->sample_rate > 0.0f)) return fail(AA_ERR_INVALID, "sample_rate must be positive");
    if (cfg->features & ~AA_FEAT_ALL) return fail(AA_ERR_INVALID, "unknown feature bits");
    if ((cfg->features & AA_FEAT_TRACKER) && !(cfg->features & AA_FEAT_PITCH))
        return fail(AA_ERR_INVALID, "AA_FEAT_TRACKER requires AA_FEAT_PITCH");
    return AA_OK;
}

extern "C" AA_API int64_t aa_num_frames(const aa_config *cfg, int64_t clip_len)
{
    if (!cfg || cfg->n <= 0 || cfg->hop <= 0 || clip_len < cfg->n) return 0;
    return (clip_len - cfg->n) / cfg->hop + 1;
}

static void fill_params(const aa_config &cfg, const Tables &tab, AnalyzeParams *p)
{
    const int half = cfg.n / 2 + 1;
    p->tab = tab;
    p->n = cfg.n;
    p->hop = cfg.hop;
    p->half = half;
    p->bin_width = cfg.sample_rate / (float)cfg.n;                          // stft.rs:320
    p->min_freq = cfg.min_freq;
    p->max_freq = cfg.max_freq;
    p->global_floor = powf(10.0f, cfg.noise_floor_db / 20.0f) * (float)half / 2.0f;  // stft.rs:323-324
    // stft.rs:454-455 (`as usize` saturates negatives to 0)
    float mb = ceilf(cfg.min_freq / p->bin_width);
    long long minb = mb > 0.0f ? (mb < 1e9f ? (long long)mb : 1000000000LL) : 0;
    if (minb < 1) minb = 1;
    float xb = floorf(cfg.max_freq / p->bin_width);
    long long maxb = xb > 0.0f ? (xb < 1e9f ? (long long)xb : 1000000000LL) : 0;
    const long long hs2 = half >= 2 ? half - 2 : 0;
    if (maxb > hs2) maxb = hs2;
    p->min_bin = (int)minb;
    p->max_bin = (int)maxb;
    p->features_mask = cfg.features;
}

struct StageSlot {
    float *in = nullptr;
    size_t in_cap = 0;
    uint8_t *raw = nullptr;      // interleaved PCM as it came from the host (aa_analyze_host_pcm)
    size_t raw_cap = 0;
    float *mags = nullptr;
    size_t mags_cap = 0;
    aa_frame_features *feat = nullptr;
    aa_stable_pitches *stable = nullptr;
    aa_clip_summary *summ = nullptr;
    uint8_t *onset = nullptr;
    float *dbg_floor = nullptr;
    uint8_t *dbg_peaks = nullptr;
    size_t frames_cap = 0, clips_cap = 0, dbgf_cap = 0, dbgp_cap = 0, onset_cap = 0;
    cudaEvent_t h2d_done = nullptr, k_done = nullptr, d2h_done = nullptr;
};

struct aa_analyzer {
    aa_config cfg{};
    int device = 0, num_sms = 0;
    DeviceTables dt;
    cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    StageSlot slot[2];
    unsigned char *scratch = nullptr;   // per-CTA overflow scratch of the analysis kernel
    unsigned long long *work_counter = nullptr;   // work queue of the persistent CTAs: [16 B counter][n_clips x 2 u32 flags]
    size_t flag_clips_cap = 0;                    // clips the segment flags behind the counter have room for
    float *seg_state = nullptr;                   // [n_clips][state_floats]: state hand-off between time segments
    size_t seg_state_cap = 0;                     // in floats
    int max_grid = 0;
    int64_t launches = 0;
    // time-sliced host pipeline (analyze_host_sliced): whole-batch device arrays, raw PCM staging, carried state
    struct Sliced {
        float *in = nullptr, *mags = nullptr, *state = nullptr;
        aa_frame_features *feat = nullptr;
        aa_stable_pitches *stab = nullptr;
        aa_clip_summary *summ = nullptr;
        uint8_t *onset = nullptr, *raw[2] = {nullptr, nullptr};
        size_t in_cap = 0, mags_cap = 0, state_cap = 0, feat_cap = 0, stab_cap = 0, summ_cap = 0, onset_cap = 0,
               raw_cap[2] = {0, 0};
        cudaEvent_t h2d_done[16] = {}, k_done[16] = {};
        bool events = false;
    } sl;
};

extern "C" AA_API aa_status aa_analyzer_create(const aa_config *cfg, aa_analyzer **out)
{
    if (!out) return fail(AA_ERR_INVALID, "out is null");
    *out = nullptr;
    aa_status st = validate_config(cfg);
    if (st != AA_OK) return st;
    int sms = 0;
    st = check_device(&sms);
    if (st != AA_OK) return st;
    aa_analyzer *h = new aa_analyzer();
    h->cfg = *cfg;
    h->device = g_device;
    h->num_sms = sms;
    st = build_tables(cfg->n, &h->dt);
    if (st != AA_OK) { delete h; return st; }
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking)) != cudaSuccess) {
        aa_analyzer_destroy(h);
        return fail_cuda(e, "cudaStreamCreate");
    }
    h->max_grid = sms * analyze_ctas_per_sm(cfg->n);
    if ((e = cudaMalloc(&h->scratch, (size_t)h->max_grid * analyze_scratch_bytes(cfg->n))) != cudaSuccess ||
        (e = cudaMalloc(&h->work_counter, 16)) != cudaSuccess) {
        aa_analyzer_destroy(h);
        return fail_cuda(e, "cudaMalloc(scratch)");
    }
    for (int i = 0; i < 2; ++i) {
        cudaEventCreateWithFlags(&h->slot[i].h2d_done, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->slot[i].k_done, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->slot[i].d2h_done, cudaEventDisableTiming);
    }
    *out = h;
    return AA_OK;
}

static void free_slot(StageSlot &s)
{
    cudaFree(s.in); cudaFree(s.raw); cudaFree(s.mags); cudaFree(s.feat); cudaFree(s.stable); cudaFree(s.summ);
    cudaFree(s.onset); cudaFree(s.dbg_floor); cudaFree(s.dbg_peaks);
    if (s.h2d_done) cudaEventDestroy(s.h2d_done);
    if (s.k_done) cudaEventDestroy(s.k_done);
    if (s.d2h_done) cudaEventDestroy(s.d2h_done);
    s = StageSlot();
}

extern "C" AA_API aa_status aa_analyzer_destroy(aa_analyzer *h)
{
    if (!h) return AA_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; ++i) free_slot(h->slot[i]);
    if (h->s_compute) cudaStreamDestroy(h->s_compute);
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    cudaFree(h->sl.in); cudaFree(h->sl.mags); cudaFree(h->sl.state); cudaFree(h->sl.feat); cudaFree(h->sl.stab);
    cudaFree(h->sl.summ); cudaFree(h->sl.onset); cudaFree(h->sl.raw[0]); cudaFree(h->sl.raw[1]);
    if (h->sl.events)
        for (int i = 0; i < 16; ++i) { cudaEventDestroy(h->sl.h2d_done[i]); cudaEventDestroy(h->sl.k_done[i]); }
    cudaFree(h->scratch);
    cudaFree(h->work_counter);
    cudaFree(h->seg_state);
    cudaFree(h->dt.mem);
    delete h;
    return AA_OK;
}

extern "C" AA_API int64_t aa_analyzer_last_launches(const aa_analyzer *h) { return h ? h->launches : 0; }

template <typename Tp>
static aa_status grow(Tp **p, size_t *cap, size_t need)
{
    if (need <= *cap) return AA_OK;
    cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    CU(cudaMalloc(p, need * sizeof(Tp)));
    *cap = need;
    return AA_OK;
}

// Time segments of a clip for the batch path (see AnalyzeParams::n_seg).  With more clips than resident CTAs the
// clips are dealt from a queue, and a batch ends with CTAs idling for up to one clip while the last ones finish
// (1024 clips on 444 CTAs: 5-6 % of the run).  Cutting the clips into segments that halve towards the end makes
// that ragged end one short segment; a segment costs one state hand-off through HBM (33 KB each way at n = 4096)
// and a refill of the hop ring.  AA_SEG_MIN (frames, environment, read once) overrides the shortest segment;
// 0 turns segmentation off.
static int plan_segments(int64_t T, int64_t n_clips, int grid, int *start)
{
    static const int min_seg = [] {
        const char *e = getenv("AA_SEG_MIN");
        return e ? atoi(e) : 32;
    }();
    start[0] = 0;
    start[1] = (int)T;
    if (min_seg <= 0 || n_clips <= grid || n_clips >= 16 * (int64_t)grid) return 1;
    int n = 0;
    int64_t pos = 0;
    while (n < aa::AA_MAX_SEG - 1 && T - pos > 2 * (int64_t)min_seg) {
        start[n++] = (int)pos;
        pos += (T - pos) / 2;
    }
    start[n++] = (int)pos;
    start[n] = (int)T;
    return n;
}

extern "C" AA_API int aa_plan_segments(int64_t T, int64_t n_clips, int resident_ctas, int32_t *starts)
{
    static_assert(AA_MAX_SEGMENTS == aa::AA_MAX_SEG, "header and kernel disagree on the segment limit");
    int tmp[aa::AA_MAX_SEG + 1];
    if (T < 0) T = 0;
    const int n = plan_segments(T, n_clips, resident_ctas, tmp);
    if (starts)
        for (int i = 0; i <= n; ++i) starts[i] = tmp[i];
    return n;
}

// out_T / out_f0: see AnalyzeParams (a time slice of longer clips writes frame f at clip * out_T + out_f0 + f); the
// summaries are not computed for a slice (out_T != T).
static aa_status analyze_device_impl(aa_analyzer *h, const float *clips_dev, int64_t n_clips, int64_t clip_len,
                                     int64_t clip_stride, const uint8_t *onset_in_dev, const aa_outputs *out,
                                     float *state, cudaStream_t s, int64_t *launches, int64_t out_T = 0,
                                     int64_t out_f0 = 0, unsigned long long *done_flag = nullptr,
                                     unsigned long long done_value = 0)
{
    const int64_t T = aa_num_frames(&h->cfg, clip_len);
    if (n_clips == 0 || T == 0) return AA_OK;
    if (((uintptr_t)clips_dev & 15u) || (clip_stride & 3))
        return fail(AA_ERR_INVALID,
                    "clips must be 16-byte aligned and clip_stride a multiple of 4 samples (TMA bulk copy)");
    if (out->summaries && !out->features) return fail(AA_ERR_INVALID, "summaries need the features output");
    AnalyzeParams p{};
    fill_params(h->cfg, h->dt.tab, &p);
    p.clips = clips_dev;
    p.n_clips = n_clips;
    p.clip_len = clip_len;
    p.clip_stride = clip_stride;
    p.T = T;
    p.out_T = out_T > 0 ? out_T : T;
    p.out_f0 = out_T > 0 ? out_f0 : 0;
    p.done_flag = done_flag;
    p.done_value = done_value;
    p.onset_in = onset_in_dev;
    p.mags = out->mags;
    p.features = out->features;
    p.stable = out->stable;
    p.dbg_floor = out->dbg_floor;
    p.dbg_peaks = out->dbg_peaks;
    p.state = state;
    p.scratch = h->scratch;
    p.grid = (int)std::min<int64_t>(n_clips, h->max_grid);
    // a grid that fills the GPU may run the resident sub-blocks of an SM as one CTA (analyze_kernel's SUBS; identical
    // results) IF the library was built with packed sub-blocks (AA_DEF_SUBS_4096 / AA_DEF_SUBS_2048: an experiment
    // build, off by default -- measured neutral).  AA_NO_PACK=1 in the environment (read per full-grid launch) is the
    // other arm of that A/B: one sub-block per CTA in the same build.
    p.packed = 0;
    if (p.grid == h->max_grid) {
        const char *np = getenv("AA_NO_PACK");
        p.packed = (np && np[0] == '1') ? 0 : 1;
    }
    // (with a carried state and more clips than resident CTAs the clips are dealt whole from the queue: every clip
    // loads its state block when it is taken and stores it when it is done)
    // more clips than resident CTAs: hand them out through a device-wide queue so that every SM ends up
    // with the same amount of work to within one clip
    p.work_counter = nullptr;
    p.n_seg = 1;
    p.seg_start[0] = 0;
    p.seg_start[1] = (int)T;
    if (n_clips > p.grid) {
        if (!state) p.n_seg = plan_segments(T, n_clips, p.grid, p.seg_start);
        if (p.n_seg > 1) {
            if ((size_t)n_clips > h->flag_clips_cap) {
                cudaFree(h->work_counter);
                h->work_counter = nullptr;
                h->flag_clips_cap = 0;
                CU(cudaMalloc(&h->work_counter, 16 + (size_t)n_clips * 2 * sizeof(unsigned)));
                h->flag_clips_cap = (size_t)n_clips;
            }
            aa_status st = grow(&h->seg_state, &h->seg_state_cap, (size_t)n_clips * state_floats(p.half));
            if (st != AA_OK) return st;
            p.seg_state = h->seg_state;
            p.seg_flags = reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(h->work_counter) + 16);
        }
        CU(cudaMemsetAsync(h->work_counter, 0, p.n_seg > 1 ? 16 + (size_t)n_clips * 2 * sizeof(unsigned) : 16, s));
        p.work_counter = h->work_counter;
    }
    CU(launch_analyze(p, s));
    ++*launches;
#ifdef AA_CHECKED
    {   // checked build: the kernel range-checks its data-dependent indices (aa_analyze.cu: AA_CHK)
        unsigned word = 0u;
        CU(analyze_check_word(&word, s));
        if (word) {
            char msg[96];
            snprintf(msg, sizeof msg, "checked build: index check failed in analyze_kernel, mask 0x%x", word);
            return fail(AA_ERR_CUDA, msg);
        }
    }
#endif
    if (out->summaries && p.out_T == T) {
        CU(launch_summaries(out->features, n_clips, T, out->summaries, s));
        ++*launches;
    }
    return AA_OK;
}

extern "C" AA_API aa_status aa_analyze_device(aa_analyzer *h, const float *clips_dev, int64_t n_clips,
                                              int64_t clip_len, int64_t clip_stride,
                                              const uint8_t *onset_in_dev, const aa_outputs *out_dev,
                                              void *stream)
{
    if (!h || !clips_dev || !out_dev || n_clips < 0 || clip_len < 0 || clip_stride < 0)
        return fail(AA_ERR_INVALID, "aa_analyze_device: bad argument");
    CU(cudaSetDevice(h->device));
    h->launches = 0;
    cudaStream_t s = (cudaStream_t)stream;   // NULL = the CUDA default stream
    return analyze_device_impl(h, clips_dev, n_clips, clip_len, clip_stride, onset_in_dev, out_dev, nullptr, s,
                               &h->launches);
}

extern "C" AA_API int64_t aa_state_floats(const aa_config *cfg)
{
    if (!cfg || !valid_n(cfg->n)) return 0;
    return (int64_t)state_floats(cfg->n / 2 + 1);
}

extern "C" AA_API aa_status aa_analyze_device_carry(aa_analyzer *h, const float *clips_dev, int64_t n_clips,
                                                    int64_t clip_len, int64_t clip_stride,
                                                    const uint8_t *onset_in_dev, const aa_outputs *out_dev,
                                                    float *state_dev, void *stream)
{
    if (!h || !clips_dev || !out_dev || !state_dev || n_clips < 0 || clip_len < 0 || clip_stride < 0)
        return fail(AA_ERR_INVALID, "aa_analyze_device_carry: bad argument");
    CU(cudaSetDevice(h->device));
    h->launches = 0;
    return analyze_device_impl(h, clips_dev, n_clips, clip_len, clip_stride, onset_in_dev, out_dev, state_dev,
                               (cudaStream_t)stream, &h->launches);
}

// format / channels: what the host buffer holds (AA_PCM_*, interleaved); mono f32 is copied straight into the
// analysis input, anything else goes through the ingest kernel (mod.rs:765-792) on the device
static aa_status analyze_host_body(aa_analyzer *h, const void *clips_host, int format, int channels, int64_t n_clips,
                                   int64_t clip_len, int64_t clip_stride, const uint8_t *onset_in_host,
                                   const aa_outputs *out_host);

// On any failure the copies of earlier clip groups may still be in flight to / from the caller's host buffers:
// drain the three streams before the error is returned, so the caller may free or reuse its buffers.
static aa_status analyze_host_impl(aa_analyzer *h, const void *clips_host, int format, int channels, int64_t n_clips,
                                   int64_t clip_len, int64_t clip_stride, const uint8_t *onset_in_host,
                                   const aa_outputs *out_host)
{
    const aa_status st = analyze_host_body(h, clips_host, format, channels, n_clips, clip_len, clip_stride,
                                           onset_in_host, out_host);
    if (st != AA_OK && h) {
        const std::string keep = g_err;
        cudaStreamSynchronize(h->s_h2d);
        cudaStreamSynchronize(h->s_compute);
        cudaStreamSynchronize(h->s_d2h);
        cudaGetLastError();
        g_err = keep;
    }
    return st;
}

// Time-sliced host pipeline.  The clip-group pipeline below overlaps copies and kernels across GROUPS OF CLIPS, so
// what stays exposed after the last copy is a whole clip's serial latency (a clip is walked frame by frame: 6 ms
// for 30 s at n = 4096) plus the records of a third of the batch.  Here the batch is cut in TIME instead: slice j
// holds the frames [f0_j, f1_j) of EVERY clip; its new samples go up with one strided copy per slice (each input byte
// still crosses PCIe once), its kernel starts from the analyzer state the previous slice left
// (the carry mechanism of aa_analyze_device_carry, so the records are byte-identical to a whole-clip run) and writes
// its records where a whole-clip launch would have put them (AnalyzeParams::out_T / out_f0); what stays exposed is
// one slice's kernel.  The device holds the whole batch (inputs, records, optionally magnitudes).
static aa_status analyze_host_sliced(aa_analyzer *h, const void *clips_host, int format, int channels, int64_t n_clips,
                                     int64_t clip_len, int64_t clip_stride, const uint8_t *onset_in_host,
                                     const aa_outputs *out_host, int64_t T, int n_slices)
{
    auto &S = h->sl;
    const bool direct = format == AA_PCM_F32 && channels == 1;
    const size_t fb = (format == AA_PCM_F32 ? 4 : 2) * (size_t)channels;       // bytes per input frame
    const int n = h->cfg.n, hop = h->cfg.hop, half = n / 2 + 1;
    const size_t span = (size_t)((n_clips - 1) * clip_stride + clip_len);
    const size_t frames = (size_t)(n_clips * T);
    const size_t sf = state_floats(half);
    aa_status st;
    if (!S.events) {
        for (int i = 0; i < 16; ++i) {
            CU(cudaEventCreateWithFlags(&S.h2d_done[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&S.k_done[i], cudaEventDisableTiming));
        }
        S.events = true;
    }
    if ((st = grow(&S.in, &S.in_cap, (span + 3) & ~(size_t)3)) != AA_OK) return st;
    if ((st = grow(&S.state, &S.state_cap, (size_t)n_clips * sf)) != AA_OK) return st;
    const bool want_feat = out_host->features || out_host->summaries;
    if (want_feat && (st = grow(&S.feat, &S.feat_cap, frames)) != AA_OK) return st;
    if (out_host->stable && (st = grow(&S.stab, &S.stab_cap, frames)) != AA_OK) return st;
    if (out_host->mags && (st = grow(&S.mags, &S.mags_cap, frames * half)) != AA_OK) return st;
    if (out_host->summaries && (st = grow(&S.summ, &S.summ_cap, (size_t)n_clips)) != AA_OK) return st;
    if (onset_in_host && (st = grow(&S.onset, &S.onset_cap, frames)) != AA_OK) return st;
    CU(cudaMemsetAsync(S.state, 0, sizeof(float) * (size_t)n_clips * sf, h->s_compute));   // fresh analyzers

    for (int j = 0; j < n_slices; ++j) {
        const int64_t f0 = T * j / n_slices, f1 = T * (j + 1) / n_slices, nf = f1 - f0;
        // samples of a clip this slice brings: everything up to the end of its last window that is not there yet
        const int64_t a = j == 0 ? 0 : f0 * hop + (n - hop);
        const int64_t b = j + 1 == n_slices ? clip_len : (f1 - 1) * hop + n;
        const int64_t width = b - a;
        // ---- H2D ----
        if (direct) {
            CU(cudaMemcpy2DAsync(S.in + a, sizeof(float) * (size_t)clip_stride,
                                 static_cast<const float *>(clips_host) + a, sizeof(float) * (size_t)clip_stride,
                                 sizeof(float) * (size_t)width, (size_t)n_clips, cudaMemcpyHostToDevice, h->s_h2d));
        } else {
            const int r = j & 1;
            if (j >= 2) CU(cudaStreamWaitEvent(h->s_h2d, S.k_done[j - 2], 0));      // its ingest kernel has read raw[r]
            if ((st = grow(&S.raw[r], &S.raw_cap[r], (size_t)n_clips * (size_t)width * fb + 16)) != AA_OK) return st;
            CU(cudaMemcpy2DAsync(S.raw[r], (size_t)width * fb,
                                 static_cast<const uint8_t *>(clips_host) + (size_t)a * fb, (size_t)clip_stride * fb,
                                 (size_t)width * fb, (size_t)n_clips, cudaMemcpyHostToDevice, h->s_h2d));
        }
        if (onset_in_host)
            CU(cudaMemcpy2DAsync(S.onset + f0, (size_t)T, onset_in_host + f0, (size_t)T, (size_t)nf, (size_t)n_clips,
                                 cudaMemcpyHostToDevice, h->s_h2d));
        CU(cudaEventRecord(S.h2d_done[j], h->s_h2d));
        // ---- kernels ----
        CU(cudaStreamWaitEvent(h->s_compute, S.h2d_done[j], 0));
        if (!direct) {
            CU(launch_ingest(S.raw[j & 1], format, channels, n_clips, width, width, clip_stride, S.in + a, h->num_sms,
                             h->s_compute));
            ++h->launches;
        }
        aa_outputs od{};
        od.mags = out_host->mags ? S.mags : nullptr;
        od.features = want_feat ? S.feat : nullptr;
        od.stable = out_host->stable ? S.stab : nullptr;
        st = analyze_device_impl(h, S.in + f0 * hop, n_clips, (nf - 1) * hop + n, clip_stride,
                                 onset_in_host ? S.onset : nullptr, &od, S.state, h->s_compute, &h->launches, T, f0);
        if (st != AA_OK) return st;
        CU(cudaEventRecord(S.k_done[j], h->s_compute));
        // ---- D2H: the columns [f0, f1) of every clip's rows ----
        CU(cudaStreamWaitEvent(h->s_d2h, S.k_done[j], 0));
        if (out_host->features)
            CU(cudaMemcpy2DAsync(out_host->features + f0, sizeof(aa_frame_features) * (size_t)T, S.feat + f0,
                                 sizeof(aa_frame_features) * (size_t)T, sizeof(aa_frame_features) * (size_t)nf,
                                 (size_t)n_clips, cudaMemcpyDeviceToHost, h->s_d2h));
        if (out_host->stable)
            CU(cudaMemcpy2DAsync(out_host->stable + f0, sizeof(aa_stable_pitches) * (size_t)T, S.stab + f0,
                                 sizeof(aa_stable_pitches) * (size_t)T, sizeof(aa_stable_pitches) * (size_t)nf,
                                 (size_t)n_clips, cudaMemcpyDeviceToHost, h->s_d2h));
        if (out_host->mags)
            CU(cudaMemcpy2DAsync(out_host->mags + (size_t)f0 * half, sizeof(float) * (size_t)T * half,
                                 S.mags + (size_t)f0 * half, sizeof(float) * (size_t)T * half,
                                 sizeof(float) * (size_t)nf * half, (size_t)n_clips, cudaMemcpyDeviceToHost, h->s_d2h));
    }
    if (out_host->summaries) {
        CU(launch_summaries(S.feat, n_clips, T, S.summ, h->s_compute));
        ++h->launches;
        CU(cudaEventRecord(S.k_done[15], h->s_compute));
        CU(cudaStreamWaitEvent(h->s_d2h, S.k_done[15], 0));
        CU(cudaMemcpyAsync(out_host->summaries, S.summ, sizeof(aa_clip_summary) * (size_t)n_clips,
                           cudaMemcpyDeviceToHost, h->s_d2h));
    }
    CU(cudaStreamSynchronize(h->s_d2h));
    CU(cudaStreamSynchronize(h->s_compute));
    return AA_OK;
}

static aa_status analyze_host_body(aa_analyzer *h, const void *clips_host, int format, int channels, int64_t n_clips,
                                   int64_t clip_len, int64_t clip_stride, const uint8_t *onset_in_host,
                                   const aa_outputs *out_host)
{
    if (!h || !clips_host || !out_host || n_clips < 0 || clip_len < 0 || clip_stride < 0)
        return fail(AA_ERR_INVALID, "aa_analyze_host: bad argument");
    const bool direct = format == AA_PCM_F32 && channels == 1;
    const size_t sample_bytes = format == AA_PCM_F32 ? 4 : 2;
    CU(cudaSetDevice(h->device));
    h->launches = 0;
    const int64_t T = aa_num_frames(&h->cfg, clip_len);
    if (n_clips == 0 || T == 0) return AA_OK;
    if (clip_stride & 3) return fail(AA_ERR_INVALID, "clip_stride must be a multiple of 4 samples");
    if (out_host->summaries && !out_host->features)
        return fail(AA_ERR_INVALID, "summaries need the features output");
    const int half = h->cfg.n / 2 + 1;

    // Large batches of non-overlapping clips go through the time-sliced pipeline (see analyze_host_sliced) as long as
    // the device can hold the batch; AA_HOST_SLICES in the environment overrides the slice count (0 = never).
    {
        const char *env = getenv("AA_HOST_SLICES");
        const int env_slices = env ? atoi(env) : -1;
        const size_t in_bytes = sizeof(float) * (size_t)((n_clips - 1) * clip_stride + clip_len);
        const size_t mags_bytes = out_host->mags ? sizeof(float) * (size_t)(n_clips * T) * half : 0;
        int n_slices = env_slices >= 0 ? env_slices : (int)std::min<int64_t>(12, T / 32);
        if (n_slices > 14) n_slices = 14;
        if (n_slices > T) n_slices = (int)T;          // (an override may ask for more slices than there are frames)
        if (n_slices >= 2 && !out_host->dbg_floor && !out_host->dbg_peaks && clip_stride >= clip_len &&
            (env_slices >= 2 || in_bytes >= ((size_t)32 << 20)) && in_bytes + mags_bytes <= ((size_t)64 << 30))
            return analyze_host_sliced(h, clips_host, format, channels, n_clips, clip_len, clip_stride, onset_in_host,
                                       out_host, T, n_slices);
    }

    // clip groups: at least ~2 waves of CTAs per group when there is enough work, at most 8 groups
    int64_t n_groups = n_clips / (2 * (int64_t)h->num_sms);
    n_groups = std::max<int64_t>(1, std::min<int64_t>(8, n_groups));
    const int64_t G = (n_clips + n_groups - 1) / n_groups;

    for (int64_t g = 0, c0 = 0; c0 < n_clips; ++g, c0 += G) {
        const int64_t nc = std::min(G, n_clips - c0);
        StageSlot &sl = h->slot[g & 1];
        // slot reuse: its previous D2H must have drained
        if (g >= 2) CU(cudaEventSynchronize(sl.d2h_done));
        const size_t span = (size_t)((nc - 1) * clip_stride + clip_len);
        const size_t span_pad = (span + 3) & ~(size_t)3;
        const size_t frames = (size_t)(nc * T);
        aa_status st;
        if ((st = grow(&sl.in, &sl.in_cap, span_pad)) != AA_OK) return st;
        if (out_host->mags && (st = grow(&sl.mags, &sl.mags_cap, frames * half)) != AA_OK) return st;
        if (frames > sl.frames_cap) {
            cudaFree(sl.feat); cudaFree(sl.stable);
            sl.feat = nullptr; sl.stable = nullptr; sl.frames_cap = 0;
            CU(cudaMalloc(&sl.feat, frames * sizeof(aa_frame_features)));
            CU(cudaMalloc(&sl.stable, frames * sizeof(aa_stable_pitches)));
            sl.frames_cap = frames;
        }
        if (out_host->summaries && (st = grow(&sl.summ, &sl.clips_cap, (size_t)nc)) != AA_OK) return st;
        if (out_host->dbg_floor && (st = grow(&sl.dbg_floor, &sl.dbgf_cap, frames * half)) != AA_OK) return st;
        if (out_host->dbg_peaks && (st = grow(&sl.dbg_peaks, &sl.dbgp_cap, frames * half)) != AA_OK) return st;
        if (onset_in_host && (st = grow(&sl.onset, &sl.onset_cap, frames)) != AA_OK) return st;

        // H2D (the kernel that last read this slot's input has finished: d2h_done above implies k_done)
        if (direct) {
            CU(cudaMemcpyAsync(sl.in, static_cast<const float *>(clips_host) + c0 * clip_stride, span * sizeof(float),
                               cudaMemcpyHostToDevice, h->s_h2d));
        } else {
            const size_t frame_bytes = sample_bytes * (size_t)channels;
            if ((st = grow(&sl.raw, &sl.raw_cap, span * frame_bytes + 16)) != AA_OK) return st;
            CU(cudaMemcpyAsync(sl.raw, static_cast<const uint8_t *>(clips_host) + (size_t)(c0 * clip_stride) * frame_bytes,
                               span * frame_bytes, cudaMemcpyHostToDevice, h->s_h2d));
        }
        if (onset_in_host)
            CU(cudaMemcpyAsync(sl.onset, onset_in_host + c0 * T, frames, cudaMemcpyHostToDevice, h->s_h2d));
        CU(cudaEventRecord(sl.h2d_done, h->s_h2d));

        // kernels
        CU(cudaStreamWaitEvent(h->s_compute, sl.h2d_done, 0));
        if (!direct) {
            CU(launch_ingest(sl.raw, format, channels, nc, clip_len, clip_stride, clip_stride, sl.in, h->num_sms,
                             h->s_compute));
            ++h->launches;
        }
        aa_outputs od{};
        od.mags = out_host->mags ? sl.mags : nullptr;
        od.features = out_host->features || out_host->summaries ? sl.feat : nullptr;
        od.stable = out_host->stable ? sl.stable : nullptr;
        od.summaries = out_host->summaries ? sl.summ : nullptr;
        od.dbg_floor = out_host->dbg_floor ? sl.dbg_floor : nullptr;
        od.dbg_peaks = out_host->dbg_peaks ? sl.dbg_peaks : nullptr;
        st = analyze_device_impl(h, sl.in, nc, clip_len, clip_stride, onset_in_host ? sl.onset : nullptr, &od,
                                 nullptr, h->s_compute, &h->launches);
        if (st != AA_OK) return st;
        CU(cudaEventRecord(sl.k_done, h->s_compute));

        // D2H
        CU(cudaStreamWaitEvent(h->s_d2h, sl.k_done, 0));
        const size_t f0 = (size_t)(c0 * T);
        if (out_host->mags)
            CU(cudaMemcpyAsync(out_host->mags + f0 * half, sl.mags, frames * half * sizeof(float),
                               cudaMemcpyDeviceToHost, h->s_d2h));
        if (out_host->features)
            CU(cudaMemcpyAsync(out_host->features + f0, sl.feat, frames * sizeof(aa_frame_features),
                               cudaMemcpyDeviceToHost, h->s_d2h));
        if (out_host->stable)
            CU(cudaMemcpyAsync(out_host->stable + f0, sl.stable, frames * sizeof(aa_stable_pitches),
                               cudaMemcpyDeviceToHost, h->s_d2h));
        if (out_host->summaries)
            CU(cudaMemcpyAsync(out_host->summaries + c0, sl.summ, (size_t)nc * sizeof(aa_clip_summary),
                               cudaMemcpyDeviceToHost, h->s_d2h));
        if (out_host->dbg_floor)
            CU(cudaMemcpyAsync(out_host->dbg_floor + f0 * half, sl.dbg_floor, frames * half * sizeof(float),
                               cudaMemcpyDeviceToHost, h->s_d2h));
        if (out_host->dbg_peaks)
            CU(cudaMemcpyAsync(out_host->dbg_peaks + f0 * half, sl.dbg_peaks, frames * half,
                               cudaMemcpyDeviceToHost, h->s_d2h));
        CU(cudaEventRecord(sl.d2h_done, h->s_d2h));
        // the next H2D into the *other* slot may start immediately; H2D into this slot
        // waits for d2h_done at the top of the loop
    }
    CU(cudaStreamSynchronize(h->s_d2h));
    CU(cudaStreamSynchronize(h->s_compute));
    return AA_OK;
}

extern "C" AA_API aa_status aa_analyze_host(aa_analyzer *h, const float *clips_host, int64_t n_clips,
                                            int64_t clip_len, int64_t clip_stride,
                                            const uint8_t *onset_in_host, const aa_outputs *out_host)
{
    return analyze_host_impl(h, clips_host, AA_PCM_F32, 1, n_clips, clip_len, clip_stride, onset_in_host, out_host);
}

static aa_status pcm_check(int32_t format, int32_t channels, const char *who)
{
    if (format != AA_PCM_F32 && format != AA_PCM_I16 && format != AA_PCM_U16)
        return fail(AA_ERR_INVALID, std::string(who) + ": format must be AA_PCM_F32, AA_PCM_I16 or AA_PCM_U16");
    if (channels < 1 || channels > 64) return fail(AA_ERR_INVALID, std::string(who) + ": 1 <= channels <= 64");
    return AA_OK;
}

extern "C" AA_API aa_status aa_analyze_host_pcm(aa_analyzer *h, const void *pcm_host, int32_t format, int32_t channels,
                                                int64_t n_clips, int64_t clip_len, int64_t clip_stride,
                                                const uint8_t *onset_in_host, const aa_outputs *out_host)
{
    aa_status st = pcm_check(format, channels, "aa_analyze_host_pcm");
    if (st != AA_OK) return st;
    return analyze_host_impl(h, pcm_host, format, channels, n_clips, clip_len, clip_stride, onset_in_host, out_host);
}

extern "C" AA_API aa_status aa_ingest_device(const void *pcm_dev, int32_t format, int32_t channels, int64_t n_clips,
                                             int64_t clip_len, int64_t in_stride, int64_t out_stride, float *mono_dev,
                                             void *stream)
{
    aa_status st = pcm_check(format, channels, "aa_ingest_device");
    if (st != AA_OK) return st;
    if (!pcm_dev || !mono_dev || n_clips < 0 || clip_len < 0 || in_stride < 0 || out_stride < 0 ||
        (reinterpret_cast<uintptr_t>(pcm_dev) & 15) || (reinterpret_cast<uintptr_t>(mono_dev) & 15))
        return fail(AA_ERR_INVALID, "aa_ingest_device: null, negative or not 16-byte aligned argument");
    int sms = 0;
    st = check_device(&sms);
    if (st != AA_OK) return st;
    CU(launch_ingest(pcm_dev, format, channels, n_clips, clip_len, in_stride, out_stride, mono_dev, sms,
                     (cudaStream_t)stream));
    return AA_OK;
}

// ---------------------------------------------------------------------------
// streaming
// ---------------------------------------------------------------------------
struct aa_stream {
    aa_analyzer *an = nullptr;
    int n = 0, hop = 0, half = 0;
    // sample ring: two linear buffers (ping-pong compaction) in pinned host memory MAPPED into the device address
    // space -- a push is a memcpy into the ring and a kernel launch; the kernel's TMA bulk copies read the hops
    // straight from host memory (4 KB per frame: a copy-engine transfer would cost more than the reads)
    float *h_buf[2] = {nullptr, nullptr};   // host addresses
    float *d_buf[2] = {nullptr, nullptr};   // the same memory as the device sees it
    int cur = 0;
    int64_t cap = 0;          // samples per buffer
    int64_t rd = 0, wr = 0;   // read / write positions in buf[cur]
    // completion: the kernel stores the launch's sequence number here (mapped host memory) after its last record
    unsigned long long *h_done = nullptr, *m_done = nullptr;
    unsigned long long launch_seq = 0;
    bool need_sync = false;   // a launch went through the device-side record buffers + copies (ring wrap-around)
    // analyzer state, outputs
    float *d_state = nullptr;
    aa_frame_features *d_feat = nullptr;
    aa_stable_pitches *d_stab = nullptr;
    uint8_t *d_onset = nullptr;          // device address of h_onset
    uint8_t *h_onset = nullptr;          // pinned, mapped
    aa_frame_features *h_feat = nullptr; // pinned ring (mapped)
    aa_stable_pitches *h_stab = nullptr; // pinned ring (mapped)
    aa_frame_features *m_feat = nullptr; // device addresses of the mapped rings
    aa_stable_pitches *m_stab = nullptr;
    int64_t max_frames_per_push = 0;
    int64_t out_cap = 0;                 // frames in the host result ring
    int64_t out_head = 0, out_count = 0; // ring of completed frames
    int64_t frame_index = 0;             // total frames produced
    bool onset_pending = false;
    cudaStream_t s = nullptr;
};

extern "C" AA_API aa_status aa_stream_destroy(aa_stream *h)
{
    if (!h) return AA_OK;
    if (h->an) cudaSetDevice(h->an->device);
    if (h->s) { cudaStreamSynchronize(h->s); cudaStreamDestroy(h->s); }
    cudaFree(h->d_state); cudaFree(h->d_feat); cudaFree(h->d_stab);
    cudaFreeHost(h->h_buf[0]); cudaFreeHost(h->h_buf[1]); cudaFreeHost(h->h_done);
    cudaFreeHost(h->h_onset); cudaFreeHost(h->h_feat); cudaFreeHost(h->h_stab);
    if (h->an) aa_analyzer_destroy(h->an);
    delete h;
    return AA_OK;
}

extern "C" AA_API aa_status aa_stream_create(const aa_config *cfg, aa_stream **out)
{
    if (!out) return fail(AA_ERR_INVALID, "out is null");
    *out = nullptr;
    aa_analyzer *an = nullptr;
    aa_status st = aa_analyzer_create(cfg, &an);
    if (st != AA_OK) return st;
    aa_stream *h = new aa_stream();
    h->an = an;
    h->n = cfg->n;
    h->hop = cfg->hop;
    h->half = cfg->n / 2 + 1;
    // the reference ring is max(8192, 4*window) samples (stft.rs:171, onset.rs:124); one push may carry
    // up to that much, a buffer holds 8x as much so compaction is rare
    const int64_t ring = std::max<int64_t>(8192, 4 * (int64_t)cfg->n);
    h->cap = 8 * ring;
    h->max_frames_per_push = ring / cfg->hop + 4;
    h->out_cap = 4 * h->max_frames_per_push;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    ok(cudaStreamCreateWithFlags(&h->s, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
        ok(cudaHostAlloc(&h->h_buf[b], sizeof(float) * h->cap, cudaHostAllocMapped));
        if (e == cudaSuccess) ok(cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->d_buf[b]), h->h_buf[b], 0));
    }
    ok(cudaHostAlloc(&h->h_done, 16 * sizeof(unsigned long long), cudaHostAllocMapped));   // [0..3] completion words, [4..] profiling stamps
    if (e == cudaSuccess) {
        std::memset(h->h_done, 0, 16 * sizeof(unsigned long long));
        ok(cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->m_done), h->h_done, 0));
    }
    ok(cudaMalloc(&h->d_state, sizeof(float) * state_floats(h->half)));
    ok(cudaMalloc(&h->d_feat, sizeof(aa_frame_features) * h->max_frames_per_push));
    ok(cudaMalloc(&h->d_stab, sizeof(aa_stable_pitches) * h->max_frames_per_push));
    ok(cudaHostAlloc(&h->h_onset, h->max_frames_per_push, cudaHostAllocMapped));
    if (e == cudaSuccess) ok(cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->d_onset), h->h_onset, 0));
    // the result ring is mapped into the device address space: the kernel writes the records of a push
    // straight into it (no device-to-host copies on the latency path) whenever they do not wrap around
    ok(cudaHostAlloc(&h->h_feat, sizeof(aa_frame_features) * h->out_cap, cudaHostAllocMapped));
    ok(cudaHostAlloc(&h->h_stab, sizeof(aa_stable_pitches) * h->out_cap, cudaHostAllocMapped));
    ok(cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->m_feat), h->h_feat, 0));
    ok(cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->m_stab), h->h_stab, 0));
    if (e == cudaSuccess) ok(cudaMemsetAsync(h->d_state, 0, sizeof(float) * state_floats(h->half), h->s));
    if (e == cudaSuccess) ok(cudaStreamSynchronize(h->s));
    if (e != cudaSuccess) {
        aa_stream_destroy(h);
        return fail_cuda(e, "aa_stream_create");
    }
    *out = h;
    return AA_OK;
}

// Every launch of this stream so far has written its records.  The usual case is a short spin on the completion word the
// kernel stores into mapped host memory (PCIe writes arrive in order: records first, then the word); a launch that
// went through device-side copies, or a word that does not arrive in 20 ms (a failed launch never stores it), takes
// cudaStreamSynchronize, which also reports the error.
static aa_status stream_wait(aa_stream *h)
{
    if (!h->need_sync) {
        volatile unsigned long long *f = h->h_done;
        const int words = aa::analyze_tail_warps();          // one completion word per record-writing warp
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            for (int i = 0; i < 256; ++i) {
                bool all = true;
                for (int w = 0; w < words; ++w) all = all && f[w] >= h->launch_seq;
                if (all) {
                    std::atomic_thread_fence(std::memory_order_acquire);
                    return AA_OK;
                }
#if defined(__x86_64__) || defined(__i386__)
                __builtin_ia32_pause();
#endif
            }
            if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(20)) break;
        }
    }
    CU(cudaStreamSynchronize(h->s));
    h->need_sync = false;
    return AA_OK;
}

extern "C" AA_API aa_status aa_stream_reset(aa_stream *h)
{
    if (!h) return fail(AA_ERR_INVALID, "stream is null");
    CU(cudaSetDevice(h->an->device));
    CU(cudaStreamSynchronize(h->s));
    h->need_sync = false;
    h->launch_seq = 0;
    std::memset(h->h_done, 0, 16 * sizeof(unsigned long long));
    CU(cudaMemsetAsync(h->d_state, 0, sizeof(float) * state_floats(h->half), h->s));
    CU(cudaStreamSynchronize(h->s));
    h->rd = h->wr = 0;
    h->cur = 0;
    h->out_head = h->out_count = 0;
    h->frame_index = 0;
    h->onset_pending = false;
    return AA_OK;
}

extern "C" AA_API aa_status aa_stream_set_noise_floor_db(aa_stream *h, float db)
{
    if (!h) return fail(AA_ERR_INVALID, "stream is null");
    h->an->cfg.noise_floor_db = db;   // picked up by fill_params at the next push (stft.rs:322)
    return AA_OK;
}

extern "C" AA_API aa_status aa_stream_signal_onset(aa_stream *h)
{
    if (!h) return fail(AA_ERR_INVALID, "stream is null");
    h->onset_pending = true;          // consumed by the next frame (stft.rs:387 swap(false))
    return AA_OK;
}

extern "C" AA_API aa_status aa_stream_push(aa_stream *h, const float *samples, int32_t count)
{
    if (!h || (!samples && count > 0) || count < 0) return fail(AA_ERR_INVALID, "aa_stream_push: bad argument");
    if (count == 0) return AA_OK;
    const int64_t ring = h->cap / 8;
    if (count > ring) return fail(AA_ERR_OVERFLOW, "aa_stream_push: more samples than the ring holds in one push");
    CU(cudaSetDevice(h->an->device));
    // Validate before anything is mutated: a push that fails consumed nothing, so the caller may poll
    // and retry the same samples (the frames this push completes must fit the result ring).
    {
        const int64_t avail = h->wr - h->rd + count;
        int64_t T = avail >= h->n ? (avail - h->n) / h->hop + 1 : 0;
        if (T > h->max_frames_per_push) T = h->max_frames_per_push;
        if (h->out_count + T > h->out_cap)
            return fail(AA_ERR_OVERFLOW, "aa_stream_push: result ring full, call aa_stream_poll (nothing was consumed)");
        const int64_t lead = h->rd - (h->rd & ~(int64_t)3);
        if (h->wr + count > h->cap && lead + avail > h->cap)
            return fail(AA_ERR_OVERFLOW, "aa_stream_push: sample ring full (nothing was consumed)");
    }
    // compact: move the unread tail to the other buffer when the linear buffer would overflow (the launches so far
    // may still be reading this one: wait for them first -- once per 7 ring lengths of samples)
    if (h->wr + count > h->cap) {
        aa_status stw = stream_wait(h);
        if (stw != AA_OK) return stw;
        const int64_t avail = h->wr - h->rd;
        // keep the read position 16-byte aligned for the TMA bulk copy
        const int64_t rd_al = h->rd & ~(int64_t)3;
        const int64_t lead = h->rd - rd_al;
        std::memcpy(h->h_buf[h->cur ^ 1], h->h_buf[h->cur] + rd_al, sizeof(float) * (size_t)(avail + lead));
        h->cur ^= 1;
        h->rd = lead;
        h->wr = lead + avail;
    }
    // (earlier launches read below wr only, so the new samples can be written while one is still running)
    std::memcpy(h->h_buf[h->cur] + h->wr, samples, sizeof(float) * (size_t)count);
    h->wr += count;

    const int64_t avail = h->wr - h->rd;
    if (avail < h->n) return AA_OK;
    int64_t T = (avail - h->n) / h->hop + 1;
    if (T > h->max_frames_per_push) T = h->max_frames_per_push;
    if (h->rd & 3) return fail(AA_ERR_INVALID, "aa_stream: hop must keep the read position 16-byte aligned");

    // onset_pending (stft.rs:387) reaches the tracker as a per-frame flag array; without a pending onset the
    // kernel takes a null array as "no onsets" and the copy is skipped
    const bool use_onset = (h->an->cfg.features & AA_FEAT_TRACKER) != 0 && h->onset_pending;
    if (use_onset) {
        aa_status stw = stream_wait(h);          // an earlier launch may still be reading the flag array
        if (stw != AA_OK) return stw;
        std::memset(h->h_onset, 0, (size_t)T);
        h->h_onset[0] = 1;
        h->onset_pending = false;
    }
    // results: straight into the mapped host ring when the T records are contiguous there, else through the
    // device buffers and two-segment copies
    const int64_t tail = (h->out_head + h->out_count) % h->out_cap;
    const int64_t first = std::min(T, h->out_cap - tail);
    const bool direct = first == T;
    aa_outputs od{};
    od.features = direct ? h->m_feat + tail : h->d_feat;
    od.stable = direct ? h->m_stab + tail : h->d_stab;
    const int64_t clip_len = h->n + (T - 1) * h->hop;
    int64_t launches = 0;
    aa_status st = analyze_device_impl(h->an, h->d_buf[h->cur] + h->rd, 1, clip_len, (clip_len + 3) & ~(int64_t)3,
                                       use_onset ? h->d_onset : nullptr, &od, h->d_state, h->s, &launches, 0, 0,
                                       h->m_done, h->launch_seq + 1);
    if (st != AA_OK) return st;
    ++h->launch_seq;
    if (!direct) {
        h->need_sync = true;
        CU(cudaMemcpyAsync(h->h_feat + tail, h->d_feat, sizeof(aa_frame_features) * (size_t)first,
                           cudaMemcpyDeviceToHost, h->s));
        CU(cudaMemcpyAsync(h->h_stab + tail, h->d_stab, sizeof(aa_stable_pitches) * (size_t)first,
                           cudaMemcpyDeviceToHost, h->s));
        CU(cudaMemcpyAsync(h->h_feat, h->d_feat + first, sizeof(aa_frame_features) * (size_t)(T - first),
                           cudaMemcpyDeviceToHost, h->s));
        CU(cudaMemcpyAsync(h->h_stab, h->d_stab + first, sizeof(aa_stable_pitches) * (size_t)(T - first),
                           cudaMemcpyDeviceToHost, h->s));
    }
    h->out_count += T;
    h->rd += T * h->hop;
    return AA_OK;
}

extern "C" AA_API aa_status aa_stream_poll(aa_stream *h, aa_stream_frame *out, int32_t max, int32_t *n_out)
{
    if (!h || !n_out || (max > 0 && !out)) return fail(AA_ERR_INVALID, "aa_stream_poll: bad argument");
    *n_out = 0;
    if (h->out_count == 0 || max <= 0) return AA_OK;
    CU(cudaSetDevice(h->an->device));
    aa_status stw = stream_wait(h);
    if (stw != AA_OK) return stw;
    int32_t n = (int32_t)std::min<int64_t>(max, h->out_count);
    for (int32_t i = 0; i < n; ++i) {
        const int64_t idx = (h->out_head + i) % h->out_cap;
        out[i].frame_index = h->frame_index + i;
        out[i].features = h->h_feat[idx];
        out[i].stable = h->h_stab[idx];
    }
    h->out_head = (h->out_head + n) % h->out_cap;
    h->out_count -= n;
    h->frame_index += n;
    *n_out = n;
    return AA_OK;
}

#ifdef AA_STREAM_PROF      // experiment only: the %globaltimer stamps of the last launch
extern "C" AA_API const unsigned long long *aa_stream_debug_stamps(aa_stream *h) { return h ? h->h_done : nullptr; }
#endif

extern "C" AA_API aa_status aa_stream_probe_latency(aa_stream *h, const float *samples, int32_t count, int32_t n_pushes,
                                                    double *latency_us, int64_t *frames_out)
{
    if (!h || !samples || !latency_us || count <= 0 || n_pushes < 0)
        return fail(AA_ERR_INVALID, "aa_stream_probe_latency: bad argument");
    std::vector<aa_stream_frame> buf((size_t)h->max_frames_per_push + 1);
    int64_t frames = 0;
    for (int32_t i = 0; i < n_pushes; ++i) {
        const auto t0 = std::chrono::steady_clock::now();
        aa_status st = aa_stream_push(h, samples + (size_t)i * (size_t)count, count);
        if (st != AA_OK) return st;
        int32_t got = 0;
        st = aa_stream_poll(h, buf.data(), (int32_t)buf.size(), &got);
        if (st != AA_OK) return st;
        latency_us[i] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        frames += got;
    }
    if (frames_out) *frames_out = frames;
    return AA_OK;
}

// ---------------------------------------------------------------------------
// note identification (theory.rs:195-209)
// ---------------------------------------------------------------------------
extern "C" AA_API aa_status aa_notes_from_stable_device(const aa_stable_pitches *stable_dev, int64_t n_frames,
                                                        float base_freq, aa_note_record *notes_dev, void *stream)
{
    if (!stable_dev || !notes_dev || n_frames < 0 || !(base_freq > 0.0f))
        return fail(AA_ERR_INVALID, "aa_notes_from_stable_device: bad argument");
    aa_status st = check_device(nullptr);
    if (st != AA_OK) return st;
    const float base_c0 = base_freq * powf(2.0f, -4.75f);   // theory.rs:197
    CU(launch_notes(stable_dev, n_frames, base_c0, notes_dev, (cudaStream_t)stream));
    return AA_OK;
}

extern "C" AA_API aa_status aa_notes_from_stable_host(const aa_stable_pitches *stable_host, int64_t n_frames,
                                                      float base_freq, aa_note_record *notes_host)
{
    if (!stable_host || !notes_host || n_frames < 0 || !(base_freq > 0.0f))
        return fail(AA_ERR_INVALID, "aa_notes_from_stable_host: bad argument");
    if (n_frames == 0) return AA_OK;
    aa_status st = check_device(nullptr);
    if (st != AA_OK) return st;
    aa_stable_pitches *d_in = nullptr;
    aa_note_record *d_out = nullptr;
    CU(cudaMalloc(&d_in, sizeof(aa_stable_pitches) * (size_t)n_frames));
    cudaError_t e = cudaMalloc(&d_out, sizeof(aa_note_record) * (size_t)n_frames);
    if (e == cudaSuccess) e = cudaMemcpy(d_in, stable_host, sizeof(aa_stable_pitches) * (size_t)n_frames, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_notes(d_in, n_frames, base_freq * powf(2.0f, -4.75f), d_out, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(notes_host, d_out, sizeof(aa_note_record) * (size_t)n_frames, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail_cuda(e, "aa_notes_from_stable_host");
    return AA_OK;
}

// ---------------------------------------------------------------------------
// YIN-style lag search (a14, NEW)
// ---------------------------------------------------------------------------
static aa_status yin_check(const aa_yin_config *cfg, int64_t clip_len, int64_t *T)
{
    if (!cfg) return fail(AA_ERR_INVALID, "yin config is null");
    if (cfg->n < 8 || cfg->n > 8192 || cfg->hop <= 0) return fail(AA_ERR_UNSUPPORTED, "yin: 8 <= n <= 8192, hop > 0");
    if (cfg->min_lag < 1 || cfg->max_lag < cfg->min_lag || cfg->max_lag >= cfg->n)
        return fail(AA_ERR_INVALID, "yin: need 1 <= min_lag <= max_lag < n");
    *T = clip_len < cfg->n ? 0 : (clip_len - cfg->n) / cfg->hop + 1;
    return AA_OK;
}

extern "C" AA_API aa_status aa_yin_device(const aa_yin_config *cfg, const float *clips_dev, int64_t n_clips,
                                          int64_t clip_len, int64_t clip_stride, int32_t *lag_dev, float *cmnd_dev,
                                          void *stream)
{
    int64_t T = 0;
    aa_status st = yin_check(cfg, clip_len, &T);
    if (st != AA_OK) return st;
    if (!clips_dev || !lag_dev || n_clips < 0 || clip_stride < 0) return fail(AA_ERR_INVALID, "aa_yin_device: bad argument");
    int sms = 0;
    st = check_device(&sms);
    if (st != AA_OK) return st;
    CU(launch_yin(clips_dev, n_clips, clip_stride, T, cfg->n, cfg->hop, cfg->min_lag, cfg->max_lag, cfg->threshold,
                  lag_dev, cmnd_dev, sms, (cudaStream_t)stream));
    return AA_OK;
}

extern "C" AA_API aa_status aa_yin_host(const aa_yin_config *cfg, const float *clips_host, int64_t n_clips,
                                        int64_t clip_len, int64_t clip_stride, int32_t *lag_host, float *cmnd_host)
{
    int64_t T = 0;
    aa_status st = yin_check(cfg, clip_len, &T);
    if (st != AA_OK) return st;
    if (!clips_host || !lag_host || n_clips < 0 || clip_stride < 0) return fail(AA_ERR_INVALID, "aa_yin_host: bad argument");
    if (n_clips == 0 || T == 0) return AA_OK;
    int sms = 0;
    st = check_device(&sms);
    if (st != AA_OK) return st;
    const size_t span = (size_t)((n_clips - 1) * clip_stride + clip_len);
    const size_t frames = (size_t)(n_clips * T);
    float *d_in = nullptr, *d_c = nullptr;
    int32_t *d_lag = nullptr;
    cudaError_t e = cudaMalloc(&d_in, span * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_lag, frames * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_c, frames * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(d_in, clips_host, span * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = launch_yin(d_in, n_clips, clip_stride, T, cfg->n, cfg->hop, cfg->min_lag, cfg->max_lag, cfg->threshold, d_lag,
                       d_c, sms, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(lag_host, d_lag, frames * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && cmnd_host) e = cudaMemcpy(cmnd_host, d_c, frames * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_lag); cudaFree(d_c);
    if (e != cudaSuccess) return fail_cuda(e, "aa_yin_host");
    return AA_OK;
}

// ---------------------------------------------------------------------------
// synthetic clips
// ---------------------------------------------------------------------------
extern "C" AA_API aa_status aa_synth_clips_device(float *clips_dev, int64_t n_clips, int64_t clip_len,
                                                  int64_t clip_stride, float sample_rate, uint64_t seed,
                                                  void *stream)
{
    if (!clips_dev || n_clips < 0 || clip_len < 0 || clip_stride < clip_len || !(sample_rate > 0.0f))
        return fail(AA_ERR_INVALID, "aa_synth_clips_device: bad argument");
    aa_status st = check_device(nullptr);
    if (st != AA_OK) return st;
    CU(launch_synth(clips_dev, n_clips, clip_len, clip_stride, sample_rate, seed, (cudaStream_t)stream));
    return AA_OK;
}

// ---------------------------------------------------------------------------
// input conditioning chain: reducer-thread filters + gate (mod.rs:351-487) and
// DynamicsTracker (dynamics.rs:156-360)
// ---------------------------------------------------------------------------
struct aa_conditioner {
    aa_cond_config cfg;
    aa::CondParams p;
    int device = 0, sms = 0;
    float *stats = nullptr, *gains = nullptr, *agc_state = nullptr, *carry = nullptr;
    size_t stats_cap = 0, gains_cap = 0, agc_cap = 0, carry_cap = 0;     // in floats
    int64_t carry_clips = 0;                                            // clips the carried state belongs to
};

// mod.rs:357-385, in f32 with the reference's operation order (host libm, like the reference)
static void calc_biquad(float freq, bool is_lpf, float sample_rate, float out[5])
{
    const float PI_F = 3.14159265358979323846f;
    const float w0 = 2.0f * PI_F * freq / sample_rate;
    const float cos_w0 = cosf(w0), sin_w0 = sinf(w0);
    const float alpha = sin_w0 / (2.0f * 0.707f);
    float b0, b1, b2;
    if (is_lpf) { b0 = (1.0f - cos_w0) / 2.0f; b1 = 1.0f - cos_w0; b2 = (1.0f - cos_w0) / 2.0f; }
    else        { b0 = (1.0f + cos_w0) / 2.0f; b1 = -(1.0f + cos_w0); b2 = (1.0f + cos_w0) / 2.0f; }
    const float a0 = 1.0f + alpha, a1 = -2.0f * cos_w0, a2 = 1.0f - alpha;
    out[0] = b0 / a0; out[1] = b1 / a0; out[2] = b2 / a0; out[3] = a1 / a0; out[4] = a2 / a0;
}

extern "C" AA_API int64_t aa_cond_num_slots(const aa_cond_config *cfg, int64_t clip_len)
{
    if (!cfg || cfg->slot_len <= 0 || clip_len < 0) return 0;
    return clip_len / cfg->slot_len;
}

extern "C" AA_API aa_status aa_conditioner_create(const aa_cond_config *cfg, aa_conditioner **out)
{
    if (!cfg || !out) return fail(AA_ERR_INVALID, "aa_conditioner_create: null argument");
    *out = nullptr;
    if (!(cfg->sample_rate > 0.0f) || cfg->slot_len < 4 || (cfg->slot_len & 3))
        return fail(AA_ERR_INVALID, "aa_conditioner_create: sample_rate must be > 0 and slot_len a positive multiple of 4");
    int sms = 0;
    aa_status st = check_device(&sms);
    if (st != AA_OK) return st;
    aa_conditioner *h = new (std::nothrow) aa_conditioner();
    if (!h) return fail(AA_ERR_INVALID, "out of host memory");
    h->cfg = *cfg;
    h->device = g_device;
    h->sms = sms;
    aa::CondParams &p = h->p;
    const float sr = cfg->sample_rate;
    calc_biquad(40.0f, false, sr, p.hp);                                   // mod.rs:387
    calc_biquad(14000.0f, true, sr, p.lp);                                 // mod.rs:388
    p.gate_threshold_linear = powf(10.0f, -60.0f / 20.0f);                 // mod.rs:400-401
    p.release_coeff = expf(-1.0f / (0.040f * sr));                         // mod.rs:408
    p.gate_hold_samples = (int32_t)(0.020f * sr);                          // mod.rs:413
    const float slot_rate = sr / (float)cfg->slot_len;                     // dynamics.rs:164
    p.target_db = -18.0f;                                                  // mod.rs:347-355
    p.max_boost_db = 100.0f;
    p.smooth_alpha = 1.0f - expf(-1.0f / (240.0f * slot_rate));            // dynamics.rs:174
    p.silence_decay_alpha = 1.0f - expf(-1.0f / (10.0f * slot_rate));      // dynamics.rs:175
    p.active_snr_db = 20.0f;                                               // dynamics.rs:188
    p.bootstrap_floor_db = -55.0f;                                         // dynamics.rs:189
    p.slot_len = cfg->slot_len;
    *out = h;
    return AA_OK;
}

extern "C" AA_API aa_status aa_conditioner_destroy(aa_conditioner *h)
{
    if (!h) return AA_OK;
    cudaSetDevice(h->device);
    cudaFree(h->stats); cudaFree(h->gains); cudaFree(h->agc_state); cudaFree(h->carry);
    delete h;
    return AA_OK;
}

extern "C" AA_API aa_status aa_conditioner_reset(aa_conditioner *h)
{
    if (!h) return fail(AA_ERR_INVALID, "aa_conditioner_reset: null handle");
    h->carry_clips = 0;       // the next call re-initialises the carried state
    return AA_OK;
}

static aa_status cond_grow(float **buf, size_t *cap, size_t need)
{
    if (*cap >= need) return AA_OK;
    cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    CU(cudaMalloc(buf, need * sizeof(float)));
    *cap = need;
    return AA_OK;
}

extern "C" AA_API aa_status aa_condition_device(aa_conditioner *h, float *clips_dev, int64_t n_clips, int64_t clip_len,
                                                int64_t clip_stride, aa_dynamics *dyn_dev, void *stream)
{
    if (!h || !clips_dev) return fail(AA_ERR_INVALID, "aa_condition_device: null argument");
    if (n_clips < 0 || clip_len < 0 || clip_stride < 0 || (clip_stride & 3) || (reinterpret_cast<uintptr_t>(clips_dev) & 15))
        return fail(AA_ERR_INVALID, "aa_condition_device: clips must be 16-byte aligned with clip_stride % 4 == 0");
    if (n_clips > 1 && clip_stride < clip_len)
        return fail(AA_ERR_INVALID, "aa_condition_device: overlapping clips cannot be conditioned in place");
    const int64_t n_slots = clip_len / h->cfg.slot_len;
    if (n_clips == 0 || n_slots == 0) return AA_OK;
    CU(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const bool agc = (h->cfg.flags & AA_COND_AGC) != 0;
    const bool carry = (h->cfg.flags & AA_COND_CARRY) != 0;
    aa_status st;
    if (carry) {
        if ((st = cond_grow(&h->carry, &h->carry_cap, (size_t)n_clips * 16)) != AA_OK) return st;
        if (h->carry_clips != n_clips) {     // first use (or a different batch shape): start from the initial state
            CU(cudaMemsetAsync(h->carry, 0, (size_t)n_clips * 16 * sizeof(float), s));
        }
    }
    if (agc) {
        const size_t per = aa::cond_agc_state_floats();
        if ((st = cond_grow(&h->stats, &h->stats_cap, (size_t)(n_clips * n_slots) * 4)) != AA_OK) return st;
        if ((st = cond_grow(&h->gains, &h->gains_cap, (size_t)(n_clips * n_slots))) != AA_OK) return st;
        const size_t old_cap = h->agc_cap;
        if ((st = cond_grow(&h->agc_state, &h->agc_cap, (size_t)n_clips * per)) != AA_OK) return st;
        if (carry && (h->carry_clips != n_clips || old_cap != h->agc_cap))   // scalars (incl. the valid mark) to zero
            CU(cudaMemsetAsync(h->agc_state, 0, (size_t)n_clips * per * sizeof(float), s));
    }
    if (carry) h->carry_clips = n_clips;
    CU(aa::launch_cond_filter_gate(clips_dev, n_clips, clip_stride, n_slots, h->p, agc ? h->stats : nullptr,
                                   carry ? h->carry : nullptr, s));
    if (agc) {
        CU(aa::launch_cond_agc(h->stats, n_clips, n_slots, h->p, h->agc_state, carry ? 1 : 0, h->gains, dyn_dev, s));
        CU(aa::launch_cond_apply_gain(clips_dev, n_clips, clip_stride, n_slots, h->cfg.slot_len, h->gains, h->sms, s));
    }
    return AA_OK;
}

extern "C" AA_API aa_status aa_condition_host(aa_conditioner *h, float *clips_host, int64_t n_clips, int64_t clip_len,
                                              int64_t clip_stride, aa_dynamics *dyn_host)
{
    if (!h || !clips_host) return fail(AA_ERR_INVALID, "aa_condition_host: null argument");
    if (n_clips < 0 || clip_len < 0 || clip_stride < 0 || (clip_stride & 3))
        return fail(AA_ERR_INVALID, "aa_condition_host: clip_stride must be a non-negative multiple of 4");
    const int64_t n_slots = clip_len / h->cfg.slot_len;
    if (n_clips == 0 || n_slots == 0) return AA_OK;
    CU(cudaSetDevice(h->device));
    const size_t span = (size_t)((n_clips - 1) * clip_stride + clip_len);
    const size_t nd = (size_t)(n_clips * n_slots);
    float *d = nullptr;
    aa_dynamics *dd = nullptr;
    cudaError_t e = cudaMalloc(&d, span * sizeof(float));
    if (e == cudaSuccess && dyn_host && (h->cfg.flags & AA_COND_AGC)) e = cudaMalloc(&dd, nd * sizeof(aa_dynamics));
    if (e == cudaSuccess) e = cudaMemcpy(d, clips_host, span * sizeof(float), cudaMemcpyHostToDevice);
    aa_status st = AA_OK;
    if (e == cudaSuccess) st = aa_condition_device(h, d, n_clips, clip_len, clip_stride, dd, nullptr);
    if (e == cudaSuccess && st == AA_OK) e = cudaMemcpy(clips_host, d, span * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && st == AA_OK && dd) e = cudaMemcpy(dyn_host, dd, nd * sizeof(aa_dynamics), cudaMemcpyDeviceToHost);
    cudaFree(d); cudaFree(dd);
    if (st != AA_OK) return st;
    if (e != cudaSuccess) return fail_cuda(e, "aa_condition_host");
    return AA_OK;
}

// ---------------------------------------------------------------------------
// Tuner::run per-frame branch + Interval::new: tuner.rs:148-193, theory.rs:306-382
// ---------------------------------------------------------------------------
extern "C" AA_API aa_status aa_tuner_from_stable_device(const aa_stable_pitches *stable_dev, int64_t n_frames,
                                                        float base_freq, int32_t system, int32_t single_pitch_mode,
                                                        aa_tuner_record *out_dev, void *stream)
{
    if (!stable_dev || !out_dev || n_frames < 0 || !(base_freq > 0.0f) || system < 0 || system > 2)
        return fail(AA_ERR_INVALID, "aa_tuner_from_stable_device: bad argument (system is 0, 1 or 2)");
    aa_status st = check_device(nullptr);
    if (st != AA_OK) return st;
    CU(launch_tuner(stable_dev, n_frames, base_freq * powf(2.0f, -4.75f), system, single_pitch_mode ? 1 : 0, out_dev,
                    (cudaStream_t)stream));
    return AA_OK;
}

extern "C" AA_API aa_status aa_tuner_from_stable_host(const aa_stable_pitches *stable_host, int64_t n_frames,
                                                      float base_freq, int32_t system, int32_t single_pitch_mode,
                                                      aa_tuner_record *out_host)
{
    if (!stable_host || !out_host || n_frames < 0 || !(base_freq > 0.0f) || system < 0 || system > 2)
        return fail(AA_ERR_INVALID, "aa_tuner_from_stable_host: bad argument (system is 0, 1 or 2)");
    if (n_frames == 0) return AA_OK;
    aa_status st = check_device(nullptr);
    if (st != AA_OK) return st;
    aa_stable_pitches *d_in = nullptr;
    aa_tuner_record *d_out = nullptr;
    CU(cudaMalloc(&d_in, sizeof(aa_stable_pitches) * (size_t)n_frames));
    cudaError_t e = cudaMalloc(&d_out, sizeof(aa_tuner_record) * (size_t)n_frames);
    if (e == cudaSuccess) e = cudaMemcpy(d_in, stable_host, sizeof(aa_stable_pitches) * (size_t)n_frames, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = launch_tuner(d_in, n_frames, base_freq * powf(2.0f, -4.75f), system, single_pitch_mode ? 1 : 0, d_out, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(out_host, d_out, sizeof(aa_tuner_record) * (size_t)n_frames, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail_cuda(e, "aa_tuner_from_stable_host");
    return AA_OK;
}

// ---------------------------------------------------------------------------
// offline onset events: onset.rs:383-456, timing.rs:311-337
// ---------------------------------------------------------------------------
static aa_status onset_events_check(const aa_config *cfg, int64_t n_clips, int64_t clip_len, float bpm, int32_t max_events,
                                    int64_t *T)
{
    if (!cfg || n_clips < 0 || clip_len < 0 || !(bpm > 0.0f) || max_events < 1 || cfg->n <= 0 || cfg->hop <= 0 ||
        !(cfg->sample_rate > 0.0f))
        return fail(AA_ERR_INVALID, "aa_onset_events: bad argument");
    *T = aa_num_frames(cfg, clip_len);
    return AA_OK;
}

extern "C" AA_API aa_status aa_onset_events_device(const aa_config *cfg, const aa_frame_features *features_dev,
                                                   int64_t n_clips, int64_t clip_len, float bpm, int32_t max_events,
                                                   aa_onset_event *events_dev, int32_t *counts_dev, void *stream)
{
    int64_t T = 0;
    aa_status st = onset_events_check(cfg, n_clips, clip_len, bpm, max_events, &T);
    if (st != AA_OK) return st;
    if (!features_dev || !events_dev || !counts_dev) return fail(AA_ERR_INVALID, "aa_onset_events_device: null pointer");
    st = check_device(nullptr);
    if (st != AA_OK) return st;
    const double bps = (double)bpm / (60.0 * (double)cfg->sample_rate);                 // timing.rs:313-315
    CU(launch_onset_events(features_dev, n_clips, T, cfg->n, cfg->hop, bps, max_events, events_dev, counts_dev,
                           (cudaStream_t)stream));
    return AA_OK;
}

extern "C" AA_API aa_status aa_onset_events_host(const aa_config *cfg, const aa_frame_features *features_host,
                                                 int64_t n_clips, int64_t clip_len, float bpm, int32_t max_events,
                                                 aa_onset_event *events_host, int32_t *counts_host)
{
    int64_t T = 0;
    aa_status st = onset_events_check(cfg, n_clips, clip_len, bpm, max_events, &T);
    if (st != AA_OK) return st;
    if (!features_host || !events_host || !counts_host) return fail(AA_ERR_INVALID, "aa_onset_events_host: null pointer");
    if (n_clips == 0) return AA_OK;
    st = check_device(nullptr);
    if (st != AA_OK) return st;
    const size_t nf = (size_t)(n_clips * T), ne = (size_t)n_clips * (size_t)max_events;
    aa_frame_features *d_f = nullptr;
    aa_onset_event *d_e = nullptr;
    int32_t *d_c = nullptr;
    cudaError_t e = cudaMalloc(&d_f, std::max<size_t>(nf, 1) * sizeof(aa_frame_features));
    if (e == cudaSuccess) e = cudaMalloc(&d_e, ne * sizeof(aa_onset_event));
    if (e == cudaSuccess) e = cudaMalloc(&d_c, (size_t)n_clips * sizeof(int32_t));
    if (e == cudaSuccess && nf) e = cudaMemcpy(d_f, features_host, nf * sizeof(aa_frame_features), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d_e, 0, ne * sizeof(aa_onset_event));
    if (e == cudaSuccess)
        e = launch_onset_events(d_f, n_clips, T, cfg->n, cfg->hop, (double)bpm / (60.0 * (double)cfg->sample_rate),
                                max_events, d_e, d_c, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(events_host, d_e, ne * sizeof(aa_onset_event), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(counts_host, d_c, (size_t)n_clips * sizeof(int32_t), cudaMemcpyDeviceToHost);
    cudaFree(d_f); cudaFree(d_e); cudaFree(d_c);
    if (e != cudaSuccess) return fail_cuda(e, "aa_onset_events_host");
    return AA_OK;
}

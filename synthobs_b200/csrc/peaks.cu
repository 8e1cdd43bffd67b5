// Measurement support (bench.py): (1) register-only micro-benchmarks of the pipes the imaging and dust-spectra
// kernels are bound by -- FP32 FFMA, packed FFMA2, MUFU.EX2, and the dust kernel's own per-particle-wavelength
// arithmetic with and without its polynomial share -- so roofline denominators are MEASURED on the box the bench
// runs on ("the builder must measure them"); (2) per-kernel timing: when enabled, every
// labelled launch inside the library is bracketed by CUDA events on its own stream (SO_PROF in common.cuh) and the
// per-label totals are reported as text.  Neither touches user data.
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "exp2.cuh"

namespace {

constexpr int kPeakThreads = 256;

// 16 independent FFMA chains per thread
__global__ void __launch_bounds__(kPeakThreads) peak_ffma_kernel(int64_t iters, float* out) {
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = 1.0f + 1e-3f * (float)(threadIdx.x + j);
    const float b = 0.999f, c = 1e-3f;
    for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = fmaf(a[j], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += a[j];
    if (s == 123.456f) out[0] = s;
}

// 16 independent packed chains (FFMA2: two FP32 multiply-adds per lane and instruction)
__global__ void __launch_bounds__(kPeakThreads) peak_ffma2_kernel(int64_t iters, float* out) {
    float2 a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = make_float2(1.0f + 1e-3f * (float)(threadIdx.x + j), 1.0f + 2e-3f * (float)j);
    const float2 b = make_float2(0.999f, 0.998f), c = make_float2(1e-3f, 2e-3f);
    for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = __ffma2_rn(a[j], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += a[j].x + a[j].y;
    if (s == 123.456f) out[0] = s;
}

// 16 independent ex2.approx chains (x -> 2^-x stays in (0.5, 1))
__global__ void __launch_bounds__(kPeakThreads) peak_ex2_kernel(int64_t iters, float* out) {
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = 0.5f + 1e-3f * (float)((threadIdx.x + j) & 63);
    for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = ex2a(-a[j]);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += a[j];
    if (s == 123.456f) out[0] = s;
}

// the inner loop of dust_spectra_kernel (old particles: one exponential per particle and wavelength, sed_dust.cu
// old_one) on register operands only: per "particle" and wavelength pair  d = m * 2^(tau * k) + d.
// HYBRID: every fourth particle takes the FMA-pipe polynomial, exactly as in the kernel.
template <bool HYBRID>
__global__ void __launch_bounds__(kPeakThreads) peak_dust_kernel(int64_t iters, float* out) {
    float2 d[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    const float2 k[2] = {make_float2(-0.7f, -0.9f), make_float2(-1.1f, -1.3f)};
    float t = 0.1f + 1e-3f * (float)(threadIdx.x & 31);
    const float2 m2 = make_float2(1.0f, 1.0f);
    for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 t2 = make_float2(t, t);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 a = __fmul2_rn(t2, k[h]);
                d[h] = __ffma2_rn(m2, (HYBRID && q == 3) ? ex2_poly2(a) : ex2_mufu2(a), d[h]);
            }
            t += 1e-4f;
        }
    }
    const float s = d[0].x + d[0].y + d[1].x + d[1].y;
    if (s == 123.456f) out[0] = s;
}

struct ProfMark {
    std::string label;
    cudaEvent_t e0, e1;
};
struct ProfState {
    bool on = false;
    std::vector<ProfMark> marks;
};
ProfState& prof() {
    static ProfState p;
    return p;
}

}  // namespace

int so_prof_is_on() { return prof().on ? 1 : 0; }

int so_prof_open(const char* label, cudaStream_t st) {
    ProfState& p = prof();
    if (!p.on || p.marks.size() >= (1u << 20)) return -1;
    ProfMark m;
    m.label = label;
    if (cudaEventCreate(&m.e0) != cudaSuccess || cudaEventCreate(&m.e1) != cudaSuccess) return -1;
    cudaEventRecord(m.e0, st);
    p.marks.push_back(m);
    return (int)p.marks.size() - 1;
}

void so_prof_close(int id, cudaStream_t st) {
    ProfState& p = prof();
    if (id < 0 || id >= (int)p.marks.size()) return;
    cudaEventRecord(p.marks[id].e1, st);
}

extern "C" int so_prof_enable(int32_t on) {
    ProfState& p = prof();
    for (auto& m : p.marks) { cudaEventDestroy(m.e0); cudaEventDestroy(m.e1); }
    p.marks.clear();
    p.on = on != 0;
    return SO_OK;
}

extern "C" int64_t so_prof_report(char* buf, int64_t buf_bytes) {
    ProfState& p = prof();
    std::map<std::string, std::pair<double, long>> acc;
    std::vector<std::string> order;
    for (auto& m : p.marks) {
        if (cudaEventSynchronize(m.e1) != cudaSuccess) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, m.e0, m.e1) != cudaSuccess) continue;
        if (!acc.count(m.label)) order.push_back(m.label);
        acc[m.label].first += ms;
        acc[m.label].second += 1;
    }
    std::string text;
    char line[256];
    for (auto& l : order) {
        snprintf(line, sizeof line, "%s\t%ld\t%.6f\n", l.c_str(), acc[l].second, acc[l].first);
        text += line;
    }
    if (buf && buf_bytes > 0) {
        const size_t ncopy = text.size() < (size_t)buf_bytes - 1 ? text.size() : (size_t)buf_bytes - 1;
        memcpy(buf, text.data(), ncopy);
        buf[ncopy] = 0;
    }
    return (int64_t)text.size() + 1;
}

extern "C" int so_peak_kernel(int32_t kind, int64_t iters, int32_t ctas_per_sm, float* out, double* ops_host, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (iters < 1 || ctas_per_sm < 1 || !out) return SO_ERR_BAD_ARG;
    const unsigned grid = (unsigned)(so_num_sms() * ctas_per_sm);
    const double threads = (double)grid * kPeakThreads;
    double per_thread_iter = 0.0;
    switch (kind) {
    case 0: peak_ffma_kernel<<<grid, kPeakThreads, 0, st>>>(iters, out); per_thread_iter = 16.0; break;      // FMAs
    case 1: peak_ffma2_kernel<<<grid, kPeakThreads, 0, st>>>(iters, out); per_thread_iter = 32.0; break;     // FMAs
    case 2: peak_ex2_kernel<<<grid, kPeakThreads, 0, st>>>(iters, out); per_thread_iter = 16.0; break;       // exponentials
    case 3: peak_dust_kernel<false><<<grid, kPeakThreads, 0, st>>>(iters, out); per_thread_iter = 16.0; break;
    case 4: peak_dust_kernel<true><<<grid, kPeakThreads, 0, st>>>(iters, out); per_thread_iter = 16.0; break;
    default: return SO_ERR_BAD_ARG;
    }
    SO_LAUNCH_CHECK();
    if (ops_host) *ops_host = threads * (double)iters * per_thread_iter;
    return SO_OK;
}

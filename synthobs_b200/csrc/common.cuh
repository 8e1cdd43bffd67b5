// Shared helpers for the synthobs_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/synthobs_b200.h"

#define SO_CUDA(call)                                  \
    do {                                               \
        cudaError_t _e = (call);                       \
        if (_e != cudaSuccess) return (int)_e;         \
    } while (0)

#define SO_LAUNCH_CHECK()                              \
    do {                                               \
        cudaError_t _e = cudaGetLastError();           \
        if (_e != cudaSuccess) return (int)_e;         \
    } while (0)

// optional per-kernel timing (peaks.cu; bench.py switches it on for one profiling pass): a labelled scope records
// CUDA events on the launching stream before and after the launches it encloses
int so_prof_is_on();
int so_prof_open(const char* label, cudaStream_t st);
void so_prof_close(int id, cudaStream_t st);
struct SoProfScope {
    int id;
    cudaStream_t st;
    SoProfScope(const char* label, cudaStream_t s) : id(-1), st(s) { if (so_prof_is_on()) id = so_prof_open(label, s); }
    ~SoProfScope() { if (id >= 0) so_prof_close(id, st); }
};
#define SO_PROF_CAT2(a, b) a##b
#define SO_PROF_CAT(a, b) SO_PROF_CAT2(a, b)
#define SO_PROF(label, st) SoProfScope SO_PROF_CAT(so_prof_scope_, __LINE__)(label, st)

static inline int64_t so_align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// bump allocator over a caller-provided workspace
struct SoArena {
    char* base;
    int64_t size;
    int64_t used;
    __host__ SoArena(void* p, int64_t n) : base((char*)p), size(n), used(0) {}
    template <typename T>
    __host__ T* take(int64_t count) {
        int64_t bytes = so_align_up(count * (int64_t)sizeof(T), 256);
        if (base == nullptr) { used += bytes; return nullptr; }
        if (used + bytes > size) { used += bytes; return nullptr; }
        T* r = (T*)(base + used);
        used += bytes;
        return r;
    }
    __host__ bool ok() const { return base == nullptr || used <= size; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming (read-once) global loads: bypass L1 allocation
__device__ __forceinline__ double ldg_stream(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldg_stream(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

static inline int so_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

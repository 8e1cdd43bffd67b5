// Library identification, device probe and the FP32 (FMA-pipe) GEMM used to validate the
// tensor-core kernel on the device.
#include "common.cuh"

namespace {

// C[M,N] = A[M,K] * B[N,K]^T, 64x64 tile, 256 threads, 4x4 per thread, K step 16
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int64_t lda,
                                                       const float* __restrict__ B, int64_t ldb,
                                                       float* __restrict__ C, int64_t ldc,
                                                       int64_t M, int64_t N, int64_t K) {
    __shared__ float sA[16][64 + 4];
    __shared__ float sB[16][64 + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.y * 64, n0 = (int64_t)blockIdx.x * 64;
    const int ty = tid / 16, tx = tid % 16;
    float acc[4][4] = {};
    for (int64_t k0 = 0; k0 < K; k0 += 16) {
        for (int e = tid; e < 64 * 16; e += 256) {
            const int r = e / 16, kk = e % 16;
            const int64_t m = m0 + r, n = n0 + r, k = k0 + kk;
            sA[kk][r] = (m < M && k < K) ? A[m * lda + k] : 0.f;
            sB[kk][r] = (n < N && k < K) ? B[n * ldb + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const int64_t m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) C[m * ldc + n] = acc[i][j];
        }
}

// ---- np.median of float64 arrays by 8-bit radix selection (no sort) --------------------------------
// order-preserving map double -> uint64 (negative values: all bits flipped, others: sign bit set)
__device__ __forceinline__ unsigned long long f64_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

struct SelectState {
    unsigned long long prefix;      // high bytes of the selected key found so far
    unsigned long long rank;        // rank of the wanted element among the keys that match the prefix
    unsigned long long below_max;   // largest key below the selected one (0 if none)
    unsigned int ticket;            // blocks of the current pass that have added their histogram
    unsigned int hist[256];
};

// pass p: histogram of byte (7 - p) over the keys whose p high bytes equal the prefix; the block that adds
// its histogram last (ticket counter) picks the byte value whose bin holds the wanted rank, descends into
// it and clears the histogram for the next pass
__global__ void __launch_bounds__(512) select_hist_kernel(const double* __restrict__ x, int64_t n, int64_t stride,
                                                          SelectState* __restrict__ st, int pass) {
    __shared__ unsigned int s_h[256];
    __shared__ bool s_last;
    const double* xa = x + (size_t)blockIdx.y * stride;
    SelectState* sa = st + blockIdx.y;
    if (threadIdx.x < 256) s_h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = sa->prefix;
    const int hs = 64 - 8 * pass;                   // bits above the byte of this pass
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = f64_key(xa[i]);
        if (pass == 0 || (k >> hs) == (prefix >> hs)) atomicAdd(&s_h[(k >> (hs - 8)) & 255], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 256 && s_h[threadIdx.x]) atomicAdd(&sa->hist[threadIdx.x], s_h[threadIdx.x]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&sa->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < 256) s_h[threadIdx.x] = *(volatile unsigned int*)&sa->hist[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long r = sa->rank, acc = 0;
        int b = 0;
        for (; b < 255; ++b) {
            if (acc + s_h[b] > r) break;
            acc += s_h[b];
        }
        sa->rank = r - acc;
        sa->prefix = prefix | ((unsigned long long)b << (56 - 8 * pass));
        sa->ticket = 0;
    }
    if (threadIdx.x < 256) sa->hist[threadIdx.x] = 0;
}

// largest key strictly below the selected one (the lower middle element when n is even and distinct)
__global__ void __launch_bounds__(512) select_below_kernel(const double* __restrict__ x, int64_t n, int64_t stride,
                                                           SelectState* __restrict__ st) {
    const double* xa = x + (size_t)blockIdx.y * stride;
    SelectState* sa = st + blockIdx.y;
    const unsigned long long sel = sa->prefix;
    unsigned long long m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = f64_key(xa[i]);
        if (k < sel && k > m) m = k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&sa->below_max, m);
}

// np.median: element n//2 for odd n, mean of elements n//2 - 1 and n//2 for even n.  After the
// eight passes sa->rank is the rank of the selected value among its duplicates: when it is > 0
// (or n is odd) the lower middle element equals the upper one
__global__ void select_finish_kernel(const SelectState* __restrict__ st, int64_t n, double* __restrict__ out) {
    const SelectState* sa = st + threadIdx.x;
    const double hi = key_f64(sa->prefix);
    double lo = hi;
    if ((n & 1) == 0 && sa->rank == 0) lo = key_f64(sa->below_max);
    out[threadIdx.x] = (n & 1) ? hi : __dmul_rn(__dadd_rn(lo, hi), 0.5);
}

__global__ void select_init_kernel(SelectState* __restrict__ st, int64_t n) {
    SelectState* sa = st + blockIdx.x;
    if (threadIdx.x == 0) { sa->prefix = 0; sa->rank = (unsigned long long)(n / 2); sa->below_max = 0; sa->ticket = 0; }
    sa->hist[threadIdx.x] = 0;
}

// np.median per CSR segment: one CTA walks its segment nine times (eight digit passes + the lower middle)
__global__ void __launch_bounds__(256) median_segmented_kernel(const double* __restrict__ x,
                                                               const int64_t* __restrict__ seg, double* __restrict__ out) {
    __shared__ unsigned int s_h[256];
    __shared__ unsigned long long s_prefix, s_rank, s_below;
    const int64_t a = seg[blockIdx.x], b = seg[blockIdx.x + 1];
    const int64_t n = b - a;
    if (n <= 0) {
        if (threadIdx.x == 0) out[blockIdx.x] = __longlong_as_double(0x7ff8000000000000LL);   // NaN, like np.median([])
        return;
    }
    if (threadIdx.x == 0) { s_prefix = 0; s_rank = (unsigned long long)(n / 2); s_below = 0; }
    for (int pass = 0; pass < 8; ++pass) {
        s_h[threadIdx.x] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const int hs = 64 - 8 * pass;
        for (int64_t i = a + threadIdx.x; i < b; i += 256) {
            const unsigned long long k = f64_key(x[i]);
            if (pass == 0 || (k >> hs) == (prefix >> hs)) atomicAdd(&s_h[(k >> (hs - 8)) & 255], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long r = s_rank, acc = 0;
            int bin = 0;
            for (; bin < 255; ++bin) {
                if (acc + s_h[bin] > r) break;
                acc += s_h[bin];
            }
            s_rank = r - acc;
            s_prefix = prefix | ((unsigned long long)bin << (56 - 8 * pass));
        }
        __syncthreads();
    }
    const unsigned long long sel = s_prefix;
    unsigned long long m = 0;
    for (int64_t i = a + threadIdx.x; i < b; i += 256) {
        const unsigned long long k = f64_key(x[i]);
        if (k < sel && k > m) m = k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&s_below, m);
    __syncthreads();
    if (threadIdx.x == 0) {
        const double hi = key_f64(sel);
        double lo = hi;
        if ((n & 1) == 0 && s_rank == 0) lo = key_f64(s_below);
        out[blockIdx.x] = (n & 1) ? hi : __dmul_rn(__dadd_rn(lo, hi), 0.5);
    }
}

// x[i] += shift - med[segment of i]   (recentring of every object of a batch, images.py:183-187)
__global__ void __launch_bounds__(256) recentre_segmented_kernel(double* __restrict__ x, const int64_t* __restrict__ seg,
                                                                 const double* __restrict__ med, double shift) {
    const int64_t a = seg[blockIdx.x], b = seg[blockIdx.x + 1];
    const double m = med[blockIdx.x];
    for (int64_t i = a + threadIdx.x; i < b; i += 256) x[i] = __dadd_rn(__dsub_rn(x[i], m), shift);
}

}  // namespace

extern "C" int so_median_segmented(const double* x, const int64_t* seg_offsets, int64_t n_seg, double* out, void* stream) {
    if (!x || !seg_offsets || !out || n_seg < 1 || n_seg >= ((int64_t)1 << 31)) return SO_ERR_BAD_ARG;
    median_segmented_kernel<<<(unsigned)n_seg, 256, 0, (cudaStream_t)stream>>>(x, seg_offsets, out);
    SO_LAUNCH_CHECK();
    return SO_OK;
}

extern "C" int so_recentre_segmented(double* x, const int64_t* seg_offsets, int64_t n_seg, const double* med, double shift,
                                     void* stream) {
    if (!x || !seg_offsets || !med || n_seg < 1 || n_seg >= ((int64_t)1 << 31)) return SO_ERR_BAD_ARG;
    recentre_segmented_kernel<<<(unsigned)n_seg, 256, 0, (cudaStream_t)stream>>>(x, seg_offsets, med, shift);
    SO_LAUNCH_CHECK();
    return SO_OK;
}

extern "C" int64_t so_median_workspace_bytes(int32_t n_arrays) {
    return so_align_up((int64_t)n_arrays * (int64_t)sizeof(SelectState), 256) + 256;
}

extern "C" int so_median_f64(const double* x, int64_t n, int32_t n_arrays, int64_t stride, double* out,
                             void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !out || n < 1 || n_arrays < 1 || n_arrays > 1024 || stride < n) return SO_ERR_BAD_ARG;
    SoArena a(workspace, workspace_bytes);
    SelectState* state = a.take<SelectState>(n_arrays);
    if (!workspace || !a.ok()) return SO_ERR_WORKSPACE;
    int64_t blocks = (n + 512 * 8 - 1) / (512 * 8);
    if (blocks > 2 * so_num_sms()) blocks = 2 * so_num_sms();
    if (blocks < 1) blocks = 1;
    dim3 grid((unsigned)blocks, (unsigned)n_arrays);
    select_init_kernel<<<n_arrays, 256, 0, st>>>(state, n);
    SO_LAUNCH_CHECK();
    for (int pass = 0; pass < 8; ++pass) {
        select_hist_kernel<<<grid, 512, 0, st>>>(x, n, stride, state, pass);
        SO_LAUNCH_CHECK();
    }
    select_below_kernel<<<grid, 512, 0, st>>>(x, n, stride, state);
    SO_LAUNCH_CHECK();
    select_finish_kernel<<<1, n_arrays, 0, st>>>(state, n, out);
    SO_LAUNCH_CHECK();
    return SO_OK;
}

extern "C" const char* so_version(void) { return "synthobs_b200 0.1 (sm_100a, " __DATE__ ")"; }

extern "C" int so_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) return 0;
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    return major == 10 ? 1 : 0;
}

extern "C" int so_gemm_f32(const float* A, int64_t lda, const float* B, int64_t ldb,
                           float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, void* stream) {
    if (M < 1 || N < 1 || K < 1 || !A || !B || !C) return SO_ERR_BAD_ARG;
    dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64));
    gemm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, B, ldb, C, ldc, M, N, K);
    SO_LAUNCH_CHECK();
    return SO_OK;
}

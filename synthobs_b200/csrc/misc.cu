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

// ---- np.median of float64 arrays by radix selection (no sort) ---------------------------------------
// order-preserving map double -> uint64 (negative values: all bits flipped, others: sign bit set)
__device__ __forceinline__ unsigned long long f64_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// Selection state of one array.  The first two passes histogram 11-bit digits of ALL keys (sign +
// exponent, then the top mantissa bits), the third pass copies the few keys that share those 22
// bits into a buffer, and one CTA finishes the remaining 42 bits on that buffer.
constexpr int kSelBits = 11;
constexpr int kSelBins = 1 << kSelBits;
constexpr int kSelThreads = 512;
constexpr int kSelTailThreads = 1024;

struct SelectState {
    unsigned long long prefix;      // high digits of the selected key found so far
    unsigned long long rank;        // rank of the wanted element among the keys that match the prefix
    unsigned long long below_max;   // largest key below the selected prefix group (0 if none)
    unsigned long long count;       // keys copied to the buffer
    unsigned int ticket;            // blocks of the current pass that have added their histogram
    unsigned int pad;
    unsigned int hist[kSelBins];
};

// All threads of the CTA: the bin of s_h[0 .. nbins) that holds rank r -> (*bin, *rank inside the bin)
// written by the one thread that owns it.  nbins <= 4 * blockDim.x.
__device__ __forceinline__ void select_pick(const unsigned int* s_h, int nbins, unsigned long long r,
                                            unsigned int* s_scan, int* bin, unsigned long long* rank_in) {
    const int t = threadIdx.x, per = (nbins + blockDim.x - 1) / blockDim.x;
    unsigned int mine = 0;
    for (int j = 0; j < per; ++j) { const int bi = t * per + j; if (bi < nbins) mine += s_h[bi]; }
    // inclusive scan of the per-thread sums (warp shuffles + one shared pass over the warp totals)
    unsigned int x = mine;
    const int lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_scan[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned int w = (lane < (int)(blockDim.x >> 5)) ? s_scan[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        s_scan[lane] = w;
    }
    __syncthreads();
    unsigned long long before = (unsigned long long)(x - mine) + (warp > 0 ? s_scan[warp - 1] : 0);
    if (r >= before && r < before + mine) {
        for (int j = 0; j < per; ++j) {
            const int bi = t * per + j;
            const unsigned int c = s_h[bi];
            if (r < before + c) { *bin = bi; *rank_in = r - before; break; }
            before += c;
        }
    }
    __syncthreads();
}

// pass p (0 or 1): histogram of digit p over the keys whose higher digits equal the prefix; the block that
// adds its histogram last (ticket counter) picks the digit whose bin holds the wanted rank and clears the
// histogram for the next pass
__global__ void __launch_bounds__(kSelThreads) select_hist_kernel(const double* __restrict__ x, int64_t n, int64_t stride,
                                                                  SelectState* __restrict__ st, int pass) {
    __shared__ unsigned int s_h[kSelBins];
    __shared__ unsigned int s_scan[32];
    __shared__ bool s_last;
    __shared__ int s_bin;
    __shared__ unsigned long long s_rank;
    const double* xa = x + (size_t)blockIdx.y * stride;
    SelectState* sa = st + blockIdx.y;
    for (int i = threadIdx.x; i < kSelBins; i += kSelThreads) s_h[i] = 0;
    __syncthreads();
    const unsigned long long prefix = sa->prefix;
    const int hs = 64 - kSelBits * pass;            // bits above the digit of this pass
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = f64_key(xa[i]);
        if (pass == 0 || (k >> hs) == (prefix >> hs)) atomicAdd(&s_h[(k >> (hs - kSelBits)) & (kSelBins - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSelBins; i += kSelThreads)
        if (s_h[i]) atomicAdd(&sa->hist[i], s_h[i]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&sa->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = threadIdx.x; i < kSelBins; i += kSelThreads) {
        s_h[i] = *(volatile unsigned int*)&sa->hist[i];
        sa->hist[i] = 0;
    }
    __syncthreads();
    select_pick(s_h, kSelBins, sa->rank, s_scan, &s_bin, &s_rank);
    if (threadIdx.x == 0) {
        sa->rank = s_rank;
        sa->prefix = prefix | ((unsigned long long)s_bin << (hs - kSelBits));
        sa->ticket = 0;
    }
}

// pass 2: the keys that share the 22 selected bits go to the buffer (any order); the largest key below
// that group is the lower middle element when the selected key turns out to be the smallest of its group
__global__ void __launch_bounds__(kSelThreads) select_compact_kernel(const double* __restrict__ x, int64_t n, int64_t stride,
                                                                     SelectState* __restrict__ st,
                                                                     unsigned long long* __restrict__ buf) {
    const double* xa = x + (size_t)blockIdx.y * stride;
    SelectState* sa = st + blockIdx.y;
    unsigned long long* out = buf + (size_t)blockIdx.y * n;
    constexpr int hs = 64 - 2 * kSelBits;
    const unsigned long long want = sa->prefix >> hs;
    unsigned long long below = 0;
    const int lane = threadIdx.x & 31;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (n + step - 1) / step * step;            // whole warps stay in the loop together
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += step) {
        unsigned long long k = 0;
        bool hit = false;
        if (i < n) {
            k = f64_key(xa[i]);
            const unsigned long long top = k >> hs;
            hit = (top == want);
            if (top < want && k > below) below = k;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal) {
            unsigned long long base = 0;
            if (lane == (__ffs(bal) - 1)) base = atomicAdd(&sa->count, (unsigned long long)__popc(bal));
            base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
            if (hit) out[base + __popc(bal & ((1u << lane) - 1u))] = k;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, below, o);
        below = t > below ? t : below;
    }
    if (lane == 0 && below) atomicMax(&sa->below_max, below);
}

// the remaining 42 bits on the buffered keys (one CTA per array), then np.median: element n//2 for odd n,
// mean of elements n//2 - 1 and n//2 for even n.  rank > 0 after the last digit = the selected value has
// duplicates below the wanted rank: then the lower middle element equals the upper one
__global__ void __launch_bounds__(kSelTailThreads) select_tail_kernel(SelectState* __restrict__ st, int64_t n,
                                                                       const unsigned long long* __restrict__ buf,
                                                                       double* __restrict__ out, double* __restrict__ mid) {
    __shared__ unsigned int s_h[kSelBins];
    __shared__ unsigned int s_scan[32];
    __shared__ int s_bin;
    __shared__ unsigned long long s_rank, s_prefix, s_below;
    SelectState* sa = st + blockIdx.x;
    const unsigned long long* keys = buf + (size_t)blockIdx.x * n;
    const int64_t m = (int64_t)sa->count;
    if (threadIdx.x == 0) { s_rank = sa->rank; s_prefix = sa->prefix; s_below = sa->below_max; }
    __syncthreads();
    int hs = 64 - 2 * kSelBits;                      // bits above the next digit: 42 left
    while (hs > 0) {
        const int bits = hs >= kSelBits ? kSelBits : hs;
        for (int i = threadIdx.x; i < kSelBins; i += kSelTailThreads) s_h[i] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        for (int64_t i = threadIdx.x; i < m; i += kSelTailThreads) {
            const unsigned long long k = keys[i];
            if ((k >> hs) == (prefix >> hs)) atomicAdd(&s_h[(k >> (hs - bits)) & ((1u << bits) - 1u)], 1u);
        }
        __syncthreads();
        select_pick(s_h, 1 << bits, s_rank, s_scan, &s_bin, &s_rank);
        if (threadIdx.x == 0) s_prefix = prefix | ((unsigned long long)s_bin << (hs - bits));
        __syncthreads();
        hs -= bits;
    }
    const unsigned long long sel = s_prefix;
    unsigned long long below = 0;
    for (int64_t i = threadIdx.x; i < m; i += kSelTailThreads) {
        const unsigned long long k = keys[i];
        if (k < sel && k > below) below = k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, below, o);
        below = t > below ? t : below;
    }
    if ((threadIdx.x & 31) == 0 && below) atomicMax(&s_below, below);
    __syncthreads();
    if (threadIdx.x == 0) {
        const double hi = key_f64(sel);
        double lo = hi;
        if ((n & 1) == 0 && s_rank == 0) lo = key_f64(s_below);
        out[blockIdx.x] = (n & 1) ? hi : __dmul_rn(__dadd_rn(lo, hi), 0.5);
        if (mid) { mid[2 * blockIdx.x] = (n & 1) ? hi : lo; mid[2 * blockIdx.x + 1] = hi; }
    }
}

__global__ void __launch_bounds__(kSelThreads) select_init_kernel(SelectState* __restrict__ st, int64_t n) {
    SelectState* sa = st + blockIdx.x;
    if (threadIdx.x == 0) {
        sa->prefix = 0; sa->rank = (unsigned long long)(n / 2); sa->below_max = 0; sa->count = 0; sa->ticket = 0;
    }
    for (int i = threadIdx.x; i < kSelBins; i += kSelThreads) sa->hist[i] = 0;
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

// The reference's images.particle recentres the caller's X, Y once per filter (images.py:183-187 inside the loop of
// :277-284): X_f = (X_{f-1} - median(X_{f-1})) + offset_f.  Every step is a monotone map, so the two middle elements
// of X_f are the images of the two middle elements of X: the whole chain of medians follows from (lo, hi) of the
// original array by scalar arithmetic, and all steps are applied to an element in one pass.
constexpr int kChainMax = 32;
struct ChainParams {
    const double* in[2];
    const double* mid;              // [4]: x_lo, x_hi, y_lo, y_hi
    double off[2][kChainMax];
    int n_steps, n_out, odd;
    int out_step[8];                // 1-based, ascending
    double* out[2][8];
    int64_t n;
};

__global__ void __launch_bounds__(256) recentre_chain_kernel(const ChainParams p) {
    const int axis = blockIdx.y;
    double lo = p.mid[2 * axis], hi = p.mid[2 * axis + 1];
    double med[kChainMax];
#pragma unroll 1
    for (int s = 0; s < p.n_steps; ++s) {
        const double m = p.odd ? hi : __dmul_rn(__dadd_rn(lo, hi), 0.5);
        med[s] = m;
        lo = __dadd_rn(__dsub_rn(lo, m), p.off[axis][s]);
        hi = __dadd_rn(__dsub_rn(hi, m), p.off[axis][s]);
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = p.in[axis][i];
        int j = 0;
#pragma unroll 1
        for (int s = 0; s < p.n_steps; ++s) {
            v = __dadd_rn(__dsub_rn(v, med[s]), p.off[axis][s]);       // A -= np.median(A); A += offset
            if (j < p.n_out && p.out_step[j] == s + 1) { p.out[axis][j][i] = v; ++j; }
        }
    }
}

}  // namespace

extern "C" int so_recentre_chain(const double* x, const double* y, int64_t n, const double* mid, int32_t n_steps,
                                 const double* off_x, const double* off_y, int32_t n_out, const int32_t* out_step,
                                 double* const* out_x, double* const* out_y, void* stream) {
    if (!x || !y || !mid || n < 0 || n_steps < 1 || n_steps > kChainMax || n_out < 1 || n_out > 8 || !off_x || !off_y ||
        !out_step || !out_x || !out_y)
        return SO_ERR_BAD_ARG;
    if (n == 0) return SO_OK;
    ChainParams p{};
    p.in[0] = x; p.in[1] = y; p.mid = mid; p.n_steps = n_steps; p.n_out = n_out; p.odd = (int)(n & 1); p.n = n;
    for (int s = 0; s < n_steps; ++s) { p.off[0][s] = off_x[s]; p.off[1][s] = off_y[s]; }
    for (int j = 0; j < n_out; ++j) {
        if (out_step[j] < 1 || out_step[j] > n_steps || (j > 0 && out_step[j] <= out_step[j - 1]) || !out_x[j] || !out_y[j])
            return SO_ERR_BAD_ARG;
        p.out_step[j] = out_step[j]; p.out[0][j] = out_x[j]; p.out[1][j] = out_y[j];
    }
    int64_t blocks = (n + 255) / 256;
    if (blocks > 8 * so_num_sms()) blocks = 8 * so_num_sms();
    { SO_PROF("recentre_chain_kernel", (cudaStream_t)stream); recentre_chain_kernel<<<dim3((unsigned)blocks, 2), 256, 0, (cudaStream_t)stream>>>(p); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

extern "C" int so_median_segmented(const double* x, const int64_t* seg_offsets, int64_t n_seg, double* out, void* stream) {
    if (!x || !seg_offsets || !out || n_seg < 1 || n_seg >= ((int64_t)1 << 31)) return SO_ERR_BAD_ARG;
    { SO_PROF("median_segmented_kernel", (cudaStream_t)stream); median_segmented_kernel<<<(unsigned)n_seg, 256, 0, (cudaStream_t)stream>>>(x, seg_offsets, out); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

extern "C" int so_recentre_segmented(double* x, const int64_t* seg_offsets, int64_t n_seg, const double* med, double shift,
                                     void* stream) {
    if (!x || !seg_offsets || !med || n_seg < 1 || n_seg >= ((int64_t)1 << 31)) return SO_ERR_BAD_ARG;
    { SO_PROF("recentre_segmented_kernel", (cudaStream_t)stream); recentre_segmented_kernel<<<(unsigned)n_seg, 256, 0, (cudaStream_t)stream>>>(x, seg_offsets, med, shift); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

extern "C" int64_t so_median_workspace_bytes(int64_t n, int32_t n_arrays) {
    return so_align_up((int64_t)n_arrays * (int64_t)sizeof(SelectState), 256) +
           so_align_up((int64_t)n_arrays * n * (int64_t)sizeof(unsigned long long), 256) + 512;
}

extern "C" int so_median_f64(const double* x, int64_t n, int32_t n_arrays, int64_t stride, double* out,
                             void* workspace, int64_t workspace_bytes, void* stream) {
    return so_median_f64_mid(x, n, n_arrays, stride, out, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int so_median_f64_mid(const double* x, int64_t n, int32_t n_arrays, int64_t stride, double* out, double* mid,
                                 void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !out || n < 1 || n_arrays < 1 || n_arrays > 1024 || stride < n) return SO_ERR_BAD_ARG;
    SoArena a(workspace, workspace_bytes);
    SelectState* state = a.take<SelectState>(n_arrays);
    unsigned long long* buf = a.take<unsigned long long>((size_t)n_arrays * n);
    if (!workspace || !a.ok()) return SO_ERR_WORKSPACE;
    int64_t blocks = (n + kSelThreads * 8 - 1) / (kSelThreads * 8);
    if (blocks > 2 * so_num_sms()) blocks = 2 * so_num_sms();
    if (blocks < 1) blocks = 1;
    dim3 grid((unsigned)blocks, (unsigned)n_arrays);
    { SO_PROF("select_init_kernel", st); select_init_kernel<<<n_arrays, kSelThreads, 0, st>>>(state, n); }
    SO_LAUNCH_CHECK();
    for (int pass = 0; pass < 2; ++pass) {
        { SO_PROF("select_hist_kernel", st); select_hist_kernel<<<grid, kSelThreads, 0, st>>>(x, n, stride, state, pass); }
        SO_LAUNCH_CHECK();
    }
    { SO_PROF("select_compact_kernel", st); select_compact_kernel<<<grid, kSelThreads, 0, st>>>(x, n, stride, state, buf); }
    SO_LAUNCH_CHECK();
    { SO_PROF("select_tail_kernel", st); select_tail_kernel<<<n_arrays, kSelTailThreads, 0, st>>>(state, n, buf, out, mid); }
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
    { SO_PROF("gemm_f32_kernel", (cudaStream_t)stream); gemm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, B, ldb, C, ldc, M, N, K); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

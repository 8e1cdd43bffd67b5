// Particle shares of ONE huge image over several GPUs (BASELINE config 5; no upstream equivalent: the reference images
// one object in one Python loop, synthobs/morph/images.py:83-97).
//
// Every rank holds all positions, but a rank only needs the search tree AROUND its own share.  so_share_plan splits the
// selected particles into n_parts spatially compact shares of (almost) equal size and tells each particle whether it
// is OWN (this rank answers its k-NN query and deposits it) and/or LOCAL (own or halo: a member of this rank's search
// tree), such that the k-th neighbour distance computed among the LOCAL particles is EXACT for every OWN particle:
//
//   1. histogram of the selected particles on a coarse 2^cb x 2^cb grid over the image (so_share_hist: integer atomics
//      on counters -- never on pixels -- so it is order-independent; with several ranks each one counts 1/N of the
//      particles and the counts are summed with an all-reduce of the small table, identical on every rank);
//   2. shares = contiguous ranges of the coarse cells in Morton order with equal cumulative counts;
//   3. halo: a query in coarse cell c has at least k particles inside the smallest block B_m(c) of cells within
//      Chebyshev distance m of c that holds >= k particles (summed-area table), so its k-th neighbour is at most
//      diag(B_m) = sqrt(2) (2m + 1) cells away: every cell within ceil(sqrt(2) (2m + 1)) of c is marked LOCAL;
//   4. one pass over the particles writes the flags.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int kMaxCb = 9;

__host__ __device__ __forceinline__ uint32_t sh_spread(uint32_t v) {
    v &= 0xffff;
    v = (v | (v << 8)) & 0x00ff00ff;
    v = (v | (v << 4)) & 0x0f0f0f0f;
    v = (v | (v << 2)) & 0x33333333;
    v = (v | (v << 1)) & 0x55555555;
    return v;
}
__device__ __forceinline__ uint32_t sh_compact(uint32_t c) {
    c &= 0x55555555u;
    c = (c | (c >> 1)) & 0x33333333u; c = (c | (c >> 2)) & 0x0f0f0f0fu;
    c = (c | (c >> 4)) & 0x00ff00ffu; c = (c | (c >> 8)) & 0xffffu;
    return c;
}

// coarse cell (Morton index) of a position, -1 when the particle is not selected (images.py:47, strict)
__device__ __forceinline__ int coarse_cell(double x, double y, double hw, double inv_c, int nc) {
    if (!(fabs(x) < hw) || !(fabs(y) < hw)) return -1;
    int cx = (int)((x + hw) * inv_c), cy = (int)((y + hw) * inv_c);
    cx = max(0, min(nc - 1, cx));
    cy = max(0, min(nc - 1, cy));
    return (int)(sh_spread((uint32_t)cx) | (sh_spread((uint32_t)cy) << 1));
}

// counts per coarse cell in MORTON order (the order of the shares); integer atomics on counters, order-independent
__global__ void __launch_bounds__(512) share_hist_kernel(const double* __restrict__ X, const double* __restrict__ Y, int64_t n,
                                                         double hw, double inv_c, int nc, int32_t* __restrict__ hist) {
    // neighbours in memory are often neighbours in space (a simulation writes its particles clump by clump): the lanes
    // of a warp that hit the same cell add ONE count together
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (n + stride - 1) / stride;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t r = 0; r < rounds; ++r, i += stride) {
        const int c = i < n ? coarse_cell(X[i], Y[i], hw, inv_c, nc) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        if (c >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[c], __popc(peers));
    }
}

// Morton-ordered counts -> summed-area table sat[(y + 1) * (nc + 1) + (x + 1)] = sum of cells (<= x, <= y):
// row prefix sums (one thread per row), then column prefix sums (one thread per column, coalesced)
__global__ void __launch_bounds__(128) share_sat_rows_kernel(const int32_t* __restrict__ hist, int nc, unsigned long long* __restrict__ sat) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = nc + 1;
    if (y > nc) return;
    if (y == nc) { for (int x = 0; x < w; ++x) sat[x] = 0; return; }
    unsigned long long acc = 0;
    const uint32_t ybits = sh_spread((uint32_t)y) << 1;
    sat[(size_t)(y + 1) * w] = 0;
#pragma unroll 8
    for (int x = 0; x < nc; ++x) {
        acc += (unsigned long long)hist[sh_spread((uint32_t)x) | ybits];
        sat[(size_t)(y + 1) * w + x + 1] = acc;
    }
}
__global__ void __launch_bounds__(128) share_sat_cols_kernel(int nc, unsigned long long* __restrict__ sat) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = nc + 1;
    if (x >= nc) return;
    unsigned long long acc = 0;
#pragma unroll 8
    for (int y = 0; y < nc; ++y) {
        acc += sat[(size_t)(y + 1) * w + x + 1];
        sat[(size_t)(y + 1) * w + x + 1] = acc;
    }
}

__device__ __forceinline__ unsigned long long sat_block(const unsigned long long* sat, int w, int x0, int y0, int x1, int y1) {
    // cells [x0, x1] x [y0, y1], clipped by the caller
    return sat[(size_t)(y1 + 1) * w + x1 + 1] - sat[(size_t)y0 * w + x1 + 1] - sat[(size_t)(y1 + 1) * w + x0] + sat[(size_t)y0 * w + x0];
}

// cum = inclusive Morton-order prefix of the cells' WORK (share_weight_kernel).  A cell belongs to the share [lo, hi)
// (fractions of the total work) that holds the MIDDLE of the cell's interval: contiguous ranges of cells, the
// requested work up to one cell.
__device__ __forceinline__ bool cell_in_share(const unsigned long long* cum, int c, unsigned long long total, double lo, double hi) {
    const unsigned long long a = c ? cum[c - 1] : 0, b = cum[c];
    if (b == a) return false;                                 // empty cell: nobody's
    const double mid = 0.5 * ((double)a + (double)b) / (double)total;
    return mid >= lo && mid < hi;
}

// work of a cell = its particles, counted 2.5 times where the density is below dense_count per cell: there the k-th
// neighbour lies beyond the FWHM floor, so a query walks the tree instead of being settled by its sorted window
// (knn.cu) and its support spans many pixels
__global__ void __launch_bounds__(256) share_weight_kernel(const int32_t* __restrict__ hist, int cells, double dense_count,
                                                           unsigned long long* __restrict__ wgt) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const unsigned long long n = (unsigned long long)hist[c];
    wgt[c] = ((double)n < dense_count) ? 5 * n : 2 * n;
}

// every cell of share `part` marks the cells its queries may need (see header); mark: 1 = local, 2 = own
__global__ void __launch_bounds__(256) share_mark_kernel(const int32_t* __restrict__ hist, const unsigned long long* __restrict__ cum,
                                                         const unsigned long long* __restrict__ sat, int nc, int k, double lo,
                                                         double hi, unsigned char* __restrict__ mark) {
    // one WARP per cell: a sparse cell far out may need a block of thousands of cells around it
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= nc * nc) return;
    const unsigned long long total = cum[nc * nc - 1];
    if (total == 0 || hist[c] == 0 || !cell_in_share(cum, c, total, lo, hi)) return;
    const int cx = (int)sh_compact((uint32_t)c), cy = (int)sh_compact((uint32_t)c >> 1);
    const int w = nc + 1;
    int m = 0;
    if (sat[(size_t)w * w - 1] < (unsigned long long)k) {
        m = nc;                                               // fewer than k particles at all: every distance is inf
    } else {
        while (sat_block(sat, w, max(0, cx - m), max(0, cy - m), min(nc - 1, cx + m), min(nc - 1, cy + m)) < (unsigned long long)k) ++m;
    }
    // k-th neighbour within diag(B_m) = sqrt(2) (2m + 1) cell edges: Chebyshev ring M = ceil of that
    const int M = min(nc, (int)ceil(1.41421356237309515 * (double)(2 * m + 1)));
    const int x0 = max(0, cx - M), x1 = min(nc - 1, cx + M), y0 = max(0, cy - M), y1 = min(nc - 1, cy + M);
    const int bw = x1 - x0 + 1, cnt = bw * (y1 - y0 + 1);
    for (int i = lane; i < cnt; i += 32) {
        const int x = x0 + i % bw, y = y0 + i / bw;
        mark[sh_spread((uint32_t)x) | (sh_spread((uint32_t)y) << 1)] = 1;           // racing writers all store the same byte
    }
}

__global__ void __launch_bounds__(256) share_own_kernel(const unsigned long long* __restrict__ cum, int nc, double lo, double hi,
                                                        unsigned char* __restrict__ mark) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc * nc) return;
    const unsigned long long total = cum[nc * nc - 1];
    if (total && cell_in_share(cum, c, total, lo, hi)) mark[c] |= 3;
}

__global__ void __launch_bounds__(512) share_flag_kernel(const double* __restrict__ X, const double* __restrict__ Y, int64_t n,
                                                         double hw, double inv_c, int nc, const unsigned char* __restrict__ mark,
                                                         unsigned char* __restrict__ flags, unsigned long long* __restrict__ counts) {
    __shared__ unsigned int s_cnt[2];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    unsigned int loc = 0, own = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = coarse_cell(X[i], Y[i], hw, inv_c, nc);
        const unsigned char f = c >= 0 ? mark[c] : 0;
        flags[i] = f;
        loc += f & 1;
        own += (f >> 1) & 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { loc += __shfl_xor_sync(0xffffffffu, loc, o); own += __shfl_xor_sync(0xffffffffu, own, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt[0], loc); atomicAdd(&s_cnt[1], own); }
    __syncthreads();
    if (threadIdx.x < 2 && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

struct FlagSet {
    const unsigned char* flags;
    unsigned char mask;
    __host__ __device__ unsigned char operator()(int i) const { return (flags[i] & mask) ? 1 : 0; }
};

__global__ void __launch_bounds__(256) share_gather_kernel(const double* __restrict__ X, const double* __restrict__ Y,
                                                           const unsigned char* __restrict__ flags, const int32_t* __restrict__ idx,
                                                           int64_t m, double* __restrict__ oX, double* __restrict__ oY,
                                                           unsigned char* __restrict__ oF) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int32_t i = idx[j];
    if (oX) oX[j] = X[i];
    if (oY) oY[j] = Y[i];
    if (oF) oF[j] = flags[i];
}

struct ShareLayout {
    unsigned long long* cum;
    unsigned long long* wgt;
    unsigned long long* sat;
    unsigned char* mark;
    void* cub_tmp;
    size_t cub_bytes;
    int64_t total;
};

static ShareLayout carve_share(void* ws, int64_t bytes, int cb, bool* ok) {
    ShareLayout l;
    const int nc = 1 << cb;
    const int64_t cells = (int64_t)nc * nc;
    SoArena a(ws, bytes);
    l.cum = a.take<unsigned long long>(cells);
    l.wgt = a.take<unsigned long long>(cells);
    l.sat = a.take<unsigned long long>((int64_t)(nc + 1) * (nc + 1));
    l.mark = a.take<unsigned char>(cells);
    l.cub_bytes = 0;
    cub::DeviceScan::InclusiveSum((void*)nullptr, l.cub_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)cells);
    l.cub_tmp = a.take<char>((int64_t)l.cub_bytes);
    l.total = a.used;
    if (ok) *ok = a.ok();
    return l;
}

static size_t compact_tmp_bytes(int64_t n) {
    size_t b = 0;
    cub::TransformInputIterator<unsigned char, FlagSet, cub::CountingInputIterator<int32_t>> fl(cub::CountingInputIterator<int32_t>(0),
                                                                                                 FlagSet{nullptr, 1});
    cub::DeviceSelect::Flagged((void*)nullptr, b, cub::CountingInputIterator<int32_t>(0), fl, (int32_t*)nullptr, (int64_t*)nullptr,
                               (int)(n > 0 ? n : 1));
    return b;
}

}  // namespace

extern "C" int so_share_hist(const double* X, const double* Y, int64_t begin, int64_t end, double half_width, int32_t cb,
                             int32_t* hist, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (begin < 0 || end < begin || !(half_width > 0.0) || cb < 1 || cb > kMaxCb || !hist) return SO_ERR_BAD_ARG;
    const int nc = 1 << cb;
    SO_CUDA(cudaMemsetAsync(hist, 0, sizeof(int32_t) * (size_t)nc * nc, st));
    const int64_t n = end - begin;
    if (n == 0) return SO_OK;
    if (!X || !Y) return SO_ERR_BAD_ARG;
    int64_t blocks = (n + 511) / 512;
    if (blocks > 16 * so_num_sms()) blocks = 16 * so_num_sms();
    { SO_PROF("share_hist_kernel", st);
      share_hist_kernel<<<(unsigned)blocks, 512, 0, st>>>(X + begin, Y + begin, n, half_width, (double)nc / (2.0 * half_width), nc, hist); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

extern "C" int64_t so_share_plan_workspace_bytes(int32_t cb) {
    if (cb < 1 || cb > kMaxCb) return 0;
    return carve_share(nullptr, 0, cb, nullptr).total + 256;
}

extern "C" int so_share_plan(const double* X, const double* Y, int64_t n, double half_width, int32_t cb, const int32_t* hist,
                             int32_t k, double floor, double share_lo, double share_hi, uint8_t* flags, uint64_t* counts,
                             void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 0 || !(half_width > 0.0) || cb < 1 || cb > kMaxCb || k < 1 || !(share_lo >= 0.0) || !(share_hi > share_lo) || !counts || !hist)
        return SO_ERR_BAD_ARG;
    if (n > 0 && (!X || !Y || !flags)) return SO_ERR_BAD_ARG;
    bool ok = false;
    ShareLayout l = carve_share(workspace, workspace_bytes, cb, &ok);
    if (!ok || !workspace) return SO_ERR_WORKSPACE;
    const int nc = 1 << cb;
    const int cells = nc * nc;
    const double inv_c = (double)nc / (2.0 * half_width);
    SO_CUDA(cudaMemsetAsync(l.mark, 0, (size_t)cells, st));
    SO_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(uint64_t), st));
    if (n == 0) return SO_OK;
    // a cell is "dense" when k particles fit inside the floor radius: count >= k * cell area / (pi floor^2)
    const double cell = 2.0 * half_width / nc;
    const double dense_count = floor > 0.0 ? (double)k * cell * cell / (3.14159265358979312 * floor * floor) : 0.0;
    { SO_PROF("share_weight_kernel", st); share_weight_kernel<<<(cells + 255) / 256, 256, 0, st>>>(hist, cells, dense_count, l.wgt); }
    SO_LAUNCH_CHECK();
    size_t tmp = l.cub_bytes;
    SO_CUDA(cub::DeviceScan::InclusiveSum(l.cub_tmp, tmp, l.wgt, l.cum, cells, st));
    { SO_PROF("share_sat_kernels", st);
      share_sat_rows_kernel<<<(nc + 1 + 127) / 128, 128, 0, st>>>(hist, nc, l.sat);
      share_sat_cols_kernel<<<(nc + 127) / 128, 128, 0, st>>>(nc, l.sat); }
    SO_LAUNCH_CHECK();
    { SO_PROF("share_mark_kernel", st); share_mark_kernel<<<(unsigned)(((int64_t)cells * 32 + 255) / 256), 256, 0, st>>>(hist, l.cum, l.sat, nc, k, share_lo, share_hi, l.mark); }
    SO_LAUNCH_CHECK();
    { SO_PROF("share_own_kernel", st); share_own_kernel<<<(cells + 255) / 256, 256, 0, st>>>(l.cum, nc, share_lo, share_hi, l.mark); }
    SO_LAUNCH_CHECK();
    int64_t blocks = (n + 511) / 512;
    if (blocks > 16 * so_num_sms()) blocks = 16 * so_num_sms();
    { SO_PROF("share_flag_kernel", st); share_flag_kernel<<<(unsigned)blocks, 512, 0, st>>>(X, Y, n, half_width, inv_c, nc, l.mark, flags,
                                                                                          (unsigned long long*)counts); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

extern "C" int64_t so_share_compact_workspace_bytes(int64_t n) { return (int64_t)so_align_up((int64_t)compact_tmp_bytes(n), 256) + 256; }

extern "C" int so_share_compact(const double* X, const double* Y, const uint8_t* flags, int64_t n, int32_t bit_mask, int64_t m,
                                int32_t* out_index, double* out_X, double* out_Y, uint8_t* out_flags, int64_t* count,
                                void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 0 || m < 0 || m > n || !out_index || !count || bit_mask < 1 || bit_mask > 255) return SO_ERR_BAD_ARG;
    if (n >= ((int64_t)1 << 31)) return SO_ERR_TOO_LARGE;
    if (n == 0) { SO_CUDA(cudaMemsetAsync(count, 0, sizeof(int64_t), st)); return SO_OK; }
    if (!flags) return SO_ERR_BAD_ARG;
    size_t tmp = compact_tmp_bytes(n);
    if (!workspace || (int64_t)tmp > workspace_bytes) return SO_ERR_WORKSPACE;
    cub::TransformInputIterator<unsigned char, FlagSet, cub::CountingInputIterator<int32_t>> fl(
        cub::CountingInputIterator<int32_t>(0), FlagSet{flags, (unsigned char)bit_mask});
    // stable: the selected particles keep their relative order.  out_index must hold m entries and m must be the number of
    // flagged particles (so_share_plan's counts): DeviceSelect writes exactly that many.
    { SO_PROF("cub select share", st);
      SO_CUDA(cub::DeviceSelect::Flagged(workspace, tmp, cub::CountingInputIterator<int32_t>(0), fl, out_index, count, (int)n, st)); }
    if (m > 0 && (out_X || out_Y || out_flags)) {
        { SO_PROF("share_gather_kernel", st);
          share_gather_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(X, Y, flags, out_index, m, out_X, out_Y, out_flags); }
        SO_LAUNCH_CHECK();
    }
    return SO_OK;
}

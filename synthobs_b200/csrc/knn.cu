// K7 — exact k-th nearest-neighbour distance in 2-D (float64) on an implicit quadtree.
//
// Replaces cKDTree(...).query(points, k)[..., -1] of synthobs/morph/images.py:83-85: the distance
// from every selected particle to its k-th neighbour, the particle itself being the first
// (distance 0).
//
// Particles are sorted by the Morton code of their cell on a 2^L x 2^L grid over the image.  In that
// order every quadtree node (any level) is one contiguous range, found in O(1) from a single
// upper-bound table over the finest cells (E[c] = #particles with code <= c).  A query walks the tree
// depth-first, nearest child first, pruning nodes whose square is farther than the current k-th best,
// and scans leaves (<= kLeaf points or finest level).  Exact for any clustering: dense cores,
// sparse outskirts and isolated particles all cost O(log) node visits plus a few leaf scans.
// The squared distance is formed with separately rounded products and sum (no FMA contraction) and
// an IEEE sqrt, i.e. the same float64 arithmetic as a straightforward C loop.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int kLeaf = 32;
constexpr int kMaxLevel = 12;
constexpr int kStack = 3 * kMaxLevel + 4;

struct KnnLayout {
    int32_t *keys0, *vals0, *keys1, *vals1;
    double *xs, *ys;
    int32_t* E;                    // [4^L] upper bounds
    unsigned long long* counter;   // [0] selected count
    void* cub_tmp;
    int64_t cub_tmp_bytes;
    int L;
    int64_t total;
};

static int knn_level(int64_t n) {
    int L = 4;
    while (L < kMaxLevel && ((int64_t)1 << (2 * (L - 1))) < n) ++L;   // ~ceil(log2(sqrt(n))) + 1
    return L;
}

static KnnLayout carve_knn(void* ws, int64_t bytes, int64_t n, bool* ok) {
    KnnLayout l;
    l.L = knn_level(n);
    const int64_t ncell = (int64_t)1 << (2 * l.L);
    SoArena a(ws, bytes);
    l.keys0 = a.take<int32_t>(n);
    l.vals0 = a.take<int32_t>(n);
    l.keys1 = a.take<int32_t>(n);
    l.vals1 = a.take<int32_t>(n);
    l.xs = a.take<double>(n);
    l.ys = a.take<double>(n);
    l.E = a.take<int32_t>(ncell);
    l.counter = a.take<unsigned long long>(2);
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs((void*)nullptr, b1, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)(n > 0 ? n : 1), 0, 32);
    cub::DeviceScan::InclusiveScan((void*)nullptr, b2, (int32_t*)nullptr, (int32_t*)nullptr, cub::Max(), (int)ncell);
    l.cub_tmp_bytes = (int64_t)(b1 > b2 ? b1 : b2);
    l.cub_tmp = a.take<char>(l.cub_tmp_bytes);
    l.total = a.used;
    if (ok) *ok = a.ok();
    return l;
}

__host__ __device__ __forceinline__ uint32_t spread_bits(uint32_t v) {   // 12 bits -> even bit positions
    v &= 0xfff;
    v = (v | (v << 8)) & 0x00ff00ff;
    v = (v | (v << 4)) & 0x0f0f0f0f;
    v = (v | (v << 2)) & 0x33333333;
    v = (v | (v << 1)) & 0x55555555;
    return v;
}
__device__ __forceinline__ uint32_t morton(uint32_t cx, uint32_t cy) { return spread_bits(cx) | (spread_bits(cy) << 1); }

__global__ void knn_key_kernel(const double* __restrict__ X, const double* __restrict__ Y,
                               const int32_t* __restrict__ ix, int64_t n, double hw, double inv_h, int nc,
                               int32_t* __restrict__ keys, int32_t* __restrict__ vals,
                               unsigned long long* __restrict__ counter) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool sel = false;
    if (i < n) {
        sel = ix[i] >= 0;
        int key = nc * nc;   // sentinel: not selected, sorts to the end
        if (sel) {
            int cx = (int)((X[i] + hw) * inv_h), cy = (int)((Y[i] + hw) * inv_h);
            cx = max(0, min(nc - 1, cx));
            cy = max(0, min(nc - 1, cy));
            key = (int)morton((uint32_t)cx, (uint32_t)cy);
        }
        keys[i] = key;
        vals[i] = (int32_t)i;
    }
    const unsigned m = __ballot_sync(0xffffffffu, sel);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(counter, (unsigned long long)__popc(m));
}

__global__ void knn_gather_kernel(const double* __restrict__ X, const double* __restrict__ Y,
                                  const int32_t* __restrict__ vals, int64_t n,
                                  double* __restrict__ xs, double* __restrict__ ys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = vals[i];
    xs[i] = X[p];
    ys[i] = Y[p];
}

// E[c] = i + 1 at the last sorted element of code c (others stay 0); a max-scan then makes
// E[c] = #selected particles with code <= c
__global__ void knn_mark_kernel(const int32_t* __restrict__ keys, int64_t n, int ncell, int32_t* __restrict__ E) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = keys[i];
    if (k >= ncell) return;
    if (i + 1 == n || keys[i + 1] != k) E[k] = (int32_t)(i + 1);
}

__device__ __forceinline__ void knn_scan_range(const double* __restrict__ xs, const double* __restrict__ ys,
                                               int a, int b, double x, double y, double* best, int k, int stride) {
    for (int q = a; q < b; ++q) {
        const double dx = xs[q] - x, dy = ys[q] - y;
        const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (d2 < best[(k - 1) * stride]) {
            int j = k - 1;
            while (j > 0 && best[(j - 1) * stride] > d2) {
                best[j * stride] = best[(j - 1) * stride];
                --j;
            }
            best[j * stride] = d2;
        }
    }
}

// squared distance (in cell units) from (u, v) to the square of node (level, code); 0 inside
__device__ __forceinline__ double node_dmin2(int level, uint32_t code, int L, double u, double v) {
    // de-interleave the node coordinates
    uint32_t cx = code & 0x55555555u, cy = (code >> 1) & 0x55555555u;
    cx = (cx | (cx >> 1)) & 0x33333333u; cx = (cx | (cx >> 2)) & 0x0f0f0f0fu; cx = (cx | (cx >> 4)) & 0x00ff00ffu; cx = (cx | (cx >> 8)) & 0xffffu;
    cy = (cy | (cy >> 1)) & 0x33333333u; cy = (cy | (cy >> 2)) & 0x0f0f0f0fu; cy = (cy | (cy >> 4)) & 0x00ff00ffu; cy = (cy | (cy >> 8)) & 0xffffu;
    const double s = (double)(1u << (L - level));
    const double x0 = cx * s, y0 = cy * s;
    const double dx = fmax(fmax(x0 - u, u - (x0 + s)), 0.0);
    const double dy = fmax(fmax(y0 - v, v - (y0 + s)), 0.0);
    return dx * dx + dy * dy;
}

__global__ void knn_query_kernel(const double* __restrict__ xs, const double* __restrict__ ys,
                                 const int32_t* __restrict__ keys, const int32_t* __restrict__ vals, int64_t n,
                                 const int32_t* __restrict__ E, int L, double hw, double inv_h, int k,
                                 const unsigned long long* __restrict__ counter, double* __restrict__ dk) {
    extern __shared__ double s_best[];
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int pid = vals[s];
    const int key = keys[s];
    const int ncell = 1 << (2 * L);
    if (key >= ncell) { dk[pid] = 0.0; return; }
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    if ((unsigned long long)k > *counter) { dk[pid] = INF; return; }
    const int stride = blockDim.x;
    double* best = s_best + threadIdx.x;
    for (int j = 0; j < k; ++j) best[j * stride] = INF;
    const double x = xs[s], y = ys[s];
    const double u = (x + hw) * inv_h, v = (y + hw) * inv_h;     // cell units, as used for the keys
    const double h2 = 1.0 / (inv_h * inv_h) * (1.0 - 1e-9);      // cell units^2 -> length^2, conservative

    uint32_t stack[kStack];       // (level << 26) | code
    int sp = 0;
    stack[sp++] = 0u;             // root: level 0, code 0
    while (sp > 0) {
        const uint32_t node = stack[--sp];
        const int level = (int)(node >> 26);
        const uint32_t code = node & 0x3ffffffu;
        const int sh = 2 * (L - level);
        const int a = (code == 0) ? 0 : E[(code << sh) - 1];
        const int b = E[((code + 1) << sh) - 1];
        if (a == b) continue;
        if (node_dmin2(level, code, L, u, v) * h2 >= best[(k - 1) * stride]) continue;
        if (b - a <= kLeaf || level == L) {
            knn_scan_range(xs, ys, a, b, x, y, best, k, stride);
            continue;
        }
        // children, pushed farthest first so the nearest is expanded next
        uint32_t cc[4];
        double dd[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            cc[c] = (code << 2) | c;
            dd[c] = node_dmin2(level + 1, cc[c], L, u, v);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3 - i; ++j)
                if (dd[j] < dd[j + 1]) {
                    const double td = dd[j]; dd[j] = dd[j + 1]; dd[j + 1] = td;
                    const uint32_t tc = cc[j]; cc[j] = cc[j + 1]; cc[j + 1] = tc;
                }
#pragma unroll
        for (int c = 0; c < 4; ++c) stack[sp++] = ((uint32_t)(level + 1) << 26) | cc[c];
    }
    dk[pid] = sqrt(best[(k - 1) * stride]);
}

}  // namespace

extern "C" int64_t so_knn_workspace_bytes(int64_t n, int32_t k) {
    (void)k;
    KnnLayout l = carve_knn(nullptr, 0, n, nullptr);
    return l.total + 256;
}

extern "C" int so_knn_kth_distance(const double* X, const double* Y, const int32_t* ix, int64_t n,
                                   int32_t k, double half_width, double* dk,
                                   void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return SO_OK;
    if (n < 0 || k < 1 || !(half_width > 0.0) || !X || !Y || !ix || !dk) return SO_ERR_BAD_ARG;
    if (n >= (int64_t)1 << 31) return SO_ERR_TOO_LARGE;
    bool ok = false;
    KnnLayout l = carve_knn(workspace, workspace_bytes, n, &ok);
    if (!ok || !workspace) return SO_ERR_WORKSPACE;
    int threads = 128;
    size_t smem = (size_t)k * threads * sizeof(double);
    if (smem > 200 * 1024) { threads = 32; smem = (size_t)k * threads * sizeof(double); }
    if (smem > 200 * 1024) return SO_ERR_UNSUPPORTED;
    const int L = l.L;
    const int nc = 1 << L;
    const int ncell = nc * nc;
    const double inv_h = nc / (2.0 * half_width);
    SO_CUDA(cudaMemsetAsync(l.counter, 0, 2 * sizeof(unsigned long long), st));
    SO_CUDA(cudaMemsetAsync(l.E, 0, sizeof(int32_t) * (size_t)ncell, st));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    knn_key_kernel<<<blocks, 256, 0, st>>>(X, Y, ix, n, half_width, inv_h, nc, l.keys0, l.vals0, l.counter);
    SO_LAUNCH_CHECK();
    size_t tmp = (size_t)l.cub_tmp_bytes;
    SO_CUDA(cub::DeviceRadixSort::SortPairs(l.cub_tmp, tmp, l.keys0, l.keys1, l.vals0, l.vals1, (int)n, 0, 2 * L + 1, st));
    knn_gather_kernel<<<blocks, 256, 0, st>>>(X, Y, l.vals1, n, l.xs, l.ys);
    SO_LAUNCH_CHECK();
    knn_mark_kernel<<<blocks, 256, 0, st>>>(l.keys1, n, ncell, l.E);
    SO_LAUNCH_CHECK();
    tmp = (size_t)l.cub_tmp_bytes;
    SO_CUDA(cub::DeviceScan::InclusiveScan(l.cub_tmp, tmp, l.E, l.E, cub::Max(), ncell, st));
    SO_CUDA(cudaFuncSetAttribute(knn_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_query_kernel<<<(unsigned)((n + threads - 1) / threads), threads, smem, st>>>(
        l.xs, l.ys, l.keys1, l.vals1, n, l.E, L, half_width, inv_h, k, l.counter, dk);
    SO_LAUNCH_CHECK();
    return SO_OK;
}

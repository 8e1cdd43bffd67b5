// K7 — exact k-th nearest-neighbour distance in 2-D (float64), uniform-grid ring search.
//
// Replaces cKDTree(...).query(points, k)[..., -1] of synthobs/morph/images.py:83-85:
// the distance from every selected particle to its k-th neighbour, the particle itself
// being the first (distance 0).  The squared distance is formed with separately rounded
// products and sum (no FMA contraction) and an IEEE sqrt, i.e. the same float64 arithmetic
// as a straightforward C loop.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

struct KnnLayout {
    int32_t *keys0, *vals0, *keys1, *vals1;
    double *xs, *ys;
    int32_t* cell_start;
    unsigned long long* counter;
    void* cub_tmp;
    int64_t cub_tmp_bytes;
    int nc;
    int64_t total;
};

static int knn_cells_per_axis(int64_t n) {
    double c = sqrt((double)n / 2.0);
    int nc = (int)c;
    if (nc < 1) nc = 1;
    if (nc > 2048) nc = 2048;
    return nc;
}

static KnnLayout carve_knn(void* ws, int64_t bytes, int64_t n, bool* ok) {
    KnnLayout l;
    l.nc = knn_cells_per_axis(n);
    SoArena a(ws, bytes);
    l.keys0 = a.take<int32_t>(n);
    l.vals0 = a.take<int32_t>(n);
    l.keys1 = a.take<int32_t>(n);
    l.vals1 = a.take<int32_t>(n);
    l.xs = a.take<double>(n);
    l.ys = a.take<double>(n);
    l.cell_start = a.take<int32_t>((int64_t)l.nc * l.nc + 2);
    l.counter = a.take<unsigned long long>(1);
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs((void*)nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)(n > 0 ? n : 1), 0, 32);
    l.cub_tmp_bytes = (int64_t)b;
    l.cub_tmp = a.take<char>(l.cub_tmp_bytes);
    l.total = a.used;
    if (ok) *ok = a.ok();
    return l;
}

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
            key = cy * nc + cx;
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

__global__ void knn_cell_start_kernel(const int32_t* __restrict__ keys, int64_t n, int ncell,
                                      int32_t* __restrict__ cell_start) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > ncell) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < c) lo = mid + 1; else hi = mid;
    }
    cell_start[c] = (int32_t)lo;
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

__global__ void knn_query_kernel(const double* __restrict__ xs, const double* __restrict__ ys,
                                 const int32_t* __restrict__ keys, const int32_t* __restrict__ vals, int64_t n,
                                 const int32_t* __restrict__ cell_start, int nc, double hw, double h, int k,
                                 const unsigned long long* __restrict__ counter, double* __restrict__ dk) {
    extern __shared__ double s_best[];
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int pid = vals[s];
    const int key = keys[s];
    if (key >= nc * nc) { dk[pid] = 0.0; return; }
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    if ((unsigned long long)k > *counter) { dk[pid] = INF; return; }
    const int stride = blockDim.x;
    double* best = s_best + threadIdx.x;
    for (int j = 0; j < k; ++j) best[j * stride] = INF;
    const double x = xs[s], y = ys[s];
    const int cx = key % nc, cy = key / nc;
    for (int r = 0; r < nc; ++r) {
        const int xa = max(0, cx - r), xb = min(nc - 1, cx + r);
        if (r == 0) {
            knn_scan_range(xs, ys, cell_start[key], cell_start[key + 1], x, y, best, k, stride);
        } else {
            if (cy - r >= 0) {
                const int row = (cy - r) * nc;
                knn_scan_range(xs, ys, cell_start[row + xa], cell_start[row + xb + 1], x, y, best, k, stride);
            }
            if (cy + r < nc) {
                const int row = (cy + r) * nc;
                knn_scan_range(xs, ys, cell_start[row + xa], cell_start[row + xb + 1], x, y, best, k, stride);
            }
            const int ya = max(0, cy - r + 1), yb = min(nc - 1, cy + r - 1);
            for (int yy = ya; yy <= yb; ++yy) {
                if (cx - r >= 0) {
                    const int c = yy * nc + cx - r;
                    knn_scan_range(xs, ys, cell_start[c], cell_start[c + 1], x, y, best, k, stride);
                }
                if (cx + r < nc) {
                    const int c = yy * nc + cx + r;
                    knn_scan_range(xs, ys, cell_start[c], cell_start[c + 1], x, y, best, k, stride);
                }
            }
        }
        // every unseen point lies outside the searched square of cells; stop when the k-th best is
        // safely inside it (sides that coincide with the grid boundary have nothing beyond them)
        double lim = INF;
        if (cx - r > 0) lim = fmin(lim, (x + hw) - (double)(cx - r) * h);
        if (cx + r < nc - 1) lim = fmin(lim, (double)(cx + r + 1) * h - (x + hw));
        if (cy - r > 0) lim = fmin(lim, (y + hw) - (double)(cy - r) * h);
        if (cy + r < nc - 1) lim = fmin(lim, (double)(cy + r + 1) * h - (y + hw));
        if (!(lim < INF)) break;   // whole grid searched
        if (lim > 0.0 && best[(k - 1) * stride] < lim * lim * (1.0 - 1e-12)) break;
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
    const int nc = l.nc;
    const double h = 2.0 * half_width / nc, inv_h = nc / (2.0 * half_width);
    SO_CUDA(cudaMemsetAsync(l.counter, 0, sizeof(unsigned long long), st));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    knn_key_kernel<<<blocks, 256, 0, st>>>(X, Y, ix, n, half_width, inv_h, nc, l.keys0, l.vals0, l.counter);
    SO_LAUNCH_CHECK();
    int bits = 1;
    while ((1ll << bits) < (int64_t)nc * nc + 1) ++bits;
    size_t tmp = (size_t)l.cub_tmp_bytes;
    SO_CUDA(cub::DeviceRadixSort::SortPairs(l.cub_tmp, tmp, l.keys0, l.keys1, l.vals0, l.vals1, (int)n, 0, bits, st));
    knn_gather_kernel<<<blocks, 256, 0, st>>>(X, Y, l.vals1, n, l.xs, l.ys);
    SO_LAUNCH_CHECK();
    const int ncell = nc * nc;
    knn_cell_start_kernel<<<(ncell + 1 + 255) / 256, 256, 0, st>>>(l.keys1, n, ncell, l.cell_start);
    SO_LAUNCH_CHECK();
    SO_CUDA(cudaFuncSetAttribute(knn_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_query_kernel<<<(unsigned)((n + threads - 1) / threads), threads, smem, st>>>(
        l.xs, l.ys, l.keys1, l.vals1, n, l.cell_start, nc, half_width, h, k, l.counter, dk);
    SO_LAUNCH_CHECK();
    return SO_OK;
}

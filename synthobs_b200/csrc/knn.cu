// K7 — exact k-th nearest-neighbour distance in 2-D (float64) on an implicit quadtree.
//
// Replaces cKDTree(...).query(points, k)[..., -1] of synthobs/morph/images.py:83-85: the distance
// from every selected particle to its k-th neighbour, the particle itself being the first
// (distance 0).
//
// Particles are sorted by the 30-bit Morton code of their cell on a 32768 x 32768 grid over the image.
// In that order every quadtree node (any of 15 levels) is one contiguous range: found in O(1) from an
// upper-bound table over the cells of level Lt (E[c] = #particles in table cells <= c, Lt ~ log4 n),
// and by bisection of the sorted keys inside one table cell for the deeper levels.  A query walks the
// tree depth-first, nearest child first, pruning nodes whose square is farther than the current k-th
// best, and scans leaves (<= kLeaf points or finest level).  Exact for any clustering: dense cores,
// sparse outskirts and isolated particles all cost O(log) node visits plus a few leaf scans.
// The squared distance is formed with separately rounded products and sum (no FMA contraction) and
// an IEEE sqrt, i.e. the same float64 arithmetic as a straightforward C loop.
#include <stdlib.h>

#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int kLeaf = 32;           // nodes with at most this many points are scanned
constexpr int kBits = 15;           // Morton bits per axis: the implicit quadtree has levels 0..15
constexpr int kTableLevelMax = 11;  // node ranges down to this level come from the table E, deeper ones by bisection
constexpr int kStack = 3 * kBits + 8;
constexpr int kMaxTableBits = 26;   // at most 2^26 table cells over all images of a batch

// Sort key = (image id << 30) | 30-bit Morton code; particles outside their image get image id n_img and
// sort last.  A batch of images (galaxies) is searched in one pass: the node tables of the images are
// simply concatenated, because (key >> shift) enumerates (image, cell) pairs in sorted order.
struct KnnLayout {
    void *keys0, *keys1;            // uint32 keys for one image (31 bits), uint64 for a batch of images
    int32_t *vals0, *vals1;
    double2* pts;                   // positions in sorted (Morton) order
    double2* xy;                    // positions packed in input order: one sector per particle for the sorted gather
    int32_t* E;                    // [4^Lt] upper bounds per table cell
    unsigned long long* counter;   // [0] selected count
    void* cub_tmp;
    int64_t cub_tmp_bytes;
    int Lt;
    int64_t total;
};

static int knn_table_level(int64_t n, int64_t n_img) {
    const int64_t per = (n + n_img - 1) / n_img;
    int L = 2;
    while (L < kTableLevelMax && ((int64_t)1 << (2 * (L - 1))) < per) ++L;   // ~ one particle per 4 table cells
    while (L > 0 && (n_img << (2 * L)) > ((int64_t)1 << kMaxTableBits)) --L;
    return L;
}

static KnnLayout carve_knn(void* ws, int64_t bytes, int64_t n, int64_t n_img, bool* ok) {
    KnnLayout l;
    l.Lt = knn_table_level(n, n_img);
    const int64_t ncell = n_img << (2 * l.Lt);
    SoArena a(ws, bytes);
    l.keys0 = a.take<unsigned long long>(n);     // sized for the wider key type
    l.vals0 = a.take<int32_t>(n);
    l.keys1 = a.take<unsigned long long>(n);
    l.vals1 = a.take<int32_t>(n);
    l.pts = a.take<double2>(n);
    l.xy = a.take<double2>(n);
    l.E = a.take<int32_t>(ncell);
    l.counter = a.take<unsigned long long>(2);
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs((void*)nullptr, b1, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)(n > 0 ? n : 1), 0, 64);
    cub::DeviceScan::InclusiveScan((void*)nullptr, b2, (int32_t*)nullptr, (int32_t*)nullptr, cub::Max(), (int)ncell);
    l.cub_tmp_bytes = (int64_t)(b1 > b2 ? b1 : b2);
    l.cub_tmp = a.take<char>(l.cub_tmp_bytes);
    l.total = a.used;
    if (ok) *ok = a.ok();
    return l;
}

__host__ __device__ __forceinline__ uint32_t spread_bits(uint32_t v) {   // 16 bits -> even bit positions
    v &= 0xffff;
    v = (v | (v << 8)) & 0x00ff00ff;
    v = (v | (v << 4)) & 0x0f0f0f0f;
    v = (v | (v << 2)) & 0x33333333;
    v = (v | (v << 1)) & 0x55555555;
    return v;
}
__device__ __forceinline__ uint32_t morton(uint32_t cx, uint32_t cy) { return spread_bits(cx) | (spread_bits(cy) << 1); }
__device__ __forceinline__ uint32_t compact_bits(uint32_t c) {   // even bit positions -> 16 bits
    c &= 0x55555555u;
    c = (c | (c >> 1)) & 0x33333333u; c = (c | (c >> 2)) & 0x0f0f0f0fu;
    c = (c | (c >> 4)) & 0x00ff00ffu; c = (c | (c >> 8)) & 0xffffu;
    return c;
}

// sort keys (image id, Morton code of the fine cell) + the positions packed as (x, y) pairs.  Grid-stride with
// one counter update per CTA: a same-address atomic per warp serialises (3.3 ms at 1e8 particles).
constexpr int kKeyThreads = 512;
template <typename KeyT>
__global__ void __launch_bounds__(kKeyThreads) knn_key_kernel(
        const double* __restrict__ X, const double* __restrict__ Y, const int32_t* __restrict__ ix,
        const int32_t* __restrict__ img, int64_t n, int64_t n_img, double hw, double inv_h,
        KeyT* __restrict__ keys, int32_t* __restrict__ vals, double2* __restrict__ xy,
        unsigned long long* __restrict__ counter) {
    __shared__ unsigned int s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    unsigned int mine = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool sel = ix[i] >= 0;
        const double x = X[i], y = Y[i];
        unsigned long long key = (unsigned long long)n_img << (2 * kBits);
        if (sel) {
            const int nc = 1 << kBits;
            int cx = (int)((x + hw) * inv_h), cy = (int)((y + hw) * inv_h);
            cx = max(0, min(nc - 1, cx));
            cy = max(0, min(nc - 1, cy));
            key = ((unsigned long long)(img ? img[i] : 0) << (2 * kBits)) | morton((uint32_t)cx, (uint32_t)cy);
            ++mine;
        }
        keys[i] = (KeyT)key;
        vals[i] = (int32_t)i;
        xy[i] = make_double2(x, y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_count, mine);
    __syncthreads();
    if (threadIdx.x == 0 && s_count) atomicAdd(counter, (unsigned long long)s_count);
}

__global__ void knn_gather_kernel(const double2* __restrict__ xy, const int32_t* __restrict__ vals, int64_t n,
                                  double2* __restrict__ pts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pts[i] = __ldg(xy + vals[i]);
}

// E[c] = i + 1 at the last sorted element of table cell c (others stay 0); a max-scan then makes
// E[c] = #selected particles whose table cell is <= c
template <typename KeyT>
__global__ void knn_mark_kernel(const KeyT* __restrict__ keys, int64_t n, int64_t n_img, int tshift,
                                int32_t* __restrict__ E) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    if ((k >> (2 * kBits)) >= (unsigned long long)n_img) return;
    if (i + 1 == n || ((unsigned long long)keys[i + 1] >> tshift) != (k >> tshift)) E[k >> tshift] = (int32_t)(i + 1);
}

// scan sorted positions [a, b) minus the window [wa, wb) that is already in the list
__device__ __forceinline__ void knn_scan_range(const double2* __restrict__ pts,
                                               int a, int b, int wa, int wb, double x, double y, double* best, int k,
                                               int stride) {
    double kth = best[(k - 1) * stride];
    if (a >= wa && a < wb) a = wb;                        // range starts inside the pre-scanned window
    for (int q = a; q < b; ++q) {
        if (q == wa) { q = wb; if (q >= b) break; }       // skip the pre-scanned window
        const double2 c = __ldg(pts + q);
        const double dx = c.x - x, dy = c.y - y;
        const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (d2 < kth) {
            int j = k - 1;
            while (j > 0 && best[(j - 1) * stride] > d2) {
                best[j * stride] = best[(j - 1) * stride];
                --j;
            }
            best[j * stride] = d2;
            kth = best[(k - 1) * stride];
        }
    }
}

// conservative lower bound (fine-cell units, FP32) of the distance from (u, v) to the square of the
// node with cell coordinates (cx, cy) and edge 2^shift: the true distance minus more than the FP32
// rounding of the inputs; 0 inside
__device__ __forceinline__ float node_dmin(int shift, uint32_t cx, uint32_t cy, float u, float v) {
    const float s = (float)(1u << shift);
    const float x0 = (float)cx * s, y0 = (float)cy * s;
    const float dx = fmaxf(fmaxf(x0 - u, u - (x0 + s)), 0.f);
    const float dy = fmaxf(fmaxf(y0 - v, v - (y0 + s)), 0.f);
    return fmaxf(sqrtf(dx * dx + dy * dy) * (1.f - 1e-5f) - 1.6e-2f, 0.f);
}

// first position in [a, b) whose key is >= key (keys ascending)
template <typename KeyT>
__device__ __forceinline__ int lower_bound_key(const KeyT* __restrict__ keys, int a, int b, unsigned long long key) {
    while (a < b) {
        const int mid = (a + b) >> 1;
        if (keys[mid] < key) a = mid + 1; else b = mid;
    }
    return a;
}

template <typename KeyT>
struct KnnTree {
    const KeyT* keys;
    const int32_t* E;
    int Lt;                    // table level
    unsigned long long img;    // image (galaxy) of the query
};

// sorted range of node (level, code) of image t.img: table look-up down to level Lt, bisection inside the
// table cell below.  (image << 2 level) | code enumerates the nodes of a level over the whole batch.
template <typename KeyT>
__device__ __forceinline__ void node_range(const KnnTree<KeyT>& t, int level, uint32_t code, int& a, int& b) {
    const unsigned long long xc = (t.img << (2 * level)) | code;
    if (level <= t.Lt) {
        const int sh = 2 * (t.Lt - level);
        const unsigned long long lo = xc << sh;
        a = (lo == 0) ? 0 : t.E[lo - 1];
        b = t.E[((xc + 1) << sh) - 1];
    } else {
        const unsigned long long tc = xc >> (2 * (level - t.Lt));
        const int ta = (tc == 0) ? 0 : t.E[tc - 1], tb = t.E[tc];
        const int sh = 2 * (kBits - level);
        a = lower_bound_key(t.keys, ta, tb, xc << sh);
        b = lower_bound_key(t.keys, a, tb, (xc + 1) << sh);
    }
}

// One thread per query, in Morton order (neighbouring lanes walk neighbouring parts of the tree).
//  1. the 2k+1 sorted neighbours of the query (itself included) give a first k-best list: an upper
//     bound r on the k-th distance that is already tight in the bulk of a smooth distribution;
//  2. the walk starts at the (at most four) tree nodes of the level whose cells are >= 2r wide and
//     that cover the box [u-r, u+r] x [v-r, v+r], instead of at the root;
//  3. depth-first, nearest child first, pruning on the current k-th best; leaves (<= kLeaf points
//     or finest level) are scanned, skipping the window of step 1.
// The tree is 15 levels deep whatever the particle number, so the dense cores of a clustered
// catalogue (hundreds of particles per table cell) are resolved instead of being scanned linearly.
// (a register-resident list for k == 8 — compile-time K, compare-exchange insertion — was slower: 60 registers, 0.71 vs
// 0.65 ms per 1e6 queries)
template <typename KeyT>
__global__ void __launch_bounds__(128, 12) knn_query_kernel(const double2* __restrict__ pts,
                                 const KeyT* __restrict__ keys, const int32_t* __restrict__ vals, int64_t n,
                                 int64_t n_img, const int32_t* __restrict__ E, int Lt, double hw, double inv_h, int k,
                                 int part, int n_parts, int leaf, int win, double floor2, const uint8_t* __restrict__ qmask,
                                 double* __restrict__ dk) {
    extern __shared__ double s_best[];
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int pid = vals[s];
    const unsigned long long key = keys[s];
    if ((key >> (2 * kBits)) >= (unsigned long long)n_img) { dk[pid] = 0.0; return; }
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    KnnTree<KeyT> t{keys, E, Lt, key >> (2 * kBits)};
    // sorted range of the query's own image: its neighbours are searched there only
    int g_lo, g_hi;
    node_range(t, 0, 0u, g_lo, g_hi);
    const int nsel = g_hi - g_lo;
    // this call's share of the queries: a contiguous range of the Morton order (spatially compact)
    if (s - g_lo < (int64_t)nsel * part / n_parts || s - g_lo >= (int64_t)nsel * (part + 1) / n_parts) return;
    if (qmask && !(qmask[pid] & 2)) return;
    if (k > nsel) { dk[pid] = INF; return; }
    const int stride = blockDim.x;
    double* best = s_best + threadIdx.x;
    for (int j = 0; j < k; ++j) best[j * stride] = INF;
    auto kth = [&]() -> double { return best[(k - 1) * stride]; };
    const double2 self = pts[s];
    const double x = self.x, y = self.y;
    auto scan = [&](int a, int b, int sa, int sb) { knn_scan_range(pts, a, b, sa, sb, x, y, best, k, stride); };
    const double ud = (x + hw) * inv_h, vd = (y + hw) * inv_h;   // fine-cell units, as used for the keys
    const float u = (float)ud, v = (float)vd;
    const double h = 1.0 / inv_h;                                 // fine-cell units -> length

    // 1. window of sorted neighbours
    int wb = min(g_hi, (int)s + win + 1);
    const int wa = max(g_lo, wb - (2 * win + 1));
    wb = min(g_hi, wa + 2 * win + 1);
    scan(wa, wb, -1, -1);
    // the caller only needs max(dk, floor): k neighbours inside the floor radius settle the answer
    if (kth() <= floor2) { dk[pid] = sqrt(kth()); return; }

    // 2. start nodes
    uint32_t stack[kStack];       // node = (1 << 2 level) | code: the leading one marks the level
    int sp = 0;
    {
        const double rc = sqrt(kth()) * inv_h * (1.0 + 1e-9) + 1e-9;   // fine-cell units
        int e = 0;
        if (rc >= 0.5) e = ilogb(2.0 * rc) + 1;                 // 2^e >= 2 rc
        if (!(rc < (double)(1 << kBits)) || e >= kBits) {
            stack[sp++] = 1u;                                   // root
        } else {
            const int level = kBits - e;
            const double sc = (double)(1 << e);
            const int top = (1 << level) - 1;
            const int cx0 = max(0, min(top, (int)floor((ud - rc) / sc))), cx1 = max(0, min(top, (int)floor((ud + rc) / sc)));
            const int cy0 = max(0, min(top, (int)floor((vd - rc) / sc))), cy1 = max(0, min(top, (int)floor((vd + rc) / sc)));
            for (int cy = cy0; cy <= cy1; ++cy)
                for (int cx = cx0; cx <= cx1; ++cx) {
                    stack[sp++] = (1u << (2 * level)) | morton((uint32_t)cx, (uint32_t)cy);
                }
        }
    }

    // 3. pruned depth-first walk
    while (sp > 0) {
        const uint32_t node = stack[--sp];
        const int level = (31 - __clz(node)) >> 1;
        const uint32_t code = node ^ (1u << (2 * level));
        const uint32_t ncx = compact_bits(code), ncy = compact_bits(code >> 1);
        const double dm = (double)node_dmin(kBits - level, ncx, ncy, u, v) * h;
        if (dm * dm >= kth()) continue;
        int a, b;
        node_range(t, level, code, a, b);
        if (a == b) continue;
        if (b - a <= leaf || level == kBits) {
            scan(a, b, wa, wb);
            continue;
        }
        // children, pushed farthest first so the nearest is expanded next
        uint32_t cc[4];
        float dd[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            cc[c] = (code << 2) | c;
            dd[c] = node_dmin(kBits - level - 1, (ncx << 1) + (c & 1), (ncy << 1) + (c >> 1), u, v);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3 - i; ++j)
                if (dd[j] < dd[j + 1]) {
                    const float td = dd[j]; dd[j] = dd[j + 1]; dd[j + 1] = td;
                    const uint32_t tc = cc[j]; cc[j] = cc[j + 1]; cc[j + 1] = tc;
                }
#pragma unroll
        for (int c = 0; c < 4; ++c) stack[sp++] = (1u << (2 * level + 2)) | cc[c];
    }
    dk[pid] = sqrt(kth());
}

// Warp-cooperative variant (default): one warp answers 32 consecutive queries of the Morton order, which
// are neighbours in space.  The tree walk is shared (uniform control flow): a node is opened when ANY
// lane still needs it (its own lower bound vs its own k-th best, __any_sync), a leaf is loaded once,
// coalesced, into shared memory and every lane tests every candidate against its own list.  Per query
// this scans more candidates than the private walk above but with all 32 lanes busy (the private walk
// runs at ~9 of 32 active lanes), and the node tests, table look-ups and point loads are paid once per warp.
constexpr int kWarpsPerCta = 4;

template <typename KeyT>
__global__ void __launch_bounds__(kWarpsPerCta * 32) knn_query_warp_kernel(
        const double2* __restrict__ pts, const KeyT* __restrict__ keys,
        const int32_t* __restrict__ vals, int64_t n, int64_t n_img, const int32_t* __restrict__ E, int Lt, double hw,
        double inv_h, int k, int part, int n_parts, int leaf, int win, double floor2, const uint8_t* __restrict__ qmask,
        double* __restrict__ dk) {
    extern __shared__ double s_dyn[];
    __shared__ double s_px[kWarpsPerCta][32], s_py[kWarpsPerCta][32];
    __shared__ uint32_t s_stack[kWarpsPerCta][kStack];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = blockDim.x;
    double* best = s_dyn + threadIdx.x;
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const unsigned FULL = 0xffffffffu;

    // per-lane query
    bool valid = s < n;
    unsigned long long key = valid ? (unsigned long long)keys[s] : ~0ull;
    const int pid = valid ? vals[s] : 0;
    unsigned long long image = key >> (2 * kBits);
    if (valid && image >= (unsigned long long)n_img) { dk[pid] = 0.0; valid = false; }
    double x = 0.0, y = 0.0;
    if (valid) { const double2 self = pts[s]; x = self.x; y = self.y; }
    const double ud = (x + hw) * inv_h, vd = (y + hw) * inv_h;   // fine-cell units, as used for the keys
    const float u = (float)ud, v = (float)vd;
    const double h = 1.0 / inv_h;

    unsigned todo = __ballot_sync(FULL, valid);
    while (todo) {                                   // one pass per image present in the warp (almost always one)
        const int leader = __ffs(todo) - 1;
        const unsigned long long g = __shfl_sync(FULL, image, leader);
        bool mine = valid && image == g;
        todo &= ~__ballot_sync(FULL, mine);
        KnnTree<KeyT> t{keys, E, Lt, g};
        int g_lo, g_hi;
        node_range(t, 0, 0u, g_lo, g_hi);
        const int nsel = g_hi - g_lo;
        if (mine && (s - g_lo < (int64_t)nsel * part / n_parts || s - g_lo >= (int64_t)nsel * (part + 1) / n_parts)) mine = false;
        if (mine && qmask && !(qmask[pid] & 2)) mine = false;
        if (mine && k > nsel) { dk[pid] = INF; mine = false; }
        if (!__any_sync(FULL, mine)) continue;

        // 1. private window of sorted neighbours -> first k-best list
        for (int j = 0; j < k; ++j) best[j * stride] = INF;
        int wa = 0, wb = 0;
        if (mine) {
            wb = min(g_hi, (int)s + win + 1);
            wa = max(g_lo, wb - (2 * win + 1));
            wb = min(g_hi, wa + 2 * win + 1);
            knn_scan_range(pts, wa, wb, -1, -1, x, y, best, k, stride);
        }
        double kth = best[(k - 1) * stride];
        if (mine && kth <= floor2) { dk[pid] = sqrt(kth); mine = false; }   // settled below the caller's floor
        if (!__any_sync(FULL, mine)) continue;

        // 2. start nodes: the level whose cells are at least as wide as the box around all the lanes' balls
        const float rc = mine ? (float)(sqrt(kth) * inv_h) * (1.f + 1e-5f) + 1.6e-2f : 0.f;
        float x0 = mine ? u - rc : 3.0e38f, x1 = mine ? u + rc : -3.0e38f;
        float y0 = mine ? v - rc : 3.0e38f, y1 = mine ? v + rc : -3.0e38f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x0 = fminf(x0, __shfl_xor_sync(FULL, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(FULL, x1, o));
            y0 = fminf(y0, __shfl_xor_sync(FULL, y0, o)); y1 = fmaxf(y1, __shfl_xor_sync(FULL, y1, o));
        }
        int sp = 0;
        {
            const float ext = fmaxf(x1 - x0, y1 - y0);
            int e = 0;
            if (ext >= 1.f) e = ilogbf(ext) + 1;                // 2^e > ext: the box spans at most 2 cells per axis
            if (!(ext < (float)(1 << kBits)) || e >= kBits) {
                if (lane == 0) s_stack[warp][0] = 1u;           // root
                sp = 1;
            } else {
                const int level = kBits - e;
                const float sc = (float)(1 << e);
                const int top = (1 << level) - 1;
                const int cx0 = max(0, min(top, (int)floorf(x0 / sc))), cx1 = max(0, min(top, (int)floorf(x1 / sc)));
                const int cy0 = max(0, min(top, (int)floorf(y0 / sc))), cy1 = max(0, min(top, (int)floorf(y1 / sc)));
                for (int cy = cy0; cy <= cy1; ++cy)
                    for (int cx = cx0; cx <= cx1; ++cx) {
                        if (lane == 0) s_stack[warp][sp] = (1u << (2 * level)) | morton((uint32_t)cx, (uint32_t)cy);
                        ++sp;
                    }
            }
        }
        __syncwarp();

        // 3. shared depth-first walk
        while (sp > 0) {
            const uint32_t node = s_stack[warp][--sp];
            __syncwarp();
            const int level = (31 - __clz(node)) >> 1;
            const uint32_t code = node ^ (1u << (2 * level));
            const uint32_t ncx = compact_bits(code), ncy = compact_bits(code >> 1);
            const double dm = (double)node_dmin(kBits - level, ncx, ncy, u, v) * h;
            const bool need = mine && dm * dm < kth;
            if (!__any_sync(FULL, need)) continue;
            int a, b;
            node_range(t, level, code, a, b);
            if (a == b) continue;
            if (b - a <= leaf || level == kBits) {
                for (int base = a; base < b; base += 32) {
                    const int m = min(32, b - base);
                    if (lane < m) { const double2 c = pts[base + lane]; s_px[warp][lane] = c.x; s_py[warp][lane] = c.y; }
                    __syncwarp();
                    if (need) {
                        for (int c = 0; c < m; ++c) {
                            const int q = base + c;
                            if (q >= wa && q < wb) continue;            // already in the list (window)
                            const double dx = s_px[warp][c] - x, dy = s_py[warp][c] - y;
                            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                            if (d2 < kth) {
                                int j = k - 1;
                                while (j > 0 && best[(j - 1) * stride] > d2) {
                                    best[j * stride] = best[(j - 1) * stride];
                                    --j;
                                }
                                best[j * stride] = d2;
                                kth = best[(k - 1) * stride];
                            }
                        }
                    }
                    __syncwarp();
                }
                continue;
            }
            if (lane < 4) s_stack[warp][sp + lane] = (1u << (2 * level + 2)) | (code << 2) | (uint32_t)lane;
            sp += 4;
            __syncwarp();
        }
        if (mine) dk[pid] = sqrt(kth);
        __syncwarp();
    }
}

__global__ void knn_range_kernel(const unsigned long long* __restrict__ counter, int part, int n_parts,
                                 int64_t* __restrict__ range_out) {
    const int64_t nsel = (int64_t)*counter;
    range_out[0] = nsel * part / n_parts;
    range_out[1] = nsel * (part + 1) / n_parts;
    range_out[2] = nsel;
}

}  // namespace

extern "C" int64_t so_knn_workspace_bytes(int64_t n, int32_t k) {
    (void)k;
    KnnLayout l = carve_knn(nullptr, 0, n, 1, nullptr);
    return l.total + 256;
}

// everything after the argument checks, for one key width: uint32 keys hold the 30-bit Morton code and the
// "outside the image" flag of ONE image (halves the radix-sort traffic of the huge-image case), uint64 keys add
// the image id of a batch
template <typename KeyT>
static int knn_run_keys(const double* X, const double* Y, const int32_t* ix, const int32_t* img, int64_t n, int64_t n_img,
                        int32_t k, double half_width, int32_t part, int32_t n_parts, double floor, const uint8_t* qmask, double* dk,
                        int32_t* order_out, int64_t* range_out, const KnnLayout& l, int threads, size_t smem,
                        cudaStream_t st) {
    KeyT* keys0 = reinterpret_cast<KeyT*>(l.keys0);
    KeyT* keys1 = reinterpret_cast<KeyT*>(l.keys1);
    const int Lt = l.Lt;
    const int64_t ncell = n_img << (2 * Lt);
    const double inv_h = (double)(1 << kBits) / (2.0 * half_width);
    int img_bits = 1;
    while (((int64_t)1 << img_bits) <= n_img) ++img_bits;
    SO_CUDA(cudaMemsetAsync(l.counter, 0, 2 * sizeof(unsigned long long), st));
    SO_CUDA(cudaMemsetAsync(l.E, 0, sizeof(int32_t) * (size_t)ncell, st));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    int64_t key_blocks = (n + kKeyThreads - 1) / kKeyThreads;
    if (key_blocks > 8 * so_num_sms()) key_blocks = 8 * so_num_sms();
    { SO_PROF("knn_key_kernel", st); knn_key_kernel<KeyT><<<(unsigned)key_blocks, kKeyThreads, 0, st>>>(X, Y, ix, img, n, n_img, half_width, inv_h, keys0,
                                                                     l.vals0, l.xy, l.counter); }
    SO_LAUNCH_CHECK();
    size_t tmp = (size_t)l.cub_tmp_bytes;
    { SO_PROF("cub sort Morton keys", st); SO_CUDA(cub::DeviceRadixSort::SortPairs(l.cub_tmp, tmp, keys0, keys1, l.vals0, l.vals1, (int)n, 0,
                                            2 * kBits + img_bits, st)); }
    { SO_PROF("knn_gather_kernel", st); knn_gather_kernel<<<blocks, 256, 0, st>>>(l.xy, l.vals1, n, l.pts); }
    SO_LAUNCH_CHECK();
    { SO_PROF("knn_mark_kernel", st); knn_mark_kernel<KeyT><<<blocks, 256, 0, st>>>(keys1, n, n_img, 2 * (kBits - Lt), l.E); }
    SO_LAUNCH_CHECK();
    tmp = (size_t)l.cub_tmp_bytes;
    SO_CUDA(cub::DeviceScan::InclusiveScan(l.cub_tmp, tmp, l.E, l.E, cub::Max(), (int)ncell, st));
    // measured on B200 (1e6 clustered particles): the private walk wins for k <= 8 (0.71 vs 0.76 ms), the shared
    // walk for larger k (k = 16: 1.17 vs 1.27 ms).  SO_KNN_PRIVATE=0/1 and SO_KNN_LEAF override (A/B aids).
    static int force_private = -2, leaf = 0;
    if (force_private == -2) {
        const char* e = getenv("SO_KNN_PRIVATE");
        force_private = e ? (e[0] == '1' ? 1 : 0) : -1;
        const char* lf = getenv("SO_KNN_LEAF");
        leaf = lf ? atoi(lf) : kLeaf;
        if (leaf < 1) leaf = kLeaf;
    }
    const bool private_walk = force_private >= 0 ? force_private == 1 : k <= 8;
    const double floor2 = floor > 0.0 ? floor * floor : -1.0;
    // half-width of the window of sorted neighbours scanned first (SO_KNN_WIN = multiple of k, default 1)
    static int win_mul = 0;
    if (win_mul == 0) { const char* e = getenv("SO_KNN_WIN"); win_mul = e ? atoi(e) : 1; if (win_mul < 1) win_mul = 1; }
    const int win = k * win_mul;
    if (private_walk || threads != kWarpsPerCta * 32) {
        SO_CUDA(cudaFuncSetAttribute(knn_query_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        { SO_PROF("knn_query_kernel", st); knn_query_kernel<KeyT><<<(unsigned)((n + threads - 1) / threads), threads, smem, st>>>(
            l.pts, keys1, l.vals1, n, n_img, l.E, Lt, half_width, inv_h, k, part, n_parts, leaf, win, floor2, qmask, dk); }
    } else {
        SO_CUDA(cudaFuncSetAttribute(knn_query_warp_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        { SO_PROF("knn_query_warp_kernel", st); knn_query_warp_kernel<KeyT><<<(unsigned)((n + threads - 1) / threads), threads, smem, st>>>(
            l.pts, keys1, l.vals1, n, n_img, l.E, Lt, half_width, inv_h, k, part, n_parts, leaf, win, floor2, qmask, dk); }
    }
    SO_LAUNCH_CHECK();
    if (order_out) SO_CUDA(cudaMemcpyAsync(order_out, l.vals1, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, st));
    if (range_out) {
        { SO_PROF("knn_range_kernel", st); knn_range_kernel<<<1, 1, 0, st>>>(l.counter, part, n_parts, range_out); }
        SO_LAUNCH_CHECK();
    }
    return SO_OK;
}

static int knn_run(const double* X, const double* Y, const int32_t* ix, const int32_t* img, int64_t n, int64_t n_img,
                   int32_t k, double half_width, int32_t part, int32_t n_parts, double floor, const uint8_t* qmask, double* dk,
                   int32_t* order_out,
                   int64_t* range_out, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
    if (n == 0) return SO_OK;
    if (n < 0 || k < 1 || !(half_width > 0.0) || !X || !Y || !ix || !dk || n_img < 1) return SO_ERR_BAD_ARG;
    if (n_parts < 1 || part < 0 || part >= n_parts || (n_img > 1 && (n_parts != 1 || !img))) return SO_ERR_BAD_ARG;
    if (n >= (int64_t)1 << 31 || n_img >= (int64_t)1 << 30) return SO_ERR_TOO_LARGE;
    bool ok = false;
    KnnLayout l = carve_knn(workspace, workspace_bytes, n, n_img, &ok);
    if (!ok || !workspace) return SO_ERR_WORKSPACE;
    int threads = 128;
    size_t smem = (size_t)k * threads * sizeof(double);
    if (smem > 200 * 1024) { threads = 32; smem = (size_t)k * threads * sizeof(double); }
    if (smem > 200 * 1024) return SO_ERR_UNSUPPORTED;
    if (n_img == 1)
        return knn_run_keys<uint32_t>(X, Y, ix, img, n, n_img, k, half_width, part, n_parts, floor, qmask, dk, order_out, range_out,
                                      l, threads, smem, st);
    return knn_run_keys<unsigned long long>(X, Y, ix, img, n, n_img, k, half_width, part, n_parts, floor, qmask, dk, order_out,
                                            range_out, l, threads, smem, st);
}

extern "C" int so_knn_kth_distance(const double* X, const double* Y, const int32_t* ix, int64_t n,
                                   int32_t k, double half_width, double* dk,
                                   void* workspace, int64_t workspace_bytes, void* stream) {
    return knn_run(X, Y, ix, nullptr, n, 1, k, half_width, 0, 1, 0.0, nullptr, dk, nullptr, nullptr, workspace, workspace_bytes,
                   (cudaStream_t)stream);
}

extern "C" int so_knn_kth_distance_part(const double* X, const double* Y, const int32_t* ix, int64_t n,
                                        int32_t k, double half_width, int32_t part, int32_t n_parts,
                                        double* dk, int32_t* order_out, int64_t* range_out,
                                        void* workspace, int64_t workspace_bytes, void* stream) {
    return knn_run(X, Y, ix, nullptr, n, 1, k, half_width, part, n_parts, 0.0, nullptr, dk, order_out, range_out, workspace,
                   workspace_bytes, (cudaStream_t)stream);
}

extern "C" int64_t so_knn_batch_workspace_bytes(int64_t n, int64_t n_img, int32_t k) {
    (void)k;
    KnnLayout l = carve_knn(nullptr, 0, n, n_img < 1 ? 1 : n_img, nullptr);
    return l.total + 256;
}

extern "C" int so_knn_kth_distance_batch(const double* X, const double* Y, const int32_t* ix, const int32_t* img,
                                         int64_t n, int64_t n_img, int32_t k, double half_width, double* dk,
                                         void* workspace, int64_t workspace_bytes, void* stream) {
    return knn_run(X, Y, ix, img, n, n_img, k, half_width, 0, 1, 0.0, nullptr, dk, nullptr, nullptr, workspace, workspace_bytes,
                   (cudaStream_t)stream);
}

extern "C" int so_knn_kth_distance_ex(const double* X, const double* Y, const int32_t* ix, const int32_t* img,
                                      int64_t n, int64_t n_img, int32_t k, double half_width, int32_t part, int32_t n_parts,
                                      double floor, double* dk, int32_t* order_out, int64_t* range_out,
                                      void* workspace, int64_t workspace_bytes, void* stream) {
    return knn_run(X, Y, ix, img, n, n_img < 1 ? 1 : n_img, k, half_width, part, n_parts, floor, nullptr, dk, order_out, range_out,
                   workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int so_knn_kth_distance_q(const double* X, const double* Y, const int32_t* ix, int64_t n, int32_t k, double half_width,
                                     const uint8_t* query_mask, double floor, double* dk, void* workspace,
                                     int64_t workspace_bytes, void* stream) {
    return knn_run(X, Y, ix, nullptr, n, 1, k, half_width, 0, 1, floor, query_mask, dk, nullptr, nullptr, workspace, workspace_bytes,
                   (cudaStream_t)stream);
}

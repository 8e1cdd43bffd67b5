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
//
// Calls of at most kTileAutoMax particles use a second search instead (knn_tile_kernel, further down): one warp per 32
// consecutive queries of the same sorted order, candidate tiles from a hierarchy of bucket boxes, FP32 screening and a
// float64 decision for the survivors - equally exact, faster where a launch is latency-bound.
#include <stdio.h>
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
// Bucket hierarchy over the sorted order (search structure of the tile kernel): level 0 = buckets of kBucket consecutive
// points, level l + 1 = groups of four level-l nodes.  Node i of level l covers the sorted positions
// [i * 8 * 4^l, (i + 1) * 8 * 4^l): no child pointers, no ranges to look up.
constexpr int kBucket = 8;
constexpr int kMaxLevels = 16;
struct KnnLevels {
    int nlev;                       // levels 0 .. nlev-1, the last one has a single node
    int32_t cnt[kMaxLevels];
    int64_t off[kMaxLevels];        // first box of the level in bb
    int64_t total;
};
struct __align__(32) KnnBox { double x0, y0, x1, y1; };

static KnnLevels knn_levels(int64_t n) {
    KnnLevels lv;
    for (int l = 0; l < kMaxLevels; ++l) { lv.cnt[l] = 0; lv.off[l] = 0; }
    int64_t c = (n + kBucket - 1) / kBucket;
    if (c < 1) c = 1;
    int l = 0;
    int64_t off = 0;
    for (;;) {
        lv.cnt[l] = (int32_t)c;
        lv.off[l] = off;
        off += c;
        ++l;
        if (c == 1 || l == kMaxLevels) break;
        c = (c + 3) / 4;
    }
    lv.nlev = l;
    lv.total = off;
    return lv;
}

struct KnnLayout {
    void *keys0, *keys1;            // uint32 keys for one image (31 bits), uint64 for a batch of images
    int32_t *vals0, *vals1;
    double2* pts;                   // positions in sorted (Morton) order
    double2* xy;                    // positions packed in input order: one sector per particle for the sorted gather
    int32_t* E;                    // [4^Lt] upper bounds per table cell
    unsigned long long* counter;   // [0] selected count
    struct KnnBox* bb;             // bounding boxes of the bucket hierarchy, level after level
    int32_t* img_start;            // [n_img + 2] first sorted position of every image; [n_img] = #selected
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
    l.bb = a.take<KnnBox>(knn_levels(n).total);
    l.img_start = a.take<int32_t>(n_img + 2);
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
        const double x = X[i], y = Y[i];
        unsigned long long key = (unsigned long long)n_img << (2 * kBits);
        // an image id outside [0, n_img) would leave key bits above the sorted range: such a particle counts as outside
        const bool sel = ix[i] >= 0 && (!img || (unsigned)img[i] < (unsigned long long)n_img);
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
                                int32_t* __restrict__ E, const unsigned long long* __restrict__ gate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (gate && *gate == 0) return;                       // fallback of the tile kernel and nobody asked
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
                                 double* __restrict__ dk, const unsigned long long* __restrict__ gate) {
    extern __shared__ double s_best[];
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    if (gate && *gate == 0) return;                       // fallback of the tile kernel and nobody asked
    const int pid = vals[s];
    if (gate && dk[pid] == dk[pid]) return;               // only the queries the tile kernel marked with NaN
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

// ---- bucket hierarchy over the Morton order (search structure of the tile kernel) ---------------------------------
// The sorted order is cut into buckets of kBucket = 8 points with their TIGHT float64 bounding boxes, grouped four by
// four up to a single root: node i of level l covers the sorted positions [i 8 4^l, (i + 1) 8 4^l), so node ranges are
// index arithmetic (no table, no bisection) and a level-1 node is exactly one tile of 32 queries.  Boxes are min/max of
// the members, and box distances are formed with the same separately rounded, monotone operations as a candidate's
// distance: box_d2 <= d2 of every member, so pruning on it never drops a candidate that would have been kept.
// (A per-query depth-first walk of this hierarchy - one thread per query as in knn_query_kernel - was measured first:
// exact, 1.05 ms against 0.66 ms per 1e6 queries for the quadtree walk: the boxes of consecutive runs of the Morton
// order overlap, and without a nearest-first order more buckets are opened.)

constexpr int kBvhThreads = 1024;   // one CTA of the build kernel covers levels 0..3 of 1024 sorted points

template <typename KeyT>
__global__ void __launch_bounds__(kBvhThreads) knn_bvh_leaf_kernel(
        const double2* __restrict__ pts, const KeyT* __restrict__ keys, int64_t n, int64_t n_img,
        KnnBox* __restrict__ bb, KnnLevels lv, int32_t* __restrict__ img_start) {
    __shared__ KnnBox s_box[kBvhThreads / 32];
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const int64_t s = (int64_t)blockIdx.x * kBvhThreads + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t im = n_img;
    if (s < n) { im = (int64_t)((unsigned long long)keys[s] >> (2 * kBits)); if (im > n_img) im = n_img; }
    const bool valid = s < n && im < n_img;
    double x0 = INF, y0 = INF, x1 = -INF, y1 = -INF;
    if (valid) { const double2 p = pts[s]; x0 = x1 = p.x; y0 = y1 = p.y; }
    // first sorted position of every image (images without selected particles share their successor's)
    if (s < n) {
        int64_t prev = -1;
        if (s > 0) { prev = (int64_t)((unsigned long long)keys[s - 1] >> (2 * kBits)); if (prev > n_img) prev = n_img; }
        for (int64_t j = prev + 1; j <= im; ++j) img_start[j] = (int32_t)s;
        if (s == n - 1) for (int64_t j = im + 1; j <= n_img; ++j) img_start[j] = (int32_t)n;
    }
    const unsigned FULL = 0xffffffffu;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        x0 = fmin(x0, __shfl_xor_sync(FULL, x0, o)); y0 = fmin(y0, __shfl_xor_sync(FULL, y0, o));
        x1 = fmax(x1, __shfl_xor_sync(FULL, x1, o)); y1 = fmax(y1, __shfl_xor_sync(FULL, y1, o));
    }
    if ((lane & 7) == 0 && (s >> 3) < lv.cnt[0]) bb[lv.off[0] + (s >> 3)] = KnnBox{x0, y0, x1, y1};
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
        x0 = fmin(x0, __shfl_xor_sync(FULL, x0, o)); y0 = fmin(y0, __shfl_xor_sync(FULL, y0, o));
        x1 = fmax(x1, __shfl_xor_sync(FULL, x1, o)); y1 = fmax(y1, __shfl_xor_sync(FULL, y1, o));
    }
    if (lane == 0) {
        if ((s >> 5) < lv.cnt[1]) bb[lv.off[1] + (s >> 5)] = KnnBox{x0, y0, x1, y1};
        s_box[warp] = KnnBox{x0, y0, x1, y1};
    }
    __syncthreads();
    // levels 2 (4 warps) and 3 (16 warps)
    for (int level = 2; level <= 3; ++level) {
        const int per = level == 2 ? 4 : 16, cnt = (kBvhThreads / 32) / per;
        if ((int)threadIdx.x < cnt) {
            KnnBox b = s_box[threadIdx.x * per];
            for (int j = 1; j < per; ++j) {
                const KnnBox c = s_box[threadIdx.x * per + j];
                b.x0 = fmin(b.x0, c.x0); b.y0 = fmin(b.y0, c.y0); b.x1 = fmax(b.x1, c.x1); b.y1 = fmax(b.y1, c.y1);
            }
            const int64_t idx = (int64_t)blockIdx.x * cnt + threadIdx.x;
            if (idx < lv.cnt[level]) bb[lv.off[level] + idx] = b;
        }
    }
}

// levels 4 and up (n / 2048 nodes and fewer): one CTA, level after level
__global__ void __launch_bounds__(1024) knn_bvh_upper_kernel(KnnBox* __restrict__ bb, KnnLevels lv) {
    for (int level = 4; level < lv.nlev; ++level) {
        const KnnBox* child = bb + lv.off[level - 1];
        const int nchild = lv.cnt[level - 1];
        for (int idx = threadIdx.x; idx < lv.cnt[level]; idx += blockDim.x) {
            KnnBox b = child[4 * (int64_t)idx];
            for (int j = 1; j < 4; ++j) {
                if (4 * (int64_t)idx + j >= nchild) break;
                const KnnBox c = child[4 * (int64_t)idx + j];
                b.x0 = fmin(b.x0, c.x0); b.y0 = fmin(b.y0, c.y0); b.x1 = fmax(b.x1, c.x1); b.y1 = fmax(b.y1, c.y1);
            }
            bb[lv.off[level] + idx] = b;
        }
        __syncthreads();
    }
}

// ---- tiles: one warp answers the 32 queries of one tile (level-1 node) ---------------------------------------------
// lane = query.  The candidate tiles are found ONCE per tile: the siblings of every ancestor of the tile are tested
// against the bounding box of the open queries widened by their largest radius (one node per lane, float64), needed
// nodes are opened in batches of up to 32, and a leaf tile is used only when some lane's own ball reaches its box.
// A needed tile is loaded once (one coalesced 16-byte load per lane, coordinates relative to the tile's first point
// rounded to FP32) and screened TRANSPOSED: lane = candidate, loop over the queries that need the tile (their position
// and threshold broadcast by shuffles), so 32 candidates cost one query ~16 instructions.  A candidate inside the
// query's threshold is only APPENDED to that query's list in shared memory (key = bit pattern of the FP32 squared
// distance, sorted position) - no float64, no insertion under divergence.  After each tile every lane passes its new
// entries through a branch-free compare-exchange chain of KC registers (the k smallest keys), which tightens its
// threshold.  At the end the entries inside the final threshold - the k nearest, plus, rarely, near-ties - are
// evaluated in float64 exactly like a C loop would (separately rounded products and sum), and the k-th smallest is the
// answer.  Exactness: a coordinate difference formed from the rounded relative coordinates is off by at most
// 2^-23 ext (ext = largest relative coordinate loaded so far), so with E = 2.5e-7 ext every point within the true k-th
// distance has an FP32 squared distance below thr = (sqrt(a) (1 + 1e-6) + 2 E)^2 (1 + 1e-6), a = k-th smallest key so
// far; boxes are float64 min/max of their members and box distances are formed with the same monotone operations as
// a candidate's, so pruning on box_d2 >= thr never drops such a point.  A lane whose list cannot be compacted below
// its capacity (more than kTileSpare near-ties of the k-th distance: stacks of identical positions) is handed to the
// quadtree walk (dk = NaN + hard_count), which is launched behind this kernel and exits at once when nobody asked.
constexpr int kTileWarps = 4;
constexpr int64_t kTileAutoMax = 32768;
constexpr int kTileStack = 192;    // batches stop growing the stack 48 entries (3 per level of a depth-first descent) below
constexpr int kTileSpare = 24;
__host__ __device__ constexpr int kTileCapOf(int kc) { return (kc + kTileSpare + 32) | 1; }     // odd row length: conflict-free both ways

template <typename KeyT, int KC, bool BATCH, bool STATS>
__global__ void __launch_bounds__(kTileWarps * 32) knn_tile_kernel(
        const double2* __restrict__ pts, const KeyT* __restrict__ keys, const int32_t* __restrict__ vals, int64_t n,
        int64_t n_img, const int32_t* __restrict__ img_start, const KnnBox* __restrict__ bb, KnnLevels lv, int k,
        int part, int n_parts, double floor2, int tighten_every, const uint8_t* __restrict__ qmask, double* __restrict__ dk,
        unsigned long long* __restrict__ hard_count, unsigned long long* __restrict__ stats) {
    constexpr int CAP = kTileCapOf(KC);                          // entries per query; a tile adds at most 32
    constexpr int ROOM = CAP - 32;
    unsigned long long st_[16];
    if (STATS) for (int i = 0; i < 16; ++i) st_[i] = 0;
    long long t0_ = 0;
#define ST_ADD(i, v) do { if (STATS) st_[i] += (v); } while (0)
#define ST_T0() do { if (STATS) t0_ = clock64(); } while (0)
#define ST_T(i) do { if (STATS) { const long long t1_ = clock64(); st_[i] += t1_ - t0_; t0_ = t1_; } } while (0)
    extern __shared__ uint32_t s_list[];                        // [warp][keys | positions][32 queries][CAP]
    __shared__ uint32_t s_stack[kTileWarps][kTileStack];
    __shared__ float4 s_box[kTileWarps][32];                     // frame-relative FP32 boxes (rounded outwards) of one batch
    __shared__ float s_bext[kTileWarps][32];                     // their largest |coordinate|
    __shared__ float4 s_q[kTileWarps][32];                       // per query: x, y (frame-relative FP32), threshold, #entries
    __shared__ int2 s_rng[BATCH ? kTileWarps : 1][32];           // per query: sorted range of its image
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t base = ((int64_t)blockIdx.x * kTileWarps + warp) * 32;
    if (base >= n) return;
    uint32_t* wk = s_list + (size_t)warp * 2 * CAP * 32;          // query q, entry j: key wk[q * CAP + j]
    int32_t* wi = reinterpret_cast<int32_t*>(wk + CAP * 32);      //                   sorted position wi[q * CAP + j]
    uint32_t* lk = wk + lane * CAP;
    int32_t* li = wi + lane * CAP;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t INFB = 0x7f800000u;
    const float FINF = __int_as_float(INFB);
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const int64_t s = base + lane;

    // ---- the lane's query
    bool act = s < n;
    int pid = 0;
    int64_t image = n_img;
    if (act) {
        pid = vals[s];
        image = (int64_t)((unsigned long long)keys[s] >> (2 * kBits));
        if (image >= n_img) { dk[pid] = 0.0; act = false; }      // outside its image
    }
    int g_lo = 0, g_hi = 0;
    if (act) { g_lo = img_start[image]; g_hi = img_start[image + 1]; }
    {
        const int nsel = g_hi - g_lo;
        if (act && (s - g_lo < (int64_t)nsel * part / n_parts || s - g_lo >= (int64_t)nsel * (part + 1) / n_parts)) act = false;
        if (act && qmask && !(qmask[pid] & 2)) act = false;
        if (act && k > nsel) { dk[pid] = INF; act = false; }
    }
    if (!__any_sync(FULL, act)) return;
    ST_ADD(0, 1); ST_T0();
    // sorted range that can hold candidates of any open lane
    int lo = act ? g_lo : 0x7fffffff, hi = act ? g_hi : 0;
    if (BATCH) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(FULL, lo, o)); hi = max(hi, __shfl_xor_sync(FULL, hi, o)); }
    } else {
        lo = 0;
        hi = img_start[1];
    }
    double x = 0.0, y = 0.0;
    const bool self_ok = s >= lo && s < hi;
    if (self_ok) { const double2 self = pts[s]; x = self.x; y = self.y; }
    const double ox = __shfl_sync(FULL, x, 0), oy = __shfl_sync(FULL, y, 0);   // frame of the FP32 screen
    const float qx = (float)(x - ox), qy = (float)(y - oy);
    // bounding box of the open queries
    double qx0 = act ? x : INF, qx1 = act ? x : -INF, qy0 = act ? y : INF, qy1 = act ? y : -INF;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        qx0 = fmin(qx0, __shfl_xor_sync(FULL, qx0, o)); qx1 = fmax(qx1, __shfl_xor_sync(FULL, qx1, o));
        qy0 = fmin(qy0, __shfl_xor_sync(FULL, qy0, o)); qy1 = fmax(qy1, __shfl_xor_sync(FULL, qy1, o));
    }

    uint32_t c[KC];                                               // the KC smallest keys seen, ascending
#pragma unroll
    for (int i = 0; i < KC; ++i) c[i] = INFB;
    float ext = 0.f, akth = FINF, thr = FINF;
    int cnt = 0, done = 0;
    bool hard = false;

    // sorted insertion without a dependent chain: new[i] = max(old[i - 1], min(old[i], key))
    auto chain_insert = [&](uint32_t key) {
#pragma unroll
        for (int i = KC - 1; i > 0; --i) c[i] = max(c[i - 1], min(c[i], key));
        c[0] = min(c[0], key);
    };
    auto refresh_thr = [&]() {
        if (akth < FINF) {
            const float E = 2.5e-7f * ext + 1e-22f;
            // sqrt(a) as a * rsqrt(a) (2^-22 relative, inside the 1e-6 allowance); a = 0: r = 2 E
            const float r = (akth > 0.f ? akth * __frsqrt_rn(akth) : 0.f) * (1.f + 1e-6f) + 2.f * E;
            thr = r * r * (1.f + 1e-6f) + 1e-37f;
        } else {
            thr = FINF;
        }
        s_q[warp][lane].z = act ? thr : 0.f;                     // what the screen and the node tests read
    };
    auto kth_key = [&]() {
        uint32_t kk = c[KC - 1];
        if (k != KC) {
            kk = c[0];
#pragma unroll
            for (int i = 1; i < KC - 1; ++i) if (i == k - 1) kk = c[i];
        }
        akth = __uint_as_float(kk);
    };
    // lane = candidate (sorted position tb + lane, frame-relative (cx, cy), cok = it exists); every query of `needm`
    // gets the candidates inside its threshold appended to its list
    auto screen = [&](int64_t tb, unsigned needm, float cx, float cy, bool cok) {
        __syncwarp();
        while (needm) {                                           // four queries per round: independent chains
            int qi[4];
            float4 rec[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                qi[u] = needm ? __ffs(needm) - 1 : -1;
                needm &= needm - 1;
                rec[u] = s_q[warp][qi[u] < 0 ? 0 : qi[u]];
                if (qi[u] < 0) rec[u].z = 0.f;
                if (BATCH) {
                    const int2 rg = s_rng[warp][qi[u] < 0 ? 0 : qi[u]];
                    if (!(tb + lane >= rg.x && tb + lane < rg.y)) rec[u].z = 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float dx = cx - rec[u].x, dy = cy - rec[u].y;
                const float d2 = fmaf(dx, dx, dy * dy);
                const bool pass = cok && d2 < rec[u].z;
                const unsigned b = __ballot_sync(FULL, pass);
                const int cq = __float_as_int(rec[u].w);
                if (pass) {
                    const int pos = qi[u] * CAP + cq + __popc(b & lt_mask);
                    wk[pos] = __float_as_uint(d2);
                    wi[pos] = (int32_t)(tb + lane);
                }
                if (lane == 0 && b) s_q[warp][qi[u]].w = __int_as_float(cq + __popc(b));
            }
        }
        __syncwarp();
        cnt = __float_as_int(s_q[warp][lane].w);
    };
    // new list entries -> chain, new threshold; lanes whose bound is inside the caller's floor are finished
    auto tighten = [&]() {
        int j1 = act ? cnt - done : 0;                              // longest run of new entries
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) j1 = max(j1, __shfl_xor_sync(FULL, j1, o));
        ST_ADD(6, j1);
        for (int j = 0; j < j1; ++j) {
            uint32_t key = INFB;
            if (act && done + j < cnt) key = lk[done + j];
            chain_insert(key);
        }
        done = cnt;
        kth_key();
        refresh_thr();
        if (floor2 > 0.0 && act && (double)thr <= floor2) { dk[pid] = sqrt((double)thr); act = false; s_q[warp][lane].z = 0.f; }
    };
    // drop the entries beyond the threshold (all entries are in the chain: call after tighten)
    auto compact = [&]() {
        int j1 = act ? cnt : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) j1 = max(j1, __shfl_xor_sync(FULL, j1, o));
        int w = 0;
        for (int j = 0; j < j1; ++j) {
            if (act && j < cnt) {
                const uint32_t key = lk[j];
                if (__uint_as_float(key) < thr) { const int32_t id = li[j]; lk[w] = key; li[w] = id; ++w; }
            }
        }
        if (act) {
            cnt = done = w;
            if (cnt > ROOM) { hard = true; act = false; s_q[warp][lane].z = 0.f; }
            s_q[warp][lane].w = __int_as_float(cnt);
        }
    };

    // ---- own tile: chain first (tight threshold), then the lists
    s_q[warp][lane] = make_float4(qx, qy, 0.f, __int_as_float(0));
    if (BATCH) s_rng[warp][lane] = make_int2(g_lo, g_hi);
    {
        float m = self_ok ? fmaxf(fabsf(qx), fabsf(qy)) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
        ext = m;
    }
    const float ext_q = ext;                                     // largest |coordinate| of a query of the tile
    {
        const unsigned okm = __ballot_sync(FULL, self_ok);
#pragma unroll 8
        for (int ci = 0; ci < 32; ++ci) {
            const float4 rc = s_q[warp][ci];
            bool ok = act && ((okm >> ci) & 1u);
            if (BATCH) ok = ok && base + ci >= g_lo && base + ci < g_hi;
            const float dx = rc.x - qx, dy = rc.y - qy;
            chain_insert(ok ? __float_as_uint(fmaf(dx, dx, dy * dy)) : INFB);
        }
        kth_key();
        refresh_thr();
        screen(base, __ballot_sync(FULL, act), qx, qy, self_ok);
        done = cnt;
        tighten();
    }
    ST_T(8);

    // ---- the siblings of every ancestor, nearest level on top of the stack
    int sp = 0;
    {
        int top = 1;                                               // first level whose ancestor holds every candidate
        while (top < lv.nlev && !(((base >> (3 + 2 * top)) << (3 + 2 * top)) <= lo && (((base >> (3 + 2 * top)) + 1) << (3 + 2 * top)) >= hi)) ++top;
        for (int level = top - 1; level >= 1; --level) {
            const int sh = 3 + 2 * level;
            const int64_t node = base >> sh;
            const int64_t sb = (node & ~(int64_t)3) + lane;
            const bool push = lane < 4 && sb != node && sb < lv.cnt[level] && (sb << sh) < hi && ((sb + 1) << sh) > lo;
            const unsigned pm = __ballot_sync(FULL, push);
            if (push && sp + 4 <= kTileStack) s_stack[warp][sp + __popc(pm & lt_mask)] = ((uint32_t)level << 28) | (uint32_t)sb;
            sp += __popc(pm);
        }
        __syncwarp();
    }
    int since = 0;
    while (sp > 0 && __any_sync(FULL, act)) {
        int P = min(32, sp);
        P = min(P, max(1, (kTileStack - 48 - sp) / 3));
        uint32_t e = 0xffffffffu;
        if (lane < P) e = s_stack[warp][sp - 1 - lane];
        sp -= P;
        __syncwarp();
        bool need = false;
        int lvl = 0;
        int64_t idx = 0;
        double R2max;
        {
            float tm = act ? thr : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tm = fmaxf(tm, __shfl_xor_sync(FULL, tm, o));
            R2max = (double)tm;
        }
        if (e != 0xffffffffu) {
            lvl = (int)(e >> 28);
            idx = (int64_t)(e & 0x0fffffffu);
            const int nsh = 3 + 2 * lvl;
            if ((idx << nsh) < hi && ((idx + 1) << nsh) > lo) {
                const KnnBox* nb = bb + lv.off[lvl] + idx;
                const double2 blo = __ldg(reinterpret_cast<const double2*>(nb));
                const double2 bhi = __ldg(reinterpret_cast<const double2*>(nb) + 1);
                const double dx = qx1 < blo.x ? blo.x - qx1 : (qx0 > bhi.x ? qx0 - bhi.x : 0.0);
                const double dy = qy1 < blo.y ? blo.y - qy1 : (qy0 > bhi.y ? qy0 - bhi.y : 0.0);
                need = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) < R2max;
                if (need) {
                    // the box in the FP32 frame, pushed outwards by more than the rounding of its corners, of the
                    // queries' coordinates and of the subtraction below: the lanes' own tests stay conservative
                    const float fx0 = (float)(blo.x - ox), fy0 = (float)(blo.y - oy);
                    const float fx1 = (float)(bhi.x - ox), fy1 = (float)(bhi.y - oy);
                    const float ma = fmaxf(fmaxf(fabsf(fx0), fabsf(fx1)), fmaxf(fabsf(fy0), fabsf(fy1)));
                    const float mg = 2.5e-7f * fmaxf(ma, ext_q) + 1e-30f;
                    s_box[warp][lane] = make_float4(fx0 - mg, fy0 - mg, fx1 + mg, fy1 + mg);
                    s_bext[warp][lane] = ma + mg;
                }
            }
        }
        // nodes that pass the tile-wide test, in stack order (nearest level first); each must also be reached by
        // the ball of at least one query: leaf tiles are then used at once (which tightens the radii for the nodes
        // after them), internal nodes have their children pushed at the end of the batch
        unsigned pm = __ballot_sync(FULL, need);
        const unsigned leaf_mask = __ballot_sync(FULL, need && lvl == 1);
        unsigned pushm = 0;
        ST_ADD(1, 1); ST_ADD(2, P); ST_ADD(3, __popc(pm));
        __syncwarp();
        ST_T(9);
        // the candidates of the next leaf tile are loaded while the current node is processed
        int nleaf = (pm & leaf_mask) ? __ffs(pm & leaf_mask) - 1 : -1;
        int64_t tb_next = 0;
        double2 p_next = make_double2(0.0, 0.0);
        if (nleaf >= 0) {
            tb_next = __shfl_sync(FULL, idx, nleaf) << 5;
            if (tb_next + lane >= lo && tb_next + lane < hi) p_next = __ldg(pts + tb_next + lane);
        }
        while (pm) {
            const int cur = __ffs(pm) - 1;
            pm &= pm - 1;
            const bool is_leaf = (leaf_mask >> cur) & 1u;
            const int64_t tb = tb_next;
            const double2 p = p_next;
            if (is_leaf) {
                nleaf = (pm & leaf_mask) ? __ffs(pm & leaf_mask) - 1 : -1;
                if (nleaf >= 0) {
                    tb_next = __shfl_sync(FULL, idx, nleaf) << 5;
                    if (tb_next + lane >= lo && tb_next + lane < hi) p_next = __ldg(pts + tb_next + lane);
                }
            }
            bool mneed = false;
            {
                const float4 bx = s_box[warp][cur];
                const float dx = fmaxf(fmaxf(bx.x - qx, qx - bx.z), 0.f), dy = fmaxf(fmaxf(bx.y - qy, qy - bx.w), 0.f);
                mneed = act && fmaf(dx, dx, dy * dy) * (1.f - 1e-6f) < thr;
            }
            if (BATCH && is_leaf) mneed = mneed && tb < g_hi && tb + 32 > g_lo;
            const unsigned needm = __ballot_sync(FULL, mneed);
            ST_T(10);
            if (!needm) continue;
            if (!is_leaf) { pushm |= 1u << cur; continue; }
            ST_ADD(4, 1); ST_ADD(7, __popc(needm));
            // the tile's 32 candidates, one per lane
            const bool cok = tb + lane >= lo && tb + lane < hi;
            float cx = 0.f, cy = 0.f;
            if (cok) { cx = (float)(p.x - ox); cy = (float)(p.y - oy); }
            {
                const float m = s_bext[warp][cur];                // bounds every |coordinate| of the tile
                if (m > ext) { ext = m; refresh_thr(); }
            }
            ST_T(11);
            if (__any_sync(FULL, act && cnt > ROOM)) { if (since) tighten(); compact(); since = 0; ST_ADD(5, 1); ST_T(14); }
            screen(tb, needm & __ballot_sync(FULL, act), cx, cy, cok);
            ST_T(12);
            if (++since >= tighten_every) { tighten(); since = 0; }
            ST_T(13);
        }
        if (pushm) {
            const bool internal = (pushm >> lane) & 1u;
            if (internal) {
                const int at = sp + 4 * __popc((pushm >> lane) >> 1);      // the nearest node's children on top
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const int64_t ci = 4 * idx + ch;
                    s_stack[warp][at + ch] = ci < lv.cnt[lvl - 1] ? (((uint32_t)(lvl - 1) << 28) | (uint32_t)ci) : 0xffffffffu;
                }
            }
            sp += 4 * __popc(pushm);
        }
        __syncwarp();
    }
    ST_T(9);

    // ---- exact float64 distances of the entries inside the final threshold
    {
        if (since) tighten();
        compact();
        int j1 = act ? cnt : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) j1 = max(j1, __shfl_xor_sync(FULL, j1, o));
        double dmax = 0.0;
        for (int j = 0; j < j1; ++j) {
            if (act && j < cnt) {
                const double2 p = __ldg(pts + li[j]);
                const double dx = p.x - x, dy = p.y - y;
                dmax = fmax(dmax, __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            }
        }
        if (act) {
            if (cnt == k) {
                dk[pid] = sqrt(dmax);
            } else if (cnt > k) {
                // near-ties of the k-th distance: the k-th smallest of the exact values by rank
                double ans = dmax;
                for (int a = 0; a < cnt; ++a) {
                    const double2 pa = __ldg(pts + li[a]);
                    const double ax = pa.x - x, ay = pa.y - y;
                    const double da = __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay));
                    int lt = 0, eq = 0;
                    for (int b = 0; b < cnt; ++b) {
                        const double2 pb = __ldg(pts + li[b]);
                        const double bx = pb.x - x, by = pb.y - y;
                        const double db = __dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by));
                        lt += db < da;
                        eq += db == da;
                    }
                    if (lt < k && k <= lt + eq) { ans = da; break; }
                }
                dk[pid] = sqrt(ans);
            } else {
                hard = true;
            }
        }
    }
    ST_T(15);
    if (hard) {
        dk[pid] = __longlong_as_double(0x7ff8000000000000LL);
        atomicAdd(hard_count, 1ull);
    }
    if (STATS && lane == 0) for (int i = 0; i < 16; ++i) atomicAdd(stats + i, st_[i]);
#undef ST_ADD
#undef ST_T0
#undef ST_T
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
    static int dbg = -1;      // SO_KNN_DEBUG=1: synchronise and report after every stage (locates a failing or hanging stage)
    if (dbg < 0) { const char* e = getenv("SO_KNN_DEBUG"); dbg = (e && e[0] == '1') ? 1 : 0; }
#define KNN_DBG(what) do { if (dbg) { cudaError_t e_ = cudaStreamSynchronize(st); fprintf(stderr, "[knn] %s: %s\n", what, cudaGetErrorString(e_)); } } while (0)
    SO_CUDA(cudaMemsetAsync(l.counter, 0, 2 * sizeof(unsigned long long), st));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    int64_t key_blocks = (n + kKeyThreads - 1) / kKeyThreads;
    if (key_blocks > 8 * so_num_sms()) key_blocks = 8 * so_num_sms();
    { SO_PROF("knn_key_kernel", st); knn_key_kernel<KeyT><<<(unsigned)key_blocks, kKeyThreads, 0, st>>>(X, Y, ix, img, n, n_img, half_width, inv_h, keys0,
                                                                     l.vals0, l.xy, l.counter); }
    SO_LAUNCH_CHECK();
    size_t tmp = (size_t)l.cub_tmp_bytes;
    { SO_PROF("cub sort Morton keys", st); SO_CUDA(cub::DeviceRadixSort::SortPairs(l.cub_tmp, tmp, keys0, keys1, l.vals0, l.vals1, (int)n, 0,
                                            2 * kBits + img_bits, st)); }
    KNN_DBG("sort");
    { SO_PROF("knn_gather_kernel", st); knn_gather_kernel<<<blocks, 256, 0, st>>>(l.xy, l.vals1, n, l.pts); }
    SO_LAUNCH_CHECK();
    KNN_DBG("gather");
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
    // SO_KNN_ALGO=quadtree / tiles forces one search for every call (A/B aid).  Default: the tile kernel for calls of at
    // most kTileAutoMax particles, where a launch is latency-bound and a warp per 32 queries finishes sooner than a thread
    // per query, the quadtree walks above that (measured on B200, SO_KNN_ALGO A/B in tools/ab_knn.py)
    static int algo = -1;
    if (algo < 0) { const char* e = getenv("SO_KNN_ALGO"); algo = !e ? 0 : (e[0] == 'q' ? 1 : (e[0] == 't' ? 2 : 0)); }
    static int64_t auto_max = -1;
    if (auto_max < 0) { const char* e = getenv("SO_KNN_TILE_MAX"); auto_max = e ? atoll(e) : kTileAutoMax; }
    const bool tiles = k <= 32 && (algo == 2 || (algo == 0 && n <= auto_max));
    const double floor2 = floor > 0.0 ? floor * floor : -1.0;
    const unsigned long long* gate = nullptr;
    if (tiles) {
        const KnnLevels lv = knn_levels(n);
        { SO_PROF("knn_bvh_leaf_kernel", st); knn_bvh_leaf_kernel<KeyT><<<(unsigned)((n + kBvhThreads - 1) / kBvhThreads), kBvhThreads, 0, st>>>(
            l.pts, keys1, n, n_img, l.bb, lv, l.img_start); }
        SO_LAUNCH_CHECK();
        if (lv.nlev > 4) {
            SO_PROF("knn_bvh_upper_kernel", st); knn_bvh_upper_kernel<<<1, 1024, 0, st>>>(l.bb, lv);
        }
        SO_LAUNCH_CHECK();
        const unsigned tb = (unsigned)((n + 32 * kTileWarps - 1) / (32 * kTileWarps));
#define SO_KNN_TILE_(KC, BATCH, STATS)                                                                                              \
        do {                                                                                                                \
            const int sm = kTileWarps * 2 * kTileCapOf(KC) * 32 * (int)sizeof(uint32_t);                                    \
            SO_CUDA(cudaFuncSetAttribute(knn_tile_kernel<KeyT, KC, BATCH, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)); \
            SO_PROF("knn_tile_kernel", st);                                                                                 \
            knn_tile_kernel<KeyT, KC, BATCH, STATS><<<tb, kTileWarps * 32, sm, st>>>(l.pts, keys1, l.vals1, n, n_img, l.img_start, \
                l.bb, lv, k, part, n_parts, floor2, tighten_every, qmask, dk, l.counter + 1, stats);                                       \
        } while (0)
        static int tighten_every = 0;    // SO_KNN_TIGHTEN: tiles between two threshold updates (A/B aid)
        if (tighten_every == 0) { const char* e = getenv("SO_KNN_TIGHTEN"); tighten_every = e ? atoi(e) : 1; if (tighten_every < 1) tighten_every = 1; }
        static int want_stats = -1;      // SO_KNN_STATS=1: per-phase cycle and event counters of the tile kernel (k <= 8, one image)
        if (want_stats < 0) { const char* e = getenv("SO_KNN_STATS"); want_stats = (e && e[0] == '1') ? 1 : 0; }
        unsigned long long* stats = nullptr;
#define SO_KNN_TILE(KC, BATCH) SO_KNN_TILE_(KC, BATCH, false)
        if (want_stats && n_img == 1 && k <= 8) {
            SO_CUDA(cudaMalloc(&stats, 16 * sizeof(unsigned long long)));
            SO_CUDA(cudaMemsetAsync(stats, 0, 16 * sizeof(unsigned long long), st));
            SO_KNN_TILE_(8, false, true);
            unsigned long long h[16];
            SO_CUDA(cudaMemcpyAsync(h, stats, sizeof(h), cudaMemcpyDeviceToHost, st));
            SO_CUDA(cudaStreamSynchronize(st));
            SO_CUDA(cudaFree(stats));
            const double t = (double)(h[0] ? h[0] : 1);
            fprintf(stderr, "[knn tiles] tiles %llu | per tile: batches %.2f nodes %.2f leaves tested %.2f staged %.2f (lanes needing %.1f) "
                    "compactions %.3f chain iterations %.1f | cycles: own %.0f walk %.0f leaf test %.0f stage %.0f screen %.0f "
                    "tighten %.0f compact %.0f final %.0f\n", h[0], h[1] / t, h[2] / t, h[3] / t, h[4] / t, h[4] ? (double)h[7] / h[4] : 0.0,
                    h[5] / t, h[6] / t, h[8] / t, h[9] / t, h[10] / t, h[11] / t, h[12] / t, h[13] / t, h[14] / t, h[15] / t);
        } else
        if (n_img > 1) { if (k <= 8) SO_KNN_TILE(8, true); else if (k <= 16) SO_KNN_TILE(16, true); else SO_KNN_TILE(32, true); }
        else { if (k <= 8) SO_KNN_TILE(8, false); else if (k <= 16) SO_KNN_TILE(16, false); else SO_KNN_TILE(32, false); }
#undef SO_KNN_TILE
#undef SO_KNN_TILE_
        SO_LAUNCH_CHECK();
        gate = l.counter + 1;      // the quadtree walk below answers the queries the tile kernel gave up on (usually none)
    }
    // quadtree walks: table of node ranges
    SO_CUDA(cudaMemsetAsync(l.E, 0, sizeof(int32_t) * (size_t)ncell, st));
    { SO_PROF("knn_mark_kernel", st); knn_mark_kernel<KeyT><<<blocks, 256, 0, st>>>(keys1, n, n_img, 2 * (kBits - Lt), l.E, gate); }
    SO_LAUNCH_CHECK();
    tmp = (size_t)l.cub_tmp_bytes;
    SO_CUDA(cub::DeviceScan::InclusiveScan(l.cub_tmp, tmp, l.E, l.E, cub::Max(), (int)ncell, st));
    KNN_DBG("table");
    const bool private_walk = gate || (force_private >= 0 ? force_private == 1 : k <= 8);
    // half-width of the window of sorted neighbours scanned first (SO_KNN_WIN = multiple of k, default 1)
    static int win_mul = 0;
    if (win_mul == 0) { const char* e = getenv("SO_KNN_WIN"); win_mul = e ? atoi(e) : 1; if (win_mul < 1) win_mul = 1; }
    const int win = k * win_mul;
    if (private_walk || threads != kWarpsPerCta * 32) {
        SO_CUDA(cudaFuncSetAttribute(knn_query_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        { SO_PROF("knn_query_kernel", st); knn_query_kernel<KeyT><<<(unsigned)((n + threads - 1) / threads), threads, smem, st>>>(
            l.pts, keys1, l.vals1, n, n_img, l.E, Lt, half_width, inv_h, k, part, n_parts, leaf, win, floor2, qmask, dk, gate); }
    } else {
        SO_CUDA(cudaFuncSetAttribute(knn_query_warp_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        { SO_PROF("knn_query_warp_kernel", st); knn_query_warp_kernel<KeyT><<<(unsigned)((n + threads - 1) / threads), threads, smem, st>>>(
            l.pts, keys1, l.vals1, n, n_img, l.E, Lt, half_width, inv_h, k, part, n_parts, leaf, win, floor2, qmask, dk); }
    }
    SO_LAUNCH_CHECK();
    KNN_DBG("query");
#undef KNN_DBG
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

// Imaging hot path (sm_100a): nearest-pixel index, tile binning, kernel-smoothed
// deposition, rebin.
//
// Reference behaviour reproduced (paths relative to the reference root):
//   synthobs/morph/images.py:47-50   selection |x|,|y| < width/2
//   synthobs/morph/images.py:66-69   NGP histogram (simple, hist)
//   synthobs/morph/images.py:87-97   adaptive Gaussian deposit, normalised by its sum over the image
//   synthobs/morph/images.py:19-22   rebin (block sum)
//
// Design: the Gaussian is separable, so one particle's contribution to a
// block of pixels is a rank-1 update  block += (w * ey) (x) ex.  Particles are binned per 16x16
// sub-tile with a stable radix sort (deterministic lists).  One WARP owns one work item (sub-tile,
// slice of its list): lane l owns rows 2*(l/4), +1 and columns 4*(l%4)..+3 for all C channels in
// registers; per particle the 16 column factors and 16 row factors are evaluated once (one ex2 per
// lane), exchanged through shared memory, and every lane does 8*C packed multiply-adds.  The
// particle data of a list entry is gathered as two 32-byte sectors (plan record, float32 weight
// row).  Each sub-tile is written exactly
// once with plain stores; sub-tiles split over several items are summed by a second kernel in
// fixed order.  No global atomics touch pixels; results are bitwise reproducible.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int kT = SO_TILE;          // sub-tile edge (16)
constexpr int kDepWarps = 4;         // warps (= work items) per deposit CTA
constexpr int kP = 32;               // particles staged per phase (one per lane)
constexpr int kWideW = 48;           // supports wider than this get a warp for their normalisation sums

__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Gaussian factor of pixel offset k from the nearest pixel, relative to the nearest-pixel term
// (exponent shifted so narrow particles survive FP32): exp(-((k - u)^2 - u^2) / 2 sigma^2 * pitch^2)
__device__ __forceinline__ float gauss_factor(float cpx, int k, float u) {
    const float fk = (float)k;
    return ex2_ftz(-cpx * fk * (fk - 2.f * u));
}

// per-particle deposit plan: one 32-byte record = one memory sector per gather
struct __align__(32) PlanRec {
    float ux, uy;     // (x - G[ix]) / pitch, (y - G[iy]) / pitch
    float cpx;        // log2(e) * pitch^2 / (2 sigma^2)
    float wnorm;      // 1 / (Sx * Sy), 0 for dropped particles
    int32_t ix, iy;   // nearest pixel (ix < 0: not selected)
    int32_t W;        // support half width in pixels
    int32_t pad;
};

struct PlanView {
    PlanRec* rec;     // [n]
    int64_t* offs;    // [n+1] pair offsets (exclusive scan of the tile counts)
    unsigned long long* stats;  // [4] pairs, deposits, selected, dropped
    void* cub_tmp;
    int64_t cub_tmp_bytes;
};

__device__ __forceinline__ void plan_load(const PlanRec* __restrict__ rec, int64_t p, float4& a, int4& b) {
    a = __ldg(reinterpret_cast<const float4*>(rec + p));
    b = __ldg(reinterpret_cast<const int4*>(rec + p) + 1);
}

static int64_t scan_tmp_bytes(int64_t n) {
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum((void*)nullptr, b, (int64_t*)nullptr, (int64_t*)nullptr, (int)(n + 1));
    return (int64_t)b;
}

static PlanView carve_plan(void* ws, int64_t bytes, int64_t n, bool* ok) {
    SoArena a(ws, bytes);
    PlanView v;
    v.rec = a.take<PlanRec>(n);
    v.offs = a.take<int64_t>(n + 1);
    v.stats = a.take<unsigned long long>(4);
    v.cub_tmp_bytes = scan_tmp_bytes(n);
    v.cub_tmp = a.take<char>(v.cub_tmp_bytes);
    if (ok) *ok = a.ok();
    return v;
}

__global__ void ngp_index_kernel(const double* __restrict__ X, const double* __restrict__ Y, int64_t n,
                                 const double* __restrict__ G, int ndim, double hw,
                                 int32_t* __restrict__ ix, int32_t* __restrict__ iy) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = X[i], y = Y[i];
    const bool sel = (fabs(x) < hw) && (fabs(y) < hw);   // images.py:47 (strict)
    int rx = -1, ry = -1;
    if (sel) {
        const double g0 = G[0];
        const double inv = (ndim > 1) ? (double)(ndim - 1) / (G[ndim - 1] - g0) : 0.0;
        const double v[2] = {x, y};
        int r[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            int k0 = (int)floor((v[a] - g0) * inv + 0.5);
            k0 = max(0, min(ndim - 1, k0));
            // literal first-minimum argmin over the neighbourhood of the closed-form guess,
            // on the same float64 G values the reference compares (images.py:67)
            int lo = max(0, k0 - 2), hi = min(ndim - 1, k0 + 2);
            int best = lo;
            double bd = fabs(G[lo] - v[a]);
            for (int k = lo + 1; k <= hi; ++k) {
                const double d = fabs(G[k] - v[a]);
                if (d < bd) { bd = d; best = k; }
            }
            r[a] = best;
        }
        rx = r[0]; ry = r[1];
    }
    ix[i] = rx; iy[i] = ry;
}

// one thread per particle: support box, normalisation sums, sub-tile count.  Particles whose support is
// wider than kWideW pixels only get their box here; plan_wide_kernel sums their factors with a warp.
__global__ void __launch_bounds__(256) plan_kernel(const int32_t* __restrict__ ix_in, const int32_t* __restrict__ iy_in,
                                                   const double* __restrict__ X, const double* __restrict__ Y,
                                                   const double* __restrict__ dk, int64_t n,
                                                   const double* __restrict__ G, int ndim, double R,
                                                   PlanView pv) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long my_dep = 0, my_sel = 0, my_drop = 0, my_pairs = 0;
    if (p < n) {
        const int ip = ix_in[p], jp = iy_in[p];
        int64_t count = 0;
        float ux = 0.f, uy = 0.f, cpx = 0.f, wn = 0.f;
        int W = 0;
        if (ip >= 0) {
            const double pitch = (ndim > 1) ? (G[ndim - 1] - G[0]) / (double)(ndim - 1) : 1.0;
            const double dx0 = X[p] - G[ip], dy0 = Y[p] - G[jp];
            bool dropped = false;
            if (dk) {
                const double fwhm = fmax(dk[p], 0.01);               // images.py:89
                const double sigma = __ddiv_rn(fwhm, 2.355);         // images.py:91
                const double two_s2 = __dmul_rn(2.0, __dmul_rn(sigma, sigma));
                // the reference drops the particle when every float64 term underflows (images.py:95-97)
                const double amin = __ddiv_rn(__dadd_rn(__dmul_rn(dx0, dx0), __dmul_rn(dy0, dy0)), two_s2);
                // np.exp(-a) == 0.0 exactly for a >= 0x1.74910d52d3052p+9 (745.1332191019412: below half the smallest
                // denormal); established by bisection against NumPy, so no device exp is needed (or trusted) here
                dropped = !(amin < 745.1332191019412);
                const double c = 1.4426950408889634 * pitch * pitch / two_s2;
                cpx = (float)c;
                if (R > 0.0) {
                    const double w = ceil(__ddiv_rn(__dmul_rn(R, sigma), pitch));
                    W = (w >= (double)ndim) ? ndim : (int)w;
                } else {
                    W = ndim;
                }
                if (!(sigma < __longlong_as_double(0x7ff0000000000000LL))) W = ndim;
            }
            ux = (float)(dx0 / pitch);
            uy = (float)(dy0 / pitch);
            if (dropped || !dk) W = 0;
            const int kx0 = max(-W, -ip), kx1 = min(W, ndim - 1 - ip);
            const int ky0 = max(-W, -jp), ky1 = min(W, ndim - 1 - jp);
            if (dk && !dropped) {
                if (W <= kWideW) {
                    float sx = 0.f, sy = 0.f;
                    for (int k = kx0; k <= kx1; ++k) sx += gauss_factor(cpx, k, ux);
                    for (int k = ky0; k <= ky1; ++k) sy += gauss_factor(cpx, k, uy);
                    wn = 1.0f / (sx * sy);
                } else {
                    wn = -1.0f;     // marker: plan_wide_kernel fills it in
                }
                my_dep = (unsigned long long)(kx1 - kx0 + 1) * (unsigned long long)(ky1 - ky0 + 1);
            }
            const int tx0 = (ip + kx0) / kT, tx1 = (ip + kx1) / kT;
            const int ty0 = (jp + ky0) / kT, ty1 = (jp + ky1) / kT;
            count = (int64_t)(tx1 - tx0 + 1) * (ty1 - ty0 + 1);
            my_sel = 1;
            my_drop = dropped ? 1 : 0;
            my_pairs = (unsigned long long)count;
        }
        float4* out = reinterpret_cast<float4*>(pv.rec + p);
        out[0] = make_float4(ux, uy, cpx, wn);
        reinterpret_cast<int4*>(out)[1] = make_int4(ip, jp, W, 0);
        pv.offs[p] = count;
        if (p == n - 1) pv.offs[n] = 0;
    }
    // statistics (integer atomics on 4 counters; never on pixels)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_dep += __shfl_xor_sync(0xffffffffu, my_dep, o);
        my_sel += __shfl_xor_sync(0xffffffffu, my_sel, o);
        my_drop += __shfl_xor_sync(0xffffffffu, my_drop, o);
        my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, o);
    }
    // one update of the four counters per CTA (same-address atomics from every warp serialise)
    __shared__ unsigned long long s_stats[4];
    if (threadIdx.x < 4) s_stats[threadIdx.x] = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && my_sel) {
        atomicAdd(&s_stats[0], my_pairs); atomicAdd(&s_stats[1], my_dep);
        atomicAdd(&s_stats[2], my_sel); atomicAdd(&s_stats[3], my_drop);
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_stats[threadIdx.x]) atomicAdd(&pv.stats[threadIdx.x], s_stats[threadIdx.x]);
}

// supports wider than kWideW (marked wnorm < 0): each warp looks at 32 consecutive particles and sums the
// factors of the marked ones cooperatively, one particle at a time
__global__ void __launch_bounds__(256) plan_wide_kernel(int64_t n, int ndim, PlanView pv) {
    const int lane = threadIdx.x & 31;
    const int64_t base = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 5;
    if (base >= n) return;
    const int64_t mine = base + lane;
    unsigned todo = __ballot_sync(0xffffffffu, mine < n && pv.rec[mine].wnorm < 0.f);
    while (todo) {
        const int64_t p = base + (__ffs(todo) - 1);
        todo &= todo - 1;
        const PlanRec r = pv.rec[p];
        const int ip = r.ix, jp = r.iy, W = r.W;
        const float cpx = r.cpx, ux = r.ux, uy = r.uy;
        const int kx0 = max(-W, -ip), kx1 = min(W, ndim - 1 - ip);
        const int ky0 = max(-W, -jp), ky1 = min(W, ndim - 1 - jp);
        float sx = 0.f, sy = 0.f;
        for (int k = kx0 + lane; k <= kx1; k += 32) sx += gauss_factor(cpx, k, ux);
        for (int k = ky0 + lane; k <= ky1; k += 32) sy += gauss_factor(cpx, k, uy);
        sx = warp_sum(sx);
        sy = warp_sum(sy);
        __syncwarp();
        if (lane == 0) pv.rec[p].wnorm = 1.0f / (sx * sy);
    }
}

// (tile, pid) pairs at offs[p].. (row-major over the particle's tile box).  One thread per particle writes
// short lists itself; the particles of a warp with more than kExpandSerial pairs are then expanded by
// the whole warp, one particle at a time
constexpr int kExpandSerial = 8;
__global__ void __launch_bounds__(256) expand_kernel(PlanView pv, int64_t n, int ndim, int ntx,
                                                     const int32_t* __restrict__ img, int ntiles_img,
                                                     int32_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int tx0 = 0, ty0 = 0, nx = 0, cnt = 0, kb = 0;     // kb: first tile key of the particle's image
    int64_t o = 0;
    if (p < n) {
        const int4 r = __ldg(reinterpret_cast<const int4*>(pv.rec + p) + 1);
        const int ip = r.x;
        if (ip >= 0) {
            kb = img ? img[p] * ntiles_img : 0;
            const int jp = r.y, W = r.z;
            const int kx0 = max(-W, -ip), kx1 = min(W, ndim - 1 - ip);
            const int ky0 = max(-W, -jp), ky1 = min(W, ndim - 1 - jp);
            tx0 = (ip + kx0) / kT;
            ty0 = (jp + ky0) / kT;
            nx = (ip + kx1) / kT - tx0 + 1;
            cnt = nx * ((jp + ky1) / kT - ty0 + 1);
            o = pv.offs[p];
            if (cnt <= kExpandSerial)
                for (int q = 0; q < cnt; ++q) {
                    keys[o + q] = kb + (ty0 + q / nx) * ntx + tx0 + q % nx;
                    vals[o + q] = (int32_t)p;
                }
        }
    }
    unsigned todo = __ballot_sync(0xffffffffu, cnt > kExpandSerial);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int c = __shfl_sync(0xffffffffu, cnt, src), w = __shfl_sync(0xffffffffu, nx, src);
        const int x0 = __shfl_sync(0xffffffffu, tx0, src), y0 = __shfl_sync(0xffffffffu, ty0, src);
        const int k0 = __shfl_sync(0xffffffffu, kb, src);
        const int64_t oo = __shfl_sync(0xffffffffu, o, src);
        const int32_t pid = (int32_t)(p - lane + src);
        for (int q = lane; q < c; q += 32) {
            keys[oo + q] = k0 + (y0 + q / w) * ntx + x0 + q % w;
            vals[oo + q] = pid;
        }
    }
}

// tile_start[t] = first sorted pair with key >= t  (t in [0, ntiles])
__global__ void tile_start_kernel(const int32_t* __restrict__ keys, int64_t n_pairs, int ntiles,
                                  int64_t* __restrict__ tile_start) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    int64_t lo = 0, hi = n_pairs;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < t) lo = mid + 1; else hi = mid;
    }
    tile_start[t] = lo;
}

// items per tile -> exclusive scan (device-wide; [ntiles + 1] entries, the last one = number of items) ; item -> tile map
__global__ void tile_item_count_kernel(const int64_t* __restrict__ tile_start, int ntiles, int64_t chunk,
                                       int64_t* __restrict__ item_start) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    item_start[t] = t < ntiles ? (tile_start[t + 1] - tile_start[t] + chunk - 1) / chunk : 0;
}

__global__ void item_tile_kernel(const int64_t* __restrict__ item_start, int ntiles, int64_t max_items,
                                 int32_t* __restrict__ item_tile) {
    const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= max_items || it >= item_start[ntiles]) return;
    int lo = 0, hi = ntiles;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (item_start[mid] <= it) lo = mid; else hi = mid;
    }
    item_tile[it] = lo;
}

struct DepositParams {
    PlanView pv;
    const double* L;          // [C_total, n]
    const float* wts;         // [n, C_total] float32 weights L * wnorm (written by weights_kernel)
    int64_t n;
    int ndim, ntx, ntiles;
    const int32_t* pids;      // sorted pair list
    const int64_t* tile_start;
    const int64_t* item_start;
    const int32_t* item_tile;
    int64_t chunk;
    float* data;              // [n_img, C_total] images of ndim x ndim, row stride ld, image stride plane (floats)
    int64_t ld, plane;
    float* partial;           // [max_items, C_total, kT*kT]
    int n_chan;               // C_total
    int chan0;                // first channel of this launch
    int ntiles_img;           // sub-tiles per image; tile key = image * ntiles_img + tile (batch of images)
    // several GPUs, one image (so_img_deposit_push): a sub-tile only THIS rank touches is also stored into every rank's
    // result image over NVLink while the deposit runs; push_n = 0: off
    int push_n;
    float* push_result[SO_MAX_PEERS];          // the ranks' result images, same layout as data
    const unsigned char* push_masks;           // [push_n][ntiles]: which ranks touch which sub-tile (this rank's copy)
    unsigned long long* visits;                // so_img_deposit_count: [0] += (lane, particle) visits, [1] += rounds x 32
};

// true when no other rank of a push deposit touches the sub-tile
__device__ __forceinline__ bool tile_exclusive(const DepositParams& p, int tile) {
    int cnt = 0;
    for (int r = 0; r < p.push_n; ++r) cnt += p.push_masks[(size_t)r * p.ntiles + tile] ? 1 : 0;
    return cnt == 1;
}

// wts[p][c] = (float)(L[c][p] * wnorm[p]): the weights of a particle in one sector, so the deposit gathers
// two sectors per (particle, sub-tile) pair (plan record + weights) instead of one per array and channel
__global__ void __launch_bounds__(256) weights_kernel(const double* __restrict__ L, const PlanRec* __restrict__ rec,
                                                      int64_t n, int n_chan, float* __restrict__ wts) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double wn = (double)rec[p].wnorm;
    for (int c = 0; c < n_chan; ++c) wts[(size_t)p * n_chan + c] = (float)(L[(size_t)c * n + p] * wn);
}

// 32 x 32 bit-matrix transpose across the lanes of a warp: bit j of the result in lane i = bit i of x in lane j
__device__ __forceinline__ unsigned warp_transpose32(unsigned x, int lane) {
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const unsigned m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
        const unsigned y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
    }
    return x;
}

// One warp = one work item (sub-tile, slice of its particle list), C channels in registers.
// Per particle the 16 column factors and 16 row factors of the sub-tile are evaluated one per lane and
// exchanged through shared memory (all 32 staged particles at once, one __syncwarp per phase); a lane
// owns a 2 x 4 pixel block, so it reads 4 column and 2 row factors per particle, and all multiply-adds are
// packed (FMUL2 / FFMA2).
// A narrow particle (support of a few pixels: the bulk of a dense image) overlaps only a few of the 32 pixel
// blocks, so each lane walks ITS OWN list of the staged particles whose support box meets its block: the lane
// that stages particle q computes the 32-bit set of lanes the particle touches from its support box, a bit-matrix
// transpose turns the 32 sets into one particle mask per lane, and the multiply-add loop runs until the longest
// list is done (about 14 instead of 32 rounds at ~18 pixels per particle; all 32 for wide supports).  The order
// in which a pixel receives its contributions is still the order of the list: bitwise reproducible.
template <int C, bool PUSH, bool COUNT = false>
__global__ void __launch_bounds__(kDepWarps * 32) deposit_kernel(const DepositParams p) {
    constexpr int CP = (C == 8) ? 12 : C;                      // padded weight row: lanes read different rows at once
    constexpr int FP = 2 * kT + 4;                             // padded factor row
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * kDepWarps + warp;
    if (item >= p.item_start[p.ntiles]) return;
    const int tile = p.item_tile[item];
    const int64_t t0 = p.tile_start[tile], t1 = p.tile_start[tile + 1];
    const int64_t start = t0 + (item - p.item_start[tile]) * p.chunk;
    const int64_t end = min(t1, start + p.chunk);
    const bool single = (p.item_start[tile + 1] - p.item_start[tile]) == 1;
    const bool push = PUSH && single && tile_exclusive(p, tile);      // (PUSH = false: the single-GPU code without these branches)
    const int image = tile / p.ntiles_img, t_in = tile - image * p.ntiles_img;
    const int py0 = (t_in / p.ntx) * kT, px0 = (t_in % p.ntx) * kT;

    __shared__ float4 s_rec[kDepWarps][2][kP];                 // per axis: {-2u, -cpx, tile origin - nearest pixel, W}
    __shared__ __align__(16) float s_w[kDepWarps][kP][CP];     // weights
    __shared__ __align__(16) float s_f[kDepWarps][kP][FP];     // factors per staged particle: 16 columns, then 16 rows

    // lane -> one factor to evaluate (lanes 0-15: column px0+lane, lanes 16-31: row py0+lane-16) and
    // a block of 2 rows x 4 columns of owned pixels
    const int axis = lane >> 4, fl = lane & 15;                // 0: x (columns), 1: y (rows)
    const bool fvalid = (axis ? py0 : px0) + fl < p.ndim;
    const float flf = (float)fl;
    const int br = lane >> 2, bc = lane & 3;                   // rows 2 br, 2 br + 1; columns 4 bc .. 4 bc + 3
    float2 acc[C][4];                                          // [channel][2 * row + column pair]
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[c][j] = make_float2(0.f, 0.f);

    unsigned n_visits = 0, n_rounds = 0;                       // COUNT only
    int pid_next = (start + lane < end) ? p.pids[start + lane] : 0;     // list entries are fetched one phase ahead
    for (int64_t base = start; base < end; base += kP) {
        const int np = (int)min((int64_t)kP, end - base);
        const int pid = pid_next;
        if (base + kP + lane < end) pid_next = p.pids[base + kP + lane];
        __syncwarp();
        unsigned touched = 0;                                  // lanes (pixel blocks) the particle of this lane overlaps
        if (lane < np) {
            float4 ra;
            int4 rb;
            plan_load(p.pv.rec, pid, ra, rb);
            const float Wf = (float)rb.z;
            s_rec[warp][0][lane] = make_float4(-2.f * ra.x, -ra.z, (float)(px0 - rb.x), Wf);
            s_rec[warp][1][lane] = make_float4(-2.f * ra.y, -ra.z, (float)(py0 - rb.y), Wf);
            const float* wp = p.wts + (size_t)pid * p.n_chan + p.chan0;
            if constexpr (C % 4 == 0) {
                if ((p.n_chan & 3) == 0) {          // chan0 is a multiple of 4 then: 16-byte loads
#pragma unroll
                    for (int c = 0; c < C; c += 4)
                        *reinterpret_cast<float4*>(&s_w[warp][lane][c]) = __ldg(reinterpret_cast<const float4*>(wp + c));
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) s_w[warp][lane][c] = __ldg(wp + c);
                }
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) s_w[warp][lane][c] = __ldg(wp + c);
            }
            // support box [nearest - W, nearest + W]^2 in sub-tile coordinates -> pixel blocks (2 rows x 4 columns)
            const int W = rb.z, cx = rb.x - px0, cy = rb.y - py0;      // (the nearest pixel may be far outside this sub-tile)
            const int cA = max(0, cx - W), cB = min(kT - 1, cx + W), rA = max(0, cy - W), rB = min(kT - 1, cy + W);
            if (cA <= cB && rA <= rB) {
                const unsigned cols = ((2u << (cB >> 2)) - 1u) & ~((1u << (cA >> 2)) - 1u);           // 4-bit pattern
                const int bA = rA >> 1, bB = rB >> 1;
                const unsigned hi = bB == 7 ? 0xffffffffu : ((1u << (4 * (bB + 1))) - 1u);
                touched = cols * (hi & ~((1u << (4 * bA)) - 1u) & 0x11111111u);
            }
        }
        unsigned mine = warp_transpose32(touched, lane);       // bit q: staged particle q overlaps my pixel block
        __syncwarp();
        // all factors of the phase first (independent iterations: the ex2 chain is pipelined), one __syncwarp, then
        // the multiply-adds with nothing but loads in between: neither loop waits on a per-particle exchange
#pragma unroll 4
        for (int q = 0; q < np; ++q) {
            // gauss_factor(cpx, k, u) with k = pixel - nearest pixel, zero outside the support / the image
            const float4 r = s_rec[warp][axis][q];
            const float fk = r.z + flf;
            float f = ex2_ftz((r.y * fk) * (fk + r.x));
            f = (fabsf(fk) <= r.w && fvalid) ? f : 0.f;
            s_f[warp][q][lane] = f;
        }
        __syncwarp();
        while (__any_sync(0xffffffffu, mine != 0)) {
            if (COUNT) { ++n_rounds; n_visits += mine != 0; }
            if (mine) {
                const int q = __ffs(mine) - 1;
                mine &= mine - 1;
                const float4 fx = *reinterpret_cast<const float4*>(&s_f[warp][q][4 * bc]);
                const float2 fy = *reinterpret_cast<const float2*>(&s_f[warp][q][kT + 2 * br]);
                float2 g[4];
                g[0] = __fmul2_rn(make_float2(fy.x, fy.x), make_float2(fx.x, fx.y));
                g[1] = __fmul2_rn(make_float2(fy.x, fy.x), make_float2(fx.z, fx.w));
                g[2] = __fmul2_rn(make_float2(fy.y, fy.y), make_float2(fx.x, fx.y));
                g[3] = __fmul2_rn(make_float2(fy.y, fy.y), make_float2(fx.z, fx.w));
                float w[C];
                if constexpr (C % 4 == 0) {
#pragma unroll
                    for (int c = 0; c < C; c += 4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(&s_w[warp][q][c]);
                        w[c] = w4.x; w[c + 1] = w4.y; w[c + 2] = w4.z; w[c + 3] = w4.w;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) w[c] = s_w[warp][q][c];
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float2 w2 = make_float2(w[c], w[c]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[c][j] = __ffma2_rn(w2, g[j], acc[c][j]);
                }
            }
        }
    }

    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_visits += __shfl_xor_sync(0xffffffffu, n_visits, o);
        if (lane == 0) { atomicAdd(p.visits, (unsigned long long)n_visits); atomicAdd(p.visits + 1, 32ull * n_rounds); }
    }
    const int col = px0 + 4 * bc;
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const float4 v = make_float4(acc[c][2 * rr].x, acc[c][2 * rr].y, acc[c][2 * rr + 1].x, acc[c][2 * rr + 1].y);
            const int r = py0 + 2 * br + rr;
            if (single) {
                if (r < p.ndim) {
                    const size_t off = ((size_t)image * p.n_chan + p.chan0 + c) * p.plane + (size_t)r * p.ld + col;
                    float* out = p.data + off;
                    if (col + 4 <= p.ndim && ((p.ld | p.plane) & 3) == 0) {
                        *reinterpret_cast<float4*>(out) = v;
                        if (push)
                            for (int q = 0; q < p.push_n; ++q) *reinterpret_cast<float4*>(p.push_result[q] + off) = v;
                    } else {
                        if (push)
                            for (int q = 0; q < p.push_n; ++q)
                                for (int j = 0; j < 4; ++j)
                                    if (col + j < p.ndim) p.push_result[q][off + j] = j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w;
                        if (col < p.ndim) out[0] = v.x;
                        if (col + 1 < p.ndim) out[1] = v.y;
                        if (col + 2 < p.ndim) out[2] = v.z;
                        if (col + 3 < p.ndim) out[3] = v.w;
                    }
                }
            } else {
                float* out = p.partial + ((size_t)item * p.n_chan + p.chan0 + c) * (kT * kT) + (2 * br + rr) * kT + 4 * bc;
                *reinterpret_cast<float4*>(out) = v;
            }
        }
    }
}

// Sub-tiles with one work item were written by the deposit kernel, empty ones keep the zeros the image was cleared
// with; a sub-tile split over several items is the sum of their partial sub-tiles in a FIXED order (four interleaved
// running sums, combined pairwise: deterministic, and the loads of four items are in flight).
// One CTA looks at 32 consecutive work items (one per lane of its first warp) and finishes the sub-tiles whose FIRST
// item is among them (blockIdx.y = channel), so the launch is max_items / 32 x channels CTAs whatever the image size.
constexpr int kReduceItems = 32;
template <bool PUSH>
__global__ void __launch_bounds__(kT * kT) tile_reduce_kernel(const DepositParams p, int n_chan) {
    __shared__ unsigned s_todo;
    __shared__ int s_tile[kReduceItems];
    if (threadIdx.x < 32) {
        const int64_t item = (int64_t)blockIdx.x * kReduceItems + threadIdx.x;
        bool first = false;
        int tile = 0;
        if (item < p.item_start[p.ntiles]) {
            tile = p.item_tile[item];
            first = p.item_start[tile] == item && p.item_start[tile + 1] - item > 1;
        }
        s_tile[threadIdx.x] = tile;
        const unsigned m = __ballot_sync(0xffffffffu, first);
        if (threadIdx.x == 0) s_todo = m;
    }
    __syncthreads();
    unsigned todo = s_todo;
    const int e = threadIdx.x;
    while (todo) {
        const int tile = s_tile[__ffs(todo) - 1];
        todo &= todo - 1;
        const int64_t i0 = p.item_start[tile], i1 = p.item_start[tile + 1];
        const int image = tile / p.ntiles_img, t_in = tile - image * p.ntiles_img;
        const int row = (t_in / p.ntx) * kT + e / kT, col = (t_in % p.ntx) * kT + e % kT;
        const size_t stride = (size_t)n_chan * (kT * kT);
        {
            const int chan = blockIdx.y;
            const float* src = p.partial + ((size_t)i0 * n_chan + chan) * (kT * kT) + e;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int64_t it = i0;
            for (; it + 4 <= i1; it += 4) {
                s0 += src[0]; s1 += src[stride]; s2 += src[2 * stride]; s3 += src[3 * stride];
                src += 4 * stride;
            }
            for (; it < i1; ++it) { s0 += src[0]; src += stride; }
            if (row < p.ndim && col < p.ndim) {
                const size_t off = ((size_t)image * n_chan + chan) * p.plane + (size_t)row * p.ld + col;
                const float v = (s0 + s1) + (s2 + s3);
                p.data[off] = v;
                if (PUSH && tile_exclusive(p, tile))
                    for (int q = 0; q < p.push_n; ++q) p.push_result[q][off] = v;
            }
        }
    }
}

// NGP products: one CTA per sub-tile; shared-memory accumulation (float64 sums, integer counts)
__global__ void __launch_bounds__(64) ngp_tile_kernel(const DepositParams p, int n_chan, float* __restrict__ simple,
                                                      float* __restrict__ hist) {
    __shared__ double s_sum[kT * kT];
    __shared__ unsigned int s_cnt[kT * kT];
    const int tile = blockIdx.x;
    const int image = tile / p.ntiles_img, t_in = tile - image * p.ntiles_img;
    const int py0 = (t_in / p.ntx) * kT, px0 = (t_in % p.ntx) * kT;
    const int64_t t0 = p.tile_start[tile], t1 = p.tile_start[tile + 1];
    for (int chan = 0; chan < (simple ? n_chan : 1); ++chan) {
        for (int e = threadIdx.x; e < kT * kT; e += 64) { s_sum[e] = 0.0; s_cnt[e] = 0; }
        __syncthreads();
        const double* Lc = p.L ? p.L + (size_t)chan * p.n : nullptr;
        for (int64_t q = t0 + threadIdx.x; q < t1; q += 64) {
            const int pid = p.pids[q];
            const int4 r = __ldg(reinterpret_cast<const int4*>(p.pv.rec + pid) + 1);
            const int lx = r.x - px0, ly = r.y - py0;
            if (lx >= 0 && lx < kT && ly >= 0 && ly < kT) {   // the particle's own sub-tile only
                if (simple) atomicAdd(&s_sum[ly * kT + lx], Lc[pid]);
                atomicAdd(&s_cnt[ly * kT + lx], 1u);
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < kT * kT; e += 64) {
            const int row = py0 + e / kT, col = px0 + e % kT;
            if (row < p.ndim && col < p.ndim) {
                if (simple) simple[(((size_t)image * n_chan + chan) * p.ndim + row) * p.ndim + col] = (float)s_sum[e];
                if (hist && chan == 0) hist[((size_t)image * p.ndim + row) * p.ndim + col] = (float)s_cnt[e];
            }
        }
        __syncthreads();
    }
}

__global__ void rebin_kernel(const float* __restrict__ in, int64_t nin, int64_t plane, float* __restrict__ out,
                             int n_out, int s) {
    const int img = blockIdx.z;
    const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ox >= n_out || oy >= n_out) return;
    const float* src = in + (size_t)img * plane;          // nin = row stride of the input (>= n_out * s)
    // images.py:22 sums the last axis first (columns within the block), then the rows
    float tot = 0.f;
    for (int r = 0; r < s; ++r) {
        float rs = 0.f;
        for (int c = 0; c < s; ++c) rs += src[((size_t)oy * s + r) * nin + (size_t)ox * s + c];
        tot += rs;
    }
    out[((size_t)img * n_out + oy) * n_out + ox] = tot;
}

// Sersic2D profile on the pixel grid (astropy.modeling.models.Sersic2D.evaluate, amplitude 1):
//   exp(-b_n ((z)^(1/n) - 1)),  z = sqrt((x_maj/a)^2 + (x_min/b)^2),  a = r_eff, b = (1 - ellip) r_eff
// float64 throughout; per-block partial sums (fixed order) for the normalisation of images.py:233
__global__ void __launch_bounds__(256) sersic_kernel(const double* __restrict__ G, int ndim, double r_eff, double n,
                                                     double x0, double y0, double ellip, double theta, double bn,
                                                     double* __restrict__ out, double* __restrict__ block_sums) {
    __shared__ double s[256];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    if (i < (int64_t)ndim * ndim) {
        const int row = (int)(i / ndim), col = (int)(i % ndim);      // row = y, column = x (np.meshgrid(g, g))
        const double x = G[col], y = G[row];
        const double a = r_eff, b = (1.0 - ellip) * r_eff;
        const double ct = cos(theta), st = sin(theta);
        const double x_maj = (x - x0) * ct + (y - y0) * st;
        const double x_min = -(x - x0) * st + (y - y0) * ct;
        const double qa = x_maj / a, qb = x_min / b;
        const double z = sqrt(qa * qa + qb * qb);
        v = exp(-bn * (pow(z, 1.0 / n) - 1.0));
        out[i] = v;
    }
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s[0];
}

// out[c] = (float)(L[c] * profile / sum(block_sums))    (images.py:233,238)
__global__ void __launch_bounds__(256) sersic_scale_kernel(const double* __restrict__ prof, const double* __restrict__ block_sums,
                                                           int n_blocks, const double* __restrict__ L, int n_chan, int64_t npix,
                                                           float* __restrict__ out) {
    __shared__ double s_tot;
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int b = 0; b < n_blocks; ++b) t += block_sums[b];     // fixed order
        s_tot = t;
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const double v = prof[i] / s_tot;
    for (int c = 0; c < n_chan; ++c) out[(size_t)c * npix + i] = (float)(L[c] * v);
}

__global__ void complex_mul_kernel(const float2* __restrict__ a, const float2* __restrict__ b,
                                   float2* __restrict__ out, int64_t nc, int b_per_image) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    const float2 x = a[(size_t)blockIdx.y * nc + i], y = b[(b_per_image ? (size_t)blockIdx.y * nc : 0) + i];
    out[(size_t)blockIdx.y * nc + i] = make_float2(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x);
}

// ---- several GPUs, one image: mask exchange, barrier and overlap fix over peer memory (so_img_deposit_push) ----
struct PeerPtrs {
    void* p[SO_MAX_PEERS];
};

// this rank's row of the sub-tile masks (a sub-tile is touched when it has a work item), stored into every rank's copy
__global__ void __launch_bounds__(256) push_mask_kernel(const int64_t* __restrict__ item_start, int ntiles, PeerPtrs masks, int n_ranks,
                                                        int rank) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const unsigned char m = item_start[t + 1] > item_start[t] ? 1 : 0;
    for (int q = 0; q < n_ranks; ++q) reinterpret_cast<unsigned char*>(masks.p[q])[(size_t)rank * ntiles + t] = m;
}

// Barrier between the ranks' streams over NVLink: lane q publishes `epoch` in rank q's flag array and waits until rank q's
// epoch has arrived in this rank's.  Kernels of a stream run in order, so everything the peers enqueued before their
// barrier has completed (and is visible: release / acquire at system scope) when this kernel ends.  A rank that never
// arrives is a trap after ~4 s, not a hang.
__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerPtrs flags, int n_ranks, int rank, unsigned long long epoch) {
    const int q = threadIdx.x;
    if (q >= n_ranks) return;
    __threadfence_system();
    unsigned long long* theirs = reinterpret_cast<unsigned long long*>(flags.p[q]) + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(flags.p[rank]) + q;
    const long long t0 = clock64();
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= epoch) break;
        if (clock64() - t0 > 8000000000LL) __trap();
    }
    __threadfence_system();
    // how long this rank waited for the slowest one [ns], per barrier of the call (slots SO_MAX_PEERS, SO_MAX_PEERS + 1 of
    // its own flag array): what a caller needs to balance the shares by measured busy time
    __syncwarp((1u << n_ranks) - 1u);                 // every lane has seen its rank arrive
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (q == 0) reinterpret_cast<unsigned long long*>(flags.p[rank])[SO_MAX_PEERS + (int)(epoch & 1ull)] = g1 - g0;
}

// sub-tiles several ranks touched are the sum of their partial sub-tiles in rank order (deterministic, the same on every
// rank); sub-tiles of one rank arrived with its deposit, untouched ones keep the zeros the result was cleared with.
// One WARP per sub-tile: almost all of them leave after the mask test.
__global__ void __launch_bounds__(256) push_fix_kernel(PeerPtrs partial, const unsigned char* __restrict__ masks, int n_ranks,
                                                       int ntiles, int ntx, int ndim, int n_chan, float* __restrict__ result) {
    const int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (tile >= ntiles) return;
    const unsigned touch = __ballot_sync(0xffffffffu, lane < n_ranks && masks[(size_t)lane * ntiles + tile]);
    if (__popc(touch) < 2) return;
    const int py0 = (tile / ntx) * kT, px0 = (tile % ntx) * kT;
    for (int c = 0; c < n_chan; ++c)
        for (int e = lane; e < kT * kT; e += 32) {
            const int row = py0 + e / kT, col = px0 + e % kT;
            if (row >= ndim || col >= ndim) continue;
            const size_t off = ((size_t)c * ndim + row) * ndim + col;
            float s = 0.f;
            for (unsigned t = touch; t; t &= t - 1) s += reinterpret_cast<const float*>(partial.p[__ffs(t) - 1])[off];
            result[off] = s;
        }
}

struct DepositLayout {
    int32_t *keys0, *vals0, *keys1, *vals1;
    int64_t *tile_start, *item_start;
    int32_t* item_tile;
    float* partial;
    float* wts;
    void* cub_tmp;
    int64_t cub_tmp_bytes;
    int64_t chunk, max_items;
    int ntx, ntiles;
};

static int64_t sort_tmp_bytes(int64_t n_pairs) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs((void*)nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)n_pairs, 0, 32);
    return (int64_t)b;
}

static DepositLayout carve_deposit(void* ws, int64_t bytes, int64_t n, int64_t n_pairs, int ndim, int n_chan, int n_img,
                                   bool* ok) {
    DepositLayout d;
    d.ntx = (ndim + kT - 1) / kT;
    d.ntiles = d.ntx * d.ntx * n_img;       // over the whole batch of images
    // one warp per work item: aim for ~32 items (warps) per SM, never below two staging phases
    const int64_t target = 32 * 148;        // measured flat between 16 and 64 items per SM
    int64_t chunk = (n_pairs + target - 1) / target;
    chunk = (chunk + kP - 1) / kP * kP;
    if (chunk < 2 * kP) chunk = 2 * kP;
    d.chunk = chunk;
    d.max_items = n_pairs / chunk + d.ntiles + 1;
    SoArena a(ws, bytes);
    d.keys0 = a.take<int32_t>(n_pairs);
    d.vals0 = a.take<int32_t>(n_pairs);
    d.keys1 = a.take<int32_t>(n_pairs);
    d.vals1 = a.take<int32_t>(n_pairs);
    d.tile_start = a.take<int64_t>(d.ntiles + 1);
    d.item_start = a.take<int64_t>(d.ntiles + 1);
    d.item_tile = a.take<int32_t>(d.max_items);
    d.partial = a.take<float>(d.max_items * n_chan * kT * kT);
    d.wts = a.take<float>(n * n_chan);
    d.cub_tmp_bytes = sort_tmp_bytes(n_pairs > 0 ? n_pairs : 1);
    {
        size_t b = 0;
        cub::DeviceScan::ExclusiveSum((void*)nullptr, b, (int64_t*)nullptr, (int64_t*)nullptr, d.ntiles + 1);
        if ((int64_t)b > d.cub_tmp_bytes) d.cub_tmp_bytes = (int64_t)b;
    }
    d.cub_tmp = a.take<char>(d.cub_tmp_bytes);
    if (ok) *ok = a.ok();
    if (!ws) d.cub_tmp_bytes = a.used;   // total, for the size query
    return d;
}

}  // namespace

// so_img_deposit_count: device counters the next deposits add their (lane, particle) visits to (measurement aid)
static unsigned long long* g_deposit_visits = nullptr;
extern "C" int so_img_deposit_count(unsigned long long* counters) {
    g_deposit_visits = counters;
    return SO_OK;
}

extern "C" int so_img_ngp_index(const double* X, const double* Y, int64_t n, const double* G, int32_t ndim,
                                double half_width, int32_t* ix, int32_t* iy, void* stream) {
    if (n < 0 || ndim < 1 || !G) return SO_ERR_BAD_ARG;
    if (n == 0) return SO_OK;
    { SO_PROF("ngp_index_kernel", (cudaStream_t)stream); ngp_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, Y, n, G, ndim, half_width, ix, iy); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

extern "C" int64_t so_img_plan_workspace_bytes(int64_t n) {
    SoArena a(nullptr, 0);
    bool ok;
    (void)ok;
    // replicate carve_plan's accounting with a null base
    a.take<PlanRec>(n);
    a.take<int64_t>(n + 1);
    a.take<unsigned long long>(4);
    a.take<char>(scan_tmp_bytes(n));
    return a.used + 256;
}

extern "C" int so_img_plan(const int32_t* ix, const int32_t* iy, const double* X, const double* Y,
                           const double* dk, int64_t n, const double* G, int32_t ndim,
                           double support_sigmas, void* plan_workspace, int64_t plan_workspace_bytes,
                           int64_t* n_pairs, double* stats, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 0 || ndim < 1 || !G || !n_pairs || (n > 0 && (!ix || !iy || !X || !Y))) return SO_ERR_BAD_ARG;
    if (n >= (int64_t)1 << 31) return SO_ERR_TOO_LARGE;
    bool ok = false;
    PlanView pv = carve_plan(plan_workspace, plan_workspace_bytes, n, &ok);
    if (!ok || !plan_workspace) return SO_ERR_WORKSPACE;
    SO_CUDA(cudaMemsetAsync(pv.stats, 0, 4 * sizeof(unsigned long long), st));
    *n_pairs = 0;
    if (n == 0) {
        if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0.0;
        return SO_OK;
    }
    { SO_PROF("plan_kernel", st); plan_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ix, iy, X, Y, dk, n, G, ndim, support_sigmas, pv); }
    SO_LAUNCH_CHECK();
    if (dk) {
        { SO_PROF("plan_wide_kernel", st); plan_wide_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, ndim, pv); }
        SO_LAUNCH_CHECK();
    }
    size_t tmp = (size_t)pv.cub_tmp_bytes;
    SO_CUDA(cub::DeviceScan::ExclusiveSum(pv.cub_tmp, tmp, pv.offs, pv.offs, (int)(n + 1), st));
    unsigned long long h[4];
    int64_t total = 0;
    SO_CUDA(cudaMemcpyAsync(&total, pv.offs + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SO_CUDA(cudaMemcpyAsync(h, pv.stats, sizeof(h), cudaMemcpyDeviceToHost, st));
    SO_CUDA(cudaStreamSynchronize(st));
    *n_pairs = total;
    if (stats) for (int i = 0; i < 4; ++i) stats[i] = (double)h[i];
    return SO_OK;
}

extern "C" int64_t so_img_deposit_batch_workspace_bytes(int64_t n, int64_t n_pairs, int32_t ndim, int32_t n_chan,
                                                        int32_t n_img) {
    DepositLayout d = carve_deposit(nullptr, 0, n, n_pairs, ndim, n_chan, n_img < 1 ? 1 : n_img, nullptr);
    return d.cub_tmp_bytes + 256;
}

extern "C" int64_t so_img_deposit_workspace_bytes(int64_t n, int64_t n_pairs, int32_t ndim, int32_t n_chan) {
    return so_img_deposit_batch_workspace_bytes(n, n_pairs, ndim, n_chan, 1);
}

extern "C" int so_img_deposit_batch(const void* plan_workspace, const double* L, const int32_t* img, int64_t n,
                                    int32_t n_chan, int32_t n_img, int32_t ndim, int64_t n_pairs,
                                    float* data, float* simple, float* hist,
                                    int32_t* tile_ids, int32_t* pair_pids,
                                    void* workspace, int64_t workspace_bytes, void* stream) {
    return so_img_deposit_batch_ld(plan_workspace, L, img, n, n_chan, n_img, ndim, n_pairs, data, ndim,
                                   (int64_t)ndim * ndim, simple, hist, tile_ids, pair_pids, workspace, workspace_bytes, stream);
}

extern "C" int so_img_deposit_batch_ld(const void* plan_workspace, const double* L, const int32_t* img, int64_t n,
                                       int32_t n_chan, int32_t n_img, int32_t ndim, int64_t n_pairs,
                                       float* data, int64_t data_ld, int64_t data_plane, float* simple, float* hist,
                                       int32_t* tile_ids, int32_t* pair_pids,
                                       void* workspace, int64_t workspace_bytes, void* stream) {
    return so_img_deposit_push(plan_workspace, L, img, n, n_chan, n_img, ndim, n_pairs, data, data_ld, data_plane, simple, hist,
                               tile_ids, pair_pids, workspace, workspace_bytes, nullptr, stream);
}

extern "C" int so_img_deposit_push(const void* plan_workspace, const double* L, const int32_t* img, int64_t n,
                                   int32_t n_chan, int32_t n_img, int32_t ndim, int64_t n_pairs,
                                   float* data, int64_t data_ld, int64_t data_plane, float* simple, float* hist,
                                   int32_t* tile_ids, int32_t* pair_pids,
                                   void* workspace, int64_t workspace_bytes, const so_push_t* push, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (push) {
        if (push->n_ranks < 2 || push->n_ranks > SO_MAX_PEERS || push->rank < 0 || push->rank >= push->n_ranks || n_img != 1 ||
            !data || simple || hist)
            return SO_ERR_BAD_ARG;
        for (int r = 0; r < push->n_ranks; ++r)
            if (!push->result[r] || !push->partial[r] || !push->masks[r] || !push->flags[r]) return SO_ERR_BAD_ARG;
        if (push->partial[push->rank] != (void*)data) return SO_ERR_BAD_ARG;
    }
    if (n < 0 || ndim < 1 || n_chan < 1 || n_pairs < 0 || !plan_workspace || n_img < 1) return SO_ERR_BAD_ARG;
    if (data && (data_ld < ndim || data_plane < data_ld * ndim)) return SO_ERR_BAD_ARG;
    const bool dense = data_ld == ndim && data_plane == (int64_t)ndim * ndim;
    if (n_img > 1 && !img && n > 0) return SO_ERR_BAD_ARG;
    if (n_pairs >= (int64_t)1 << 31) return SO_ERR_TOO_LARGE;
    if (n > 0 && (data || simple) && !L) return SO_ERR_BAD_ARG;
    const int64_t ntx0 = (ndim + kT - 1) / kT;
    if (ntx0 * ntx0 * (int64_t)n_img >= ((int64_t)1 << 31) - 1) return SO_ERR_TOO_LARGE;
    bool ok = false;
    PlanView pv = carve_plan(const_cast<void*>(plan_workspace), (int64_t)1 << 60, n, &ok);
    DepositLayout d = carve_deposit(workspace, workspace_bytes, n, n_pairs, ndim, n_chan, n_img, &ok);
    if (!ok || !workspace) return SO_ERR_WORKSPACE;
    const size_t img_bytes = (size_t)ndim * ndim * sizeof(float) * n_img;
    if (push && !(data_ld == ndim && data_plane == (int64_t)ndim * ndim)) return SO_ERR_BAD_ARG;
    if (n_pairs == 0 && !push) {
        if (data) {
            if (dense) SO_CUDA(cudaMemsetAsync(data, 0, img_bytes * n_chan, st));
            else
                for (int64_t i = 0; i < (int64_t)n_img * n_chan; ++i)
                    SO_CUDA(cudaMemset2DAsync(data + i * data_plane, data_ld * sizeof(float), 0, ndim * sizeof(float), ndim, st));
        }
        if (simple) SO_CUDA(cudaMemsetAsync(simple, 0, img_bytes * n_chan, st));
        if (hist) SO_CUDA(cudaMemsetAsync(hist, 0, img_bytes, st));
        return SO_OK;
    }
    const int ntiles_img = d.ntx * d.ntx;
    // sub-tiles nobody touches stay zero: the result is cleared before this rank passes the first barrier, i.e. before
    // any peer stores into it
    if (push) SO_CUDA(cudaMemsetAsync(push->result[push->rank], 0, img_bytes * n_chan, st));
    if (n > 0 && n_pairs > 0)
    { SO_PROF("expand_kernel", st); expand_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pv, n, ndim, d.ntx, img, ntiles_img, d.keys0, d.vals0); }
    SO_LAUNCH_CHECK();
    int bits = 1;
    while (((int64_t)1 << bits) < d.ntiles) ++bits;
    size_t tmp = (size_t)d.cub_tmp_bytes;
    if (n_pairs > 0)
    { SO_PROF("cub sort (tile, particle) pairs", st); SO_CUDA(cub::DeviceRadixSort::SortPairs(d.cub_tmp, tmp, d.keys0, d.keys1, d.vals0, d.vals1, (int)n_pairs, 0, bits, st)); }
    { SO_PROF("tile_start_kernel", st); tile_start_kernel<<<(d.ntiles + 1 + 255) / 256, 256, 0, st>>>(d.keys1, n_pairs, d.ntiles, d.tile_start); }
    SO_LAUNCH_CHECK();
    if (tile_ids) SO_CUDA(cudaMemcpyAsync(tile_ids, d.keys1, sizeof(int32_t) * n_pairs, cudaMemcpyDeviceToDevice, st));
    if (pair_pids) SO_CUDA(cudaMemcpyAsync(pair_pids, d.vals1, sizeof(int32_t) * n_pairs, cudaMemcpyDeviceToDevice, st));

    DepositParams p{};
    p.pv = pv; p.L = L; p.n = n; p.ndim = ndim; p.ntx = d.ntx; p.ntiles = d.ntiles; p.ntiles_img = ntiles_img;
    p.pids = d.vals1; p.tile_start = d.tile_start; p.item_start = d.item_start; p.item_tile = d.item_tile;
    p.ld = data_ld; p.plane = data_plane;
    p.chunk = d.chunk; p.data = data; p.partial = d.partial; p.n_chan = n_chan; p.chan0 = 0; p.wts = d.wts;
    p.push_n = 0;
    p.visits = g_deposit_visits;
    PeerPtrs pp_partial{}, pp_masks{}, pp_flags{};
    if (push) {
        p.push_n = push->n_ranks;
        p.push_masks = (const unsigned char*)push->masks[push->rank];
        for (int r = 0; r < push->n_ranks; ++r) {
            p.push_result[r] = (float*)push->result[r];
            pp_partial.p[r] = push->partial[r]; pp_masks.p[r] = push->masks[r]; pp_flags.p[r] = push->flags[r];
        }
    }
    if (data) {
        // empty sub-tiles are never written: a dense image is cleared here (a push deposit zeroes them in its result
        // instead); a window of a larger (padded) buffer is cleared by the caller together with its padding
        if (dense && !push) SO_CUDA(cudaMemsetAsync(data, 0, img_bytes * n_chan, st));
        if (n > 0)
        { SO_PROF("weights_kernel", st); weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L, pv.rec, n, n_chan, d.wts); }
        SO_LAUNCH_CHECK();
        { SO_PROF("tile_item_count_kernel", st); tile_item_count_kernel<<<(d.ntiles + 1 + 255) / 256, 256, 0, st>>>(d.tile_start, d.ntiles, d.chunk, d.item_start); }
        SO_LAUNCH_CHECK();
        size_t stmp = (size_t)d.cub_tmp_bytes;
        SO_CUDA(cub::DeviceScan::ExclusiveSum(d.cub_tmp, stmp, d.item_start, d.item_start, d.ntiles + 1, st));
        { SO_PROF("item_tile_kernel", st); item_tile_kernel<<<(unsigned)((d.max_items + 255) / 256), 256, 0, st>>>(d.item_start, d.ntiles, d.max_items,
                                                                              d.item_tile); }
        SO_LAUNCH_CHECK();
        if (push) {
            // every rank learns which sub-tiles the others touch before anybody deposits
            { SO_PROF("push_mask_kernel", st);
              push_mask_kernel<<<(d.ntiles + 255) / 256, 256, 0, st>>>(d.item_start, d.ntiles, pp_masks, push->n_ranks, push->rank); }
            { SO_PROF("peer_barrier_kernel", st); peer_barrier_kernel<<<1, 32, 0, st>>>(pp_flags, push->n_ranks, push->rank, 2 * push->epoch); }
            SO_LAUNCH_CHECK();
        }
        const unsigned grid = (unsigned)((d.max_items + kDepWarps - 1) / kDepWarps);
        for (int c0 = 0; c0 < n_chan && n_pairs > 0;) {       // channel groups of 8 / 4 / 2 / 1 share one pass over the lists
            SO_PROF("deposit_kernel", st);
            p.chan0 = c0;
            const int left = n_chan - c0;
#define SO_DEPOSIT(C)                                                                                 \
            do {                                                                                      \
                if (p.push_n > 0) deposit_kernel<C, true><<<grid, kDepWarps * 32, 0, st>>>(p);        \
                else if (p.visits) deposit_kernel<C, false, true><<<grid, kDepWarps * 32, 0, st>>>(p); \
                else deposit_kernel<C, false><<<grid, kDepWarps * 32, 0, st>>>(p);                    \
                c0 += C;                                                                              \
            } while (0)
            if (left >= 8) SO_DEPOSIT(8);
            else if (left >= 4) SO_DEPOSIT(4);
            else if (left >= 2) SO_DEPOSIT(2);
            else SO_DEPOSIT(1);
#undef SO_DEPOSIT
            SO_LAUNCH_CHECK();
        }
        if (n_pairs > 0)
        { SO_PROF("tile_reduce_kernel", st);
          const dim3 rg((unsigned)((d.max_items + kReduceItems - 1) / kReduceItems), (unsigned)n_chan);
          if (p.push_n > 0) tile_reduce_kernel<true><<<rg, kT * kT, 0, st>>>(p, n_chan);
          else tile_reduce_kernel<false><<<rg, kT * kT, 0, st>>>(p, n_chan); }
        SO_LAUNCH_CHECK();
        if (push) {
            // all ranks have deposited (and pushed their own sub-tiles): finish the shared ones
            { SO_PROF("peer_barrier_kernel", st); peer_barrier_kernel<<<1, 32, 0, st>>>(pp_flags, push->n_ranks, push->rank, 2 * push->epoch + 1); }
            { SO_PROF("push_fix_kernel", st);
              push_fix_kernel<<<(unsigned)(((int64_t)d.ntiles * 32 + 255) / 256), 256, 0, st>>>(pp_partial, p.push_masks, push->n_ranks, d.ntiles, d.ntx, ndim, n_chan,
                                                                       (float*)push->result[push->rank]); }
            SO_LAUNCH_CHECK();
        }
    }
    if (simple || hist) {
        { SO_PROF("ngp_tile_kernel", st); ngp_tile_kernel<<<(unsigned)d.ntiles, 64, 0, st>>>(p, n_chan, simple, hist); }
        SO_LAUNCH_CHECK();
    }
    return SO_OK;
}

extern "C" int so_img_deposit(const void* plan_workspace, const double* L, int64_t n, int32_t n_chan,
                              int32_t ndim, int64_t n_pairs,
                              float* data, float* simple, float* hist,
                              int32_t* tile_ids, int32_t* pair_pids,
                              void* workspace, int64_t workspace_bytes, void* stream) {
    return so_img_deposit_batch(plan_workspace, L, nullptr, n, n_chan, 1, ndim, n_pairs, data, simple, hist, tile_ids,
                                pair_pids, workspace, workspace_bytes, stream);
}

extern "C" int so_img_rebin(const float* in, float* out, int32_t n_img, int32_t n_out, int32_t s, void* stream) {
    return so_img_rebin_ld(in, (int64_t)n_out * s, (int64_t)n_out * s * n_out * s, out, n_img, n_out, s, stream);
}

extern "C" int so_img_rebin_ld(const float* in, int64_t in_ld, int64_t in_plane, float* out, int32_t n_img,
                               int32_t n_out, int32_t s, void* stream) {
    if (n_img < 1 || n_out < 1 || s < 1 || in_ld < (int64_t)n_out * s) return SO_ERR_BAD_ARG;
    dim3 block(32, 8), grid((n_out + 31) / 32, (n_out + 7) / 8, n_img);
    { SO_PROF("rebin_kernel", (cudaStream_t)stream); rebin_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, in_ld, in_plane, out, n_out, s); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

extern "C" int so_complex_mul_bcast(const float* a, const float* b, float* out, int64_t n_complex, int32_t n_img,
                                    int32_t b_per_image, void* stream) {
    if (n_complex < 1 || n_img < 1) return SO_ERR_BAD_ARG;
    dim3 grid((unsigned)((n_complex + 255) / 256), (unsigned)n_img);
    { SO_PROF("complex_mul_kernel", (cudaStream_t)stream); complex_mul_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float2*)a, (const float2*)b, (float2*)out, n_complex,
                                                               b_per_image); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

extern "C" int64_t so_img_sersic_workspace_bytes(int32_t ndim) {
    const int64_t npix = (int64_t)ndim * ndim;
    return so_align_up(npix * 8, 256) + so_align_up(((npix + 255) / 256) * 8, 256) + 256;
}

extern "C" int so_img_sersic(const double* G, int32_t ndim, double r_eff, double n, double x0, double y0,
                             double ellip, double theta, double bn, const double* L, int32_t n_chan,
                             float* out, void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!G || ndim < 1 || n_chan < 1 || !L || !out || !(r_eff > 0.0) || !(n > 0.0)) return SO_ERR_BAD_ARG;
    const int64_t npix = (int64_t)ndim * ndim;
    const int64_t nblk = (npix + 255) / 256;
    SoArena a(workspace, workspace_bytes);
    double* prof = a.take<double>(npix);
    double* sums = a.take<double>(nblk);
    if (!workspace || !a.ok()) return SO_ERR_WORKSPACE;
    { SO_PROF("sersic_kernel", st); sersic_kernel<<<(unsigned)nblk, 256, 0, st>>>(G, ndim, r_eff, n, x0, y0, ellip, theta, bn, prof, sums); }
    SO_LAUNCH_CHECK();
    { SO_PROF("sersic_scale_kernel", st); sersic_scale_kernel<<<(unsigned)nblk, 256, 0, st>>>(prof, sums, (int)nblk, L, n_chan, npix, out); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

// Line-of-sight metal column density of star particles: the quantity the reference's loaders
// read as MetSurfaceDensities.npy (synthobs/core.py:17-21, :33-37) and its examples turn into the ISM optical depth
// (examples/SED_luminosities.py:69).  The reference does not compute it; the specification is oracle/los_oracle.py:
//
//   Sigma_i = scale * sum_{j : z_j > z_i}  m_j Z_j / h_j^2 * F(b_ij / h_j),   b_ij = |(x_i - x_j, y_i - y_j)|,
//
// F = the cubic-spline SPH kernel (support r < h) projected along the line of sight, tabulated at NT + 1 nodes and
// interpolated linearly.
//
// One thread = four stars, one CTA = 512 stars of one object (and, when the stars alone do not fill the GPU, one of up to
// 16 interleaved chunks of the object's gas tiles: the chunks' partial columns are added in order by a second kernel); the object's gas particles pass through shared memory in
// tiles of 128 (one coalesced load per thread and tile: position relative to the object's first gas particle and
// as FP32 head + tail, 1 / h^2, m Z / h^2 in FP32, z in float64, and the four coefficients of (b/h)^2 as a polynomial in
// the star's coordinates with its FP32 error bound taken off the constant), and every thread walks the tile with broadcast
// loads: three multiply-adds and a compare per pair; the few pairs inside a kernel's support (b < h, in front) are noted
// and then decided exactly: difference from heads and tails (coordinates far from the origin lose nothing), one rsqrt
// and two table loads.  The z test is made on the float64 inputs (a particle counts or does not: no FP32
// rounding decides that), the smooth part is FP32, summed per tile in FP32 and over the tiles in float64 in a fixed
// order: bitwise reproducible.  Compute-bound: N_star x N_gas pairs at ~9 instructions (bench.py `los` record).
#include <vector>

#include "common.cuh"

namespace {

constexpr int kLosThreads = 128;      // threads per CTA = gas particles per tile
constexpr int kLosStars = 4;          // stars per thread: one broadcast load of a gas particle serves four pairs
constexpr int kLosMaxTable = 2048;
constexpr int kLosList = 32;          // noted candidates per thread and tile before they are worked off
constexpr int kLosMaxChunks = 16;

// blockIdx.x = block of kLosThreads * kLosStars stars of one object, blockIdx.y = chunk of the object's gas tiles
// (tiles c, c + n_chunk, ...): with one chunk the result goes straight to `out`, otherwise to partial[chunk][star] and
// los_sum_kernel adds the chunks in order.
__global__ void __launch_bounds__(kLosThreads) los_kernel(
        const double* __restrict__ xs, const double* __restrict__ ys, const double* __restrict__ zs,
        const int64_t* __restrict__ star_off, const double* __restrict__ xg, const double* __restrict__ yg,
        const double* __restrict__ zg, const double* __restrict__ mg, const double* __restrict__ Zg,
        const double* __restrict__ hg, const int64_t* __restrict__ gas_off, const int64_t* __restrict__ blk_start,
        int64_t n_obj, const float* __restrict__ table, int nt, double scale, int n_chunk, int64_t n_star,
        double* __restrict__ out) {
    __shared__ float4 s_g[kLosThreads];          // screening form of (b / h)^2 = A + B sx + C sy + D (sx^2 + sy^2): A', B, C, D
    __shared__ float4 s_p[kLosThreads];          // x, y relative to the object's origin as FP32 head + tail (2^-48 together)
    __shared__ float s_w[kLosThreads];           // m Z / h^2
    __shared__ double s_z[kLosThreads];
    __shared__ float s_tab[kLosMaxTable + 1];
    __shared__ unsigned short s_list[kLosThreads][kLosList];
    __shared__ float s_red[kLosThreads / 32];
    for (int i = threadIdx.x; i <= nt; i += kLosThreads) s_tab[i] = table[i];
    // object of this CTA: last g with blk_start[g] <= blockIdx.x
    int64_t lo = 0, hi = n_obj;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (blk_start[mid] <= (int64_t)blockIdx.x) lo = mid; else hi = mid;
    }
    const int64_t g = lo;
    const int64_t s0 = star_off[g], s1 = star_off[g + 1], g0 = gas_off[g], g1 = gas_off[g + 1];
    const int64_t i0 = s0 + ((int64_t)blockIdx.x - blk_start[g]) * (kLosThreads * kLosStars) + threadIdx.x;
    double ox = 0.0, oy = 0.0;
    if (g1 > g0) { ox = xg[g0]; oy = yg[g0]; }
    float sx[kLosStars], sy[kLosStars], s2[kLosStars], sxl[kLosStars], syl[kLosStars], acc[kLosStars];
    double sz[kLosStars], total[kLosStars];
    float smax = 0.f;                             // largest |star position| of the CTA: enters the screening error bound
#pragma unroll
    for (int u = 0; u < kLosStars; ++u) {
        const int64_t i = i0 + (int64_t)u * kLosThreads;
        sx[u] = sy[u] = 0.f;
        s2[u] = 3.0e30f;                          // no star: D s2 keeps it outside every kernel
        sxl[u] = syl[u] = 0.f;
        sz[u] = 0.0;
        total[u] = 0.0;
        if (i < s1) {
            const double rx = xs[i] - ox, ry = ys[i] - oy;
            sx[u] = (float)rx; sy[u] = (float)ry;
            sxl[u] = (float)(rx - (double)sx[u]); syl[u] = (float)(ry - (double)sy[u]);
            s2[u] = fmaf(sx[u], sx[u], sy[u] * sy[u]);
            smax = fmaxf(smax, sqrtf(s2[u]) * 1.0001f);
            sz[u] = zs[i];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = smax;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kLosThreads / 32; ++w) smax = fmaxf(smax, s_red[w]);
    const float fnt = (float)nt;
    unsigned short* mine = s_list[threadIdx.x];
    for (int64_t t0 = g0 + (int64_t)blockIdx.y * kLosThreads; t0 < g1; t0 += (int64_t)n_chunk * kLosThreads) {
        const int cnt = (int)min((int64_t)kLosThreads, g1 - t0);
        __syncthreads();
        if (threadIdx.x < cnt) {
            const int64_t j = t0 + threadIdx.x;
            const double h = hg[j];
            const double ih2 = 1.0 / (h * h);
            const double rx = xg[j] - ox, ry = yg[j] - oy;
            const float fx = (float)rx, fy = (float)ry;
            // (b/h)^2 expanded in the star's coordinates: three multiply-adds per pair.  The FP32 evaluation is off by at
            // most ~8 roundings of terms no larger than (|g| + |s|)^2 / h^2: that bound is taken off A, so the screen
            // only ever errs towards MORE candidates (they are decided exactly below)
            const double gn = sqrt(rx * rx + ry * ry) + (double)smax;
            const double err = 1.0e-6 * gn * gn * ih2;
            s_g[threadIdx.x] = make_float4((float)((rx * rx + ry * ry) * ih2 - err), (float)(-2.0 * rx * ih2), (float)(-2.0 * ry * ih2),
                                           (float)ih2);
            s_p[threadIdx.x] = make_float4(fx, fy, (float)(rx - (double)fx), (float)(ry - (double)fy));
            s_w[threadIdx.x] = (float)(mg[j] * Zg[j] * ih2);
            s_z[threadIdx.x] = zg[j];
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kLosStars; ++u) acc[u] = 0.f;
        // a candidate pair (star u of this thread, gas j of the tile): in front of the star?  then the difference from
        // heads and tails (exact to ~1e-7 of h whatever the distance from the origin) and the interpolated kernel
        auto inside = [&](int j, int u) {
            float usx = sx[0], usy = sy[0], usxl = sxl[0], usyl = syl[0];
            double usz = sz[0];
#pragma unroll
            for (int v = 1; v < kLosStars; ++v)
                if (u == v) { usx = sx[v]; usy = sy[v]; usxl = sxl[v]; usyl = syl[v]; usz = sz[v]; }
            if (!(s_z[j] > usz)) return;
            const float4 p = s_p[j];
            const float dx = (p.x - usx) + (p.z - usxl), dy = (p.y - usy) + (p.w - usyl);
            const float q2 = fmaf(dx, dx, dy * dy) * s_g[j].w;          // (b / h)^2
            if (q2 < 1.f) {
                const float t = q2 * rsqrtf(fmaxf(q2, 1e-30f)) * fnt;    // u NT, u = sqrt(q2)
                const int k = min((int)t, nt - 1);
                const float f = t - (float)k;
                const float a = s_tab[k];
                const float c = s_w[j] * fmaf(f, s_tab[k + 1] - a, a);
#pragma unroll
                for (int v = 0; v < kLosStars; ++v)
                    if (u == v) acc[v] += c;
            }
        };
        auto entry = [&](int e) {                                     // e = gas | 4-bit set of this thread's stars << 7
            const int j = e & (kLosThreads - 1);
#pragma unroll
            for (int u = 0; u < kLosStars; ++u)
                if ((e >> (7 + u)) & 1) inside(j, u);
        };
        // Pass 1, all pairs of the tile: one broadcast load of a gas particle's four screening coefficients, three
        // multiply-adds and a compare per star of the thread; a gas particle with a candidate among the thread's stars is
        // only NOTED in the thread's list (one store), so the rare exact work does not make a whole warp leave the pair
        // stream for it.  Pass 2, the lists: all lanes step through theirs together.
        int n_mine = 0;
#pragma unroll 2
        for (int j = 0; j < cnt; ++j) {
            const float4 p = s_g[j];
            int m = 0;
#pragma unroll
            for (int u = 0; u < kLosStars; ++u) {
                const float q = fmaf(p.y, sx[u], fmaf(p.z, sy[u], fmaf(p.w, s2[u], p.x)));
                m |= (q < 1.f ? 1 : 0) << (7 + u);
            }
            if (m) mine[n_mine++] = (unsigned short)(m | j);
            // a full list anywhere in the warp: ALL lanes work theirs off now, together (a lane-private flush would
            // serialise the warp: with the cell grid a third of the tested pairs are candidates)
            if (__any_sync(0xffffffffu, n_mine == kLosList)) {
                for (int k = 0; k < kLosList; ++k)
                    if (k < n_mine) entry(mine[k]);
                n_mine = 0;
            }
        }
        for (int k = 0; k < kLosList; ++k) {
            if (!__any_sync(0xffffffffu, k < n_mine)) break;
            if (k < n_mine) entry(mine[k]);
        }
#pragma unroll
        for (int u = 0; u < kLosStars; ++u) total[u] += (double)acc[u];
    }
#pragma unroll
    for (int u = 0; u < kLosStars; ++u) {
        const int64_t i = i0 + (int64_t)u * kLosThreads;
        if (i < s1) {
            if (n_chunk == 1) out[i] = scale * total[u];
            else out[(int64_t)blockIdx.y * n_star + i] = total[u];      // out = the partial buffer
        }
    }
}

__global__ void los_sum_kernel(const double* __restrict__ partial, int n_chunk, int64_t n_star, double scale, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_star) return;
    double t = 0.0;
    for (int c = 0; c < n_chunk; ++c) t += partial[(int64_t)c * n_star + i];
    out[i] = scale * t;
}

// gas chunks per object: enough CTAs to fill the GPU when the stars alone do not (one big object)
static int los_chunks(int64_t star_blocks) {
    const int64_t want = 8LL * so_num_sms();
    int64_t c = (want + star_blocks - 1) / (star_blocks > 0 ? star_blocks : 1);
    if (c < 1) c = 1;
    if (c > kLosMaxChunks) c = kLosMaxChunks;
    return (int)c;
}

}  // namespace

// chunks never exceed what the fewest possible star blocks (no padding between objects) would get
static int64_t los_partial_count(int64_t n_star) {
    if (n_star <= 0) return 0;
    const int c = los_chunks((n_star + kLosThreads * kLosStars - 1) / (kLosThreads * kLosStars));
    return c > 1 ? (int64_t)c * n_star : 0;
}

extern "C" int64_t so_los_workspace_bytes(int64_t n_obj, int64_t n_star) {
    return 3 * so_align_up((n_obj + 1) * (int64_t)sizeof(int64_t), 256) +
           so_align_up(los_partial_count(n_star) * (int64_t)sizeof(double), 256) + 256;
}

extern "C" int so_los_column(const double* xs, const double* ys, const double* zs, const int64_t* star_off_host,
                             const double* xg, const double* yg, const double* zg, const double* mg, const double* Zg,
                             const double* hg, const int64_t* gas_off_host, int64_t n_obj, const float* table, int32_t nt,
                             double scale, double* out, void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_obj == 0) return SO_OK;
    if (n_obj < 0 || !star_off_host || !gas_off_host || !table || nt < 2 || nt > kLosMaxTable) return SO_ERR_BAD_ARG;
    for (int64_t g = 0; g < n_obj; ++g)
        if (star_off_host[g + 1] < star_off_host[g] || gas_off_host[g + 1] < gas_off_host[g]) return SO_ERR_BAD_ARG;
    const int64_t n_star = star_off_host[n_obj] - star_off_host[0];
    if (n_star == 0) return SO_OK;                       // nothing to answer
    if (!xs || !ys || !zs || !out) return SO_ERR_BAD_ARG;
    if (!workspace || workspace_bytes < so_los_workspace_bytes(n_obj, n_star)) return SO_ERR_WORKSPACE;
    if (gas_off_host[n_obj] > gas_off_host[0] && (!xg || !yg || !zg || !mg || !Zg || !hg)) return SO_ERR_BAD_ARG;
    SoArena a(workspace, workspace_bytes);
    int64_t* d_star = a.take<int64_t>(n_obj + 1);
    int64_t* d_gas = a.take<int64_t>(n_obj + 1);
    int64_t* d_blk = a.take<int64_t>(n_obj + 1);
    double* d_partial = a.take<double>(los_partial_count(n_star));
    // first CTA of every object (the offsets are host arrays; pageable sources are staged by the runtime before
    // cudaMemcpyAsync returns, so the vector may go out of scope)
    std::vector<int64_t> h_blk((size_t)n_obj + 1);
    int64_t blocks = 0;
    for (int64_t g = 0; g < n_obj; ++g) {
        h_blk[g] = blocks;
        blocks += (star_off_host[g + 1] - star_off_host[g] + kLosThreads * kLosStars - 1) / (kLosThreads * kLosStars);
    }
    h_blk[n_obj] = blocks;
    if (blocks >= ((int64_t)1 << 31)) return SO_ERR_TOO_LARGE;
    SO_CUDA(cudaMemcpyAsync(d_star, star_off_host, sizeof(int64_t) * (size_t)(n_obj + 1), cudaMemcpyHostToDevice, st));
    SO_CUDA(cudaMemcpyAsync(d_gas, gas_off_host, sizeof(int64_t) * (size_t)(n_obj + 1), cudaMemcpyHostToDevice, st));
    SO_CUDA(cudaMemcpyAsync(d_blk, h_blk.data(), sizeof(int64_t) * (size_t)(n_obj + 1), cudaMemcpyHostToDevice, st));
    const int n_chunk = los_chunks(blocks);
    if (star_off_host[0] != 0) return SO_ERR_BAD_ARG;    // the partial buffer is indexed by the star's position
    {
        SO_PROF("los_kernel", st);
        los_kernel<<<dim3((unsigned)blocks, (unsigned)n_chunk), kLosThreads, 0, st>>>(
            xs, ys, zs, d_star, xg, yg, zg, mg, Zg, hg, d_gas, d_blk, n_obj, table, nt, scale, n_chunk, n_star,
            n_chunk == 1 ? out : d_partial);
    }
    if (n_chunk > 1) {
        SO_LAUNCH_CHECK();
        SO_PROF("los_sum_kernel", st);
        los_sum_kernel<<<(unsigned)((n_star + 255) / 256), 256, 0, st>>>(d_partial, n_chunk, n_star, scale, out);
    }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

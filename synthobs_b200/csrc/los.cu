// Line-of-sight metal column density of star particles: the quantity the reference's loaders
// read as MetSurfaceDensities.npy (synthobs/core.py:17-21, :33-37) and its examples turn into the ISM optical depth
// (examples/SED_luminosities.py:69).  The reference does not compute it; the specification is oracle/los_oracle.py:
//
//   Sigma_i = scale * sum_{j : z_j > z_i}  m_j Z_j / h_j^2 * F(b_ij / h_j),   b_ij = |(x_i - x_j, y_i - y_j)|,
//
// F = the cubic-spline SPH kernel (support r < h) projected along the line of sight, tabulated at NT + 1 nodes and
// interpolated linearly.
//
// One thread = one star, one CTA = 128 stars of one object; the object's gas particles pass through shared memory in
// tiles of 128 (one coalesced load per thread and tile: position relative to the object's first gas particle and
// as FP32 head + tail, 1 / h^2, m Z / h^2 in FP32, z in float64), and every thread walks the tile with broadcast loads:
// 2 subtractions, a multiply-add and one multiply per pair on the heads, four pairs per round with one branch; the few pairs inside a kernel's
// support (b < h, in front) redo the difference with the tails (coordinates far from the origin lose nothing), then take
// one rsqrt and two table loads.  The z test is made on the float64 inputs (a particle counts or does not: no FP32
// rounding decides that), the smooth part is FP32, summed per tile in FP32 and over the tiles in float64 in a fixed
// order: bitwise reproducible.  Compute-bound: N_star x N_gas pairs at ~9 instructions (bench.py `los` record).
#include <vector>

#include "common.cuh"

namespace {

constexpr int kLosThreads = 128;
constexpr int kLosMaxTable = 2048;

__global__ void __launch_bounds__(kLosThreads) los_kernel(
        const double* __restrict__ xs, const double* __restrict__ ys, const double* __restrict__ zs,
        const int64_t* __restrict__ star_off, const double* __restrict__ xg, const double* __restrict__ yg,
        const double* __restrict__ zg, const double* __restrict__ mg, const double* __restrict__ Zg,
        const double* __restrict__ hg, const int64_t* __restrict__ gas_off, const int64_t* __restrict__ blk_start,
        int64_t n_obj, const float* __restrict__ table, int nt, double scale, double* __restrict__ out) {
    __shared__ float4 s_g[kLosThreads];          // x, y (relative to the object's origin, FP32 head), 1 / h^2, m Z / h^2
    __shared__ float2 s_lo[kLosThreads];         // FP32 tails of x, y: head + tail carries the float64 coordinate to 2^-48
    __shared__ double s_z[kLosThreads];
    __shared__ float s_tab[kLosMaxTable + 1];
    for (int i = threadIdx.x; i <= nt; i += kLosThreads) s_tab[i] = table[i];
    // object of this CTA: last g with blk_start[g] <= blockIdx.x
    int64_t lo = 0, hi = n_obj;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (blk_start[mid] <= (int64_t)blockIdx.x) lo = mid; else hi = mid;
    }
    const int64_t g = lo;
    const int64_t s0 = star_off[g], s1 = star_off[g + 1], g0 = gas_off[g], g1 = gas_off[g + 1];
    const int64_t i = s0 + ((int64_t)blockIdx.x - blk_start[g]) * kLosThreads + threadIdx.x;
    const bool valid = i < s1;
    double ox = 0.0, oy = 0.0;
    if (g1 > g0) { ox = xg[g0]; oy = yg[g0]; }
    float sx = 0.f, sy = 0.f, sxl = 0.f, syl = 0.f;
    double sz = 0.0;
    if (valid) {
        const double rx = xs[i] - ox, ry = ys[i] - oy;
        sx = (float)rx; sy = (float)ry;
        sxl = (float)(rx - (double)sx); syl = (float)(ry - (double)sy);
        sz = zs[i];
    }
    const float fnt = (float)nt;
    double total = 0.0;
    for (int64_t t0 = g0; t0 < g1; t0 += kLosThreads) {
        const int cnt = (int)min((int64_t)kLosThreads, g1 - t0);
        __syncthreads();
        if (threadIdx.x < cnt) {
            const int64_t j = t0 + threadIdx.x;
            const double h = hg[j];
            const double ih2 = 1.0 / (h * h);
            const double rx = xg[j] - ox, ry = yg[j] - oy;
            const float fx = (float)rx, fy = (float)ry;
            s_g[threadIdx.x] = make_float4(fx, fy, (float)ih2, (float)(mg[j] * Zg[j] * ih2));
            s_lo[threadIdx.x] = make_float2((float)(rx - (double)fx), (float)(ry - (double)fy));
            s_z[threadIdx.x] = zg[j];
        }
        __syncthreads();
        float acc = 0.f;
        // a pair inside (or at the edge of) a kernel's support, in front of the star: redo the difference with the tails
        // (exact to ~1e-7 of h whatever the distance from the origin), then the interpolated kernel
        auto inside = [&](int j, const float4 p, float dxc, float dyc) {
            if (!(s_z[j] > sz)) return;
            const float2 l = s_lo[j];
            const float dx = dxc + (l.x - sxl), dy = dyc + (l.y - syl);
            const float q2 = fmaf(dx, dx, dy * dy) * p.z;               // (b / h)^2
            if (q2 < 1.f) {
                const float t = q2 * rsqrtf(fmaxf(q2, 1e-30f)) * fnt;    // u NT, u = sqrt(q2)
                const int k = min((int)t, nt - 1);
                const float f = t - (float)k;
                const float a = s_tab[k];
                acc = fmaf(p.w, fmaf(f, s_tab[k + 1] - a, a), acc);
            }
        };
        // four pairs per round on the FP32 heads (off by <= 2^-23 of the coordinates: the test is widened), ONE branch
        // for the round: the four loads and distance chains are independent and overlap
        int j = 0;
        for (; j + 4 <= cnt; j += 4) {
            float4 p[4];
            float dxc[4], dyc[4], q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = s_g[j + u];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                dxc[u] = p[u].x - sx; dyc[u] = p[u].y - sy;
                q[u] = fmaf(dxc[u], dxc[u], dyc[u] * dyc[u]) * p[u].z;
            }
            if (fminf(fminf(q[0], q[1]), fminf(q[2], q[3])) < 1.002f) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (q[u] < 1.002f) inside(j + u, p[u], dxc[u], dyc[u]);
            }
        }
        for (; j < cnt; ++j) {
            const float4 p = s_g[j];
            const float dxc = p.x - sx, dyc = p.y - sy;
            if (fmaf(dxc, dxc, dyc * dyc) * p.z < 1.002f) inside(j, p, dxc, dyc);
        }
        total += (double)acc;
    }
    if (valid) out[i] = scale * total;
}

}  // namespace

extern "C" int64_t so_los_workspace_bytes(int64_t n_obj) {
    return 3 * so_align_up((n_obj + 1) * (int64_t)sizeof(int64_t), 256) + 256;
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
    if (!workspace || workspace_bytes < so_los_workspace_bytes(n_obj)) return SO_ERR_WORKSPACE;
    if (gas_off_host[n_obj] > gas_off_host[0] && (!xg || !yg || !zg || !mg || !Zg || !hg)) return SO_ERR_BAD_ARG;
    SoArena a(workspace, workspace_bytes);
    int64_t* d_star = a.take<int64_t>(n_obj + 1);
    int64_t* d_gas = a.take<int64_t>(n_obj + 1);
    int64_t* d_blk = a.take<int64_t>(n_obj + 1);
    // first CTA of every object (the offsets are host arrays; pageable sources are staged by the runtime before
    // cudaMemcpyAsync returns, so the vector may go out of scope)
    std::vector<int64_t> h_blk((size_t)n_obj + 1);
    int64_t blocks = 0;
    for (int64_t g = 0; g < n_obj; ++g) {
        h_blk[g] = blocks;
        blocks += (star_off_host[g + 1] - star_off_host[g] + kLosThreads - 1) / kLosThreads;
    }
    h_blk[n_obj] = blocks;
    if (blocks >= ((int64_t)1 << 31)) return SO_ERR_TOO_LARGE;
    SO_CUDA(cudaMemcpyAsync(d_star, star_off_host, sizeof(int64_t) * (size_t)(n_obj + 1), cudaMemcpyHostToDevice, st));
    SO_CUDA(cudaMemcpyAsync(d_gas, gas_off_host, sizeof(int64_t) * (size_t)(n_obj + 1), cudaMemcpyHostToDevice, st));
    SO_CUDA(cudaMemcpyAsync(d_blk, h_blk.data(), sizeof(int64_t) * (size_t)(n_obj + 1), cudaMemcpyHostToDevice, st));
    {
        SO_PROF("los_kernel", st);
        los_kernel<<<(unsigned)blocks, kLosThreads, 0, st>>>(xs, ys, zs, d_star, xg, yg, zg, mg, Zg, hg, d_gas, d_blk, n_obj, table, nt,
                                                              scale, out);
    }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

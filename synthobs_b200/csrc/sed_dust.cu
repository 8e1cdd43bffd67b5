// K6 — dust-attenuated spectra of generate_SED (sm_100a).
//
// Reference behaviour reproduced: synthobs/sed/models.py:285-310.  Per particle i and
// wavelength l the reference forms T_ISM = exp(-tauV_ISM_i k_ISM(l)), T_BC = exp(-tauV_BC_i k_BC(l))
// and accumulates
//   BC           += M (fesc inc + T_BC (1-fesc)(tra+neb))     young      | M inc   old
//   total        += T_ISM * (the BC term)
//   no_HII_no_BC += T_ISM * M inc
// T does not factor through the cell histogram, so this is N x n_lam elementwise work
// (MUFU/FP32 bound).  Particles arrive sorted by (young, cell); per run of equal keys a
// thread keeps D1 = sum M T_ISM, D2 = sum M T_ISM T_BC, D3 = sum M T_BC, Ms = sum M for its
// wavelengths in registers and folds them into the grid rows once per run, so each SSP row is
// read once per (run, particle block) instead of once per particle.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "exp2.cuh"

namespace {

constexpr int kDustThreads = 256;
constexpr int kStage = 512;

struct DustParams {
    const double* masses;
    const double* tau_ism;
    const double* tau_bc;
    const int32_t* key;      // sorted: cell + (young << 16)
    const int32_t* order;    // sorted position -> particle
    int64_t n;
    const float* inc;        // [ncell, n_lam]
    const float* tn;
    int32_t ncell, n_lam;
    const float* k_ism;      // [n_lam] or null
    const float* k_bc;
    float fesc;
    float* partial;          // [n_seg * pblocks, 3, n_lam]
    int32_t pblocks;         // particle blocks per galaxy
    const int64_t* seg;      // CSR offsets of the galaxies in the sorted order, or null (one galaxy)
    int64_t y_base;          // first (galaxy, particle block) pair of this launch (gridDim.y <= 65535)
};

// kLamPerThread wavelengths per thread (strided by the CTA width, so grid rows are read coalesced).
// Particles are staged in shared memory as 16-byte records {M, tauV_ISM, tauV_BC, key} (one
// broadcast LDS.128 per particle) together with a bit mask of the run ends (ballot), so the CTA
// walks the sorted list run by run: the inner loops carry no key test, the fold into the grid rows
// sits between runs, and what remains per particle and wavelength is the exponential itself plus
// half a packed FMUL2/FFMA2.
// HYBRID: every fourth particle of a run evaluates its exponentials with the FMA-pipe polynomial
// instead of MUFU.EX2 (the XU pipe is the roof; the two pipes overlap).
constexpr int kLamPerThread = 4;

template <bool HYBRID>
__global__ void __launch_bounds__(kDustThreads) dust_spectra_kernel(const DustParams p) {
    __shared__ float4 s_rec[kStage];
    __shared__ unsigned s_end[kStage / 32];
    const int lam0 = blockIdx.x * (kDustThreads * kLamPerThread) + threadIdx.x;
    const int64_t yb = p.y_base + blockIdx.y;
    const int gal = (int)(yb / p.pblocks), pb = (int)(yb % p.pblocks);
    const int64_t g0 = p.seg ? p.seg[gal] : 0, g1 = p.seg ? p.seg[gal + 1] : p.n;
    const int64_t per = (g1 - g0 + p.pblocks - 1) / p.pblocks;
    const int64_t p0 = min(g1, g0 + (int64_t)pb * per), p1 = min(g1, p0 + per);
    const float log2e = 1.4426950408889634f;
    // -k log2(e) per owned wavelength, as pairs (0 where the wavelength is out of range or no dust)
    float2 kI[kLamPerThread / 2], kB[kLamPerThread / 2];
    bool ok[kLamPerThread];
#pragma unroll
    for (int j = 0; j < kLamPerThread; ++j) {
        const int lam = lam0 + j * kDustThreads;
        ok[j] = lam < p.n_lam;
        const float a = (p.k_ism && ok[j]) ? -p.k_ism[lam] * log2e : 0.f;
        const float b = (p.k_bc && ok[j]) ? -p.k_bc[lam] * log2e : 0.f;
        if (j & 1) { kI[j / 2].y = a; kB[j / 2].y = b; } else { kI[j / 2].x = a; kB[j / 2].x = b; }
    }
    const float fesc = p.fesc, ofe = 1.f - p.fesc;

    float2 bc[kLamPerThread / 2], tot[kLamPerThread / 2], noh[kLamPerThread / 2];
    float2 d1[kLamPerThread / 2], d2[kLamPerThread / 2], d3[kLamPerThread / 2];
    const float2 z2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < kLamPerThread / 2; ++h) { bc[h] = tot[h] = noh[h] = d1[h] = d2[h] = d3[h] = z2; }
    float ms = 0.f;
    int cur = -1;

    // fold the run's sums into the grid rows:  D1 = sum M T_ISM, D2 = sum M T_ISM T_BC, D3 = sum M T_BC
    auto flush = [&]() {
        if (cur >= 0) {
            const int cell = cur & 65535;
            const bool young = (cur >> 16) != 0;
#pragma unroll
            for (int j = 0; j < kLamPerThread; ++j) {
                if (!ok[j]) continue;
                const int lam = lam0 + j * kDustThreads;
                const float inc = p.inc[(size_t)cell * p.n_lam + lam];
                float& rbc = (j & 1) ? bc[j / 2].y : bc[j / 2].x;
                float& rtot = (j & 1) ? tot[j / 2].y : tot[j / 2].x;
                float& rnoh = (j & 1) ? noh[j / 2].y : noh[j / 2].x;
                const float v1 = (j & 1) ? d1[j / 2].y : d1[j / 2].x;
                const float v2 = (j & 1) ? d2[j / 2].y : d2[j / 2].x;
                const float v3 = (j & 1) ? d3[j / 2].y : d3[j / 2].x;
                rnoh = fmaf(v1, inc, rnoh);
                if (young) {
                    const float tn = p.tn[(size_t)cell * p.n_lam + lam];
                    rbc += fesc * ms * inc + ofe * v3 * tn;
                    rtot += fesc * v1 * inc + ofe * v2 * tn;
                } else {
                    rbc = fmaf(ms, inc, rbc);
                    rtot = fmaf(v1, inc, rtot);
                }
            }
        }
        ms = 0.f;
#pragma unroll
        for (int h = 0; h < kLamPerThread / 2; ++h) { d1[h] = d2[h] = d3[h] = z2; }
    };
    auto old_one = [&](const float4 r, const bool poly) {
        const float2 m2 = make_float2(r.x, r.x), ti2 = make_float2(r.y, r.y);
        ms += r.x;
#pragma unroll
        for (int h = 0; h < kLamPerThread / 2; ++h) {
            const float2 a = __fmul2_rn(ti2, kI[h]);
            d1[h] = __ffma2_rn(m2, poly ? ex2_poly2(a) : ex2_mufu2(a), d1[h]);
        }
    };
    auto young_one = [&](const float4 r, const bool poly) {
        const float2 m2 = make_float2(r.x, r.x), ti2 = make_float2(r.y, r.y), tb2 = make_float2(r.z, r.z);
        ms += r.x;
#pragma unroll
        for (int h = 0; h < kLamPerThread / 2; ++h) {
            const float2 a = __fmul2_rn(ti2, kI[h]), b = __fmul2_rn(tb2, kB[h]);
            const float2 TI = poly ? ex2_poly2(a) : ex2_mufu2(a);
            d1[h] = __ffma2_rn(m2, TI, d1[h]);
            const float2 mTB = __fmul2_rn(m2, poly ? ex2_poly2(b) : ex2_mufu2(b));
            d3[h] = __fadd2_rn(d3[h], mTB);
            d2[h] = __ffma2_rn(mTB, TI, d2[h]);
        }
    };

    for (int64_t base = p0; base < p1; base += kStage) {
        const int cnt = (int)min((int64_t)kStage, p1 - base);
        __syncthreads();
        for (int t = threadIdx.x; t < kStage; t += kDustThreads) {
            const int64_t s = base + t;
            int key = -2, next = -3;
            if (t < cnt) {
                const int i = p.order[s];
                key = p.key[s];
                next = (t + 1 < cnt) ? p.key[s + 1] : -3;
                s_rec[t] = make_float4((float)p.masses[i], p.tau_ism ? (float)p.tau_ism[i] : 0.f,
                                       p.tau_bc ? (float)p.tau_bc[i] : 0.f, __int_as_float(key));
            }
            const unsigned ends = __ballot_sync(0xffffffffu, t < cnt && key != next);   // last particle of a run
            if ((t & 31) == 0) s_end[t >> 5] = ends;
        }
        __syncthreads();
        int q = 0;
        while (q < cnt) {                                  // one iteration per run (uniform across the CTA)
            const int key = __float_as_int(s_rec[q].w);
            if (key != cur) { flush(); cur = key; }
            int w = q >> 5;
            unsigned bits = s_end[w] >> (q & 31);
            int e = q;
            if (bits) {
                e = q + __ffs(bits);
            } else {
                do { ++w; bits = s_end[w]; } while (!bits);
                e = (w << 5) + __ffs(bits);
            }
            if (cur >> 16) {
                for (; q + 4 <= e; q += 4) {
                    const float4 r0 = s_rec[q], r1 = s_rec[q + 1], r2 = s_rec[q + 2], r3 = s_rec[q + 3];
                    young_one(r0, false); young_one(r1, false); young_one(r2, false); young_one(r3, HYBRID);
                }
#pragma unroll 1
                for (; q < e; ++q) young_one(s_rec[q], false);
            } else {
                for (; q + 4 <= e; q += 4) {
                    const float4 r0 = s_rec[q], r1 = s_rec[q + 1], r2 = s_rec[q + 2], r3 = s_rec[q + 3];
                    old_one(r0, false); old_one(r1, false); old_one(r2, false); old_one(r3, HYBRID);
                }
#pragma unroll 1
                for (; q < e; ++q) old_one(s_rec[q], false);
            }
        }
    }
    flush();
    float* out = p.partial + (size_t)yb * 3 * p.n_lam;
#pragma unroll
    for (int j = 0; j < kLamPerThread; ++j) {
        if (!ok[j]) continue;
        const int lam = lam0 + j * kDustThreads;
        out[lam] = (j & 1) ? bc[j / 2].y : bc[j / 2].x;
        out[p.n_lam + lam] = (j & 1) ? tot[j / 2].y : tot[j / 2].x;
        out[2 * p.n_lam + lam] = (j & 1) ? noh[j / 2].y : noh[j / 2].x;
    }
}

__global__ void dust_reduce_kernel(const float* __restrict__ partial, int pblocks, int n_lam,
                                   float* __restrict__ o_bc, float* __restrict__ o_tot, float* __restrict__ o_noh,
                                   int64_t gal_base) {
    const int lam = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gal = gal_base + blockIdx.y;
    if (lam >= n_lam) return;
    double a = 0, b = 0, c = 0;
    for (int pb = 0; pb < pblocks; ++pb) {
        const float* s = partial + ((size_t)gal * pblocks + pb) * 3 * n_lam;
        a += s[lam]; b += s[n_lam + lam]; c += s[2 * n_lam + lam];
    }
    const size_t o = (size_t)gal * n_lam + lam;
    if (o_bc) o_bc[o] = (float)a;
    if (o_tot) o_tot[o] = (float)b;
    if (o_noh) o_noh[o] = (float)c;
}

// SO_DUST_HYBRID=0 selects the MUFU-only variant (profiling aid); default is the hybrid kernel
static bool so_dust_hybrid() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SO_DUST_HYBRID"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

static int dust_pblocks(int64_t n, int64_t n_seg, int n_lam) {
    const int64_t lam_blocks = (n_lam + kDustThreads * 4 - 1) / (kDustThreads * 4);
    int64_t pb = (4 * 148 + lam_blocks * n_seg - 1) / (lam_blocks * n_seg);
    const int64_t max_pb = (n / (n_seg > 0 ? n_seg : 1) + 63) / 64;
    if (pb > max_pb) pb = (int)max_pb;
    if (pb < 1) pb = 1;
    return (int)pb;
}

}  // namespace

extern "C" int64_t so_sed_dust_spectra_workspace_bytes(int64_t n, int64_t n_seg, int32_t n_lam) {
    return so_align_up(n_seg * dust_pblocks(n, n_seg, n_lam) * 3 * n_lam * (int64_t)sizeof(float), 256) + 256;
}

extern "C" int so_sed_dust_spectra(const double* masses, const double* tau_ism, const double* tau_bc,
                                   const int32_t* cell_sorted, const int32_t* order, int64_t n,
                                   const int64_t* seg_offsets, int64_t n_seg,
                                   const float* grid_inc, const float* grid_tn, int32_t ncell, int32_t n_lam,
                                   const float* k_ism_lam, const float* k_bc_lam, double fesc,
                                   float* out_bc, float* out_total, float* out_nohii,
                                   void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 0 || ncell < 1 || n_lam < 1 || n_seg < 1 || !grid_inc || !grid_tn) return SO_ERR_BAD_ARG;
    if (!seg_offsets && n_seg != 1) return SO_ERR_BAD_ARG;
    if (n > 0 && (!masses || !cell_sorted || !order)) return SO_ERR_BAD_ARG;
    if (n == 0) {
        if (out_bc) SO_CUDA(cudaMemsetAsync(out_bc, 0, sizeof(float) * n_lam * n_seg, st));
        if (out_total) SO_CUDA(cudaMemsetAsync(out_total, 0, sizeof(float) * n_lam * n_seg, st));
        if (out_nohii) SO_CUDA(cudaMemsetAsync(out_nohii, 0, sizeof(float) * n_lam * n_seg, st));
        return SO_OK;
    }
    const int pb = dust_pblocks(n, n_seg, n_lam);
    SoArena a(workspace, workspace_bytes);
    float* partial = a.take<float>(n_seg * pb * 3 * n_lam);
    if (!a.ok() || !partial) return SO_ERR_WORKSPACE;
    DustParams p{};
    p.masses = masses; p.tau_ism = tau_ism; p.tau_bc = tau_bc; p.key = cell_sorted; p.order = order; p.n = n;
    p.inc = grid_inc; p.tn = grid_tn; p.ncell = ncell; p.n_lam = n_lam; p.k_ism = k_ism_lam; p.k_bc = k_bc_lam;
    p.fesc = (float)fesc; p.partial = partial; p.pblocks = pb; p.seg = seg_offsets;
    const unsigned gx = (n_lam + kDustThreads * kLamPerThread - 1) / (kDustThreads * kLamPerThread);
    for (int64_t y0 = 0; y0 < n_seg * pb; y0 += 65535) {        // gridDim.y <= 65535: (galaxy, particle block) pairs in slabs
        p.y_base = y0;
        dim3 grid(gx, (unsigned)std::min<int64_t>(65535, n_seg * pb - y0));
        if (so_dust_hybrid()) { SO_PROF("dust_spectra_kernel", st); dust_spectra_kernel<true><<<grid, kDustThreads, 0, st>>>(p); }
        else { SO_PROF("dust_spectra_kernel", st); dust_spectra_kernel<false><<<grid, kDustThreads, 0, st>>>(p); }
        SO_LAUNCH_CHECK();
    }
    for (int64_t g0 = 0; g0 < n_seg; g0 += 65535) {
        dim3 rgrid((n_lam + 255) / 256, (unsigned)std::min<int64_t>(65535, n_seg - g0));
        { SO_PROF("dust_reduce_kernel", st); dust_reduce_kernel<<<rgrid, 256, 0, st>>>(partial, pb, n_lam, out_bc, out_total, out_nohii, g0); }
        SO_LAUNCH_CHECK();
    }
    return SO_OK;
}

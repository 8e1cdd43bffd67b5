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
#include "common.cuh"

namespace {

constexpr int kDustThreads = 256;
constexpr int kStage = 256;

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
};

__device__ __forceinline__ float ex2a(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// kLamPerThread wavelengths per thread (strided by the CTA width, so grid rows are read coalesced):
// the particle record is fetched from shared memory once per 4 wavelengths and the arithmetic runs
// on wavelength pairs with the packed FMUL2/FFMA2 instructions; what remains per particle and
// wavelength is essentially the ex2 itself (MUFU-bound).
constexpr int kLamPerThread = 4;

__global__ void __launch_bounds__(kDustThreads) dust_spectra_kernel(const DustParams p) {
    __shared__ float s_m[kStage], s_ti[kStage], s_tb[kStage];
    __shared__ int s_key[kStage];
    const int lam0 = blockIdx.x * (kDustThreads * kLamPerThread) + threadIdx.x;
    const int gal = blockIdx.y / p.pblocks, pb = blockIdx.y % p.pblocks;
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

    auto flush = [&]() {
        if (cur < 0) return;
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
    };

    for (int64_t base = p0; base < p1; base += kStage) {
        const int cnt = (int)min((int64_t)kStage, p1 - base);
        __syncthreads();
        if (threadIdx.x < cnt) {
            const int64_t s = base + threadIdx.x;
            const int i = p.order[s];
            s_m[threadIdx.x] = (float)p.masses[i];
            s_ti[threadIdx.x] = p.tau_ism ? (float)p.tau_ism[i] : 0.f;
            s_tb[threadIdx.x] = p.tau_bc ? (float)p.tau_bc[i] : 0.f;
            s_key[threadIdx.x] = p.key[s];
        }
        __syncthreads();
        for (int q = 0; q < cnt; ++q) {
            const int key = s_key[q];
            if (key != cur) {           // uniform across the CTA
                flush();
                cur = key; ms = 0.f;
#pragma unroll
                for (int h = 0; h < kLamPerThread / 2; ++h) { d1[h] = d2[h] = d3[h] = z2; }
            }
            const float m = s_m[q];
            const float2 m2 = make_float2(m, m), ti2 = make_float2(s_ti[q], s_ti[q]);
            ms += m;
            if (key >> 16) {
                const float2 tb2 = make_float2(s_tb[q], s_tb[q]);
#pragma unroll
                for (int h = 0; h < kLamPerThread / 2; ++h) {
                    const float2 aI = __fmul2_rn(ti2, kI[h]);
                    const float2 TI = make_float2(ex2a(aI.x), ex2a(aI.y));
                    d1[h] = __ffma2_rn(m2, TI, d1[h]);
                    const float2 aB = __fmul2_rn(tb2, kB[h]);
                    const float2 mTB = __fmul2_rn(m2, make_float2(ex2a(aB.x), ex2a(aB.y)));
                    d3[h] = __fadd2_rn(d3[h], mTB);
                    d2[h] = __ffma2_rn(mTB, TI, d2[h]);
                }
            } else {
#pragma unroll
                for (int h = 0; h < kLamPerThread / 2; ++h) {
                    const float2 aI = __fmul2_rn(ti2, kI[h]);
                    d1[h] = __ffma2_rn(m2, make_float2(ex2a(aI.x), ex2a(aI.y)), d1[h]);
                }
            }
        }
    }
    flush();
    float* out = p.partial + (size_t)blockIdx.y * 3 * p.n_lam;
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
                                   float* __restrict__ o_bc, float* __restrict__ o_tot, float* __restrict__ o_noh) {
    const int lam = blockIdx.x * blockDim.x + threadIdx.x;
    const int gal = blockIdx.y;
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
    dim3 grid((n_lam + kDustThreads * kLamPerThread - 1) / (kDustThreads * kLamPerThread), (unsigned)(n_seg * pb));
    dust_spectra_kernel<<<grid, kDustThreads, 0, st>>>(p);
    SO_LAUNCH_CHECK();
    dim3 rgrid((n_lam + 255) / 256, (unsigned)n_seg);
    dust_reduce_kernel<<<rgrid, 256, 0, st>>>(partial, pb, n_lam, out_bc, out_total, out_nohii);
    SO_LAUNCH_CHECK();
    return SO_OK;
}

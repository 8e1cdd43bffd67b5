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

__global__ void __launch_bounds__(kDustThreads) dust_spectra_kernel(const DustParams p) {
    __shared__ float s_m[kStage], s_ti[kStage], s_tb[kStage];
    __shared__ int s_key[kStage];
    const int lam = blockIdx.x * kDustThreads + threadIdx.x;
    const bool lam_ok = lam < p.n_lam;
    const int gal = blockIdx.y / p.pblocks, pb = blockIdx.y % p.pblocks;
    const int64_t g0 = p.seg ? p.seg[gal] : 0, g1 = p.seg ? p.seg[gal + 1] : p.n;
    const int64_t per = (g1 - g0 + p.pblocks - 1) / p.pblocks;
    const int64_t p0 = min(g1, g0 + (int64_t)pb * per), p1 = min(g1, p0 + per);
    const float log2e = 1.4426950408889634f;
    const float kI = (p.k_ism && lam_ok) ? p.k_ism[lam] * log2e : 0.f;
    const float kB = (p.k_bc && lam_ok) ? p.k_bc[lam] * log2e : 0.f;
    const float fesc = p.fesc, ofe = 1.f - p.fesc;

    float bc = 0.f, tot = 0.f, noh = 0.f;
    float d1 = 0.f, d2 = 0.f, d3 = 0.f, ms = 0.f;
    int cur = -1;

    auto flush = [&]() {
        if (cur < 0 || !lam_ok) return;
        const int cell = cur & 65535;
        const bool young = (cur >> 16) != 0;
        const float inc = p.inc[(size_t)cell * p.n_lam + lam];
        noh = fmaf(d1, inc, noh);
        if (young) {
            const float tn = p.tn[(size_t)cell * p.n_lam + lam];
            bc += fesc * ms * inc + ofe * d3 * tn;
            tot += fesc * d1 * inc + ofe * d2 * tn;
        } else {
            bc = fmaf(ms, inc, bc);
            tot = fmaf(d1, inc, tot);
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
                cur = key; d1 = d2 = d3 = ms = 0.f;
            }
            const float m = s_m[q];
            const float TI = exp2f(-s_ti[q] * kI);
            d1 = fmaf(m, TI, d1);
            ms += m;
            if (key >> 16) {
                const float TB = exp2f(-s_tb[q] * kB);
                const float mTB = m * TB;
                d3 += mTB;
                d2 = fmaf(mTB, TI, d2);
            }
        }
    }
    flush();
    if (lam_ok) {
        float* out = p.partial + (size_t)blockIdx.y * 3 * p.n_lam;
        out[lam] = bc;
        out[p.n_lam + lam] = tot;
        out[2 * p.n_lam + lam] = noh;
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
    const int64_t lam_blocks = (n_lam + kDustThreads - 1) / kDustThreads;
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
    dim3 grid((n_lam + kDustThreads - 1) / kDustThreads, (unsigned)(n_seg * pb));
    dust_spectra_kernel<<<grid, kDustThreads, 0, st>>>(p);
    SO_LAUNCH_CHECK();
    dim3 rgrid((n_lam + 255) / 256, (unsigned)n_seg);
    dust_reduce_kernel<<<rgrid, 256, 0, st>>>(partial, pb, n_lam, out_bc, out_total, out_nohii);
    SO_LAUNCH_CHECK();
    return SO_OK;
}

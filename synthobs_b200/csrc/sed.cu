// SED hot path, particle kernels (sm_100a).
//
// K1+K3  so_sed_broadband : NGP SSP-cell lookup + per-particle broadband values for
//                           all filters + per-galaxy sums (HBM-bound streaming kernel)
// K2     so_sed_hist      : per-galaxy grid-weight histogram (left GEMM operand)
//        so_split_tf32    : hi/lo split for the 3xTF32 GEMM
//
// Reference behaviour reproduced (paths relative to the reference root):
//   synthobs/sed/models.py:88-135, :153-200  per-particle loop of generate_{L,F}nu_array
//   synthobs/sed/models.py:74-85, :139-150   np.sum of generate_{L,F}nu
//   synthobs/sed/models.py:108-114           NGP argmin (via float64 thresholds, see header)
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kThreads = 512;                 // threads per persistent CTA, 2 CTAs per SM
constexpr int kPerThread = 2;                 // particles per thread per work item
constexpr int kItem = kThreads * kPerThread;  // particles per work item
constexpr int kMaxThr = 127;                  // max thresholds per axis (kernel-parameter resident)
constexpr int kLut = 6144;                    // index table entries per axis (256 per octave, 24 octaves)
constexpr int kLutShift = 12;                 // bin = (high 32 bits of the double >> 12): exponent + 8 mantissa bits
constexpr int kPrefetchItems = 1;               // L2 prefetch distance in work items, sums-only variant (measured: 1 best, 0.260 -> 0.241 ms)
constexpr int kYWarp = 32 * (kPerThread + 1); // per-warp buffer of compacted young particles (< 32 left over + one item)

// One NGP axis: ascending float64 thresholds + an index table over bins of the double's own bit
// pattern (sign/exponent/top 8 mantissa bits: 256 bins per octave, exact edges, integer ALU only).
// lut[b] & 127 = #{thr <= lower edge of bin b}; bit 7 is set when a threshold lies inside the bin:
// only then (< 1 % of the particles) are exact float64 compares against thr[] needed.
struct Axis {
    double thr[kMaxThr + 1];
    unsigned char lut[kLut];
    int base;                  // bin = (hi32(v) >> kLutShift) - base, clamped to [0, kLut)
    int n;                     // number of thresholds
};

struct BroadbandParams {
    const double* masses;
    const double* ages;
    const double* metals;
    const double* tau_ism;
    const double* tau_bc;
    int64_t n;
    const int64_t* seg_offsets;   // device, null = single segment
    int64_t n_seg;
    const int64_t* item_start;    // [n_seg+1] device (null when single segment)
    int64_t n_items;              // exact when single segment, else read from item_start[n_seg]
    int32_t na, nz;
    double young_thr;
    const float* tab_inc;         // [F_total, ncell]
    const float* tab_tn;
    int32_t n_filters;
    int32_t f0;                   // first filter of this launch
    int32_t n_filters_total;
    float k_ism[SO_MAX_FILTERS];  // -k * log2(e), 0 when no dust
    float k_bc[SO_MAX_FILTERS];
    float fesc;
    float* out_pf;                // [F_total, n]
    double* partial;              // [n_items, F_total]
    int32_t* out_cell;
    int32_t prefetch;             // L2 prefetch distance in work items (0 = off)
    Axis age, z;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// index = #{thr <= v} (ascending thresholds); NaN / +inf -> 0 like np.argmin over all-NaN/inf distances.
// Split in two so a thread can issue the table loads of all its particles back to back and take the
// (rare) exact branch once: lut_entry() is the load, locate_exact() the float64 decision for flagged bins.
__device__ __forceinline__ int lut_entry(const unsigned char* __restrict__ lut, int base, double v) {
    int b = (__double2hiint(v) >> kLutShift) - base;     // negative v -> negative -> bin 0
    b = max(0, min(kLut - 1, b));                        // +inf / NaN -> last bin; both end bins are flagged
    return lut[b];
}

__device__ __noinline__ int locate_exact(const double* __restrict__ thr, int n, int g, double v) {
    if (!(v < __longlong_as_double(0x7ff0000000000000LL))) return 0;
    while (g > 0 && !(v >= thr[g - 1])) --g;
    while (g < n && v >= thr[g]) ++g;
    return g;
}

__device__ __forceinline__ int locate(const double* __restrict__ thr, const unsigned char* __restrict__ lut, int n,
                                      int base, double v) {
    const int e = lut_entry(lut, base, v);
    return (e & 128) ? locate_exact(thr, n, e & 127, v) : e;
}

__device__ __forceinline__ float2 ex2_approx2(float2 x) { return make_float2(ex2_approx(x.x), ex2_approx(x.y)); }

// Persistent streaming kernel: each CTA walks a contiguous range of work items (kItem particles of
// one galaxy).  The per-filter arithmetic runs on pairs of filters with the sm_100 packed FP32
// instructions (FMUL2 / FFMA2): the kernel is issue/XU-bound, not FLOP-bound.
// Sums-only variant (WRITE_PF = false): only ~1/4 of the particles are young, so their extra
// birth-cloud term (a second ex2 per filter) is not evaluated under a predicate by every warp:
// each warp compacts its young particles (ballot/popc, fixed order, no block barrier) into a small
// private shared-memory ring and evaluates the term 32 particles at a time with all lanes active.
template <int FP, bool WRITE_PF>
__global__ void __launch_bounds__(kThreads, 2) sed_broadband_kernel(const __grid_constant__ BroadbandParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // row stride in float4 units is odd so random cells spread over all bank groups
    constexpr int kStride = ((FP / 4) | 1) * 4;
    const int ncell = p.na * p.nz;
    float* s_inc = reinterpret_cast<float*>(smem_raw);
    float* s_tn = s_inc + (size_t)ncell * kStride;
    double* s_age_thr = reinterpret_cast<double*>(s_tn + (size_t)ncell * kStride);
    double* s_z_thr = s_age_thr + (kMaxThr + 1);
    double* s_red = s_z_thr + (kMaxThr + 1);                       // [kThreads/32][FP]
    unsigned char* s_age_lut = reinterpret_cast<unsigned char*>(s_red + (kThreads / 32) * FP);
    unsigned char* s_z_lut = s_age_lut + kLut;
    float4* s_y = reinterpret_cast<float4*>(s_z_lut + kLut);       // [warps][kYWarp] {m(1-fesc), tauI, tauB, cell}

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    float4* my_y = s_y + warp * kYWarp;
    int ytail = 0;                 // young particles waiting in this warp's buffer (warp-uniform, < 32 between items)
    const unsigned lane_lt = (1u << lane) - 1u;

    // stage the per-filter tables as [cell][filter] (transposed from [filter][cell])
    for (int idx = tid; idx < ncell * FP; idx += kThreads) {
        const int f = idx / ncell, c = idx - f * ncell;   // consecutive threads -> consecutive cells: coalesced
        float a = 0.f, b = 0.f;
        if (f < p.n_filters) {
            a = p.tab_inc[(size_t)(p.f0 + f) * ncell + c];
            b = p.tab_tn[(size_t)(p.f0 + f) * ncell + c];
        }
        s_inc[c * kStride + f] = a;
        s_tn[c * kStride + f] = b;
    }
    for (int i = tid; i <= kMaxThr; i += kThreads) { s_age_thr[i] = p.age.thr[i]; s_z_thr[i] = p.z.thr[i]; }
    for (int i = tid; i < kLut / 4; i += kThreads) {
        reinterpret_cast<unsigned*>(s_age_lut)[i] = reinterpret_cast<const unsigned*>(p.age.lut)[i];
        reinterpret_cast<unsigned*>(s_z_lut)[i] = reinterpret_cast<const unsigned*>(p.z.lut)[i];
    }
    __syncthreads();
    // programmatic dependent launch: everything above (table staging) overlaps the item-scan kernel; its
    // output (item_start) is read from here on
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // contiguous range of work items per CTA so consecutive items share a galaxy
    const bool segmented = p.item_start != nullptr;
    const int64_t n_items = segmented ? p.item_start[p.n_seg] : p.n_items;
    const int64_t per = (n_items + gridDim.x - 1) / gridDim.x;
    const int64_t item_lo = (int64_t)blockIdx.x * per;
    const int64_t item_hi = min(n_items, item_lo + per);
    if (item_lo >= item_hi) return;

    // running (galaxy, segment end, chunk start): advanced incrementally, no per-item lookups
    int64_t gal = 0, s1 = p.n, begin = item_lo * kItem;
    if (segmented) {
        int64_t lo = 0, hi = p.n_seg;             // last galaxy whose first item is <= item_lo
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (p.item_start[mid] <= item_lo) lo = mid; else hi = mid;
        }
        gal = lo;
        const int64_t s0 = p.seg_offsets[gal];
        s1 = p.seg_offsets[gal + 1];
        begin = s0 + (item_lo - p.item_start[gal]) * kItem;
    }

    float2 acc[FP / 2];
#pragma unroll
    for (int f = 0; f < FP / 2; ++f) acc[f] = make_float2(0.f, 0.f);
    const float fesc = p.fesc, ofe = 1.0f - p.fesc;
    const int nf = p.n_filters;

    // L2 prefetch: thread t < 5 * kLines owns line (t % kLines) of input array (t / kLines) of a later item
    constexpr int kLines = kItem * 8 / 128;              // 128-byte lines per array per item
    static_assert(5 * kLines <= kThreads, "one prefetch line per thread");
    const double* pf_line = nullptr;
    int64_t pf_off = 0;
    if (p.prefetch > 0 && tid < 5 * kLines) {
        const int arr = tid / kLines, line = tid - arr * kLines;
        const double* base = arr == 0 ? p.ages : arr == 1 ? p.metals : arr == 2 ? p.masses : arr == 3 ? p.tau_ism : p.tau_bc;
        pf_off = (int64_t)p.prefetch * kItem + line * 16;
        if (base) pf_line = base + pf_off;
    }

    for (int64_t item = item_lo; item < item_hi; ++item) {
        const int count = (int)min((int64_t)kItem, s1 - begin);

        double age[kPerThread], zz[kPerThread];
        float mm[kPerThread], ti[kPerThread], tb[kPerThread];
        if (pf_line && item + p.prefetch < item_hi) {
            // pull a later item's 128-byte lines towards L2 while this one is being processed (contiguous in
            // memory unless a galaxy boundary intervenes: then the hint is merely wasted).  One line per thread,
            // spread over 320 threads: issuing the five lines of a column from 64 threads only was slower.
            if (begin + pf_off < p.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf_line + begin));
        }
        if (count == kItem) {
            // full item (all but the last of a galaxy): plain strided loads off one base offset
            const int64_t i0 = begin + tid;
            const double* pa = p.ages + i0;
            const double* pz = p.metals + i0;
            const double* pm = p.masses + i0;
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) {
                age[j] = ldg_stream(pa + j * kThreads);
                zz[j] = ldg_stream(pz + j * kThreads);
                mm[j] = (float)ldg_stream(pm + j * kThreads);
            }
            if (p.tau_ism) {
                const double* pt = p.tau_ism + i0;
#pragma unroll
                for (int j = 0; j < kPerThread; ++j) ti[j] = (float)ldg_stream(pt + j * kThreads);
            } else {
#pragma unroll
                for (int j = 0; j < kPerThread; ++j) ti[j] = 0.f;
            }
            if (p.tau_bc) {
                const double* pt = p.tau_bc + i0;
#pragma unroll
                for (int j = 0; j < kPerThread; ++j) tb[j] = (float)ldg_stream(pt + j * kThreads);
            } else {
#pragma unroll
                for (int j = 0; j < kPerThread; ++j) tb[j] = 0.f;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) {
                const int64_t i = begin + j * kThreads + tid;
                const bool on = (j * kThreads + tid) < count;
                age[j] = on ? ldg_stream(p.ages + i) : 1.0;
                zz[j] = on ? ldg_stream(p.metals + i) : 1.0;
                mm[j] = on ? (float)ldg_stream(p.masses + i) : 0.f;
                ti[j] = (on && p.tau_ism) ? (float)ldg_stream(p.tau_ism + i) : 0.f;
                tb[j] = (on && p.tau_bc) ? (float)ldg_stream(p.tau_bc + i) : 0.f;
            }
        }
        if constexpr (!WRITE_PF) {
            // sums only.  Phase A: NGP cells of the thread's particles, young ones appended to the warp's buffer
            int cellj[kPerThread];
            float wj[kPerThread];
            int iaj[kPerThread], izj[kPerThread];
            int flagged = 0;
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) {
                iaj[j] = lut_entry(s_age_lut, p.age.base, age[j]);
                izj[j] = lut_entry(s_z_lut, p.z.base, zz[j]);
                flagged |= iaj[j] | izj[j];
            }
            if (flagged & 128) {        // a threshold lies inside one of the bins hit: exact float64 compares
#pragma unroll
                for (int j = 0; j < kPerThread; ++j) {
                    if (iaj[j] & 128) iaj[j] = locate_exact(s_age_thr, p.age.n, iaj[j] & 127, age[j]);
                    if (izj[j] & 128) izj[j] = locate_exact(s_z_thr, p.z.n, izj[j] & 127, zz[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) {
                const int64_t i = begin + j * kThreads + tid;
                const bool on = (j * kThreads + tid) < count;
                const int ia = iaj[j], iz = izj[j];
                const bool young = (age[j] < p.young_thr) && (age[j] >= 0.0);
                const int cell = ia * p.nz + iz;
                if (p.out_cell && on) p.out_cell[i] = cell + (young ? 65536 : 0);
                const bool isy = young && on;
                const unsigned bal = __ballot_sync(0xffffffffu, isy);      // fixed order: slot, lane
                if (isy) my_y[ytail + __popc(bal & lane_lt)] = make_float4(mm[j] * ofe, ti[j], tb[j], __int_as_float(cell));
                ytail += __popc(bal);
                cellj[j] = cell;
                wj[j] = young ? mm[j] * fesc : mm[j];          // m == 0 for lanes past the end
            }
            // Phase B: acc += w 2^(tI kI) inc for all filters, the particles interleaved (independent chains);
            // the birth-cloud term of the young ones follows in the dense pass below
#pragma unroll
            for (int q = 0; q < FP / 4; ++q) {
#pragma unroll
                for (int j = 0; j < kPerThread; ++j) {
                    const float4 vi = reinterpret_cast<const float4*>(s_inc + cellj[j] * kStride)[q];
                    const float2 inc2[2] = {make_float2(vi.x, vi.y), make_float2(vi.z, vi.w)};
                    const float2 w2 = make_float2(wj[j], wj[j]), tI2 = make_float2(ti[j], ti[j]);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int pr = q * 2 + h;                       // filter pair (2 pr, 2 pr + 1)
                        if (q < FP / 4 - 1 || h == 0 || 2 * pr < nf) {  // uniform; only the last quad can be partial
                            const float2 argI = __fmul2_rn(tI2, make_float2(p.k_ism[2 * pr], p.k_ism[2 * pr + 1]));
                            acc[pr] = __ffma2_rn(__fmul2_rn(w2, ex2_approx2(argI)), inc2[h], acc[pr]);
                        }
                    }
                }
            }
        } else {
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) {
            const int64_t i = begin + j * kThreads + tid;
            [[maybe_unused]] const bool on = (j * kThreads + tid) < count;
            // --- NGP cell: float64 compares against the host-built thresholds
            const int ia = locate(s_age_thr, s_age_lut, p.age.n, p.age.base, age[j]);
            const int iz = locate(s_z_thr, s_z_lut, p.z.n, p.z.base, zz[j]);
            const bool young = (age[j] < p.young_thr) && (age[j] >= 0.0);
            const int cell = ia * p.nz + iz;
            if (p.out_cell && on) p.out_cell[i] = cell + (young ? 65536 : 0);
            if constexpr (!WRITE_PF) {
                // warp-level compaction of the young particles (fixed order: slot, lane)
                const bool isy = young && on;
                const unsigned bal = __ballot_sync(0xffffffffu, isy);
                if (isy) my_y[ytail + __popc(bal & lane_lt)] = make_float4(mm[j] * ofe, ti[j], tb[j], __int_as_float(cell));
                ytail += __popc(bal);
            }

            // l = m T_ISM inc                                       (old,   models.py:133)
            // l = m T_ISM (fesc inc + T_BC (1-fesc)(tra+neb))       (young, models.py:131)
            //   = w T_ISM inc + [young] m (1-fesc) 2^(argI+argB) tn,   w = young ? m fesc : m
            const float m = mm[j];                             // m == 0 for lanes past the end
            const float wv = young ? m * fesc : m;
            const float2 w2 = make_float2(wv, wv), mo2 = make_float2(m * ofe, m * ofe);
            const float2 tI2 = make_float2(ti[j], ti[j]), tB2 = make_float2(tb[j], tb[j]);
            const float4* rinc = reinterpret_cast<const float4*>(s_inc + cell * kStride);
            const float4* rtn = reinterpret_cast<const float4*>(s_tn + cell * kStride);
            [[maybe_unused]] float* po = nullptr;
            if constexpr (WRITE_PF) po = p.out_pf + (size_t)p.f0 * p.n + i;
            // FP = n_filters rounded up to 4: every quad but the last is full
#pragma unroll
            for (int q = 0; q < FP / 4; ++q) {
                const float4 vi = rinc[q];
                float4 vt = make_float4(0.f, 0.f, 0.f, 0.f);
                if constexpr (WRITE_PF) { if (young) vt = rtn[q]; }
                const float2 inc2[2] = {make_float2(vi.x, vi.y), make_float2(vi.z, vi.w)};
                const float2 tn2[2] = {make_float2(vt.x, vt.y), make_float2(vt.z, vt.w)};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int pr = q * 2 + h;                       // filter pair (2 pr, 2 pr + 1)
                    if (q < FP / 4 - 1 || h == 0 || 2 * pr < nf) {  // uniform; only the last quad can be partial
                        const float2 argI = __fmul2_rn(tI2, make_float2(p.k_ism[2 * pr], p.k_ism[2 * pr + 1]));
                        const float2 t = __fmul2_rn(w2, ex2_approx2(argI));
                        if constexpr (WRITE_PF) {
                            float2 l = __fmul2_rn(t, inc2[h]);
                            if (young)
                                l = __ffma2_rn(__fmul2_rn(mo2, ex2_approx2(__ffma2_rn(tB2, make_float2(p.k_bc[2 * pr], p.k_bc[2 * pr + 1]), argI))),
                                               tn2[h], l);
                            acc[pr] = __fadd2_rn(acc[pr], l);
                            if (on) {
                                po[(size_t)(2 * pr) * p.n] = l.x;
                                if (2 * pr + 1 < nf) po[(size_t)(2 * pr + 1) * p.n] = l.y;
                            }
                        } else {
                            acc[pr] = __ffma2_rn(t, inc2[h], acc[pr]);     // young term: dense pass below
                        }
                    }
                }
            }
        }

        }
        // advance to the next work item; a run ends with the galaxy (or with this CTA's range)
        int64_t next_begin = begin + kItem;
        bool run_end = (item + 1 == item_hi);
        if (segmented && !run_end && next_begin >= s1) {
            int64_t s0n;
            do { ++gal; s0n = p.seg_offsets[gal]; s1 = p.seg_offsets[gal + 1]; } while (s1 == s0n);
            next_begin = s0n;
            run_end = true;
        }
        begin = next_begin;

        if constexpr (!WRITE_PF) {
            // dense evaluation of the birth-cloud term  m (1-fesc) 2^(tI kI + tB kB) (tra+neb)
            __syncwarp();
            int yhead = 0;
            while (ytail - yhead >= 32 || (run_end && ytail > yhead)) {
                const int cnt = min(32, ytail - yhead);
                if (lane < cnt) {
                    const float4 rec = my_y[yhead + lane];
                    const float2 mo2 = make_float2(rec.x, rec.x), tI2 = make_float2(rec.y, rec.y);
                    const float2 tB2 = make_float2(rec.z, rec.z);
                    const float4* rtn = reinterpret_cast<const float4*>(s_tn + __float_as_int(rec.w) * kStride);
#pragma unroll
                    for (int q = 0; q < FP / 4; ++q) {
                        const float4 vt = rtn[q];
                        const float2 tn2[2] = {make_float2(vt.x, vt.y), make_float2(vt.z, vt.w)};
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int pr = q * 2 + h;
                            if (q < FP / 4 - 1 || h == 0 || 2 * pr < nf) {
                                const float2 arg = __ffma2_rn(tB2, make_float2(p.k_bc[2 * pr], p.k_bc[2 * pr + 1]),
                                                              __fmul2_rn(tI2, make_float2(p.k_ism[2 * pr], p.k_ism[2 * pr + 1])));
                                acc[pr] = __ffma2_rn(__fmul2_rn(mo2, ex2_approx2(arg)), tn2[h], acc[pr]);
                            }
                        }
                    }
                }
                yhead += cnt;
            }
            // fewer than 32 records are left: move them to the front of the buffer
            if (yhead > 0) {
                const int left = ytail - yhead;
                float4 keep = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane < left) keep = my_y[yhead + lane];
                __syncwarp();
                if (lane < left) my_y[lane] = keep;
                ytail = left;
            }
            __syncwarp();
        }

        if (p.partial) {
            if (run_end) {
                // block reduction of the per-thread sums of this run of items
#pragma unroll
                for (int f = 0; f < FP; ++f) {
                    const float v = (f & 1) ? acc[f / 2].y : acc[f / 2].x;
                    const float ws = warp_sum(v);
                    if (lane == 0) s_red[warp * FP + f] = (double)ws;
                }
#pragma unroll
                for (int f = 0; f < FP / 2; ++f) acc[f] = make_float2(0.f, 0.f);
                __syncthreads();
                if (tid < nf) {
                    double sum = 0.0;
                    for (int ww = 0; ww < kThreads / 32; ++ww) sum += s_red[ww * FP + tid];
                    p.partial[item * p.n_filters_total + p.f0 + tid] = sum;
                }
                __syncthreads();
            } else if (tid < nf) {
                p.partial[item * p.n_filters_total + p.f0 + tid] = 0.0;
            }
        }
    }
}

// items per segment -> exclusive scan (single CTA; n_seg is at most a few 1e6)
__global__ void __launch_bounds__(1024) item_scan_kernel(const int64_t* __restrict__ seg, int64_t n_seg,
                                                         int64_t* __restrict__ item_start) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the broadband kernel may stage its tables now
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_seg; base += 1024) {
        const int64_t g = base + tid;
        int64_t v = 0;
        if (g < n_seg) v = (seg[g + 1] - seg[g] + kItem - 1) / kItem;
        int64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int64_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int64_t carry = s_carry;
        const int64_t incl = x + (warp > 0 ? s_warp[warp - 1] : 0) + carry;
        if (g < n_seg) item_start[g] = incl - v;
        __syncthreads();
        if (tid == 1023) s_carry = incl;
        __syncthreads();
    }
    if (tid == 0) item_start[n_seg] = s_carry;
}

// per-galaxy sum of the per-item partials, fixed order -> deterministic
__global__ void __launch_bounds__(256) seg_reduce_kernel(const double* __restrict__ partial,
                                                         const int64_t* __restrict__ item_start,
                                                         int64_t n_items_single, int32_t F,
                                                         double* __restrict__ out) {
    __shared__ double s[256];
    asm volatile("griddepcontrol.wait;" ::: "memory");       // launched ahead of the broadband kernel's end
    const int64_t g = blockIdx.x;
    const int64_t i0 = item_start ? item_start[g] : 0;
    const int64_t i1 = item_start ? item_start[g + 1] : n_items_single;
    for (int f = 0; f < F; ++f) {
        double a = 0.0;
        for (int64_t it = i0 + threadIdx.x; it < i1; it += 256) a += partial[it * F + f];
        s[threadIdx.x] = a;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[g * F + f] = s[0];
        __syncthreads();
    }
}

template <int FP, bool WRITE_PF>
int launch_broadband_impl(const BroadbandParams& p, cudaStream_t st) {
    constexpr int kStride = ((FP / 4) | 1) * 4;
    const size_t ncell = (size_t)p.na * p.nz;
    size_t smem = 2 * ncell * kStride * sizeof(float) + 2 * (size_t)(kMaxThr + 1) * sizeof(double) +
                  (size_t)(kThreads / 32) * FP * sizeof(double) + 2 * kLut + 64 +
                  (WRITE_PF ? 0 : (size_t)(kThreads / 32) * kYWarp * sizeof(float4));
    if (smem > 227 * 1024) return SO_ERR_TOO_LARGE;
    SO_CUDA(cudaFuncSetAttribute(sed_broadband_kernel<FP, WRITE_PF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    int per_sm = (smem <= 113 * 1024) ? 2 : 1;
    int64_t grid = (int64_t)so_num_sms() * per_sm;
    if (grid > p.n_items) grid = p.n_items;
    if (grid < 1) grid = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // only behind our own item-scan kernel (segmented path): the table staging must not overtake a foreign
    // predecessor that might still be writing the tables
    cfg.numAttrs = p.item_start ? 1 : 0;
    { SO_PROF("sed_broadband_kernel", cfg.stream); SO_CUDA(cudaLaunchKernelEx(&cfg, sed_broadband_kernel<FP, WRITE_PF>, p)); }
    return SO_OK;
}

template <int FP>
int launch_broadband(const BroadbandParams& p, cudaStream_t st) {
    return p.out_pf ? launch_broadband_impl<FP, true>(p, st) : launch_broadband_impl<FP, false>(p, st);
}

// host: thresholds (ascending, may end with +inf) -> Axis with its bit-pattern index table
static void fill_axis(Axis& ax, const double* thr, int n) {
    ax.n = n;
    for (int i = 0; i <= kMaxThr; ++i) ax.thr[i] = (i < n) ? thr[i] : __builtin_inf();
    auto hi_bin = [](double v) -> int64_t {
        int64_t bits;
        memcpy(&bits, &v, sizeof(bits));
        return (int64_t)((int32_t)(bits >> 32) >> kLutShift);
    };
    // centre the table on the finite positive thresholds
    double tmin = 0.0, tmax = 0.0;
    bool any = false;
    for (int i = 0; i < n; ++i)
        if (thr[i] > 0 && thr[i] < __builtin_inf()) {
            if (!any) { tmin = tmax = thr[i]; any = true; }
            if (thr[i] < tmin) tmin = thr[i];
            if (thr[i] > tmax) tmax = thr[i];
        }
    if (!any) { tmin = tmax = 1.0; }
    const int64_t b_lo = hi_bin(tmin), b_hi = hi_bin(tmax);
    int64_t base = b_lo - 16;
    if (b_hi - base + 16 >= kLut) base = b_lo - 1;   // very wide grids: bins above the table share the last entry
    ax.base = (int)base;
    int g = 0;
    for (int b = 0; b < kLut; ++b) {
        // lower edge of bin b = the double whose high word is (base + b) << kLutShift, low word 0
        const int64_t hb = (base + b) << kLutShift;
        double e0, e1;
        int64_t bits0 = hb << 32, bits1 = ((base + b + 1) << kLutShift) << 32;
        memcpy(&e0, &bits0, sizeof(e0));
        memcpy(&e1, &bits1, sizeof(e1));
        while (g < n && e0 >= thr[g]) ++g;                   // #{thr <= lower edge}
        int g1 = g;
        while (g1 < n && e1 > thr[g1]) ++g1;                 // thresholds in [e0, e1)
        bool check = (g1 != g);
        if (b == 0 || b == kLut - 1 || hb <= 0) check = true;    // clamped bins cover half-lines; zero/denormal edge
        ax.lut[b] = (unsigned char)(g | (check ? 128 : 0));
    }
}

// The index table of an axis takes ~50 us of host time to build and only depends on the thresholds: keep the
// last one per axis and thread (a model is normally used for many calls in a row).
static void fill_axis_cached(int which, Axis& ax, const double* thr, int n) {
    struct Slot { bool valid; int n; double thr[kMaxThr + 1]; Axis ax; };
    static thread_local Slot slots[2] = {};
    Slot& c = slots[which];
    if (c.valid && c.n == n && (n == 0 || memcmp(c.thr, thr, sizeof(double) * (size_t)n) == 0)) {
        ax = c.ax;
        return;
    }
    fill_axis(ax, thr, n);
    c.valid = true;
    c.n = n;
    if (n > 0) memcpy(c.thr, thr, sizeof(double) * (size_t)n);
    c.ax = ax;
}

__global__ void hist_kernel(const double* __restrict__ masses, const int32_t* __restrict__ cell,
                            const int64_t* __restrict__ seg, int64_t n, int32_t ncell, int32_t chunks_per_seg,
                            double* __restrict__ partial) {
    // one CTA = one (galaxy, chunk); private float64 histogram in shared memory
    extern __shared__ double s_h[];
    const int64_t g = blockIdx.x / chunks_per_seg;
    const int c = blockIdx.x % chunks_per_seg;
    const int64_t s0 = seg ? seg[g] : 0, s1 = seg ? seg[g + 1] : n;
    const int64_t len = s1 - s0;
    const int64_t per = (len + chunks_per_seg - 1) / chunks_per_seg;
    const int64_t b = s0 + c * per, e = min(s1, b + per);
    for (int i = threadIdx.x; i < 2 * ncell; i += blockDim.x) s_h[i] = 0.0;
    __syncthreads();
    for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
        const int cc = cell[i];
        const int idx = (cc & 65535) + ((cc >> 16) ? ncell : 0);
        atomicAdd(&s_h[idx], masses[i]);
    }
    __syncthreads();
    double* out = partial + (size_t)blockIdx.x * 2 * ncell;
    for (int i = threadIdx.x; i < 2 * ncell; i += blockDim.x) out[i] = s_h[i];
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__global__ void hist_finish_kernel(const double* __restrict__ partial, int32_t ncell2, int32_t chunks_per_seg,
                                   float* __restrict__ H, float* __restrict__ H_lo, int64_t ld, int64_t g_base) {
    const int64_t g = g_base + blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell2) return;
    double a = 0.0;
    for (int c = 0; c < chunks_per_seg; ++c) a += partial[((size_t)g * chunks_per_seg + c) * ncell2 + i];
    const float v = (float)a;
    if (H_lo) {
        const float hi = tf32_hi(v);
        H[g * ld + i] = hi;
        H_lo[g * ld + i] = tf32_hi(v - hi);
    } else {
        H[g * ld + i] = v;
    }
}

__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                                  int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    const float h = tf32_hi(v);
    hi[i] = h;
    lo[i] = tf32_hi(v - h);
}

}  // namespace

extern "C" int64_t so_sed_broadband_workspace_bytes(int64_t n, int64_t n_seg, int32_t n_filters) {
    const int64_t max_items = (n + kItem - 1) / kItem + n_seg;
    SoArena a(nullptr, 0);
    a.take<int64_t>(n_seg + 1);
    a.take<int32_t>(max_items);
    a.take<double>(max_items * n_filters);
    return a.used + 256;
}

extern "C" int so_sed_broadband(const double* masses, const double* ages, const double* metals,
                                const double* tau_ism, const double* tau_bc, int64_t n,
                                const int64_t* seg_offsets, int64_t n_seg,
                                const double* age_thr, int32_t na, const double* z_thr, int32_t nz, double young_thr,
                                const float* tab_inc, const float* tab_tn, int32_t n_filters,
                                const double* k_ism, const double* k_bc, double fesc,
                                float* out_pf, double* out_sums, int32_t* out_cell,
                                void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 0 || n_seg < 1 || na < 1 || nz < 1 || n_filters < 1) return SO_ERR_BAD_ARG;
    if (na - 1 > kMaxThr || nz - 1 > kMaxThr || (int64_t)na * nz > 65535) return SO_ERR_TOO_LARGE;
    if (n > 0 && (!masses || !ages || !metals || !tab_inc || !tab_tn || (na > 1 && !age_thr) || (nz > 1 && !z_thr)))
        return SO_ERR_BAD_ARG;
    if (n == 0) {
        if (out_sums) SO_CUDA(cudaMemsetAsync(out_sums, 0, sizeof(double) * n_seg * n_filters, st));
        return SO_OK;
    }
    const bool segmented = seg_offsets != nullptr;
    if (!segmented && n_seg != 1) return SO_ERR_BAD_ARG;
    const int64_t max_items = (n + kItem - 1) / kItem + (segmented ? n_seg : 0);
    SoArena a(workspace, workspace_bytes);
    int64_t* item_start = a.take<int64_t>(n_seg + 1);
    int32_t* item_gal = a.take<int32_t>(max_items);
    double* partial = a.take<double>(max_items * n_filters);
    (void)item_gal;      // (slot kept so the workspace layout of earlier builds still fits)
    if (!a.ok() || (out_sums && !partial) || (segmented && !item_start)) return SO_ERR_WORKSPACE;

    BroadbandParams p{};
    p.masses = masses; p.ages = ages; p.metals = metals; p.tau_ism = tau_ism; p.tau_bc = tau_bc; p.n = n;
    p.seg_offsets = seg_offsets; p.n_seg = n_seg;
    p.na = na; p.nz = nz; p.young_thr = young_thr;
    fill_axis_cached(0, p.age, age_thr, na - 1);
    fill_axis_cached(1, p.z, z_thr, nz - 1);
    p.tab_inc = tab_inc; p.tab_tn = tab_tn; p.n_filters_total = n_filters;
    // the per-particle-output variant is store-bound and loses with the prefetch (measured): off there
    { const char* e = getenv("SO_SED_PREFETCH"); p.prefetch = e ? atoi(e) : (out_pf ? 0 : kPrefetchItems); }
    p.fesc = (float)fesc; p.out_pf = out_pf; p.partial = out_sums ? partial : nullptr; p.out_cell = out_cell;

    if (segmented) {
        { SO_PROF("item_scan_kernel", st); item_scan_kernel<<<1, 1024, 0, st>>>(seg_offsets, n_seg, item_start); }
        SO_LAUNCH_CHECK();
        p.item_start = item_start;
        p.n_items = max_items;  // upper bound for the grid; the exact count is read on the device
    } else {
        p.item_start = nullptr;
        p.n_items = (n + kItem - 1) / kItem;
    }

    // filters per launch: as many as fit the shared-memory tables (<= SO_MAX_FILTERS)
    int fmax = SO_MAX_FILTERS;
    while (fmax > 4) {
        const size_t stride = (size_t)(((fmax / 4) | 1) * 4);
        const size_t smem = 2 * (size_t)na * nz * stride * sizeof(float) + 2 * (size_t)(kMaxThr + 1) * sizeof(double) +
                            (size_t)(kThreads / 32) * fmax * sizeof(double) + 2 * kLut + 64 +
                            (size_t)(kThreads / 32) * kYWarp * sizeof(float4);
        if (smem <= 227 * 1024) break;
        fmax -= 4;
    }
    for (int f0 = 0; f0 < n_filters; f0 += fmax) {
        const int nf = (n_filters - f0 < fmax) ? (n_filters - f0) : fmax;
        p.f0 = f0; p.n_filters = nf;
        const double log2e = 1.4426950408889634;
        for (int f = 0; f < SO_MAX_FILTERS; ++f) {
            p.k_ism[f] = (f < nf && k_ism) ? (float)(-k_ism[f0 + f] * log2e) : 0.f;
            p.k_bc[f] = (f < nf && k_bc) ? (float)(-k_bc[f0 + f] * log2e) : 0.f;
        }
        // only the first chunk writes the cell ids
        if (f0 > 0) p.out_cell = nullptr;
        int rc;
        const int fp = (nf + 3) / 4 * 4;
        switch (fp) {
            case 4: rc = launch_broadband<4>(p, st); break;
            case 8: rc = launch_broadband<8>(p, st); break;
            case 12: rc = launch_broadband<12>(p, st); break;
            case 16: rc = launch_broadband<16>(p, st); break;
            case 20: rc = launch_broadband<20>(p, st); break;
            case 24: rc = launch_broadband<24>(p, st); break;
            case 28: rc = launch_broadband<28>(p, st); break;
            default: rc = launch_broadband<32>(p, st); break;
        }
        if (rc != SO_OK) return rc;
    }
    if (out_sums) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)n_seg);
        cfg.blockDim = dim3(256);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        SO_CUDA(cudaLaunchKernelEx(&cfg, seg_reduce_kernel, (const double*)partial,
                                   (const int64_t*)(segmented ? item_start : nullptr), (int64_t)p.n_items,
                                   (int32_t)n_filters, out_sums));
    }
    return SO_OK;
}

static int hist_chunks(int64_t n_seg) {
    const int64_t target = 2 * (int64_t)so_num_sms();
    int64_t c = (target + n_seg - 1) / n_seg;
    if (c < 1) c = 1;
    if (c > 64) c = 64;
    return (int)c;
}

extern "C" int64_t so_sed_hist_workspace_bytes(int64_t n_seg, int32_t ncell) {
    return so_align_up((int64_t)n_seg * hist_chunks(n_seg) * 2 * ncell * (int64_t)sizeof(double), 256) + 256;
}

extern "C" int so_sed_hist(const double* masses, const int32_t* cell, int64_t n,
                           const int64_t* seg_offsets, int64_t n_seg, int32_t ncell,
                           float* H, float* H_lo, int64_t ld_h,
                           void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 0 || n_seg < 1 || ncell < 1 || !H || ld_h < 2 * (int64_t)ncell) return SO_ERR_BAD_ARG;
    if (!seg_offsets && n_seg != 1) return SO_ERR_BAD_ARG;
    if (n > 0 && (!masses || !cell)) return SO_ERR_BAD_ARG;
    const int chunks = hist_chunks(n_seg);
    SoArena a(workspace, workspace_bytes);
    double* partial = a.take<double>((int64_t)n_seg * chunks * 2 * ncell);
    if (!a.ok() || !partial) return SO_ERR_WORKSPACE;
    const size_t smem = (size_t)2 * ncell * sizeof(double);
    if (smem > 200 * 1024) return SO_ERR_TOO_LARGE;
    SO_CUDA(cudaFuncSetAttribute(hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { SO_PROF("hist_kernel", st); hist_kernel<<<(unsigned)(n_seg * chunks), 256, smem, st>>>(masses, cell, seg_offsets, n, ncell, chunks, partial); }
    SO_LAUNCH_CHECK();
    for (int64_t g0 = 0; g0 < n_seg; g0 += 65535) {       // gridDim.y <= 65535
        dim3 grid((2 * ncell + 255) / 256, (unsigned)std::min<int64_t>(65535, n_seg - g0));
        { SO_PROF("hist_finish_kernel", st); hist_finish_kernel<<<grid, 256, 0, st>>>(partial, 2 * ncell, chunks, H, H_lo, ld_h, g0); }
        SO_LAUNCH_CHECK();
    }
    return SO_OK;
}

extern "C" int so_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream) {
    if (n <= 0) return SO_OK;
    { SO_PROF("split_tf32_kernel", (cudaStream_t)stream); split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, hi, lo, n); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

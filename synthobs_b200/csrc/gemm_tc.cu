// K4/K5 — C[M,N] = A[M,K] * B[N,K]^T on the 5th-generation tensor cores (tcgen05), FP32 accumulate
// in TMEM, error-compensated 3xTF32:  A*B ~= Ahi*Bhi + Ahi*Blo + Alo*Bhi  (hi = TF32-exact part,
// lo = TF32-exact part of the remainder; the caller supplies both, see so_split_tf32).
//
// Replaces the dense contractions of the reference's SED path:
//   hist x SSP grid            synthobs/sed/models.py:277-310 (dust-free spectra of generate_SED)
//   grid x filter weights      synthobs/sed/models.py:39-67   (create_Lnu_grid / create_Fnu_grid)
//
// Accumulation accuracy: the tensor core adds each MMA result into the FP32 TMEM accumulator with
// truncation, a bias of ~1.5e-8 x (number of accumulating MMAs) relative to the sum (measured:
// 5e-5 at K = 9000).  The K loop is therefore cut into chunks of kChunk K-blocks (24 MMAs, bias
// < 4e-7): each chunk accumulates into one of two TMEM buffers, the epilogue warps drain the finished
// chunk into round-to-nearest FP32 running sums in registers while the next chunk is being computed.
//
// Structure (persistent: one CTA per SM walks 128x128 output tiles in a grouped raster order, so the
// CTAs running at the same time share 16 A row-tiles and ~9 B column-tiles in L2; the barrier rings
// run on across tiles, so the producer prefetches the next tile while the last chunk drains; 384 threads):
//   warp 0   TMA producer: per K-block of 32 floats (=128 B, one SWIZZLE_128B atom row) four
//            cp.async.bulk.tensor loads (Ahi, Alo, Bhi, Blo tiles, 16 KB each) into a 3-stage ring
//   warp 1   MMA issuer: one elected thread issues 4 k-steps x 3 tcgen05.mma.kind::tf32 (M=128,N=128,
//            K=8) per stage, commits the stage back to the producer and finally the accumulator
//   warp 2   TMEM allocator (2 x 128 columns)
//   warps 4-11 epilogue/drain: warp w owns TMEM lanes 32*(w%4).. and column half (w-4)/4;
//            tcgen05.ld 32 lanes x 32 columns -> registers (+= running sums); at the end of a tile the
//            sums go through a swizzled 16 KB staging buffer per column half and out with TMA stores
//            (cp.async.bulk.tensor, asynchronous: the warps return to draining the next tile at once)
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int kTileBytes = BM * BK * 4;            // 16 KB: one 128-row operand tile
// kCtas = 1: one CTA per 128x128 tile, stage = {Ahi, Alo, Bhi, Blo} x 16 KB, 3 stages.
// kCtas = 2: a CTA pair (cta_group::2) per 256x128 tile: each CTA holds its own 128 A rows and HALF of the B
//            tile (64 rows), the pair's tensor cores read both halves; stage = 48 KB per CTA, 4 stages.
template <int kCtas> struct Cfg {
    static constexpr int kBTileBytes = kTileBytes / kCtas;
    static constexpr int kStageBytes = 2 * kTileBytes + 2 * kBTileBytes;
    static constexpr int kStages = kCtas == 1 ? 3 : 4;
};
constexpr int kGemmThreads = 384;
constexpr int kAccBufs = 2;                    // TMEM accumulator buffers (four were measured: no gain)
constexpr int kTmemCols = kAccBufs * 128;      // 128-column accumulators
constexpr int kChunk = 2;                      // K-blocks per TMEM accumulation chunk
constexpr int kDrainWarps = 8;
constexpr int kGroupM = 16;                    // row-tiles per raster group
constexpr int kSlabCols = 32;                  // columns per epilogue slab (one 128-byte swizzle row)
constexpr int kSlabBytes = BM * kSlabCols * 4; // 16 KB staging buffer per column half
template <int kCtas> constexpr int smem_bytes() {
    return Cfg<kCtas>::kStages * Cfg<kCtas>::kStageBytes + 2 * kSlabBytes + 1024 /*alignment*/ + 256 /*barriers*/;
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address: rank 0 of the pair

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t it = 0; it < 20000000u; ++it) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start >> 4 | LBO (ignored for swizzled K-major, 1) << 16 | SBO = 8 rows * 128 B >> 4 << 32 | version 1 << 46 | layout 2 << 61
__device__ __forceinline__ uint64_t umma_desc(const void* tile) {
    return (uint64_t)((smem_u32(tile) >> 4) & 0x3fff) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---- cta_group::2 variants (the CTA pair of one 256x128 tile) ----
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    // executed by both CTAs; the transaction bytes are counted on the barrier of rank 0
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {     // arrives on the barrier of BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_rank0(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int kCtas>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_3xtf32_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const __grid_constant__ CUtensorMap tmC,
                   float* __restrict__ C, int64_t ldc, int64_t M, int64_t N, int64_t K, int tiles_m, int tiles_n, int tma_store) {
    constexpr int kStages = Cfg<kCtas>::kStages, kStageBytes = Cfg<kCtas>::kStageBytes, kBTileBytes = Cfg<kCtas>::kBTileBytes;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* slab = smem + kStages * kStageBytes;          // [2][kSlabBytes] epilogue staging (TMA store source)
    uint32_t rank = 0;                                           // CTA rank in the pair
    if constexpr (kCtas == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    uint64_t* bars = reinterpret_cast<uint64_t*>(slab + 2 * kSlabBytes);
    uint64_t* full = bars;                 // [kStages]
    uint64_t* empty = bars + kStages;      // [kStages]
    uint64_t* tmem_full = bars + 2 * kStages;                   // [kAccBufs]
    uint64_t* tmem_empty = bars + 2 * kStages + kAccBufs;       // [kAccBufs]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAccBufs);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = (int)((K + BK - 1) / BK);
    const int nchunk = (nkb + kChunk - 1) / kChunk;
    // tiles_m counts (kCtas * 128)-row tiles.  Grouped raster: tile t -> (m tile, n tile); groups of kGroupM / kCtas
    // row-tiles are swept along n.  m0 = this CTA's first row.
    const int64_t n_tiles = (int64_t)tiles_m * tiles_n;
    const int64_t t_first = blockIdx.x / kCtas, t_step = gridDim.x / kCtas;
    auto tile_origin = [&](int64_t t, int& m0, int& n0) {
        constexpr int kG = kGroupM / kCtas;
        const int64_t per_group = (int64_t)kG * tiles_n;
        const int group = (int)(t / per_group);
        const int first_m = group * kG;
        const int gsz = min(kG, tiles_m - first_m);
        const int r = (int)(t - (int64_t)group * per_group);
        m0 = (first_m + r % gsz) * BM * kCtas + (int)rank * BM;
        n0 = (r / gsz) * BN;
    };

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        // tmem_empty of rank 0 collects the drain warps of the whole pair
        for (int b = 0; b < kAccBufs; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], kDrainWarps * kCtas); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 2) {
        if constexpr (kCtas == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (kCtas == 2) cluster_sync();       // the peer's barriers and TMEM are ready too
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;                      // K-block counter over all of this CTA's tiles (ring position)
            for (int64_t t = t_first; t < n_tiles; t += t_step) {
                int m0, n0;
                tile_origin(t, m0, n0);
                const int nb = n0 + (int)rank * (BN / kCtas);       // this CTA's part of the B tile
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (it / kStages) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    unsigned char* st = smem + s * kStageBytes;
                    if constexpr (kCtas == 2) {
                        // both CTAs' bytes are counted on rank 0's barrier, which its MMA thread waits on
                        if (rank == 0) mbar_expect_tx(&full[s], 2 * kStageBytes);
                        tma_load_2d_pair(st, &tmAh, &full[s], kb * BK, m0);
                        tma_load_2d_pair(st + kTileBytes, &tmAl, &full[s], kb * BK, m0);
                        tma_load_2d_pair(st + 2 * kTileBytes, &tmBh, &full[s], kb * BK, nb);
                        tma_load_2d_pair(st + 2 * kTileBytes + kBTileBytes, &tmBl, &full[s], kb * BK, nb);
                    } else {
                        mbar_expect_tx(&full[s], kStageBytes);
                        tma_load_2d(st, &tmAh, &full[s], kb * BK, m0);
                        tma_load_2d(st + kTileBytes, &tmAl, &full[s], kb * BK, m0);
                        tma_load_2d(st + 2 * kTileBytes, &tmBh, &full[s], kb * BK, nb);
                        tma_load_2d(st + 2 * kTileBytes + kBTileBytes, &tmBl, &full[s], kb * BK, nb);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {             // one thread of the pair's rank 0 issues the MMAs of both CTAs
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, K-major, N>>3, M>>4
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)((BM * kCtas) >> 4) << 24);
            uint32_t it = 0, cc = 0;              // ring position, chunk counter over all tiles
            for (int64_t t = t_first; t < n_tiles; t += t_step) {
                for (int c = 0; c < nchunk; ++c, ++cc) {
                    const int buf = cc % kAccBufs;
                    mbar_wait(&tmem_empty[buf], ((cc / kAccBufs) & 1) ^ 1);      // drained by the epilogue warps
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
                    uint32_t acc = 0;
                    const int kb_end = min(nkb, (c + 1) * kChunk);
                    for (int kb = c * kChunk; kb < kb_end; ++kb, ++it) {
                        const int s = it % kStages;
                        const uint32_t ph = (it / kStages) & 1;
                        mbar_wait(&full[s], ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        unsigned char* st = smem + s * kStageBytes;
                        const uint64_t dAh = umma_desc(st), dAl = umma_desc(st + kTileBytes);
                        const uint64_t dBh = umma_desc(st + 2 * kTileBytes), dBl = umma_desc(st + 2 * kTileBytes + kBTileBytes);
#pragma unroll
                        for (int k4 = 0; k4 < BK / 8; ++k4) {
                            const uint64_t adv = (uint64_t)(k4 * 2);     // 8 tf32 = 32 B = 2 x 16 B along K inside the swizzle atom
                            if constexpr (kCtas == 2) {
                                umma_tf32_pair(tmem_d, dAl + adv, dBh + adv, idesc, acc);   // small terms first
                                acc = 1;
                                umma_tf32_pair(tmem_d, dAh + adv, dBl + adv, idesc, acc);
                                umma_tf32_pair(tmem_d, dAh + adv, dBh + adv, idesc, acc);
                            } else {
                                umma_tf32(tmem_d, dAl + adv, dBh + adv, idesc, acc);
                                acc = 1;
                                umma_tf32(tmem_d, dAh + adv, dBl + adv, idesc, acc);
                                umma_tf32(tmem_d, dAh + adv, dBh + adv, idesc, acc);
                            }
                        }
                        // frees the stage (in both CTAs) once these MMAs have read it
                        if constexpr (kCtas == 2) umma_commit_pair(&empty[s]); else umma_commit(&empty[s]);
                    }
                    // chunk accumulator complete (in both CTAs' tensor memory)
                    if constexpr (kCtas == 2) umma_commit_pair(&tmem_full[buf]); else umma_commit(&tmem_full[buf]);
                }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;                   // TMEM lane quarter this warp may access
        const int half = (warp - 4) >> 2;         // which 64 of the 128 columns
        uint32_t cc = 0;
        for (int64_t t = t_first; t < n_tiles; t += t_step) {
            int m0, n0;
            tile_origin(t, m0, n0);
            const int64_t row = (int64_t)m0 + q * 32 + lane;
            float sum[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) sum[j] = 0.f;
            for (int c = 0; c < nchunk; ++c, ++cc) {
                const int buf = cc % kAccBufs;
                mbar_wait(&tmem_full[buf], (cc / kAccBufs) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // both 32-column loads go out before the single wait, and the buffer is handed back to the MMA
                // warp as soon as the values are in registers (the additions overlap the next chunk's MMAs)
                uint32_t v[64];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * 64 + h * 32);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[h * 32 + 0]), "=r"(v[h * 32 + 1]), "=r"(v[h * 32 + 2]), "=r"(v[h * 32 + 3]),
                          "=r"(v[h * 32 + 4]), "=r"(v[h * 32 + 5]), "=r"(v[h * 32 + 6]), "=r"(v[h * 32 + 7]),
                          "=r"(v[h * 32 + 8]), "=r"(v[h * 32 + 9]), "=r"(v[h * 32 + 10]), "=r"(v[h * 32 + 11]),
                          "=r"(v[h * 32 + 12]), "=r"(v[h * 32 + 13]), "=r"(v[h * 32 + 14]), "=r"(v[h * 32 + 15]),
                          "=r"(v[h * 32 + 16]), "=r"(v[h * 32 + 17]), "=r"(v[h * 32 + 18]), "=r"(v[h * 32 + 19]),
                          "=r"(v[h * 32 + 20]), "=r"(v[h * 32 + 21]), "=r"(v[h * 32 + 22]), "=r"(v[h * 32 + 23]),
                          "=r"(v[h * 32 + 24]), "=r"(v[h * 32 + 25]), "=r"(v[h * 32 + 26]), "=r"(v[h * 32 + 27]),
                          "=r"(v[h * 32 + 28]), "=r"(v[h * 32 + 29]), "=r"(v[h * 32 + 30]), "=r"(v[h * 32 + 31])
                        : "r"(taddr));
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // hand the buffer back to the MMA warp
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if constexpr (kCtas == 2) mbar_arrive_rank0(&tmem_empty[buf]);
                    else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
                }
#pragma unroll
                for (int j = 0; j < 64; ++j) sum[j] += __uint_as_float(v[j]);
            }
            if (tma_store) {
                // tile -> global through shared memory and TMA: each column half (4 warps) owns one 16 KB
                // staging buffer laid out as the 128-byte-swizzled box [128 rows][32 columns]
                unsigned char* buf = slab + half * kSlabBytes;
                const int r = q * 32 + lane;
                const bool leader = (q == 0 && lane == 0);
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    const int col = n0 + (half * 2 + sl) * kSlabCols;
                    if (col >= N) break;                                   // uniform
                    if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // buffer no longer being read
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int j = sl * 32 + c * 4;
                        *reinterpret_cast<float4*>(buf + r * 128 + ((c ^ (r & 7)) << 4)) =
                            make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
                    if (leader) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(&tmC), "r"(smem_u32(buf)), "r"(col), "r"(m0) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            } else if (row < M) {
                const int64_t col0 = (int64_t)n0 + half * 64;
                float* out = C + row * ldc + col0;
#pragma unroll
                for (int j = 0; j < 64; ++j)
                    if (col0 + j < N) out[j] = sum[j];
            }
        }
        if (tma_store && q == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (kCtas == 2) cluster_sync();      // neither CTA leaves while the other may still read its memory
    if (warp == 2) {
        if constexpr (kCtas == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D K-major operand [rows, K] with row stride ld (floats): box = 32 floats x 128 rows, 128 B swizzle
static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t K, int64_t ld, int box_k = BK,
                    int box_rows = BM) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return SO_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SO_OK : SO_ERR_BAD_ARG;
}

}  // namespace

extern "C" int so_gemm_3xtf32(const float* A_hi, const float* A_lo, int64_t lda,
                              const float* B_hi, const float* B_lo, int64_t ldb,
                              float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, void* stream) {
    if (M < 1 || N < 1 || K < 1 || !A_hi || !A_lo || !B_hi || !B_lo || !C) return SO_ERR_BAD_ARG;
    if ((lda & 3) || (ldb & 3) || lda < K || ldb < K || ldc < N) return SO_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(A_hi) | reinterpret_cast<uintptr_t>(A_lo) | reinterpret_cast<uintptr_t>(B_hi) |
         reinterpret_cast<uintptr_t>(B_lo)) & 15)
        return SO_ERR_BAD_ARG;
    // SO_GEMM_PAIR=1: a CTA pair (cta_group::2) per 256x128 tile when there are at least two 128-row tiles.
    // Correct and as fast as single CTAs (2.17 vs 2.16 ms), not faster: the kernel is bound by the per-chunk
    // drain of the accumulator (tcgen05.ld blocks the tensor pipe for ~500 cycles per chunk), not by shared memory.
    static int allow_pair = -1;
    if (allow_pair < 0) { const char* e = getenv("SO_GEMM_PAIR"); allow_pair = (e && e[0] == '1') ? 1 : 0; }
    const bool pair = allow_pair && M > BM;
    const int kc = pair ? 2 : 1;
    CUtensorMap mAh, mAl, mBh, mBl;
    int rc;
    if ((rc = make_map(&mAh, A_hi, M, K, lda)) != SO_OK) return rc;
    if ((rc = make_map(&mAl, A_lo, M, K, lda)) != SO_OK) return rc;
    if ((rc = make_map(&mBh, B_hi, N, K, ldb, BK, BN / kc)) != SO_OK) return rc;
    if ((rc = make_map(&mBl, B_lo, N, K, ldb, BK, BN / kc)) != SO_OK) return rc;
    // the epilogue stores C with TMA when its rows are 16-byte aligned; plain stores otherwise
    CUtensorMap mC = mAh;
    const int tma_store = ((ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) ? 1 : 0;
    if (tma_store && (rc = make_map(&mC, C, M, N, ldc, kSlabCols)) != SO_OK) return rc;
    const int64_t tiles_m = (M + BM * kc - 1) / (BM * kc), tiles_n = (N + BN - 1) / BN;
    if (tiles_m > (1 << 24) || tiles_n > (1 << 24)) return SO_ERR_TOO_LARGE;
    int64_t groups = tiles_m * tiles_n;                       // CTAs (kc = 1) or CTA pairs (kc = 2)
    if (groups > so_num_sms() / kc) groups = so_num_sms() / kc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(groups * kc));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kc;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SO_PROF("gemm_3xtf32_kernel", cfg.stream);
    if (pair) {
        cfg.dynamicSmemBytes = smem_bytes<2>();
        SO_CUDA(cudaFuncSetAttribute(gemm_3xtf32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<2>()));
        SO_CUDA(cudaLaunchKernelEx(&cfg, gemm_3xtf32_kernel<2>, mAh, mAl, mBh, mBl, mC, C, ldc, M, N, K, (int)tiles_m,
                                   (int)tiles_n, tma_store));
    } else {
        cfg.dynamicSmemBytes = smem_bytes<1>();
        SO_CUDA(cudaFuncSetAttribute(gemm_3xtf32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<1>()));
        SO_CUDA(cudaLaunchKernelEx(&cfg, gemm_3xtf32_kernel<1>, mAh, mAl, mBh, mBl, mC, C, ldc, M, N, K, (int)tiles_m,
                                   (int)tiles_n, tma_store));
    }
    return SO_OK;
}

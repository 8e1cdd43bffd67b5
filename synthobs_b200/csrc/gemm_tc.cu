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
// Structure (one CTA per 128x128 output tile, 384 threads):
//   warp 0   TMA producer: per K-block of 32 floats (=128 B, one SWIZZLE_128B atom row) four
//            cp.async.bulk.tensor loads (Ahi, Alo, Bhi, Blo tiles, 16 KB each) into a 3-stage ring
//   warp 1   MMA issuer: one elected thread issues 4 k-steps x 3 tcgen05.mma.kind::tf32 (M=128,N=128,
//            K=8) per stage, commits the stage back to the producer and finally the accumulator
//   warp 2   TMEM allocator (2 x 128 columns)
//   warps 4-11 epilogue/drain: warp w owns TMEM lanes 32*(w%4).. and column half (w-4)/4;
//            tcgen05.ld 32 lanes x 32 columns -> registers (+= running sums) -> global stores
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int kStages = 3;
constexpr int kTileBytes = BM * BK * 4;            // 16 KB
constexpr int kStageBytes = 4 * kTileBytes;        // Ahi, Alo, Bhi, Blo
constexpr int kGemmThreads = 384;
constexpr int kTmemCols = 256;                 // two 128-column accumulators
constexpr int kChunk = 2;                      // K-blocks per TMEM accumulation chunk
constexpr int kDrainWarps = 8;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment*/ + 256 /*barriers*/;

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

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_3xtf32_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                   float* __restrict__ C, int64_t ldc, int64_t M, int64_t N, int64_t K) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* full = bars;                 // [kStages]
    uint64_t* empty = bars + kStages;      // [kStages]
    uint64_t* tmem_full = bars + 2 * kStages;          // [2]
    uint64_t* tmem_empty = bars + 2 * kStages + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nkb = (int)((K + BK - 1) / BK);
    const int nchunk = (nkb + kChunk - 1) / kChunk;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], kDrainWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                unsigned char* st = smem + s * kStageBytes;
                mbar_expect_tx(&full[s], kStageBytes);
                tma_load_2d(st + 0 * kTileBytes, &tmAh, &full[s], kb * BK, m0);
                tma_load_2d(st + 1 * kTileBytes, &tmAl, &full[s], kb * BK, m0);
                tma_load_2d(st + 2 * kTileBytes, &tmBh, &full[s], kb * BK, n0);
                tma_load_2d(st + 3 * kTileBytes, &tmBl, &full[s], kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, K-major, N>>3, M>>4
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int c = 0; c < nchunk; ++c) {
                const int buf = c & 1;
                mbar_wait(&tmem_empty[buf], ((c >> 1) & 1) ^ 1);      // drained by the epilogue warps
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
                uint32_t acc = 0;
                const int kb_end = min(nkb, (c + 1) * kChunk);
                for (int kb = c * kChunk; kb < kb_end; ++kb) {
                    const int s = kb % kStages;
                    const uint32_t ph = (kb / kStages) & 1;
                    mbar_wait(&full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    unsigned char* st = smem + s * kStageBytes;
                    const uint64_t dAh = umma_desc(st + 0 * kTileBytes), dAl = umma_desc(st + 1 * kTileBytes);
                    const uint64_t dBh = umma_desc(st + 2 * kTileBytes), dBl = umma_desc(st + 3 * kTileBytes);
#pragma unroll
                    for (int k4 = 0; k4 < BK / 8; ++k4) {
                        const uint64_t adv = (uint64_t)(k4 * 2);     // 8 tf32 = 32 B = 2 x 16 B along K inside the swizzle atom
                        umma_tf32(tmem_d, dAl + adv, dBh + adv, idesc, acc);   // small terms first
                        acc = 1;
                        umma_tf32(tmem_d, dAh + adv, dBl + adv, idesc, acc);
                        umma_tf32(tmem_d, dAh + adv, dBh + adv, idesc, acc);
                    }
                    umma_commit(&empty[s]);       // frees the stage once these MMAs have read it
                }
                umma_commit(&tmem_full[buf]);     // chunk accumulator complete
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;                   // TMEM lane quarter this warp may access
        const int half = (warp - 4) >> 2;         // which 64 of the 128 columns
        const int64_t row = (int64_t)m0 + q * 32 + lane;
        float sum[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) sum[j] = 0.f;
        for (int c = 0; c < nchunk; ++c) {
            const int buf = c & 1;
            mbar_wait(&tmem_full[buf], (c >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * 64 + h * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) sum[h * 32 + j] += __uint_as_float(v[j]);
            }
            // hand the buffer back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
        }
        if (row < M) {
            const int64_t col0 = (int64_t)n0 + half * 64;
            float* out = C + row * ldc + col0;
            if (col0 + 64 <= N && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 64; j += 4)
                    *reinterpret_cast<float4*>(out + j) = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 64; ++j)
                    if (col0 + j < N) out[j] = sum[j];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
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
static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t K, int64_t ld) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return SO_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
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
    CUtensorMap mAh, mAl, mBh, mBl;
    int rc;
    if ((rc = make_map(&mAh, A_hi, M, K, lda)) != SO_OK) return rc;
    if ((rc = make_map(&mAl, A_lo, M, K, lda)) != SO_OK) return rc;
    if ((rc = make_map(&mBh, B_hi, N, K, ldb)) != SO_OK) return rc;
    if ((rc = make_map(&mBl, B_lo, N, K, ldb)) != SO_OK) return rc;
    SO_CUDA(cudaFuncSetAttribute(gemm_3xtf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM));
    gemm_3xtf32_kernel<<<grid, kGemmThreads, kSmemBytes, (cudaStream_t)stream>>>(mAh, mAl, mBh, mBl, C, ldc, M, N, K);
    SO_LAUNCH_CHECK();
    return SO_OK;
}

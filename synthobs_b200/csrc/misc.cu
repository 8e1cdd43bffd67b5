// Library identification, device probe and the FP32 (FMA-pipe) GEMM used to validate the
// tensor-core kernel on the device.
#include "common.cuh"

namespace {

// C[M,N] = A[M,K] * B[N,K]^T, 64x64 tile, 256 threads, 4x4 per thread, K step 16
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int64_t lda,
                                                       const float* __restrict__ B, int64_t ldb,
                                                       float* __restrict__ C, int64_t ldc,
                                                       int64_t M, int64_t N, int64_t K) {
    __shared__ float sA[16][64 + 4];
    __shared__ float sB[16][64 + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.y * 64, n0 = (int64_t)blockIdx.x * 64;
    const int ty = tid / 16, tx = tid % 16;
    float acc[4][4] = {};
    for (int64_t k0 = 0; k0 < K; k0 += 16) {
        for (int e = tid; e < 64 * 16; e += 256) {
            const int r = e / 16, kk = e % 16;
            const int64_t m = m0 + r, n = n0 + r, k = k0 + kk;
            sA[kk][r] = (m < M && k < K) ? A[m * lda + k] : 0.f;
            sB[kk][r] = (n < N && k < K) ? B[n * ldb + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const int64_t m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) C[m * ldc + n] = acc[i][j];
        }
}

}  // namespace

extern "C" const char* so_version(void) { return "synthobs_b200 0.1 (sm_100a, " __DATE__ ")"; }

extern "C" int so_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) return 0;
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    return major == 10 ? 1 : 0;
}

extern "C" int so_gemm_f32(const float* A, int64_t lda, const float* B, int64_t ldb,
                           float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, void* stream) {
    if (M < 1 || N < 1 || K < 1 || !A || !B || !C) return SO_ERR_BAD_ARG;
    dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64));
    gemm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, B, ldb, C, ldc, M, N, K);
    SO_LAUNCH_CHECK();
    return SO_OK;
}

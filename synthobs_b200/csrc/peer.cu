// Sum of the ranks' partial images of ONE huge image over NVLink peer memory (BASELINE config 5; the reference has no
// multi-GPU path: synthobs/morph/images.py:83-97 images one object in one loop).
//
// A plain all-reduce moves the whole image 2 (N-1)/N times per rank although every partial image is spatially compact
// (a rank deposits one share of the particles) and nearly disjoint from the others.  Here the image is cut into
// 32 x 32 pixel exchange blocks; every rank publishes which blocks its partial image touches (one byte per block),
// and after one all-gather of those masks two kernels finish the job over peer-mapped memory (cudaIpc, NVLink/NVSwitch
// loads):
//   so_peer_reduce   the OWNER of a block -- the lowest rank that touched it, so its own contribution is local --
//                    adds the touching ranks' partial blocks in rank order (deterministic) into its result image;
//   so_peer_gather   every rank copies the blocks it does not own from their owners' result images (untouched blocks
//                    are zeroed locally, nothing is transferred for them).
// Only overlapping blocks cross the links in the first step and every block crosses at most once per reader in the
// second.  The ordering between the steps is the caller's (a tiny collective between them; synthobs_b200/dist.py).
#include <string.h>

#include "common.cuh"

namespace {

constexpr int kXB = 32;              // exchange block edge in pixels (rows of 128 bytes)
constexpr int kMaxRanks = 16;

struct PeerPtrs {
    const float* p[kMaxRanks];
};

__global__ void __launch_bounds__(256) block_mask_kernel(const float* __restrict__ img, int n_planes, int ndim, int64_t ld, int64_t plane,
                                                         int nbx, unsigned char* __restrict__ mask) {
    const int b = blockIdx.x;
    const int y0 = (b / nbx) * kXB, x0 = (b % nbx) * kXB;
    bool any = false;
    if (((ld | plane | ndim) & 3) == 0) {
        const int y = y0 + threadIdx.x / 8, x = x0 + (threadIdx.x % 8) * 4;
        if (y < ndim && x < ndim)
            for (int c = 0; c < n_planes; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(img + (size_t)c * plane + (size_t)y * ld + x);
                any |= (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f) | (v.w != 0.f);
            }
    } else {
        for (int c = 0; c < n_planes; ++c)
            for (int e = threadIdx.x; e < kXB * kXB; e += 256) {
                const int y = y0 + e / kXB, x = x0 + e % kXB;
                if (y < ndim && x < ndim) any |= img[(size_t)c * plane + (size_t)y * ld + x] != 0.f;
            }
    }
    const int r = __syncthreads_or(any ? 1 : 0);
    if (threadIdx.x == 0) mask[b] = r ? 1 : 0;
}

__device__ __forceinline__ int block_owner(const unsigned char* __restrict__ masks, int n_ranks, int64_t n_blocks, int b) {
    for (int r = 0; r < n_ranks; ++r)
        if (masks[(size_t)r * n_blocks + b]) return r;
    return -1;
}

// one 16-byte load per thread and plane (a 32 x 32 block = 256 float4): peer loads need many bytes in flight
template <bool VEC>
__global__ void __launch_bounds__(256) peer_reduce_kernel(PeerPtrs part, const unsigned char* __restrict__ masks, int n_ranks, int rank,
                                                          int n_planes, int ndim, int nbx, int64_t n_blocks, float* __restrict__ out) {
    const int b = blockIdx.x;
    if (block_owner(masks, n_ranks, n_blocks, b) != rank) return;
    const int y0 = (b / nbx) * kXB, x0 = (b % nbx) * kXB;
    const size_t plane = (size_t)ndim * ndim;
    unsigned touch = 0;
    for (int r = 0; r < n_ranks; ++r) touch |= (masks[(size_t)r * n_blocks + b] ? 1u : 0u) << r;
    if (VEC) {
        const int y = y0 + threadIdx.x / 8, x = x0 + (threadIdx.x % 8) * 4;
        if (y >= ndim || x >= ndim) return;
        for (int c = 0; c < n_planes; ++c) {
            const size_t o = (size_t)c * plane + (size_t)y * ndim + x;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < n_ranks; ++r)
                if (touch >> r & 1u) {                                             // rank order: deterministic
                    const float4 v = *reinterpret_cast<const float4*>(part.p[r] + o);
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                }
            *reinterpret_cast<float4*>(out + o) = s;
        }
    } else {
        for (int c = 0; c < n_planes; ++c)
            for (int e = threadIdx.x; e < kXB * kXB; e += 256) {
                const int y = y0 + e / kXB, x = x0 + e % kXB;
                if (y >= ndim || x >= ndim) continue;
                const size_t o = (size_t)c * plane + (size_t)y * ndim + x;
                float s = 0.f;
                for (int r = 0; r < n_ranks; ++r)
                    if (touch >> r & 1u) s += part.p[r][o];
                out[o] = s;
            }
    }
}

template <bool VEC>
__global__ void __launch_bounds__(256) peer_gather_kernel(PeerPtrs fin, const unsigned char* __restrict__ masks, int n_ranks, int rank,
                                                          int n_planes, int ndim, int nbx, int64_t n_blocks, float* __restrict__ out) {
    // every rank starts its sweep at a different block: otherwise all readers would pull from the same owner at the
    // same time (the blocks of an owner are clustered) and its outgoing links would serialise the whole exchange
    const int b = (int)(((int64_t)blockIdx.x + (int64_t)rank * (n_blocks / n_ranks)) % n_blocks);
    const int owner = block_owner(masks, n_ranks, n_blocks, b);
    if (owner == rank) return;
    const int y0 = (b / nbx) * kXB, x0 = (b % nbx) * kXB;
    const size_t plane = (size_t)ndim * ndim;
    const float* src = owner >= 0 ? fin.p[owner] : nullptr;
    if (VEC) {
        const int y = y0 + threadIdx.x / 8, x = x0 + (threadIdx.x % 8) * 4;
        if (y >= ndim || x >= ndim) return;
        float4 v[4];
        for (int c0 = 0; c0 < n_planes; c0 += 4) {                                  // up to four planes' loads in flight
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + j < n_planes)
                    v[j] = src ? *reinterpret_cast<const float4*>(src + (size_t)(c0 + j) * plane + (size_t)y * ndim + x)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + j < n_planes) *reinterpret_cast<float4*>(out + (size_t)(c0 + j) * plane + (size_t)y * ndim + x) = v[j];
        }
    } else {
        for (int c = 0; c < n_planes; ++c)
            for (int e = threadIdx.x; e < kXB * kXB; e += 256) {
                const int y = y0 + e / kXB, x = x0 + e % kXB;
                if (y >= ndim || x >= ndim) continue;
                const size_t o = (size_t)c * plane + (size_t)y * ndim + x;
                out[o] = src ? src[o] : 0.f;
            }
    }
}

}  // namespace

extern "C" int32_t so_peer_block_edge(void) { return kXB; }

extern "C" int so_peer_alloc(int64_t bytes, void** ptr) {
    if (bytes < 1 || !ptr) return SO_ERR_BAD_ARG;
    SO_CUDA(cudaMalloc(ptr, (size_t)bytes));
    return SO_OK;
}

extern "C" int so_peer_free(void* ptr) {
    if (ptr) SO_CUDA(cudaFree(ptr));
    return SO_OK;
}

extern "C" int so_ipc_get_handle(const void* ptr, void* handle64) {
    if (!ptr || !handle64) return SO_ERR_BAD_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    SO_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
    memcpy(handle64, &h, sizeof h);
    return SO_OK;
}

extern "C" int so_ipc_open(const void* handle64, void** ptr) {
    if (!handle64 || !ptr) return SO_ERR_BAD_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    SO_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SO_OK;
}

extern "C" int so_ipc_close(void* ptr) {
    if (ptr) SO_CUDA(cudaIpcCloseMemHandle(ptr));
    return SO_OK;
}

extern "C" int so_peer_block_mask(const float* img, int32_t n_planes, int32_t ndim, int64_t ld, int64_t plane, uint8_t* mask,
                                  void* stream) {
    if (!img || !mask || n_planes < 1 || ndim < 1 || ld < ndim || plane < ld * ndim) return SO_ERR_BAD_ARG;
    const int nbx = (ndim + kXB - 1) / kXB;
    { SO_PROF("block_mask_kernel", (cudaStream_t)stream);
      block_mask_kernel<<<(unsigned)(nbx * nbx), 256, 0, (cudaStream_t)stream>>>(img, n_planes, ndim, ld, plane, nbx, mask); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

static int peer_args(const void* const* ptrs, int32_t n_ranks, int32_t rank, int32_t n_planes, int32_t ndim, PeerPtrs* pp) {
    if (!ptrs || n_ranks < 1 || n_ranks > kMaxRanks || rank < 0 || rank >= n_ranks || n_planes < 1 || ndim < 1) return SO_ERR_BAD_ARG;
    for (int r = 0; r < n_ranks; ++r) {
        if (!ptrs[r]) return SO_ERR_BAD_ARG;
        pp->p[r] = (const float*)ptrs[r];
    }
    return SO_OK;
}

extern "C" int so_peer_reduce(const void* const* partial_ptrs, const uint8_t* masks, int32_t n_ranks, int32_t rank,
                              int32_t n_planes, int32_t ndim, float* out, void* stream) {
    PeerPtrs pp{};
    int rc = peer_args(partial_ptrs, n_ranks, rank, n_planes, ndim, &pp);
    if (rc != SO_OK || !masks || !out) return SO_ERR_BAD_ARG;
    const int nbx = (ndim + kXB - 1) / kXB;
    { SO_PROF("peer_reduce_kernel", (cudaStream_t)stream);
      if ((ndim & 3) == 0)
          peer_reduce_kernel<true><<<(unsigned)(nbx * nbx), 256, 0, (cudaStream_t)stream>>>(pp, masks, n_ranks, rank, n_planes, ndim, nbx,
                                                                                           (int64_t)nbx * nbx, out);
      else
          peer_reduce_kernel<false><<<(unsigned)(nbx * nbx), 256, 0, (cudaStream_t)stream>>>(pp, masks, n_ranks, rank, n_planes, ndim, nbx,
                                                                                            (int64_t)nbx * nbx, out); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

extern "C" int so_peer_gather(const void* const* final_ptrs, const uint8_t* masks, int32_t n_ranks, int32_t rank,
                              int32_t n_planes, int32_t ndim, float* out, void* stream) {
    PeerPtrs pp{};
    int rc = peer_args(final_ptrs, n_ranks, rank, n_planes, ndim, &pp);
    if (rc != SO_OK || !masks || !out) return SO_ERR_BAD_ARG;
    const int nbx = (ndim + kXB - 1) / kXB;
    { SO_PROF("peer_gather_kernel", (cudaStream_t)stream);
      if ((ndim & 3) == 0)
          peer_gather_kernel<true><<<(unsigned)(nbx * nbx), 256, 0, (cudaStream_t)stream>>>(pp, masks, n_ranks, rank, n_planes, ndim, nbx,
                                                                                           (int64_t)nbx * nbx, out);
      else
          peer_gather_kernel<false><<<(unsigned)(nbx * nbx), 256, 0, (cudaStream_t)stream>>>(pp, masks, n_ranks, rank, n_planes, ndim, nbx,
                                                                                            (int64_t)nbx * nbx, out); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

// Sum of the ranks' partial images of ONE huge image over NVLink peer memory (BASELINE config 5; the reference has no
// multi-GPU path: synthobs/morph/images.py:83-97 images one object in one loop).
//
// A plain all-reduce moves the whole image 2 (N-1)/N times per rank although every partial image is spatially compact
// (a rank deposits one share of the particles) and nearly disjoint from the others.  Here the image is cut into
// 32 x 32 pixel exchange blocks; every rank publishes which blocks its partial image touches (one byte per block),
// and after one all-gather of those masks ONE kernel finishes the job over peer-mapped memory (cudaIpc, NVLink/NVSwitch
// loads): so_peer_sum adds, for every block, the partial blocks of the ranks that touched it, in rank order
// (deterministic and identical on every rank); untouched blocks are zeroed locally.  A block crosses the links once per
// reader and touching rank -- (N-1)/N of the image plus the few overlapping blocks, against 2 (N-1)/N for the all-reduce
// -- and the all-gather of the masks is the only synchronisation (it also marks the point after which every rank's
// partial image is complete); the caller alternates two partial buffers so that nobody overwrites what a slower
// peer may still be reading (synthobs_b200/dist.py).
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

// out = sum over the ranks whose mask has the block, in rank order (deterministic, the same on every rank); blocks
// nobody touched are zeroed locally.  One 16-byte load per thread, plane and touching rank (a 32 x 32 block = 256
// float4): peer loads need many bytes in flight.  Every rank starts its sweep at a different block: otherwise all
// readers would pull from the same rank at the same time (a rank's blocks are clustered) and its outgoing links would
// serialise the exchange.
template <bool VEC>
__global__ void __launch_bounds__(256) peer_sum_kernel(PeerPtrs part, const unsigned char* __restrict__ masks, int n_ranks, int rank,
                                                       int n_planes, int ndim, int nbx, int64_t n_blocks, float* __restrict__ out) {
    const int b = (int)(((int64_t)blockIdx.x + (int64_t)rank * (n_blocks / n_ranks)) % n_blocks);
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
            for (unsigned t = touch; t; t &= t - 1) {
                const float4 v = *reinterpret_cast<const float4*>(part.p[__ffs(t) - 1] + o);
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
                for (unsigned t = touch; t; t &= t - 1) s += part.p[__ffs(t) - 1][o];
                out[o] = s;
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

extern "C" int so_peer_sum(const void* const* partial_ptrs, const uint8_t* masks, int32_t n_ranks, int32_t rank,
                           int32_t n_planes, int32_t ndim, float* out, void* stream) {
    PeerPtrs pp{};
    int rc = peer_args(partial_ptrs, n_ranks, rank, n_planes, ndim, &pp);
    if (rc != SO_OK || !masks || !out) return SO_ERR_BAD_ARG;
    const int nbx = (ndim + kXB - 1) / kXB;
    { SO_PROF("peer_sum_kernel", (cudaStream_t)stream);
      if ((ndim & 3) == 0)
          peer_sum_kernel<true><<<(unsigned)(nbx * nbx), 256, 0, (cudaStream_t)stream>>>(pp, masks, n_ranks, rank, n_planes, ndim, nbx,
                                                                                        (int64_t)nbx * nbx, out);
      else
          peer_sum_kernel<false><<<(unsigned)(nbx * nbx), 256, 0, (cudaStream_t)stream>>>(pp, masks, n_ranks, rank, n_planes, ndim, nbx,
                                                                                         (int64_t)nbx * nbx, out); }
    SO_LAUNCH_CHECK();
    return SO_OK;
}

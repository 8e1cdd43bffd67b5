// Entry points declared in include/synthobs_b200.h whose kernels land in later commits of this
// round; they fail loudly (SO_ERR_UNSUPPORTED) rather than fall back to anything.
#include "common.cuh"

extern "C" int so_gemm_3xtf32(const float*, const float*, int64_t, const float*, const float*, int64_t,
                              float*, int64_t, int64_t, int64_t, int64_t, void*) { return SO_ERR_UNSUPPORTED; }

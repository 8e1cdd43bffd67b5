// Entry points declared in include/synthobs_b200.h whose kernels land in later commits of this
// round; they fail loudly (SO_ERR_UNSUPPORTED) rather than fall back to anything.
#include "common.cuh"

extern "C" int so_gemm_3xtf32(const float*, const float*, int64_t, const float*, const float*, int64_t,
                              float*, int64_t, int64_t, int64_t, int64_t, void*) { return SO_ERR_UNSUPPORTED; }
extern "C" int64_t so_sed_dust_spectra_workspace_bytes(int64_t, int32_t) { return 0; }
extern "C" int so_sed_dust_spectra(const double*, const double*, const double*, const int32_t*, const int32_t*, int64_t,
                                   const float*, const float*, int32_t, int32_t, const float*, const float*, double,
                                   float*, float*, float*, void*, int64_t, void*) { return SO_ERR_UNSUPPORTED; }

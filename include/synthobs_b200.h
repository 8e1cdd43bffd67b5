/* synthobs_b200 — C-ABI of the B200 (sm_100a) hot paths of SynthObs.
 *
 * The reference (stephenmwilkins/SynthObs) is pure Python/NumPy and has no FFI of
 * its own: the drop-in boundary is the Python call surface of
 * synthobs.sed.models / synthobs.morph.images.  This header is the boundary one
 * level below it: every entry point replaces one Python hot loop of the reference
 * (cited per function, paths relative to the reference root) and is what a
 * ctypes/cffi binding in the reference would bind (INTEGRATION.md shows the stub).
 *
 * Conventions
 *  - plain C: raw pointers, int64 sizes, no torch / C++ types.
 *  - pointers are DEVICE pointers unless the parameter is documented "host".
 *  - `stream` is a cudaStream_t passed as void*; every call is stream-ordered and
 *    asynchronous unless documented "synchronises".
 *  - the caller allocates every input, output and workspace buffer; nothing is
 *    allocated, retained or freed behind the ABI (except cached TMA descriptors
 *    and kernel attributes, which hold no user memory).
 *  - return value: 0 = ok, >0 = cudaError_t of the failing CUDA call,
 *    <0 = SO_ERR_* argument error.  Nothing throws across the ABI.
 *  - particle inputs are float64, exactly what the reference API receives
 *    (np.load of X, Y, Ages, ..., synthobs/core.py:17-22); index decisions are
 *    made in float64 (bit-exact against np.argmin), values are computed in FP32.
 */
#ifndef SYNTHOBS_B200_H
#define SYNTHOBS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO_OK 0
#define SO_ERR_BAD_ARG (-1)
#define SO_ERR_TOO_LARGE (-2)
#define SO_ERR_WORKSPACE (-3)
#define SO_ERR_UNSUPPORTED (-4)

#define SO_MAX_FILTERS 32 /* filters per so_sed_broadband launch (smem tables) */
#define SO_TILE 16        /* image sub-tile edge in pixels (so_img_*) */
#define SO_MAX_PEERS 16   /* ranks of one node that may share a huge image (so_img_deposit_push) */

/* library / build identification; returns a static string */
const char* so_version(void);
/* 1 if a CUDA device of compute capability 10.x is present, else 0 */
int so_device_ok(void);

/* ------------------------------------------------------------------------- */
/* SED path                                                                   */
/* ------------------------------------------------------------------------- */

/* K1+K3 — nearest-grid-point SSP lookup + per-particle broadband luminosity/flux
 * + per-galaxy sums.  Replaces the particle loops of generate_Lnu_array /
 * generate_Fnu_array (synthobs/sed/models.py:88-135, :153-200) for ALL filters in
 * one pass, and the np.sum of generate_Lnu / generate_Fnu (:74-85, :139-150).
 *
 * NGP index: the reference does ia = argmin|log10age_grid - (log10(Age)+6)|
 * (:108-114).  That decision is monotone in Age, so the host pre-computes, with
 * the SAME NumPy/libm expression, the float64 thresholds age_thr[na-1] (smallest
 * Age that maps to node i+1) and z_thr[nz-1]; the kernel only compares float64
 * values, which reproduces np.argmin bit-exactly without a device log10.
 * young_thr = smallest Age with log10(Age)+6 >= log10t_BC (:124).
 *
 *   masses, ages, metals : [n] float64
 *   tau_ism, tau_bc      : [n] float64 or NULL (= zeros, :92-93)
 *   seg_offsets          : [n_seg+1] int64 CSR offsets of the galaxies, or NULL
 *                          with n_seg = 1 (one galaxy = all particles)
 *   age_thr [na-1], z_thr [nz-1] : HOST float64 thresholds (na, nz <= 128); they
 *                          travel to the kernel as launch parameters together with
 *                          a log2-spaced first-guess table built from them
 *   tab_inc, tab_tn      : [n_filters, na*nz] float32: stellar_incident and
 *                          (stellar_transmitted + nebular) per filter (:131-133)
 *   k_ism, k_bc          : HOST [n_filters] float64 dust-curve value at the filter
 *                          pivot wavelength (:95-101), or NULL when the model has
 *                          no such dust (T = 1)
 *   out_pf               : [n_filters, n] float32 per-particle values or NULL
 *   out_sums             : [n_seg, n_filters] float64 per-galaxy sums or NULL
 *   out_cell             : [n] int32 = ia*nz + iZ + (young ? 65536 : 0) or NULL
 *   workspace            : so_sed_broadband_workspace_bytes(n, n_seg, n_filters)
 */
int64_t so_sed_broadband_workspace_bytes(int64_t n, int64_t n_seg, int32_t n_filters);
int so_sed_broadband(const double* masses, const double* ages, const double* metals,
                     const double* tau_ism, const double* tau_bc, int64_t n,
                     const int64_t* seg_offsets, int64_t n_seg,
                     const double* age_thr, int32_t na, const double* z_thr, int32_t nz, double young_thr,
                     const float* tab_inc, const float* tab_tn, int32_t n_filters,
                     const double* k_ism, const double* k_bc, double fesc,
                     float* out_pf, double* out_sums, int32_t* out_cell,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* K2 — per-galaxy grid-weight histogram H[n_seg, 2*na*nz] (old cells then young
 * cells), H = sum of particle masses per SSP cell: the left operand of the
 * dust-free contractions of generate_SED (synthobs/sed/models.py:277,281,295,
 * 306-307) and of generate_log10Q (:206-221).  `cell` is so_sed_broadband's
 * out_cell.  ld_h >= 2*na*nz is the row stride of H in floats.  H_lo (optional)
 * receives the TF32 residual of H for the 3xTF32 GEMM. */
int64_t so_sed_hist_workspace_bytes(int64_t n_seg, int32_t ncell);
int so_sed_hist(const double* masses, const int32_t* cell, int64_t n,
                const int64_t* seg_offsets, int64_t n_seg, int32_t ncell,
                float* H, float* H_lo, int64_t ld_h,
                void* workspace, int64_t workspace_bytes, void* stream);

/* split x into TF32-exact hi and lo = x - hi (elementwise), for the 3xTF32 GEMM */
int so_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);

/* K4/K5 — C[M,N] = A[M,K] * B[N,K]^T in error-compensated 3xTF32 on the tcgen05
 * tensor cores (FP32 accumulate in TMEM): A*B = Ahi*Bhi + Ahi*Blo + Alo*Bhi.
 * Replaces the dense contractions  hist x SSP-grid  (generate_SED dust-free
 * spectra, synthobs/sed/models.py:277-310),  grid x filter-weights
 * (create_Lnu_grid / create_Fnu_grid, :39-67) and  spectra x filter-weights.
 * All operands are K-major (row-major with K contiguous), float32, 16-byte
 * aligned; lda/ldb/ldc are row strides in elements, multiples of 4.
 * M, N, K arbitrary (TMA zero-fills out-of-range tiles). */
int so_gemm_3xtf32(const float* A_hi, const float* A_lo, int64_t lda,
                   const float* B_hi, const float* B_lo, int64_t ldb,
                   float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, void* stream);
/* same product on the FP32 FMA pipe (one operand set, no split); used to validate
 * the tensor-core kernel on the device and for very small problems */
int so_gemm_f32(const float* A, int64_t lda, const float* B, int64_t ldb,
                float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, void* stream);

/* K6 — dust-attenuated spectra of generate_SED (synthobs/sed/models.py:285-310):
 * BC, total and no_HII_no_BC need T_i(lambda) = exp(-tauV_i * k(lambda)) per
 * particle and wavelength, which does not factor through the histogram.
 * Particles are given in an order sorted by (galaxy, young, cell): `order` =
 * sorted position -> particle index, `cell_sorted` = the cell codes
 * (so_sed_broadband's out_cell) in that order, `seg_offsets` = CSR offsets of the
 * galaxies in the sorted order (NULL with n_seg = 1 for one galaxy).
 * grid_inc / grid_tn: [ncell, n_lam] float32 rows (stellar_incident,
 * transmitted+nebular).  k_ism_lam / k_bc_lam: [n_lam] float32 or NULL (T = 1).
 * Outputs [n_seg, n_lam] float32 each (NULL = not wanted). */
int64_t so_sed_dust_spectra_workspace_bytes(int64_t n, int64_t n_seg, int32_t n_lam);
int so_sed_dust_spectra(const double* masses, const double* tau_ism, const double* tau_bc,
                        const int32_t* cell_sorted, const int32_t* order, int64_t n,
                        const int64_t* seg_offsets, int64_t n_seg,
                        const float* grid_inc, const float* grid_tn, int32_t ncell, int32_t n_lam,
                        const float* k_ism_lam, const float* k_bc_lam, double fesc,
                        float* out_bc, float* out_total, float* out_nohii,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- */
/* Imaging path                                                               */
/* ------------------------------------------------------------------------- */

/* K8a — selection + nearest-pixel index.  Replaces images.core's selection
 * (synthobs/morph/images.py:47-50) and the per-particle argmin of :67.
 * G = the float64 np.linspace pixel centres (uploaded by the host so the device
 * compares against the very same values).  ix/iy = argmin|G - x| (first minimum)
 * or -1 when the particle is outside |x| < half_width & |y| < half_width. */
int so_img_ngp_index(const double* X, const double* Y, int64_t n,
                     const double* G, int32_t ndim, double half_width,
                     int32_t* ix, int32_t* iy, void* stream);

/* K7 — distance to the k-th nearest neighbour in 2-D, counting the particle
 * itself as the first (cKDTree.query(k)[..., -1], synthobs/morph/images.py:83-85).
 * Only particles with ix >= 0 take part; dk = 0 for the others.  dk = +inf when
 * k exceeds the number of selected particles.  Exact, float64: the squared distance
 * is formed with separately rounded products and sum and an IEEE sqrt, like a C loop
 * (equal to cKDTree to <= 4e-16 relative, ties and identical positions included).
 * The particles are sorted by the Morton code of their cell; calls of more than 32768
 * particles walk the implicit quadtree of that order with one thread per query, smaller
 * ones answer 32 consecutive queries per warp (FP32 screening of shared candidate
 * tiles, float64 decision; csrc/knn.cu).  Stream-ordered, no host synchronisation. */
int64_t so_knn_workspace_bytes(int64_t n, int32_t k);
int so_knn_kth_distance(const double* X, const double* Y, const int32_t* ix, int64_t n,
                        int32_t k, double half_width, double* dk,
                        void* workspace, int64_t workspace_bytes, void* stream);
/* The same search over ALL positions, answered only for part `part` of `n_parts` equal shares of
 * the selected particles taken in Morton order (a spatially compact share: the partition used when
 * one huge image is split over GPUs, BASELINE config 5).  dk is written for this share only.
 * order_out (optional): [n] int32, the Morton-sorted permutation (selected particles first);
 * range_out (optional): DEVICE int64[3] = {begin, end} of this share in that order, and the
 * number of selected particles. */
int so_knn_kth_distance_part(const double* X, const double* Y, const int32_t* ix, int64_t n,
                             int32_t k, double half_width, int32_t part, int32_t n_parts,
                             double* dk, int32_t* order_out, int64_t* range_out,
                             void* workspace, int64_t workspace_bytes, void* stream);

/* Batch of images (galaxies) in one call, BASELINE config 4: `img` [n] int32 gives each particle's
 * image; neighbours are searched inside the particle's own image only (the image id is the high part
 * of the sort key, the node tables of the images are concatenated).  Same arithmetic as above. */
int64_t so_knn_batch_workspace_bytes(int64_t n, int64_t n_img, int32_t k);
int so_knn_kth_distance_batch(const double* X, const double* Y, const int32_t* ix, const int32_t* img,
                              int64_t n, int64_t n_img, int32_t k, double half_width, double* dk,
                              void* workspace, int64_t workspace_bytes, void* stream);

/* General form of the three calls above (img NULL / n_img 1: one image; part 0 of 1: all queries) with a
 * `floor`: the reference uses FWHM = max(dk, 0.01) (synthobs/morph/images.py:89), so a particle that already
 * has k neighbours within `floor` needs no exact answer: for it dk is only guaranteed to be an upper bound
 * <= floor (max(dk, floor) is unchanged), and its tree walk is skipped.  floor = 0: exact everywhere.
 * Workspace: so_knn_batch_workspace_bytes(n, n_img, k). */
int so_knn_kth_distance_ex(const double* X, const double* Y, const int32_t* ix, const int32_t* img,
                           int64_t n, int64_t n_img, int32_t k, double half_width, int32_t part, int32_t n_parts,
                           double floor, double* dk, int32_t* order_out, int64_t* range_out,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* Measurement aid (bench.py's useful-FMA fraction): while `counters` (DEVICE uint64[2], zeroed by the caller) is set, every
 * single-GPU deposit adds [0] += (lane, staged particle) visits of its multiply-add loop -- each evaluates one 2 x 4 pixel
 * block for all channels -- and [1] += 32 x the rounds of that loop (the issue slots it occupied).  NULL switches it off
 * (the counting instantiation of the kernel is a separate one; the normal one is unchanged). */
int so_img_deposit_count(unsigned long long* counters);

/* K8b..K9 — kernel-smoothed deposition.  Replaces the adaptive loop of
 * images.core (synthobs/morph/images.py:87-97): per particle a Gaussian of
 * FWHM = max(dk, 0.01), normalised by its SUM OVER THE IMAGE PIXELS, times L.
 * Also produces the NGP products of :66-69 (`simple`, `hist`) when requested.
 *
 * Truncation: the Gaussian is evaluated on the pixel box +-W around the NGP
 * pixel, W = ceil(support_sigmas * sigma / pitch), clipped to the image, and
 * normalised over that same box (flux is conserved exactly as in the
 * reference); support_sigmas <= 0 means full image support (reference
 * semantics; affordable for small images).  A particle whose largest float64
 * term exp(-d^2/2 sigma^2) underflows to 0 is dropped, as the reference does
 * (`if sgauss > 0`, :97).
 *
 * Two phases because the number of (particle, tile) pairs is data dependent:
 *   so_img_plan      computes per-particle records + pair counts; synchronises
 *                    and returns the pair count in *n_pairs (host).
 *   so_img_deposit   bins the pairs per 16x16 sub-tile (stable radix sort: bit-exact,
 *                    order-deterministic lists), accumulates each sub-tile in one
 *                    warp's registers (no global atomics) and writes every pixel once.
 *
 *   L        : [n_chan, n] float64 weights (n_chan images share positions/sigma)
 *   dk       : [n] float64 k-th neighbour distance, or NULL for NGP only
 *   data     : [n_chan, ndim, ndim] float32 smoothed images (row = y) or NULL
 *   simple   : [n_chan, ndim, ndim] float32 NGP images or NULL
 *   hist     : [ndim, ndim] float32 NGP counts or NULL
 *   tile_ids / pair_pids (optional, for the tile-assignment parity test):
 *              [n_pairs] int32 sorted pair list as used by the kernel
 *   stats    : HOST double[4]: pairs, in-support (particle,pixel) deposits
 *              evaluated, selected particles, dropped particles  (or NULL)
 */
int64_t so_img_plan_workspace_bytes(int64_t n);
int so_img_plan(const int32_t* ix, const int32_t* iy, const double* X, const double* Y,
                const double* dk, int64_t n, const double* G, int32_t ndim,
                double support_sigmas, void* plan_workspace, int64_t plan_workspace_bytes,
                int64_t* n_pairs, double* stats, void* stream);
int64_t so_img_deposit_workspace_bytes(int64_t n, int64_t n_pairs, int32_t ndim, int32_t n_chan);
int so_img_deposit(const void* plan_workspace, const double* L, int64_t n, int32_t n_chan,
                   int32_t ndim, int64_t n_pairs,
                   float* data, float* simple, float* hist,
                   int32_t* tile_ids, int32_t* pair_pids,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* The same for a batch of n_img images sharing geometry (all ndim x ndim on the same pixel grid):
 * `img` [n] int32 = image of each particle (NULL when n_img == 1); outputs gain a leading image axis:
 * data / simple [n_img, n_chan, ndim, ndim], hist [n_img, ndim, ndim].  so_img_plan is per particle
 * and is shared. */
int64_t so_img_deposit_batch_workspace_bytes(int64_t n, int64_t n_pairs, int32_t ndim, int32_t n_chan, int32_t n_img);
int so_img_deposit_batch(const void* plan_workspace, const double* L, const int32_t* img, int64_t n,
                         int32_t n_chan, int32_t n_img, int32_t ndim, int64_t n_pairs,
                         float* data, float* simple, float* hist,
                         int32_t* tile_ids, int32_t* pair_pids,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* so_img_deposit_batch writing `data` into a larger buffer: image (i, c) starts at
 * data + (i * n_chan + c) * data_plane and its rows are data_ld floats apart (data_ld >= ndim).  Used to
 * deposit straight into the zero-padded buffer of the PSF transform (no pad copy); the caller zeroes the WHOLE
 * buffer (padding and image windows: sub-tiles no particle reaches are not written); a dense `data` (data_ld == ndim,
 * data_plane == ndim * ndim) is cleared by the call itself. */
int so_img_deposit_batch_ld(const void* plan_workspace, const double* L, const int32_t* img, int64_t n,
                            int32_t n_chan, int32_t n_img, int32_t ndim, int64_t n_pairs,
                            float* data, int64_t data_ld, int64_t data_plane, float* simple, float* hist,
                            int32_t* tile_ids, int32_t* pair_pids,
                            void* workspace, int64_t workspace_bytes, void* stream);

/* K11 — rebin by block SUM (synthobs/morph/images.py:19-22): in [n_img, n*s, n*s]
 * -> out [n_img, n, n].  _ld: the input images are windows of a larger buffer (rows in_ld floats apart,
 * images in_plane floats apart), e.g. the uncropped result of the PSF convolution. */
int so_img_rebin(const float* in, float* out, int32_t n_img, int32_t n_out, int32_t s, void* stream);
int so_img_rebin_ld(const float* in, int64_t in_ld, int64_t in_plane, float* out, int32_t n_img,
                    int32_t n_out, int32_t s, void* stream);

/* elementwise helper for the PSF step (K10): out = a * b on interleaved complex
 * float32 spectra; a, out: [n_img, n_complex]; b: [n_complex] broadcast over the
 * images (b_per_image = 0) or [n_img, n_complex] (b_per_image = 1).  The FFTs
 * themselves are cuFFT through torch.fft (convolve_fft of images.py:78,202). */
int so_complex_mul_bcast(const float* a, const float* b, float* out, int64_t n_complex, int32_t n_img,
                         int32_t b_per_image, void* stream);

/* Analytic source — observed.Sersic (synthobs/morph/images.py:219-238): the Sersic2D
 * profile (amplitude 1; r_eff, n, ellip, theta; centre x0, y0 in the units of G) on the
 * pixel grid G x G (row = y), normalised to unit sum over the image and scaled per
 * channel: out[c] = L[c] * profile / sum(profile), float32 [n_chan, ndim, ndim].
 * bn = gammaincinv(2n, 0.5) is supplied by the host; L is a DEVICE float64 [n_chan].
 * float64 evaluation and fixed-order sum. */
int64_t so_img_sersic_workspace_bytes(int32_t ndim);
int so_img_sersic(const double* G, int32_t ndim, double r_eff, double n, double x0, double y0,
                  double ellip, double theta, double bn, const double* L, int32_t n_chan,
                  float* out, void* workspace, int64_t workspace_bytes, void* stream);

/* np.median of n_arrays float64 arrays of length n (array a starts at x + a*stride), the
 * recentring step of observed.particle (synthobs/morph/images.py:183-184): element n//2 for
 * odd n, mean of the two middle elements for even n, bit-identical to NumPy for finite data.
 * Radix selection on the order-preserving integer image of the doubles, no sort: two
 * histogram passes over 11-bit digits, one pass that copies the keys sharing those 22 bits
 * into the workspace, the remaining 42 bits on that (small) copy.  out: DEVICE double[n_arrays]. */
int64_t so_median_workspace_bytes(int64_t n, int32_t n_arrays);
int so_median_f64(const double* x, int64_t n, int32_t n_arrays, int64_t stride, double* out,
                  void* workspace, int64_t workspace_bytes, void* stream);
/* same, and mid[2a], mid[2a+1] (DEVICE double[2 n_arrays], may be null) receive the two middle elements of
 * array a (both the median element for odd n) */
int so_median_f64_mid(const double* x, int64_t n, int32_t n_arrays, int64_t stride, double* out, double* mid,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* The per-filter recentring of images.particle (synthobs/morph/images.py:277-284 calling :183-187 once per
 * filter): X_f = (X_{f-1} - np.median(X_{f-1})) + off_x[f], likewise Y, for f = 1..n_steps, bit-identical to the
 * reference's repeated in-place updates.  mid = the two middle elements of x and of y (so_median_f64_mid on the
 * two arrays): every later median follows from them by scalar arithmetic, so the arrays are read once.  The
 * positions after step out_step[j] (1-based, ascending) are written to out_x[j], out_y[j] (n_out <= 8).
 * off_x, off_y, out_step, out_x, out_y are HOST arrays (out_x[j], out_y[j] DEVICE pointers). */
int so_recentre_chain(const double* x, const double* y, int64_t n, const double* mid, int32_t n_steps,
                      const double* off_x, const double* off_y, int32_t n_out, const int32_t* out_step,
                      double* const* out_x, double* const* out_y, void* stream);


/* np.median per CSR segment (one object of a batch per segment; out: DEVICE double[n_seg], NaN for
 * an empty segment) and the in-place recentring x[i] += shift - med[segment(i)] of
 * observed.particle (synthobs/morph/images.py:183-187) for every object of the batch. */
int so_median_segmented(const double* x, const int64_t* seg_offsets, int64_t n_seg, double* out, void* stream);
int so_recentre_segmented(double* x, const int64_t* seg_offsets, int64_t n_seg, const double* med, double shift,
                          void* stream);

/* Line-of-sight metal column density of star particles: the input the reference loads as
 * MetSurfaceDensities.npy (synthobs/core.py:17-21, :33-37) and turns into tauVs_ISM = 10^5.2 * MetSurfaceDensities
 * (examples/SED_luminosities.py:69).  The reference does not compute it; specification: oracle/los_oracle.py.
 *   out[i] = scale * sum over the gas particles j of star i's object with zg[j] > zs[i] of
 *            mg[j] Zg[j] / hg[j]^2 * F(b_ij / hg[j]),  b_ij = distance in the (x, y) plane,
 * F = cubic-spline SPH kernel (support r < h) projected along z: `table` = F(i / nt), i = 0..nt (DEVICE float32[nt + 1],
 * nt <= 2048), interpolated linearly, 0 beyond.  Batch of objects: stars and gas concatenated, star_off / gas_off =
 * HOST int64[n_obj + 1] CSR offsets (a star only sees its own object's gas).  All other arrays DEVICE float64.
 * The z test uses the float64 inputs; pairs are screened in FP32 (conservatively) and the in-support ones decided from
 * FP32 head + tail coordinates; the kernel sum is FP32 per tile of 128 gas particles, float64 over the tiles,
 * in a fixed order (bitwise reproducible).  star_off[0] must be 0.  Workspace: so_los_workspace_bytes(n_obj, n_star). */
int64_t so_los_workspace_bytes(int64_t n_obj, int64_t n_star);
int so_los_column(const double* xs, const double* ys, const double* zs, const int64_t* star_off_host,
                  const double* xg, const double* yg, const double* zg, const double* mg, const double* Zg,
                  const double* hg, const int64_t* gas_off_host, int64_t n_obj, const float* table, int32_t nt,
                  double scale, double* out, void* workspace, int64_t workspace_bytes, void* stream);

/* One huge image over several GPUs (BASELINE config 5; the reference images one object in one Python loop,
 * synthobs/morph/images.py:83-97): split the selected particles (|x|, |y| < half_width) into n_parts spatially compact
 * shares of equal WORK (+- one coarse cell; a particle of a cell too sparse to hold k particles within `floor`, the
 * FWHM floor of synthobs/morph/images.py:89, counts 2.5 times: its query walks the tree and its support is wide) --
 * contiguous Morton ranges of a 2^cb x 2^cb grid over the image -- and flag
 * every particle for the share [share_lo, share_hi) of the total work (share p of N equal ones: p / N, (p + 1) / N;
 * unequal fractions let the caller balance measured times): bit 1 (value 2) = OWN (this share answers its k-NN query and deposits it),
 * bit 0 = LOCAL (own or halo: member of this share's search tree).  The halo is sized from the coarse histogram so
 * that the k-th neighbour distance among the LOCAL particles is exact for every OWN particle (share.cu header).
 * so_share_hist counts the selected particles [begin, end) per coarse cell (DEVICE int32[4^cb], Morton order): with
 * several ranks each counts its 1/N of the particles and the tables are summed (ncclAllReduce) before so_share_plan.
 * counts: DEVICE uint64[2] = number of local / own particles.  Identical on every rank given identical positions. */
int so_share_hist(const double* X, const double* Y, int64_t begin, int64_t end, double half_width, int32_t cb,
                  int32_t* hist, void* stream);
int64_t so_share_plan_workspace_bytes(int32_t cb);
int so_share_plan(const double* X, const double* Y, int64_t n, double half_width, int32_t cb, const int32_t* hist,
                  int32_t k, double floor, double share_lo, double share_hi, uint8_t* flags, uint64_t* counts,
                  void* workspace, int64_t workspace_bytes, void* stream);

/* Stable compaction of the particles whose flag byte has a bit of bit_mask set: out_index[j] = index of the j-th
 * such particle (DEVICE int32[m]), and, where the pointers are not null, their positions and flag bytes.  m = the number
 * of flagged particles, known from so_share_plan's counts (the caller sizes the outputs with it); *count (DEVICE) receives
 * it again. */
int64_t so_share_compact_workspace_bytes(int64_t n);
int so_share_compact(const double* X, const double* Y, const uint8_t* flags, int64_t n, int32_t bit_mask, int64_t m,
                     int32_t* out_index, double* out_X, double* out_Y, uint8_t* out_flags, int64_t* count,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* so_knn_kth_distance_ex for one image where only the particles whose query_mask byte has bit 1 set are answered
 * (query_mask: DEVICE uint8[n] or null = all); every selected particle is a member of the tree.  dk of the others
 * is left untouched. */
int so_knn_kth_distance_q(const double* X, const double* Y, const int32_t* ix, int64_t n, int32_t k, double half_width,
                          const uint8_t* query_mask, double floor, double* dk, void* workspace,
                          int64_t workspace_bytes, void* stream);

/* so_img_deposit_batch_ld for ONE image shared by the GPUs of a node, with the exchange of the partial images fused into
 * the deposit (config 5; replaces deposit + ncclAllReduce).  Every pointer array holds one DEVICE pointer per rank
 * (the peers' mapped with so_ipc_open; entry `rank` is this rank's own buffer):
 *   partial  the ranks' partial images [n_chan, ndim, ndim]; partial[rank] must be `data` (dense)
 *   result   the ranks' result images, same layout
 *   masks    uint8 [n_ranks][ntiles] per rank, ntiles = ceil(ndim / SO_TILE)^2
 *   flags    uint64 [SO_MAX_PEERS + 2] per rank, zero before the first call (the last two slots receive the time [ns] this
 *            rank spent waiting at the two barriers of the call)
 *   epoch    generation of this call: starts at 1, +1 per call, the same on every rank
 * Steps, all stream-ordered: binning; every rank stores which sub-tiles it touches into everybody's masks; barrier
 * over the flags (release / acquire stores and loads at system scope through NVLink); deposit -- a sub-tile only this
 * rank touches is stored straight into EVERY rank's result while the deposit runs, so the transfer overlaps the
 * arithmetic tile by tile; barrier; sub-tiles several ranks touch are summed from their partial images in rank order, by
 * every rank for itself (sub-tiles nobody touches keep the zeros the result was cleared with before the first barrier).  Deterministic and identical on all ranks.  The caller
 * alternates two result (and mask) buffers between calls so that nobody overwrites what a slower rank still uses.  A rank
 * that never reaches a barrier makes the others trap after ~4 s instead of hanging. */
typedef struct {
    int32_t n_ranks, rank;
    uint64_t epoch;
    void* partial[SO_MAX_PEERS];
    void* result[SO_MAX_PEERS];
    void* masks[SO_MAX_PEERS];
    void* flags[SO_MAX_PEERS];
} so_push_t;
int so_img_deposit_push(const void* plan_workspace, const double* L, const int32_t* img, int64_t n,
                        int32_t n_chan, int32_t n_img, int32_t ndim, int64_t n_pairs,
                        float* data, int64_t data_ld, int64_t data_plane, float* simple, float* hist,
                        int32_t* tile_ids, int32_t* pair_pids,
                        void* workspace, int64_t workspace_bytes, const so_push_t* push, void* stream);

/* Sum of the ranks' partial images over NVLink peer memory (config 5; instead of ncclAllReduce of the whole image).
 * The image [n_planes, ndim, ndim] (dense float32) is cut into so_peer_block_edge() = 32 pixel square exchange blocks.
 *   so_peer_alloc / so_peer_free          device buffers that other processes may map (plain cudaMalloc)
 *   so_ipc_get_handle / so_ipc_open / so_ipc_close   cudaIpc handle (64 bytes, HOST) of such a buffer / mapping of a peer's
 *   so_peer_block_mask   mask[b] = 1 iff block b of the image holds a non-zero pixel (DEVICE uint8[nb * nb])
 *   so_peer_sum          masks = the N ranks' masks, gathered (DEVICE uint8[N][nb * nb]); partial_ptrs = HOST array of the N
 *                        partial images (DEVICE pointers, the peers' mapped).  out (DEVICE, this rank's own memory) = for
 *                        every block the sum over the ranks that touched it, in rank order; zero where nobody did.
 * The caller orders the steps across ranks: all partial images complete before so_peer_sum (the all-gather of the masks
 * does that), and no rank rewrites a partial image a peer may still be reading (two alternating buffers). */
int32_t so_peer_block_edge(void);
int so_peer_alloc(int64_t bytes, void** ptr);
int so_peer_free(void* ptr);
int so_ipc_get_handle(const void* ptr, void* handle64);
int so_ipc_open(const void* handle64, void** ptr);
int so_ipc_close(void* ptr);
int so_peer_block_mask(const float* img, int32_t n_planes, int32_t ndim, int64_t ld, int64_t plane, uint8_t* mask, void* stream);
int so_peer_sum(const void* const* partial_ptrs, const uint8_t* masks, int32_t n_ranks, int32_t rank,
                int32_t n_planes, int32_t ndim, float* out, void* stream);

/* ------------------------------------------------------------------------- */
/* measurement support (bench.py); no upstream equivalent                      */
/* ------------------------------------------------------------------------- */

/* Register-only micro-benchmark of one pipe on every SM (ctas_per_sm CTAs of 256 threads per SM, `iters`
 * iterations): kind 0 = FP32 FFMA, 1 = packed FFMA2, 2 = MUFU.EX2, 3 / 4 = the per-particle-wavelength arithmetic
 * of so_sed_dust_spectra with MUFU only / with its FMA-pipe polynomial share.  *ops_host (HOST pointer) receives the
 * operations one launch executes (multiply-adds for kinds 0-1, exponentials for 2-4); time the launch with CUDA
 * events on `stream`.  `out` is a device float the kernels never actually write. */
int so_peak_kernel(int32_t kind, int64_t iters, int32_t ctas_per_sm, float* out, double* ops_host, void* stream);

/* Per-kernel timing: while enabled, every labelled launch inside the library is bracketed by CUDA events on its
 * own stream.  so_prof_enable(0/1) also discards the marks collected so far.  so_prof_report synchronises on the
 * recorded events and writes "label<TAB>launches<TAB>total_ms" lines into buf (HOST); returns the bytes needed. */
int so_prof_enable(int32_t on);
int64_t so_prof_report(char* buf, int64_t buf_bytes);

#ifdef __cplusplus
}
#endif
#endif /* SYNTHOBS_B200_H */

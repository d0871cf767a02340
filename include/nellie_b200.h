/* nellie_b200 — C ABI of the B200-native structure-enhancement hot path.
 *
 * Drop-in boundary for aelefebv/nellie @ 54bf227: the entry points below are what a binding
 * of the reference's `Filter` (nellie/segmentation/filtering.py) and `Label`
 * (nellie/segmentation/labelling.py) stage classes calls instead of numpy / scipy.ndimage /
 * cupy (INTEGRATION.md shows the ctypes stub).  Each function names the reference lines it
 * replaces.  Conventions:
 *
 *   - every pointer is a DEVICE pointer owned by the caller (contiguous, C order, T[Z]YX
 *     frame layout), except arguments documented "host";
 *   - no allocation happens inside the library; scratch space is passed in;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no host sync);
 *   - return 0 on success, a negative NB200_ERR_* otherwise, text via nb200_last_error();
 *   - scalars that the reference derives from the data (gamma, thresholds, max|H|) stay
 *     in device memory between calls so a whole frame is enqueued without a round trip.
 *
 * Z window (multi-GPU slabs): a 3-D buffer holds `nz_buf` planes; buffer plane b is global
 * plane b + zg_off of a frame with `nz_glob` planes; kernels compute planes [zc0, zc1)
 * (buffer coordinates) and may read neighbouring planes that the caller filled by halo
 * exchange.  Reflection (Gaussian) and one-sided differences (Hessian) apply at the GLOBAL
 * frame border only.  Single GPU: nz_buf = nz_glob, zg_off = 0, zc0 = 0, zc1 = nz_buf.
 * 2-D frames: nz = 1 with the `_2d` entry points.
 */
#ifndef NELLIE_B200_H
#define NELLIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB200_ABI_VERSION 1

#define NB200_OK 0
#define NB200_ERR_ARG (-1)
#define NB200_ERR_CUDA (-2)
#define NB200_ERR_OOM (-3)          /* text contains "out of memory": nellie/utils/adaptive_run.py:116-127 */
#define NB200_ERR_UNSUPPORTED (-4)

/* geometry shared by the 3-D kernels */
typedef struct nb200_vol {
    int nz_buf, ny, nx; /* buffer extent */
    int zc0, zc1;       /* planes to compute, buffer coordinates */
    int zg_off;         /* global plane index of buffer plane 0 (may be negative) */
    int nz_glob;        /* planes in the whole frame */
} nb200_vol;

/* value transforms applied while scanning a sample buffer */
#define NB200_TF_NONE 0   /* v                       (gamma: gauss samples, filtering.py:369) */
#define NB200_TF_DIV 1    /* v / *divisor            (frob: sqrt(frob_sq)/max_abs, filtering.py:562) */
#define NB200_TF_LOG10 2  /* log10f(v)               (Label, labelling.py:450) */

/* int64[NB200_HIST_WORDS] device state of one 256-bin histogram threshold */
#define NB200_HIST_NBINS 256
#define NB200_HIST_MIN 0    /* ordered key of the minimum kept value */
#define NB200_HIST_MAX 1    /* ordered key of the maximum kept value */
#define NB200_HIST_COUNT 2  /* number of kept values */
#define NB200_HIST_BINS 3   /* 256 counts follow */
#define NB200_HIST_WORDS (3 + NB200_HIST_NBINS)

/* double[NB200_SP_WORDS] device record of the per-sigma scalars (filtering.py:839-848) */
#define NB200_SP_GAMMA 0      /* min(triangle, otsu) of positive gauss samples, eps fallback */
#define NB200_SP_GAMMA_SQ 1   /* 2*gamma^2 (f64, used as fl32) */
#define NB200_SP_FROB_THR 2   /* min(triangle, otsu) of frob samples (before division) */
#define NB200_SP_FROB_CUT 3   /* thr / frob_thresh_division: mask = frob > fl32(cut) */
#define NB200_SP_MAX_ABS 4    /* max |Hessian component| (1.0 if <= 0) */
#define NB200_SP_SKIP 5       /* 1.0 when the frob mask is empty: sigma skipped (filtering.py:843-844) */
#define NB200_SP_STATUS 6     /* 0 ok, 1 degenerate histogram (reference would raise) */
#define NB200_SP_TRI 7
#define NB200_SP_OTSU 8
#define NB200_SP_UNSAFE 9     /* 1.0 when the blurred volume holds values outside the exponent range the fast
                                 constant-divisor division was verified on: kernels fall back to IEEE division */
#define NB200_SP_FROBSQ_MIN 10 /* smallest frob_sq whose sqrt(.)/max_abs exceeds fl32(cut): mask = frob_sq >= this */
/* words written by nb200_finalize_frob_fast for nb200_frangi_fast (approximate classification with proven margins) */
#define NB200_SP_AMBIG 11     /* 1.0 when the bounds frob_max in [1, 3] cannot decide whether the mask is empty:
                                 nb200_hessian_stats_ambig + nb200_finalize_frob_resolve settle it exactly */
#define NB200_SP_FS_LO 12     /* approximate frob_sq below this: the voxel surely fails the mask */
#define NB200_SP_FS_HI 13     /* approximate frob_sq at or above this: the voxel surely passes the mask */
#define NB200_SP_ZT_C 14      /* margin subtracted from the approximate diagonal pair sums (provably-zero test) */
#define NB200_SP_DELTA 15     /* proven bound of |approximate - reference| for every Hessian entry of this sigma */
#define NB200_SP_WORDS 20

/* int64[NB200_HS_WORDS] device record reduced by nb200_hessian_stats */
#define NB200_HS_MAX_ABS_BITS 0   /* float bits of max|H| */
#define NB200_HS_MAX_FROBSQ_BITS 1 /* float bits of max frob_sq */
#define NB200_HS_MIN_NZ_COMPL 2   /* 0x7f800000 - float bits of the smallest non-zero |blurred value| (0 = none seen);
                                     complemented so that every word of the record reduces with MAX */
#define NB200_HS_MAX_G_BITS 3     /* float bits of the largest |blurred value| */
#define NB200_HS_APPROX_MAX_BITS 4 /* nb200_hessian_stats_fast: float bits of the approximate max|H| of the interior */
#define NB200_HS_FALLBACK 5       /* nb200_hessian_stats_fast: 1 when its exactness argument does not apply to this
                                     volume (nb200_finalize_max_abs then sets sp[UNSAFE]: the exact passes take over) */
#define NB200_HS_WORDS 8

/* int64[NB200_STATE_WORDS]: histogram state followed by the Hessian stats record, contiguous, so that a Z-sharded
 * caller moves both with ONE all-gather per reduction point (nb200_fold_records) */
#define NB200_STATE_WORDS (NB200_HIST_WORDS + NB200_HS_WORDS)

int nb200_abi_version(void);
const char* nb200_last_error(void);
/* number of SMs of the current device (grid sizing is a multiple of it) */
int nb200_sm_count(void);

/* ---- F1: cascaded Gaussian --------------------------------------------------------------
 * One axis of scipy.ndimage.gaussian_filter(mode="reflect") as called at filtering.py:828-835:
 * float32 in, float64 accumulate in scipy's pair order, float32 out.  `weights` (HOST, double)
 * holds w[0..radius], w[0] the centre tap, normalised as scipy's _gaussian_kernel1d does.
 * axis: 0 = Z, 1 = Y, 2 = X.  src != dst. */
int nb200_gauss_axis(const float* src, float* dst, const nb200_vol* vol, int axis,
                     const double* weights, int radius, void* stream);
/* The Y pass followed by the X pass of the same call (axes 1 and 2 of scipy's loop, with the float32
 * intermediate scipy stores between them) fused into one kernel: 8 B/voxel of HBM traffic for two axes
 * instead of 16.  Same radius on both axes, 1 <= radius <= 8 (NB200_ERR_UNSUPPORTED otherwise: use two
 * nb200_gauss_axis calls).  wy / wx: HOST double[radius+1] taps of the two axes. */
int nb200_gauss_yx(const float* src, float* dst, const nb200_vol* vol, const double* wy, const double* wx,
                   int radius, void* stream);

/* ---- F2: threshold sampling lattice -----------------------------------------------------
 * arr[::sz, ::sy, ::sx] on the GLOBAL lattice (filtering.py:328-363); every lattice point of
 * the planes [zc0,zc1) is written (consumers keep values > 0).  `out` holds
 * n_lattice_planes * ceil(ny/sy) * ceil(nx/sx) floats, planes in ascending z. */
int nb200_lattice_sample(const float* src, const nb200_vol* vol, int sz, int sy, int sx,
                         float* out, void* stream);
/* flat[offset::step] of a contiguous array (labelling.py:412), n_out = ceil((n-offset)/step);
 * if gate != NULL, values whose gate[i] <= gate_thresh are written as 0 (labelling.py:417). */
int nb200_strided_sample(const float* src, long long n, long long offset, long long step,
                         const float* gate, float gate_thresh, float* out, void* stream);

/* ---- U1/U2: 256-bin histogram thresholds (utils/gpu_functions.py:23-94) -------------------
 * hist_reset zeroes the state; hist_minmax folds min/max/count of the kept (> 0) transformed
 * values; hist_bins adds np.histogram(bins=256, range=(min,max)) counts (float32 edges with
 * numpy's edge correction).  Between the calls a multi-GPU caller all-reduces the state
 * (MIN/MAX/SUM per field). `divisor` is a device double* (NB200_TF_DIV) or NULL. */
int nb200_hist_reset(long long* state, void* stream);
int nb200_hist_minmax(const float* vals, long long n, int transform, const double* divisor,
                      long long* state, void* stream);
int nb200_hist_bins(const float* vals, long long n, int transform, const double* divisor,
                    long long* state, void* stream);
/* Multi-GPU reduction of the threshold state ("pack into <= 2 calls"): `gathered` holds the
 * NB200_STATE_WORDS-word records of all `world` ranks (rank-major, e.g. from ncclAllGather); folds them into `state`.
 * stage 0: HIST_MIN -> min, HIST_MAX -> max, Hessian stats -> max (every stats word reduces with MAX);
 * stage 1: HIST_COUNT and the 256 bins -> sum, Hessian stats -> max.  Other words of `state` are left alone. */
int nb200_fold_records(const long long* gathered, int world, int stage, long long* state, void* stream);
/* The same for `count` records per rank (rank r's records start at gathered + r * count * NB200_STATE_WORDS): a Z-sharded
 * frame that runs a reduction point for all its sigmas at once (one all-gather instead of one per sigma). */
int nb200_fold_records_n(const long long* gathered, int world, int count, int stage, long long* state, void* stream);
/* gamma = min(triangle, otsu) (filtering.py:365-380) -> sp[GAMMA], sp[GAMMA_SQ] */
int nb200_finalize_gamma(const long long* state, double* sp, void* stream);
/* Frobenius threshold (filtering.py:407-444): consumes the histogram of frob samples and the
 * reduced Hessian stats; fixed_thresh = NaN for auto. -> sp[FROB_THR..SKIP] */
int nb200_finalize_frob(const long long* state, const long long* hstats, double fixed_thresh,
                        double division, double* sp, void* stream);
/* max|H| -> sp[MAX_ABS] (must run before hist_* with NB200_TF_DIV on &sp[MAX_ABS]) */
int nb200_finalize_max_abs(const long long* hstats, double* sp, void* stream);
/* Label threshold (labelling.py:440-455): out (device double[7]): out[0] = min(10**tri, 10**otsu) as float32 value,
 * out[1] = 10**tri, out[2] = 10**otsu (float64 pow rounded once to float32), out[3] = 1 if no samples (None),
 * out[4] = status, out[5] / out[6] = triangle / Otsu threshold in the histogram's own domain (log10): a host caller
 * that wants the reference's bits applies `10 ** np.float32(.)` to these itself (numpy's scalar float32 power is
 * libm powf, which is not correctly rounded for ~0.05 % of arguments).  The samples are transformed with numpy's own
 * float32 log10 (devmath.cuh np_log10f).  With log_domain = 0 it is the plain Otsu of labelling.py:457-465. */
int nb200_finalize_label_threshold(const long long* state, int log_domain, double* out, void* stream);

/* Otsu threshold of an INTEGER frame's positive samples, as numpy evaluates it (labelling.py:457-465 ->
 * gpu_functions.py:23-50): np.histogram bins integer data with FLOAT64 edges and the threshold is a float64 bin centre.
 * vals: the samples as float32 (exact for uint8 / uint16); state: after nb200_hist_reset + nb200_hist_minmax(TF_NONE).
 * nb200_finalize_otsu_f64: out[0] = threshold, out[3] != 0 -> no samples, out[4] != 0 -> degenerate (NaN variance). */
int nb200_hist_bins_f64(const float* vals, long long n, long long* state, void* stream);
int nb200_finalize_otsu_f64(const long long* state, double* out, void* stream);

/* ---- F4: Hessian statistics -------------------------------------------------------------
 * Finite-difference Hessian of numpy.gradient(numpy.gradient(.)) (filtering.py:446-562) at
 * every voxel of [zc0,zc1): reduces max|component| and max frob_sq into hstats and writes
 * sqrt(frob_sq) at the lattice points into `frob_samples` (layout of nb200_lattice_sample).
 * spacing[6] (HOST, float): {fl32(hz), fl32(2hz), fl32(hy), fl32(2hy), fl32(hx), fl32(2hx)}. */
int nb200_hessian_stats(const float* gauss, const nb200_vol* vol, const float* spacing, int div_mode,
                        const double* sp, int sz, int sy, int sx, float* frob_samples, long long* hstats,
                        void* stream);
/* Same, and additionally leaves one float per voxel in `code` (volume of the buffer shape, planes [zc0,zc1)
 * written) for nb200_frangi_sparse: |code| = frob_sq as the mask tests it, sign bit set when the vesselness of
 * the voxel is provably zero (a pair of diagonal Hessian entries with a clearly positive sum: lambda_2 or
 * lambda_3 > 0, filtering.py:759-761).  code may be NULL (plain nb200_hessian_stats). */
int nb200_hessian_stats_code(const float* gauss, const nb200_vol* vol, const float* spacing, int div_mode,
                             const double* sp, int sz, int sy, int sx, float* frob_samples, long long* hstats,
                             float* code, void* stream);
/* Safety net of the fast constant-divisor division (NB200_DIV_FAST): nb200_hessian_stats[_code] also reduces the
 * range of the non-zero blurred values; nb200_finalize_max_abs sets sp[NB200_SP_UNSAFE] when that range could
 * produce a numerator outside the exponents the division was verified on.  This call then recomputes the
 * statistics (and code) with IEEE division — it enqueues kernels that return immediately when the flag is clear,
 * and is a no-op for the other division modes.  Call it after nb200_finalize_max_abs, then reduce hstats and
 * finalize again. */
int nb200_hessian_stats_redo(const float* gauss, const nb200_vol* vol, const float* spacing, int div_mode,
                             const double* sp, int sz, int sy, int sx, float* frob_samples, long long* hstats,
                             float* code, void* stream);
/* Exact statistics pass that only runs when nb200_finalize_frob_fast left sp[NB200_SP_AMBIG] set (and the redo pass
 * has not run): provides the exact max frob_sq for nb200_finalize_frob_resolve. */
int nb200_hessian_stats_ambig(const float* gauss, const nb200_vol* vol, const float* spacing, int div_mode,
                              const double* sp, int sz, int sy, int sx, float* frob_samples, long long* hstats,
                              void* stream);

/* ---- F4, fast form: the same statistics from an APPROXIMATE Hessian with proven error bounds ---------------
 * (filtering.py:446-562).  The interior march evaluates the six second differences with plain float32
 * subtractions (no divisions), keeps per-warp maxima, and appends every sub-chunk whose maximum is within 2^-10 of
 * the running maximum to a work list; a second kernel re-evaluates the listed sub-chunks that can hold the true
 * maximum (approximate maximum minus twice the proven error bound) with the reference's exact arithmetic, so
 * hstats[MAX_ABS_BITS] ends up bit-identical to nb200_hessian_stats.  sqrt(frob_sq) at the lattice points is
 * evaluated exactly inside the march; the border shell runs through the exact per-voxel kernel.  max frob_sq is NOT
 * produced (see nb200_finalize_frob_fast).  When the argument does not apply (value range outside the verified
 * exponents, error bound not small against the maximum, work list overflow) hstats[FALLBACK] is set and
 * nb200_finalize_max_abs raises sp[UNSAFE], which makes nb200_hessian_stats_redo recompute everything exactly.
 * Requires nx % 4 == 0, 16-byte aligned gauss, div_mode FAST or POW2 (NB200_ERR_UNSUPPORTED otherwise: use
 * nb200_hessian_stats_code).  workspace: nb200_hessian_fast_workspace_bytes() bytes of device memory. */
size_t nb200_hessian_fast_workspace_bytes(void);
int nb200_hessian_stats_fast(const float* gauss, const nb200_vol* vol, const float* spacing, int div_mode,
                             int sz, int sy, int sx, float* frob_samples, long long* hstats, void* workspace,
                             void* stream);
/* Frobenius threshold for the fast path: as nb200_finalize_frob, but without max frob_sq.  Emptiness of the mask
 * follows from the bounds 1 <= max frob <= 3 (the voxel attaining max|H| has frob >= 1; frob_sq <= 9 max|H|^2);
 * when the cut falls between them sp[AMBIG] is set.  Also writes the classification thresholds
 * sp[FS_LO], sp[FS_HI], sp[ZT_C], sp[DELTA] from the proven error bound delta = 40 * 2^-24 * max|g| * max_scale
 * (max_scale = largest 1/(fl32(2h_a) * fl32(2h_b))).  mask_enabled = 0: every voxel passes (Filter mask=False).
 * max_scale = 0 selects the plain nb200_finalize_frob behaviour (exact max frob_sq in hstats) with the mask switch. */
int nb200_finalize_frob_fast(const long long* state, const long long* hstats, double fixed_thresh, double division,
                             double max_scale, int mask_enabled, double* sp, void* stream);
/* After nb200_hessian_stats_ambig: sp[SKIP] from the exact max frob_sq when sp[AMBIG] was set. */
int nb200_finalize_frob_resolve(const long long* hstats, double* sp, void* stream);

/* Division mode for one grid-spacing divisor d = fl32(h) or fl32(2h) (synchronous, init time only):
 *   2 (NB200_DIV_POW2)  d is a power of two: multiply by the exact reciprocal;
 *   1 (NB200_DIV_FAST)  q = fma(fma(-n*r, d, n), r, n*r), r = RN(1/d), verified HERE bit-for-bit against IEEE
 *                       division for every numerator with exponent in [-90, 90] and +-0;
 *   0 (NB200_DIV_IEEE)  plain correctly rounded division.
 * A launch uses one mode for all six divisors: the weakest of their modes. */
#define NB200_DIV_IEEE 0
#define NB200_DIV_FAST 1
#define NB200_DIV_POW2 2
int nb200_divisor_mode(float d, int* mode_out, void* stream);
int nb200_hstats_reset(long long* hstats, void* stream);
/* The six second derivatives themselves (tests / diagnostics): out6 = 6 volumes of the buffer shape in the
 * order d0d0, d1d0, d2d0, d1d1, d2d1, d2d2 (the reference's hxx, hxy, hxz, hyy, hyz, hzz; axes 0,1,2 = Z,Y,X). */
int nb200_hessian_components(const float* gauss, const nb200_vol* vol, const float* spacing, int div_mode,
                             float* out6, void* stream);

/* ---- F4-F9 fused: Hessian + Frobenius mask + eigenvalues + vesselness + max/AND ------------
 * (filtering.py:842-851).  acc holds max-over-sigma vesselness for live voxels and -1 for
 * voxels that failed the mask at any non-skipped sigma; acc must be zero-filled before the
 * first sigma.  Reads gamma_sq / frob cut / max_abs / skip from the device record `sp`. */
int nb200_frangi_accumulate(const float* gauss, float* acc, const nb200_vol* vol, const float* spacing,
                            int div_mode, float alpha_sq, float beta_sq, const double* sp, void* stream);
/* The same step from the record of nb200_hessian_stats_code: the dense part only streams code and acc
 * (12 B/voxel); Hessian, eigenvalues and vesselness are evaluated for the few percent of voxels that are
 * alive, pass the mask and are not provably zero, compacted through shared-memory work queues.  Results are
 * bit-identical to nb200_frangi_accumulate. */
int nb200_frangi_sparse(const float* gauss, const float* code, float* acc, const nb200_vol* vol,
                        const float* spacing, int div_mode, float alpha_sq, float beta_sq, const double* sp,
                        unsigned* list, long long list_capacity, unsigned long long* counter, void* stream);
/* ---- F4-F9 fused, fast form ---------------------------------------------------------------------------------
 * One Z-march over the blurred volume (TMA ring, no CTA barrier): the approximate Hessian classifies every live
 * voxel as surely failing the mask (acc = -1), surely passing with a provably zero response (nothing to do), or
 * candidate; candidates get the reference's exact Hessian from the staged planes, the exact mask test, the exact
 * provably-zero tests, and the survivors the float64 eigenvalues + vesselness, all inside the same kernel
 * (per-warp queues, full warps).  Results are bit-identical to nb200_frangi_accumulate.  Returns at once when
 * sp[SKIP] or sp[UNSAFE] is set (the caller then runs nb200_frangi_sparse_gated).  Same requirements as
 * nb200_hessian_stats_fast.  diag: optional device uint64[8] counters (tests / profiling). */
int nb200_frangi_fast(const float* gauss, float* acc, const nb200_vol* vol, const float* spacing, int div_mode,
                      float alpha_sq, float beta_sq, const double* sp, unsigned long long* diag, void* stream);
/* nb200_frangi_sparse that only runs when sp[UNSAFE] is set (fallback of nb200_frangi_fast). */
int nb200_frangi_sparse_gated(const float* gauss, const float* code, float* acc, const nb200_vol* vol,
                              const float* spacing, int div_mode, float alpha_sq, float beta_sq, const double* sp,
                              unsigned* list, long long list_capacity, unsigned long long* counter, void* stream);
/*   list / counter (optional scratch): device buffer of `list_capacity` uint32 (>= voxels of [zc0,zc1)) and one
 *   device uint64.  With them (and a buffer below 2^32 voxels) the step runs as a barrier-free stream kernel that
 *   appends the candidates to the list plus a solve kernel that walks it; without, as one kernel with
 *   shared-memory queues.  Same results either way. */
/* 2-D variant (closed-form 2x2 eigenvalues, filtering.py:676-690, :737-741); spacing[4] = y,x */
int nb200_frangi_accumulate_2d(const float* gauss, float* acc, int ny, int nx, const float* spacing,
                               float beta_sq, const double* sp, void* stream);
int nb200_hessian_stats_2d(const float* gauss, int ny, int nx, const float* spacing, int sy, int sx,
                           float* frob_samples, long long* hstats, void* stream);

/* ---- F11: percentile + opening ----------------------------------------------------------
 * np.percentile(positive lattice sample, 1) (filtering.py:963) by radix select; scratch is
 * int64[2048+8] device words; out[0] = threshold, out[1] = number of positive samples. */
int nb200_percentile(const float* samples, long long n, double q_percent, long long* scratch,
                     double* out, void* stream);
/* out = V * binary_opening(V > thr) with V = max(acc, 0) (filtering.py:926, :964-966); cross
 * structuring element, border_value 0.  If out_thr[1] == 0 (no positive sample) V passes
 * through unchanged (filtering.py:959-960).  Needs 2 halo planes of acc on interior slab sides. */
int nb200_finalize_opening(const float* acc, float* out, const nb200_vol* vol, const double* thr, void* stream);
int nb200_finalize_opening_2d(const float* v, float* out, int ny, int nx, const double* thr, void* stream);

/* ---- F12: Filter._remove_edges (filtering.py:969-1000, :227-250; off by default) -------------------------------
 * In every slice of v (nz slices of ny x nx; nz = 1 for a 2-D frame) the rows of the bounding box of the positive
 * response are found and min(margin, height) rows are zeroed at its top and bottom; the reference uses margin = 15.
 * In place; entries <= 0 count as empty (the accumulator keeps -1 for dead voxels). */
int nb200_remove_edges(float* v, int nz, int ny, int nx, int margin, void* stream);

/* ---- F10: 2-D multi-scale LoG blobness (filtering.py:772-795, :927-930) ------------------------
 * t0/t1: the two separable second-derivative Gaussians of scipy.ndimage.gaussian_laplace (built with
 * nb200_gauss_axis and order-2 taps); acc: the sigma-loop accumulator (>= 0 alive, -1 dead).
 * accumulate: L = first ? cur : max(L, cur), cur = (-(t0+t1)) * sigma_sq * alive.
 * combine: V = max(max(acc,0), max((max(L,0) / (max L + 1e-12)) / 10, 0)); max_bits: device int64 scratch. */
int nb200_log2d_accumulate(const float* t0, const float* t1, const float* acc, float sigma_sq, int first,
                           long long n, float* L, void* stream);
int nb200_log2d_combine(const float* acc, const float* L, long long n, long long* max_bits, float* v, void* stream);

/* ---- L4/L5: Label._get_labels (labelling.py:467-509, :546-556) -------------------------------
 * threshold (strict >, optional intensity gate on `raw`) -> fill holes (3-D only) -> 26-/8-connected
 * components -> drop components smaller than min_area -> 3^d majority smoothing -> components again.
 * labels: int32, 0 = background, ids 1..n in raster order of each component's first voxel
 * (scipy.ndimage.label numbering).  thr: device double[>= 5] as written by
 * nb200_finalize_label_threshold (thr[0] = threshold, thr[3] != 0 -> "None": empty mask).
 * nz = 1 selects the 2-D path.  workspace: nb200_label_workspace_bytes() bytes of device memory.
 * n_labels: device int64 receiving the component count. */
size_t nb200_label_workspace_bytes(int nz, int ny, int nx);
int nb200_label_frame(const float* frangi, const float* raw, int use_intensity, float intensity_thresh,
                      const double* thr, int nz, int ny, int nx, long long min_area, int fill_holes,
                      int* labels, void* workspace, long long* n_labels, void* stream);
/* scipy.ndimage.label alone on a uint8 mask; connectivity_full: 1 = 26/8-connected, 0 = 6/4. */
int nb200_ccl_label(const unsigned char* mask, int nz, int ny, int nx, int connectivity_full, int* labels,
                    void* workspace, long long* n_labels, void* stream);

/* ---- Network stage, the array kernels its GPU backend runs ------------------------------------------------
 * skel / labels: int32 frames (nz = 1 for 2-D).  Everything else of the stage (skeletonize, _add_missing_skeleton_labels,
 * _relabel_objects) runs on the host in the reference too and is not part of this library.
 * nb200_pixel_class: networking.py:669-680 — out (uint8) = skel > 0 ? min(4, set voxels in the 3^d window, zero outside) : 0.
 * nb200_branch_labels: networking.py:758-797 — scipy.ndimage.label(ones(3^d)) of (pixel_class > 0) & (pixel_class != 4);
 *   workspace = nb200_label_workspace_bytes(), ids in raster order of each component's first voxel.
 * nb200_remove_connected_label_pixels: networking.py:261-296 — a labelled voxel off the frame boundary whose 3^d window
 *   holds two different positive labels becomes 0 (out != labels). */
int nb200_pixel_class(const int* skel, int nz, int ny, int nx, unsigned char* out, void* stream);
int nb200_branch_labels(const unsigned char* pixel_class, int nz, int ny, int nx, int* labels, void* workspace,
                        long long* n_labels, void* stream);
int nb200_remove_connected_label_pixels(const int* labels, int nz, int ny, int nx, int* out, void* stream);

/* ---- Markers stage: nellie/segmentation/mocap_marking.py, full-volume branch --------------------------------
 * Frames are (nz, ny, nx), nz = 1 for 2-D.  All kernels are exact (integer / comparison work, one float32 multiply), so the
 * three outputs of Markers._run_frame_impl (mocap_marking.py:648-703) are bit-identical to the reference's.
 * nb200_markers_mask_border: mask = labels > 0 (:657) and border = binary_dilation(mask, cross, 1 iteration) ^ mask (:440),
 *   both uint8 0/1.
 * nb200_markers_edt: scipy.ndimage.distance_transform_edt(mask).astype(float32) clamped with np.minimum(., clamp) (:444-447).
 *   Exact squared distances by three windowed min-plus passes (X, Y, Z) with early exit; a voxel whose nearest background
 *   voxel is further than `window` in some axis has a distance > window, so with window >= clamp the clamped result is
 *   exact (NB200_ERR_ARG otherwise).  1 <= window <= 147 (squared distances are carried as uint16).  Voxels outside the
 *   frame are NOT background (scipy semantics).  scratch: 2 * nz*ny*nx uint16.
 * nb200_markers_log_response: resp = -(d0 + d1 [+ d2]) * sigma_sq, negatives set to 0 (:489-494); d0, d1, d2 are the
 *   separable second-derivative Gaussians of scipy.ndimage.gaussian_laplace in axis order (built with nb200_gauss_axis /
 *   nb200_gauss_yx and order-2 taps), d2 = NULL for 2-D; resp may alias d0.
 * nb200_markers_peak_update: one scale of :496-505 — a voxel with mask != 0, distance > 0, resp equal to the maximum of its
 *   3^d neighbourhood (clamped at the frame border) and resp > best gets best = resp, peak = 1.  best / peak are zeroed by
 *   the caller before the first scale.
 * nb200_markers_peak_update_fused: nb200_markers_log_response + nb200_markers_peak_update in one kernel without the response
 *   volume (the response is re-evaluated from d0, d1, d2 where it is needed: mask voxels and the neighbours of candidates);
 *   identical best / peak.  This is what the stage runs; the two-step form remains for callers that want the response.
 * nb200_markers_nms: :595-606 — marker = 1 where peak != 0, intensity > 0 and no peak inside the (2*radius+1)^d window
 *   (clamped at the frame border) has a larger intensity; equal intensities keep each other, as the reference's
 *   score == maximum_filter(score) does.  intensity: the raw frame as float32 (the cast of score_img[...] = intensity). */
int nb200_markers_mask_border(const int* labels, int nz, int ny, int nx, unsigned char* mask, unsigned char* border,
                              void* stream);
int nb200_markers_edt(const unsigned char* mask, int nz, int ny, int nx, int window, float clamp,
                      unsigned short* scratch, float* distance, void* stream);
int nb200_markers_log_response(const float* d0, const float* d1, const float* d2, long long n, float sigma_sq,
                               float* resp, void* stream);
int nb200_markers_peak_update(const float* resp, const unsigned char* mask, const float* distance, int nz, int ny, int nx,
                              float* best, unsigned char* peak, void* stream);
int nb200_markers_peak_update_fused(const float* d0, const float* d1, const float* d2, float sigma_sq,
                                    const unsigned char* mask, const float* distance, int nz, int ny, int nx, float* best,
                                    unsigned char* peak, void* stream);
int nb200_markers_nms(const unsigned char* peak, const float* intensity, int nz, int ny, int nx, int radius,
                      unsigned char* marker, void* stream);

/* ---- HuMomentTracking, per-frame feature extraction: nellie/tracking/hu_tracking.py ---------------------------
 * Frames are (nz, ny, nx) float32 device copies (nz = 1 for 2-D); markers are rows of `coords`, int64 (n, ndim) voxel indices
 * in raster order (argwhere of the marker frame); `bounds` is int32 (n, 6) = lo / hi for Z, Y, X.
 * nb200_hu_frangi_transform: hu_tracking.py:604-612 — out = log10 of the positive values (numpy's float32 log10, bit for
 *   bit), other values copied; then every negative value minus the minimum of the negatives.  scratch: one uint32 that the
 *   caller sets to 0xFFFFFFFF before the call.
 * nb200_hu_distance_max: :614-616 — 2 * maximum over the 3^d neighbourhood (border voxels repeated).
 * nb200_hu_bounds: :392-421 (_get_im_bounds) — half width r = ceil(distance_max[marker]); lo = clip(m - r, 0, size),
 *   hi = clip(m + r + 1, 0, size); max_half (device int, zeroed by the caller) receives the largest r: the reference's ROI
 *   cube has side 2 * max_half + 1 (:633).
 * nb200_hu_roi_stats: :341-390 (_calculate_mean_and_variance) per marker -> stats float32 (n, 2) = mean and variance of the
 *   non-zero ROI voxels.  cube > 0: the dense path's zero-padded ROI cube of that side is what numpy reduces; cube = 0: the
 *   ROI box itself (streaming path, :682-750).  int_bits: 0 = float32 frame (numpy's pairwise float32 summation restated),
 *   8 / 16 = unsigned integer frame of that width (exact integer sums, squares wrapped to the width as numpy does).
 *   Bit-identical to the reference in all four combinations.
 * nb200_hu_log_moments: :225-325, :544-571 — per marker the log-Hu features of the ROI (2-D: 6) or of its three maximum
 *   projections along Z, Y, X (3-D: 18) -> out float64 (n, 6 | 18).  proj: float32 scratch (n, 1 | 3, side, side) with
 *   side >= the largest box extent; cube as above (0 joins the maximum where the box is shorter than the cube);
 *   integer_frame != 0: exact integer raw moments.  Agrees with the reference to float64 rounding (numpy's SIMD pow is not
 *   correctly rounded and not reproduced), not to the bit. */
int nb200_hu_frangi_transform(const float* frangi, long long n, float* out, unsigned int* scratch, void* stream);
int nb200_hu_distance_max(const float* distance, int nz, int ny, int nx, float* out, void* stream);
int nb200_hu_bounds(const long long* coords, long long n, int ndim, const float* distance_max, int nz, int ny, int nx,
                    int* bounds, int* max_half, void* stream);
int nb200_hu_roi_stats(const float* frame, int nz, int ny, int nx, const int* bounds, long long n, int ndim, int cube,
                       int int_bits, float* stats, void* stream);
int nb200_hu_log_moments(const float* frame, int nz, int ny, int nx, const int* bounds, long long n, int ndim, int side,
                         int cube, int integer_frame, float* proj, double* out, void* stream);

/* ---- Network stage, host steps of the reference moved to the device -------------------------------------------
 * Frames are int32 (nz, ny, nx), nz = 1 for 2-D, fewer than 2^32 voxels; object ids 1..max_label.
 * nb200_network_add_missing: networking.py:315-392 — an object whose label does not occur in `skel` gets skel[p] = label at
 *   the voxel p of its largest Frangi response (first voxel in raster order among equal values; scipy's own choice among
 *   exactly equal values follows an unstable sort).  key: uint64[max_label + 1], in_skel: uint8[max_label + 1], both zeroed
 *   by the caller.
 * nb200_network_skeleton_labels: :835 — out = skel > 0 ? labels : 0.
 * nb200_network_object_boxes: scipy.ndimage.find_objects — boxes int32 (max_label + 1, 6) = min z, y, x (caller fills
 *   INT_MAX) and max z, y, x (caller fills -1); seeded[lab] (uint8, zeroed) = 1 when the object holds a voxel with branch > 0.
 * nb200_network_relabel: :485-577 (_relabel_objects) — for every row of `crops` (int64 (m, 8) = label, z0, y0, x0, extents
 *   ez, ey, ex, offset of the crop in crop space; ascending offsets) the feature transform of the object's seeds
 *   (label == row label and branch > 0) over its bounding box, exactly as scipy.ndimage.distance_transform_edt(~seeds,
 *   sampling, return_indices=True) computes it (tie-breaking included), then out[v] = branch[nearest seed of v] for the
 *   object's voxels; other voxels of `out` (uint32, zeroed by the caller) are left alone.  crop_voxels = sum of the box
 *   volumes; line_starts int64 (3, m) = per axis the exclusive prefix sum of the number of crop lines along that axis,
 *   n_lines (HOST, 3) their totals; sampling (HOST double[3]) = voxel size along Z, Y, X; ft_a, ft_b: int32 (3, crop_voxels),
 *   stack: int32 (crop_voxels). */
int nb200_network_add_missing(const int* labels, const float* frangi, int* skel, int nz, int ny, int nx, int max_label,
                              unsigned long long* key, unsigned char* in_skel, void* stream);
int nb200_network_skeleton_labels(const int* skel, const int* labels, long long n, int* out, void* stream);
int nb200_network_object_boxes(const int* labels, const int* branch, int nz, int ny, int nx, int max_label, int* boxes,
                               unsigned char* seeded, void* stream);
int nb200_network_relabel(const int* labels, const int* branch, int nz, int ny, int nx, const long long* crops, long long m,
                          long long crop_voxels, const long long* line_starts, const long long* n_lines,
                          const double* sampling, int* ft_a, int* ft_b, int* stack, unsigned int* out, void* stream);

/* ---- Label thresholds with histogram_nbins != 256 (labelling.py:23-35, :440-465; gpu_functions.py:23-94) ---------------------
 * General-bin-count form of nb200_hist_bins + nb200_finalize_label_threshold / nb200_hist_bins_f64 + nb200_finalize_otsu_f64
 * (those are specialised for the reference default of 256 and stay the production path).  state: after nb200_hist_reset +
 * nb200_hist_minmax over the same `vals` (transform NONE, or LOG10 when log_domain != 0).  f64_edges != 0: the samples are an
 * integer frame's values (np.histogram then bins with float64 edges; log_domain must be 0).  otsu_only != 0: out[0] = the Otsu
 * bin centre (the intensity threshold of labelling.py:457-465); otherwise out[] is laid out as nb200_finalize_label_threshold
 * writes it.  workspace: nb200_histn_workspace_bytes(nbins) bytes, 8-byte aligned.  2 <= nbins <= 2^24. */
size_t nb200_histn_workspace_bytes(int nbins);
int nb200_histn_threshold(const float* vals, long long n, int log_domain, int f64_edges, int otsu_only, int nbins,
                          const long long* state, void* workspace, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NELLIE_B200_H */

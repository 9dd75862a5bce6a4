/*
 * rvs_b200.h -- C ABI of librvs_b200.so, the sm_100a implementation of the
 * rvspecfit per-spectrum likelihood hot path.
 *
 * Conventions
 *   - Plain pointers and sizes only.  Pointers named d_* are DEVICE pointers,
 *     h_* are HOST pointers.  `stream` is a cudaStream_t passed as void*
 *     (NULL = default stream).  Every call is stream-ordered and re-entrant;
 *     the library keeps no global state besides cached cuFFT plans.
 *   - Return value: 0 on success, a negative RVS_E_* code on a call-level
 *     error (rvs_last_error() gives the text).  Per-item conditions are
 *     reported through int32 status arrays (RVS_ST_* bits), never by aborting:
 *     the Python mirror turns them into the exceptions the reference raises.
 *   - All arithmetic is fp64; the template grid is stored fp32 (or fp64) as
 *     the reference stores it (make_interpol.py:363-364) and promoted on use.
 *
 * Reference interfaces replaced (paths under /root/reference/py/rvspecfit/):
 *   rvs_spline_construct / rvs_spline_eval  <- cffi `_spliner.construct`,
 *       `_spliner.evaler` (ffibuilder.py:10-17, src/spliner.c:7-108)
 *   rvs_locate_grid      <- GridInterp.__call__ vertex and weight location
 *       (spec_inter.py:153-194)
 *   rvs_template_build   <- GridInterp.__call__ / TriInterp.__call__ weighted
 *       sum + exp (spec_inter.py:134-194, 35-59), convolve_vsini +
 *       compute_vsini_kernel (spec_fit.py:565-682), Spline.__init__
 *       (spliner.py:10-32)
 *   rvs_obs_prepare, rvs_basis_build <- SpecData.__init__ products and
 *       get_poly_basis (spec_fit.py:103-108, 148-176)
 *   rvs_chisq_scan, rvs_chisq_fused <- evalRV + get_chisq0 per arm inside
 *       get_chisq / find_best (spec_fit.py:707-727, 205-249, 879-975, 1062-1071)
 *   rvs_scan_stats       <- find_best tail (spec_fit.py:1072-1092)
 *   rvs_ccf_*            <- fitter_ccf.fit hot loop (fitter_ccf.py:126-232)
 */
#ifndef RVS_B200_H
#define RVS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVS_E_ARG -1      /* bad argument (size, null pointer, unsupported npoly) */
#define RVS_E_CUDA -2     /* CUDA runtime / cuFFT error */
#define RVS_E_NODEVICE -3 /* no CUDA device */
#define RVS_E_LIMIT -4    /* problem exceeds a compiled-in limit */

/* per-item status bits */
#define RVS_ST_OK 0
#define RVS_ST_TEMPLATE_BAD 1 /* template not finite or |value| > 1e100 */
#define RVS_ST_NOT_PD 2       /* continuum normal matrix not positive definite /
                                 result not finite: caller takes the SVD route */
#define RVS_ST_RANGE 4        /* evaluation wavelength outside the template */
#define RVS_ST_TAPS 8         /* vsini kernel longer than the tap capacity (truncated) */
#define RVS_ST_LIMIT 16       /* item does not fit the fused path's shared-memory window:
                                 re-evaluate it with rvs_template_build + rvs_chisq_scan */

#define RVS_MAX_NPOLY 16
#define RVS_MAX_TAPS 2048      /* one-sided vsini taps, rvs_template_build */
#define RVS_MAX_FUSED_TAPS 128 /* one-sided vsini taps, rvs_chisq_fused */

const char *rvs_last_error(void);
int rvs_version(void);
/* number of CUDA kernels this library has launched in this process (for
 * bench.py's gpu_launches) */
int64_t rvs_launch_count(void);
/* Optional per-stage timing of the fused evaluation (bench.py --stage-profile):
 * when enabled, CUDA events bracket every kernel on its launching stream.
 * rvs_profile_read synchronises the device, writes the summed milliseconds and
 * launch counts of stages 0..nstage-1 (locate, prep, chunk, gram, solve,
 * resid), clears the records and returns the number of stages. */
void rvs_profile_enable(int on);
int rvs_profile_active(void);
int rvs_profile_read(double *ms_total, int64_t *launches, int nstage);
/* Same records as a timeline: out[3i..] = stage, start ms, end ms (relative to the
 * first record's start), in launch order; returns the count and clears them. */
int rvs_profile_timeline(double *out, int max_records);

/* ---- native spline: drop-in for the reference's cffi module ------------- */
/* Host-buffer entry points with the reference's exact signatures and status
 * codes (0 / -1 evaluation point outside knots / -2 knots not uniform).  They
 * run on the GPU (copy in, kernel, copy out). */
void rvs_spline_construct(double *h_xs, double *h_ys, int N, double *h_A, double *h_B,
                          double *h_C, double *h_D, double *h_h);
int rvs_spline_eval(double *h_evalx, int nevalx, int N, double *h_xs, double *h_hs,
                    double *h_As, double *h_Bs, double *h_Cs, double *h_Ds, int log_step,
                    double *h_ret);

/* ---- descriptors (plain C structs of pointers and sizes) ----------------- */
/* Knot grid of one spectral setup (template wavelength grid) and its Thomas
 * tables.  Filled on the host by rvs_knot_tables / rvs_knot_info; the d_*
 * members are device copies of the host arrays those functions produce. */
typedef struct {
  const double *d_lam_t; /* [npix_t] knots */
  const double *d_h;     /* [npix_t-1] x[i+1]-x[i] */
  const double *d_hinv;  /* [npix_t-1] 1/h */
  const double *d_cp;    /* [npix_t-2] modified super-diagonal (spliner.c:33-37) */
  const double *d_winv;  /* [npix_t-2] reciprocal pivots */
  int32_t npix_t;
  int32_t log_step;      /* 1: knots uniform in ln(x); 0: uniform in x */
  double x0, xlast;      /* first / last knot */
  double q0, qstep_inv;  /* ln(x0) (or x0) and 1/step of the uniform coordinate */
  double lnstep;         /* ln(x[1]/x[0]) (vsini kernel scale) */
  double ratio;          /* h[k+1]/h[k] of the ideal grid: (x[n-1]/x[0])^(1/(n-1)), or 1 */
  double ratio_dev;      /* max_k |h[k+1] / (ratio h[k]) - 1| of the actual knots */
} rvs_knots;

/* Ragged batch of observed spectra of one setup and their derived products
 * (rvs_obs_prepare, rvs_basis_build).  Object i owns pixels [off[i], off[i+1])
 * of the OBJECT pools (dn, einv).  Its wavelength grid -- shared with every
 * other object observed on the same pixels, as all DESI spectra of an arm are --
 * starts at goff[i] in the GRID pools (lam, loglam, P).  The continuum basis is
 * pixel-major: value r of grid pixel p is d_P[(goff[i] + p) * npp + r], npp =
 * npoly rounded up to even (16-byte rows). */
typedef struct {
  const double *d_lam, *d_loglam; /* grid pools */
  const double *d_P;              /* grid pool, [ntot_grid][npp] */
  const int64_t *d_goff;          /* [B] */
  const double *d_dn, *d_einv;    /* object pools */
  const double *d_sumlog2;        /* [B] 2*sum ln sigma */
  const int64_t *d_off;           /* [B+1] */
  int32_t npoly;
  int32_t npp;
  int32_t nobj;
  int32_t shared_grid; /* 1: every object has the same pixels (one entry in the grid pools):
                          enables the batched GEMM form of the continuum solve */
  /* Resolution matrices (SpecData.resolution / resol_params, spec_fit.py:410-492,
   * 922-929; DESI's banded matrices, desi/desi_fit.py:723-748): the resampled template
   * T of object i is replaced by R_i T before the continuum fit.  R_i is banded with
   * the nresol diagonals d_resol_offs[] (ascending, shared by the batch), stored by
   * OUTPUT pixel in an object pool: (R_i T)[p] = sum_k d_resol[off[i]*nresol +
   * k*npix_i + p] * T[p + offs[k]], terms with p + offs[k] outside [0, npix_i) skipped.
   * resol_halfwidth = max |d_resol_offs[]| (selects the shared-memory staged RV-scan
   * kernel for narrow bands).  d_resol == NULL (nresol == 0): no resolution matrices. */
  const double *d_resol;
  const int32_t *d_resol_offs;
  int32_t nresol;
  int32_t resol_halfwidth;
} rvs_obs;

/* Regular template grid in mapped parameter space (spec_inter.py:97-132):
 * unique node coordinates of every dimension, concatenated (dimension i starts
 * at uoff[i] and has len[i] entries), and the C-order table of node ids
 * (-1 = no template at that grid point). */
#define RVS_MAX_GRID_DIM 5
typedef struct {
  const double *d_uvec;
  const int32_t *d_idgrid;
  int32_t ndim;
  int32_t len[RVS_MAX_GRID_DIM];
  int32_t uoff[RVS_MAX_GRID_DIM];
  int32_t nnode;
  const double *d_vnorm;        /* [nnode][ndim] node coordinates / ptp (the KD-tree's points) */
  double ptp[RVS_MAX_GRID_DIM]; /* peak-to-peak of every coordinate (spec_inter.py:128) */
} rvs_gridmap;

/* ---- template evaluation ------------------------------------------------- */
/* Polylinear vertex location (GridInterp.__call__, spec_inter.py:153-194) for
 * K mapped parameter vectors, coordinate i of item k at d_q[i*q_stride + k].
 * Writes the 2^ndim corner ids and weights (the d_ids / d_w of
 * rvs_template_build and rvs_chisq_fused) and d_flag[k] = 1 for the points the
 * reference resolves through its KD-tree (off-grid, missing corner) or that are
 * non-finite: those get the placeholder "row 0 alone".  With d_outside != NULL
 * the flagged finite points are resolved here as the reference does -- nearest
 * node in ptp-normalised space as a single-row item, d_outside[k] = its distance
 * (GridOutsideCheck, spec_inter.py:77-92), flag cleared; d_outside[k] = 0 for
 * points inside, NaN (flag kept) for non-finite ones. */
int rvs_locate_grid(const rvs_gridmap *gm, const double *d_q, int64_t q_stride, int K,
                    int32_t *d_ids, double *d_w, int32_t *d_flag, double *d_outside,
                    void *stream);

/* Host helpers (plain C, no device work).  rvs_knot_tables: h, hinv (n-1
 * each), cp, winv (n-2 each).  rvs_knot_info fills the scalar members of
 * rvs_knots from the host knot array and returns 0, or -2 if the knots are
 * not uniform to 1e-10 (the reference evaler's validation, spliner.c:84-96). */
void rvs_knot_tables(const double *h_x, int n, double *h_h, double *h_hinv, double *h_cp,
                     double *h_winv);
int rvs_knot_info(const double *h_x, int n, int log_step, rvs_knots *out);

/* For each item k < K:
 *   s[p]  = sum_j w[k,j] * grid[ids[k,j], p]       (fp64 accumulate)
 *   y     = log_spec ? exp(s) : s
 *   y     = vsini[k] > 0 ? rotational broadening of y : y
 *   z     = second derivatives of the natural cubic spline through (lam_t, y)
 * Output d_yz[(k*yz_stride + p)*2 + {0,1}] = y[p], z[p].
 * d_grid: fp32 (grid_f64=0) or fp64 rows of length npix_t, row stride `ld`
 * elements (ld % 4 == 0, base 16-byte aligned).  d_ids int32 [K,nvert];
 * d_w fp64 [K,nvert].  An item with ids[k,1] < 0 is a single-row item: row
 * ids[k,0] alone, exp rounded to the row's precision (the reference's off-grid
 * nearest-node value, spec_inter.py:160-167).  d_vsini may be NULL.  d_status int32[K] receives
 * RVS_ST_TEMPLATE_BAD / RVS_ST_TAPS. */
int rvs_template_build(const void *d_grid, int grid_f64, int64_t ld, const rvs_knots *knots,
                       const int32_t *d_ids, const double *d_w, int nvert,
                       const double *d_vsini, int log_spec, int K, double *d_yz,
                       int64_t yz_stride, int32_t *d_status, void *stream);

/* ---- observed-spectrum products ------------------------------------------ */
/* Object pools: dn = spec/sigma, einv = 1/sigma with sigma = sqrt(espec^2 +
 * sys^2) (sys may be 0), and sumlog2[i] = 2*sum ln sigma (the
 * 2*log(espec).sum() term of spec_fit.py:247). */
int rvs_obs_prepare(const double *d_spec, const double *d_espec, const int64_t *d_off, int B,
                    double espec_sys, double *d_dn, double *d_einv, double *d_sumlog2,
                    void *stream);

/* Grid pools of G wavelength grids: grid g owns pixels [gstart[g], gstart[g+1])
 * of d_lam (ntot = gstart[G] pixels in all).  Writes loglam = ln(lam) and the
 * continuum basis rows d_P[p*npp + r], r < npoly (zero for npoly <= r < npp).
 * rbf=1: {1,t,t^2} + Gaussian RBFs; rbf=0: Chebyshev (spec_fit.py:148-176). */
int rvs_basis_build(const double *d_lam, const int64_t *d_gstart, int G, int64_t ntot, int npoly,
                    int rbf, int npp, double *d_loglam, double *d_P, void *stream);

/* ---- chi-square ----------------------------------------------------------- */
/* For item k < K and trial j < nv:  resample template row tix[k] of d_yz at
 * velocity vels[k*nv+j] onto object oix[k]'s wavelengths, solve the continuum
 * normal equations and return
 *   chisq[k*nv+j] = 2 sum ln L_ii + sumlog2 + |D - a^T G|^2
 * and status[k*nv+j] (RVS_ST_NOT_PD, RVS_ST_RANGE).
 * Optional outputs (may be NULL; nv must be 1): d_coeffs [K,npoly]; d_raw,
 * d_model: resampled template and continuum-multiplied model written at
 * d_moff[k] + p.  fast_interp != 0: the template value at a pixel is that of the
 * first knot >= the rest wavelength (get_chisq's fast_interp switch,
 * spec_fit.py:913-918) instead of the spline value.  With resolution matrices
 * (obs->d_resol) the resampled template T of every (item, trial) is replaced by
 * R_obj T before the continuum fit (spec_fit.py:922-929); d_raw then receives R T,
 * the reference's raw_models. */
int rvs_chisq_scan(const double *d_yz, int64_t yz_stride, const int32_t *d_tix,
                   const rvs_knots *knots, const rvs_obs *obs, const int32_t *d_oix,
                   const double *d_vels, int nv, int K, double *d_chisq, int32_t *d_status,
                   double *d_coeffs, double *d_raw, double *d_model, const int64_t *d_moff,
                   int fast_interp, void *stream);
/* The same for ragged scans: item k wants its first d_nv[k] <= nv trials only (refinement
 * grids of different lengths padded to nv columns; the padding columns must hold valid
 * velocities).  The GEMM form (nv >= 4, no model output) skips the trials past an item's
 * count and writes chisq = 0, status = 0 there; d_nv == NULL: every trial. */
int rvs_chisq_scan_ragged(const double *d_yz, int64_t yz_stride, const int32_t *d_tix,
                          const rvs_knots *knots, const rvs_obs *obs, const int32_t *d_oix,
                          const double *d_vels, int nv, const int32_t *d_nv, int K,
                          double *d_chisq, int32_t *d_status, double *d_coeffs, double *d_raw,
                          double *d_model, const int64_t *d_moff, int fast_interp, void *stream);

/* Fused optimiser-phase evaluation: template build (as rvs_template_build) and
 * chi-square at ONE velocity per item without the HBM round trip of the
 * spline.  Item k uses object oix[k] (oix[k] < 0: the object
 * has no spectrum in this setup; chisq[k] = 0, status[k] = 0).  Only the part of the template that the
 * object covers at that velocity is gathered.  vsini_max: upper bound of
 * d_vsini (sizes the tap buffer; 0 if d_vsini is NULL).  d_tn: workspace
 * [K, tn_stride] doubles, tn_stride >= the longest object.  d_work: workspace
 * (16-byte aligned) of rvs_fused_workspace(K, tapcap, knots->npix_t) doubles, tapcap = ceil(vsini_max /
 * (c lnstep) + 1) + 1 (0 when d_vsini is NULL or vsini_max <= 0).  With resolution
 * matrices (obs->d_resol) d_tn holds [2K, tn_stride] doubles: the resampled template
 * goes to the second half and R T / sigma to the first.  Outputs chisq[K],
 * status[K] (template bits | RVS_ST_NOT_PD | RVS_ST_RANGE | RVS_ST_LIMIT).
 * Needs knots->ratio_dev < 1e-8 (exactly uniform or log-uniform knots) and
 * tapcap <= RVS_MAX_FUSED_TAPS, else RVS_E_LIMIT: use the general path.
 * rvs_fused_chunks: how many warps share one item.
 *
 * box (may be NULL): TMA descriptor of a DENSE regular 4-D fp32 grid -- row id =
 * C-order index of the node's grid position, no missing nodes (rvs_gridbox_init).
 * With it the 16 corner rows of a polylinear item (the 2x2x2x2 box at the grid
 * position of d_ids[k*16]) are fetched by the copy engine as two 5-D tensor
 * tiles per block of template pixels instead of 16 per-lane row gathers; results
 * are identical. */
typedef struct {
  unsigned char tmap[128]; /* CUtensorMap {ld, len[3], len[2], len[1], len[0]} fp32 */
  int32_t len[4];          /* grid lengths, slowest dimension first */
  int32_t cols;            /* template pixels per tile the descriptor was built for */
  int32_t rows;            /* corner rows per tile (4 or 8) */
} rvs_gridbox;
/* Fills *box for the grid rows at d_grid (row stride ld floats, ld a multiple of
 * 4, prod(len) rows).  RVS_E_ARG unless ndim == 4; RVS_E_CUDA if the driver
 * refuses the descriptor. */
int rvs_gridbox_init(rvs_gridbox *box, const void *d_grid, int64_t ld, int ndim,
                     const int32_t *len);
int rvs_fused_chunks(int npix_t, int tapcap);
int64_t rvs_fused_workspace(int K, int tapcap, int npix_t);
int rvs_chisq_fused(const void *d_grid, int grid_f64, int64_t ld, const rvs_knots *knots,
                    const int32_t *d_ids, const double *d_w, int nvert, const double *d_vsini,
                    double vsini_max, int log_spec, const rvs_obs *obs, const int32_t *d_oix,
                    const double *d_vels, int K, double *d_tn, int64_t tn_stride, double *d_work,
                    double *d_chisq, int32_t *d_status, const rvs_gridbox *box, void *stream);

/* One arm (spectral setup) of a multi-arm evaluation call: the per-arm arguments of
 * rvs_chisq_fused. */
typedef struct {
  const void *d_grid;
  int32_t grid_f64, log_spec;
  int64_t ld;
  const rvs_knots *knots;
  const rvs_obs *obs;
  const int32_t *d_oix;
  double *d_tn;
  int64_t tn_stride;
  double *d_work;
  double *d_chisq;
  int32_t *d_status;
  const rvs_gridbox *box;
} rvs_fused_arm;
/* rvs_chisq_fused for the narm (<= 4) arms of a call whose items share vertex ids and
 * weights (banks with one node table, located once), rotation and velocity: when the arms
 * run the same kernel instantiations (same grid kind, npoly, all on shared pixel grids, no
 * resolution matrices) every kernel of the call is launched ONCE for all arms -- 6 launches
 * per call instead of 5 per arm, the small latency-bound kernels paid once; otherwise the
 * arms are enqueued one after the other.  Values are identical to per-arm calls. */
int rvs_chisq_fused_multi(const rvs_fused_arm *arms, int narm, const int32_t *d_ids,
                          const double *d_w, int nvert, const double *d_vsini, double vsini_max,
                          const double *d_vels, int K, void *stream);

/* RV-grid statistics of find_best for S scans: scan s has nv velocities
 * vels[s*nv..] and chi-squares chisq[(s*npar+q)*nv + j] for npar templates.
 * out[s*8..] = best_chi, best_vel, vel_err, skewness, kurtosis, i_vel, i_par, flags
 * (bit 0: the parabola vertex is not strictly inside its bracket -- the reference's
 * assertion spec_fit.py:1014 would fail; bit 1: a chi-square is NaN); probs (may be NULL)
 * [S,nv].  rvs_scan_stats_ragged: scan s uses only its first d_nv[s] velocities; rows keep
 * the stride nv_stride (refinement scans of many objects, vel_fit.py:358-439). */
int rvs_scan_stats(const double *d_vels, const double *d_chisq, int S, int npar, int nv,
                   int quadratic, double *d_out, double *d_probs, void *stream);
int rvs_scan_stats_ragged(const double *d_vels, const double *d_chisq, int S, int npar,
                          int nv_stride, const int32_t *d_nv, int quadratic, double *d_out,
                          double *d_probs, void *stream);

/* ---- cross-correlation first guess (fitter_ccf.py:126-232) ---------------- */
/* One arm's CCF template bank and the lag -> velocity-grid table.  d_fft,
 * d_fft2: rfft of the preprocessed models and of their squares, complex128
 * interleaved [ntempl][npoints/2+1] (CCFCache.ccfs / ccf2s, fitter_ccf.py:51-56).
 * Velocity-grid point j is interpolated linearly between the CCF pixels lo[j]
 * and hi[j] (indices into the length-npoints inverse transform, i.e. the
 * reference's subind[idx-1], subind[idx], fitter_ccf.py:132-150,205):
 *   value = (y_hi - y_lo) / dx[j] * dxn[j] + y_lo,
 * dx = vels[hi]-vels[lo], dxn = vel_grid[j]-vels[lo] (scipy interp1d).
 * continuum=1: chi2 = -2 ccf0 + ccf1; 0: -ccf0^2/ccf1 (fitter_ccf.py:198-201). */
typedef struct {
  const double *d_fft, *d_fft2;
  const int32_t *d_lo, *d_hi; /* [nvel] */
  const double *d_dxn, *d_dx; /* [nvel] */
  int32_t npoints, ntempl, continuum, nvel;
} rvs_ccf_arm;

/* Bytes of workspace that let rvs_ccf_accumulate process nb objects per pass. */
int64_t rvs_ccf_workspace(const rvs_ccf_arm *arm, int nb);

/* Adds one arm's contribution for B objects.  d_pspec, d_pivar: [B][npoints]
 * preprocessed spectrum and inverse variance (make_ccf.preprocess_data).
 * Object b accumulates into row d_row[b] (NULL: b) of d_chisq
 * [nrow][ntempl][nvel] and d_sse[nrow] (+= sum pspec^2 pivar); the caller
 * zeroes both before the first arm.  d_work: 256-byte aligned workspace of
 * work_bytes (>= rvs_ccf_workspace(arm, 1)); objects are processed in as
 * large passes as it allows. */
int rvs_ccf_accumulate(const rvs_ccf_arm *arm, const double *d_pspec, const double *d_pivar,
                       int B, const int32_t *d_row, double *d_chisq, double *d_sse, void *d_work,
                       int64_t work_bytes, void *stream);

/* Per row: all_chisqs = chisq + sse; best template = argmin_t min_v, best pixel,
 * parabola vertex if it opens upwards (fitter_ccf.py:206-222).
 * out[row*8..] = best_id, best_pix, best_vel, best value, finite flag, 0,0,0;
 * d_best_ccf (may be NULL) [nrow][nvel] = all_chisqs[best_id]. */
int rvs_ccf_best(const double *d_chisq, const double *d_sse, const double *d_velgrid, int nrow,
                 int ntempl, int nvel, double *d_out, double *d_best_ccf, void *stream);

/* Data preparation in front of the CCF, on the device: make_ccf.preprocess_data
 * (reference make_ccf.py:330-414, with interp_masker :287-327 and get_continuum
 * :105-164) for n spectra observed on ONE pixel grid d_lam[npix] (increasing).
 * d_spec, d_espec: [n][npix]; d_bad: [n][npix] bytes or NULL.  Host-built tables of the
 * pixel grid (rvspecfit_b200/make_ccf.py DevicePrep): d_basis[npix][nn] = value at every
 * pixel of the quadratic interpolating spline through unit node values (the continuum is
 * exp(d_basis . p)), d_bin[nn+1] = pixel ranges whose medians start the node values,
 * d_left[npoints] / d_wr[npoints] = left bracketing pixel (-1: outside the spectrum) and
 * right-hand weight of every CCF pixel.  continuum = 0 skips the continuum (ccfconf
 * without splinestep).  Outputs d_pspec, d_pivar [n][npoints] (what rvs_ccf_accumulate
 * takes), optionally d_cont [n][npix] and d_info [n][2] (solver iterations, 1 = every
 * pixel was masked).  rvs_ccf_prep_smem: bytes of shared memory a spectrum of npix pixels
 * needs (RVS_E_LIMIT above ~220 kB: preprocess such spectra on the host). */
int64_t rvs_ccf_prep_smem(int npix, int nn);
int rvs_ccf_preprocess(const double *d_lam, const double *d_spec, const double *d_espec,
                       const uint8_t *d_bad, int n, int npix, const double *d_basis, int nn,
                       const int32_t *d_bin, const int32_t *d_left, const double *d_wr,
                       int npoints, int continuum, double maxerr, double *d_pspec,
                       double *d_pivar, double *d_cont, int32_t *d_info, void *stream);

/* ---- host-side lock-step Nelder-Mead stepper (nm_host.cpp; no device work) ----
 * The optimiser loop of vel_fit.process (reference vel_fit.py:628-650:
 * scipy.optimize.minimize(method='Nelder-Mead', options={initial_simplex, xatol, fatol,
 * maxiter, maxfev=inf})) for B simplices at once, scipy's decision rules and
 * floating-point expressions.  Protocol: n = rvs_nm_request(...) writes the problem
 * index and the coordinates of the next n trial points (n > cap: nothing beyond cap was
 * written, call again with larger buffers is NOT possible -- size them B * max(N + 1, 4));
 * the caller evaluates them and passes the n values to rvs_nm_feed; n == 0 means every
 * problem has stopped and rvs_nm_result may be read.  With at most `speculate_below`
 * live problems a round asks for the reflection AND the three candidate second points
 * of every problem (one round per iteration instead of two; each problem still uses
 * exactly the values scipy would have computed). */
void *rvs_nm_create(int B, int N, const double *h_sims /* [B][N+1][N] */, double xatol,
                    double fatol, int64_t maxiter);
void rvs_nm_destroy(void *nm);
int64_t rvs_nm_request(void *nm, int speculate_below, int32_t *h_idx, double *h_X, int64_t cap);
int rvs_nm_feed(void *nm, const double *h_f, int64_t n);
/* Problems still iterating (h_active [B], may be NULL, receives their flags).  The rows
 * rvs_nm_result gives for a stopped problem are final while the others go on, so a driver
 * may hand finished problems to the next stage of the fit without waiting for the rest. */
int64_t rvs_nm_live(void *nm, uint8_t *h_active);
int rvs_nm_result(void *nm, double *h_x /* [B][N] */, double *h_fun, uint8_t *h_success,
                  double *h_final_simplex, int64_t *h_nit, int64_t *h_nfev);

/* ---- host side of an optimiser-phase evaluation call of the batched fit (fit_host.cpp) --
 * Layout of the fitted vector and the per-object constants of a batch_fit.BatchObjective:
 * vector = [vel, (vsini if fit_vsini), free atmospheric parameters in grid order]
 * (reference vel_fit.py:119-207 ParamMapper). */
typedef struct {
  int32_t nfit;      /* length of the fitted vector */
  int32_t nspec;     /* atmospheric parameters of the grid */
  int32_t fit_vsini; /* 1: vector[1] is vsini (clipped to [0, max_vsini], quadratic penalty) */
  int32_t has_vsini; /* 1 (and !fit_vsini): vsini fixed at h_vsini0[object] */
  int32_t fixmask;   /* bit j: parameter j is fixed at h_p0[object][j] */
  int32_t logmask;   /* bit j: parameter j enters the grid as log10 (read_grid.py:127-145) */
  int32_t priormask; /* bit j: Gaussian prior (h_prior_mu[j], h_prior_sig[j]) on parameter j */
  int32_t narm, nobj;
  int32_t pad_;
  double min_vel, max_vel, max_vsini;
  const double *h_p0;        /* [nobj][nspec] start / fixed values */
  const double *h_q0;        /* [nobj][nspec] the same, grid-mapped (log10 where logmask) */
  const double *h_vsini0;    /* [nobj] or NULL */
  const double *h_prior_mu;  /* [nspec] */
  const double *h_prior_sig; /* [nspec] */
  const int32_t *h_oix;      /* [narm][nobj] row of the object in each arm's batch, -1 absent */
  const double *h_badchi;    /* [nobj] 10 x total pixels (spec_fit.py:863) */
  const uint8_t *h_cover;    /* [nobj] template covers the object over [min_vel, max_vel] */
} rvs_fit_layout;
/* K fitted vectors h_X[K][nfit] of objects h_obj[K] -> rows (vel, vsini, q_0..) of
 * h_in[2+nspec][Kp] and arm rows h_oix[narm][Kp] (the call's pinned upload buffers; items
 * K..Kp-1 are padding, absent on every arm), the additive host terms h_prior[K], h_pen[K],
 * and h_wall[K] = 1 for vectors behind the hard walls (vel outside [min_vel, max_vel] or a
 * non-finite parameter: not evaluated, value 1e30, vel_fit.py:252-254).  h_logvals
 * [nlog][K]: log10 of the FITTED log-mapped parameters, in parameter order, computed by
 * the caller (numpy's log10, so that every path of the package maps identically).
 * *h_vsini_max = largest vsini of an evaluated item. */
int rvs_fit_pack(const rvs_fit_layout *L, int64_t K, int64_t Kp, const int32_t *h_obj,
                 const double *h_X, const double *h_logvals, double *h_in, int32_t *h_oix,
                 double *h_prior, double *h_pen, uint8_t *h_wall, double *h_vsini_max);
/* The call's downloads h_chi[2][narm][Kp] (chi-square | off-grid measure) and
 * h_flags[2][narm][Kp] (evaluation | location status) -> objective values h_out[K] =
 * prior + sum over arms (chi-square + off-grid penalty) + penalty, 1e30 behind the walls;
 * h_redo[K] = 1 where the fused path could not settle the item (any flag, a non-finite
 * value, template not covering the object): the caller re-evaluates those through the
 * general path.  Returns the number of such items (negative: RVS_E_*). */
int64_t rvs_fit_collect(const rvs_fit_layout *L, int64_t K, int64_t Kp, const int32_t *h_obj,
                        const double *h_in, const double *h_chi, const int32_t *h_flags,
                        int shared_locate, int outside_penalty, const double *h_prior,
                        const double *h_pen, const uint8_t *h_wall, double *h_out,
                        uint8_t *h_redo);

/* ---- native round loop of a lock-step Nelder-Mead stage (drive_host.cpp) -------------
 * The optimiser loop of vel_fit.process (reference vel_fit.py:628-650) asks for one
 * objective value after the other; the batched fit asks for one evaluation CALL after the
 * other, and between two calls sits the host: stepper, packing, launch, wait, reduction.
 * rvs_nm_drive runs those rounds without the interpreter on the slot the caller holds
 * (stream, pinned staging, captured graphs of the call's launch configurations) and
 * returns only for what the caller alone can do (the RVS_DRIVE_* codes). */
#define RVS_DRIVE_DONE 0      /* every problem has stopped */
#define RVS_DRIVE_PEEL 1      /* >= stop_stopped problems have stopped, others go on */
#define RVS_DRIVE_LAUNCH 2    /* inputs packed (K, Kp, vmax set); no captured graph for this
                                 configuration: launch it, set state = RVS_DRIVE_LAUNCHED */
#define RVS_DRIVE_REDO 3      /* f_out collected, f_redo marks items for the general path:
                                 patch f_out, call again */
#define RVS_DRIVE_PYEVAL 4    /* the call does not fit the fused path: evaluate the request
                                 (rvs_drive_request), write f_out, state = RVS_DRIVE_COLLECTED */
#define RVS_DRIVE_IDLE 0
#define RVS_DRIVE_PACKED 1
#define RVS_DRIVE_LAUNCHED 2
#define RVS_DRIVE_COLLECTED 3
typedef struct {
  /* the evaluation slot the caller holds for the stage */
  void *stream;                 /* cudaStream_t */
  double *h_in;                 /* pinned [2+nspec][Kp] */
  int32_t *h_oix;               /* pinned [narm][Kp] */
  const double *h_chi;          /* pinned [2][narm][Kp] */
  const int32_t *h_flags;       /* pinned [2][narm][Kp] */
  double *f_prior, *f_pen;      /* [cap] */
  uint8_t *f_wall;
  double *f_out;
  uint8_t *f_redo;
  int64_t cap;                  /* items the buffers above hold */
  int32_t shared_locate;        /* one vertex location for all arms (row 0 stands for all) */
  int32_t ngraph;               /* captured graphs of the slot: configuration -> executable */
  const int64_t *g_kp;
  const double *g_vmax;
  void *const *g_exec;          /* cudaGraphExec_t */
  const int32_t *g_nk;          /* kernels in the graph (launch accounting) */
  /* the stage */
  const int32_t *objmap;        /* stepper problem -> object of the engine */
  int64_t nprob;
  int32_t speculate_below;
  int32_t state;                /* RVS_DRIVE_IDLE ... (in/out) */
  int64_t stop_stopped;         /* 0: never return RVS_DRIVE_PEEL */
  double fused_vmax;            /* largest vsini the fused path takes */
  /* the round in progress (out) */
  int64_t K, Kp;
  double vmax;
  /* counters (accumulated) */
  int64_t rounds, items, graph_launches, graph_kernels, h2d_bytes, d2h_bytes;
  /* optional timing: (t0, t1, K) of every graph launch, ms since epoch_event */
  void *epoch_event;            /* cudaEvent_t recorded by the caller, NULL: no timing */
  double *t_rec;
  int64_t t_cap, t_n;
  int32_t timed, pad_;
} rvs_drive;
/* A non-blocking stream owned by the caller (cudaStream_t; NULL on failure).
 * high_priority: 0 = least priority (the default stream's), 1 = halfway, 2 = greatest. */
void *rvs_stream_create(int high_priority);
void rvs_stream_destroy(void *stream);
/* Item count of an evaluation call rounded up to a launch configuration. */
int64_t rvs_fit_round_items(int64_t K);
void *rvs_drive_create(int64_t cap /* most points of a request */, int nfit);
void rvs_drive_destroy(void *drive);
/* The request in progress: stepper problems, fitted vectors [K][nfit], objects. */
int rvs_drive_request(void *drive, const int32_t **idx, const double **X, const int32_t **obj);
int rvs_nm_drive(void *nm, void *drive, const rvs_fit_layout *lay, rvs_drive *io);
/* The same loop around the BFGS stepper below (speculate_below is ignored). */
int rvs_bfgs_drive(void *bfgs, void *drive, const rvs_fit_layout *lay, rvs_drive *io);

/* ---- host-side lock-step BFGS stepper (bfgs_host.cpp; no device work) ----------------
 * The polish step of vel_fit.process (reference vel_fit.py:653-658:
 * scipy.optimize.minimize(method='BFGS', options=dict(hess_inv0=...)), forward-difference
 * gradient) for B problems at once: scipy's `_minimize_bfgs` with the MINPACK line search
 * DCSRCH and its fallback line_search_wolfe2, same constants and decision rules; matrix
 * products summed in index order (scipy's go through BLAS: values agree to rounding).
 * Protocol as rvs_nm_*: a request is N + 1 consecutive points per searching problem (a point
 * and its forward-difference neighbours); buffers of B * (N + 1) points always suffice.
 * status: scipy's warnflag (0 converged, 1 maxiter, 2 precision loss, 3 NaN). */
void *rvs_bfgs_create(int B, int N, const double *h_x0 /* [B][N] */,
                      const double *h_hess_inv0 /* [N][N] or NULL */, double gtol,
                      int64_t maxiter /* <= 0: 200 N */);
void rvs_bfgs_destroy(void *bfgs);
int64_t rvs_bfgs_request(void *bfgs, int32_t *h_idx, double *h_X, int64_t cap);
int rvs_bfgs_feed(void *bfgs, const double *h_f, int64_t n);
int64_t rvs_bfgs_live(void *bfgs, uint8_t *h_active);
int rvs_bfgs_result(void *bfgs, double *h_x, double *h_fun, int64_t *h_nit, int32_t *h_status,
                    int64_t *h_rounds);

#ifdef __cplusplus
}
#endif
#endif

// Data preparation in front of the CCF kernels, on the device: the reference's
// make_ccf.preprocess_data (make_ccf.py:330-414) with its helpers interp_masker
// (:287-327) and get_continuum (:105-164) for a batch of spectra observed on ONE
// pixel grid.  One CTA per spectrum; everything a spectrum needs lives in the CTA's
// shared memory, so HBM sees the inputs once (24 B per pixel) and the outputs once
// (16 B per CCF pixel).
//
//   1. smooth = medfilt(flux, 11) enters only through (smooth <= 0): the median of
//      an 11-pixel window (zero padded) is <= 0 iff at least 6 of its values are.
//   2. typical_err = nanmedian(err): bitonic sort in shared memory.
//   3. bad |= err > maxerr typical_err | smooth <= 0;  err[bad] = 1e9 typical_err.
//   4. interp_masker: masked pixels bridged linearly between the nearest good
//      neighbours (max / min scans give the neighbours), flat beyond the ends.
//   5. get_continuum: exp of a quadratic interpolating spline through nn nodes, fitted
//      with the soft-L1 loss rho(z) = 2 (sqrt(1 + z) - 1).  The spline is linear in its
//      node values, so its value at every pixel is a row of a cardinal-basis matrix
//      Cb[npix][nn] the host computes once per pixel grid with the reference's own
//      FITPACK routine; start values are the per-bin medians as in the reference.
//      The reference minimises with scipy.optimize.least_squares (trust-region
//      reflective, 2-point Jacobian, ftol = xtol = gtol = 1e-8); here a damped
//      Gauss-Newton iteration with the analytic Jacobian and the same robust
//      rescaling of residuals and Jacobian (scipy _lsq/common.py
//      scale_for_robust_loss_function) runs to a tighter stop.  Both land in the same
//      minimum; they agree to the accuracy the reference's own stop leaves in the
//      continuum (~1e-5 relative; tests/test_gpu_ccf.py states the tolerance).
//   6. normalisation, inverse-variance propagation and the linear resampling onto the
//      CCF pixels (make_ccf.py:388-412), with the bracketing pixel and weight of every
//      CCF pixel from a host-built table.
#include "common.cuh"

namespace rvs {

#define CP_THREADS 256
#define CP_MAXNODES 24

struct CcfPrepArgs {
  const double *lam;        // [npix]
  const double *spec, *espec;  // [n][npix]
  const uint8_t *bad;       // [n][npix] or NULL
  const double *Cb;         // [npix][nn]
  const int32_t *bin;       // [nn + 1] pixel ranges of the start-value bins
  const int32_t *left;      // [npoints] left bracketing pixel, -1: outside
  const double *wr;         // [npoints] weight of the right pixel
  double *pspec, *pivar;    // [n][npoints]
  double *cont;             // [n][npix] or NULL (diagnostic output)
  int32_t *info;            // [n][2]: iterations, flags
  int npix, npix2, nn, npoints, continuum;
  double maxerr;
};

__device__ __forceinline__ void bitonic_sort(double *s, int n2) {
  for (int k = 2; k <= n2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const double a = s[i], b = s[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s[i] = b; s[l] = a; }
        }
      }
      __syncthreads();
    }
}

// median of the m smallest entries of the sorted buffer (numpy: mean of the two middle ones)
__device__ __forceinline__ double sorted_median(const double *s, int m) {
  if (m <= 0) return nan("");
  return (m & 1) ? s[m >> 1] : (s[(m >> 1) - 1] + s[m >> 1]) / 2.0;
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Cholesky solve of the nn x nn system A x = b by one thread (nn <= 24; runs a few dozen
// times per spectrum).  Returns false if A is not positive definite.
__device__ bool chol_solve(double *A, double *b, int n) {
  for (int j = 0; j < n; j++) {
    double d = A[j * n + j];
    for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0)) return false;
    d = sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double v = A[i * n + j];
      for (int k = 0; k < j; k++) v -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = v / d;
    }
  }
  for (int i = 0; i < n; i++) {
    double v = b[i];
    for (int k = 0; k < i; k++) v -= A[i * n + k] * b[k];
    b[i] = v / A[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double v = b[i];
    for (int k = i + 1; k < n; k++) v -= A[k * n + i] * b[k];
    b[i] = v / A[i * n + i];
  }
  return true;
}

// block-wide sum in a fixed order (warp trees, then warp 0 over the warp sums)
__device__ double block_sum(double v, double *red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
  for (int w = 0; w < CP_THREADS / 32; w++) t += red[w];
  return t;
}

__global__ void __launch_bounds__(CP_THREADS) ccf_prep_kernel(CcfPrepArgs a) {
  extern __shared__ double sm[];
  const int npix = a.npix, nn = a.nn, tid = threadIdx.x;
  double *flux = sm;                 // [npix]  bridged flux
  double *err = flux + npix;         // [npix]
  double *buf = err + npix;          // [npix2] sort buffer; later per-pixel weights
  double *buf2 = buf + a.npix2;      // [npix]  second per-pixel array of the solver
  double *mat = buf2 + npix;         // [MAXN^2] normal matrix
  double *g = mat + CP_MAXNODES * CP_MAXNODES;   // [MAXN] gradient
  double *p = g + CP_MAXNODES;       // [MAXN] node values (log continuum)
  double *ptry = p + CP_MAXNODES;    // [MAXN] trial point
  double *Awork = ptry + CP_MAXNODES;            // [MAXN^2 + MAXN] damped copy + step
  double *red = Awork + CP_MAXNODES * CP_MAXNODES + CP_MAXNODES;   // [16] reduction scratch
  __shared__ int s_prev[CP_THREADS], s_next[CP_THREADS];
  __shared__ int s_flag;
  uint8_t *badm = reinterpret_cast<uint8_t *>(red + 16);   // [npix]
  const int64_t row = (int64_t)blockIdx.x * npix;
  const double *spec0 = a.spec + row;

  // ---- 1-3: typical error, masks
  int nfin = 0;
  for (int i = tid; i < a.npix2; i += CP_THREADS) {
    double e = i < npix ? a.espec[row + i] : INFINITY;
    if (e != e) e = INFINITY;           // nanmedian ignores NaNs
    buf[i] = e;
  }
  for (int i = tid; i < npix; i += CP_THREADS) {
    flux[i] = spec0[i];
    const double e = a.espec[row + i];
    err[i] = e;
    nfin += (e == e);
  }
  __syncthreads();
  bitonic_sort(buf, a.npix2);
  const int nvalid = (int)(block_sum((double)nfin, red) + 0.5);
  const double typ = sorted_median(buf, nvalid);
  __syncthreads();
  for (int i = tid; i < npix; i += CP_THREADS) {
    bool b = a.bad ? a.bad[row + i] != 0 : false;
    if (a.continuum) {
      int npos = 0;       // window values <= 0 (zero padding counts)
      for (int d = -5; d <= 5; d++) {
        const int j = i + d;
        npos += (j < 0 || j >= npix) ? 1 : (spec0[j] <= 0);
      }
      b = b || (err[i] > a.maxerr * typ) || npos >= 6;
    }
    badm[i] = b;
  }
  __syncthreads();
  for (int i = tid; i < npix; i += CP_THREADS)
    if (badm[i]) err[i] = 1e9 * typ;
  // ---- 4: bridge masked pixels (each thread owns a contiguous segment)
  {
    const int seg = (npix + CP_THREADS - 1) / CP_THREADS;
    const int i0 = min(npix, tid * seg), i1 = min(npix, i0 + seg);
    int last = -1, first = npix;
    for (int i = i0; i < i1; i++)
      if (!badm[i]) { last = i; if (first == npix) first = i; }
    s_prev[tid] = last;      // last good pixel of the segment
    s_next[tid] = first;     // first good pixel of the segment
    __syncthreads();
    int pg = -1, ng = npix;  // nearest good pixel before / after the segment
    for (int t = tid - 1; t >= 0; t--) if (s_prev[t] >= 0) { pg = s_prev[t]; break; }
    for (int t = tid + 1; t < CP_THREADS; t++) if (s_next[t] < npix) { ng = s_next[t]; break; }
    if (tid == 0) s_flag = 0;
    __syncthreads();
    // per pixel: previous good = running value inside the segment; next good by a
    // backwards pass
    int nxt = ng;
    for (int i = i1 - 1; i >= i0; i--) {
      if (!badm[i]) nxt = i;
      else buf2[i] = (double)nxt;          // stash the next good index
    }
    int prv = pg;
    for (int i = i0; i < i1; i++) {
      if (!badm[i]) { prv = i; continue; }
      const int ia = prv, ib = (int)buf2[i];
      double v;
      if (ia < 0 && ib >= npix) {          // nothing good at all
        v = spec0[i];
        if (!isfinite(v)) v = 1;
        s_flag = 1;
      } else if (ia < 0) v = spec0[ib];
      else if (ib >= npix) v = spec0[ia];
      else {
        const double la = a.lam[ia], lb = a.lam[ib], l0 = a.lam[i];
        v = (-(la - l0) * spec0[ib] + (lb - l0) * spec0[ia]) / (lb - la);
      }
      flux[i] = v;
    }
  }
  __syncthreads();
  // ---- median of the bridged flux
  for (int i = tid; i < a.npix2; i += CP_THREADS) buf[i] = i < npix ? flux[i] : INFINITY;
  __syncthreads();
  bitonic_sort(buf, a.npix2);
  const double med = sorted_median(buf, npix);
  __syncthreads();
  int iters = 0;
  if (a.continuum) {
    // ---- 5a: start values = log of the per-bin medians (make_ccf.py:139-150)
    double medc = med;
    if (medc <= 0) medc = fabs(medc) > 0 ? fabs(medc) : 1;
    for (int k = 0; k < nn; k++) {
      const int b0 = a.bin[k], b1 = a.bin[k + 1], m = b1 - b0;
      const int m2 = next_pow2(max(m, 1));
      for (int i = tid; i < m2; i += CP_THREADS) buf[i] = i < m ? flux[b0 + i] : INFINITY;
      __syncthreads();
      bitonic_sort(buf, m2);
      if (tid == 0) {
        double v = log(fmax(sorted_median(buf, m), 1e-3 * medc));
        if (!isfinite(v)) v = log(medc);
        p[k] = v;
      }
      __syncthreads();
    }
    // ---- 5b: damped Gauss-Newton on the soft-L1 cost
    auto cost_of = [&](const double *q) -> double {
      double c = 0;
      for (int i = tid; i < npix; i += CP_THREADS) {
        double lc = 0;
        const double *cb = a.Cb + (int64_t)i * nn;
        for (int k = 0; k < nn; k++) lc += cb[k] * q[k];
        const double f = (exp(fmin(fmax(lc, -100.0), 100.0)) - flux[i]) / err[i];
        c += 2 * (sqrt(1 + f * f) - 1);
      }
      return 0.5 * block_sum(c, red);
    };
    double cost = cost_of(p);
    double lambda = 1e-3;
    const int nent = nn * (nn + 1) / 2;
    for (iters = 0; iters < 100; iters++) {
      // per-pixel scaled Jacobian factor and residual (scale_for_robust_loss_function)
      for (int i = tid; i < npix; i += CP_THREADS) {
        double lc = 0;
        const double *cb = a.Cb + (int64_t)i * nn;
        for (int k = 0; k < nn; k++) lc += cb[k] * p[k];
        const bool clipped = lc < -100.0 || lc > 100.0;
        const double m = exp(fmin(fmax(lc, -100.0), 100.0));
        const double f = (m - flux[i]) / err[i];
        const double z = f * f, r1 = 1 / sqrt(1 + z), r2 = -0.5 * r1 * r1 * r1;
        double js = r1 + 2 * r2 * z;
        if (js < 2.220446049250313e-16) js = 2.220446049250313e-16;
        js = sqrt(js);
        const double fs = f * r1 / js;
        const double dj = clipped ? 0.0 : js * m / err[i];   // d f_scaled / d (log cont)
        buf[i] = dj * dj;
        buf2[i] = dj * fs;
      }
      __syncthreads();
      // normal equations: thread t owns one entry (fixed summation order over pixels)
      for (int e = tid; e < nent + nn; e += CP_THREADS) {
        double acc = 0;
        if (e < nent) {
          int r = 0, rem = e;
          while (rem > r) { rem -= r + 1; r++; }
          const int c = rem;
          for (int i = 0; i < npix; i++)
            acc += buf[i] * a.Cb[(int64_t)i * nn + r] * a.Cb[(int64_t)i * nn + c];
          mat[r * nn + c] = acc;
          mat[c * nn + r] = acc;
        } else {
          const int r = e - nent;
          for (int i = 0; i < npix; i++) acc += buf2[i] * a.Cb[(int64_t)i * nn + r];
          g[r] = acc;
        }
      }
      __syncthreads();
      bool accepted = false, converged = false;
      for (int attempt = 0; attempt < 12 && !accepted; attempt++) {
        double *A = Awork;
        if (tid == 0) {
          for (int r = 0; r < nn; r++)
            for (int c = 0; c < nn; c++)
              A[r * nn + c] = mat[r * nn + c] * (r == c ? 1 + lambda : 1.0);
          double *st = A + nn * nn;
          for (int r = 0; r < nn; r++) st[r] = -g[r];
          const bool ok = chol_solve(A, st, nn);
          for (int r = 0; r < nn; r++) ptry[r] = ok ? p[r] + st[r] : p[r];
          s_flag = (s_flag & 1) | (ok ? 0 : 2);
        }
        __syncthreads();
        const bool ok = (s_flag & 2) == 0;
        double cnew = ok ? cost_of(ptry) : INFINITY;
        if (ok && cnew <= cost) {
          double dmax = 0;
          for (int r = 0; r < nn; r++) dmax = fmax(dmax, fabs(ptry[r] - p[r]));
          converged = (cost - cnew) <= 1e-15 * cost || dmax < 1e-13;
          __syncthreads();
          if (tid == 0) for (int r = 0; r < nn; r++) p[r] = ptry[r];
          cost = cnew;
          lambda = fmax(lambda * 0.2, 1e-12);
          accepted = true;
        } else {
          lambda *= 8;
        }
        __syncthreads();
      }
      if (!accepted || converged) break;
    }
  }
  // ---- 5c/6: continuum, normalisation, inverse variance (make_ccf.py:384-396)
  for (int i = tid; i < npix; i += CP_THREADS) {
    double c = 1;
    if (a.continuum) {
      double lc = 0;
      const double *cb = a.Cb + (int64_t)i * nn;
      for (int k = 0; k < nn; k++) lc += cb[k] * p[k];
      c = exp(fmin(fmax(lc, -100.0), 100.0));
    }
    c = med > 0 ? fmax(1e-2 * med, c) : fmax(c, 1.0);
    if (a.cont) a.cont[row + i] = c;
    const bool b = badm[i];
    const double iv = b ? 0.0 : 1.0 / (err[i] * err[i]);
    buf[i] = b ? 0.0 : spec0[i] / c;      // norm
    buf2[i] = c * c * iv;                  // ivar
  }
  __syncthreads();
  const int64_t orow = (int64_t)blockIdx.x * a.npoints;
  for (int gp = tid; gp < a.npoints; gp += CP_THREADS) {
    const int li = a.left[gp];
    double os = 0, oi = 0;
    if (li >= 0) {
      const double wr = a.wr[gp], wl = 1 - wr;
      os = wl * buf[li] + wr * buf[li + 1];
      const double il = buf2[li], ir = buf2[li + 1];
      oi = il * ir / (wl * wl * ir + wr * wr * il + ((il * ir) == 0 ? 1.0 : 0.0));
    }
    a.pspec[orow + gp] = os;
    a.pivar[orow + gp] = oi;
  }
  if (tid == 0 && a.info) {
    a.info[2 * blockIdx.x] = iters;
    a.info[2 * blockIdx.x + 1] = s_flag & 1;
  }
}

}  // namespace rvs

extern "C" int64_t rvs_ccf_prep_smem(int npix, int nn) {
  int npix2 = 1;
  while (npix2 < npix) npix2 <<= 1;
  (void)nn;
  const int64_t doubles = 3 * (int64_t)npix + npix2 + 2 * CP_MAXNODES * CP_MAXNODES +
                          4 * CP_MAXNODES + 16 + (npix + 7) / 8 + 1;
  return doubles * 8;
}

extern "C" int rvs_ccf_preprocess(const double *d_lam, const double *d_spec, const double *d_espec,
                                  const uint8_t *d_bad, int n, int npix, const double *d_basis,
                                  int nn, const int32_t *d_bin, const int32_t *d_left,
                                  const double *d_wr, int npoints, int continuum, double maxerr,
                                  double *d_pspec, double *d_pivar, double *d_cont,
                                  int32_t *d_info, void *stream) {
  using namespace rvs;
  if (n == 0) return 0;
  RVS_REQUIRE(d_lam && d_spec && d_espec && d_left && d_wr && d_pspec && d_pivar, RVS_E_ARG,
              "rvs_ccf_preprocess: null pointer");
  RVS_REQUIRE(npix >= 12 && npoints > 0, RVS_E_ARG, "rvs_ccf_preprocess: npix=%d", npix);
  RVS_REQUIRE(!continuum || (d_basis && d_bin && nn >= 1 && nn <= CP_MAXNODES), RVS_E_LIMIT,
              "rvs_ccf_preprocess: %d continuum nodes (limit %d)", nn, CP_MAXNODES);
  const int64_t smem = rvs_ccf_prep_smem(npix, continuum ? nn : 0);
  RVS_REQUIRE(smem <= 220 * 1024, RVS_E_LIMIT,
              "rvs_ccf_preprocess: %d pixels need %lld B of shared memory", npix, (long long)smem);
  RVS_CUDA_OK(ensure_dyn_smem(ccf_prep_kernel, (size_t)smem));
  CcfPrepArgs a;
  a.lam = d_lam; a.spec = d_spec; a.espec = d_espec; a.bad = d_bad; a.Cb = d_basis; a.bin = d_bin;
  a.left = d_left; a.wr = d_wr; a.pspec = d_pspec; a.pivar = d_pivar; a.cont = d_cont;
  a.info = d_info; a.npix = npix; a.nn = continuum ? nn : 0; a.npoints = npoints;
  a.continuum = continuum; a.maxerr = maxerr;
  a.npix2 = 1;
  while (a.npix2 < npix) a.npix2 <<= 1;
  ccf_prep_kernel<<<n, CP_THREADS, smem, (cudaStream_t)stream>>>(a);
  RVS_LAUNCH_OK();
  return 0;
}

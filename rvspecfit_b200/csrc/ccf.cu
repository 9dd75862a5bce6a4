// Cross-correlation first guess (reference fitter_ccf.py:126-232), batched over
// objects.  cuFFT does the transforms (the only library primitive on the path);
// everything around them is hand-written:
//   ccf_prep_kernel    S/E^2 products and the sum S^2/E^2 of every object
//   ccf_mult_kernel    Z[b,t,:] = (-2 That[t,:] conj(Shat[b,:]) + That2[t,:] conj(Ihat[b,:])) / n
//                      (continuum mode: by linearity ONE inverse transform per
//                      (object, template) instead of the reference's two;
//                      ratio mode -ccf0^2/ccf1 keeps both)
//   ccf_gather_kernel  lag window -> linear interpolation onto the common
//                      velocity grid (scipy interp1d arithmetic) -> += over arms
//   ccf_best_kernel    + total_sse, argmin over templates and velocities,
//                      parabola vertex
// The inverse real transform of n points is taken as ONE complex transform of n/2 points:
// with A[k] = X[k] + conj(X[n/2-k]) and B[k] = (X[k] - conj(X[n/2-k])) e^{2 pi i k/n},
// z = IFFT_{n/2}(A + iB) is x[2m] + i x[2m+1], i.e. the real row itself.  ccf_mult_kernel
// writes A + iB directly, so the product spectrum crosses HBM once on the way in and the
// correlation once on the way out; cuFFT's own Z2D spends two more passes over the data
// on the same packing and unpacking (its preprocess / unpackC2R kernels were 40 % of the
// accumulate stage in the launch list).
#include <cufft.h>
#include <math.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace rvs {

// ---------------------------------------------------------------- plan cache
static std::mutex g_plan_mu;
static std::map<std::tuple<int, int, int>, cufftHandle> g_plans;  // (kind, n, batch)

static int get_plan(int kind, int n, int batch, cufftHandle *out) {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  const auto key = std::make_tuple(kind + 4 * dev, n, batch);
  auto it = g_plans.find(key);
  if (it == g_plans.end()) {
    cufftHandle p;
    int nn[1] = {n};
    const int nfreq = n / 2 + 1;
    cufftResult r;
    if (kind == 0) {  // D2Z, out of place, packed
      int in_e[1] = {n}, out_e[1] = {nfreq};
      r = cufftPlanMany(&p, 1, nn, in_e, 1, n, out_e, 1, nfreq, CUFFT_D2Z, batch);
    } else if (kind == 1) {  // Z2D in place
      int in_e[1] = {nfreq}, out_e[1] = {2 * nfreq};
      r = cufftPlanMany(&p, 1, nn, in_e, 1, nfreq, out_e, 1, 2 * nfreq, CUFFT_Z2D, batch);
    } else {  // Z2Z of n/2 points in place, rows of n/2 complex
      int nh[1] = {n / 2};
      r = cufftPlanMany(&p, 1, nh, nh, 1, n / 2, nh, 1, n / 2, CUFFT_Z2Z, batch);
    }
    if (r != CUFFT_SUCCESS) {
      set_error("cufftPlanMany(kind=%d, n=%d, batch=%d) failed: %d", kind, n, batch, (int)r);
      return RVS_E_CUDA;
    }
    it = g_plans.emplace(key, p).first;
  }
  *out = it->second;
  return 0;
}

// -------------------------------------------------------------------- kernels
// one CTA per object: a = ps * pi ; sse[row] += sum (ps*ps)*pi  (fixed tree)
__global__ void __launch_bounds__(256) ccf_prep_kernel(const double *ps, const double *pi, int n,
                                                       const int32_t *row, int b0, double *a,
                                                       double *sse) {
  const int b = blockIdx.x;
  const double *s = ps + (int64_t)(b0 + b) * n, *w = pi + (int64_t)(b0 + b) * n;
  double *o = a + (int64_t)b * n;
  double acc = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double sv = s[i], wv = w[i];
    o[i] = sv * wv;
    acc += (sv * sv) * wv;
  }
  __shared__ double red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w2 = 0; w2 < 8; w2++) t += red[w2];
    const int r = row ? row[b0 + b] : b0 + b;
    sse[r] += t;
  }
}

// grid (ceil(nfreq/256), ntempl); each thread keeps its template bins in
// registers and streams over the objects of the chunk
template <bool CONT>
__global__ void __launch_bounds__(256) ccf_mult_kernel(const double2 *T, const double2 *T2,
                                                       const double2 *SF, const double2 *IF,
                                                       int nfreq, int ntempl, int nb, double inv_n,
                                                       double2 *Z0, double2 *Z1) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (f >= nfreq) return;
  const double2 a = T[(int64_t)t * nfreq + f], c = T2[(int64_t)t * nfreq + f];
  const bool edge = (f == 0) || (f == nfreq - 1);  // DC / Nyquist: imaginary part unused
  for (int b = 0; b < nb; b++) {
    const double2 s = SF[(int64_t)b * nfreq + f], w = IF[(int64_t)b * nfreq + f];
    // x * conj(y) = (xr yr + xi yi) + i (xi yr - xr yi)
    double2 p0 = make_double2(a.x * s.x + a.y * s.y, a.y * s.x - a.x * s.y);
    double2 p1 = make_double2(c.x * w.x + c.y * w.y, c.y * w.x - c.x * w.y);
    const int64_t o = ((int64_t)b * ntempl + t) * nfreq + f;
    if (CONT) {
      double2 z = make_double2((p1.x - 2 * p0.x) * inv_n, (p1.y - 2 * p0.y) * inv_n);
      if (edge) z.y = 0;
      Z0[o] = z;
    } else {
      p0.x *= inv_n; p0.y *= inv_n; p1.x *= inv_n; p1.y *= inv_n;
      if (edge) { p0.y = 0; p1.y = 0; }
      Z0[o] = p0;
      Z1[o] = p1;
    }
  }
}

// The same products packed for the half-size complex transform (see the file comment):
// grid (ceil((n/4 + 1)/256), ntempl); thread k <= n/4 forms X[k] and X[n/2-k] of every
// object of the chunk from the template bins it keeps in registers and writes BOTH
// Z[k] = A + iB and Z[n/2-k] = conj(A) + i conj(B) (rows of n/2 complex = the n real values
// of the correlation after the transform), so every product is formed once.
template <bool CONT>
__global__ void __launch_bounds__(256) ccf_mult_half_kernel(const double2 *T, const double2 *T2,
                                                            const double2 *SF, const double2 *IF,
                                                            int nfreq, int ntempl, int nb,
                                                            double inv_n, double2 *Z0, double2 *Z1) {
  const int nh = nfreq - 1;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (2 * k > nh) return;
  const int kp = nh - k;                       // partner bin (k = 0: the Nyquist bin)
  const double2 a = T[(int64_t)t * nfreq + k], c = T2[(int64_t)t * nfreq + k];
  const double2 ap = T[(int64_t)t * nfreq + kp], cp = T2[(int64_t)t * nfreq + kp];
  const bool edge = k == 0;                    // DC and Nyquist: imaginary parts unused
  const bool both = k > 0 && kp != k;          // Z[n/2-k] is a row element of its own
  double tw_s, tw_c;
  sincospi(2.0 * (double)k / (double)(2 * nh), &tw_s, &tw_c);
  auto pack = [&](double2 x, double2 xp, double2 *Z, int64_t row) {
    if (edge) { x.y = 0; xp.y = 0; }
    const double ar = x.x + xp.x, ai = x.y - xp.y;          // A = X + conj(Xp)
    const double dr = x.x - xp.x, di = x.y + xp.y;          // X - conj(Xp)
    const double br = dr * tw_c - di * tw_s, bi = dr * tw_s + di * tw_c;   // B
    Z[row + k] = make_double2(ar - bi, ai + br);             // A + iB
    if (both) Z[row + kp] = make_double2(ar + bi, br - ai);  // conj(A) + i conj(B)
  };
  for (int b = 0; b < nb; b++) {
    const double2 s = SF[(int64_t)b * nfreq + k], w = IF[(int64_t)b * nfreq + k];
    const double2 sp = SF[(int64_t)b * nfreq + kp], wp = IF[(int64_t)b * nfreq + kp];
    // x * conj(y) = (xr yr + xi yi) + i (xi yr - xr yi)
    double2 p0 = make_double2(a.x * s.x + a.y * s.y, a.y * s.x - a.x * s.y);
    double2 p1 = make_double2(c.x * w.x + c.y * w.y, c.y * w.x - c.x * w.y);
    double2 q0 = make_double2(ap.x * sp.x + ap.y * sp.y, ap.y * sp.x - ap.x * sp.y);
    double2 q1 = make_double2(cp.x * wp.x + cp.y * wp.y, cp.y * wp.x - cp.x * wp.y);
    const int64_t row = ((int64_t)b * ntempl + t) * nh;
    if (CONT) {
      const double2 x = make_double2((p1.x - 2 * p0.x) * inv_n, (p1.y - 2 * p0.y) * inv_n);
      const double2 xp = make_double2((q1.x - 2 * q0.x) * inv_n, (q1.y - 2 * q0.y) * inv_n);
      pack(x, xp, Z0, row);
    } else {
      p0.x *= inv_n; p0.y *= inv_n; p1.x *= inv_n; p1.y *= inv_n;
      q0.x *= inv_n; q0.y *= inv_n; q1.x *= inv_n; q1.y *= inv_n;
      pack(p0, q0, Z0, row);
      pack(p1, q1, Z1, row);
    }
  }
}

// one CTA per (object, template) row
template <bool CONT>
__global__ void __launch_bounds__(128) ccf_gather_kernel(const double *R0, const double *R1,
                                                         int64_t rstride, int ntempl, int nvel,
                                                         const int32_t *lo, const int32_t *hi,
                                                         const double *dxn, const double *dx,
                                                         const int32_t *row, int b0,
                                                         double *chisq) {
  const int64_t bt = blockIdx.x;
  const int b = (int)(bt / ntempl), t = (int)(bt - (int64_t)b * ntempl);
  const double *r0 = R0 + bt * rstride, *r1 = CONT ? nullptr : R1 + bt * rstride;
  const int r = row ? row[b0 + b] : b0 + b;
  double *out = chisq + ((int64_t)r * ntempl + t) * nvel;
  for (int j = threadIdx.x; j < nvel; j += blockDim.x) {
    const int l = lo[j], h = hi[j];
    double yl, yh;
    if (CONT) {
      yl = r0[l];
      yh = r0[h];
    } else {
      const double a0 = r0[l], a1 = r0[h];
      yl = -(a0 * a0) / r1[l];
      yh = -(a1 * a1) / r1[h];
    }
    // scipy interp1d._call_linear: slope * (x_new - x_lo) + y_lo
    const double slope = (yh - yl) / dx[j];
    out[j] += slope * dxn[j] + yl;
  }
}

// NaN-aware "less" of numpy argmin: the first NaN wins, else the first minimum
__device__ __forceinline__ bool np_less(double a, int ia, double b, int ib) {
  const bool na = isnan(a), nb = isnan(b);
  if (na || nb) return na && (!nb || ia < ib);
  return a < b || (a == b && ia < ib);
}

// one CTA (256 threads) per object row
__global__ void __launch_bounds__(256) ccf_best_kernel(const double *chisq, const double *sse,
                                                       const double *vgrid, int ntempl, int nvel,
                                                       double *out, double *best_ccf) {
  const int r = blockIdx.x;
  const double *c = chisq + (int64_t)r * ntempl * nvel;
  const double add = sse[r];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __shared__ double sval[8];
  __shared__ int sidx[8];
  // best template: argmin over t of min_j (NaN-propagating row minimum)
  double bval = INFINITY;
  int bt = 0x7fffffff;
  for (int t = wid; t < ntempl; t += nw) {
    const double *row = c + (int64_t)t * nvel;
    double m = INFINITY;
    bool nan_seen = false;
    for (int j = lane; j < nvel; j += 32) {
      const double v = row[j] + add;
      nan_seen |= isnan(v);
      m = fmin(m, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
      m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
      nan_seen |= __shfl_xor_sync(0xffffffffu, (int)nan_seen, o) != 0;
    }
    if (nan_seen) m = nan("");
    if (np_less(m, t, bval, bt)) { bval = m; bt = t; }
  }
  if (lane == 0) { sval[wid] = bval; sidx[wid] = bt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nw; w++)
      if (np_less(sval[w], sidx[w], sval[0], sidx[0])) { sval[0] = sval[w]; sidx[0] = sidx[w]; }
  }
  __syncthreads();
  const int best_t = sidx[0];
  __syncthreads();
  // best pixel of that template
  const double *row = c + (int64_t)best_t * nvel;
  double pv = INFINITY;
  int pj = 0x7fffffff;
  for (int j = threadIdx.x; j < nvel; j += blockDim.x) {
    const double v = row[j] + add;
    if (best_ccf) best_ccf[(int64_t)r * nvel + j] = v;
    if (np_less(v, j, pv, pj)) { pv = v; pj = j; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, pv, o);
    const int oj = __shfl_xor_sync(0xffffffffu, pj, o);
    if (np_less(ov, oj, pv, pj)) { pv = ov; pj = oj; }
  }
  if (lane == 0) { sval[wid] = pv; sidx[wid] = pj; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nw; w++)
      if (np_less(sval[w], sidx[w], sval[0], sidx[0])) { sval[0] = sval[w]; sidx[0] = sidx[w]; }
    const int bp = sidx[0];
    const double bc = sval[0];
    double bv = vgrid[bp];
    if (bp != 0 && bp != nvel - 1) {
      // parabola through the three points around the minimum (fitter_ccf.py:210-218)
      const double x0 = vgrid[bp - 1] - vgrid[bp], x2 = vgrid[bp + 1] - vgrid[bp];
      const double y0 = (row[bp - 1] + add) - bc, y2 = (row[bp + 1] + add) - bc;
      const double den = x0 * x2 * (x0 - x2);
      const double a2 = (y0 * x2 - y2 * x0) / den;
      const double a1 = (y2 * x0 * x0 - y0 * x2 * x2) / den;
      if (a2 > 0) bv = vgrid[bp] - a1 / (2 * a2);
    }
    double *o = out + (int64_t)r * 8;
    o[0] = best_t; o[1] = bp; o[2] = bv; o[3] = bc; o[4] = isfinite(bc) ? 1.0 : 0.0;
    o[5] = 0; o[6] = 0; o[7] = 0;
  }
}

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// workspace bytes for a chunk of nb objects: A | SF | IF | Z0 (| Z1)
static int64_t ccf_bytes(const rvs_ccf_arm *arm, int64_t nb) {
  const int64_t n = arm->npoints, nfreq = n / 2 + 1;
  const int64_t zrows = nb * arm->ntempl;
  return align_up(nb * n * 8, 256) + 2 * align_up(nb * nfreq * 16, 256) +
         (arm->continuum ? 1 : 2) * align_up(zrows * nfreq * 16, 256);
}

}  // namespace rvs

extern "C" int64_t rvs_ccf_workspace(const rvs_ccf_arm *arm, int nb) {
  if (!arm || nb < 1) return 0;
  return rvs::ccf_bytes(arm, nb);
}

extern "C" int rvs_ccf_accumulate(const rvs_ccf_arm *arm, const double *d_pspec,
                                  const double *d_pivar, int B, const int32_t *d_row,
                                  double *d_chisq, double *d_sse, void *d_work,
                                  int64_t work_bytes, void *stream) {
  using namespace rvs;
  if (B == 0) return 0;
  RVS_REQUIRE(arm && d_pspec && d_pivar && d_chisq && d_sse && d_work, RVS_E_ARG,
              "rvs_ccf_accumulate: null pointer");
  RVS_REQUIRE(arm->d_fft && arm->d_fft2 && arm->d_lo && arm->d_hi && arm->d_dxn && arm->d_dx,
              RVS_E_ARG, "rvs_ccf_accumulate: null pointer in rvs_ccf_arm");
  const int n = arm->npoints, nt = arm->ntempl, nvel = arm->nvel;
  RVS_REQUIRE(n >= 4 && n % 2 == 0 && nt >= 1 && nvel >= 1, RVS_E_ARG,
              "rvs_ccf_accumulate: npoints=%d (even, >= 4), ntempl=%d, nvel=%d", n, nt, nvel);
  RVS_REQUIRE(((uintptr_t)d_work & 255) == 0, RVS_E_ARG, "rvs_ccf_accumulate: d_work alignment");
  const int64_t nfreq = n / 2 + 1;
  int64_t nbmax = B;
  while (nbmax > 1 && ccf_bytes(arm, nbmax) > work_bytes) nbmax = nbmax > 8 ? nbmax * 3 / 4 : nbmax - 1;
  RVS_REQUIRE(ccf_bytes(arm, nbmax) <= work_bytes, RVS_E_LIMIT,
              "rvs_ccf_accumulate: workspace of %lld B is below the %lld B one object needs",
              (long long)work_bytes, (long long)ccf_bytes(arm, 1));
  while (nbmax * nt > (1 << 20)) nbmax = (nbmax + 1) / 2;  // bound the cuFFT batch
  cudaStream_t st = (cudaStream_t)stream;
  char *w = static_cast<char *>(d_work);
  double *A = reinterpret_cast<double *>(w);
  w += align_up(nbmax * n * 8, 256);
  double2 *SF = reinterpret_cast<double2 *>(w);
  w += align_up(nbmax * nfreq * 16, 256);
  double2 *IF = reinterpret_cast<double2 *>(w);
  w += align_up(nbmax * nfreq * 16, 256);
  double2 *Z0 = reinterpret_cast<double2 *>(w);
  w += align_up(nbmax * nt * nfreq * 16, 256);
  double2 *Z1 = arm->continuum ? nullptr : reinterpret_cast<double2 *>(w);
  const double2 *T = reinterpret_cast<const double2 *>(arm->d_fft);
  const double2 *T2 = reinterpret_cast<const double2 *>(arm->d_fft2);
  for (int64_t b0 = 0; b0 < B; b0 += nbmax) {
    const int nb = (int)((B - b0 < nbmax) ? B - b0 : nbmax);
    ccf_prep_kernel<<<nb, 256, 0, st>>>(d_pspec, d_pivar, n, d_row, (int)b0, A, d_sse);
    RVS_LAUNCH_OK();
    cufftHandle fwd, inv;
    int rc = get_plan(0, n, nb, &fwd);
    if (rc) return rc;
    rc = get_plan(2, n, nb * nt, &inv);
    if (rc) return rc;
    RVS_REQUIRE(cufftSetStream(fwd, st) == CUFFT_SUCCESS && cufftSetStream(inv, st) == CUFFT_SUCCESS,
                RVS_E_CUDA, "cufftSetStream failed");
    RVS_REQUIRE(cufftExecD2Z(fwd, A, reinterpret_cast<cufftDoubleComplex *>(SF)) == CUFFT_SUCCESS,
                RVS_E_CUDA, "cufftExecD2Z failed");
    RVS_REQUIRE(cufftExecD2Z(fwd, const_cast<double *>(d_pivar) + b0 * n,
                             reinterpret_cast<cufftDoubleComplex *>(IF)) == CUFFT_SUCCESS,
                RVS_E_CUDA, "cufftExecD2Z failed");
    dim3 grid((unsigned)((n / 4 + 1 + 255) / 256), (unsigned)nt);
    if (arm->continuum)
      ccf_mult_half_kernel<true><<<grid, 256, 0, st>>>(T, T2, SF, IF, (int)nfreq, nt, nb, 1.0 / n, Z0, Z1);
    else
      ccf_mult_half_kernel<false><<<grid, 256, 0, st>>>(T, T2, SF, IF, (int)nfreq, nt, nb, 1.0 / n, Z0, Z1);
    RVS_LAUNCH_OK();
    RVS_REQUIRE(cufftExecZ2Z(inv, reinterpret_cast<cufftDoubleComplex *>(Z0),
                             reinterpret_cast<cufftDoubleComplex *>(Z0), CUFFT_INVERSE) == CUFFT_SUCCESS,
                RVS_E_CUDA, "cufftExecZ2Z failed");
    if (!arm->continuum)
      RVS_REQUIRE(cufftExecZ2Z(inv, reinterpret_cast<cufftDoubleComplex *>(Z1),
                               reinterpret_cast<cufftDoubleComplex *>(Z1), CUFFT_INVERSE) == CUFFT_SUCCESS,
                  RVS_E_CUDA, "cufftExecZ2Z failed");
    const unsigned rows = (unsigned)(nb * nt);
    if (arm->continuum)
      ccf_gather_kernel<true><<<rows, 128, 0, st>>>(
          reinterpret_cast<const double *>(Z0), nullptr, (int64_t)n, nt, nvel, arm->d_lo, arm->d_hi,
          arm->d_dxn, arm->d_dx, d_row, (int)b0, d_chisq);
    else
      ccf_gather_kernel<false><<<rows, 128, 0, st>>>(
          reinterpret_cast<const double *>(Z0), reinterpret_cast<const double *>(Z1), (int64_t)n, nt,
          nvel, arm->d_lo, arm->d_hi, arm->d_dxn, arm->d_dx, d_row, (int)b0, d_chisq);
    RVS_LAUNCH_OK();
  }
  return 0;
}

extern "C" int rvs_ccf_best(const double *d_chisq, const double *d_sse, const double *d_velgrid,
                            int nrow, int ntempl, int nvel, double *d_out, double *d_best_ccf,
                            void *stream) {
  using namespace rvs;
  if (nrow == 0) return 0;
  RVS_REQUIRE(d_chisq && d_sse && d_velgrid && d_out && ntempl >= 1 && nvel >= 1, RVS_E_ARG,
              "rvs_ccf_best: bad arguments");
  ccf_best_kernel<<<nrow, 256, 0, (cudaStream_t)stream>>>(d_chisq, d_sse, d_velgrid, ntempl, nvel,
                                                          d_out, d_best_ccf);
  RVS_LAUNCH_OK();
  return 0;
}

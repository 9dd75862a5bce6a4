// Observed-spectrum products, continuum basis and the RV-scan chi-square kernel
// (Doppler spline resampling + continuum normal equations + residual norm).
//
// Reference behaviour reproduced (paths under /root/reference/py/rvspecfit/):
//   spec_fit.py:103-108, 148-176   per-spectrum products, continuum basis
//   spec_fit.py:707-727 + src/spliner.c:71-108   Doppler shift + spline evaluation
//   spec_fit.py:205-249            normal equations, Cholesky, residual norm
#include <math.h>

#include "chisq_device.cuh"

namespace rvs {

// ------------------------------------------------------------- obs products
__global__ void obs_prepare_kernel(const double *spec, const double *espec, const int64_t *off,
                                   double sys, double *dn, double *einv, double *sumlog2) {
  const int i = blockIdx.x;
  const int64_t p0 = off[i], p1 = off[i + 1];
  double s = 0;
  for (int64_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
    double e = espec[p];
    if (sys != 0) e = sqrt(sys * sys + e * e);
    dn[p] = spec[p] / e;
    einv[p] = 1.0 / e;
    s += log(e);
  }
  __shared__ double red[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
    sumlog2[i] = 2.0 * t;
  }
}

__global__ void basis_kernel(const double *lam, const int64_t *goff, int G, int64_t ntot,
                             int npoly, int rbf, int npp, double *loglam, double *P) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ntot) return;
  int lo = 0, hi = G;  // grid g with goff[g] <= p < goff[g+1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (goff[mid] <= p) lo = mid; else hi = mid;
  }
  const int64_t p0 = goff[lo], p1 = goff[lo + 1];
  const double l0 = lam[p0], l1 = lam[p1 - 1];
  const double t = (lam[p] - l0) / (l1 - l0) * 2 - 1;
  loglam[p] = log(lam[p]);
  double *row = P + p * npp;
  for (int r = npoly; r < npp; r++) row[r] = 0;
  if (rbf) {
    double pw = 1;
    for (int r = 0; r < min(3, npoly); r++) {
      row[r] = pw;
      pw *= t;
    }
    const int nr = npoly - 3;
    if (nr > 0) {
      const double sig = 1.0 / nr;
      const double step = nr > 1 ? 2.0 / (nr - 1) : 0.0;
      for (int r = 0; r < nr; r++) {
        // numpy.linspace(-1, 1, nr): start + r*step, last point forced to stop
        const double c = (nr > 1 && r == nr - 1) ? 1.0 : -1.0 + r * step;
        const double dlt = t - c;
        row[3 + r] = exp(-0.5 * (dlt * dlt) / (sig * sig));
      }
    }
  } else {
    double tm = 1, tc = t;
    for (int r = 0; r < npoly; r++) {
      row[r] = (r == 0) ? 1.0 : tc;
      if (r >= 1) {
        const double tn = 2 * t * tc - tm;
        tm = tc;
        tc = tn;
      }
    }
  }
}

template <int NP>
__global__ void __launch_bounds__(SCAN_WARPS * 32) chisq_scan_kernel(ScanArgs a) {
  constexpr int NTRI = NP * (NP + 1) / 2;
  constexpr int RSPLIT = NP > 10 ? 10 : NP;  // rows handled in the first sweep
  __shared__ double sM[SCAN_WARPS][NTRI];
  __shared__ double sV[SCAN_WARPS][NP];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int k = blockIdx.x;
  const int obj = a.oix[k];
  const int64_t p0 = a.off[obj];
  const int npix = (int)(a.off[obj + 1] - p0);
  const int64_t b0 = a.goff[obj];
  const double2 *yz = a.yz + (int64_t)a.tix[k] * a.yz_stride;
  const double *lam = a.lam + b0, *ql = (a.log_step ? a.loglam : a.lam) + b0;
  const double *Pb = a.P + b0 * a.npp;
  const double *dn = a.dn + p0, *einv = a.einv + p0;
  const double *rb = a.resol ? a.resol + p0 * a.nresol : nullptr;
  auto ev = [&](double x, double q) { return spline_eval(a, yz, x, q); };

  for (int j = blockIdx.y * SCAN_WARPS + wid; j < a.nv; j += gridDim.y * SCAN_WARPS) {
    const double beta = a.vels[(int64_t)k * a.nv + j] / RVS_C_KMS;
    const double f = sqrt((1 - beta) / (1 + beta));
    const double qf = a.log_step ? log(f) : 0.0;
    int st = 0;
    // the reference checks the first and last evaluation points (spliner.c:78-83)
    {
      const double xa = lam[0] * f, xb = lam[npix - 1] * f;
      if (xa < a.x0 || xb < a.x0 || xa >= a.xlast || xb >= a.xlast) st |= RVS_ST_RANGE;
    }
    // ---- sweep 1: rows [0,RSPLIT) of M and v
    {
      GramAcc<NP, 0, RSPLIT> acc;
      acc.zero();
      double v[NP];
#pragma unroll
      for (int i = 0; i < NP; i++) v[i] = 0;
      for (int p = lane; p < npix; p += 32) {
        const double tn = template_at(a, lam, ql, rb, npix, p, f, qf, ev) * einv[p];
        double g[NP];
        load_basis<NP>(Pb + (int64_t)p * a.npp, tn, g);
        const double d = dn[p];
#pragma unroll
        for (int i = 0; i < NP; i++) v[i] = fma(g[i], d, v[i]);
        acc.add(g);
      }
      acc.reduce_store(sM[wid], lane);
      warp_reduce_store<NP>(v, sV[wid], lane);
    }
    if (NP > RSPLIT) {  // ---- sweep 1b: remaining rows
      GramAcc<NP, RSPLIT, NP> acc;
      acc.zero();
      for (int p = lane; p < npix; p += 32) {
        const double tn = template_at(a, lam, ql, rb, npix, p, f, qf, ev) * einv[p];
        double g[NP];
        load_basis<NP>(Pb + (int64_t)p * a.npp, tn, g);
        acc.add(g);
      }
      acc.reduce_store(sM[wid], lane);
    }
    __syncwarp();
    const double ldet = chol_solve<NP>(sM[wid], sV[wid], lane);
    double co[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) co[i] = sV[wid][i];
    // ---- sweep 2: residual norm |D - a^T G|^2
    double rss = 0;
    const bool want_model = (a.raw != nullptr) && j == 0;
    for (int p = lane; p < npix; p += 32) {
      const double tv = template_at(a, lam, ql, rb, npix, p, f, qf, ev);
      const double tn = tv * einv[p];
      double g[NP];
      load_basis<NP>(Pb + (int64_t)p * a.npp, tn, g);
      double mval = 0;
#pragma unroll
      for (int i = 0; i < NP; i++) mval = fma(co[i], g[i], mval);
      const double r = dn[p] - mval;
      rss = fma(r, r, rss);
      if (want_model) {
        a.raw[a.moff[k] + p] = tv;
        load_basis<NP>(Pb + (int64_t)p * a.npp, tv, g);
        double cont = 0;
#pragma unroll
        for (int i = 0; i < NP; i++) cont = fma(co[i], g[i], cont);
        a.model[a.moff[k] + p] = cont;
      }
    }
    rss = warp_sum(rss);
    const double chi = ldet + a.sumlog2[obj] + rss;
    if (!isfinite(chi)) st |= RVS_ST_NOT_PD;
    if (lane == 0) {
      a.chisq[(int64_t)k * a.nv + j] = chi;
      a.status[(int64_t)k * a.nv + j] = st;
      if (a.coeffs != nullptr && j == 0) {
#pragma unroll
        for (int i = 0; i < NP; i++) a.coeffs[(int64_t)k * NP + i] = co[i];
      }
    }
    __syncwarp();
  }
}

template <int NP>
static int launch_scan(const ScanArgs &a, cudaStream_t st) {
  int by = (a.nv + SCAN_WARPS - 1) / SCAN_WARPS;
  if (by > 65535) by = 65535;
  dim3 grid(a.K, by);
  chisq_scan_kernel<NP><<<grid, SCAN_WARPS * 32, 0, st>>>(a);
  RVS_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------ scan statistics
// One CTA per scan.  spec_fit.py:1072-1092.
// Ragged form: `nvs` (may be NULL) gives the number of velocities of every scan; rows of
// vels / chisq / probs keep the stride nv_stride.
__global__ void scan_stats_kernel(const double *vels, const double *chisq, int npar,
                                  int nv_stride, const int32_t *nvs, int quadratic, double *out,
                                  double *probs) {
  const int s = blockIdx.x;
  const int nv = nvs ? nvs[s] : nv_stride;
  const double *v = vels + (int64_t)s * nv_stride;
  const double *c = chisq + (int64_t)s * npar * nv_stride;
  __shared__ double sval[32];
  __shared__ int sidx[32];
  __shared__ double sred[4][32];
  __shared__ double bc[4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // argmin in numpy's flattened (velocity-major) order: index = i*npar + q
  double best = INFINITY;
  int bidx = 0x7fffffff;
  bool anynan = false;
  for (int t = threadIdx.x; t < npar * nv; t += blockDim.x) {
    const int q = t / nv, i = t - q * nv;
    const double x = c[(int64_t)q * nv_stride + i];
    const int flat = i * npar + q;
    if (isnan(x)) anynan = true;
    if (x < best || (x == best && flat < bidx)) { best = x; bidx = flat; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ob < best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
  }
  if (lane == 0) { sval[wid] = best; sidx[wid] = bidx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nw; w++)
      if (sval[w] < sval[0] || (sval[w] == sval[0] && sidx[w] < sidx[0])) {
        sval[0] = sval[w]; sidx[0] = sidx[w];
      }
    if (sidx[0] == 0x7fffffff) sidx[0] = 0;
  }
  __syncthreads();
  anynan = __syncthreads_or(anynan);
  const int i1 = sidx[0] / npar, i2 = sidx[0] % npar;
  const double *col = c + (int64_t)i2 * nv_stride;
  const double cmin = col[i1];
  // parabola vertex through the three points around the minimum
  double bv = v[i1];
  int vertex_bad = 0;
  if (quadratic && i1 > 0 && i1 < nv - 1) {
    const double x0 = v[i1 - 1] - v[i1], x2 = v[i1 + 1] - v[i1];
    const double y0 = col[i1 - 1] - cmin, y2 = col[i1 + 1] - cmin;
    // y = a2 x^2 + a1 x through (x0,y0),(0,0),(x2,y2)
    const double den = x0 * x2 * (x0 - x2);
    const double a2 = (y0 * x2 - y2 * x0) / den;
    const double a1 = (y2 * x0 * x0 - y0 * x2 * x2) / den;
    bv = v[i1] - a1 / (2 * a2);
    // the reference asserts that the vertex lies strictly inside the bracket
    // (spec_fit.py:1014); a flat or non-finite triple fails it -- reported in out[7]
    if (!(bv > v[i1 - 1] && bv < v[i1 + 1])) vertex_bad = 1;
  }
  if (anynan) vertex_bad |= 2;
  // moments
  auto block4 = [&](double e0, double e1, double e2, double e3) {
    e0 = warp_sum(e0); e1 = warp_sum(e1); e2 = warp_sum(e2); e3 = warp_sum(e3);
    __syncthreads();
    if (lane == 0) { sred[0][wid] = e0; sred[1][wid] = e1; sred[2][wid] = e2; sred[3][wid] = e3; }
    __syncthreads();
    if (threadIdx.x < 4) {
      double t = 0;
      for (int w = 0; w < nw; w++) t += sred[threadIdx.x][w];
      bc[threadIdx.x] = t;
    }
    __syncthreads();
  };
  double psum = 0;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) psum += exp(-0.5 * (col[i] - cmin));
  block4(psum, 0, 0, 0);
  const double tot = bc[0];
  double m2 = 0, m3 = 0, m4 = 0;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const double pr = exp(-0.5 * (col[i] - cmin)) / tot;
    if (probs) probs[(int64_t)s * nv_stride + i] = pr;
    const double dv = v[i] - bv;
    m2 += pr * dv * dv;
    m3 += pr * dv * dv * dv;
    m4 += pr * dv * dv * dv * dv;
  }
  block4(m2, m3, m4, 0);
  if (threadIdx.x == 0) {
    const double err = sqrt(bc[0]);
    double sk = 0, ku = 0;
    if (!(err < 1e-10)) {
      ku = bc[2] / (err * err * err * err);
      sk = bc[1] / (err * err * err);
    }
    double *o = out + (int64_t)s * 8;
    o[0] = cmin; o[1] = bv; o[2] = err; o[3] = sk; o[4] = ku; o[5] = i1; o[6] = i2; o[7] = vertex_bad;
  }
}

}  // namespace rvs

extern "C" int rvs_obs_prepare(const double *d_spec, const double *d_espec, const int64_t *d_off,
                               int B, double espec_sys, double *d_dn, double *d_einv,
                               double *d_sumlog2, void *stream) {
  using namespace rvs;
  if (B == 0) return 0;
  RVS_REQUIRE(d_spec && d_espec && d_off && d_dn && d_einv && d_sumlog2, RVS_E_ARG,
              "rvs_obs_prepare: null pointer");
  obs_prepare_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(d_spec, d_espec, d_off, espec_sys,
                                                          d_dn, d_einv, d_sumlog2);
  RVS_LAUNCH_OK();
  return 0;
}

extern "C" int rvs_basis_build(const double *d_lam, const int64_t *d_gstart, int G, int64_t ntot,
                               int npoly, int rbf, int npp, double *d_loglam, double *d_P,
                               void *stream) {
  using namespace rvs;
  if (G == 0 || ntot == 0) return 0;
  RVS_REQUIRE(d_lam && d_gstart && d_P && d_loglam, RVS_E_ARG, "rvs_basis_build: null pointer");
  RVS_REQUIRE(npoly >= 1 && npoly <= RVS_MAX_NPOLY, RVS_E_ARG,
              "rvs_basis_build: npoly=%d outside 1..%d", npoly, RVS_MAX_NPOLY);
  RVS_REQUIRE(npp >= npoly && npp % 2 == 0 && ((uintptr_t)d_P & 15) == 0, RVS_E_ARG,
              "rvs_basis_build: npp=%d must be even and >= npoly, d_P 16-byte aligned", npp);
  const int bx = (int)((ntot + 255) / 256);
  basis_kernel<<<bx, 256, 0, (cudaStream_t)stream>>>(d_lam, d_gstart, G, ntot, npoly, rbf, npp,
                                                     d_loglam, d_P);
  RVS_LAUNCH_OK();
  return 0;
}

namespace rvs {
int launch_scan_mma_group0(const ScanArgs &, int, cudaStream_t);
int launch_scan_mma_group1(const ScanArgs &, int, cudaStream_t);
int launch_scan_mma_group2(const ScanArgs &, int, cudaStream_t);
int launch_scan_mma_group3(const ScanArgs &, int, cudaStream_t);

int fill_scan_args(ScanArgs &a, const rvs_knots *kn, const rvs_obs *obs) {
  RVS_REQUIRE(kn && obs && kn->d_lam_t && kn->d_h && kn->d_hinv && obs->d_lam &&
                  obs->d_loglam && obs->d_dn && obs->d_einv && obs->d_sumlog2 && obs->d_off &&
                  obs->d_P && obs->d_goff,
              RVS_E_ARG, "chisq: null pointer in descriptor");
  RVS_REQUIRE(obs->npoly >= 1 && obs->npoly <= RVS_MAX_NPOLY, RVS_E_ARG,
              "chisq: npoly=%d outside 1..%d", obs->npoly, RVS_MAX_NPOLY);
  a.lam_t = kn->d_lam_t; a.h = kn->d_h; a.hinv = kn->d_hinv; a.npix_t = kn->npix_t;
  a.log_step = kn->log_step; a.x0 = kn->x0; a.xlast = kn->xlast; a.q0 = kn->q0;
  a.qstep_inv = kn->qstep_inv;
  a.lam = obs->d_lam; a.loglam = obs->d_loglam; a.dn = obs->d_dn; a.einv = obs->d_einv;
  a.sumlog2 = obs->d_sumlog2; a.off = obs->d_off; a.goff = obs->d_goff; a.P = obs->d_P;
  a.npp = obs->npp;
  a.fast_interp = 0;
  a.nby = 0;
  a.nvk = nullptr;
  RVS_REQUIRE(obs->nresol >= 0 && (obs->nresol == 0) == (obs->d_resol == nullptr) &&
                  (obs->nresol == 0 || obs->d_resol_offs),
              RVS_E_ARG, "chisq: d_resol, d_resol_offs and nresol must be set together");
  a.resol = obs->d_resol; a.resol_offs = obs->d_resol_offs; a.nresol = obs->nresol;
  a.resol_hw = obs->resol_halfwidth;
  RVS_REQUIRE(obs->nresol == 0 || obs->resol_halfwidth >= 0, RVS_E_ARG,
              "chisq: resol_halfwidth must be max |d_resol_offs[]|");
  RVS_REQUIRE(obs->npp >= obs->npoly && obs->npp % 2 == 0 && ((uintptr_t)obs->d_P & 15) == 0,
              RVS_E_ARG, "chisq: basis rows must be npp = even >= npoly doubles, 16-byte aligned");
  return 0;
}
}  // namespace rvs

extern "C" int rvs_chisq_scan(const double *d_yz, int64_t yz_stride, const int32_t *d_tix,
                              const rvs_knots *knots, const rvs_obs *obs, const int32_t *d_oix,
                              const double *d_vels, int nv, int K, double *d_chisq,
                              int32_t *d_status, double *d_coeffs, double *d_raw,
                              double *d_model, const int64_t *d_moff, int fast_interp,
                              void *stream) {
  return rvs_chisq_scan_ragged(d_yz, yz_stride, d_tix, knots, obs, d_oix, d_vels, nv, nullptr, K,
                               d_chisq, d_status, d_coeffs, d_raw, d_model, d_moff, fast_interp,
                               stream);
}

extern "C" int rvs_chisq_scan_ragged(const double *d_yz, int64_t yz_stride, const int32_t *d_tix,
                                     const rvs_knots *knots, const rvs_obs *obs,
                                     const int32_t *d_oix, const double *d_vels, int nv,
                                     const int32_t *d_nv, int K, double *d_chisq,
                                     int32_t *d_status, double *d_coeffs, double *d_raw,
                                     double *d_model, const int64_t *d_moff, int fast_interp,
                                     void *stream) {
  using namespace rvs;
  if (K == 0 || nv == 0) return 0;
  ScanArgs a;
  const int rc = fill_scan_args(a, knots, obs);
  if (rc) return rc;
  RVS_REQUIRE(d_yz && d_tix && d_oix && d_vels && d_chisq && d_status, RVS_E_ARG,
              "rvs_chisq_scan: null pointer");
  RVS_REQUIRE(!(d_raw || d_model) || (d_raw && d_model && d_moff && nv == 1), RVS_E_ARG,
              "rvs_chisq_scan: model output needs raw, model, moff and nv == 1");
  RVS_REQUIRE(!d_coeffs || nv == 1, RVS_E_ARG, "rvs_chisq_scan: coeffs output needs nv == 1");
  a.yz = reinterpret_cast<const double2 *>(d_yz); a.yz_stride = yz_stride; a.tix = d_tix;
  a.oix = d_oix; a.vels = d_vels; a.nv = nv; a.K = K; a.chisq = d_chisq; a.status = d_status;
  a.coeffs = d_coeffs; a.raw = d_raw; a.model = d_model; a.moff = d_moff;
  a.fast_interp = fast_interp ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  // per-item trial counts are honoured by the GEMM kernel; the per-trial kernel evaluates
  // every column (the padding columns hold valid velocities)
  a.nvk = d_nv;
  if (nv >= 4 && !d_coeffs && !d_raw && !d_model && (!a.resol || a.resol_hw <= RS_HW)) {
    // several trials per template: the trials are the columns of an FP64 GEMM (scan_mma.cuh)
    const int np = obs->npoly;
    if (np <= 7) return launch_scan_mma_group0(a, np, st);
    if (np <= 10) return launch_scan_mma_group1(a, np, st);
    if (np <= 13) return launch_scan_mma_group2(a, np, st);
    return launch_scan_mma_group3(a, np, st);
  }
  switch (obs->npoly) {
#define RVS_CASE(N) case N: return launch_scan<N>(a, st);
    RVS_CASE(1) RVS_CASE(2) RVS_CASE(3) RVS_CASE(4) RVS_CASE(5) RVS_CASE(6) RVS_CASE(7)
    RVS_CASE(8) RVS_CASE(9) RVS_CASE(10) RVS_CASE(11) RVS_CASE(12) RVS_CASE(13)
    RVS_CASE(14) RVS_CASE(15) RVS_CASE(16)
#undef RVS_CASE
  }
  return RVS_E_ARG;
}

extern "C" int rvs_scan_stats(const double *d_vels, const double *d_chisq, int S, int npar,
                              int nv, int quadratic, double *d_out, double *d_probs,
                              void *stream) {
  using namespace rvs;
  if (S == 0) return 0;
  RVS_REQUIRE(d_vels && d_chisq && d_out && npar >= 1 && nv >= 1, RVS_E_ARG,
              "rvs_scan_stats: bad arguments");
  scan_stats_kernel<<<S, 128, 0, (cudaStream_t)stream>>>(d_vels, d_chisq, npar, nv, nullptr,
                                                         quadratic, d_out, d_probs);
  RVS_LAUNCH_OK();
  return 0;
}

extern "C" int rvs_scan_stats_ragged(const double *d_vels, const double *d_chisq, int S, int npar,
                                     int nv_stride, const int32_t *d_nv, int quadratic,
                                     double *d_out, double *d_probs, void *stream) {
  using namespace rvs;
  if (S == 0) return 0;
  RVS_REQUIRE(d_vels && d_chisq && d_out && d_nv && npar >= 1 && nv_stride >= 1, RVS_E_ARG,
              "rvs_scan_stats_ragged: bad arguments");
  scan_stats_kernel<<<S, 128, 0, (cudaStream_t)stream>>>(d_vels, d_chisq, npar, nv_stride, d_nv,
                                                         quadratic, d_out, d_probs);
  RVS_LAUNCH_OK();
  return 0;
}

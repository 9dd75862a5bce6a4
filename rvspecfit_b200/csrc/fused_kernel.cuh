// Fused optimiser-phase kernel: one CTA per (item, arm).  Template build
// (gather + exp + vsini + spline) entirely in shared memory, then Doppler
// resampling, continuum normal equations and residual norm at one velocity.
// HBM traffic per item = the gathered grid rows + the object's pixels; the
// spline never leaves the SM.  (algorithmic bytes.)
#pragma once
#include "chisq_device.cuh"
#include "template_device.cuh"

namespace rvs {

struct FusedArgs {
  TemplateArgs t;
  ScanArgs s;
};

template <typename GT, int NV, int NP>
__global__ void __launch_bounds__(TB_THREADS) chisq_fused_kernel(FusedArgs fa) {
  constexpr int NTRI = NP * (NP + 1) / 2;
  constexpr int RSPLIT = NP > 10 ? 10 : NP;
  constexpr int NW = TB_THREADS / 32;
  const TemplateArgs &a = fa.t;
  const ScanArgs &s = fa.s;
  extern __shared__ double sm[];
  double *ya = sm, *yb = sm + a.npad, *yc = sm + 2 * a.npad, *taps = sm + 3 * a.npad;
  __shared__ int32_t s_ids[32];
  __shared__ double s_w[32];
  __shared__ double red[NW];
  __shared__ int s_bad, s_taps;
  __shared__ double sM[NW][NTRI];
  __shared__ double sV[NW][NP];
  __shared__ double s_ldet;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int k = blockIdx.x;
  if (tid < a.nvert) {
    s_ids[tid] = a.ids[(int64_t)k * a.nvert + tid];
    s_w[tid] = a.w[(int64_t)k * a.nvert + tid];
  }
  if (tid == 0) { s_bad = 0; s_taps = 0; }
  __syncthreads();
  const bool f32row = single_row_item(a, s_ids, s_w);
  int bad = 0;
  gather_rows<GT, NV>(a, s_ids, s_w, ya, bad, f32row);
  if (bad) s_bad = 1;
  __syncthreads();
  double *py, *pz;
  broaden_and_spline(a, a.vsini ? a.vsini[k] : 0.0, ya, yb, yc, taps, red, py, pz, &s_taps);
  double *tnbuf = yb;  // free after the spline solve

  const int obj = s.oix[k];
  const int64_t p0 = s.off[obj];
  const int npix = (int)(s.off[obj + 1] - p0);
  const int64_t b0 = s.boff[obj];
  const double *lam = s.lam + p0, *ql = (s.log_step ? s.loglam : s.lam) + p0;
  const double *dn = s.dn + p0, *einv = s.einv + p0;
  const bool stash = npix <= a.npad;
  const double beta = s.vels[k] / RVS_C_KMS;
  const double f = sqrt((1 - beta) / (1 + beta));
  const double qf = s.log_step ? log(f) : 0.0;
  int st = (s_bad ? RVS_ST_TEMPLATE_BAD : 0) | (s_taps ? RVS_ST_TAPS : 0);
  {
    const double xa = lam[0] * f, xb = lam[npix - 1] * f;
    if (xa < s.x0 || xb < s.x0 || xa >= s.xlast || xb >= s.xlast) st |= RVS_ST_RANGE;
  }
  auto templ_at = [&](int p) -> double {
    const double x = lam[p] * f;
    const double q = s.log_step ? ql[p] + qf : x;
    int pos = (int)((q - s.q0) * s.qstep_inv);
    pos = max(0, min(pos, s.npix_t - 2));
    const double y0 = py[pos], y1 = py[pos + 1], z0 = pz[pos], z1 = pz[pos + 1];
    const double xl = __ldg(s.lam_t + pos), xr = __ldg(s.lam_t + pos + 1);
    const double hh = __ldg(s.h + pos), hi = __ldg(s.hinv + pos);
    const double t1 = hi * (1. / 6), t2 = hh * (1. / 6);
    const double A = z1 * t1, B = z0 * t1;
    const double C = y1 * hi - z1 * t2, D = y0 * hi - z0 * t2;
    const double dl = x - xl, dr = xr - x;
    return A * dl * dl * dl + B * dr * dr * dr + C * dl + D * dr;
  };
  // ---- sweep 1
  {
    GramAcc<NP, 0, RSPLIT> acc;
    acc.zero();
    double v[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) v[i] = 0;
    for (int p = tid; p < npix; p += TB_THREADS) {
      const double tn = templ_at(p) * einv[p];
      if (stash) tnbuf[p] = tn;
      double g[NP];
      load_basis<NP>(s.P, s.pstride, b0 + p, tn, g);
      const double d = dn[p];
#pragma unroll
      for (int i = 0; i < NP; i++) v[i] = fma(g[i], d, v[i]);
      acc.add(g);
    }
    acc.reduce_store(sM[wid], lane);
    warp_reduce_store<NP>(v, sV[wid], lane);
  }
  __syncthreads();
  if (NP > RSPLIT) {
    GramAcc<NP, RSPLIT, NP> acc;
    acc.zero();
    for (int p = tid; p < npix; p += TB_THREADS) {
      const double tn = stash ? tnbuf[p] : templ_at(p) * einv[p];
      double g[NP];
      load_basis<NP>(s.P, s.pstride, b0 + p, tn, g);
      acc.add(g);
    }
    acc.reduce_store(sM[wid], lane);
    __syncthreads();
  }
  // cross-warp sums in fixed order (NTRI + NP <= 152 < TB_THREADS)
  {
    double t = 0;
    if (tid < NTRI) {
      for (int w = 0; w < NW; w++) t += sM[w][tid];
    } else if (tid < NTRI + NP) {
      for (int w = 0; w < NW; w++) t += sV[w][tid - NTRI];
    }
    __syncthreads();
    if (tid < NTRI) sM[0][tid] = t;
    else if (tid < NTRI + NP) sV[0][tid - NTRI] = t;
    __syncthreads();
  }
  if (wid == 0) {
    const double ld = chol_solve<NP>(sM[0], sV[0], lane);
    if (lane == 0) s_ldet = ld;
  }
  __syncthreads();
  double co[NP];
#pragma unroll
  for (int i = 0; i < NP; i++) co[i] = sV[0][i];
  double rss = 0;
  for (int p = tid; p < npix; p += TB_THREADS) {
    const double tn = stash ? tnbuf[p] : templ_at(p) * einv[p];
    double mval = 0;
#pragma unroll
    for (int i = 0; i < NP; i++) mval = fma(co[i], __ldg(s.P + i * s.pstride + b0 + p) * tn, mval);
    const double r = dn[p] - mval;
    rss = fma(r, r, rss);
  }
  rss = block_sum(rss, red);
  if (tid == 0) {
    const double chi = s_ldet + s.sumlog2[obj] + rss;
    if (!isfinite(chi)) st |= RVS_ST_NOT_PD;
    s.chisq[k] = chi;
    s.status[k] = st;
  }
}

template <typename GT, int NV, int NP>
static int launch_fused_one(const FusedArgs &fa, int K, size_t smem, cudaStream_t st) {
  auto kern = chisq_fused_kernel<GT, NV, NP>;
  RVS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<K, TB_THREADS, smem, st>>>(fa);
  RVS_LAUNCH_OK();
  return 0;
}

template <int NP>
static int launch_fused_np(const FusedArgs &fa, int grid_f64, int K, size_t smem,
                           cudaStream_t st) {
  if (grid_f64) return launch_fused_one<double, 0, NP>(fa, K, smem, st);
  switch (fa.t.nvert) {
    case 16: return launch_fused_one<float, 16, NP>(fa, K, smem, st);
    case 5: return launch_fused_one<float, 5, NP>(fa, K, smem, st);
    default: return launch_fused_one<float, 0, NP>(fa, K, smem, st);
  }
}

}  // namespace rvs

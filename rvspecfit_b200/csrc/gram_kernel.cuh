// Fused optimiser-phase evaluation, stage B: one CTA per item.  Reads the
// resampled, error-normalised template T/sigma written by stage A, accumulates
// the continuum normal equations (per-thread register accumulators, transposed
// warp reduction, fixed-order cross-warp sum), solves them by Cholesky and
// returns 2 sum ln L_ii + 2 sum ln sigma + |D - a^T G|^2 (spec_fit.py:205-249).
#pragma once
#include "chisq_device.cuh"

namespace rvs {

constexpr int GR_THREADS = 128;
constexpr int GR_WARPS = GR_THREADS / 32;

struct GramArgs {
  const double *tn;
  int64_t tn_stride;
  const double *dn, *sumlog2;
  const int64_t *off;
  const int32_t *oix;
  const double *P;
  int64_t pstride;
  const int64_t *boff;
  double *chisq;
  int32_t *status;
};

template <int NP>
__global__ void __launch_bounds__(GR_THREADS) gram_kernel(GramArgs a) {
  constexpr int NTRI = NP * (NP + 1) / 2;
  constexpr int RSPLIT = NP > 10 ? 10 : NP;
  __shared__ double sM[GR_WARPS][NTRI];
  __shared__ double sV[GR_WARPS][NP];
  __shared__ double red[GR_WARPS];
  __shared__ double s_ldet;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int k = blockIdx.x;
  const int obj = a.oix[k];
  const int64_t p0 = a.off[obj];
  const int npix = (int)(a.off[obj + 1] - p0);
  const int64_t b0 = a.boff[obj];
  const double *tn = a.tn + (int64_t)k * a.tn_stride;
  const double *dn = a.dn + p0;
  {
    GramAcc<NP, 0, RSPLIT> acc;
    acc.zero();
    double v[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) v[i] = 0;
    for (int p = tid; p < npix; p += GR_THREADS) {
      double g[NP];
      load_basis<NP>(a.P, a.pstride, b0 + p, tn[p], g);
      const double d = dn[p];
#pragma unroll
      for (int i = 0; i < NP; i++) v[i] = fma(g[i], d, v[i]);
      acc.add(g);
    }
    acc.reduce_store(sM[wid], lane);
    warp_reduce_store<NP>(v, sV[wid], lane);
  }
  if (NP > RSPLIT) {
    GramAcc<NP, RSPLIT, NP> acc;
    acc.zero();
    for (int p = tid; p < npix; p += GR_THREADS) {
      double g[NP];
      load_basis<NP>(a.P, a.pstride, b0 + p, tn[p], g);
      acc.add(g);
    }
    acc.reduce_store(sM[wid], lane);
  }
  __syncthreads();
  {  // cross-warp sums in fixed order
    double t = 0;
    for (int e = tid; e < NTRI + NP; e += GR_THREADS) {
      t = 0;
      if (e < NTRI) {
        for (int w = 0; w < GR_WARPS; w++) t += sM[w][e];
      } else {
        for (int w = 0; w < GR_WARPS; w++) t += sV[w][e - NTRI];
      }
      // each entry e is read (all warps' copies) and rewritten (warp 0's copy) by
      // this thread only, so no barrier is needed between the two
      if (e < NTRI) sM[0][e] = t; else sV[0][e - NTRI] = t;
    }
  }
  __syncthreads();
  if (wid == 0) {
    const double ld = chol_solve<NP>(sM[0], sV[0], lane);
    if (lane == 0) s_ldet = ld;
  }
  __syncthreads();
  double co[NP];
#pragma unroll
  for (int i = 0; i < NP; i++) co[i] = sV[0][i];
  double rss = 0;
  for (int p = tid; p < npix; p += GR_THREADS) {
    const double t = tn[p];
    double mval = 0;
#pragma unroll
    for (int i = 0; i < NP; i++) mval = fma(co[i], __ldg(a.P + i * a.pstride + b0 + p) * t, mval);
    const double r = dn[p] - mval;
    rss = fma(r, r, rss);
  }
  rss = warp_sum(rss);
  if (lane == 0) red[wid] = rss;
  __syncthreads();
  if (tid == 0) {
    double t = 0;
    for (int w = 0; w < GR_WARPS; w++) t += red[w];
    const double chi = s_ldet + a.sumlog2[obj] + t;
    a.chisq[k] = chi;
    if (!isfinite(chi)) a.status[k] |= RVS_ST_NOT_PD;
  }
}

}  // namespace rvs

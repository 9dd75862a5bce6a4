// Fused optimiser-phase evaluation, stage B: one CTA per item.  Reads the
// resampled, error-normalised template T/sigma written by stage A, accumulates
// the continuum normal equations (per-thread register accumulators, transposed
// warp reduction, fixed-order cross-warp sum), solves them by Cholesky and
// returns 2 sum ln L_ii + 2 sum ln sigma + |D - a^T G|^2 (spec_fit.py:205-249).
//
// Every sweep over the pixels is fed by a register-free prefetch queue: each
// thread issues cp.async (LDGSTS) copies of its next GR_DEPTH pixels -- the
// pixel-major basis row (16-byte copies), T/sigma and D -- into its own
// shared-memory slots and consumes them in order, so the L2/HBM latency of one
// pixel hides behind the arithmetic of the previous ones without spending
// registers on loads in flight (the 65 accumulators already take 130).  A thread
// only ever reads what it copied itself: no barriers in the sweeps.
#pragma once
#include "chisq_device.cuh"

namespace rvs {

constexpr int GR_THREADS = 128;
constexpr int GR_WARPS = GR_THREADS / 32;
constexpr int GR_DEPTH = 4;

struct GramArgs {
  const double *tn;
  int64_t tn_stride;
  const double *dn, *sumlog2;
  const int64_t *off, *goff;
  const int32_t *oix;
  const double *P;  // pixel-major [pixel][npp]
  int npp;
  double *chisq;
  int32_t *status;
};

// doubles per (thread, stage) slot: basis row, T/sigma, D; (slot/2) odd so that the
// 16-byte shared-memory accesses of a quarter-warp fall in distinct banks
__host__ __device__ constexpr int gram_slot(int np) {
  const int npp = (np + 1) & ~1;
  return ((npp + 2) / 2) % 2 ? npp + 2 : npp + 4;
}

__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(double *dst_smem, const double *src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ring layout: [stage][thread][slot]; body(g, d) gets g[i] = P_i * T/sigma and D
template <int NP, typename F>
__device__ __forceinline__ void pixel_sweep(const double *Pb, const double *tn, const double *dn,
                                            int npix, double *ring, F &&body) {
  constexpr int NPP = (NP + 1) & ~1;
  constexpr int SLOT = gram_slot(NP);
  const int tid = threadIdx.x;
  const int niter = (npix + GR_THREADS - 1) / GR_THREADS;
  double *mine = ring + (size_t)tid * SLOT;
  auto issue = [&](int it, int stage) {
    const int p = it * GR_THREADS + tid;
    if (p < npix) {
      double *slot = mine + (size_t)stage * GR_THREADS * SLOT;
      const double *src = Pb + (int64_t)p * NPP;
#pragma unroll
      for (int i = 0; i < NPP / 2; i++) cp_async16(slot + 2 * i, src + 2 * i);
      cp_async8(slot + NPP, tn + p);
      cp_async8(slot + NPP + 1, dn + p);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int u = 0; u < GR_DEPTH - 1; u++) issue(u, u);
  for (int it0 = 0; it0 < niter; it0 += GR_DEPTH) {
#pragma unroll
    for (int u = 0; u < GR_DEPTH; u++) {
      const int it = it0 + u;
      issue(it + GR_DEPTH - 1, (u + GR_DEPTH - 1) % GR_DEPTH);
      cp_async_wait<GR_DEPTH - 1>();
      const int p = it * GR_THREADS + tid;
      if (p < npix) {
        const double2 *slot = reinterpret_cast<const double2 *>(mine + (size_t)u * GR_THREADS * SLOT);
        const double2 td = slot[NPP / 2];
        double g[NP];
#pragma unroll
        for (int i = 0; i < NP / 2; i++) {
          const double2 v = slot[i];
          g[2 * i] = v.x * td.x;
          g[2 * i + 1] = v.y * td.x;
        }
        if (NP & 1) g[NP - 1] = slot[NP / 2].x * td.x;
        body(g, td.y);
      }
    }
  }
  cp_async_wait<0>();
}

template <int NP>
__global__ void __launch_bounds__(GR_THREADS) gram_kernel(GramArgs a) {
  constexpr int NTRI = NP * (NP + 1) / 2;
  constexpr int RSPLIT = NP > 10 ? 10 : NP;
  extern __shared__ __align__(16) double ring[];
  __shared__ double sM[GR_WARPS][NTRI];
  __shared__ double sV[GR_WARPS][NP];
  __shared__ double red[GR_WARPS];
  __shared__ double s_ldet;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int k = blockIdx.x;
  const int obj = a.oix[k];
  if (obj < 0) {  // the object has no spectrum in this setup (uniform over the CTA)
    if (tid == 0) a.chisq[k] = 0;
    return;
  }
  const int64_t p0 = a.off[obj];
  const int npix = (int)(a.off[obj + 1] - p0);
  const double *Pb = a.P + a.goff[obj] * a.npp;
  const double *tn = a.tn + (int64_t)k * a.tn_stride;
  const double *dn = a.dn + p0;
  {
    GramAcc<NP, 0, RSPLIT> acc;
    acc.zero();
    double v[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) v[i] = 0;
    pixel_sweep<NP>(Pb, tn, dn, npix, ring, [&](const double (&g)[NP], double d) {
#pragma unroll
      for (int i = 0; i < NP; i++) v[i] = fma(g[i], d, v[i]);
      acc.add(g);
    });
    acc.reduce_store(sM[wid], lane);
    warp_reduce_store<NP>(v, sV[wid], lane);
  }
  if (NP > RSPLIT) {
    GramAcc<NP, RSPLIT, NP> acc;
    acc.zero();
    pixel_sweep<NP>(Pb, tn, dn, npix, ring, [&](const double (&g)[NP], double) { acc.add(g); });
    acc.reduce_store(sM[wid], lane);
  }
  __syncthreads();
  {  // cross-warp sums in fixed order
    for (int e = tid; e < NTRI + NP; e += GR_THREADS) {
      double t = 0;
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
  pixel_sweep<NP>(Pb, tn, dn, npix, ring, [&](const double (&g)[NP], double d) {
    double mval = 0;
#pragma unroll
    for (int i = 0; i < NP; i++) mval = fma(co[i], g[i], mval);
    const double r = d - mval;
    rss = fma(r, r, rss);
  });
  rss = warp_sum(rss);
  if (lane == 0) red[wid] = rss;
  __syncthreads();
  if (tid == 0) {
    double t = 0;
    for (int w = 0; w < GR_WARPS; w++) t += red[w];
    const double chi = s_ldet + a.sumlog2[obj] + t;
    a.chisq[k] = chi;
    if (!isfinite(chi)) atomicOr(a.status + k, RVS_ST_NOT_PD);
  }
}

}  // namespace rvs

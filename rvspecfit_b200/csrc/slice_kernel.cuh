// Fused optimiser-phase evaluation, stage A: one CTA per (slice, item).
// A slice is a contiguous range of template knots (and the observed pixels that
// fall on them at the item's velocity).  The CTA gathers only its window of the
// grid rows, exponentiates, broadens, solves the spline on the window (halo of
// SPL_HALO knots each side, exact natural boundary where the window touches the
// template ends), resamples onto its pixels and writes T/sigma.  Stage B
// (gram_kernel.cuh) does the continuum solve.  Many small CTAs of different
// phases are resident per SM, which is what hides the latency of each phase.
#pragma once
#include "chisq_device.cuh"
#include "template_device.cuh"

namespace rvs {

constexpr int SL_THREADS = 128;
constexpr int SL_WARPS = SL_THREADS / 32;

struct SliceArgs {
  // template side
  const void *grid;
  int64_t ld;
  int npix_t;
  const int32_t *ids;
  const double *w;
  int nvert;
  const double *vsini;
  const double *lam_t, *h, *hinv, *cp, *winv;
  double lnstep;
  int log_spec, log_step;
  double x0, xlast, q0, qstep_inv;
  // observed side
  const double *lam, *loglam, *einv;
  const int64_t *off;
  const int32_t *oix;
  const double *vels;
  // outputs
  double *tn;
  int64_t tn_stride;
  int32_t *status;
  int wcap, tapcap;
  double *dbg;  // optional dump of slice 0 / item 0 (rvs_set_debug_buffer), else NULL
};

// inclusive scan over the CTA of affine maps x -> A + B x, composed left to
// right (thread t's map is applied after thread t-1's).  Returns the value the
// map of all earlier threads gives for input `xin` (the exclusive prefix).
__device__ __forceinline__ double affine_exclusive_scan(double A, double B, double xin,
                                                        double *sA, double *sB) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double Ap = __shfl_up_sync(0xffffffffu, A, o);
    const double Bp = __shfl_up_sync(0xffffffffu, B, o);
    if (lane >= o) {
      A = fma(B, Ap, A);
      B = B * Bp;
    }
  }
  __syncthreads();
  if (lane == 31) { sA[wid] = A; sB[wid] = B; }
  __syncthreads();
  // value entering this warp
  double x = xin;
  for (int w = 0; w < wid; w++) x = fma(sB[w], x, sA[w]);
  // value entering this thread: inclusive result of the previous lane
  const double Ai = __shfl_up_sync(0xffffffffu, A, 1);
  const double Bi = __shfl_up_sync(0xffffffffu, B, 1);
  return lane == 0 ? x : fma(Bi, x, Ai);
}

template <typename GT, int NV>
__global__ void __launch_bounds__(SL_THREADS, 4) slice_kernel(SliceArgs a) {
  extern __shared__ double sm[];
  double *B0 = sm, *B1 = sm + a.wcap, *taps = sm + 2 * a.wcap;
  __shared__ int32_t s_ids[32];
  __shared__ double s_w[32];
  __shared__ double red[SL_WARPS], sA[SL_WARPS], sB[SL_WARPS];
  __shared__ int s_flag;
  const int tid = threadIdx.x;
  const int k = blockIdx.y, S = gridDim.x, s = blockIdx.x;
  const int n = a.npix_t;
  const int obj = a.oix[k];
  const int64_t p0 = a.off[obj];
  const int npix = (int)(a.off[obj + 1] - p0);
  const double *lam = a.lam + p0, *ql = (a.log_step ? a.loglam : a.lam) + p0;
  const double beta = a.vels[k] / RVS_C_KMS;
  const double f = sqrt((1 - beta) / (1 + beta));
  const double qf = a.log_step ? log(f) : 0.0;
  auto pos_of = [&](int p) -> int {
    const double q = a.log_step ? ql[p] + qf : lam[p] * f;
    const int pos = (int)((q - a.q0) * a.qstep_inv);
    return max(0, min(pos, n - 2));
  };
  const int posmin = pos_of(0), posmax = pos_of(npix - 1);
  const int nk = posmax + 1 - posmin;
  const int c0 = posmin + (int)((int64_t)nk * s / S), c1 = posmin + (int)((int64_t)nk * (s + 1) / S);
  if (tid < a.nvert) {
    s_ids[tid] = a.ids[(int64_t)k * a.nvert + tid];
    s_w[tid] = a.w[(int64_t)k * a.nvert + tid];
  }
  if (tid == 0) s_flag = 0;
  if (s == 0 && tid == 0) {
    const double xa = lam[0] * f, xb = lam[npix - 1] * f;
    if (xa < a.x0 || xb < a.x0 || xa >= a.xlast || xb >= a.xlast)
      atomicOr(a.status + k, RVS_ST_RANGE);
  }
  if (c1 <= c0) return;  // empty slice (block-uniform)
  // pixel range of the slice: pos is non-decreasing in p.  Two rounds of a
  // block-wide vote instead of a serial binary search.
  auto first_px_with_pos_ge = [&](int c) -> int {
    const int pt = (int)((int64_t)npix * tid / SL_THREADS);
    const int nfalse = __syncthreads_count(pos_of(pt) < c);
    int lo = (nfalse == 0) ? 0 : (int)((int64_t)npix * (nfalse - 1) / SL_THREADS) + 1;
    const int hi = (nfalse == SL_THREADS) ? npix : (int)((int64_t)npix * nfalse / SL_THREADS);
    while (true) {
      const int p = lo + tid;
      const int cnt = __syncthreads_count(p < hi && pos_of(p) < c);
      lo += cnt;
      if (cnt < SL_THREADS || lo >= hi) break;
    }
    return lo;
  };
  const int plo = first_px_with_pos_ge(c0);
  const int phi = (s == S - 1) ? npix : first_px_with_pos_ge(c1);
  // rotation kernel size
  const double vs = a.vsini ? a.vsini[k] : 0.0;
  double R = 0;
  int kmax = 0;
  bool conv = false;
  if (vs > 0) {
    R = (vs / RVS_C_KMS) / a.lnstep;
    if (R >= 1e-9) {
      conv = true;
      kmax = (int)ceil(R + 1);
      if (kmax > a.tapcap) { kmax = a.tapcap; if (tid == 0) s_flag |= RVS_ST_TAPS; }
    }
  }
  // knot windows (global indices, inclusive)
  const int ya0 = max(0, c0 - SPL_HALO - 1), ya1 = min(n - 1, c1 + SPL_HALO + 2);
  const int g0 = max(0, ya0 - kmax), g1 = min(n - 1, ya1 + kmax);
  const int g0a = g0 & ~3;
  const int W0 = ((g1 + 1 - g0a) + 3) & ~3;
  if (W0 > a.wcap) {  // does not fit: the caller re-evaluates this item on the general path
    if (tid == 0) atomicOr(a.status + k, RVS_ST_LIMIT);
    return;
  }
  __syncthreads();
  const bool f32row = a.nvert > 1 && s_ids[1] < 0;
  __syncthreads();
  if (f32row && tid >= 1 && tid < a.nvert) { s_ids[tid] = s_ids[0]; s_w[tid] = 0; }
  __syncthreads();
  // ---- gather + exp over the window [g0a, g0a + W0)
  {
    constexpr int VEC = RowLoader<GT>::VEC;
    const bool round32 = f32row && sizeof(GT) == 4 && a.log_spec;
    const GT *base = static_cast<const GT *>(a.grid) + g0a;
    const int nv = NV > 0 ? NV : a.nvert;
    int bad = 0;
    for (int q = tid; q < W0 / VEC; q += SL_THREADS) {
      double acc[4] = {0, 0, 0, 0};
      if (NV > 0) {
        double r[NV > 0 ? NV : 1][4];
#pragma unroll
        for (int j = 0; j < NV; j++) RowLoader<GT>::load(base + (int64_t)s_ids[j] * a.ld, q, r[j]);
#pragma unroll
        for (int j = 0; j < NV; j++) {
#pragma unroll
          for (int e = 0; e < VEC; e++) acc[e] = fma(s_w[j], r[j][e], acc[e]);
        }
      } else {
        for (int j = 0; j < nv; j++) {
          double r[4];
          RowLoader<GT>::load(base + (int64_t)s_ids[j] * a.ld, q, r);
#pragma unroll
          for (int e = 0; e < VEC; e++) acc[e] = fma(s_w[j], r[e], acc[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < VEC; e++) {
        const int g = g0a + q * VEC + e;
        double y = a.log_spec ? exp(acc[e]) : acc[e];
        if (round32) y = (double)(float)y;
        if (g >= n) y = 0;  // row padding, never a knot
        else if (!(fabs(y) <= 1e100)) bad = 1;
        B0[q * VEC + e] = y;
      }
    }
    if (bad) s_flag |= RVS_ST_TEMPLATE_BAD;  // benign race: all writers set the same bit
  }
  __syncthreads();
  // ---- rotational broadening onto [ya0, ya1]
  const int WY = ya1 - ya0 + 1;
  double *Y, *Dz;  // Y[j] = y at knot ya0 + j ; Dz: spline scratch
  if (conv) {
    double part = 0;
    for (int t = tid; t <= kmax; t += SL_THREADS) {
      const double w = rot_weight(t, R);
      taps[t] = w;
      part += (t == 0) ? w : 2 * w;
    }
    const double tot = block_sum(part, red);
    __syncthreads();
    for (int t = tid; t <= kmax; t += SL_THREADS) taps[t] = taps[t] / tot;
    __syncthreads();
    for (int j = tid; j < WY; j += SL_THREADS) {
      const int g = ya0 + j;            // global knot
      const int c = g - g0a;            // index in B0
      double sum = taps[0] * B0[c];
      for (int t = 1; t <= kmax; t++) {
        const double lo = (g - t >= 0) ? B0[c - t] : 0.0;
        const double hi = (g + t < n) ? B0[c + t] : 0.0;
        sum = fma(taps[t], lo + hi, sum);
      }
      B1[j] = sum;
    }
    __syncthreads();
    Y = B1;
    Dz = B0 + 1;
  } else {
    Y = B0 + (ya0 - g0a);
    Dz = B1 + 1;
  }
  // Dz[r] holds d of row kr0 + r, then z at knot kr0 + r + 1; with the two guard
  // slots Dz[-1] (knot kr0) and Dz[nrow] (knot kr1 + 1), z at knot g is Dz[g - 1 - kr0]
  // for every knot the resampling can touch.  The guards are the natural boundary
  // values z[0] = z[n-1] = 0 whenever the window reaches an end of the template.
  // ---- spline rows kr0 <= k < kr1 (row k couples knots k, k+1, k+2; unknown z[k+1])
  const int m = n - 2;
  const int kr0 = ya0, kr1 = min(m, ya1 - 1);
  const int nrow = kr1 - kr0;
  const int ch = (nrow + SL_THREADS - 1) / SL_THREADS;
  const int r0 = min(nrow, tid * ch), r1 = min(nrow, r0 + ch);
  if (tid == 0) { Dz[-1] = 0; Dz[nrow] = 0; }
  {
    // forward: d_k = (u_k - h_k d_{k-1}) w_k  ==  a_k + b_k d_{k-1}
    double A = 0, Bc = 1;
    for (int r = r0; r < r1; r++) {
      const int kk = kr0 + r;
      const double bl = (Y[r + 1] - Y[r]) * __ldg(a.hinv + kk);
      const double br = (Y[r + 2] - Y[r + 1]) * __ldg(a.hinv + kk + 1);
      const double wv = __ldg(a.winv + kk);
      const double ak = 6 * (br - bl) * wv, bk = -(__ldg(a.h + kk) * wv);
      A = fma(bk, A, ak);
      Bc = bk * Bc;
    }
    double d = affine_exclusive_scan(A, Bc, 0.0, sA, sB);  // d_{kr0-1} := 0 (exact at kr0 == 0)
    for (int r = r0; r < r1; r++) {
      const int kk = kr0 + r;
      const double bl = (Y[r + 1] - Y[r]) * __ldg(a.hinv + kk);
      const double br = (Y[r + 2] - Y[r + 1]) * __ldg(a.hinv + kk + 1);
      d = (6 * (br - bl) - __ldg(a.h + kk) * d) * __ldg(a.winv + kk);
      Dz[r] = d;
    }
  }
  __syncthreads();
  {
    // backward: z_{k+1} = d_k - cp_k z_{k+2}; threads in reverse order
    const int rt = SL_THREADS - 1 - tid;  // reversed thread rank
    const int q0r = min(nrow, rt * ch), q1r = min(nrow, q0r + ch);
    double A = 0, Bc = 1;
    for (int r = q1r - 1; r >= q0r; r--) {
      const double ck = -__ldg(a.cp + kr0 + r);
      A = fma(ck, A, Dz[r]);
      Bc = ck * Bc;
    }
    // scan in reversed order: thread `tid` plays rank `tid` of the reversed sequence
    // (rank 0 owns the LAST rows), which is why the chunk above uses rt
    double zz = affine_exclusive_scan(A, Bc, 0.0, sA, sB);  // z_{kr1+1} := 0 (exact at kr1 == m)
    for (int r = q1r - 1; r >= q0r; r--) {
      zz = Dz[r] - __ldg(a.cp + kr0 + r) * zz;
      Dz[r] = zz;  // = z at knot kr0 + r + 1
    }
  }
  __syncthreads();
  if (a.dbg && k == 0 && s == 0) {
    if (tid == 0) {
      a.dbg[0] = ya0; a.dbg[1] = ya1; a.dbg[2] = kr0; a.dbg[3] = kr1; a.dbg[4] = c0;
      a.dbg[5] = c1; a.dbg[6] = plo; a.dbg[7] = phi; a.dbg[8] = g0a; a.dbg[9] = W0;
    }
    for (int j = tid; j < WY; j += SL_THREADS) a.dbg[16 + j] = Y[j];
    for (int r = tid; r < nrow; r += SL_THREADS) a.dbg[16 + a.wcap + r] = Dz[r];
  }
  // ---- resample onto the slice's pixels
  const double *einv = a.einv + p0;
  double *tn = a.tn + (int64_t)k * a.tn_stride;
  for (int p = plo + tid; p < phi; p += SL_THREADS) {
    const double x = lam[p] * f;
    const int pos = pos_of(p);
    const int j = pos - ya0;
    const double y0v = Y[j], y1v = Y[j + 1];
    const double z0v = Dz[pos - 1 - kr0], z1v = Dz[pos - kr0];
    const double xl = __ldg(a.lam_t + pos), xr = __ldg(a.lam_t + pos + 1);
    const double hh = __ldg(a.h + pos), hi = __ldg(a.hinv + pos);
    const double t1 = hi * (1. / 6), t2 = hh * (1. / 6);
    const double Ac = z1v * t1, Bq = z0v * t1;
    const double Cc = y1v * hi - z1v * t2, Dc = y0v * hi - z0v * t2;
    const double dl = x - xl, dr = xr - x;
    tn[p] = (Ac * dl * dl * dl + Bq * dr * dr * dr + Cc * dl + Dc * dr) * einv[p];
    if (a.dbg && k == 0 && s == 0 && p < 8) {
      double *o = a.dbg + 16 + 2 * a.wcap + p * 16;
      o[0] = pos; o[1] = x; o[2] = y0v; o[3] = y1v; o[4] = z0v; o[5] = z1v; o[6] = xl; o[7] = xr;
      o[8] = hh; o[9] = hi; o[10] = einv[p]; o[11] = tn[p]; o[12] = lam[p]; o[13] = f;
    }
  }
  __syncthreads();
  if (tid == 0 && s_flag) atomicOr(a.status + k, s_flag);
}

}  // namespace rvs

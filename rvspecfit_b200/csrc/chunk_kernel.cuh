// Fused optimiser-phase evaluation, stage A ("chunk kernel"): one WARP per
// (item, chunk).  A chunk is a contiguous range of at most `C` template knots
// and the observed pixels that fall on them at the item's velocity.  The warp
//   1. gathers its window of the 2^d (or d+1) grid rows -- as 5-D TMA tiles of the
//      corner box on dense regular grids, with per-lane 16-byte cp.async copies
//      otherwise --, accumulates the corner-weighted sum in fp64 and exponentiates,
//   2. applies the rotational-broadening taps (prepared by prep_kernel),
//   3. solves the natural cubic spline on the window,
//   4. resamples onto its pixels and writes T/sigma.
// Warps never synchronise with each other (only __syncwarp / shuffles), so the
// 20-odd resident warps of an SM sit in different phases and the memory phase
// of one overlaps the fp64 phases of the others.  Stage B (gram_kernel.cuh)
// does the continuum solve.
//
// Spline solve.  The template knots are uniform in x or in ln x (validated by
// rvs_knot_info, as the reference's evaler requires, spliner.c:84-96), so
// h[k+1] = r h[k] with one ratio r.  In the scaled unknown s_k = z_k h_k^2 / 6
// the reference's tridiagonal system (spliner.c:21-41)
//     h_k z_k + 2 (h_k + h_{k+1}) z_{k+1} + h_{k+1} z_{k+2} = 6 (b_{k+1} - b_k)
// becomes the constant-coefficient system
//     s_k + c1 s_{k+1} + c2 s_{k+2} = dy_{k+1} / r - dy_k,
//     c1 = 2 (1 + r) / r^2,  c2 = 1 / r^3,
// whose Thomas pivots converge to a constant within ~15 rows of knot 0
// (table wtab[] holds the exact ones there).  No per-knot tables are read.  The
// window carries CK_HALO extra rows on each side: the influence of the
// artificial window ends decays as 0.268^k per row (2e-14 after 24 rows, on a
// term that is ~1e-2 of the value: below fp64 rounding of the resampled template;
// exact natural boundary where the window touches an end of the template).
// Evaluation in [x_i, x_{i+1}):  u = (x - x_i)/h_i, v = (x_{i+1} - x)/h_i,
//     T = v y_i + u y_{i+1} + s_i (v^3 - v) + (s_{i+1}/r^2) (u^3 - u),
// algebraically the reference's A dl^3 + B dr^3 + C dl + D dr (spliner.c:97-106).
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the encoder is resolved at run time, api.cu)

#include "chisq_device.cuh"
#include "template_device.cuh"

namespace rvs {

constexpr int CK_WARPS = 4;
constexpr int CK_THREADS = CK_WARPS * 32;
constexpr int CK_NT = 32;  // rows next to knot 0 with tabulated pivots
// spline halo of a chunk window: the influence of the artificial window ends decays
// as 0.268^k per knot: 2e-14 after 24 knots, times h^2 z / y ~ 1e-2 -> below fp64
// rounding of the resampled template
#ifndef RVS_CK_HALO
#define RVS_CK_HALO 24
#endif
constexpr int CK_HALO = RVS_CK_HALO;
#ifndef RVS_CK_RING
#define RVS_CK_RING 8
#endif
constexpr int CK_RING = RVS_CK_RING;  // slots (of 512 B) of the gather prefetch ring
#ifndef RVS_CK_MINB
#define RVS_CK_MINB 6
#endif
constexpr int CK_MINB = RVS_CK_MINB;  // resident CTAs per SM the register budget allows
#ifndef RVS_CK_MINB_TMA
#define RVS_CK_MINB_TMA 7
#endif
constexpr int CK_MINB_TMA = RVS_CK_MINB_TMA;  // same for the TMA variant
// TMA gather (dense regular 4-D grids, rvs_gridbox): one cp.async.bulk.tensor of a
// [TMA_COLS px][2][2][2][1] box brings 8 of the 16 corner rows of a column block
// into one ring stage; two stages (the two halves of the first grid dimension).
#ifndef RVS_TMA_COLS
#define RVS_TMA_COLS 64
#endif
constexpr int TMA_COLS = RVS_TMA_COLS;     // knots per column block (32 lanes x 2 or 4)
#ifndef RVS_TMA_ROWS
#define RVS_TMA_ROWS 8
#endif
constexpr int TMA_ROWS = RVS_TMA_ROWS;     // corner rows per stage: 8 (box 2x2x2x1) or 4 (2x2x1x1)
constexpr int TMA_NQ = 16 / TMA_ROWS;      // stages per column block
#ifndef RVS_TMA_NSTG
#define RVS_TMA_NSTG 2
#endif
constexpr int TMA_NSTG = RVS_TMA_NSTG;     // ring stages per warp (NSTG - 1 in flight)
constexpr int TMA_STG_BYTES = TMA_COLS * TMA_ROWS * 4;
constexpr int TMA_RING_DOUBLES = TMA_NSTG * TMA_STG_BYTES / 8;
struct TrueTag { static constexpr bool value = true; };
struct FalseTag { static constexpr bool value = false; };
template <int N>
struct IntTag { static constexpr int value = N; };
constexpr unsigned FULL = 0xffffffffu;

struct ChunkArgs {
  // template side
  const void *grid;
  int64_t ld;
  int npix_t;
  const int32_t *ids;
  const double *w;
  int nvert;
  const double *taps;   // [K, tapstride] normalised one-sided weights (prep_kernel) or NULL
  int tapstride;
  // per-item records written by prep_kernel
  const double *rec;      // [K][2]  Doppler factor f, ln f (0 on linear knot grids)
  const int32_t *irec;    // [K][8]  posmin, nk, S (chunks the item really has), kmax,
                          //         grid position of the item's first corner (TMA gather)
  const int32_t *pbound;  // [K][nch+1] first pixel of every chunk, pbound[S] = npix
  const double *lam_t, *hinv;
  int log_spec, log_step;
  double x0, xlast, q0, qstep_inv;
  double rinv, r2inv, c2, winv_inf;
  double wtab[CK_NT];
  // observed side
  const double *lam, *loglam;  // grid pools (see rvs_obs)
  const double *einv;          // object pool
  const int64_t *off, *goff;
  const int32_t *oix;
  const double *vels;
  // outputs
  double *tn;
  int64_t tn_stride;
  int32_t *status;
  int wcap;   // doubles of smem buffer B0 per warp
  int wcap1;  // doubles of smem buffer B1 per warp (>= wcap; holds the gather ring)
  int nch;   // chunks per item (upper bound; surplus chunks exit)
  int C;     // knots per chunk (upper bound)
  int K;
  int raw_out;  // 1: write T, not T/sigma (a resolution matrix is applied next, resol_apply_kernel)
};

// Resolution-matrix stage of the fused evaluation (spec_fit.py:922-929): tn[k][p] =
// (R_obj T_k)[p] / sigma[p] from the resampled template `raw` of item k and the
// object's band rows (rvs_obs).  One thread per (item, pixel); the rows of T stay in L2
// between the chunk kernel and this one.
struct ResolArgs {
  const double *raw;
  double *tn;
  int64_t tn_stride;
  const double *resol, *einv;
  const int32_t *resol_offs;
  int nresol;
  const int64_t *off;
  const int32_t *oix;
};
__global__ void __launch_bounds__(256) resol_apply_kernel(ResolArgs a) {
  const int k = blockIdx.x;
  const int obj = a.oix[k];
  if (obj < 0) return;
  const int64_t p0 = a.off[obj];
  const int npix = (int)(a.off[obj + 1] - p0);
  const int p = blockIdx.y * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const double *raw = a.raw + (int64_t)k * a.tn_stride;
  const double *rb = a.resol + p0 * a.nresol + p;
  double s = 0;
  for (int d = 0; d < a.nresol; d++) {
    const int pp = p + __ldg(a.resol_offs + d);
    if (pp < 0 || pp >= npix) continue;
    s = fma(__ldg(rb + (int64_t)d * npix), raw[pp], s);
  }
  a.tn[(int64_t)k * a.tn_stride + p] = s * a.einv[p0 + p];
}

// Per-item preparation, one warp per item: Doppler factor, the knot range the
// object covers at that velocity and its split into chunks, the first pixel of
// every chunk, range / capacity status, and the one-sided normalised rotation
// taps (spec_fit.py:565-625).  Everything a chunk warp would otherwise recompute
// 32-fold per chunk.
struct PrepArgs {
  const double *vsini;  // may be NULL
  double lnstep;
  int tapcap, tapstride, K;
  double *taps;
  // geometry
  const double *lam_t;
  int npix_t, log_step;
  double x0, xlast, q0, qstep_inv, qstep;
  const double *lam, *loglam;
  const int64_t *off, *goff;
  const int32_t *oix;
  const double *vels;
  int nch, C;
  double *rec;
  int32_t *irec, *pbound;
  int32_t *status;
  // TMA gather: first corner id -> grid position (blen = lengths of dims 1..3)
  const int32_t *ids;
  int nvert, box;
  int blen[3];
};

// The arms of an evaluation call (the setups of a multi-arm object: three for DESI) go
// through ONE launch of every kernel of the call: blockIdx.y (.z for the Gram kernels)
// selects the arm's argument record in the launch's constant parameter block.  A call is
// then 6 kernels instead of 16, its small fixed-latency kernels are paid once instead of
// once per arm, and the tail of one arm's grid is filled by the others.
constexpr int RVS_MAX_ARMS = 4;
struct PrepArgsM {
  PrepArgs a[RVS_MAX_ARMS];
};

__global__ void __launch_bounds__(128) prep_kernel(const __grid_constant__ PrepArgsM m) {
  const PrepArgs &a = m.a[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= a.K) return;
  int32_t *irec = a.irec + (int64_t)k * 8;
  const int obj = a.oix[k];
  if (obj < 0) {  // no spectrum in this setup: no chunks
    if (lane < 8) irec[lane] = 0;
    if (lane == 0) a.status[k] = 0;
    return;
  }
  // ---- rotation taps
  int kmax = 0, st0 = 0;
  const double vs = a.vsini ? a.vsini[k] : 0.0;
  if (vs > 0) {
    const double R = (vs / RVS_C_KMS) / a.lnstep;
    if (R >= 1e-9) {
      kmax = (int)ceil(R + 1);
      if (kmax > a.tapcap) {
        kmax = a.tapcap;
        st0 |= RVS_ST_TAPS;
      }
      double *t = a.taps + (int64_t)k * a.tapstride;
      double part = 0;
      for (int j = lane; j <= kmax; j += 32) {
        const double wv = rot_weight(j, R);
        t[j] = wv;
        part += (j == 0) ? wv : 2 * wv;
      }
      const double tot = warp_sum(part);
      __syncwarp();
      for (int j = lane; j <= kmax; j += 32) t[j] = t[j] / tot;
    }
  }
  // ---- geometry
  const int n = a.npix_t;
  const int64_t p0 = a.off[obj];
  const int npix = (int)(a.off[obj + 1] - p0);
  const int64_t gp0 = a.goff[obj];
  const double *lam = a.lam + gp0, *ql = (a.log_step ? a.loglam : a.lam) + gp0;
  const double beta = a.vels[k] / RVS_C_KMS;
  const double f = sqrt((1 - beta) / (1 + beta));
  const double qf = a.log_step ? log(f) : 0.0;
  // knot interval of the rest-frame coordinate q (ln x or x)
  auto pos_q = [&](double q) -> int {
    const int pos = (int)((q - a.q0) * a.qstep_inv);
    return max(0, min(pos, n - 2));
  };
  auto pos_of = [&](int p) -> int { return pos_q(a.log_step ? ql[p] + qf : lam[p] * f); };
  const double lam_first = lam[0], lam_last = lam[npix - 1];
  const double ql_first = a.log_step ? ql[0] : lam_first, ql_last = a.log_step ? ql[npix - 1] : lam_last;
  const int posmin = pos_q(a.log_step ? ql_first + qf : lam_first * f);
  const int posmax = pos_q(a.log_step ? ql_last + qf : lam_last * f);
  const int nk = posmax + 1 - posmin;
  const int S = min(a.nch, (nk + a.C - 1) / a.C);  // chunks this item really has
  if (lane == 0) {
    // the reference checks the first and last evaluation points (spliner.c:78-83)
    const double xa = lam_first * f, xb = lam_last * f;
    int st = st0;
    if (xa < a.x0 || xb < a.x0 || xa >= a.xlast || xb >= a.xlast) st |= RVS_ST_RANGE;
    if ((int64_t)a.nch * a.C < nk) st |= RVS_ST_LIMIT;
    a.status[k] = st;  // first kernel of the sequence: initialises the status word
    a.rec[2 * (int64_t)k] = f;
    a.rec[2 * (int64_t)k + 1] = qf;
    irec[0] = posmin; irec[1] = nk; irec[2] = S; irec[3] = kmax;
    if (a.box) {  // C-order position of node ids[0] in the dense grid
      int id = a.ids[(int64_t)k * a.nvert];
      const int p3 = id % a.blen[2]; id /= a.blen[2];
      const int p2 = id % a.blen[1]; id /= a.blen[1];
      const int p1 = id % a.blen[0]; id /= a.blen[0];
      irec[4] = id; irec[5] = p1; irec[6] = p2; irec[7] = p3;
    }
  }
  // ---- first pixel of every chunk: the first pixel whose knot interval is >= c
  // (pos is non-decreasing in p).  One lane per chunk boundary: a linear guess in
  // ln lambda (exact to a pixel on uniform and log-uniform pixel grids), then a
  // galloping bracket and a bisection, so any monotone pixel grid is handled.
  auto first_px_with_pos_ge = [&](int c) -> int {
    double gq;
    if (a.log_step) gq = (a.q0 + c * a.qstep - qf - ql_first) / (ql_last - ql_first);
    else gq = (__ldg(a.lam_t + c) / f - lam_first) / (lam_last - lam_first);
    int g = (int)(gq * (npix - 1));
    g = max(0, min(g, npix - 1));
    int lo, hi;  // pos_of(lo) < c (or lo == -1), pos_of(hi) >= c (or hi == npix)
    if (pos_of(g) >= c) {
      hi = g;
      int step = 1;
      while (hi - step >= 0 && pos_of(hi - step) >= c) { hi -= step; step *= 2; }
      lo = max(-1, hi - step);
    } else {
      lo = g;
      int step = 1;
      while (lo + step < npix && pos_of(lo + step) < c) { lo += step; step *= 2; }
      hi = min(npix, lo + step);
    }
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (pos_of(mid) >= c) hi = mid; else lo = mid;
    }
    return hi;
  };
  int32_t *pb = a.pbound + (int64_t)k * (a.nch + 1);
  for (int s = lane; s <= S; s += 32) {
    int v;
    if (s == 0) v = 0;
    else if (s == S) v = npix;
    else v = first_px_with_pos_ge(posmin + (int)((int64_t)nk * s / S));
    pb[s] = v;
  }
}

// Scans over the warp of affine maps x -> A + B x.
// up:   maps composed in lane order (lane l acts after lanes < l); returns the
//       value entering this lane when `xin` enters lane 0.
// down: maps composed in reverse lane order (lane l acts after lanes > l); returns
//       the value entering this lane when `xin` enters lane 31.
__device__ __forceinline__ double warp_affine_up(double A, double B, double xin, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double Ap = __shfl_up_sync(FULL, A, o);
    const double Bp = __shfl_up_sync(FULL, B, o);
    if (lane >= o) {
      A = fma(B, Ap, A);
      B = B * Bp;
    }
  }
  const double Ai = __shfl_up_sync(FULL, A, 1);
  const double Bi = __shfl_up_sync(FULL, B, 1);
  return lane == 0 ? xin : fma(Bi, xin, Ai);
}
__device__ __forceinline__ double warp_affine_down(double A, double B, double xin, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double Ap = __shfl_down_sync(FULL, A, o);
    const double Bp = __shfl_down_sync(FULL, B, o);
    if (lane + o < 32) {
      A = fma(B, Ap, A);
      B = B * Bp;
    }
  }
  const double Ai = __shfl_down_sync(FULL, A, 1);
  const double Bi = __shfl_down_sync(FULL, B, 1);
  return lane == 31 ? xin : fma(Bi, xin, Ai);
}

__device__ __forceinline__ void prefetch_l1(const void *p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

struct ChunkArgsM {
  ChunkArgs a[RVS_MAX_ARMS];
  alignas(64) CUtensorMap tmap[RVS_MAX_ARMS];
};

template <typename GT, int NV, bool TMA>
__global__ void __launch_bounds__(CK_THREADS, TMA ? CK_MINB_TMA : CK_MINB)
chunk_kernel(const __grid_constant__ ChunkArgsM arms) {
  const ChunkArgs &a = arms.a[blockIdx.y];
  const CUtensorMap &tmap = arms.tmap[blockIdx.y];
  extern __shared__ __align__(128) double sm[];
  __shared__ int64_t s_off[CK_WARPS][32];  // element offset of each grid row
  __shared__ double s_w[CK_WARPS][32];
  __shared__ __align__(8) uint64_t s_bar[CK_WARPS][TMA_NSTG];  // TMA gather: stage barriers
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t wg = (int64_t)blockIdx.x * CK_WARPS + wid;
  const int k = (int)(wg / a.nch), s = (int)(wg - (int64_t)k * a.nch);
  if (k >= a.K) return;
  double *B0 = sm + (size_t)wid * (a.wcap + a.wcap1), *B1 = B0 + a.wcap;
  const unsigned bar_s = (unsigned)__cvta_generic_to_shared(&s_bar[wid][0]);
  if (TMA && lane == 0) {  // before any load is outstanding: the fence has nothing to wait for
#pragma unroll
    for (int i = 0; i < TMA_NSTG; i++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s + 8 * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int n = a.npix_t;
  const int4 ir = __ldg(reinterpret_cast<const int4 *>(a.irec) + 2 * k);  // posmin, nk, S, kmax
  int4 bp = make_int4(0, 0, 0, 0);  // grid position of the first corner (TMA gather)
  if (TMA) bp = __ldg(reinterpret_cast<const int4 *>(a.irec) + 2 * k + 1);
  const int posmin = ir.x, nk = ir.y, S = ir.z;
  if (s >= S) return;
  const int c0 = posmin + (int)((int64_t)nk * s / S), c1 = posmin + (int)((int64_t)nk * (s + 1) / S);
  if (c1 <= c0) return;
  const int kmax = ir.w;
  // knot windows (global indices, inclusive): Y on [ya0, ya1], raw y on [ya0-kmax, ya1+kmax]
  // which may stick out of the template: those knots are the zero padding of the
  // reference's 'same' convolution (spec_fit.py:677-680)
  const int ya0 = max(0, c0 - CK_HALO - 1), ya1 = min(n - 1, c1 + CK_HALO + 2);
  const int w0 = (ya0 - kmax) & ~3;  // first knot of the window (multiple of 4, may be < 0)
  const int W0 = ((ya1 + kmax + 1 - w0) + 3) & ~3;
  if (W0 + 2 > a.wcap) {  // does not fit: the caller re-evaluates this item on the general path
    if (lane == 0) atomicOr(a.status + k, RVS_ST_LIMIT);
    return;
  }
  // ---- TMA gather: the first stages go out as soon as the window is known
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(B1);
  const uint64_t tm = reinterpret_cast<uint64_t>(&tmap);
  const int ncb = (W0 + TMA_COLS - 1) / TMA_COLS;
  const int nstep = TMA_NQ * ncb;  // step t = (column block t / NQ, part t % NQ), stage t % NSTG
  auto issue = [&](int t) {  // lane 0 only
    const int slot = t % TMA_NSTG;
    const unsigned bar = bar_s + 8 * slot, dst = ring_s + slot * TMA_STG_BYTES;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TMA_STG_BYTES) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"(tm), "r"(bar),
        "r"(w0 + (t / TMA_NQ) * TMA_COLS), "r"(bp.w), "r"(bp.z),
        "r"(TMA_NQ == 4 ? bp.y + (t & 1) : bp.y),
        "r"(TMA_NQ == 4 ? bp.x + ((t >> 1) & 1) : bp.x + (t & 1))
        : "memory");
  };
  if (TMA && lane == 0)
    for (int t = 0; t < TMA_NSTG - 1 && t < nstep; t++) issue(t);
  const int obj = a.oix[k];
  const int64_t p0 = a.off[obj];
  const int64_t gp0 = a.goff[obj];
  const double *lam = a.lam + gp0, *ql = (a.log_step ? a.loglam : a.lam) + gp0;
  const double2 fq = __ldg(reinterpret_cast<const double2 *>(a.rec) + k);
  const double f = fq.x, qf = fq.y;
  auto pos_q = [&](double q) -> int {
    const int pos = (int)((q - a.q0) * a.qstep_inv);
    return max(0, min(pos, n - 2));
  };
  if (lane < a.nvert) {
    s_off[wid][lane] = (int64_t)a.ids[(int64_t)k * a.nvert + lane] * a.ld;
    s_w[wid][lane] = a.w[(int64_t)k * a.nvert + lane];
  }
  __syncwarp();
  // single-row item (off-grid nearest node): exp rounded to the row's precision
  const bool f32row = a.nvert > 1 && s_off[wid][1] < 0;
  __syncwarp();
  if (f32row && lane >= 1 && lane < a.nvert) { s_off[wid][lane] = s_off[wid][0]; s_w[wid][lane] = 0; }
  __syncwarp();
  // ---- pixel range [plo, phi) of the chunk (prep_kernel)
  const int plo = __ldg(a.pbound + (int64_t)k * (a.nch + 1) + s);
  const int phi = __ldg(a.pbound + (int64_t)k * (a.nch + 1) + s + 1);
  // bring what the resampling will read into L1 while the gather is in flight
  {
    const double *einv = a.einv + p0;
    for (int p = plo + lane * 16; p < phi; p += 32 * 16) {
      prefetch_l1(lam + p);
      prefetch_l1(einv + p);
      if (a.log_step) prefetch_l1(ql + p);
    }
    for (int c = c0 + lane * 16; c <= c1 + 1 && c < n; c += 32 * 16) {
      prefetch_l1(a.lam_t + c);
      if (c < n - 1) prefetch_l1(a.hinv + c);
    }
  }
  int flag = 0;
  // ---- gather + exp over the window: B0[i] = y at knot w0 + i (0 outside the template)
  // The rows stream through a register-free prefetch ring: every lane keeps
  // CK_RING 16-byte cp.async copies (L2 -> shared, bypassing L1 and the register
  // file) in flight, one per (column block, row) in consumption order, into its own
  // 16 bytes of each ring slot -- a lane only ever reads what it copied itself,
  // so no barrier is involved -- and issues the next copy as soon as it has
  // accumulated a slot.  Loads therefore stay in flight during the widening, the
  // FMAs and the exp, and cost no registers.  The ring lives in B1, which the
  // later stages only use after the gather.
  if constexpr (TMA) {
    // One elected lane issues, per column block of TMA_COLS knots, two tensor
    // copies of 8 corner rows each ([TMA_COLS][2][2][2][1] boxes at first-dimension
    // positions p0 and p0 + 1) into the two ring stages; the copy engine computes
    // the 16 row addresses and zero-fills what lies outside the grid (columns
    // before knot 0 or past the row, rows past the top edge).  A stage's mbarrier
    // flips when its 8 x TMA_COLS x 4 bytes have landed; the lanes then read their
    // own columns (conflict-free LDS.64/128), and a stage is re-armed as soon as the
    // warp has consumed it, so one stage is always in flight behind the FMAs.
    static_assert(sizeof(GT) == 4 && NV == 16, "TMA gather: fp32 grid, 16 corners");
    constexpr int TV = TMA_COLS / 32;  // knots per lane and column block (2 or 4)
    const bool round32 = f32row && a.log_spec;
    auto wait = [&](int slot, unsigned parity) {
      asm volatile(
          "{\n\t.reg .pred P1;\n\t"
          "LAB_WAIT:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
          "@P1 bra DONE;\n\t"
          "bra LAB_WAIT;\n\t"
          "DONE:\n\t}" ::"r"(bar_s + 8 * slot), "r"(parity) : "memory");
    };
    __syncwarp();
    const char *ring = reinterpret_cast<const char *>(B1) + lane * (TV * 4);
    int slot = 0;          // stage of step t
    unsigned parity = 0;   // (t / NSTG) & 1
    for (int cb = 0; cb < ncb; cb++) {
      const int i0 = cb * TMA_COLS + lane * TV;
      const int g = w0 + i0;
      double acc[TV];
#pragma unroll
      for (int e = 0; e < TV; e++) acc[e] = 0;
#pragma unroll
      for (int hf = 0; hf < TMA_NQ; hf++) {
        const int t = TMA_NQ * cb + hf;
        // the stage consumed at step t - 1 is free (the __syncwarp below): re-arm it
        if (lane == 0 && t + TMA_NSTG - 1 < nstep) issue(t + TMA_NSTG - 1);
        wait(slot, parity);
        const char *stg = ring + slot * TMA_STG_BYTES;
#pragma unroll
        for (int r = 0; r < TMA_ROWS; r++) {
          // a single-row item uses its row alone (no 0 x neighbour products)
          if (!f32row || (hf == 0 && r == 0)) {
            const double wj = s_w[wid][hf * TMA_ROWS + r];
            if constexpr (TV == 4) {
              const float4 v = *reinterpret_cast<const float4 *>(stg + r * (TMA_COLS * 4));
              acc[0] = fma(wj, (double)v.x, acc[0]);
              acc[1] = fma(wj, (double)v.y, acc[1]);
              acc[2] = fma(wj, (double)v.z, acc[2]);
              acc[3] = fma(wj, (double)v.w, acc[3]);
            } else {
              const float2 v = *reinterpret_cast<const float2 *>(stg + r * (TMA_COLS * 4));
              acc[0] = fma(wj, (double)v.x, acc[0]);
              acc[1] = fma(wj, (double)v.y, acc[1]);
            }
          }
        }
        __syncwarp();  // every lane has read the stage: it may be overwritten
        if (++slot == TMA_NSTG) { slot = 0; parity ^= 1; }
      }
      if (i0 < W0) {
#pragma unroll
        for (int e = 0; e < TV; e++) {
          double y = 0;
          if (g + e >= 0 && g + e < n) {  // else zero padding of the 'same' convolution
            y = a.log_spec ? exp(acc[e]) : acc[e];
            if (round32) y = (double)(float)y;
            if (!(fabs(y) <= 1e100)) flag |= RVS_ST_TEMPLATE_BAD;
          }
          acc[e] = y;
        }
#pragma unroll
        for (int e = 0; e < TV; e += 2)
          *reinterpret_cast<double2 *>(B0 + i0 + e) = make_double2(acc[e], acc[e + 1]);
      }
    }
  } else {
    constexpr int VEC = RowLoader<GT>::VEC;
    using Raw = typename RowLoader<GT>::Raw;
    const bool round32 = f32row && sizeof(GT) == 4 && a.log_spec;
    const GT *base = static_cast<const GT *>(a.grid);
    const int nv = NV > 0 ? NV : a.nvert;
    char *ring = reinterpret_cast<char *>(B1) + lane * 16;
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    auto in_grid = [&](int i0) -> bool {
      const int g = w0 + i0;
      return i0 < W0 && g >= 0 && g < a.ld;
    };
    // CK_RING divides the row count (or equals it), so with the rows unrolled every
    // slot index is a compile-time constant and there is no ring bookkeeping
    constexpr int RG = NV > 0 ? (NV % CK_RING == 0 ? CK_RING : NV) : CK_RING;
    static_assert(RG <= CK_RING, "row count must fit the ring or be a multiple of it");
    auto copy16 = [&](int slot_i, const GT *src) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_s + slot_i * 512), "l"(src) : "memory");
    };
    auto commit = [&]() { asm volatile("cp.async.commit_group;" ::: "memory"); };
    auto finish = [&](double (&acc)[4], int g, bool have, int i0) {
      if (have) {
#pragma unroll
        for (int e = 0; e < VEC; e++) {
          double y = a.log_spec ? exp(acc[e]) : acc[e];
          if (round32) y = (double)(float)y;
          if (g + e >= n) y = 0;  // row padding, never a knot
          else if (!(fabs(y) <= 1e100)) flag |= RVS_ST_TEMPLATE_BAD;
          acc[e] = y;
        }
      }
#pragma unroll
      for (int e = 0; e < VEC; e += 2)
        *reinterpret_cast<double2 *>(B0 + i0 + e) = make_double2(acc[e], acc[e + 1]);
    };
    if (NV > 0) {
      // prologue: rows 0 .. RG-2 of the lane's first column block
      {
        const bool have0 = in_grid(lane * VEC);
        const GT *col = base + (w0 + lane * VEC);
#pragma unroll
        for (int u = 0; u < RG - 1; u++) {
          if (have0) copy16(u, col + s_off[wid][u]);
          commit();
        }
      }
      for (int i0 = lane * VEC; i0 < W0; i0 += 32 * VEC) {
        const int g = w0 + i0;  // multiple of VEC
        double acc[4] = {0, 0, 0, 0};
        const bool have = in_grid(i0), have_next = in_grid(i0 + 32 * VEC);
        const GT *col = base + g, *col_next = col + 32 * VEC;
#pragma unroll
        for (int row = 0; row < NV; row++) {
          // copy number (row + RG - 1) of the stream goes into the slot freed last
          {
            const int ahead = row + RG - 1;
            if (ahead < NV) { if (have) copy16(ahead % RG, col + s_off[wid][ahead]); }
            else if (have_next) copy16(ahead % RG, col_next + s_off[wid][ahead - NV]);
            commit();
          }
          asm volatile("cp.async.wait_group %0;" ::"n"(RG - 1) : "memory");
          if (have) {
            const Raw v = *reinterpret_cast<const Raw *>(ring + (row % RG) * 512);
            double r[4];
            RowLoader<GT>::widen(v, r);
            const double wj = s_w[wid][row];
#pragma unroll
            for (int e = 0; e < VEC; e++) acc[e] = fma(wj, r[e], acc[e]);
          }
        }
        finish(acc, g, have, i0);
      }
    } else {
      // generic row count: running ring counters
      int is_i0 = lane * VEC, is_row = 0, is_slot = 0;
      auto issue = [&]() {
        if (in_grid(is_i0)) copy16(is_slot, base + (w0 + is_i0) + s_off[wid][is_row]);
        commit();
        if (++is_row == nv) { is_row = 0; is_i0 += 32 * VEC; }
        if (++is_slot == CK_RING) is_slot = 0;
      };
#pragma unroll
      for (int u = 0; u < CK_RING - 1; u++) issue();
      int slot = 0;
      for (int i0 = lane * VEC; i0 < W0; i0 += 32 * VEC) {
        const int g = w0 + i0;
        double acc[4] = {0, 0, 0, 0};
        const bool have = in_grid(i0);
        for (int row = 0; row < nv; row++) {
          issue();
          asm volatile("cp.async.wait_group %0;" ::"n"(CK_RING - 1) : "memory");
          if (have) {
            const Raw v = *reinterpret_cast<const Raw *>(ring + slot * 512);
            double r[4];
            RowLoader<GT>::widen(v, r);
            const double wj = s_w[wid][row];
#pragma unroll
            for (int e = 0; e < VEC; e++) acc[e] = fma(wj, r[e], acc[e]);
          }
          if (++slot == CK_RING) slot = 0;
        }
        finish(acc, g, have, i0);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncwarp();
  // ---- rotational broadening onto [ya0, ya1]
  double *Y, *Dz;  // Y[j] = y at knot ya0 + j ; Dz: spline scratch with one guard slot each side
  if (kmax > 0) {
    const double *taps = a.taps + (int64_t)k * a.tapstride;
    const double t0 = __ldg(taps);
    const double *src = B0 + (ya0 - w0);
    const int WY = ya1 - ya0 + 1;
    // short kernels (the usual case: kmax = ceil(R + 1) is 2..4 for vsini up to
    // ~100 km/s on a DESI-like grid) keep their taps in registers
    auto conv_small = [&](auto kt) {
      constexpr int KT = decltype(kt)::value;
      double tp[KT];
#pragma unroll
      for (int t = 0; t < KT; t++) tp[t] = __ldg(taps + 1 + t);
      for (int j = lane; j < WY; j += 32) {
        double sum = t0 * src[j];
#pragma unroll
        for (int t = 1; t <= KT; t++) sum = fma(tp[t - 1], src[j - t] + src[j + t], sum);
        B1[j] = sum;
      }
    };
    switch (kmax) {
      case 1: conv_small(IntTag<1>{}); break;
      case 2: conv_small(IntTag<2>{}); break;
      case 3: conv_small(IntTag<3>{}); break;
      case 4: conv_small(IntTag<4>{}); break;
      case 5: conv_small(IntTag<5>{}); break;
      case 6: conv_small(IntTag<6>{}); break;
      default:
        for (int j = lane; j < WY; j += 32) {
          double sum = t0 * src[j];
          for (int t = 1; t <= kmax; t++) sum = fma(__ldg(taps + t), src[j - t] + src[j + t], sum);
          B1[j] = sum;
        }
    }
    __syncwarp();
    Y = B1;
    Dz = B0 + 1;
  } else {
    Y = B0 + (ya0 - w0);
    Dz = B1 + 1;
  }
  // ---- spline rows kr0 <= kk < kr1 (row kk couples knots kk, kk+1, kk+2; unknown
  // s_{kk+1}).  Dz[r] holds the right-hand side, then d, then s at knot kr0 + r + 1.
  // Guards Dz[-1] (knot kr0) and Dz[nrow] (knot kr1 + 1) are the natural boundary
  // values 0 whenever the window reaches an end of the template.
  const int m = n - 2;
  const int kr0 = ya0, kr1 = min(m, ya1 - 1);
  const int nrow = kr1 - kr0;
  int ch = (nrow + 31) >> 5;
  ch |= 1;  // odd stride between lanes: no shared-memory bank conflicts
  const int r0 = min(nrow, lane * ch), r1 = min(nrow, r0 + ch);
  if (lane == 0) { Dz[-1] = 0; Dz[nrow] = 0; }
  auto solve = [&](auto near_start) {
    // pivot reciprocal of row kk: tabulated next to knot 0, constant elsewhere
    auto winv = [&](int kk) -> double {
      if (decltype(near_start)::value) return kk < CK_NT ? a.wtab[kk] : a.winv_inf;
      return a.winv_inf;
    };
    // forward: d_k = (rho_k - d_{k-1}) / omega_k  ==  a_k + b_k d_{k-1}
    double A = 0, Bc = 1;
    if (r0 < r1) {
      double y1 = Y[r0 + 1];
      double dy0 = y1 - Y[r0];
      for (int r = r0; r < r1; r++) {
        const double y2 = Y[r + 2];
        const double dy1 = y2 - y1;
        const double wi = winv(kr0 + r);
        const double ak = fma(dy1, a.rinv, -dy0) * wi;
        Dz[r] = ak;
        A = fma(-wi, A, ak);
        Bc = -wi * Bc;
        y1 = y2;
        dy0 = dy1;
      }
    }
    double d = warp_affine_up(A, Bc, 0.0, lane);  // d_{kr0-1} := 0 (exact at kr0 == 0)
    // forward apply, and composition of the backward maps of the same rows:
    // s_{k+1} = d_k - gamma_k s_{k+2}, gamma_k = c2 / omega_k; the lane's map takes
    // the s entering above its rows to the s leaving below them
    A = 0;
    Bc = 1;
    for (int r = r0; r < r1; r++) {
      const double wi = winv(kr0 + r);
      d = fma(-wi, d, Dz[r]);
      Dz[r] = d;
      A = fma(Bc, d, A);
      Bc = Bc * (-a.c2 * wi);
    }
    double zz = warp_affine_down(A, Bc, 0.0, lane);  // s_{kr1+1} := 0 (exact at kr1 == m)
    for (int r = r1 - 1; r >= r0; r--) {
      const double ck = -a.c2 * winv(kr0 + r);
      zz = fma(ck, zz, Dz[r]);
      Dz[r] = zz;  // = s at knot kr0 + r + 1
    }
  };
  if (kr0 < CK_NT) solve(TrueTag{}); else solve(FalseTag{});
  __syncwarp();
  // ---- resample onto the chunk's pixels, two pixels per lane and step with all
  // loads of a step issued before any is used
  const double *einv = a.einv + p0;
  double *tn = a.tn + (int64_t)k * a.tn_stride;
  auto value_at = [&](double x, int pos, double xl, double hi) -> double {
    const int j = pos - ya0;
    const double y0v = Y[j], y1v = Y[j + 1];
    const double s0 = Dz[pos - 1 - kr0], s1 = Dz[pos - kr0] * a.r2inv;
    const double u = (x - xl) * hi, v = 1.0 - u;
    return fma(u, fma(s1, fma(u, u, -1.0), y1v), v * fma(s0, fma(v, v, -1.0), y0v));
  };
  for (int pa = plo + lane; pa < phi; pa += 64) {
    const int pb = pa + 32;
    const bool two = pb < phi;
    const int pbs = two ? pb : pa;
    const double la = lam[pa], lb = lam[pbs];
    const double qa = a.log_step ? ql[pa] : 0.0, qb = a.log_step ? ql[pbs] : 0.0;
    const double ea = a.raw_out ? 1.0 : einv[pa], eb = a.raw_out ? 1.0 : einv[pbs];
    const double xa = la * f, xb = lb * f;
    const int posa = pos_q(a.log_step ? qa + qf : xa), posb = pos_q(a.log_step ? qb + qf : xb);
    const double xla = __ldg(a.lam_t + posa), xlb = __ldg(a.lam_t + posb);
    const double hia = __ldg(a.hinv + posa), hib = __ldg(a.hinv + posb);
    tn[pa] = value_at(xa, posa, xla, hia) * ea;
    if (two) tn[pb] = value_at(xb, posb, xlb, hib) * eb;
  }
  if (flag) atomicOr(a.status + k, flag);  // per lane: rare
}

}  // namespace rvs

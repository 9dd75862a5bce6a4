// Device functions of the chi-square evaluation shared by chisq.cu and fused.cu:
// spline evaluation from (y,z) pairs, Gram accumulators, warp reductions and
// the small Cholesky solve.  See chisq.cu for the reference lines they follow.
#pragma once
#include <math.h>

#include "common.cuh"

namespace rvs {

// ------------------------------------------------------------------ chi-square
struct ScanArgs {
  const double2 *yz;
  int64_t yz_stride;
  const int32_t *tix;
  const double *lam_t, *h, *hinv;
  int npix_t, log_step;
  double x0, xlast, q0, qstep_inv;  // knot-grid origin (ln or linear) and 1/step
  const double *lam, *loglam, *dn, *einv, *sumlog2;  // lam, loglam, P: grid pools
  const int64_t *off, *goff;
  const int32_t *oix;
  const double *P;  // pixel-major [pixel][npp]
  int npp;
  const double *vels;
  int nv, K;
  double *chisq;
  int32_t *status;
  double *coeffs, *raw, *model;
  const int64_t *moff;
  int fast_interp;  // 1: value of the first knot >= x instead of the spline (spec_fit.py:913-918)
  // resolution matrices (rvs_obs): band rows by output pixel, diagonal offsets
  const double *resol;
  const int32_t *resol_offs;
  int nresol;
  int resol_hw;  // max |resol_offs[]|
  int nby;       // trial blocks per item (1-D grid of the GEMM scan kernel)
  const int32_t *nvk;  // optional [K]: trials of item k actually wanted (<= nv), else all
};

constexpr int SCAN_WARPS = 8;
// widest half-bandwidth of a resolution matrix the GEMM scan kernel stages in shared
// memory (scan_mma.cuh); wider ones take the per-trial scan kernel
constexpr int RS_HW = 10;

// fast_interp: templ_spec[searchsorted(templ_lam, x)] (numpy 'left': first knot
// >= x), starting from the uniform-grid interval `pos` of x
__device__ __forceinline__ double nearest_knot_value(const ScanArgs &a, const double2 *yz, double x,
                                                     int pos) {
  int i = pos;
  while (i > 0 && __ldg(a.lam_t + i - 1) >= x) i--;
  while (i < a.npix_t - 1 && __ldg(a.lam_t + i) < x) i++;
  return __ldg(yz + i).x;
}

// spline value at rest wavelength x (spliner.c:97-106) from (y,z) pairs
__device__ __forceinline__ double spline_eval(const ScanArgs &a, const double2 *yz, double x,
                                              double q) {
  int pos = (int)((q - a.q0) * a.qstep_inv);
  pos = max(0, min(pos, a.npix_t - 2));
  if (a.fast_interp) return nearest_knot_value(a, yz, x, pos);
  const double2 c0 = __ldg(yz + pos), c1 = __ldg(yz + pos + 1);
  const double xl = __ldg(a.lam_t + pos), xr = __ldg(a.lam_t + pos + 1);
  const double hh = __ldg(a.h + pos), hi = __ldg(a.hinv + pos);
  const double t1 = hi * (1. / 6), t2 = hh * (1. / 6);
  const double A = c1.y * t1, B = c0.y * t1;
  const double C = c1.x * hi - c1.y * t2, D = c0.x * hi - c0.y * t2;
  const double dl = x - xl, dr = xr - x;
  return A * dl * dl * dl + B * dr * dr * dr + C * dl + D * dr;
}

// Template value at observed pixel p of an object with npix pixels: the spline at the
// Doppler-shifted wavelength, or with a resolution matrix (rb = the object's band rows)
// row p of R times the resampled template (spec_fit.py:922-929), diagonals in ascending
// order.  EVAL(x, q) is the spline evaluation of the calling kernel.
template <class EVAL>
__device__ __forceinline__ double template_at(const ScanArgs &a, const double *lam,
                                              const double *ql, const double *rb, int npix, int p,
                                              double f, double qf, EVAL eval) {
  if (rb == nullptr) {
    const double x = lam[p] * f;
    return eval(x, a.log_step ? ql[p] + qf : x);
  }
  double s = 0;
  for (int k = 0; k < a.nresol; k++) {
    const int pp = p + __ldg(a.resol_offs + k);
    if (pp < 0 || pp >= npix) continue;
    const double x = lam[pp] * f;
    s = fma(__ldg(rb + (int64_t)k * npix + p), eval(x, a.log_step ? ql[pp] + qf : x), s);
  }
  return s;
}

// Reduce N per-lane values over the 32 lanes of a warp with N/2+N/4+... shuffles
// instead of 5N.  After run(), slot r < ceil(N/32) of lane L holds the warp total
// of original entry index(r, L) (or index < 0: padding).  Fixed summation tree,
// hence deterministic.
template <int N, int O>
struct WarpScatterReduce {
  static constexpr int H = (N + 1) / 2;
  __device__ __forceinline__ static void run(double *v, int lane) {
    const bool up = (lane & O) != 0;
#pragma unroll
    for (int i = 0; i < H; i++) {
      const double a = v[i];
      const double b = (i + H < N) ? v[i + H] : 0.0;
      const double send = up ? a : b;
      const double keep = up ? b : a;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
    }
    if constexpr (O > 1) WarpScatterReduce<H, O / 2>::run(v, lane);
  }
  __device__ __forceinline__ static int index(int r, int lane) {
    int sub;
    if constexpr (O > 1) sub = WarpScatterReduce<H, O / 2>::index(r, lane);
    else sub = (r < H) ? r : -1;
    if (sub < 0) return -1;
    const int idx = sub + ((lane & O) ? H : 0);
    return idx < N ? idx : -1;
  }
};

template <int N>
__device__ __forceinline__ void warp_reduce_store(double (&v)[N], double *dst, int lane) {
  WarpScatterReduce<N, 16>::run(v, lane);
#pragma unroll
  for (int r = 0; r < (N + 31) / 32; r++) {
    const int idx = WarpScatterReduce<N, 16>::index(r, lane);
    if (idx >= 0) dst[idx] = v[r];
  }
}

// Lower-triangle accumulators of rows [R0,R1) of M = G G^T (+ v = G d)
template <int NP, int R0, int R1>
struct GramAcc {
  static constexpr int NM = (R1 * (R1 + 1) - R0 * (R0 + 1)) / 2;
  double m[NM > 0 ? NM : 1];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < NM; i++) m[i] = 0;
  }
  __device__ __forceinline__ void add(const double (&g)[NP]) {
    int c = 0;
#pragma unroll
    for (int i = R0; i < R1; i++) {
#pragma unroll
      for (int j = 0; j <= i; j++) {
        m[c] = fma(g[i], g[j], m[c]);
        c++;
      }
    }
  }
  // transposed butterfly reduction over the warp; the lane that ends up owning
  // entry c stores it to the packed lower triangle (rows R0.. are contiguous)
  __device__ __forceinline__ void reduce_store(double *Ms, int lane) {
    WarpScatterReduce<NM, 16>::run(m, lane);
#pragma unroll
    for (int r = 0; r < (NM + 31) / 32; r++) {
      const int idx = WarpScatterReduce<NM, 16>::index(r, lane);
      if (idx >= 0) Ms[R0 * (R0 + 1) / 2 + idx] = m[r];
    }
  }
};

// g[i] = P_i(pixel) * tn from the pixel-major basis (rows of npp doubles, npp even,
// 16-byte aligned)
template <int NP>
__device__ __forceinline__ void load_basis(const double *Prow, double tn, double (&g)[NP]) {
  const double2 *P2 = reinterpret_cast<const double2 *>(Prow);
#pragma unroll
  for (int i = 0; i < NP / 2; i++) {
    const double2 v = __ldg(P2 + i);
    g[2 * i] = v.x * tn;
    g[2 * i + 1] = v.y * tn;
  }
  if (NP & 1) g[NP - 1] = __ldg(Prow + NP - 1) * tn;
}

// Warp-cooperative Cholesky solve of M a = v.  Ms: packed lower triangle of M in
// shared memory (overwritten with L), a: v in, coefficients out.  Lane i owns
// row i in registers; the right-looking update exchanges column entries by
// shuffle, the forward substitution rides along, and the logarithms of the
// pivots are taken once, in parallel, at the end (2 sum ln L_ii = sum ln s_j).
// Returns that sum, or NaN if M is not positive definite.  All 32 lanes must call.
template <int NP>
__device__ __forceinline__ double chol_solve(double *Ms, double *a, int lane) {
  static_assert(NP <= 32, "one lane per row");
  double row[NP];
#pragma unroll
  for (int l = 0; l < NP; l++) row[l] = (lane < NP && l <= lane) ? Ms[lane * (lane + 1) / 2 + l] : 0.0;
  double rhs = lane < NP ? a[lane] : 0.0;
  double spiv = 1.0, myinv = 0.0;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < NP; j++) {
    const double s = __shfl_sync(0xffffffffu, row[j], j);
    if (!(s > 0) || isinf(s)) ok = false;
    const double inv = rsqrt(s);
    const double lij = row[j] * inv;  // L_ij for lanes i >= j (L_jj = s / sqrt(s))
    row[j] = lij;
    const double yj = __shfl_sync(0xffffffffu, rhs, j) * inv;
    if (lane > j) rhs = fma(-lij, yj, rhs);
    if (lane == j) { rhs = yj; spiv = s; myinv = inv; }
#pragma unroll
    for (int l = j + 1; l < NP; l++) {
      const double llj = __shfl_sync(0xffffffffu, lij, l);
      if (lane >= l) row[l] = fma(-lij, llj, row[l]);
    }
  }
  const double ldet = warp_sum(lane < NP ? log(spiv) : 0.0);
  // back substitution L^T a = y: every lane redundantly, operands broadcast from smem
  if (lane < NP) {
#pragma unroll
    for (int l = 0; l < NP; l++)
      if (l < lane) Ms[lane * (lane + 1) / 2 + l] = row[l];
    Ms[lane * (lane + 1) / 2 + lane] = myinv;  // reciprocal diagonal
    a[lane] = rhs;                             // y
  }
  __syncwarp();
  double y[NP];
#pragma unroll
  for (int i = NP - 1; i >= 0; i--) {
    double t = a[i];
#pragma unroll
    for (int k = i + 1; k < NP; k++) t = fma(-Ms[k * (k + 1) / 2 + i], y[k], t);
    y[i] = t * Ms[i * (i + 1) / 2 + i];
  }
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NP; i++) a[i] = y[i];
  }
  __syncwarp();
  return ok ? ldet : nan("");
}


int fill_scan_args(ScanArgs &a, const rvs_knots *kn, const rvs_obs *obs);

}  // namespace rvs

// Device functions of the template evaluation, shared by template_build.cu and
// fused.cu: corner-weighted gather + exp, rotational broadening, chunked Thomas
// solve.  See template_build.cu for the reference lines they follow.
#pragma once
#include <math.h>

#include "common.cuh"

namespace rvs {

constexpr int TB_THREADS = 256;
constexpr int SPL_HALO = 32;  // 0.268^32 ~ 5e-19: below fp64 rounding

// primitives of K(x) and x K(x), K ~ c1 sqrt(1-x^2) + c2 (1-x^2), eps = 0.6
__device__ __forceinline__ void rot_primitives(double x, double &k0, double &k1) {
  const double eps = 0.6;
  const double pi = 3.141592653589793;
  x = fmin(fmax(x, -1.0), 1.0);
  const double nrm = pi * (1 - eps / 3.0);
  const double c1 = 2 * (1 - eps) / nrm;
  const double c2 = (pi / 2.0) * eps / nrm;
  const double x2 = x * x;
  const double u = 1 - x2;
  const double root = sqrt(u);
  k0 = c1 * (0.5 * (x * root + asin(x))) + c2 * (x - (x2 * x) / 3.0);
  k1 = c1 * (-1.0 / 3.0 * u * root) + c2 * (x2 / 2.0 - (x2 * x2) / 4.0);
}

__device__ __forceinline__ double rot_segment(double xa, double xb, double slope,
                                              double icpt) {
  double k0a, k1a, k0b, k1b;
  rot_primitives(xb, k0b, k1b);
  rot_primitives(xa, k0a, k1a);
  return slope * (k1b - k1a) + icpt * (k0b - k0a);
}

__device__ __forceinline__ double clip1(double x) { return fmin(fmax(x, -1.0), 1.0); }

// one-sided weight w_k, k >= 0, of the overlap of the rotation profile with the
// triangular pixel basis
__device__ __forceinline__ double rot_weight(int k, double R) {
  double w = 0;
  double lo = clip1(k / R), hi = clip1((k + 1) / R);
  if (hi > lo) w += rot_segment(lo, hi, -R, 1.0 + k);
  lo = clip1((k - 1) / R);
  hi = clip1(k / R);
  if (hi > lo) w += rot_segment(lo, hi, R, 1.0 - k);
  return w;
}

__device__ __forceinline__ double block_sum(double v, double *scratch) {
  // deterministic: warp tree, then warp 0 sums the per-warp partials in order
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  double t = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += scratch[i];
  return t;
}

template <typename GT>
struct RowLoader;

template <>
struct RowLoader<float> {
  static constexpr int VEC = 4;
  using Raw = float4;
  __device__ static void widen(const Raw &v, double out[4]) {
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  }
  __device__ static void load(const float *row, int q, double out[4]) {
    const float4 v = ldg_stream_f4(reinterpret_cast<const float4 *>(row) + q);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  }
};

template <>
struct RowLoader<double> {
  static constexpr int VEC = 2;
  using Raw = double2;
  __device__ static void widen(const Raw &v, double out[4]) {
    out[0] = v.x; out[1] = v.y;
  }
  __device__ static void load(const double *row, int q, double out[4]) {
    const double2 v = __ldg(reinterpret_cast<const double2 *>(row) + q);
    out[0] = v.x; out[1] = v.y;
  }
};

struct TemplateArgs {
  const void *grid;
  int64_t ld;
  int npix_t;
  const int32_t *ids;
  const double *w;
  int nvert;
  const double *vsini;
  const double *h, *hinv, *cp, *winv;
  double lnstep;
  int log_spec;
  double *yz;
  int64_t yz_stride;
  int32_t *status;
  int npad;  // smem row length
};

// An item whose second vertex id is negative is a single-row item: the
// reference returns exp(dats[nearest]) for off-grid points (spec_inter.py:
// 160,167), and numpy evaluates that exp in the row's own precision (float32
// for the stored grids).  Such items are evaluated the same way: fp64 exp of
// the fp32 value, rounded to fp32.  Returns true for them and rewrites the
// vertex list to {row, w=1; row, w=0 ...} (block-uniform, syncs).
__device__ __forceinline__ bool single_row_item(const TemplateArgs &a, int32_t *s_ids,
                                                double *s_w) {
  const bool single = a.nvert > 1 && s_ids[1] < 0;
  __syncthreads();
  if (single && threadIdx.x >= 1 && threadIdx.x < a.nvert) {
    s_ids[threadIdx.x] = s_ids[0];
    s_w[threadIdx.x] = 0;
  }
  __syncthreads();
  return single;
}

template <typename GT, int NV>
__device__ __forceinline__ void gather_rows(const TemplateArgs &a, const int32_t *s_ids,
                                            const double *s_w, double *ya, int &bad,
                                            bool f32row) {
  const bool round32 = f32row && sizeof(GT) == 4 && a.log_spec;
  constexpr int VEC = RowLoader<GT>::VEC;
  const int nvec = a.npix_t / VEC;
  const GT *base = static_cast<const GT *>(a.grid);
  const int nv = NV > 0 ? NV : a.nvert;
  for (int q = threadIdx.x; q < nvec; q += TB_THREADS) {
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
      double y = a.log_spec ? exp(acc[e]) : acc[e];
      if (round32) y = (double)(float)y;
      if (!(fabs(y) <= 1e100)) bad = 1;  // also catches NaN
      ya[q * VEC + e] = y;
    }
  }
  for (int p = nvec * VEC + threadIdx.x; p < a.npix_t; p += TB_THREADS) {  // tail
    double acc = 0;
    for (int j = 0; j < nv; j++)
      acc = fma(s_w[j], (double)base[(int64_t)s_ids[j] * a.ld + p], acc);
    double y = a.log_spec ? exp(acc) : acc;
    if (round32) y = (double)(float)y;
    if (!(fabs(y) <= 1e100)) bad = 1;
    ya[p] = y;
  }
}

// Shared by the stand-alone builder and the fused chi-square kernel: after the
// call y (broadened) is in *py and z in *pz, both of length npix_t in shared
// memory; scratch buffers ya, yb, yc (npad doubles each) and taps
// (RVS_MAX_TAPS+1) are consumed.
__device__ __forceinline__ void broaden_and_spline(const TemplateArgs &a, double vs, double *&ya, double *&yb,
                                   double *yc, double *taps, double *red, double *&py,
                                   double *&pz, int *taps_over) {
  const int n = a.npix_t;
  const int tid = threadIdx.x;
  // ---- rotational broadening (spec_fit.py:650-682)
  if (vs > 0) {
    const double R = (vs / RVS_C_KMS) / a.lnstep;
    if (R >= 1e-9) {
      int kmax = (int)ceil(R + 1);
      if (kmax > RVS_MAX_TAPS) { kmax = RVS_MAX_TAPS; *taps_over = 1; }
      double part = 0;
      for (int k = tid; k <= kmax; k += TB_THREADS) {
        const double w = rot_weight(k, R);
        taps[k] = w;
        part += (k == 0) ? w : 2 * w;
      }
      const double tot = block_sum(part, red);
      __syncthreads();
      for (int k = tid; k <= kmax; k += TB_THREADS) taps[k] = taps[k] / tot;
      __syncthreads();
      for (int p = tid; p < n; p += TB_THREADS) {
        double s = taps[0] * ya[p];
        for (int k = 1; k <= kmax; k++) {
          const double lo = (p - k >= 0) ? ya[p - k] : 0.0;
          const double hi = (p + k < n) ? ya[p + k] : 0.0;
          s = fma(taps[k], lo + hi, s);
        }
        yb[p] = s;
      }
      __syncthreads();
      double *t = ya; ya = yb; yb = t;
    }
  }
  // ---- spline second derivatives (spliner.c:21-50), chunked Thomas with halo
  const int m = n - 2;
  int ch = (m + TB_THREADS - 1) / TB_THREADS;
  if ((ch & 1) == 0) ch++;  // odd stride: no shared-memory bank conflicts
  const int k0 = tid * ch, k1 = min(m, k0 + ch);
  double *dbuf = yb, *zbuf = yc;
  if (k0 < m) {
    const int ks = max(0, k0 - SPL_HALO);
    double d = 0;
    double bl = (ya[ks + 1] - ya[ks]) * __ldg(a.hinv + ks);
    for (int k = ks; k < k1; k++) {
      const double br = (ya[k + 2] - ya[k + 1]) * __ldg(a.hinv + k + 1);
      const double u = 6 * (br - bl);
      d = (u - __ldg(a.h + k) * d) * __ldg(a.winv + k);
      if (k >= k0) dbuf[k] = d;
      bl = br;
    }
  }
  __syncthreads();
  if (k0 < m) {
    const int ke = min(m, k1 + SPL_HALO);
    double zz = 0;
    for (int k = ke - 1; k >= k0; k--) {
      zz = dbuf[k] - __ldg(a.cp + k) * zz;
      if (k < k1) zbuf[k + 1] = zz;
    }
  }
  if (tid == 0) { zbuf[0] = 0; zbuf[n - 1] = 0; }
  __syncthreads();
  py = ya;
  pz = zbuf;
}


int fill_template_args(TemplateArgs &a, const void *d_grid, int64_t ld, const rvs_knots *kn,
                       const int32_t *d_ids, const double *d_w, int nvert,
                       const double *d_vsini, int log_spec);

}  // namespace rvs

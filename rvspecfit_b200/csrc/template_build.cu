// Template evaluation: corner-weighted gather over the template grid, exp,
// rotational (vsini) broadening and natural-cubic-spline second derivatives.
// One CTA per item; the whole template lives in shared memory.
//
// Reference behaviour reproduced (paths under /root/reference/py/rvspecfit/):
//   spec_inter.py:134-194 / 35-59  weighted sum of grid rows, exp
//   spec_fit.py:495-625            analytic rotation kernel weights
//   spec_fit.py:628-682            zero-padded 'same' convolution
//   src/spliner.c:7-60             Thomas solve for the spline
#include <math.h>

#include "template_device.cuh"

namespace rvs {

template <typename GT, int NV>
__global__ void __launch_bounds__(TB_THREADS) template_build_kernel(TemplateArgs a) {
  extern __shared__ double sm[];
  double *ya = sm, *yb = sm + a.npad, *yc = sm + 2 * a.npad, *taps = sm + 3 * a.npad;
  __shared__ int32_t s_ids[32];
  __shared__ double s_w[32];
  __shared__ double red[TB_THREADS / 32];
  __shared__ int s_bad, s_taps;
  const int k = blockIdx.x;
  if (threadIdx.x < a.nvert) {
    s_ids[threadIdx.x] = a.ids[(int64_t)k * a.nvert + threadIdx.x];
    s_w[threadIdx.x] = a.w[(int64_t)k * a.nvert + threadIdx.x];
  }
  if (threadIdx.x == 0) { s_bad = 0; s_taps = 0; }
  __syncthreads();
  const bool f32row = single_row_item(a, s_ids, s_w);
  int bad = 0;
  gather_rows<GT, NV>(a, s_ids, s_w, ya, bad, f32row);
  if (bad) s_bad = 1;
  __syncthreads();
  double *py, *pz;
  broaden_and_spline(a, a.vsini ? a.vsini[k] : 0.0, ya, yb, yc, taps, red, py, pz, &s_taps);
  double2 *out = reinterpret_cast<double2 *>(a.yz) + (int64_t)k * a.yz_stride;
  for (int p = threadIdx.x; p < a.npix_t; p += TB_THREADS) out[p] = make_double2(py[p], pz[p]);
  if (threadIdx.x == 0) a.status[k] = (s_bad ? RVS_ST_TEMPLATE_BAD : RVS_ST_OK) | (s_taps ? RVS_ST_TAPS : 0);
}

template <typename GT>
static int launch_template(const TemplateArgs &a, int K, size_t smem, cudaStream_t st) {
  auto go = [&](auto kern) -> int {
    RVS_CUDA_OK(ensure_dyn_smem(kern, smem));
    kern<<<K, TB_THREADS, smem, st>>>(a);
    RVS_LAUNCH_OK();
    return 0;
  };
  switch (a.nvert) {
    case 16: return go(template_build_kernel<GT, 16>);
    case 8: return go(template_build_kernel<GT, 8>);
    case 5: return go(template_build_kernel<GT, 5>);
    case 1: return go(template_build_kernel<GT, 1>);
    default: return go(template_build_kernel<GT, 0>);
  }
}

}  // namespace rvs

extern "C" void rvs_knot_tables(const double *x, int n, double *h, double *hinv, double *cp,
                                double *winv) {
  for (int i = 0; i + 1 < n; i++) {
    h[i] = x[i + 1] - x[i];
    hinv[i] = 1. / h[i];
  }
  const int m = n - 2;
  for (int k = 0; k < m; k++) {
    const double diag = 2 * (h[k + 1] + h[k]);
    const double den = (k == 0) ? diag : diag - h[k] * cp[k - 1];
    cp[k] = h[k + 1] / den;
    winv[k] = 1. / den;
  }
}

extern "C" int rvs_knot_info(const double *x, int n, int log_step, rvs_knots *out) {
  if (n < 3) return -1;
  out->npix_t = n;
  out->log_step = log_step;
  out->x0 = x[0];
  out->xlast = x[n - 1];
  out->lnstep = log(x[1] / x[0]);
  if (log_step) {
    const double s1 = log(x[1] / x[0]), s2 = log(x[2] / x[1]);
    if (fabs(s1 - s2) > 1e-10) return -2;
    out->q0 = log(x[0]);
    out->qstep_inv = 1.0 / s1;
  } else {
    const double s1 = x[1] - x[0], s2 = x[2] - x[1];
    if (fabs(s1 - s2) > 1e-10) return -2;
    out->q0 = x[0];
    out->qstep_inv = 1.0 / s1;
  }
  // constant ratio of consecutive knot spacings (rvs_chisq_fused's spline solve)
  out->ratio = log_step ? exp(log(x[n - 1] / x[0]) / (n - 1)) : 1.0;
  double dev = 0;
  for (int k = 0; k + 2 < n; k++) {
    const double h0 = x[k + 1] - x[k], h1 = x[k + 2] - x[k + 1];
    dev = fmax(dev, fabs(h1 / (out->ratio * h0) - 1));
  }
  out->ratio_dev = dev;
  return 0;
}

namespace rvs {
int fill_template_args(TemplateArgs &a, const void *d_grid, int64_t ld, const rvs_knots *kn,
                       const int32_t *d_ids, const double *d_w, int nvert,
                       const double *d_vsini, int log_spec) {
  RVS_REQUIRE(d_grid && kn && d_ids && d_w && kn->d_h && kn->d_hinv && kn->d_cp && kn->d_winv,
              RVS_E_ARG, "template: null pointer");
  RVS_REQUIRE(kn->npix_t >= 4 && nvert >= 1 && nvert <= 32, RVS_E_ARG,
              "template: bad sizes npix_t=%d nvert=%d", kn->npix_t, nvert);
  RVS_REQUIRE(ld % 4 == 0 && ((uintptr_t)d_grid & 15) == 0, RVS_E_ARG,
              "template: grid rows must be 16-byte aligned (ld %% 4 == 0)");
  a.grid = d_grid; a.ld = ld; a.npix_t = kn->npix_t; a.ids = d_ids; a.w = d_w; a.nvert = nvert;
  a.vsini = d_vsini; a.h = kn->d_h; a.hinv = kn->d_hinv; a.cp = kn->d_cp; a.winv = kn->d_winv;
  a.lnstep = kn->lnstep; a.log_spec = log_spec; a.yz = nullptr; a.yz_stride = 0;
  a.status = nullptr;
  a.npad = (kn->npix_t + 3) & ~3;
  return 0;
}
}  // namespace rvs

extern "C" int rvs_template_build(const void *d_grid, int grid_f64, int64_t ld,
                                  const rvs_knots *knots, const int32_t *d_ids,
                                  const double *d_w, int nvert, const double *d_vsini,
                                  int log_spec, int K, double *d_yz, int64_t yz_stride,
                                  int32_t *d_status, void *stream) {
  using namespace rvs;
  if (K == 0) return 0;
  TemplateArgs a;
  const int rc = fill_template_args(a, d_grid, ld, knots, d_ids, d_w, nvert, d_vsini, log_spec);
  if (rc) return rc;
  RVS_REQUIRE(d_yz && d_status && K > 0, RVS_E_ARG, "rvs_template_build: null output");
  RVS_REQUIRE(yz_stride >= a.npix_t, RVS_E_ARG, "rvs_template_build: yz_stride < npix_t");
  a.yz = d_yz; a.yz_stride = yz_stride; a.status = d_status;
  const size_t smem = sizeof(double) * (3 * (size_t)a.npad + RVS_MAX_TAPS + 1);
  RVS_REQUIRE(smem <= 227 * 1024, RVS_E_LIMIT,
              "rvs_template_build: npix_t=%d needs %zu B shared memory (limit 227 KB)",
              a.npix_t, smem);
  cudaStream_t st = (cudaStream_t)stream;
  return grid_f64 ? launch_template<double>(a, K, smem, st)
                  : launch_template<float>(a, K, smem, st);
}

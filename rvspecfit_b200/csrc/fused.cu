// C-ABI entry point of the fused optimiser-phase evaluation: stage A
// (slice_kernel.cuh: template window -> T/sigma) + stage B (gram_kernel.cuh:
// continuum solve -> chi-square).  The spline never goes to HBM; the only
// intermediate is T/sigma (8 bytes per observed pixel).
#include "gram_kernel.cuh"
#include "slice_kernel.cuh"

namespace rvs {
int launch_gram_group0(const GramArgs &, int, int, cudaStream_t);
int launch_gram_group1(const GramArgs &, int, int, cudaStream_t);
int launch_gram_group2(const GramArgs &, int, int, cudaStream_t);
int launch_gram_group3(const GramArgs &, int, int, cudaStream_t);

template <typename GT, int NV>
static int launch_slice_one(const SliceArgs &a, int S, int K, size_t smem, cudaStream_t st) {
  auto kern = slice_kernel<GT, NV>;
  RVS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(S, K), SL_THREADS, smem, st>>>(a);
  RVS_LAUNCH_OK();
  return 0;
}
}  // namespace rvs

static double *g_dbg = nullptr;
extern "C" void rvs_set_debug_buffer(double *d_buf) { g_dbg = d_buf; }

extern "C" int rvs_fused_slices(int npix_t) {
  int S = (npix_t + 640) / 1280;
  return S < 1 ? 1 : (S > 16 ? 16 : S);
}

extern "C" int rvs_chisq_fused(const void *d_grid, int grid_f64, int64_t ld,
                               const rvs_knots *knots, const int32_t *d_ids, const double *d_w,
                               int nvert, const double *d_vsini, double vsini_max, int log_spec,
                               const rvs_obs *obs, const int32_t *d_oix, const double *d_vels,
                               int K, double *d_tn, int64_t tn_stride, double *d_chisq,
                               int32_t *d_status, void *stream) {
  using namespace rvs;
  if (K == 0) return 0;
  TemplateArgs ta;
  ScanArgs sa;
  int rc = fill_template_args(ta, d_grid, ld, knots, d_ids, d_w, nvert, d_vsini, log_spec);
  if (rc) return rc;
  rc = fill_scan_args(sa, knots, obs);
  if (rc) return rc;
  RVS_REQUIRE(d_oix && d_vels && d_chisq && d_status && d_tn, RVS_E_ARG,
              "rvs_chisq_fused: null pointer");
  RVS_REQUIRE(K <= 65535, RVS_E_LIMIT, "rvs_chisq_fused: K=%d > 65535 items per call", K);
  const int n = knots->npix_t;
  const int S = rvs_fused_slices(n);
  int tapcap = 0;
  if (vsini_max > 0) {
    const double R = (vsini_max / RVS_C_KMS) / knots->lnstep;
    tapcap = (int)ceil(R + 1) + 1;
  }
  RVS_REQUIRE(tapcap <= RVS_MAX_TAPS, RVS_E_LIMIT, "rvs_chisq_fused: vsini_max=%g needs %d taps",
              vsini_max, tapcap);
  SliceArgs a;
  a.grid = d_grid; a.ld = ld; a.npix_t = n; a.ids = d_ids; a.w = d_w; a.nvert = nvert;
  a.vsini = d_vsini; a.lam_t = knots->d_lam_t; a.h = knots->d_h; a.hinv = knots->d_hinv;
  a.cp = knots->d_cp; a.winv = knots->d_winv; a.lnstep = knots->lnstep; a.log_spec = log_spec;
  a.log_step = knots->log_step; a.x0 = knots->x0; a.xlast = knots->xlast; a.q0 = knots->q0;
  a.qstep_inv = knots->qstep_inv;
  a.lam = obs->d_lam; a.loglam = obs->d_loglam; a.einv = obs->d_einv; a.off = obs->d_off;
  a.oix = d_oix; a.vels = d_vels; a.tn = d_tn; a.tn_stride = tn_stride; a.status = d_status;
  a.tapcap = tapcap;
  a.dbg = g_dbg;
  a.wcap = ((n + S - 1) / S + 2 * (SPL_HALO + 3 + tapcap) + 20 + 3) & ~3;
  const size_t smem = sizeof(double) * (2 * (size_t)a.wcap + tapcap + 1);
  RVS_REQUIRE(smem <= 200 * 1024, RVS_E_LIMIT, "rvs_chisq_fused: window needs %zu B smem", smem);
  cudaStream_t st = (cudaStream_t)stream;
  RVS_CUDA_OK(cudaMemsetAsync(d_status, 0, sizeof(int32_t) * K, st));
  if (grid_f64) rc = launch_slice_one<double, 0>(a, S, K, smem, st);
  else if (nvert == 16) rc = launch_slice_one<float, 16>(a, S, K, smem, st);
  else if (nvert == 5) rc = launch_slice_one<float, 5>(a, S, K, smem, st);
  else rc = launch_slice_one<float, 0>(a, S, K, smem, st);
  if (rc) return rc;
  GramArgs g;
  g.tn = d_tn; g.tn_stride = tn_stride; g.dn = obs->d_dn; g.sumlog2 = obs->d_sumlog2;
  g.off = obs->d_off; g.oix = d_oix; g.P = obs->d_P; g.pstride = obs->pstride;
  g.boff = obs->d_boff; g.chisq = d_chisq; g.status = d_status;
  const int np = obs->npoly;
  if (np <= 7) return launch_gram_group0(g, np, K, st);
  if (np <= 10) return launch_gram_group1(g, np, K, st);
  if (np <= 13) return launch_gram_group2(g, np, K, st);
  return launch_gram_group3(g, np, K, st);
}

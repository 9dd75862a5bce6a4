// C-ABI entry point of the fused optimiser-phase evaluation (fused_kernel.cuh).
#include "fused_kernel.cuh"

namespace rvs {
int launch_fused_group0(const FusedArgs &, int, int, int, size_t, cudaStream_t);
int launch_fused_group1(const FusedArgs &, int, int, int, size_t, cudaStream_t);
int launch_fused_group2(const FusedArgs &, int, int, int, size_t, cudaStream_t);
int launch_fused_group3(const FusedArgs &, int, int, int, size_t, cudaStream_t);
}  // namespace rvs

extern "C" int rvs_chisq_fused(const void *d_grid, int grid_f64, int64_t ld,
                               const rvs_knots *knots, const int32_t *d_ids, const double *d_w,
                               int nvert, const double *d_vsini, int log_spec,
                               const rvs_obs *obs, const int32_t *d_oix, const double *d_vels,
                               int K, double *d_chisq, int32_t *d_status, void *stream) {
  using namespace rvs;
  if (K == 0) return 0;
  FusedArgs fa;
  int rc = fill_template_args(fa.t, d_grid, ld, knots, d_ids, d_w, nvert, d_vsini, log_spec);
  if (rc) return rc;
  rc = fill_scan_args(fa.s, knots, obs);
  if (rc) return rc;
  RVS_REQUIRE(d_oix && d_vels && d_chisq && d_status, RVS_E_ARG, "rvs_chisq_fused: null pointer");
  fa.s.yz = nullptr; fa.s.yz_stride = 0; fa.s.tix = nullptr; fa.s.oix = d_oix;
  fa.s.vels = d_vels; fa.s.nv = 1; fa.s.K = K; fa.s.chisq = d_chisq; fa.s.status = d_status;
  fa.s.coeffs = nullptr; fa.s.raw = nullptr; fa.s.model = nullptr; fa.s.moff = nullptr;
  const size_t smem = sizeof(double) * (3 * (size_t)fa.t.npad + RVS_MAX_TAPS + 1);
  RVS_REQUIRE(smem <= 200 * 1024, RVS_E_LIMIT,
              "rvs_chisq_fused: npix_t=%d needs %zu B shared memory", fa.t.npix_t, smem);
  cudaStream_t st = (cudaStream_t)stream;
  const int np = obs->npoly;
  if (np <= 7) return launch_fused_group0(fa, np, grid_f64, K, smem, st);
  if (np <= 10) return launch_fused_group1(fa, np, grid_f64, K, smem, st);
  if (np <= 13) return launch_fused_group2(fa, np, grid_f64, K, smem, st);
  return launch_fused_group3(fa, np, grid_f64, K, smem, st);
}

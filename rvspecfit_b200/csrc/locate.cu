// Vertex location on the regular template grid (reference spec_inter.py:153-194):
// per dimension pos = digitize(q, u) - 1, the 2^d corner ids from idgrid and
// the weights prod_i x_i^S (1 - x_i)^(1-S), corners in itertools.product([0,1])
// order, factors multiplied in dimension order (np.prod).  Pure comparisons,
// one subtraction/division per dimension and d-1 multiplications per corner,
// so ids and weights are bit-identical to the host computation.  Points the
// reference resolves through its KD-tree (outside the grid -- the top edge counts
// as outside --, or a missing corner: nearest node in ptp-normalised space,
// spec_inter.py:128-132,156-167, and its distance as the off-grid measure,
// spec_inter.py:77-92) are flagged; nearest_node_kernel then finds that node by
// exhaustive search, one warp per flagged point.  Non-finite coordinates stay
// flagged for the host.
#include "common.cuh"

namespace rvs {

__global__ void __launch_bounds__(128) locate_grid_kernel(rvs_gridmap gm, const double *q,
                                                          int64_t qstride, int K, int32_t *ids,
                                                          double *w, int32_t *flag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int nd = gm.ndim, nv = 1 << nd;
  int pos[RVS_MAX_GRID_DIM];
  double x[RVS_MAX_GRID_DIM];
  bool out = false;
  for (int i = 0; i < nd; i++) {
    const double qi = q[(int64_t)i * qstride + k];
    const double *u = gm.d_uvec + gm.uoff[i];
    const int n = gm.len[i];
    int lo = 0, hi = n;  // number of nodes <= qi (digitize, increasing bins); NaN -> n
    if (!(qi == qi)) lo = n;
    else
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (u[mid] <= qi) lo = mid + 1; else hi = mid;
      }
    const int p = lo - 1;
    if (p < 0 || p >= n - 1 || !isfinite(qi)) { out = true; pos[i] = 0; x[i] = 0; continue; }
    pos[i] = p;
    x[i] = (qi - u[p]) / (u[p + 1] - u[p]);
  }
  int32_t *idk = ids + (int64_t)k * nv;
  double *wk = w + (int64_t)k * nv;
  if (!out) {
    for (int c = 0; c < nv; c++) {
      int64_t flat = 0;
      double wc = 1;
      for (int i = 0; i < nd; i++) {
        const int s = (c >> (nd - 1 - i)) & 1;
        flat = flat * gm.len[i] + pos[i] + s;
        const double fct = s ? x[i] : 1 - x[i];
        wc = (i == 0) ? fct : wc * fct;
      }
      const int32_t id = gm.d_idgrid[flat];
      if (id < 0) out = true;
      idk[c] = id;
      wk[c] = wc;
    }
  }
  if (out) {
    for (int c = 0; c < nv; c++) { idk[c] = (c == 1) ? -1 : 0; wk[c] = (c == 0) ? 1.0 : 0.0; }
  }
  flag[k] = out ? 1 : 0;
}

// one warp per item; only flagged items with finite coordinates do any work
__global__ void __launch_bounds__(128) nearest_node_kernel(rvs_gridmap gm, const double *q,
                                                           int64_t qstride, int K, int32_t *ids,
                                                           double *w, int32_t *flag,
                                                           double *outside) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= K) return;
  if (flag[k] == 0) {
    if (lane == 0) outside[k] = 0.0;
    return;
  }
  const int nd = gm.ndim, nv = 1 << nd;
  double qn[RVS_MAX_GRID_DIM];
  bool finite = true;
  for (int i = 0; i < nd; i++) {
    const double qi = q[(int64_t)i * qstride + k];
    finite = finite && isfinite(qi);
    qn[i] = qi / gm.ptp[i];
  }
  if (!finite) {  // the host decides (first node, no finite off-grid measure)
    if (lane == 0) outside[k] = nan("");
    return;
  }
  double best = INFINITY;
  int bidx = 0x7fffffff;
  for (int node = lane; node < gm.nnode; node += 32) {
    const double *v = gm.d_vnorm + (int64_t)node * nd;
    double d2 = 0;
    for (int i = 0; i < nd; i++) {
      const double d = qn[i] - __ldg(v + i);
      d2 += d * d;
    }
    if (d2 < best) { best = d2; bidx = node; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ob < best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
  }
  if (lane == 0) {
    int32_t *idk = ids + (int64_t)k * nv;
    double *wk = w + (int64_t)k * nv;
    for (int c = 0; c < nv; c++) { idk[c] = c == 0 ? bidx : (c == 1 ? -1 : 0); wk[c] = (c == 0) ? 1.0 : 0.0; }
    outside[k] = sqrt(best);
    flag[k] = 0;
  }
}

}  // namespace rvs

extern "C" int rvs_locate_grid(const rvs_gridmap *gm, const double *d_q, int64_t q_stride, int K,
                               int32_t *d_ids, double *d_w, int32_t *d_flag, double *d_outside,
                               void *stream) {
  using namespace rvs;
  if (K == 0) return 0;
  RVS_REQUIRE(gm && gm->d_uvec && gm->d_idgrid && d_q && d_ids && d_w && d_flag, RVS_E_ARG,
              "rvs_locate_grid: null pointer");
  RVS_REQUIRE(gm->ndim >= 1 && gm->ndim <= RVS_MAX_GRID_DIM, RVS_E_ARG,
              "rvs_locate_grid: ndim=%d outside 1..%d", gm->ndim, RVS_MAX_GRID_DIM);
  locate_grid_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*gm, d_q, q_stride, K,
                                                                        d_ids, d_w, d_flag);
  RVS_LAUNCH_OK();
  if (d_outside) {
    RVS_REQUIRE(gm->d_vnorm && gm->nnode > 0, RVS_E_ARG, "rvs_locate_grid: node table missing");
    nearest_node_kernel<<<(K + 3) / 4, 128, 0, (cudaStream_t)stream>>>(*gm, d_q, q_stride, K, d_ids,
                                                                       d_w, d_flag, d_outside);
    RVS_LAUNCH_OK();
  }
  return 0;
}

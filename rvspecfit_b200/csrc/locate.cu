// Vertex location on the regular template grid (reference spec_inter.py:153-194):
// per dimension pos = digitize(q, u) - 1, the 2^d corner ids from idgrid and
// the weights prod_i x_i^S (1 - x_i)^(1-S), corners in itertools.product([0,1])
// order, factors multiplied in dimension order (np.prod).  Pure comparisons,
// one subtraction/division per dimension and d-1 multiplications per corner,
// so ids and weights are bit-identical to the host computation.  Points the
// reference resolves through its KD-tree (outside the grid -- the top edge counts
// as outside --, or a missing corner: nearest node in ptp-normalised space,
// spec_inter.py:128-132,156-167, and its distance as the off-grid measure,
// spec_inter.py:77-92) are resolved by the same warp with an exhaustive search
// of the node table.  Non-finite coordinates stay flagged for the host.
#include "common.cuh"

namespace rvs {

// One warp per item.  Lane i < ndim searches dimension i; lane c < 2^ndim owns
// corner c (flat index, weight, idgrid lookup); an off-grid or missing-corner
// point with finite coordinates is resolved in place when `outside` is given:
// the warp scans the node table for the nearest node (ties -> lowest index, as
// a stable argmin), writes it as a single-row item (ids = {node, -1, 0...},
// w = {1, 0...}) with its distance as the off-grid measure and clears the flag.
// Non-finite coordinates stay flagged for the host (outside = NaN).
__global__ void __launch_bounds__(128) locate_grid_kernel(rvs_gridmap gm, const double *q,
                                                          int64_t qstride, int K, int32_t *ids,
                                                          double *w, int32_t *flag,
                                                          double *outside) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= K) return;
  const int nd = gm.ndim, nv = 1 << nd;
  int mypos = 0, mynear = 0;
  double myx = 0, myq = 0;
  bool myout = false;
  if (lane < nd) {
    const double qi = q[(int64_t)lane * qstride + k];
    myq = qi;
    const double *u = gm.d_uvec + gm.uoff[lane];
    const int n = gm.len[lane];
    int lo = 0, hi = n;  // number of nodes <= qi (digitize, increasing bins); NaN -> n
    if (!(qi == qi)) lo = n;
    else
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(u + mid) <= qi) lo = mid + 1; else hi = mid;
      }
    const int p = lo - 1;
    {  // nearest grid coordinate of this dimension in the KD-tree's normalised metric
       // (lower index on a tie): used when the point turns out to be off the grid
      const int ka = min(max(p, 0), n - 1), kb = min(max(p + 1, 0), n - 1);
      const double qn = qi / gm.ptp[lane];
      const double da = fabs(qn - __ldg(u + ka) / gm.ptp[lane]);
      const double db = fabs(qn - __ldg(u + kb) / gm.ptp[lane]);
      mynear = db < da ? kb : ka;
    }
    if (p < 0 || p >= n - 1 || !isfinite(qi)) myout = true;
    else {
      mypos = p;
      myx = (qi - __ldg(u + p)) / (__ldg(u + p + 1) - __ldg(u + p));
    }
  }
  bool out = __any_sync(0xffffffffu, myout);
  int32_t id = 0;
  double wc = 1;
  {
    int64_t flat = 0;
    for (int i = 0; i < nd; i++) {
      const int pi = __shfl_sync(0xffffffffu, mypos, i);
      const double xi = __shfl_sync(0xffffffffu, myx, i);
      const int s = (lane >> (nd - 1 - i)) & 1;
      flat = flat * gm.len[i] + pi + s;
      const double fct = s ? xi : 1 - xi;
      wc = (i == 0) ? fct : wc * fct;
    }
    if (!out && lane < nv) id = __ldg(gm.d_idgrid + flat);
  }
  out = out || __any_sync(0xffffffffu, lane < nv && id < 0);
  if (!out) {
    if (lane < nv) {
      ids[(int64_t)k * nv + lane] = id;
      w[(int64_t)k * nv + lane] = wc;
    }
    if (lane == 0) {
      flag[k] = 0;
      if (outside) outside[k] = 0.0;
    }
    return;
  }
  // ---- off the grid (the top edge counts as outside) or a missing corner
  int node0 = 0, fl = 1;
  if (outside) {
    bool finite = true;
    double qn[RVS_MAX_GRID_DIM];
    for (int i = 0; i < nd; i++) {
      const double qi = __shfl_sync(0xffffffffu, myq, i);
      finite = finite && isfinite(qi);
      qn[i] = qi / gm.ptp[i];
    }
    double res = nan("");  // the host decides (first node, no finite off-grid measure)
    if (finite) {
      // The squared distance is separable over the dimensions, so the nearest point of
      // the FULL regular grid is the per-dimension nearest coordinate; if a template
      // exists there it is the nearest node (an optimiser walking along a grid edge asks
      // for this on most of its calls).  Otherwise (hole) search the node table.
      int64_t flat = 0;
      for (int i = 0; i < nd; i++) flat = flat * gm.len[i] + __shfl_sync(0xffffffffu, mynear, i);
      const int cand = __ldg(gm.d_idgrid + flat);
      double best = INFINITY;
      int bidx = 0x7fffffff;
      if (cand >= 0) {
        const double *v = gm.d_vnorm + (int64_t)cand * nd;
        double d2 = 0;
        for (int i = 0; i < nd; i++) {
          const double d = qn[i] - __ldg(v + i);
          d2 = fma(d, d, d2);
        }
        best = d2;
        bidx = cand;
      } else {
        for (int node = lane; node < gm.nnode; node += 32) {
          const double *v = gm.d_vnorm + (int64_t)node * nd;
          double d2 = 0;
          for (int i = 0; i < nd; i++) {
            const double d = qn[i] - __ldg(v + i);
            d2 = fma(d, d, d2);
          }
          if (d2 < best) { best = d2; bidx = node; }
        }
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
          if (ob < best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
        }
      }
      node0 = bidx;
      res = sqrt(best);
      fl = 0;
    }
    if (lane == 0) outside[k] = res;
  }
  if (lane < nv) {
    ids[(int64_t)k * nv + lane] = lane == 0 ? node0 : (lane == 1 ? -1 : 0);
    w[(int64_t)k * nv + lane] = lane == 0 ? 1.0 : 0.0;
  }
  if (lane == 0) flag[k] = fl;
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
  RVS_REQUIRE(!d_outside || (gm->d_vnorm && gm->nnode > 0), RVS_E_ARG,
              "rvs_locate_grid: node table missing");
  prof_begin(ST_LOCATE, (cudaStream_t)stream);
  locate_grid_kernel<<<(K + 3) / 4, 128, 0, (cudaStream_t)stream>>>(*gm, d_q, q_stride, K, d_ids,
                                                                    d_w, d_flag, d_outside);
  prof_end(ST_LOCATE, (cudaStream_t)stream);
  RVS_LAUNCH_OK();
  return 0;
}

// Fused optimiser-phase evaluation, stage B for batches whose objects share one
// wavelength grid (all spectra of a DESI arm do): the continuum normal
// equations of 8*NT items at a time as an FP64 tensor-core GEMM
//     [M | v](out, item) = sum_px A(out, px) B(px, item),
//     A = P_i P_j (packed lower triangle, rows 0..NTRI-1)  with  B = (T/sigma)^2,
//     A = P_i     (rows of the second tile set)            with  B = (T/sigma)(D/sigma),
// (spec_fit.py:230-241: Minv = sum P_i P_j T^2/sigma^2, v = sum P_i T S/sigma^2)
// issued as mma.sync m8n8k4 f64 (DMMA).  The basis products are formed on the
// fly from the pixel-major basis (L1-resident), every thread computing exactly
// the A elements of its own fragments; the shared basis is read once per 8*NT
// items instead of once per item, which is what bounded the per-item kernel
// (gram_kernel.cuh) at L2 bandwidth.
//
// Work decomposition: grid = (item groups, KS pixel splits), 4 warps per CTA,
// each warp a contiguous run of pixels.  Partial sums go to global memory;
// gram_solve_kernel (one warp per item) adds them IN FIXED ORDER and runs the
// Cholesky solve.  resid_mma_kernel evaluates |D - a^T G|^2 the same way:
// cont(px, item) = sum_i P_i(px) a_i(item) by DMMA, then
// r = D/sigma - T/sigma * cont (the reference's second pass, spec_fit.py:242-249);
// the last CTA of a group to finish (atomic ticket) adds the partial norms in
// fixed order, so results do not depend on scheduling.
#pragma once
#include "chisq_device.cuh"

namespace rvs {

constexpr int GM_WARPS = 4;
constexpr int GM_THREADS = GM_WARPS * 32;
constexpr int GM_MAX_KS = 16;
constexpr int GM_TILE = 16;  // pixels staged per warp and stage (one 128-byte line per operand row)

struct GramMmaArgs {
  const double *tn;
  int64_t tn_stride;
  const double *dn, *sumlog2;
  const int64_t *off, *goff;
  const int32_t *oix;
  const double *P;  // pixel-major [pixel][npp]
  int npp;
  int K;
  int KS;            // pixel splits (CTAs per group)
  double *part;      // [groups][KS][ROWS][8*NT] partial sums
  double *coef;      // [groups*8*NT][NP]
  double *logdet;    // [groups*8*NT]
  double *rpart;     // [groups][KS][8*NT] partial residual norms
  unsigned *ticket;  // [groups] zeroed before the launch (resid_mma_kernel)
  double *chisq;
  int32_t *status;
};

struct GramMmaArgsM {     // one record per arm of the call (chunk_kernel.cuh: RVS_MAX_ARMS)
  GramMmaArgs a[4];
};

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <int NP>
struct GramTiles {
  static constexpr int NTRI = NP * (NP + 1) / 2;
  static constexpr int MT_M = (NTRI + 7) / 8;  // tiles of basis products
  static constexpr int MT_V = (NP + 7) / 8;    // tiles of the basis itself
  static constexpr int MT = MT_M + MT_V;
  static constexpr int ROWS = MT * 8;
};

// Sum of the warps' accumulator fragments, added in warp order 0,1,2,3 (hence
// deterministic), left in s[row][col] for the whole CTA.  All threads call.
template <int MT, int NT>
__device__ __forceinline__ void warp_ordered_sum(const double (&acc)[MT][NT][2],
                                                 double (*s)[8 * NT + 1], int wid, int r, int c) {
  for (int w = 0; w < GM_WARPS; w++) {
    if (wid == w) {
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          double *d = &s[mt * 8 + r][nt * 8 + 2 * c];
          if (w == 0) { d[0] = acc[mt][nt][0]; d[1] = acc[mt][nt][1]; }
          else { d[0] += acc[mt][nt][0]; d[1] += acc[mt][nt][1]; }
        }
    }
    __syncthreads();
  }
}

// rows of the packed lower triangle this thread supplies to its A fragments:
// entry o = 8 mt + r  ->  (i, j), i >= j (clamped to a valid entry past the end;
// the caller zeroes those)
template <int NP>
__device__ __forceinline__ void tri_rows(int r, int (&ia)[GramTiles<NP>::MT_M],
                                         int (&ja)[GramTiles<NP>::MT_M]) {
  int i = 0;  // row of entry o: o and i only grow from one tile to the next
#pragma unroll
  for (int mt = 0; mt < GramTiles<NP>::MT_M; mt++) {
    const int o = min(mt * 8 + r, GramTiles<NP>::NTRI - 1);
    while ((i + 1) * (i + 2) / 2 <= o) i++;
    ia[mt] = i;
    ja[mt] = o - i * (i + 1) / 2;
  }
}

// reference item of a group (first present one) defines the shared grid
struct GroupGeom {
  int npix;
  int64_t goff;
};

template <int NT>
__device__ __forceinline__ GroupGeom group_geom(const GramMmaArgs &a, int g) {
  GroupGeom gg{0, 0};
  for (int e = 0; e < 8 * NT; e++) {
    const int k = g * 8 * NT + e;
    if (k >= a.K) break;
    const int obj = a.oix[k];
    if (obj >= 0) {
      gg.npix = (int)(a.off[obj + 1] - a.off[obj]);
      gg.goff = a.goff[obj];
      break;
    }
  }
  return gg;
}

// Shared-memory staging of gram_mma_kernel: per warp and stage a tile of GM_TILE
// pixels of the basis [px][npp] and of the B operands T/sigma, D/sigma [item][px]
// (rows padded by 4 doubles: the fragment reads of a half-warp then hit 16 distinct
// 8-byte banks).
constexpr int GM_BROW = GM_TILE + 4;
constexpr int GM_NSTG = 2;
__host__ __device__ constexpr int gm_stage_doubles(int npp, int ni) {
  return GM_TILE * npp + 2 * ni * GM_BROW;
}

template <int NP, int NT>
__global__ void __launch_bounds__(GM_THREADS, NT == 2 ? 4 : 3)
gram_mma_kernel(const __grid_constant__ GramMmaArgsM m) {
  const GramMmaArgs &a = m.a[blockIdx.z];
  using TL = GramTiles<NP>;
  constexpr int NI = 8 * NT;
  extern __shared__ __align__(16) double s_dyn[];
  __shared__ int64_t s_toff[NI], s_doff[NI];  // element offsets of the items' T/sigma, D/sigma rows
  __shared__ int s_have[NI];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = blockIdx.x, ks = blockIdx.y;
  const int r = lane >> 2, c = lane & 3;
  const GroupGeom gg = group_geom<NT>(a, g);
  const int npix = gg.npix;
  const double *Pb = a.P + gg.goff * a.npp;
  if (tid < NI) {
    const int k = g * NI + tid;
    const int obj = k < a.K ? a.oix[k] : -1;
    s_have[tid] = obj >= 0;
    s_toff[tid] = (int64_t)k * a.tn_stride;
    s_doff[tid] = obj >= 0 ? a.off[obj] : 0;
  }
  __syncthreads();
  // this thread's A rows: packed-triangle entry o = 8 mt + r  ->  (i, j), i >= j
  int ia[TL::MT_M], ja[TL::MT_M];
  tri_rows<NP>(r, ia, ja);
  double acc[TL::MT][NT][2];
#pragma unroll
  for (int mt = 0; mt < TL::MT; mt++)
#pragma unroll
    for (int nt = 0; nt < NT; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0;
  // pixel run of this warp: segment (ks, wid) of 4*KS, lengths a multiple of 4
  const int nseg = GM_WARPS * a.KS;
  const int seglen = ((npix + nseg - 1) / nseg + 3) & ~3;
  const int pbeg = (ks * GM_WARPS + wid) * seglen;
  const int pend = min(npix, pbeg + seglen);
  // Operands go through a warp-private, double-buffered shared-memory stage filled by
  // cp.async one tile (GM_TILE pixels = 4 k-steps) ahead: the basis rows in 16-byte
  // units (rows are npp = even doubles), T/sigma and D/sigma in 8-byte units (object
  // rows start at any pixel offset); the fragments are then LDS with short, fixed
  // latency instead of L2 round trips in the dependency chain of every k-step.
  const int stg = gm_stage_doubles(a.npp, NI);
  double *sW = s_dyn + (size_t)wid * GM_NSTG * stg;
  const unsigned sW_s = (unsigned)__cvta_generic_to_shared(sW);
  // lane -> (row parity, pixel) of the B tiles: rows (lane >> 4) + 2 i, pixel lane & 15
  const int bpx = lane & (GM_TILE - 1), brow0 = lane >> 4;
  auto load_stage = [&](int t0, int slot) {
    const int tile_n = min(GM_TILE, pend - t0);
    if (tile_n > 0) {
      const unsigned base = sW_s + slot * stg * 8;
      const double *src = Pb + (int64_t)t0 * a.npp;
      const int nvec = tile_n * a.npp / 2;
      for (int v = lane; v < nvec; v += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + v * 16), "l"(src + 2 * v) : "memory");
      double *sT = sW + slot * stg + GM_TILE * a.npp;
      const unsigned sT_s = base + GM_TILE * a.npp * 8;
#pragma unroll
      for (int i = 0; i < NI / 2; i++) {
        const int row = brow0 + 2 * i;
        const int o = row * GM_BROW + bpx;
        if (bpx < tile_n && s_have[row]) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sT_s + o * 8), "l"(a.tn + s_toff[row] + t0 + bpx) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sT_s + (NI * GM_BROW + o) * 8), "l"(a.dn + s_doff[row] + t0 + bpx) : "memory");
        } else {
          sT[o] = 0;
          sT[NI * GM_BROW + o] = 0;
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_stage(pbeg, 0);
  int slot = 0;
  for (int t0 = pbeg; t0 < pend; t0 += GM_TILE, slot ^= 1) {
    const int tile_n = min(GM_TILE, pend - t0);
    load_stage(t0 + GM_TILE, slot ^ 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    const double *sP = sW + slot * stg;
    const double *sT = sP + GM_TILE * a.npp, *sD = sT + NI * GM_BROW;
    for (int p4 = t0; p4 < t0 + tile_n; p4 += 4) {
      const int p = p4 + c;
      const bool in = p < pend;
      const double *Prow = sP + (in ? p - t0 : 0) * a.npp;
      double pa[TL::MT_M], pb[TL::MT_M], pv[TL::MT_V];
#pragma unroll
      for (int mt = 0; mt < TL::MT_M; mt++) {
        pa[mt] = Prow[ia[mt]];
        pb[mt] = Prow[ja[mt]];
      }
#pragma unroll
      for (int mv = 0; mv < TL::MT_V; mv++) pv[mv] = Prow[min(mv * 8 + r, NP - 1)];
      double bsq[NT], btd[NT];
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        // pixels past the run and absent items were staged as zeros
        const int o = (nt * 8 + r) * GM_BROW + (p4 - t0) + c;
        const double tq = sT[o], dq = sD[o];
        bsq[nt] = tq * tq;
        btd[nt] = tq * dq;
      }
#pragma unroll
      for (int mt = 0; mt < TL::MT_M; mt++) {
        const double av = (in && mt * 8 + r < TL::NTRI) ? pa[mt] * pb[mt] : 0.0;
#pragma unroll
        for (int nt = 0; nt < NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], av, bsq[nt]);
      }
#pragma unroll
      for (int mv = 0; mv < TL::MT_V; mv++) {
        const double av = (in && mv * 8 + r < NP) ? pv[mv] : 0.0;
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
          dmma884(acc[TL::MT_M + mv][nt][0], acc[TL::MT_M + mv][nt][1], av, btd[nt]);
      }
    }
    __syncwarp();  // every lane has read this stage: the next iteration refills it
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- cross-warp sum in fixed (warp) order (the staging memory is free now), partial
  // of this CTA to global
  double (*s_red)[NI + 1] = reinterpret_cast<double (*)[NI + 1]>(s_dyn);
  warp_ordered_sum<TL::MT, NT>(acc, s_red, wid, r, c);
  double *part = a.part + ((int64_t)g * a.KS + ks) * (TL::ROWS * NI);
  for (int e = tid; e < TL::ROWS * NI; e += GM_THREADS) {
    const int row = e / NI, col = e - row * NI;
    part[e] = s_red[row][col];
  }
}

// One warp per item: total of the pixel-split partials in fixed order, Cholesky
// solve (spec_fit.py:236-241), coefficients and log-determinant to global.
template <int NP, int NT>
__global__ void __launch_bounds__(GM_THREADS)
gram_solve_kernel(const __grid_constant__ GramMmaArgsM m) {
  const GramMmaArgs &a = m.a[blockIdx.y];
  using TL = GramTiles<NP>;
  constexpr int NI = 8 * NT;
  __shared__ double sM[GM_WARPS][TL::NTRI];
  __shared__ double sV[GM_WARPS][NP];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // tickets of resid_mma_kernel (launched next): one per group; this grid has at
  // least as many CTAs as there are groups
  if (threadIdx.x == 0 && blockIdx.x * NI < a.K) a.ticket[blockIdx.x] = 0;
  const int k = blockIdx.x * GM_WARPS + wid;
  if (k >= a.K) return;
  const int obj = a.oix[k];
  if (obj < 0) return;
  const int g = k / NI, e = k - g * NI;
  const GroupGeom gg = group_geom<NT>(a, g);
  const bool same = (int)(a.off[obj + 1] - a.off[obj]) == gg.npix && a.goff[obj] == gg.goff;
  const double *gp = a.part + (int64_t)g * a.KS * (TL::ROWS * NI);
  for (int o = lane; o < TL::NTRI + NP; o += 32) {
    const int row = o < TL::NTRI ? o : TL::MT_M * 8 + (o - TL::NTRI);
    double v[GM_MAX_KS];
#pragma unroll
    for (int s = 0; s < GM_MAX_KS; s++)
      v[s] = s < a.KS ? __ldcg(gp + (int64_t)s * (TL::ROWS * NI) + row * NI + e) : 0.0;
    double t = 0;
#pragma unroll
    for (int s = 0; s < GM_MAX_KS; s++) t += v[s];
    if (o < TL::NTRI) sM[wid][o] = t; else sV[wid][o - TL::NTRI] = t;
  }
  __syncwarp();
  const double ld = chol_solve<NP>(sM[wid], sV[wid], lane);
  if (lane < NP) a.coef[(int64_t)k * NP + lane] = sV[wid][lane];
  if (lane == 0) {
    a.logdet[k] = ld;
    if (!same) atomicOr(a.status + k, RVS_ST_LIMIT);  // not on the group's grid: general path
  }
}

template <int NP, int NT>
__global__ void __launch_bounds__(GM_THREADS)
resid_mma_kernel(const __grid_constant__ GramMmaArgsM m) {
  const GramMmaArgs &a = m.a[blockIdx.z];
  constexpr int NI = 8 * NT;
  constexpr int KST = (NP + 3) / 4;  // k-steps over the basis index
  __shared__ double s_red[GM_WARPS][NI];
  __shared__ unsigned s_last;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = blockIdx.x, ks = blockIdx.y;
  const int r = lane >> 2, c = lane & 3;
  const GroupGeom gg = group_geom<NT>(a, g);
  const int npix = gg.npix;
  const double *Pb = a.P + gg.goff * a.npp;
  // B fragments: coefficient i = 4 s + c of item r (+ 8 nt)
  double bco[KST][NT];
#pragma unroll
  for (int nt = 0; nt < NT; nt++) {
    const int k = g * NI + nt * 8 + r;
    const bool have = k < a.K && a.oix[k] >= 0;
#pragma unroll
    for (int s = 0; s < KST; s++) {
      const int i = 4 * s + c;
      bco[s][nt] = (have && i < NP) ? __ldcg(a.coef + (int64_t)k * NP + i) : 0.0;
    }
  }
  // C fragment columns: items 2c, 2c+1 of every n-tile
  const double *tnp[NT][2], *dnp[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; nt++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int k = g * NI + nt * 8 + 2 * c + e;
      const int obj = k < a.K ? a.oix[k] : -1;
      tnp[nt][e] = obj >= 0 ? a.tn + (int64_t)k * a.tn_stride : nullptr;
      dnp[nt][e] = obj >= 0 ? a.dn + a.off[obj] : nullptr;
    }
  double rss[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; nt++) rss[nt][0] = rss[nt][1] = 0;
  const int nseg = GM_WARPS * a.KS;
  const int seglen = ((npix + nseg - 1) / nseg + 7) & ~7;
  const int pbeg = (ks * GM_WARPS + wid) * seglen;
  const int pend = min(npix, pbeg + seglen);
  // operands one tile (8 pixels) ahead in registers
  double pav[KST], tv[NT][2], dv[NT][2];
  auto load_tile = [&](int p8) {
    const int p = p8 + r;
    const bool in = p < pend;
    const double *Prow = Pb + (int64_t)(in ? p : 0) * a.npp;
#pragma unroll
    for (int s = 0; s < KST; s++) {
      const int i = 4 * s + c;
      pav[s] = (in && i < NP) ? __ldg(Prow + i) : 0.0;
    }
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        tv[nt][e] = 0;
        dv[nt][e] = 0;
        if (in && tnp[nt][e]) { tv[nt][e] = __ldcg(tnp[nt][e] + p); dv[nt][e] = __ldg(dnp[nt][e] + p); }
      }
  };
  load_tile(pbeg);
  for (int p8 = pbeg; p8 < pend; p8 += 8) {
    double cont[NT][2], tc[NT][2], dc[NT][2], ac[KST];
#pragma unroll
    for (int s = 0; s < KST; s++) ac[s] = pav[s];
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) { tc[nt][e] = tv[nt][e]; dc[nt][e] = dv[nt][e]; cont[nt][e] = 0; }
    load_tile(p8 + 8);
#pragma unroll
    for (int s = 0; s < KST; s++)
#pragma unroll
      for (int nt = 0; nt < NT; nt++) dmma884(cont[nt][0], cont[nt][1], ac[s], bco[s][nt]);
    // absent items / pixels past the end carry t = d = 0: no contribution
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const double res = fma(-tc[nt][e], cont[nt][e], dc[nt][e]);
        rss[nt][e] = fma(res, res, rss[nt][e]);
      }
  }
  // sum over the 8 pixel rows of the fragment (lanes with equal c), then warps
#pragma unroll
  for (int nt = 0; nt < NT; nt++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      double v = rss[nt][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (r == 0) s_red[wid][nt * 8 + 2 * c + e] = v;
    }
  __syncthreads();
  double *rp = a.rpart + ((int64_t)g * a.KS + ks) * NI;
  if (tid < NI) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < GM_WARPS; w++) t += s_red[w][tid];
    rp[tid] = t;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(a.ticket + g, 1u) == (unsigned)(a.KS - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < NI) {
    const int k = g * NI + tid;
    if (k < a.K) {
      const int obj = a.oix[k];
      if (obj < 0) {
        a.chisq[k] = 0;
      } else {
        double t = 0;
        for (int s = 0; s < a.KS; s++) t += __ldcg(a.rpart + ((int64_t)g * a.KS + s) * NI + tid);
        const double chi = a.logdet[k] + a.sumlog2[obj] + t;
        a.chisq[k] = chi;
        if (!isfinite(chi)) atomicOr(a.status + k, RVS_ST_NOT_PD);
      }
    }
  }
}

// Scratch of the pair of kernels for K items (any npoly, any NT), carved out of
// the caller's workspace.  Sizes are per 8 items, with K rounded up to 16 (a
// group of NT = 2 holds 16): KS <= 16 partials of at most 152 rows (npoly 16).
struct GramScratch {
  int64_t g8;  // 8-item units
  __host__ __device__ explicit GramScratch(int64_t K) : g8(2 * ((K + 15) / 16)) {}
  __host__ __device__ int64_t part() const { return g8 * GM_MAX_KS * 152 * 8; }
  __host__ __device__ int64_t coef() const { return g8 * 8 * RVS_MAX_NPOLY; }
  __host__ __device__ int64_t logdet() const { return g8 * 8; }
  __host__ __device__ int64_t rpart() const { return g8 * GM_MAX_KS * 8; }
  __host__ __device__ int64_t ticket() const { return g8 + 8; }  // 2 x uint32 per unit
  __host__ __device__ int64_t total() const { return part() + coef() + logdet() + rpart() + ticket(); }
};

}  // namespace rvs

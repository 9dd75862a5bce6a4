// RV-scan chi-square as an FP64 tensor-core GEMM.  One CTA = one item (template +
// object) x 8*NT velocity trials.  The continuum normal equations of all the
// trials are one product
//     [M | v](out, trial) = sum_px A(out, px) B(px, trial)
// with A built from the object's continuum basis exactly as in gram_mma.cuh and
// B(px, trial) = (T_trial(px)/sigma)^2 resp. (T_trial(px)/sigma)(D/sigma), where
// every thread evaluates the Doppler-shifted spline for precisely the (px, trial)
// elements of its own B fragments (spec_fit.py:707-727, spliner.c:97-106), so no
// resampled template is ever stored.  The four warps split the pixels; partial
// sums meet in shared memory in fixed order; one warp per trial does the
// Cholesky solve; the residual norm |D - a^T G|^2 is a second sweep in which the
// continuum cont(px, trial) = sum_i P_i(px) a_i(trial) is again a DMMA product.
#pragma once
#include "gram_mma.cuh"

namespace rvs {

// spline value from (y,z) pairs: T = y0 v + y1 u + h^2/6 [z1 (u^3 - u) + z0 (v^3 - v)],
// algebraically spliner.c:97-106's A dl^3 + B dr^3 + C dl + D dr
__device__ __forceinline__ double spline_eval_uv(const ScanArgs &a, const double2 *yz, double x,
                                                 double q) {
  int pos = (int)((q - a.q0) * a.qstep_inv);
  pos = max(0, min(pos, a.npix_t - 2));
  if (a.fast_interp) return nearest_knot_value(a, yz, x, pos);
  const double2 c0 = __ldg(yz + pos), c1 = __ldg(yz + pos + 1);
  const double xl = __ldg(a.lam_t + pos);
  const double hh = __ldg(a.h + pos), hi = __ldg(a.hinv + pos);
  const double u = (x - xl) * hi, v = 1.0 - u;
  const double h26 = hh * hh * (1. / 6);
  const double cub = fma(c1.y * u, fma(u, u, -1.0), c0.y * v * fma(v, v, -1.0));
  return fma(h26, cub, fma(c0.x, v, c1.x * u));
}

// resident CTAs per SM the register budget allows (tuning)
#ifndef RVS_SCAN_MINB
#define RVS_SCAN_MINB 4
#endif
// Resolution matrices (diagonals within RS_HW of the main one; DESI's: 5; wider bands
// take the per-trial kernel, rvs_chisq_scan) are applied from shared memory (RESOL = true): each warp resamples the template once for a
// tile of RS_TILE of its pixels plus RS_HW neighbours on each side, for all the CTA's
// trials, and every (pixel, trial) element of its MMA fragments is then `nresol` FMAs on
// those values instead of `nresol` spline evaluations (template_at).  Row stride RS_W:
// 104 words = 8 banks between trials, conflict-free for the first sweep's fragment layout.
constexpr int RS_TILE = 32;
constexpr int RS_W = RS_TILE + 2 * RS_HW;

template <int NP, int NT, bool RESOL = false>
__global__ void __launch_bounds__(GM_THREADS, RVS_SCAN_MINB) chisq_scan_mma_kernel(ScanArgs a) {
  using TL = GramTiles<NP>;
  constexpr int NI = 8 * NT;
  constexpr int KST = (NP + 3) / 4;
  __shared__ double s_raw[RESOL ? GM_WARPS : 1][RESOL ? NI : 1][RESOL ? RS_W : 1];
  __shared__ double s_red[TL::ROWS][NI + 1];
  __shared__ double sM[GM_WARPS][TL::NTRI];
  __shared__ double sV[GM_WARPS][NP];
  __shared__ double s_co[NI][NP + 1];
  __shared__ double s_ld[NI], s_f[NI], s_qf[NI];
  __shared__ double s_rss[GM_WARPS][NI];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int r = lane >> 2, c = lane & 3;
  // 1-D grid, the trial blocks of an item adjacent: CTAs that run together read the
  // same template (y, z) pairs -- L1 / L2 hits instead of one pass over HBM per trial block
  const int k = blockIdx.x / a.nby;
  const int j0 = (blockIdx.x - k * a.nby) * NI;
  // ragged scans (refinement grids of different lengths padded to a.nv columns): trials
  // past the item's own count are not evaluated; their outputs are zeroed
  const int nvk = a.nvk ? min(a.nv, a.nvk[k]) : a.nv;
  if (j0 >= nvk) {
    if (tid < NI && j0 + tid < a.nv) {
      a.chisq[(int64_t)k * a.nv + j0 + tid] = 0.0;
      a.status[(int64_t)k * a.nv + j0 + tid] = 0;
    }
    return;
  }
  const int obj = a.oix[k];
  const int64_t p0 = a.off[obj];
  const int npix = (int)(a.off[obj + 1] - p0);
  const int64_t b0 = a.goff[obj];
  const double2 *yz = a.yz + (int64_t)a.tix[k] * a.yz_stride;
  const double *lam = a.lam + b0, *ql = (a.log_step ? a.loglam : a.lam) + b0;
  const double *Pb = a.P + b0 * a.npp;
  const double *dn = a.dn + p0, *einv = a.einv + p0;
  const double *rb = a.resol ? a.resol + p0 * a.nresol : nullptr;
  auto ev = [&](double x, double q) { return spline_eval_uv(a, yz, x, q); };
  // RESOL: resampled template of pixels [tile - RS_HW, tile + RS_TILE + RS_HW) for the
  // CTA's trials into the warp's rows of s_raw (0 outside the spectrum / absent trials)
  auto stage = [&](int tile) {
    __syncwarp();
    for (int idx = lane; idx < NI * RS_W; idx += 32) {
      const int t = idx / RS_W, q = idx - t * RS_W;
      const int pp = tile - RS_HW + q;
      double v = 0;
      if (pp >= 0 && pp < npix && j0 + t < nvk) {
        const double x = lam[pp] * s_f[t];
        v = ev(x, a.log_step ? ql[pp] + s_qf[t] : x);
      }
      s_raw[RESOL ? wid : 0][RESOL ? t : 0][RESOL ? q : 0] = v;
    }
    __syncwarp();
  };
  if (tid < NI) {
    const int j = j0 + tid;
    double f = 1, qf = 0;
    if (j < nvk) {
      const double beta = a.vels[(int64_t)k * a.nv + j] / RVS_C_KMS;
      f = sqrt((1 - beta) / (1 + beta));
      qf = a.log_step ? log(f) : 0.0;
    }
    s_f[tid] = f;
    s_qf[tid] = qf;
  }
  __syncthreads();
  // ---- sweep 1: normal equations
  {
    double fB[NT], qfB[NT];
    bool onB[NT];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
      fB[nt] = s_f[nt * 8 + r];
      qfB[nt] = s_qf[nt * 8 + r];
      onB[nt] = j0 + nt * 8 + r < nvk;
    }
    int ia[TL::MT_M], ja[TL::MT_M];
    tri_rows<NP>(r, ia, ja);
    double acc[TL::MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < TL::MT; mt++)
#pragma unroll
      for (int nt = 0; nt < NT; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0;
    const int seglen = ((npix + GM_WARPS - 1) / GM_WARPS + 3) & ~3;
    const int pbeg = wid * seglen, pend = min(npix, pbeg + seglen);
    for (int p4 = pbeg; p4 < pend; p4 += 4) {
      const int p = p4 + c;
      const bool in = p < pend;
      if (RESOL && ((p4 - pbeg) & (RS_TILE - 1)) == 0) stage(p4);
      double bsq[NT], btd[NT];
      if (in) {
        const double ei = einv[p], dv = dn[p];
        double tv[NT];
        if constexpr (RESOL) {
          const int col = ((p4 - pbeg) & (RS_TILE - 1)) + c + RS_HW;
#pragma unroll
          for (int nt = 0; nt < NT; nt++) tv[nt] = 0;
          for (int d = 0; d < a.nresol; d++) {
            const int o = __ldg(a.resol_offs + d);
            if (p + o < 0 || p + o >= npix) continue;
            const double cf = __ldg(rb + (int64_t)d * npix + p);
#pragma unroll
            for (int nt = 0; nt < NT; nt++) tv[nt] = fma(cf, s_raw[wid][nt * 8 + r][col + o], tv[nt]);
          }
        } else {
          const double lp = lam[p], qp = ql[p];
#pragma unroll
          for (int nt = 0; nt < NT; nt++) {
            tv[nt] = 0;
            if (onB[nt]) {
              const double x = lp * fB[nt];
              tv[nt] = spline_eval_uv(a, yz, x, a.log_step ? qp + qfB[nt] : x);
            }
          }
        }
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          const double t = onB[nt] ? tv[nt] * ei : 0.0;
          bsq[nt] = t * t;
          btd[nt] = t * dv;
        }
      } else {
#pragma unroll
        for (int nt = 0; nt < NT; nt++) bsq[nt] = btd[nt] = 0;
      }
      const double *Prow = Pb + (int64_t)(in ? p : 0) * a.npp;
#pragma unroll
      for (int mt = 0; mt < TL::MT_M; mt++) {
        double av = 0;
        if (in && mt * 8 + r < TL::NTRI) av = __ldg(Prow + ia[mt]) * __ldg(Prow + ja[mt]);
#pragma unroll
        for (int nt = 0; nt < NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], av, bsq[nt]);
      }
#pragma unroll
      for (int mv = 0; mv < TL::MT_V; mv++) {
        const int i = mv * 8 + r;
        double av = 0;
        if (in && i < NP) av = __ldg(Prow + i);
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
          dmma884(acc[TL::MT_M + mv][nt][0], acc[TL::MT_M + mv][nt][1], av, btd[nt]);
      }
    }
    warp_ordered_sum<TL::MT, NT>(acc, s_red, wid, r, c);
  }
  // ---- solve: one warp per trial
  for (int e = wid; e < NI; e += GM_WARPS) {
    if (j0 + e >= nvk) break;
    for (int o = lane; o < TL::NTRI + NP; o += 32) {
      const int row = o < TL::NTRI ? o : TL::MT_M * 8 + (o - TL::NTRI);
      const double t = s_red[row][e];
      if (o < TL::NTRI) sM[wid][o] = t; else sV[wid][o - TL::NTRI] = t;
    }
    __syncwarp();
    const double ld = chol_solve<NP>(sM[wid], sV[wid], lane);
    if (lane < NP) s_co[e][lane] = sV[wid][lane];
    if (lane == 0) s_ld[e] = ld;
    __syncwarp();
  }
  __syncthreads();
  // ---- sweep 2: residual norm
  {
    double bco[KST][NT];
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
#pragma unroll
      for (int s = 0; s < KST; s++) {
        const int i = 4 * s + c;
        bco[s][nt] = (i < NP && j0 + nt * 8 + r < nvk) ? s_co[nt * 8 + r][i] : 0.0;
      }
    double fC[NT][2], qfC[NT][2], rss[NT][2];
    bool onC[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int t = nt * 8 + 2 * c + e;
        fC[nt][e] = s_f[t];
        qfC[nt][e] = s_qf[t];
        onC[nt][e] = j0 + t < nvk;
        rss[nt][e] = 0;
      }
    const int seglen = ((npix + GM_WARPS - 1) / GM_WARPS + 7) & ~7;
    const int pbeg = wid * seglen, pend = min(npix, pbeg + seglen);
    for (int p8 = pbeg; p8 < pend; p8 += 8) {
      const int p = p8 + r;
      const bool in = p < pend;
      if (RESOL && ((p8 - pbeg) & (RS_TILE - 1)) == 0) stage(p8);
      const double *Prow = Pb + (int64_t)(in ? p : 0) * a.npp;
      double cont[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; nt++) cont[nt][0] = cont[nt][1] = 0;
#pragma unroll
      for (int s = 0; s < KST; s++) {
        const int i = 4 * s + c;
        const double av = (in && i < NP) ? __ldg(Prow + i) : 0.0;
#pragma unroll
        for (int nt = 0; nt < NT; nt++) dmma884(cont[nt][0], cont[nt][1], av, bco[s][nt]);
      }
      if (in) {
        const double ei = einv[p], dv = dn[p];
        double tv[NT][2];
        if constexpr (RESOL) {
          const int col = ((p8 - pbeg) & (RS_TILE - 1)) + r + RS_HW;
#pragma unroll
          for (int nt = 0; nt < NT; nt++) tv[nt][0] = tv[nt][1] = 0;
          for (int d = 0; d < a.nresol; d++) {
            const int o = __ldg(a.resol_offs + d);
            if (p + o < 0 || p + o >= npix) continue;
            const double cf = __ldg(rb + (int64_t)d * npix + p);
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
              for (int e = 0; e < 2; e++)
                tv[nt][e] = fma(cf, s_raw[wid][nt * 8 + 2 * c + e][col + o], tv[nt][e]);
          }
        } else {
          const double lp = lam[p], qp = ql[p];
#pragma unroll
          for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              tv[nt][e] = 0;
              if (onC[nt][e]) {
                const double x = lp * fC[nt][e];
                tv[nt][e] = spline_eval_uv(a, yz, x, a.log_step ? qp + qfC[nt][e] : x);
              }
            }
        }
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
          for (int e = 0; e < 2; e++)
            if (onC[nt][e]) {
              const double t = tv[nt][e] * ei;
              const double res = fma(-t, cont[nt][e], dv);
              rss[nt][e] = fma(res, res, rss[nt][e]);
            }
      }
    }
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        double v = rss[nt][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (r == 0) s_rss[wid][nt * 8 + 2 * c + e] = v;
      }
  }
  __syncthreads();
  if (tid < NI && j0 + tid >= nvk && j0 + tid < a.nv) {
    a.chisq[(int64_t)k * a.nv + j0 + tid] = 0.0;
    a.status[(int64_t)k * a.nv + j0 + tid] = 0;
  }
  if (tid < NI && j0 + tid < nvk) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < GM_WARPS; w++) t += s_rss[w][tid];
    const double chi = s_ld[tid] + a.sumlog2[obj] + t;
    int st = 0;
    // the reference checks the first and last evaluation points (spliner.c:78-83)
    const double xa = lam[0] * s_f[tid], xb = lam[npix - 1] * s_f[tid];
    if (xa < a.x0 || xb < a.x0 || xa >= a.xlast || xb >= a.xlast) st |= RVS_ST_RANGE;
    if (!isfinite(chi)) st |= RVS_ST_NOT_PD;
    a.chisq[(int64_t)k * a.nv + j0 + tid] = chi;
    a.status[(int64_t)k * a.nv + j0 + tid] = st;
  }
}

}  // namespace rvs

// Host side of one optimiser-phase evaluation call of the batched fit (no CUDA):
//   rvs_fit_pack     fitted vectors -> the (vel, vsini, mapped parameters) rows and
//                    arm-index rows of the call's pinned upload buffers, plus the
//                    additive host terms of the objective
//   rvs_fit_collect  the call's downloaded per-arm chi-squares / flags -> objective
//                    values, and which items need the general path
// Together they are vel_fit.chisq_func around spec_fit.get_chisq (reference
// vel_fit.py:154-198 ParamMapper.forward, :210-257 priors, vsini penalty and hard
// walls; spec_fit.py:863,879,895-896 the per-arm sum and off-grid penalty), for K
// (object, vector) pairs in one pass over the data -- the numpy restatement
// (batch_fit.BatchObjective + LikelihoodEngine._submit_fast/_collect_fast) costs
// ~40 array operations per call, which is what bounded the fit throughput once the
// kernels were fast.  Arithmetic follows the numpy expressions term by term (no
// contraction: built with -ffp-contract=off), so both routes give identical values.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/rvs_b200.h"

extern "C" int rvs_fit_pack(const rvs_fit_layout *L, int64_t K, int64_t Kp, const int32_t *h_obj,
                            const double *h_X, const double *h_logvals, double *h_in,
                            int32_t *h_oix, double *h_prior, double *h_pen, uint8_t *h_wall,
                            double *h_vsini_max) {
  if (!L || K < 0 || Kp < K || !h_obj || !h_X || !h_in || !h_oix || !h_prior || !h_pen || !h_wall)
    return RVS_E_ARG;
  const int N = L->nfit, ns = L->nspec, narm = L->narm;
  if (ns < 1 || ns > 30 || N < 1) return RVS_E_ARG;
  double *vel = h_in, *vs = h_in + Kp, *q = h_in + 2 * Kp;
  double vmax = 0.0;
  int nlog_fit = 0;      // log-mapped FITTED parameters seen so far index h_logvals rows
  (void)nlog_fit;
  for (int64_t k = 0; k < K; k++) {
    const double *x = h_X + k * N;
    const int32_t o = h_obj[k];
    if (o < 0 || o >= L->nobj) return RVS_E_ARG;
    int pos = 1;
    const double v = x[0];
    double vsini = 0.0, pen = 0.0;
    if (L->fit_vsini) {
      const double xv = x[1];
      vsini = xv < 0 ? 0.0 : (xv > L->max_vsini ? L->max_vsini : xv);   // np.clip
      const double d = vsini - xv;
      pen = (xv < 0 ? d * d : 0.0) + (xv > L->max_vsini ? d * d : 0.0);
      pos = 2;
    } else if (L->has_vsini) {
      vsini = L->h_vsini0[o];
    }
    bool wall = v > L->max_vel || v < L->min_vel;
    double prior = 0.0;
    int lrow = 0;
    for (int j = 0; j < ns; j++) {
      double p, qj;
      if (L->fixmask >> j & 1) {
        p = L->h_p0[(int64_t)o * ns + j];
        qj = L->h_q0[(int64_t)o * ns + j];
      } else {
        p = x[pos++];
        if (L->logmask >> j & 1) qj = h_logvals[(int64_t)lrow++ * K + k];
        else qj = p;
      }
      if (!isfinite(p)) wall = true;
      if (L->priormask >> j & 1) {
        const double t = (L->h_prior_mu[j] - p) / L->h_prior_sig[j];
        prior = prior + t * t;
      }
      q[(int64_t)j * Kp + k] = qj;
    }
    if (pos != N) return RVS_E_ARG;
    h_wall[k] = wall;
    h_prior[k] = prior;
    h_pen[k] = pen;
    if (wall) {      // not evaluated: absent on every arm, harmless coordinates
      vel[k] = 0.0;
      vs[k] = 0.0;
      for (int j = 0; j < ns; j++) q[(int64_t)j * Kp + k] = L->h_q0[(int64_t)o * ns + j];
      for (int a = 0; a < narm; a++) h_oix[(int64_t)a * Kp + k] = -1;
      continue;
    }
    vel[k] = v;
    vs[k] = vsini;
    if (vsini > vmax) vmax = vsini;
    for (int a = 0; a < narm; a++) h_oix[(int64_t)a * Kp + k] = L->h_oix[(int64_t)a * L->nobj + o];
  }
  // padding items (launch configurations are rounded up so that captured graphs are
  // reused): absent everywhere, coordinates of the first item
  for (int64_t k = K; k < Kp; k++) {
    vel[k] = 0.0;
    vs[k] = 0.0;
    for (int j = 0; j < ns; j++) q[(int64_t)j * Kp + k] = K ? q[(int64_t)j * Kp] : 0.0;
    for (int a = 0; a < narm; a++) h_oix[(int64_t)a * Kp + k] = -1;
  }
  if (h_vsini_max) *h_vsini_max = vmax;
  return 0;
}

extern "C" int64_t rvs_fit_collect(const rvs_fit_layout *L, int64_t K, int64_t Kp,
                                   const int32_t *h_obj, const double *h_in, const double *h_chi,
                                   const int32_t *h_flags, int shared_locate, int outside_penalty,
                                   const double *h_prior, const double *h_pen,
                                   const uint8_t *h_wall, double *h_out, uint8_t *h_redo) {
  if (!L || !h_obj || !h_chi || !h_flags || !h_out || !h_redo) return RVS_E_ARG;
  const int narm = L->narm;
  const double *chi = h_chi, *outside = h_chi + (int64_t)narm * Kp;
  const int32_t *f0 = h_flags, *f1 = h_flags + (int64_t)narm * Kp;
  int64_t nredo = 0;
  for (int64_t k = 0; k < K; k++) {
    if (h_wall && h_wall[k]) {
      h_out[k] = 1e30;       // vel_fit.py:252-254
      h_redo[k] = 0;
      continue;
    }
    const int32_t o = h_obj[k];
    const double v = h_in[k];
    bool redo = !L->h_cover[o] || v < L->min_vel || v > L->max_vel;
    double tot = 0.0;
    for (int a = 0; a < narm; a++) {
      const int64_t i = (int64_t)a * Kp + k, i1 = shared_locate ? k : i;
      double c = chi[i];
      const double out = outside[i1];
      if (f0[i] != 0 || f1[i1] != 0 || !isfinite(c) || !isfinite(out)) redo = true;
      // off-grid points were resolved on the device (nearest node); their penalty is
      // added once per arm the object has (spec_fit.py:879,895-896)
      if (outside_penalty && out != 0.0 && L->h_oix[(int64_t)a * L->nobj + o] >= 0)
        c = out * L->h_badchi[o] + c;
      tot = a == 0 ? c : tot + c;
    }
    h_redo[k] = redo;
    nredo += redo;
    h_out[k] = (h_prior ? h_prior[k] + tot : tot) + (h_pen ? h_pen[k] : 0.0);
  }
  return nredo;
}

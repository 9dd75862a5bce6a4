// Lock-step Nelder-Mead stepper (host code, no CUDA): the optimiser loop of
// vel_fit.process (reference vel_fit.py:628-650 calls scipy.optimize.minimize(
// method='Nelder-Mead')) for B simplices at once.  The decision rules are scipy's
// `_minimize_neldermead` (adaptive=False, no bounds, maxfev=inf): rho, chi, psi,
// sigma = 1, 2, 1/2, 1/2, the same termination test, acceptance rules and
// ordering, and the same floating-point expressions, so that problem b visits
// exactly the points, in exactly the order, scipy would.  The driver asks for
// the next batch of trial points (rvs_nm_request), evaluates them in one launch
// per arm, and feeds the values back (rvs_nm_feed); between the two only this
// file runs, which replaces ~80 numpy calls per round of the Python restatement
// (batch_fit.nelder_mead_steps) and is what keeps the host ahead of the GPU.
//
// Compiled WITHOUT floating-point contraction (-ffp-contract=off): a fused
// multiply-add in `2*xbar - last` would change the visited points.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/rvs_b200.h"

namespace {

enum Phase { P_INIT, P_REFLECT, P_SPEC, P_SECOND, P_SHRINK, P_DONE };

struct Nm {
  int B, N, N1;
  double xatol, fatol;
  int64_t maxiter;
  std::vector<double> sim, fsim;       // [B][N1][N], [B][N1]
  std::vector<int64_t> nit, nfev;
  std::vector<uint8_t> active, success;
  Phase phase = P_INIT;
  // the iteration in progress
  std::vector<int32_t> a;              // live problems
  std::vector<double> xbar, xr, fxr;   // [na][N], [na][N], [na]
  std::vector<double> x2, f2;          // second point (where asked) [na][N], [na]
  std::vector<uint8_t> kind;           // per live problem: 0 accept reflection, 1 expand,
                                       // 2 contract outside, 3 contract inside
  std::vector<int32_t> second;         // positions in `a` that asked for a second point
  std::vector<int32_t> shrink;         // positions in `a` that shrink
  bool speculated = false;
  int64_t pending = 0;                 // points of the outstanding request

  double *S(int b) { return sim.data() + (size_t)b * N1 * N; }
  double *F(int b) { return fsim.data() + (size_t)b * N1; }
};

// numpy's sort order: NaNs last
inline bool np_less(double x, double y) { return x < y || (y != y && x == x); }

// stable ordering of a problem's vertices by function value (np.argsort(kind='stable'))
void sort_simplex(Nm &m, int b) {
  double *s = m.S(b), *f = m.F(b);
  const int N = m.N, N1 = m.N1;
  double tmp[64];
  for (int i = 1; i < N1; i++) {
    const double fi = f[i];
    int j = i;
    while (j > 0 && np_less(fi, f[j - 1])) j--;
    if (j == i) continue;
    memcpy(tmp, s + (size_t)i * N, sizeof(double) * N);
    memmove(s + (size_t)(j + 1) * N, s + (size_t)j * N, sizeof(double) * N * (i - j));
    memmove(f + j + 1, f + j, sizeof(double) * (i - j));
    memcpy(s + (size_t)j * N, tmp, sizeof(double) * N);
    f[j] = fi;
  }
}

// np.max over an array: NaN if any element is NaN
inline void np_max_acc(double &acc, double v) {
  if (acc != acc) return;
  if (v != v || v > acc) acc = v;
}

// loop head of an iteration: termination tests, then the live set and its centroids
void begin_iteration(Nm &m) {
  const int N = m.N, N1 = m.N1;
  m.a.clear();
  for (int b = 0; b < m.B; b++) {
    if (!m.active[b]) continue;
    const double *s = m.S(b), *f = m.F(b);
    const bool out_of_iters = m.nit[b] >= m.maxiter;
    double dx = -INFINITY, df = -INFINITY;
    for (int j = 1; j < N1; j++) {
      for (int i = 0; i < N; i++) np_max_acc(dx, fabs(s[(size_t)j * N + i] - s[i]));
      np_max_acc(df, fabs(f[0] - f[j]));
    }
    const bool conv = dx <= m.xatol && df <= m.fatol;
    if (out_of_iters || conv) {
      m.active[b] = 0;
      if (conv && !out_of_iters) m.success[b] = 1;
      continue;
    }
    m.a.push_back(b);
  }
  const size_t na = m.a.size();
  m.xbar.resize(na * N);
  m.xr.resize(na * N);
  m.fxr.resize(na);
  m.x2.resize(na * N);
  m.f2.resize(na);
  m.kind.resize(na);
  for (size_t k = 0; k < na; k++) {
    const double *s = m.S(m.a[k]);
    double *xb = m.xbar.data() + k * N, *xr = m.xr.data() + k * N;
    const double *last = s + (size_t)N * N;
    for (int i = 0; i < N; i++) {
      double acc = s[i];                         // np.add.reduce(sim[:-1], 0): row by row
      for (int j = 1; j < N; j++) acc = acc + s[(size_t)j * N + i];
      xb[i] = acc / N;
      xr[i] = 2.0 * xb[i] - 1.0 * last[i];       // (1 + rho) * xbar - rho * sim[-1]
    }
  }
  m.phase = na ? P_REFLECT : P_DONE;
}

inline void second_point(const Nm &m, int kind, const double *xb, const double *last, double *out) {
  const int N = m.N;
  for (int i = 0; i < N; i++) {
    if (kind == 1) out[i] = 3.0 * xb[i] - 2.0 * last[i];        // (1 + rho chi) xbar - rho chi last
    else if (kind == 2) out[i] = 1.5 * xb[i] - 0.5 * last[i];   // (1 + psi rho) xbar - psi rho last
    else out[i] = 0.5 * xb[i] + 0.5 * last[i];                  // (1 - psi) xbar + psi last
  }
}

// classify every live problem from its reflection value
void classify(Nm &m) {
  const size_t na = m.a.size();
  m.second.clear();
  for (size_t k = 0; k < na; k++) {
    const double *f = m.F(m.a[k]);
    const double fx = m.fxr[k];
    int kind;
    if (fx < f[0]) kind = 1;
    else if (fx < f[m.N1 - 2]) kind = 0;
    else if (fx < f[m.N1 - 1]) kind = 2;
    else kind = 3;
    m.kind[k] = (uint8_t)kind;
    m.nfev[m.a[k]] += 1;
    if (kind != 0) {
      m.second.push_back((int32_t)k);
      m.nfev[m.a[k]] += 1;
    }
  }
}

// after the second points are known: replace the worst vertex or mark for shrinking
void resolve(Nm &m) {
  const int N = m.N, N1 = m.N1;
  const size_t na = m.a.size();
  m.shrink.clear();
  for (size_t k = 0; k < na; k++) {
    const int b = m.a[k];
    double *s = m.S(b), *f = m.F(b);
    const double fx = m.fxr[k], f2 = m.f2[k];
    const int kind = m.kind[k];
    bool take2 = false, shr = false;
    if (kind == 1) take2 = f2 < fx;
    else if (kind == 2) { take2 = f2 <= fx; shr = !take2; }
    else if (kind == 3) { take2 = f2 < f[N1 - 1]; shr = !take2; }
    if (shr) { m.shrink.push_back((int32_t)k); continue; }
    const double *src = take2 ? m.x2.data() + k * N : m.xr.data() + k * N;
    memcpy(s + (size_t)N * N, src, sizeof(double) * N);
    f[N1 - 1] = take2 ? f2 : fx;
  }
  for (int32_t k : m.shrink) {   // sim[j] = sim[0] + sigma * (sim[j] - sim[0])
    double *s = m.S(m.a[k]);
    for (int j = 1; j < N1; j++)
      for (int i = 0; i < N; i++) s[(size_t)j * N + i] = s[i] + 0.5 * (s[(size_t)j * N + i] - s[i]);
  }
}

void end_iteration(Nm &m) {
  for (int32_t b : m.a) {
    m.nit[b] += 1;
    sort_simplex(m, b);
  }
  begin_iteration(m);
}

}  // namespace

extern "C" void *rvs_nm_create(int B, int N, const double *h_sims, double xatol, double fatol,
                               int64_t maxiter) {
  if (B < 0 || N < 1 || N > 63 || !h_sims) return nullptr;
  Nm *m = new Nm;
  m->B = B; m->N = N; m->N1 = N + 1;
  m->xatol = xatol; m->fatol = fatol; m->maxiter = maxiter;
  m->sim.assign(h_sims, h_sims + (size_t)B * (N + 1) * N);
  m->fsim.assign((size_t)B * (N + 1), 0.0);
  m->nit.assign(B, 1);
  m->nfev.assign(B, N + 1);
  m->active.assign(B, 1);
  m->success.assign(B, 0);
  m->phase = B ? P_INIT : P_DONE;
  return m;
}

extern "C" void rvs_nm_destroy(void *h) { delete static_cast<Nm *>(h); }

extern "C" int64_t rvs_nm_request(void *h, int speculate_below, int32_t *h_idx, double *h_X,
                                  int64_t cap) {
  Nm &m = *static_cast<Nm *>(h);
  const int N = m.N, N1 = m.N1;
  int64_t n = 0;
  auto put = [&](int b, const double *x) {
    if (n < cap) {
      h_idx[n] = b;
      memcpy(h_X + n * N, x, sizeof(double) * N);
    }
    n++;
  };
  switch (m.phase) {
    case P_INIT:
      for (int b = 0; b < m.B; b++)
        for (int j = 0; j < N1; j++) put(b, m.S(b) + (size_t)j * N);
      break;
    case P_REFLECT: {
      const size_t na = m.a.size();
      m.speculated = (int64_t)na <= speculate_below;
      for (size_t k = 0; k < na; k++) put(m.a[k], m.xr.data() + k * N);
      if (m.speculated) {
        // few problems left: launches are latency-bound, so the three candidate second
        // points go out with the reflection; each problem then uses exactly the value
        // scipy would have computed
        std::vector<double> tmp(N);
        for (int kind = 1; kind <= 3; kind++)
          for (size_t k = 0; k < na; k++) {
            second_point(m, kind, m.xbar.data() + k * N, m.S(m.a[k]) + (size_t)N * N, tmp.data());
            put(m.a[k], tmp.data());
          }
        m.phase = P_SPEC;
      }
      break;
    }
    case P_SECOND:
      for (int32_t k : m.second) put(m.a[k], m.x2.data() + (size_t)k * N);
      break;
    case P_SHRINK:
      for (int32_t k : m.shrink)
        for (int j = 1; j < N1; j++) put(m.a[k], m.S(m.a[k]) + (size_t)j * N);
      break;
    default:
      break;
  }
  m.pending = n;
  return n;
}

extern "C" int rvs_nm_feed(void *h, const double *h_f, int64_t K) {
  Nm &m = *static_cast<Nm *>(h);
  if (K != m.pending) return RVS_E_ARG;
  const int N = m.N, N1 = m.N1;
  const size_t na = m.a.size();
  switch (m.phase) {
    case P_INIT:
      memcpy(m.fsim.data(), h_f, sizeof(double) * (size_t)K);
      for (int b = 0; b < m.B; b++) sort_simplex(m, b);
      begin_iteration(m);
      return 0;
    case P_REFLECT:
    case P_SPEC: {
      for (size_t k = 0; k < na; k++) m.fxr[k] = h_f[k];
      classify(m);
      for (int32_t k : m.second)
        second_point(m, m.kind[k], m.xbar.data() + (size_t)k * N,
                     m.S(m.a[k]) + (size_t)N * N, m.x2.data() + (size_t)k * N);
      if (m.phase == P_SPEC) {
        for (int32_t k : m.second) m.f2[k] = h_f[(size_t)m.kind[k] * na + k];
      } else if (!m.second.empty()) {
        m.phase = P_SECOND;
        return 0;
      }
      break;
    }
    case P_SECOND: {
      size_t j = 0;
      for (int32_t k : m.second) m.f2[k] = h_f[j++];
      break;
    }
    case P_SHRINK: {
      size_t j = 0;
      for (int32_t k : m.shrink) {
        double *f = m.F(m.a[k]);
        for (int v = 1; v < N1; v++) f[v] = h_f[j++];
        m.nfev[m.a[k]] += N;
      }
      end_iteration(m);
      return 0;
    }
    default:
      return RVS_E_ARG;
  }
  resolve(m);
  if (!m.shrink.empty()) {
    m.phase = P_SHRINK;
    return 0;
  }
  end_iteration(m);
  return 0;
}

extern "C" int64_t rvs_nm_live(void *h, uint8_t *h_active) {
  Nm &m = *static_cast<Nm *>(h);
  int64_t n = 0;
  for (int b = 0; b < m.B; b++) {
    n += m.active[b] != 0;
    if (h_active) h_active[b] = m.active[b];
  }
  return n;
}

extern "C" int rvs_nm_result(void *h, double *h_x, double *h_fun, uint8_t *h_success,
                             double *h_final_simplex, int64_t *h_nit, int64_t *h_nfev) {
  // Before every problem has stopped the rows of the stopped ones (rvs_nm_live) are final.
  Nm &m = *static_cast<Nm *>(h);
  const int N = m.N, N1 = m.N1;
  for (int b = 0; b < m.B; b++) {
    if (h_x) memcpy(h_x + (size_t)b * N, m.S(b), sizeof(double) * N);
    if (h_fun) {
      double mn = m.F(b)[0];      // np.min: NaN if any
      for (int j = 1; j < N1; j++) {
        const double v = m.F(b)[j];
        if (mn == mn && (v != v || v < mn)) mn = v;
      }
      h_fun[b] = mn;
    }
    if (h_success) h_success[b] = m.success[b];
    if (h_nit) h_nit[b] = m.nit[b];
    if (h_nfev) h_nfev[b] = m.nfev[b];
  }
  if (h_final_simplex) memcpy(h_final_simplex, m.sim.data(), sizeof(double) * m.sim.size());
  return 0;
}

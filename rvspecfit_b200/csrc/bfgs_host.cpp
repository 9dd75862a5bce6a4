// Lock-step BFGS stepper (host code, no CUDA): the polish step of vel_fit.process
// (reference vel_fit.py:653-658 calls scipy.optimize.minimize(method='BFGS',
// options=dict(hess_inv0=...)) with a forward-difference gradient) for B problems at
// once.  The algorithm is scipy's `_minimize_bfgs` with the MINPACK line search DCSRCH
// and its fallback `line_search_wolfe2` (scipy/optimize/_optimize.py, _linesearch.py,
// _dcsrch.py, _numdiff.py), the same constants, the same tests in the same order, as
// restated over arrays of problems in batch_bfgs.bfgs_steps -- here one problem at a time
// in plain C++, so that a round of the stage costs microseconds instead of a millisecond of
// numpy and the round loop can run without the interpreter (drive_host.cpp).
//
// Protocol as the Nelder-Mead stepper's: rvs_bfgs_request writes the next batch of trial
// points (N + 1 consecutive points per searching problem: a point and its
// forward-difference neighbours), rvs_bfgs_feed takes their values.
//
// Rounding: scipy forms H g, g.p and the rank-two update through BLAS, whose summation
// order is the library's business; this file sums in index order without contraction
// (-ffp-contract=off).  Problems therefore follow scipy's decisions with values that
// differ in the last bits; batch_bfgs.bfgs_steps remains the bit-exact restatement
// (tests/test_batch_drivers.py holds both against scipy, this one to tolerance).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/rvs_b200.h"

namespace {

constexpr double P5 = 0.5, P66 = 0.66, XTRAPL = 1.1, XTRAPU = 4.0;
enum Task { T_FG = 0, T_CONV = 1, T_WARN = 2, T_ERROR = 3 };

inline double pymax(double a, double b) { return b > a ? b : a; }   // Python max(a, b)
inline double pymin(double a, double b) { return b < a ? b : a; }   // Python min(a, b)
inline double npsign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : (x == 0 ? 0.0 : x)); }

struct Dcsrch {
  double finit, ginit, gtest, width, width1, stx, fx, gx, sty, fy, gy, stmin, stmax;
  bool brackt;
  int stage;
};

// MINPACK-2 dcstep (_dcsrch.py)
void dcstep(double &stx, double &fx, double &dx, double &sty, double &fy, double &dy, double &stp,
            double fp, double dp, bool &brackt, double stpmin, double stpmax) {
  const double sgnd = npsign(dp) * npsign(dx);
  double stpf;
  if (fp > fx) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = pymax(pymax(fabs(theta), fabs(dx)), fabs(dp));
    double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    const double p = (gamma - dx) + theta;
    const double q = ((gamma - dx) + gamma) + dp;
    const double r = p / q;
    const double stpc = stx + r * (stp - stx);
    const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    if (fabs(stpc - stx) <= fabs(stpq - stx)) stpf = stpc;
    else stpf = stpc + (stpq - stpc) / 2.0;
    brackt = true;
  } else if (sgnd < 0.0) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = pymax(pymax(fabs(theta), fabs(dx)), fabs(dp));
    double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    const double p = (gamma - dp) + theta;
    const double q = ((gamma - dp) + gamma) + dx;
    const double r = p / q;
    const double stpc = stp + r * (stx - stp);
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
    else stpf = stpq;
    brackt = true;
  } else if (fabs(dp) < fabs(dx)) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    const double s = pymax(pymax(fabs(theta), fabs(dx)), fabs(dp));
    const double rad = (theta / s) * (theta / s) - (dx / s) * (dp / s);
    double gamma = s * sqrt(rad > 0 ? rad : 0.0);
    if (stp > stx) gamma = -gamma;
    const double p = (gamma - dp) + theta;
    const double q = (gamma + (dx - dp)) + gamma;
    const double r = p / q;
    double stpc;
    if (r < 0 && gamma != 0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      if (fabs(stpc - stp) < fabs(stpq - stp)) stpf = stpc;
      else stpf = stpq;
      const double lim = stp + 0.66 * (sty - stp);
      if (stp > stx) stpf = pymin(lim, stpf);
      else stpf = pymax(lim, stpf);
    } else {
      if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
      else stpf = stpq;
      // np.clip(stpf, stpmin, stpmax)
      stpf = stpf < stpmin ? stpmin : (stpf > stpmax ? stpmax : stpf);
    }
  } else {
    if (brackt) {
      const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      const double s = pymax(pymax(fabs(theta), fabs(dy)), fabs(dp));
      double gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      const double p = (gamma - dp) + theta;
      const double q = ((gamma - dp) + gamma) + dy;
      const double r = p / q;
      stpf = stp + r * (sty - stp);
    } else if (stp > stx) {
      stpf = stpmax;
    } else {
      stpf = stpmin;
    }
  }
  // update the interval that contains a minimiser
  if (fp > fx) {
    sty = stp; fy = fp; dy = dp;
  } else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  stp = stpf;
}

// fallback search: scipy's scalar_search_wolfe2 / _zoom as a state machine.  Every step
// length it asks for gets the function value AND the slope at once.
struct Wolfe2 {
  enum State { BRACKET, ZOOM } state;
  double phi0, derphi0, c1, c2, amax;
  int i;                                  // iteration of the bracketing loop
  double alpha0, alpha1, phi_a0, derphi_a0;
  // zoom
  int zi;
  double a_lo, a_hi, phi_lo, phi_hi, derphi_lo, phi_rec, a_rec, a_j;
};

bool cubicmin(double a, double fa, double fpa, double b, double fb, double c, double fc, double *out) {
  const double db = b - a, dc = c - a;
  const double denom = (db * dc) * (db * dc) * (db - dc);
  if (denom == 0 || !isfinite(denom)) return false;
  const double d00 = dc * dc, d01 = -(db * db), d10 = -(dc * dc * dc), d11 = db * db * db;
  const double v0 = fb - fa - fpa * db, v1 = fc - fa - fpa * dc;
  double A = d00 * v0 + d01 * v1, B = d10 * v0 + d11 * v1;
  A /= denom;
  B /= denom;
  const double rad = B * B - 3 * A * fpa;
  if (!(rad >= 0) || A == 0) return false;
  const double xmin = a + (-B + sqrt(rad)) / (3 * A);
  if (!isfinite(xmin)) return false;
  *out = xmin;
  return true;
}

bool quadmin(double a, double fa, double fpa, double b, double fb, double *out) {
  const double db = b - a * 1.0;
  if (db == 0) return false;
  const double B = (fb - fa - fpa * db) / (db * db);
  if (B == 0 || !isfinite(B)) return false;
  const double xmin = a - fpa / (2.0 * B);
  if (!isfinite(xmin)) return false;
  *out = xmin;
  return true;
}

// next trial step of the zoom phase
double zoom_next(Wolfe2 &w) {
  const double dalpha = w.a_hi - w.a_lo;
  double a, b;
  if (dalpha < 0) { a = w.a_hi; b = w.a_lo; } else { a = w.a_lo; b = w.a_hi; }
  double a_j = 0;
  bool have = false;
  if (w.zi > 0) {
    const double cchk = 0.2 * dalpha;
    have = cubicmin(w.a_lo, w.phi_lo, w.derphi_lo, w.a_hi, w.phi_hi, w.a_rec, w.phi_rec, &a_j);
    if (have && (a_j > b - cchk || a_j < a + cchk)) have = false;
  }
  if (!have) {
    const double qchk = 0.1 * dalpha;
    have = quadmin(w.a_lo, w.phi_lo, w.derphi_lo, w.a_hi, w.phi_hi, &a_j);
    if (!have || a_j > b - qchk || a_j < a + qchk) a_j = w.a_lo + 0.5 * dalpha;
  }
  w.a_j = a_j;
  return a_j;
}

struct Prob {
  double fval, old_old, derphi0, stp, gnorm;
  int64_t nit;
  int status, ls_it;
  int mode;       // 0 done, 1 MINPACK search, 2 fallback search
  Dcsrch ls;
  Wolfe2 fb;
};

struct Bfgs {
  int B, N;
  double gtol, eps, c1, c2, amin, amax, xtol;
  int64_t maxiter;
  std::vector<double> x, g, pk, H;      // [B][N], [B][N], [B][N], [B][N][N]
  std::vector<Prob> pr;
  std::vector<int32_t> req;             // problems of the outstanding request
  std::vector<double> den;              // [nreq][N] denominators of the differences
  bool init = true;
  int64_t rounds = 0;
  double *X(int b) { return x.data() + (size_t)b * N; }
  double *G(int b) { return g.data() + (size_t)b * N; }
  double *PK(int b) { return pk.data() + (size_t)b * N; }
  double *HH(int b) { return H.data() + (size_t)b * N * N; }
};

double dot(const double *a, const double *b, int n) {
  double s = 0;
  for (int i = 0; i < n; i++) s = s + a[i] * b[i];
  return s;
}

void to_fallback(Bfgs &m, int b) {
  Prob &p = m.pr[b];
  Wolfe2 &w = p.fb;
  w.phi0 = p.fval; w.derphi0 = p.derphi0; w.c1 = m.c1; w.c2 = m.c2; w.amax = m.amax;
  w.alpha0 = 0;
  double alpha1;
  if (p.derphi0 != 0) alpha1 = pymin(1.0, 1.01 * 2 * (p.fval - p.old_old) / p.derphi0);
  else alpha1 = 1.0;
  if (alpha1 < 0) alpha1 = 1.0;
  alpha1 = pymin(alpha1, m.amax);
  w.alpha1 = alpha1;
  w.phi_a0 = p.fval;
  w.derphi_a0 = p.derphi0;
  w.i = 0;
  w.state = Wolfe2::BRACKET;
  p.stp = alpha1;
  p.mode = 2;
}

// direction and first step of a new line search
void begin_search(Bfgs &m, int b) {
  const int N = m.N;
  Prob &p = m.pr[b];
  double *pk = m.PK(b);
  const double *H = m.HH(b), *g = m.G(b);
  for (int i = 0; i < N; i++) pk[i] = -dot(H + (size_t)i * N, g, N);
  const double d0 = dot(g, pk, N);
  p.derphi0 = d0;
  double a1 = pymin(1.0, 1.01 * 2 * (p.fval - p.old_old) / d0);
  if (a1 < 0) a1 = 1.0;
  if (!(d0 != 0)) a1 = 1.0;
  p.stp = a1;
  p.ls_it = 1;
  // DCSRCH task START
  const bool err = a1 < m.amin || a1 > m.amax || d0 >= 0;
  Dcsrch &s = p.ls;
  s.brackt = false;
  s.stage = 1;
  s.finit = p.fval; s.ginit = d0;
  s.gtest = m.c1 * d0;
  s.width = m.amax - m.amin;
  s.width1 = (m.amax - m.amin) / P5;
  s.stx = 0.0; s.fx = p.fval; s.gx = d0;
  s.sty = 0.0; s.fy = p.fval; s.gy = d0;
  s.stmin = 0.0;
  s.stmax = a1 + XTRAPU * a1;
  if (err) to_fallback(m, b);
  else p.mode = 1;
}

// one call of DCSRCH._iterate with task FG: new step and task
Task dcsrch_iterate(Bfgs &m, Dcsrch &s, double &stp, double f, double g) {
  const double ftest = s.finit + stp * s.gtest;
  if (s.stage == 1 && f <= ftest && g >= 0) s.stage = 2;
  bool warn = s.brackt && (stp <= s.stmin || stp >= s.stmax);
  warn = warn || (s.brackt && s.stmax - s.stmin <= m.xtol * s.stmax);
  warn = warn || (stp == m.amax && f <= ftest && g <= s.gtest);
  warn = warn || (stp == m.amin && (f > ftest || g >= s.gtest));
  const bool conv = f <= ftest && fabs(g) <= m.c2 * -s.ginit;
  if (conv) return T_CONV;
  if (warn) return T_WARN;
  if (s.stage == 1 && f <= s.fx && f > ftest) {
    // the modified function of stage 1
    double fm = f - stp * s.gtest, fxm = s.fx - s.stx * s.gtest, fym = s.fy - s.sty * s.gtest;
    double gm = g - s.gtest, gxm = s.gx - s.gtest, gym = s.gy - s.gtest;
    dcstep(s.stx, fxm, gxm, s.sty, fym, gym, stp, fm, gm, s.brackt, s.stmin, s.stmax);
    s.fx = fxm + s.stx * s.gtest;
    s.fy = fym + s.sty * s.gtest;
    s.gx = gxm + s.gtest;
    s.gy = gym + s.gtest;
  } else {
    dcstep(s.stx, s.fx, s.gx, s.sty, s.fy, s.gy, stp, f, g, s.brackt, s.stmin, s.stmax);
  }
  if (s.brackt) {
    if (fabs(s.sty - s.stx) >= P66 * s.width1) stp = s.stx + P5 * (s.sty - s.stx);
    s.width1 = s.width;
    s.width = fabs(s.sty - s.stx);
  }
  if (s.brackt) {
    s.stmin = pymin(s.stx, s.sty);
    s.stmax = pymax(s.stx, s.sty);
  } else {
    s.stmin = stp + XTRAPL * (stp - s.stx);
    s.stmax = stp + XTRAPU * (stp - s.stx);
  }
  stp = stp < m.amin ? m.amin : (stp > m.amax ? m.amax : stp);   // np.clip
  if ((s.brackt && (stp <= s.stmin || stp >= s.stmax)) ||
      (s.brackt && s.stmax - s.stmin <= m.xtol * s.stmax))
    stp = s.stx;
  return T_FG;
}

// feed (phi, derphi) at the last step to the fallback search.  Returns 0: another step
// (p.stp set), 1: finished with alpha / phi_star, 2: failed.
int wolfe2_feed(Prob &p, double phi, double derphi, double *alpha, double *phi_star) {
  Wolfe2 &w = p.fb;
  if (w.state == Wolfe2::ZOOM) {
    const double a_j = w.a_j;
    if (phi > w.phi0 + w.c1 * a_j * w.derphi0 || phi >= w.phi_lo) {
      w.phi_rec = w.phi_hi; w.a_rec = w.a_hi; w.a_hi = a_j; w.phi_hi = phi;
    } else {
      if (fabs(derphi) <= -w.c2 * w.derphi0) { *alpha = a_j; *phi_star = phi; return 1; }
      if (derphi * (w.a_hi - w.a_lo) >= 0) {
        w.phi_rec = w.phi_hi; w.a_rec = w.a_hi; w.a_hi = w.a_lo; w.phi_hi = w.phi_lo;
      } else {
        w.phi_rec = w.phi_lo; w.a_rec = w.a_lo;
      }
      w.a_lo = a_j; w.phi_lo = phi; w.derphi_lo = derphi;
    }
    w.zi += 1;
    if (w.zi > 10) return 2;
    p.stp = zoom_next(w);
    return 0;
  }
  // bracketing phase: value at alpha1 received, loop body i
  const double phi_a1 = phi, derphi_a1 = derphi;
  if (w.i >= 10) {       // the loop is over: the last point, without the gradient test
    *alpha = w.alpha1; *phi_star = phi_a1;
    return 1;
  }
  if (w.alpha1 == 0 || w.alpha0 > w.amax) return 2;
  auto start_zoom = [&](double a_lo, double a_hi, double phi_lo, double phi_hi, double derphi_lo) {
    w.state = Wolfe2::ZOOM;
    w.zi = 0; w.phi_rec = w.phi0; w.a_rec = 0;
    w.a_lo = a_lo; w.a_hi = a_hi; w.phi_lo = phi_lo; w.phi_hi = phi_hi; w.derphi_lo = derphi_lo;
    p.stp = zoom_next(w);
  };
  if (phi_a1 > w.phi0 + w.c1 * w.alpha1 * w.derphi0 || (phi_a1 >= w.phi_a0 && w.i > 0)) {
    start_zoom(w.alpha0, w.alpha1, w.phi_a0, phi_a1, w.derphi_a0);
    return 0;
  }
  if (fabs(derphi_a1) <= -w.c2 * w.derphi0) { *alpha = w.alpha1; *phi_star = phi_a1; return 1; }
  if (derphi_a1 >= 0) {
    start_zoom(w.alpha1, w.alpha0, phi_a1, w.phi_a0, derphi_a1);
    return 0;
  }
  const double alpha2 = pymin(2 * w.alpha1, w.amax);
  w.alpha0 = w.alpha1; w.alpha1 = alpha2; w.phi_a0 = phi_a1; w.derphi_a0 = derphi_a1;
  w.i += 1;
  p.stp = w.alpha1;
  return 0;
}

// quasi-Newton update of a problem whose search ended at step alpha with (fnew, gnew)
void accept(Bfgs &m, int b, double alpha, double fnew, const double *gnew) {
  const int N = m.N;
  Prob &p = m.pr[b];
  double *x = m.X(b), *g = m.G(b), *pk = m.PK(b), *H = m.HH(b);
  double sk[64], yk[64];
  double pkmax = -INFINITY, gn = -INFINITY;
  bool gn_nan = false, pk_nan = false;
  for (int i = 0; i < N; i++) {
    sk[i] = alpha * pk[i];
    x[i] = x[i] + sk[i];
    yk[i] = gnew[i] - g[i];
    g[i] = gnew[i];
    const double ag = fabs(gnew[i]), ap = fabs(pk[i]);
    if (ag != ag) gn_nan = true; else if (ag > gn) gn = ag;
    if (ap != ap) pk_nan = true; else if (ap > pkmax) pkmax = ap;
  }
  if (gn_nan) gn = NAN;             // np.max propagates NaN
  if (pk_nan) pkmax = NAN;
  p.old_old = p.fval;
  p.fval = fnew;
  p.nit += 1;
  p.gnorm = gn;
  bool stop = gn <= m.gtol;
  stop = stop || (alpha * pkmax <= 0);      // xrtol = 0
  if (!stop && !isfinite(fnew)) { p.status = 2; stop = true; }
  if (stop) { p.mode = 0; return; }
  const double rinv = dot(yk, sk, N);
  const double rho = rinv == 0. ? 1000.0 : 1. / rinv;
  // Hn = A1 (H A2) + rho sk sk^T, A1 = I - sk yk^T rho, A2 = I - yk sk^T rho
  double A1[64 * 8], A2[64 * 8], T[64 * 8];   // N <= 22 -> N*N <= 512
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) {
      const double e = i == j ? 1.0 : 0.0;
      A1[i * N + j] = e - sk[i] * yk[j] * rho;
      A2[i * N + j] = e - yk[i] * sk[j] * rho;
    }
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) {
      double s = 0;
      for (int k = 0; k < N; k++) s = s + H[i * N + k] * A2[k * N + j];
      T[i * N + j] = s;
    }
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) {
      double s = 0;
      for (int k = 0; k < N; k++) s = s + A1[i * N + k] * T[k * N + j];
      H[i * N + j] = s + rho * sk[i] * sk[j];
    }
  if (gn > m.gtol && p.nit < m.maxiter) begin_search(m, b);
  else p.mode = 0;
}

}  // namespace

extern "C" void *rvs_bfgs_create(int B, int N, const double *h_x0, const double *h_hess_inv0,
                                 double gtol, int64_t maxiter) {
  if (B < 0 || N < 1 || N > 22 || !h_x0) return nullptr;
  Bfgs *m = new Bfgs;
  m->B = B; m->N = N;
  m->gtol = gtol;
  m->eps = 1.4901161193847656e-08;     // sqrt(finfo(float).eps), scipy's _epsilon
  m->c1 = 1e-4; m->c2 = 0.9;
  m->amin = 1e-100; m->amax = 1e100; m->xtol = 1e-14;
  m->maxiter = maxiter > 0 ? maxiter : (int64_t)N * 200;
  m->x.assign(h_x0, h_x0 + (size_t)B * N);
  m->g.assign((size_t)B * N, 0.0);
  m->pk.assign((size_t)B * N, 0.0);
  m->H.assign((size_t)B * N * N, 0.0);
  for (int b = 0; b < B; b++)
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++)
        m->HH(b)[i * N + j] = h_hess_inv0 ? h_hess_inv0[i * N + j] : (i == j ? 1.0 : 0.0);
  m->pr.resize(B);
  for (auto &p : m->pr) {
    memset(&p, 0, sizeof(p));
    p.mode = 1;
  }
  return m;
}

extern "C" void rvs_bfgs_destroy(void *h) { delete static_cast<Bfgs *>(h); }

// points of the request: x + stp pk (the start points in the first round) and their
// forward-difference neighbours (scipy's approx_derivative, 2-point, abs_step = eps)
extern "C" int64_t rvs_bfgs_request(void *h, int32_t *h_idx, double *h_X, int64_t cap) {
  Bfgs &m = *static_cast<Bfgs *>(h);
  const int N = m.N;
  m.req.clear();
  for (int b = 0; b < m.B; b++)
    if (m.pr[b].mode != 0) m.req.push_back(b);
  const int64_t n = (int64_t)m.req.size() * (N + 1);
  if (n > cap) return n;
  m.den.resize(m.req.size() * N);
  int64_t o = 0;
  for (size_t q = 0; q < m.req.size(); q++) {
    const int b = m.req[q];
    double pt[64];
    const double *x = m.X(b), *pk = m.PK(b);
    for (int i = 0; i < N; i++) pt[i] = m.init ? x[i] : x[i] + m.pr[b].stp * pk[i];
    for (int r = 0; r <= N; r++, o++) {
      h_idx[o] = b;
      memcpy(h_X + o * N, pt, sizeof(double) * N);
    }
    for (int i = 0; i < N; i++) {
      double hh = m.eps;
      const double dx = (pt[i] + hh) - pt[i];
      if (dx == 0) hh = m.eps * (pt[i] >= 0 ? 1.0 : -1.0) * (fabs(pt[i]) > 1.0 ? fabs(pt[i]) : 1.0);
      h_X[(o - N + i) * N + i] = pt[i] + hh;
      m.den[q * N + i] = (pt[i] + hh) - pt[i];
    }
  }
  return n;
}

extern "C" int rvs_bfgs_feed(void *h, const double *h_f, int64_t K) {
  Bfgs &m = *static_cast<Bfgs *>(h);
  const int N = m.N;
  if (K != (int64_t)m.req.size() * (N + 1)) return RVS_E_ARG;
  m.rounds += 1;
  for (size_t q = 0; q < m.req.size(); q++) {
    const int b = m.req[q];
    Prob &p = m.pr[b];
    const double *f = h_f + q * (N + 1);
    const double f1 = f[0];
    double g1[64];
    for (int i = 0; i < N; i++) g1[i] = (f[1 + i] - f1) / m.den[q * N + i];
    if (m.init) {
      double *g = m.G(b);
      memcpy(g, g1, sizeof(double) * N);
      p.fval = f1;
      p.old_old = f1 + sqrt(dot(g1, g1, N)) / 2;
      p.nit = 0;
      p.status = 0;
      double gn = -INFINITY;
      bool nan_seen = false;
      for (int i = 0; i < N; i++) {
        const double ag = fabs(g1[i]);
        if (ag != ag) nan_seen = true; else if (ag > gn) gn = ag;
      }
      p.gnorm = nan_seen ? NAN : gn;
      if (p.gnorm > m.gtol && p.nit < m.maxiter) begin_search(m, b);
      else p.mode = 0;
      continue;
    }
    const double d1 = dot(g1, m.PK(b), N);
    if (p.mode == 1) {
      double nstp = p.stp;
      const Task task = dcsrch_iterate(m, p.ls, nstp, f1, d1);
      p.ls_it += 1;
      if (task == T_CONV) {
        accept(m, b, p.stp, f1, g1);
      } else if (task == T_FG && isfinite(nstp) && p.ls_it < 100) {
        p.stp = nstp;
      } else {
        to_fallback(m, b);
      }
    } else {
      double alpha = 0, phi_star = 0;
      const int rc = wolfe2_feed(p, f1, d1, &alpha, &phi_star);
      if (rc == 1) accept(m, b, alpha, phi_star, g1);
      else if (rc == 2) { p.status = 2; p.mode = 0; }
    }
  }
  m.init = false;
  return 0;
}

extern "C" int64_t rvs_bfgs_live(void *h, uint8_t *h_active) {
  Bfgs &m = *static_cast<Bfgs *>(h);
  int64_t n = 0;
  for (int b = 0; b < m.B; b++) {
    const bool on = m.pr[b].mode != 0;
    n += on;
    if (h_active) h_active[b] = on;
  }
  return n;
}

extern "C" int rvs_bfgs_result(void *h, double *h_x, double *h_fun, int64_t *h_nit,
                               int32_t *h_status, int64_t *h_rounds) {
  Bfgs &m = *static_cast<Bfgs *>(h);
  const int N = m.N;
  for (int b = 0; b < m.B; b++) {
    const Prob &p = m.pr[b];
    if (h_x) memcpy(h_x + (size_t)b * N, m.X(b), sizeof(double) * N);
    if (h_fun) h_fun[b] = p.fval;
    if (h_nit) h_nit[b] = p.nit;
    if (h_status) {
      int st = 0;
      bool xnan = false;
      for (int i = 0; i < N; i++) xnan = xnan || m.X(b)[i] != m.X(b)[i];
      if (p.status == 2) st = 2;
      else if (p.nit >= m.maxiter) st = 1;
      else if (p.gnorm != p.gnorm || p.fval != p.fval || xnan) st = 3;
      h_status[b] = st;
    }
  }
  if (h_rounds) *h_rounds = m.rounds;
  return 0;
}

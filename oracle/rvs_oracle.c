/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99, scalar, fp64) of the three numeric kernels of the
 * rvspecfit likelihood hot path.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the
 * product (rvspecfit_b200/) never does.
 *
 * Parity status: pinned.  tests/test_oracle.py checks these functions against
 *   (1) oracle/_ref/libspliner_ref.so = the reference's own spliner.c compiled
 *       where it lies (oracle/Makefile), and
 *   (2) tests/golden/*.npz produced by running the reference package itself
 *       (tests/golden/make_golden.py).
 *
 * Each function cites the reference lines it restates.
 */
#include <math.h>
#include <stdlib.h>

/* Natural cubic spline through (x_i, y_i), i < n.
 * Restates reference py/rvspecfit/src/spliner.c:7-60.  Second derivatives z
 * solve  h_{i} z_{i} + 2 (h_i + h_{i+1}) z_{i+1} + h_{i+1} z_{i+2} = 6 (b_{i+1} - b_i),
 * z_0 = z_{n-1} = 0, with b_i = (y_{i+1}-y_i)/h_i, by the Thomas algorithm.
 * Output is the reference's representation on interval i:
 *   S(x) = A_i (x-x_i)^3 + B_i (x_{i+1}-x)^3 + C_i (x-x_i) + D_i (x_{i+1}-x). */
void orc_spline_construct(const double *x, const double *y, int n, double *A,
                          double *B, double *C, double *D, double *h) {
  const int m = n - 2; /* unknowns z_1 .. z_{n-2} */
  double *hinv = malloc(sizeof(double) * (n - 1));
  double *slope = malloc(sizeof(double) * (n - 1));
  double *cp = malloc(sizeof(double) * (m > 0 ? m : 1));
  double *dp = malloc(sizeof(double) * (m > 0 ? m : 1));
  double *z = calloc(n, sizeof(double));
  const double sixth = 1. / 6;
  for (int i = 0; i + 1 < n; i++) {
    h[i] = x[i + 1] - x[i];
    hinv[i] = 1. / h[i];
    slope[i] = (y[i + 1] - y[i]) * hinv[i];
  }
  if (m > 0) {
    /* row k (0-based) : sub h[k], diag 2(h[k]+h[k+1]), super h[k+1] */
    double diag = 2 * (h[1] + h[0]);
    cp[0] = h[1] / diag;
    dp[0] = 6 * (slope[1] - slope[0]) / diag;
    for (int k = 1; k < m; k++) {
      diag = 2 * (h[k + 1] + h[k]);
      const double rhs = 6 * (slope[k + 1] - slope[k]);
      const double den = diag - h[k] * cp[k - 1];
      cp[k] = h[k + 1] / den;
      dp[k] = (rhs - h[k] * dp[k - 1]) / den;
    }
    z[m] = dp[m - 1];
    for (int k = m - 1; k >= 1; k--) z[k] = dp[k - 1] - cp[k - 1] * z[k + 1];
  }
  for (int i = 0; i + 1 < n; i++) {
    const double t1 = hinv[i] * sixth, t2 = h[i] * sixth;
    A[i] = z[i + 1] * t1;
    B[i] = z[i] * t1;
    C[i] = y[i + 1] * hinv[i] - z[i + 1] * t2;
    D[i] = y[i] * hinv[i] - z[i] * t2;
  }
  free(hinv); free(slope); free(cp); free(dp); free(z);
}

/* Evaluate the spline on a knot grid that is uniform in x or in ln x.
 * Restates reference spliner.c:71-108, including its status codes:
 * -1 first/last evaluation point outside [x_0, x_{n-1}); -2 knots not uniform. */
int orc_spline_eval(const double *ex, int nex, int n, const double *x,
                    const double *A, const double *B, const double *C,
                    const double *D, int log_step, double *out) {
  const double x0 = x[0], xl = x[n - 1];
  if (ex[0] < x0 || ex[nex - 1] < x0) return -1;
  if (ex[0] >= xl || ex[nex - 1] >= xl) return -1;
  double step, off;
  if (log_step) {
    step = log(x[1] / x0);
    if (fabs(step - log(x[2] / x[1])) > 1e-10) return -2;
    off = log(x0);
  } else {
    step = x[1] - x0;
    if (fabs(step - (x[2] - x[1])) > 1e-10) return -2;
    off = x0;
  }
  for (int i = 0; i < nex; i++) {
    const double e = ex[i];
    const int p = (int)(((log_step ? log(e) : e) - off) / step);
    const double dl = e - x[p], dr = x[p + 1] - e;
    out[i] = A[p] * dl * dl * dl + B[p] * dr * dr * dr + C[p] * dl + D[p] * dr;
  }
  return 0;
}

/* Continuum-marginalised -2 log L for one arm, Cholesky form.
 * Restates reference spec_fit.py:205-249 (_get_chisq0_numba_chol_resid):
 *   Dv = spec/espec, G_i = polys_i * (templ/espec), v = G Dv, M = G G^T = L L^T,
 *   a = M^-1 v,  result = 2 sum ln L_ii + 2 sum ln espec + |Dv - a^T G|^2.
 * polys is (npoly, npix) row-major.  Returns 0, or 1 if M is not positive
 * definite / the result is not finite (the caller then takes the SVD route,
 * spec_fit.py:337-354).  coeffs (npoly) may be NULL. */
int orc_chisq0_chol(const double *spec, const double *templ, const double *polys,
                    const double *espec, int npix, int npoly, double *result,
                    double *coeffs) {
  double M[32 * 32], v[32], a[32];
  if (npoly > 32) return 2;
  for (int i = 0; i < npoly; i++) {
    v[i] = 0;
    for (int j = 0; j < npoly; j++) M[i * npoly + j] = 0;
  }
  double slog = 0;
  for (int p = 0; p < npix; p++) {
    const double d = spec[p] / espec[p], tn = templ[p] / espec[p];
    slog += log(espec[p]);
    double g[32];
    for (int i = 0; i < npoly; i++) g[i] = polys[(size_t)i * npix + p] * tn;
    for (int i = 0; i < npoly; i++) {
      v[i] += g[i] * d;
      for (int j = 0; j <= i; j++) M[i * npoly + j] += g[i] * g[j];
    }
  }
  /* in-place lower Cholesky */
  double ldet = 0;
  for (int j = 0; j < npoly; j++) {
    double s = M[j * npoly + j];
    for (int k = 0; k < j; k++) s -= M[j * npoly + k] * M[j * npoly + k];
    if (!(s > 0) || !isfinite(s)) return 1;
    const double ljj = sqrt(s);
    M[j * npoly + j] = ljj;
    ldet += 2.0 * log(ljj);
    for (int i = j + 1; i < npoly; i++) {
      double t = M[i * npoly + j];
      for (int k = 0; k < j; k++) t -= M[i * npoly + k] * M[j * npoly + k];
      M[i * npoly + j] = t / ljj;
    }
  }
  for (int i = 0; i < npoly; i++) { /* L y = v */
    double t = v[i];
    for (int k = 0; k < i; k++) t -= M[i * npoly + k] * a[k];
    a[i] = t / M[i * npoly + i];
  }
  for (int i = npoly - 1; i >= 0; i--) { /* L^T a = y */
    double t = a[i];
    for (int k = i + 1; k < npoly; k++) t -= M[k * npoly + i] * a[k];
    a[i] = t / M[i * npoly + i];
  }
  double rss = 0;
  for (int p = 0; p < npix; p++) {
    const double d = spec[p] / espec[p], tn = templ[p] / espec[p];
    double m = 0;
    for (int i = 0; i < npoly; i++) m += a[i] * (polys[(size_t)i * npix + p] * tn);
    rss += (d - m) * (d - m);
  }
  const double r = ldet + 2.0 * slog + rss;
  if (coeffs) for (int i = 0; i < npoly; i++) coeffs[i] = a[i];
  *result = r;
  return isfinite(r) ? 0 : 1;
}

/* qoc_quad.c -- TEST INFRASTRUCTURE ONLY: a ground truth for the oracle itself.
 *
 * exp(X) of a small complex matrix in IEEE binary128 (__float128, libquadmath-free: only + - * / are used) by scaling and
 * squaring of a Taylor series, and the Frechet derivative L(X, E) by a central difference of that exponential with a step
 * of 2^-40 (truncation error ~1e-24 relative, far below double precision).  SURVEY.md 8c lists a quad-precision
 * restatement in the oracle stack: the expm arithmetic of the reference lives in an un-vendored dependency
 * (ExponentialUtilities.jl, call site src/gradient_computations.jl:24), so the oracle's own Higham-2005 and
 * Al-Mohy-Higham restatements are pinned against this independent computation instead (tests/test_oracle.py).
 * Nothing here is ever on the product path.  Matrices are row-major, interleaved (re, im) doubles at the interface. */
#include <stdlib.h>
#include <string.h>

typedef __float128 q;

static void mm(int d, const q* ar, const q* ai, const q* br, const q* bi, q* cr, q* ci) {
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) {
      q sr = 0, si = 0;
      for (int k = 0; k < d; k++) {
        const q xr = ar[i * d + k], xi = ai[i * d + k], yr = br[k * d + j], yi = bi[k * d + j];
        sr += xr * yr - xi * yi;
        si += xr * yi + xi * yr;
      }
      cr[i * d + j] = sr;
      ci[i * d + j] = si;
    }
}

/* out = exp(X); X given as quad planes */
static void expm_q(int d, const q* xr, const q* xi, q* outr, q* outi) {
  const int n = d * d;
  /* 1-norm -> number of halvings so that ||X / 2^s||_1 <= 1/4 */
  q nrm = 0;
  for (int j = 0; j < d; j++) {
    q cs = 0;
    for (int i = 0; i < d; i++) {
      q a = xr[i * d + j], b = xi[i * d + j];
      if (a < 0) a = -a;
      if (b < 0) b = -b;
      cs += a + b;
    }
    if (cs > nrm) nrm = cs;
  }
  int s = 0;
  q scale = 1;
  while (nrm * scale > (q)0.25) { scale /= 2; s++; }
  q* ar = malloc(sizeof(q) * n * 6);
  q *ai = ar + n, *tr = ai + n, *ti = tr + n, *pr = ti + n, *pi = pr + n;
  for (int e = 0; e < n; e++) { ar[e] = xr[e] * scale; ai[e] = xi[e] * scale; }
  /* Taylor: sum_{k=0}^{40} A^k / k!  (0.25^40 / 40! ~ 1e-72) */
  for (int e = 0; e < n; e++) { outr[e] = 0; outi[e] = 0; tr[e] = 0; ti[e] = 0; }
  for (int i = 0; i < d; i++) { outr[i * d + i] = 1; tr[i * d + i] = 1; }
  for (int k = 1; k <= 40; k++) {
    mm(d, tr, ti, ar, ai, pr, pi);
    for (int e = 0; e < n; e++) { tr[e] = pr[e] / k; ti[e] = pi[e] / k; outr[e] += tr[e]; outi[e] += ti[e]; }
  }
  for (int t = 0; t < s; t++) {
    mm(d, outr, outi, outr, outi, pr, pi);
    memcpy(outr, pr, sizeof(q) * n);
    memcpy(outi, pi, sizeof(q) * n);
  }
  free(ar);
}

/* X, E: d x d row-major interleaved complex doubles.  U = exp(X) and (if E != NULL) L = d/dt exp(X + t E) at t = 0. */
int qoc_quad_expm(int d, const double* X, const double* E, double* U, double* L) {
  const int n = d * d;
  q* w = malloc(sizeof(q) * n * 8);
  if (!w) return 1;
  q *xr = w, *xi = xr + n, *ur = xi + n, *ui = ur + n, *yr = ui + n, *yi = yr + n, *vr = yi + n, *vi = vr + n;
  for (int e = 0; e < n; e++) { xr[e] = X[2 * e]; xi[e] = X[2 * e + 1]; }
  expm_q(d, xr, xi, ur, ui);
  for (int e = 0; e < n; e++) { U[2 * e] = (double)ur[e]; U[2 * e + 1] = (double)ui[e]; }
  if (E && L) {
    q h = 1;
    for (int t = 0; t < 40; t++) h /= 2;   /* 2^-40 */
    for (int e = 0; e < n; e++) { yr[e] = xr[e] + h * E[2 * e]; yi[e] = xi[e] + h * E[2 * e + 1]; }
    expm_q(d, yr, yi, ur, ui);
    for (int e = 0; e < n; e++) { yr[e] = xr[e] - h * E[2 * e]; yi[e] = xi[e] - h * E[2 * e + 1]; }
    expm_q(d, yr, yi, vr, vi);
    for (int e = 0; e < n; e++) { L[2 * e] = (double)((ur[e] - vr[e]) / (2 * h)); L[2 * e + 1] = (double)((ui[e] - vi[e]) / (2 * h)); }
  }
  free(w);
  return 0;
}

/* Serial propagation x <- U_k x, k = 0..nt-1, accumulated in binary128 (the U_k themselves are given in double):
 * the ground truth against which the REASSOCIATION error of the parallel scans (segment products, two-level boundary
 * walk, time sharding) is bounded -- SURVEY.md F7.  U: nt matrices, d x d row-major interleaved complex doubles;
 * x0, x_out: d x m row-major interleaved complex doubles. */
int qoc_quad_chain(int d, int m, int nt, const double* U, const double* x0, double* x_out) {
  const int n = d * m;
  q* w = malloc(sizeof(q) * n * 4);
  if (!w) return 1;
  q *xr = w, *xi = xr + n, *yr = xi + n, *yi = yr + n;
  for (int e = 0; e < n; e++) { xr[e] = x0[2 * e]; xi[e] = x0[2 * e + 1]; }
  for (int k = 0; k < nt; k++) {
    const double* Uk = U + (size_t)k * 2 * d * d;
    for (int i = 0; i < d; i++)
      for (int c = 0; c < m; c++) {
        q sr = 0, si = 0;
        for (int j = 0; j < d; j++) {
          const q ur = Uk[2 * (i * d + j)], ui = Uk[2 * (i * d + j) + 1];
          sr += ur * xr[j * m + c] - ui * xi[j * m + c];
          si += ur * xi[j * m + c] + ui * xr[j * m + c];
        }
        yr[i * m + c] = sr;
        yi[i * m + c] = si;
      }
    q* t = xr; xr = yr; yr = t;
    t = xi; xi = yi; yi = t;
  }
  for (int e = 0; e < n; e++) { x_out[2 * e] = (double)xr[e]; x_out[2 * e + 1] = (double)xi[e]; }
  free(w);
  return 0;
}

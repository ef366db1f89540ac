/*
 * qoc_ref.c -- plain-C restatement of the reference's exp-based GRAPE path (CPU).
 *
 * TEST INFRASTRUCTURE ONLY: linked/executed only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  The product never calls it.  The Julia reference cannot run in this image (no julia),
 * so this file is the timed CPU baseline ("port") and a second oracle, cross-checked against oracle/qoc_oracle.py.
 * Parity pinning: see the header of qoc_oracle.py (expm boundary itself is "parity unpinned": the arithmetic lives
 * in the un-vendored ExponentialUtilities.jl; pinned indirectly through the reference's known answers).
 *
 * It mirrors the reference STRUCTURALLY so that the timing means what the reference's would:
 *   - per-slice generator + expm in an OpenMP parallel-for  (Threads.@threads, src/gradient_computations.jl:17-25)
 *   - serial forward sweep                                  (:27-29)
 *   - serial costate sweep                                  (:52-58)
 *   - serial Jacobian + contraction loop                    (:65-74, expm_jacobian! :177-213, :217-223)
 * Matrices are column-major double _Complex (== Julia Matrix{ComplexF64}); A0, A[j] are pre-multiplied by dt.
 * order 1..4 = the reference's truncated Taylor Jacobian; order 0 = exact Frechet derivative by the structured
 * block-triangular Pade evaluation (Al-Mohy & Higham 2009), which the reference does not have (it is here so the
 * CPU baseline can do the same work as the CUDA exact mode).  The small dense products are hand-written loops
 * (no system BLAS dev package in the image); a tuned BLAS would be faster at large d.
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC -o _build/libqoc_ref.so qoc_ref.c -lm
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double _Complex cplx;

static const double PADE3[] = {120., 60., 12., 1.};
static const double PADE5[] = {30240., 15120., 3360., 420., 30., 1.};
static const double PADE7[] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
static const double PADE9[] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
static const double PADE13[] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                                129060195264000.,   10559470521600.,    670442572800.,    33522128640.,
                                1323241920.,        40840800.,          960960.,          16380., 182., 1.};
/* degree switch points: Julia exp!/ExponentialUtilities (rounded Higham-2005 theta) and Al-Mohy-Higham l_m */
static const double THETA[] = {0.015, 0.25, 0.95, 2.1, 5.4};
static const double ELL[] = {1.08e-2, 2.00e-1, 7.83e-1, 1.78, 4.74};

/* C = A*B (beta=0) or C += alpha*A*B, column-major d x d */
static void gemm(int d, cplx alpha, const cplx* restrict A, const cplx* restrict B, double beta, cplx* restrict C) {
  for (int j = 0; j < d; j++) {
    cplx* restrict c = C + (size_t)j * d;
    if (beta == 0.0) for (int i = 0; i < d; i++) c[i] = 0.0;
    for (int k = 0; k < d; k++) {
      const cplx b = alpha * B[k + (size_t)j * d];
      const cplx* restrict a = A + (size_t)k * d;
      for (int i = 0; i < d; i++) c[i] += a[i] * b;
    }
  }
}
static void axpy(int n, double a, const cplx* restrict x, cplx* restrict y) { for (int i = 0; i < n; i++) y[i] += a * x[i]; }
static void addI(int d, double a, cplx* X) { for (int i = 0; i < d; i++) X[i + (size_t)i * d] += a; }
static double norm1(int d, const cplx* A) {
  double mx = 0;
  for (int j = 0; j < d; j++) { double s = 0; for (int i = 0; i < d; i++) s += cabs(A[i + (size_t)j * d]); if (s > mx) mx = s; }
  return mx;
}
/* LU with partial pivoting (getrf) and solve for nrhs right-hand sides (getrs), in place */
static int lu_factor(int d, cplx* A, int* piv) {
  for (int k = 0; k < d; k++) {
    int p = k; double mx = cabs(A[k + (size_t)k * d]);
    for (int i = k + 1; i < d; i++) { double v = cabs(A[i + (size_t)k * d]); if (v > mx) { mx = v; p = i; } }
    piv[k] = p;
    if (mx == 0.0) return 1;
    if (p != k) for (int j = 0; j < d; j++) { cplx t = A[k + (size_t)j * d]; A[k + (size_t)j * d] = A[p + (size_t)j * d]; A[p + (size_t)j * d] = t; }
    const cplx inv = 1.0 / A[k + (size_t)k * d];
    for (int i = k + 1; i < d; i++) A[i + (size_t)k * d] *= inv;
    for (int j = k + 1; j < d; j++) {
      const cplx a = A[k + (size_t)j * d];
      cplx* restrict c = A + (size_t)j * d; const cplx* restrict l = A + (size_t)k * d;
      for (int i = k + 1; i < d; i++) c[i] -= l[i] * a;
    }
  }
  return 0;
}
static void lu_solve(int d, const cplx* LU, const int* piv, int nrhs, cplx* B) {
  for (int j = 0; j < nrhs; j++) {
    cplx* b = B + (size_t)j * d;
    for (int k = 0; k < d; k++) { int p = piv[k]; if (p != k) { cplx t = b[k]; b[k] = b[p]; b[p] = t; } }
    for (int k = 0; k < d; k++) { const cplx v = b[k]; const cplx* l = LU + (size_t)k * d; for (int i = k + 1; i < d; i++) b[i] -= l[i] * v; }
    for (int k = d - 1; k >= 0; k--) { b[k] /= LU[k + (size_t)k * d]; const cplx v = b[k]; const cplx* l = LU + (size_t)k * d; for (int i = 0; i < k; i++) b[i] -= l[i] * v; }
  }
}

static void select_degree(double nrm, const double* tab, int* q, int* s) {
  static const int Q[] = {3, 5, 7, 9};
  for (int i = 0; i < 4; i++) if (nrm <= tab[i]) { *q = Q[i]; *s = 0; return; }
  *q = 13; *s = 0;
  if (nrm > tab[4]) { int v = (int)ceil(log2(nrm / tab[4])); *s = v > 0 ? v : 0; }
}
static const double* pade_coef(int q) { return q == 3 ? PADE3 : q == 5 ? PADE5 : q == 7 ? PADE7 : q == 9 ? PADE9 : PADE13; }

/* workspace: NW matrices of d*d */
#define NWORK 24

/* Higham-2005 scaling & squaring expm (what exponential!(.., ExpMethodHigham2005()) computes; no balancing).
 * X is overwritten; result in R.  w: >= 6 matrices. */
static int expm_h05(int d, cplx* X, cplx* R, cplx* w, int* piv, int* q_out, int* s_out) {
  const size_t dd = (size_t)d * d;
  cplx *A2 = w, *P = w + dd, *U = w + 2 * dd, *V = w + 3 * dd, *T = w + 4 * dd, *A4 = w + 5 * dd;
  int q, s;
  select_degree(norm1(d, X), THETA, &q, &s);
  if (q_out) *q_out = q;
  if (s_out) *s_out = s;
  if (q < 13) {
    const double* C = pade_coef(q);
    gemm(d, 1.0, X, X, 0.0, A2);
    memcpy(P, A2, dd * sizeof(cplx));
    for (size_t i = 0; i < dd; i++) { U[i] = C[3] * P[i]; V[i] = C[2] * P[i]; }
    addI(d, C[1], U); addI(d, C[0], V);
    for (int k = 2; k < (q + 1) / 2; k++) {
      gemm(d, 1.0, P, A2, 0.0, T); memcpy(P, T, dd * sizeof(cplx));
      axpy((int)dd, C[2 * k + 1], P, U); axpy((int)dd, C[2 * k], P, V);
    }
    gemm(d, 1.0, X, U, 0.0, T);            /* U = A*U */
    for (size_t i = 0; i < dd; i++) { R[i] = V[i] + T[i]; V[i] = V[i] - T[i]; }
    if (lu_factor(d, V, piv)) return 1;
    lu_solve(d, V, piv, d, R);
    return 0;
  }
  if (s > 0) { const double sc = ldexp(1.0, -s); for (size_t i = 0; i < dd; i++) X[i] *= sc; }
  const double* C = PADE13;
  cplx* A6 = P;
  gemm(d, 1.0, X, X, 0.0, A2); gemm(d, 1.0, A2, A2, 0.0, A4); gemm(d, 1.0, A2, A4, 0.0, A6);
  for (size_t i = 0; i < dd; i++) T[i] = C[13] * A6[i] + C[11] * A4[i] + C[9] * A2[i];
  gemm(d, 1.0, A6, T, 0.0, U);
  for (size_t i = 0; i < dd; i++) U[i] += C[7] * A6[i] + C[5] * A4[i] + C[3] * A2[i];
  addI(d, C[1], U);
  for (size_t i = 0; i < dd; i++) T[i] = C[12] * A6[i] + C[10] * A4[i] + C[8] * A2[i];
  gemm(d, 1.0, A6, T, 0.0, V);
  for (size_t i = 0; i < dd; i++) V[i] += C[6] * A6[i] + C[4] * A4[i] + C[2] * A2[i];
  addI(d, C[0], V);
  gemm(d, 1.0, X, U, 0.0, T);
  for (size_t i = 0; i < dd; i++) { R[i] = V[i] + T[i]; V[i] = V[i] - T[i]; }
  if (lu_factor(d, V, piv)) return 1;
  lu_solve(d, V, piv, d, R);
  for (int t = 0; t < s; t++) { gemm(d, 1.0, R, R, 0.0, T); memcpy(R, T, dd * sizeof(cplx)); }
  return 0;
}

/* exact Frechet derivatives L(X, E_j), j < nc, plus R = exp(X): Al-Mohy & Higham 2009 Alg 6.4 restated with the
 * A-only part shared between the controls.  w: >= NWORK matrices. */
static int expm_frechet(int d, int nc, const cplx* X0, const cplx* const* Ein, cplx* R, cplx* L, cplx* w, int* piv,
                        int* q_out, int* s_out) {
  const size_t dd = (size_t)d * d;
  cplx *A = w, *A2 = w + dd, *A4 = w + 2 * dd, *A6 = w + 3 * dd, *A8 = w + 4 * dd, *W = w + 5 * dd, *U = w + 6 * dd,
       *V = w + 7 * dd, *N = w + 8 * dd, *E = w + 9 * dd, *M2 = w + 10 * dd, *M4 = w + 11 * dd, *M6 = w + 12 * dd,
       *M8 = w + 13 * dd, *Lw = w + 14 * dd, *Lu = w + 15 * dd, *Lv = w + 16 * dd, *T = w + 17 * dd, *W1 = w + 18 * dd,
       *Z1 = w + 19 * dd, *T2 = w + 20 * dd;
  int q, s;
  select_degree(norm1(d, X0), ELL, &q, &s);
  if (q_out) *q_out = q;
  if (s_out) *s_out = s;
  const double sc = ldexp(1.0, -s);
  for (size_t i = 0; i < dd; i++) A[i] = X0[i] * sc;
  const double* b = pade_coef(q);
  const int np = (q - 1) / 2; /* number of even powers used when q < 13 */
  cplx* Ap[5] = {0, A2, A4, A6, A8};
  cplx* Mp[5] = {0, M2, M4, M6, M8};
  gemm(d, 1.0, A, A, 0.0, A2);
  if (q < 13) {
    for (int k = 2; k <= np; k++) gemm(d, 1.0, Ap[k - 1], A2, 0.0, Ap[k]);
    memset(W, 0, dd * sizeof(cplx)); memset(V, 0, dd * sizeof(cplx));
    addI(d, b[1], W); addI(d, b[0], V);
    for (int k = 1; k <= np; k++) { axpy((int)dd, b[2 * k + 1], Ap[k], W); axpy((int)dd, b[2 * k], Ap[k], V); }
    gemm(d, 1.0, A, W, 0.0, U);
  } else {
    gemm(d, 1.0, A2, A2, 0.0, A4); gemm(d, 1.0, A2, A4, 0.0, A6);
    for (size_t i = 0; i < dd; i++) { W1[i] = b[13] * A6[i] + b[11] * A4[i] + b[9] * A2[i]; Z1[i] = b[12] * A6[i] + b[10] * A4[i] + b[8] * A2[i]; }
    gemm(d, 1.0, A6, W1, 0.0, W);
    for (size_t i = 0; i < dd; i++) W[i] += b[7] * A6[i] + b[5] * A4[i] + b[3] * A2[i];
    addI(d, b[1], W);
    gemm(d, 1.0, A, W, 0.0, U);
    gemm(d, 1.0, A6, Z1, 0.0, V);
    for (size_t i = 0; i < dd; i++) V[i] += b[6] * A6[i] + b[4] * A4[i] + b[2] * A2[i];
    addI(d, b[0], V);
  }
  for (size_t i = 0; i < dd; i++) { N[i] = V[i] - U[i]; R[i] = V[i] + U[i]; }
  if (lu_factor(d, N, piv)) return 1;
  lu_solve(d, N, piv, d, R);
  for (int j = 0; j < nc; j++) {
    cplx* Lj = L + (size_t)j * dd;
    for (size_t i = 0; i < dd; i++) E[i] = Ein[j][i] * sc;
    gemm(d, 1.0, A, E, 0.0, M2); gemm(d, 1.0, E, A, 1.0, M2);
    if (q < 13) {
      for (int k = 2; k <= np; k++) { gemm(d, 1.0, Mp[k - 1], A2, 0.0, Mp[k]); gemm(d, 1.0, Ap[k - 1], M2, 1.0, Mp[k]); }
      memset(Lw, 0, dd * sizeof(cplx)); memset(Lv, 0, dd * sizeof(cplx));
      for (int k = 1; k <= np; k++) { axpy((int)dd, b[2 * k + 1], Mp[k], Lw); axpy((int)dd, b[2 * k], Mp[k], Lv); }
      gemm(d, 1.0, A, Lw, 0.0, Lu); gemm(d, 1.0, E, W, 1.0, Lu);
    } else {
      gemm(d, 1.0, A2, M2, 0.0, M4); gemm(d, 1.0, M2, A2, 1.0, M4);
      gemm(d, 1.0, A4, M2, 0.0, M6); gemm(d, 1.0, M4, A2, 1.0, M6);
      for (size_t i = 0; i < dd; i++) { T[i] = b[13] * M6[i] + b[11] * M4[i] + b[9] * M2[i]; T2[i] = b[12] * M6[i] + b[10] * M4[i] + b[8] * M2[i]; }
      gemm(d, 1.0, A6, T, 0.0, Lw); gemm(d, 1.0, M6, W1, 1.0, Lw);
      for (size_t i = 0; i < dd; i++) Lw[i] += b[7] * M6[i] + b[5] * M4[i] + b[3] * M2[i];
      gemm(d, 1.0, A6, T2, 0.0, Lv); gemm(d, 1.0, M6, Z1, 1.0, Lv);
      for (size_t i = 0; i < dd; i++) Lv[i] += b[6] * M6[i] + b[4] * M4[i] + b[2] * M2[i];
      gemm(d, 1.0, A, Lw, 0.0, Lu); gemm(d, 1.0, E, W, 1.0, Lu);
    }
    for (size_t i = 0; i < dd; i++) { T[i] = Lu[i] - Lv[i]; Lj[i] = Lu[i] + Lv[i]; }
    gemm(d, 1.0, T, R, 1.0, Lj);
    lu_solve(d, N, piv, d, Lj);
  }
  for (int t = 0; t < s; t++) {
    for (int j = 0; j < nc; j++) {
      cplx* Lj = L + (size_t)j * dd;
      gemm(d, 1.0, R, Lj, 0.0, T); gemm(d, 1.0, Lj, R, 1.0, T);
      memcpy(Lj, T, dd * sizeof(cplx));
    }
    gemm(d, 1.0, R, R, 0.0, T); memcpy(R, T, dd * sizeof(cplx));
  }
  return 0;
}

/* expm_jacobian!(dFdp, A0, A, p, tmp, order)  src/gradient_computations.jl:177-213, dt = 1, same association */
static void expm_jacobian(int d, int nc, const cplx* A0, const cplx* A, const double* p, int order, cplx* dF, cplx* w) {
  const size_t dd = (size_t)d * d;
  cplx *X = w, *AjX = w + dd, *XAj = w + 2 * dd, *X2 = w + 3 * dd;
  for (int j = 0; j < nc; j++) memcpy(dF + (size_t)j * dd, A + (size_t)j * dd, dd * sizeof(cplx));
  if (order <= 1) return;
  memcpy(X, A0, dd * sizeof(cplx));
  for (int j = 0; j < nc; j++) axpy((int)dd, p[j], A + (size_t)j * dd, X);
  for (int j = 0; j < nc; j++) {
    const cplx* Aj = A + (size_t)j * dd; cplx* out = dF + (size_t)j * dd;
    gemm(d, 1.0, Aj, X, 0.0, AjX); gemm(d, 1.0, X, Aj, 0.0, XAj);
    for (size_t i = 0; i < dd; i++) out[i] += 0.5 * (AjX[i] + XAj[i]);
    if (order >= 3) { gemm(d, 1.0 / 6.0, AjX, X, 1.0, out); gemm(d, 1.0 / 6.0, XAj, X, 1.0, out); gemm(d, 1.0 / 6.0, X, XAj, 1.0, out); }
    if (order >= 4) {
      gemm(d, 1.0, X, X, 0.0, X2); /* recomputed per control, as the reference does at :206 */
      gemm(d, 1.0 / 24.0, AjX, X2, 1.0, out); gemm(d, 1.0 / 24.0, XAj, X2, 1.0, out);
      gemm(d, 1.0 / 24.0, X2, AjX, 1.0, out); gemm(d, 1.0 / 24.0, X2, XAj, 1.0, out);
    }
  }
}

/* y (d x m) = U x   or  U' x */
static void apply(int d, int m, const cplx* U, const cplx* x, cplx* y, int adj) {
  for (int l = 0; l < m; l++) {
    const cplx* xl = x + (size_t)l * d; cplx* yl = y + (size_t)l * d;
    if (!adj) {
      for (int i = 0; i < d; i++) yl[i] = 0.0;
      for (int k = 0; k < d; k++) { const cplx v = xl[k]; const cplx* u = U + (size_t)k * d; for (int i = 0; i < d; i++) yl[i] += u[i] * v; }
    } else {
      for (int i = 0; i < d; i++) { const cplx* u = U + (size_t)i * d; cplx s = 0.0; for (int k = 0; k < d; k++) s += conj(u[k]) * xl[k]; yl[i] = s; }
    }
  }
}

/*
 * One full fidelity + gradient evaluation of ONE pulse, laid out like the reference's f followed by f_grad.
 * cost: 0 = 1-|tr(T'x)|^2/n^2, 1 = 1-|tr(T'x)|.  pen_*: running state penalty (n_pen_rows = 0 disables).
 * Outputs (any may be NULL): J, dJdu (nc x nt), Uk (d*d*nt), xs (d*m*(nt+1)), lams (d*m*(nt+1)), dU (d*d*nc*nt).
 * flops_out: algorithmic flops with the degrees actually chosen.  Returns 0 on success.
 */
int qoc_ref_eval(int d, int m, int nc, int nt, const double* A0_, const double* A_, const double* u, const double* x0_,
                 const double* T_, int cost, int n, int order, int nthreads, int want_grad, int n_pen_rows,
                 int n_pen_cols, const int* pen_rows, const int* pen_cols, double mu, double* J_out, double* dJdu,
                 double* Uk_out, double* xs_out, double* lams_out, double* dU_out, double* flops_out) {
  const cplx* A0 = (const cplx*)A0_; const cplx* A = (const cplx*)A_; const cplx* x0 = (const cplx*)x0_; const cplx* T = (const cplx*)T_;
  const size_t dd = (size_t)d * d, dm = (size_t)d * m;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  cplx* Uk = (cplx*)malloc(sizeof(cplx) * dd * nt);
  cplx* Lk = (order == 0 && want_grad) ? (cplx*)malloc(sizeof(cplx) * dd * nc * nt) : NULL;
  cplx* x = (cplx*)malloc(sizeof(cplx) * dm * (nt + 1));
  cplx* lam = (cplx*)malloc(sizeof(cplx) * dm * (nt + 1));
  cplx* work = (cplx*)malloc(sizeof(cplx) * dd * NWORK * nthreads);
  int* piv = (int*)malloc(sizeof(int) * d * nthreads);
  double* fl = (double*)calloc(nthreads, sizeof(double));
  int fail = 0;
  const double M = 8.0 * d * d * (double)d;
  const cplx** Ep = (const cplx**)malloc(sizeof(cplx*) * nc);
  for (int j = 0; j < nc; j++) Ep[j] = A + (size_t)j * dd;

  /* ---- propagate: threaded expm loop (:17-25) ---- */
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int k = 0; k < nt; k++) {
#ifdef _OPENMP
    const int tid = omp_get_thread_num();
#else
    const int tid = 0;
#endif
    cplx* w = work + (size_t)tid * NWORK * dd;
    cplx* X = w + (size_t)(NWORK - 1) * dd;
    memcpy(X, A0, dd * sizeof(cplx));
    for (int j = 0; j < nc; j++) axpy((int)dd, u[j + (size_t)nc * k], A + (size_t)j * dd, X);
    int q = 13, s = 0, rc;
    if (order == 0 && want_grad) {
      /* the exact mode produces U_k and L_kj together (they share the Pade powers), inside the threaded loop */
      rc = expm_frechet(d, nc, X, Ep, Uk + (size_t)k * dd, Lk + (size_t)k * nc * dd, w, piv + (size_t)tid * d, &q, &s);
      const int pi = q == 3 ? 2 : q == 5 ? 3 : q == 7 ? 4 : q == 9 ? 5 : 6;
      fl[tid] += M * ((pi + s + 4.0 / 3.0) + nc * (2.0 * pi + 2.0 * s + 2.0));
    } else {
      rc = expm_h05(d, X, Uk + (size_t)k * dd, w, piv + (size_t)tid * d, &q, &s);
      const int pi = q == 3 ? 2 : q == 5 ? 3 : q == 7 ? 4 : q == 9 ? 5 : 6;
      fl[tid] += M * (pi + s + 4.0 / 3.0);
    }
    if (rc) {
#pragma omp atomic write
      fail = 1;
    }
  }
  /* ---- serial forward sweep (:27-29) ---- */
  memcpy(x, x0, dm * sizeof(cplx));
  for (int k = 0; k < nt; k++) apply(d, m, Uk + (size_t)k * dd, x + (size_t)k * dm, x + (size_t)(k + 1) * dm, 0);
  /* ---- cost: Jfinal(x[end]) + sum(L, x)  (examples/ipopt_callbacks_exp.jl:18) ---- */
  cplx om = 0.0;
  const cplx* xN = x + (size_t)nt * dm;
  for (size_t i = 0; i < dm; i++) om += conj(T[i]) * xN[i];
  double J;
  cplx coef;
  if (cost == 0) { J = 1.0 - (creal(om) * creal(om) + cimag(om) * cimag(om)) / ((double)n * n); coef = -2.0 * om / ((double)n * n); }
  else { J = 1.0 - cabs(om); coef = -om / cabs(om); }
  if (n_pen_rows > 0)
    for (int k = 0; k <= nt; k++)
      for (int c = 0; c < n_pen_cols; c++)
        for (int r = 0; r < n_pen_rows; r++) { const cplx v = x[(size_t)k * dm + pen_rows[r] + (size_t)d * pen_cols[c]]; J += mu * (creal(v) * creal(v) + cimag(v) * cimag(v)); }
  if (J_out) *J_out = J;
  double flops = 0;
  for (int t = 0; t < nthreads; t++) flops += fl[t];
  flops += 8.0 * d * d * m * (double)nt;

  if (want_grad) {
    /* ---- serial costate sweep (:46-58) ---- */
    cplx* lN = lam + (size_t)nt * dm;
    for (size_t i = 0; i < dm; i++) lN[i] = coef * T[i];
#define ADD_PEN(k)                                                                                      \
  if (n_pen_rows > 0)                                                                                   \
    for (int c = 0; c < n_pen_cols; c++)                                                                \
      for (int r = 0; r < n_pen_rows; r++) {                                                            \
        const size_t ix = (size_t)(k) * dm + pen_rows[r] + (size_t)d * pen_cols[c];                    \
        lam[ix] += 2.0 * mu * x[ix];                                                                    \
      }
    ADD_PEN(nt)
    for (int k = nt - 1; k >= 0; k--) {
      apply(d, m, Uk + (size_t)k * dd, lam + (size_t)(k + 1) * dm, lam + (size_t)k * dm, 1);
      ADD_PEN(k)
    }
    /* ---- serial Jacobian + contraction loop (:61-74) ---- */
    cplx* dF = (cplx*)malloc(sizeof(cplx) * dd * nc);
    cplx* y = (cplx*)malloc(sizeof(cplx) * dm);
    for (int k = nt - 1; k >= 0; k--) {
      const cplx* dUk;
      if (order == 0) dUk = Lk + (size_t)k * nc * dd;
      else { expm_jacobian(d, nc, A0, A, u + (size_t)nc * k, order, dF, work); dUk = dF; }
      for (int j = 0; j < nc; j++) {
        apply(d, m, dUk + (size_t)j * dd, x + (size_t)k * dm, y, 0);
        const cplx* l = lam + (size_t)(k + 1) * dm;
        double s = 0;
        for (size_t i = 0; i < dm; i++) s += creal(l[i]) * creal(y[i]) + cimag(l[i]) * cimag(y[i]);
        dJdu[j + (size_t)nc * k] = s;
      }
      if (dU_out) memcpy(dU_out + 2 * (size_t)k * nc * dd, dUk, sizeof(cplx) * nc * dd);
    }
    free(dF); free(y);
    const double G = order == 0 ? 0.0 : order == 1 ? 0.0 : order == 2 ? 2.0 : order == 3 ? 5.0 : 10.0;
    flops += (M * nc * G + 8.0 * d * d * m * (1.0 + nc) + 4.0 * nc * d * d) * (double)nt;
  }
  if (flops_out) *flops_out = flops;
  if (Uk_out) memcpy(Uk_out, Uk, sizeof(cplx) * dd * nt);
  if (xs_out) memcpy(xs_out, x, sizeof(cplx) * dm * (nt + 1));
  if (lams_out && want_grad) memcpy(lams_out, lam, sizeof(cplx) * dm * (nt + 1));
  free(Uk); if (Lk) free(Lk); free(x); free(lam); free(work); free(piv); free(fl); free(Ep);
  return fail;
}

int qoc_ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* tests/abi_driver.c -- a plain C program against include/qoc_b200.h: what a host without Python (the Julia ccall shim of
 * INTEGRATION.md, a C++ optimiser) does.  Built and run by tests/test_gpu_abi_driver.py (-m gpu):
 *
 *     abi_driver <dir> <n_ranks>
 *
 * <dir> holds raw little-endian files written by the test from the oracle's zz_coupling case (tests/golden/zz_order3.npz):
 *   dims.txt (d m nc nt order cost n), A0.bin, A.bin, x0.bin, T.bin (c128 column-major), u.bin (nc x nt doubles, j fastest),
 *   J.bin (1 double), g.bin (nc x nt doubles), lam.bin (c128 d x m = dJfinal_dx(x_N) of the oracle)
 * The program mirrors the reference's call sequence (examples/ipopt_callbacks_exp.jl:11-31):
 *   setup_grape_cache -> propagate -> grape_sensitivity(dJfinal_dx as a host closure = lambda_final) -> the built-in cost ->
 *   the fused qoc_eval -> the stale-u error (src/gradient_computations.jl:37-39) -> qoc_create_sharded on n_ranks ranks
 * and prints one line per check; exit code 0 iff everything is within the north-star tolerances.                          */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qoc_b200.h"

static double* slurp(const char* dir, const char* name, size_t n) {
  char path[1024];
  snprintf(path, sizeof path, "%s/%s", dir, name);
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  double* p = (double*)malloc(n * sizeof(double));
  if (fread(p, sizeof(double), n, f) != n) { fprintf(stderr, "short read %s\n", path); exit(2); }
  fclose(f);
  return p;
}
static double maxabsdiff(const double* a, const double* b, size_t n) {
  double m = 0;
  for (size_t i = 0; i < n; i++) { double e = fabs(a[i] - b[i]); if (e > m) m = e; }
  return m;
}
static double maxabs(const double* a, size_t n) {
  double m = 0;
  for (size_t i = 0; i < n; i++) if (fabs(a[i]) > m) m = fabs(a[i]);
  return m;
}
#define CHECK(cond, ...) do { printf(__VA_ARGS__); printf(" : %s\n", (cond) ? "ok" : "FAILED"); if (!(cond)) bad++; } while (0)

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: abi_driver <dir> <n_ranks>\n"); return 2; }
  const char* dir = argv[1];
  const int n_ranks = atoi(argv[2]);
  int bad = 0, d, m, nc, nt, order, cost, n;
  {
    char path[1024];
    snprintf(path, sizeof path, "%s/dims.txt", dir);
    FILE* f = fopen(path, "r");
    if (!f || fscanf(f, "%d %d %d %d %d %d %d", &d, &m, &nc, &nt, &order, &cost, &n) != 7) { fprintf(stderr, "bad dims.txt\n"); return 2; }
    fclose(f);
  }
  double* A0 = slurp(dir, "A0.bin", 2 * (size_t)d * d);
  double* A = slurp(dir, "A.bin", 2 * (size_t)d * d * nc);
  double* x0 = slurp(dir, "x0.bin", 2 * (size_t)d * m);
  double* T = slurp(dir, "T.bin", 2 * (size_t)d * m);
  double* u = slurp(dir, "u.bin", (size_t)nc * nt);
  double* Jref = slurp(dir, "J.bin", 1);
  double* gref = slurp(dir, "g.bin", (size_t)nc * nt);
  double* lam = slurp(dir, "lam.bin", 2 * (size_t)d * m);
  const double gmax = maxabs(gref, (size_t)nc * nt);

  qoc_problem pr;
  memset(&pr, 0, sizeof pr);
  pr.d = d; pr.m = m; pr.nc = nc; pr.nt = nt; pr.batch = 1; pr.order = order; pr.cost = QOC_COST_NONE; pr.n = n; pr.device = 0;
  qoc_handle* h = NULL;
  int rc = qoc_create(&pr, A0, A, x0, NULL, &h);                       /* setup_grape_cache */
  CHECK(rc == QOC_OK, "qoc_create rc=%d (%s)", rc, rc ? qoc_last_error(NULL) : "");
  if (rc) return 1;
  double* xf = (double*)calloc(2 * (size_t)d * m, sizeof(double));
  double* g = (double*)calloc((size_t)nc * nt, sizeof(double));
  double J = 0;
  rc = qoc_propagate(h, u, NULL, xf);                                  /* propagate: the caller's Jfinal sees x[end] */
  CHECK(rc == QOC_OK, "qoc_propagate rc=%d", rc);
  rc = qoc_gradient(h, u, lam, g);                                     /* grape_sensitivity with a host closure's lambda_N */
  CHECK(rc == QOC_OK && maxabsdiff(g, gref, (size_t)nc * nt) <= 1e-8 * gmax, "qoc_gradient(lambda_final) rc=%d |dg|=%.2e (|g|max %.2e)", rc,
        maxabsdiff(g, gref, (size_t)nc * nt), gmax);
  rc = qoc_set_cost(h, cost, T, n);                                    /* setup_infidelity built in */
  CHECK(rc == QOC_OK, "qoc_set_cost rc=%d", rc);
  rc = qoc_propagate(h, u, &J, NULL);
  CHECK(rc == QOC_OK && fabs(J - Jref[0]) <= 1e-10, "qoc_propagate + built-in J rc=%d |dJ|=%.2e", rc, fabs(J - Jref[0]));
  memset(g, 0, sizeof(double) * nc * nt);
  rc = qoc_gradient(h, u, NULL, g);
  CHECK(rc == QOC_OK && maxabsdiff(g, gref, (size_t)nc * nt) <= 1e-8 * gmax, "qoc_gradient(built-in) rc=%d |dg|=%.2e", rc,
        maxabsdiff(g, gref, (size_t)nc * nt));
  memset(g, 0, sizeof(double) * nc * nt);
  J = 0;
  rc = qoc_eval(h, u, &J, g);
  CHECK(rc == QOC_OK && fabs(J - Jref[0]) <= 1e-10 && maxabsdiff(g, gref, (size_t)nc * nt) <= 1e-8 * gmax, "qoc_eval rc=%d |dJ|=%.2e |dg|=%.2e",
        rc, fabs(J - Jref[0]), maxabsdiff(g, gref, (size_t)nc * nt));
  u[3] += 1e-9;                                                        /* src/gradient_computations.jl:37-39 */
  rc = qoc_gradient(h, u, NULL, g);
  CHECK(rc == QOC_ERR_STALE_CACHE, "stale u -> rc=%d (%s)", rc, qoc_last_error(h));
  u[3] -= 1e-9;
  qoc_destroy(h);

  /* ---- the same pulse time-segment sharded over n_ranks ranks driven by this one process ---- */
  {
    int devs[16], ndev_rank = n_ranks > 16 ? 16 : n_ranks;
    const char* same = getenv("ABI_DRIVER_ONE_GPU");
    for (int i = 0; i < ndev_rank; i++) devs[i] = (same && same[0] == '1') ? 0 : i;
    pr.cost = cost;
    qoc_sharded* s = NULL;
    rc = qoc_create_sharded(&pr, A0, A, x0, T, ndev_rank, devs, QOC_SHARD_TIME, &s);
    CHECK(rc == QOC_OK, "qoc_create_sharded(time, %d ranks) rc=%d (%s)", ndev_rank, rc, rc ? qoc_sharded_last_error(NULL) : "");
    if (rc == QOC_OK) {
      memset(g, 0, sizeof(double) * nc * nt);
      J = 0;
      rc = qoc_sharded_eval(s, u, &J, g);
      CHECK(rc == QOC_OK && fabs(J - Jref[0]) <= 1e-10 && maxabsdiff(g, gref, (size_t)nc * nt) <= 1e-8 * gmax,
            "qoc_sharded_eval rc=%d |dJ|=%.2e |dg|=%.2e device ms %.3f (%s)", rc, fabs(J - Jref[0]), maxabsdiff(g, gref, (size_t)nc * nt),
            qoc_sharded_last_ms(s), rc ? qoc_sharded_last_error(s) : "");
      qoc_sharded_destroy(s);
    }
  }
  printf(bad ? "ABI_DRIVER_FAILED %d\n" : "ABI_DRIVER_OK\n", bad);
  return bad ? 1 : 0;
}

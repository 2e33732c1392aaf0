/* tests/c_caller/caller.c — a C11 caller written the way the reference's own integration tests are
 * (tests/test_lobpcg.c:349-393 in the reference: state struct from <p>_lobpcg_alloc, operator from
 * linop_create with a host matvec callback, solve through the _Generic `lobpcg(alg)`), compiled against
 * include/lobpcg.h and linked to liblobpcg_b200.so.  Exit code 0 = all checks passed.
 *
 *   case 1: 1-D Dirichlet Laplacian n=100 as a FOREIGN host callback operator (staged by the solver)
 *   case 2: same matrix as a built-in device stencil operator, float
 *   case 3: invalid parameters print a message and return without touching outputs
 */
#include <math.h>
#include <stdio.h>
#include "lobpcg.h"

typedef struct { uint64_t n; } lap_ctx;

static void lap1d_matvec(const LinearOperator_d_t *op, f64 *restrict x, f64 *restrict y) {
  const uint64_t n = ((lap_ctx *)op->ctx->data)->n;
  for (uint64_t i = 0; i < n; i++) {
    f64 v = 2.0 * x[i];
    if (i > 0) v -= x[i - 1];
    if (i + 1 < n) v -= x[i + 1];
    y[i] = v;
  }
}

static const double kPi = 3.14159265358979323846;
static double analytic(uint64_t n, int j) { return 2.0 - 2.0 * cos((j + 1) * kPi / (double)(n + 1)); }

static uint64_t rng_state = 88172645463325252ULL;
static double next_uniform(void) {
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (double)(rng_state >> 11) / 9007199254740992.0 - 0.5;
}

/* device block operator supplied by the caller (lb2_op_device).  An application would launch its own kernels on `stream`;
 * this one forwards the solver's device blocks to the kernel-level ABI (a shifted stencil, A + 0.5 I). */
typedef struct { void *inner; int calls; int widest; } fwd_ctx;
static int fwd_matmat(void *user, int ncols, const void *X, int64_t ldx, void *Y, int64_t ldy, void *stream) {
  fwd_ctx *c = (fwd_ctx *)user;
  (void)stream; /* = the default context's stream, which lb2_op_apply enqueues on */
  c->calls++;
  if (ncols > c->widest) c->widest = ncols;
  return lb2_op_apply(lb2_default_ctx(), c->inner, 'd', ncols, X, ldx, Y, ldy);
}

int main(void) {
  const uint64_t n = 100, nev = 3, k = 6;
  int fails = 0;

  /* ---- case 1: host callback operator, double ---- */
  lap_ctx lc = {n};
  linop_ctx_t ctx = {&lc, sizeof(lc)};
  LinearOperator_d_t *A = linop_create(n, n, lap1d_matvec, NULL, &ctx);
  d_lobpcg_t *alg = lobpcg_alloc(n, nev, k, d);
  alg->A = A; alg->B = NULL; alg->T = NULL;
  alg->maxIter = 2000; alg->tol = 1e-8; alg->verbosity = 0;
  for (uint64_t i = 0; i < n * k; i++) alg->S[i] = next_uniform();
  lobpcg(alg);
  printf("case1: iter=%lu converged=%lu\n", (unsigned long)alg->iter, (unsigned long)alg->converged);
  if (alg->converged != nev) fails++;
  for (int j = 0; j < (int)nev; j++) {
    const double rel = fabs(alg->eigVals[j] - analytic(n, j)) / analytic(n, j);
    printf("  lambda[%d]=%.15e rel.err=%.2e res=%.2e\n", j, alg->eigVals[j], rel, alg->resNorm[j]);
    if (rel > 1e-10 || alg->resNorm[j] > alg->tol) fails++;
  }
  lobpcg_free(&alg);
  if (alg != NULL) fails++;

  /* ---- case 2: built-in device stencil, float ---- */
  LinearOperator_s_t *As = (LinearOperator_s_t *)lb2_op_stencil('s', (int64_t)n, 1, 1, 2.0, -1.0, NULL);
  s_lobpcg_t *algs = lobpcg_alloc(n, nev, k, s);
  algs->A = As; algs->maxIter = 2000; algs->tol = 1e-4f;
  for (uint64_t i = 0; i < n * k; i++) algs->S[i] = (f32)next_uniform();
  lobpcg(algs);
  printf("case2: iter=%lu converged=%lu\n", (unsigned long)algs->iter, (unsigned long)algs->converged);
  if (algs->converged != nev) fails++;
  for (int j = 0; j < (int)nev; j++) {
    const double rel = fabs((double)algs->eigVals[j] - analytic(n, j)) / analytic(n, j);
    printf("  lambda[%d]=%.7e rel.err=%.2e\n", j, (double)algs->eigVals[j], rel);
    if (rel > 1e-2) fails++;   /* reference float test accepts 1 %% (tests/test_lobpcg.c:398-434) */
  }
  /* the built-in operator is still an ordinary single-vector operator for host callers */
  {
    f32 x[100], y[100];
    for (uint64_t i = 0; i < n; i++) x[i] = 1.0f;
    linop_apply(As, x, y);
    if (fabsf(y[0] - 1.0f) > 1e-6f || fabsf(y[50]) > 1e-6f) fails++;
  }

  /* ---- case 4: caller-supplied DEVICE block operator, eigenpair write-out ---- */
  {
    LinearOperator_d_t *Ain = (LinearOperator_d_t *)lb2_op_stencil('d', (int64_t)n, 1, 1, 2.5, -1.0, NULL);
    fwd_ctx sc = {Ain, 0, 0};
    LinearOperator_d_t *Ash = (LinearOperator_d_t *)lb2_op_device('d', (int64_t)n, fwd_matmat, &sc, 4.5);
    d_lobpcg_t *a4 = lobpcg_alloc(n, nev, k, d);
    a4->A = Ash; a4->maxIter = 2000; a4->tol = 1e-8;
    for (uint64_t i = 0; i < n * k; i++) a4->S[i] = next_uniform();
    lobpcg(a4);
    printf("case4: iter=%lu converged=%lu callback calls=%d widest block=%d\n", (unsigned long)a4->iter,
           (unsigned long)a4->converged, sc.calls, sc.widest);
    if (a4->converged != nev || sc.widest < (int)k) fails++;
    for (int j = 0; j < (int)nev; j++) {
      const double ex = analytic(n, j) + 0.5, rel = fabs(a4->eigVals[j] - ex) / ex;
      printf("  lambda[%d]=%.15e rel.err=%.2e\n", j, a4->eigVals[j], rel);
      if (rel > 1e-10) fails++;
    }
    if (lb2_write_mtx("caller_eigvecs.mtx", 'd', (int64_t)n, (int64_t)nev, a4->S, (int64_t)n)) fails++;
    if (lb2_write_mtx("caller_eigvals.mtx", 'd', (int64_t)nev, 1, a4->eigVals, (int64_t)nev)) fails++;
    remove("caller_eigvecs.mtx"); remove("caller_eigvals.mtx");
    lobpcg_free(&a4);
    lb2_op_destroy(Ash);
    lb2_op_destroy(Ain);
  }

  /* ---- case 3: parameter validation (reference lobpcg_impl.inc:66-75): returns, outputs untouched ---- */
  algs->sizeSub = 50; /* 3*sizeSub > size */
  algs->eigVals[0] = -7.0f;
  lobpcg(algs);
  if (algs->eigVals[0] != -7.0f) fails++;
  algs->sizeSub = k;
  lobpcg_free(&algs);
  lb2_op_destroy(As);
  linop_destroy(&A);

  printf(fails ? "FAIL (%d)\n" : "PASS\n", fails);
  return fails;
}

/* oracle/ref_harness.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin driver around the UNMODIFIED reference library (sources compiled where they lie under
 * /root/reference by oracle/Makefile, objects only under oracle/_ref/).  It supplies what the
 * reference leaves to its callers — operator callbacks (the reference ships none, all of its
 * examples live in tests/, e.g. tests/test_lobpcg.c:48-62) — and flat C entry points a ctypes
 * caller can bind:
 *
 *   ref_<p>_op_stencil / _op_csr / _op_diag / _op_bdg   build LinearOperator_<p>_t (linop.h:20-26)
 *   ref_<p>_op_cheb                                     polynomial preconditioner over another operator (for alg->T)
 *   ref_<p>_op_apply                                    apply_block_op (src/gram/gram_impl.inc:29-33)
 *   ref_<p>_solve                                       <p>_lobpcg / <p>_ilobpcg (src/core/*_impl.inc)
 *
 * p in {s,d,c,z}.  The reference's own helpers (d_gram_self, d_svqb, d_ortho_drop,
 * d_rayleigh_ritz_modified, d_get_residual ...) are exported from the same shared object and are
 * bound directly.
 *
 * Determinism: estimate_norm seeds libc rand() with time(NULL) on first use
 * (src/residual/estimate_norm_impl.inc:20-24); this file overrides time() (linked -Bsymbolic) so
 * that runs are bit-reproducible.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <complex.h>
#include <omp.h>

#include "lobpcg.h"
#include "lobpcg/linop.h"

time_t time(time_t *t) {
  const time_t fixed = (time_t)1234567;
  if (t) *t = fixed;
  return fixed;
}

typedef struct {
  int kind; /* 0 stencil, 1 csr, 2 diag, 3 bdg */
  int64_t gx, gy, gz, n;
  double cdiag, coff, shift, dre, dim;
  const void *v;    /* per-row real addend to the diagonal (stencil) or the diagonal itself (diag) */
  const int64_t *rowptr;
  const int32_t *col;
  const void *val;
  void *scratch;
} href_ctx_t;

#define HARNESS(P, CT, RT, OPT, ISCPLX)                                                          \
  static void P##_stencil_core(const href_ctx_t *c, const CT *x, CT *y, CT extra_diag) {         \
    const int64_t gx = c->gx, gy = c->gy, gz = c->gz;                                            \
    const RT cd = (RT)c->cdiag, co = (RT)c->coff;                                                \
    const RT *v = (const RT *)c->v;                                                              \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                     \
    for (int64_t z = 0; z < gz; z++)                                                             \
      for (int64_t yy = 0; yy < gy; yy++) {                                                      \
        const int64_t base = (z * gy + yy) * gx;                                                 \
        for (int64_t xx = 0; xx < gx; xx++) {                                                    \
          const int64_t i = base + xx;                                                           \
          CT nb = 0;                                                                             \
          if (xx > 0) nb += x[i - 1];                                                            \
          if (xx + 1 < gx) nb += x[i + 1];                                                       \
          if (yy > 0) nb += x[i - gx];                                                           \
          if (yy + 1 < gy) nb += x[i + gx];                                                      \
          if (z > 0) nb += x[i - gx * gy];                                                       \
          if (z + 1 < gz) nb += x[i + gx * gy];                                                  \
          const RT d = cd + (v ? v[i] : (RT)0);                                                  \
          y[i] = (d + extra_diag) * x[i] + co * nb;                                              \
        }                                                                                        \
      }                                                                                          \
  }                                                                                              \
  static void P##_mv_stencil(const OPT *op, CT *restrict x, CT *restrict y) {                    \
    P##_stencil_core((const href_ctx_t *)op->ctx->data, x, y, (CT)0);                            \
  }                                                                                              \
  static void P##_mv_csr(const OPT *op, CT *restrict x, CT *restrict y) {                        \
    const href_ctx_t *c = (const href_ctx_t *)op->ctx->data;                                     \
    const CT *val = (const CT *)c->val;                                                          \
    _Pragma("omp parallel for schedule(static)")                                                 \
    for (int64_t i = 0; i < c->n; i++) {                                                         \
      CT acc = 0;                                                                                \
      for (int64_t p = c->rowptr[i]; p < c->rowptr[i + 1]; p++) acc += val[p] * x[c->col[p]];    \
      y[i] = acc;                                                                                \
    }                                                                                            \
  }                                                                                              \
  static void P##_mv_diag(const OPT *op, CT *restrict x, CT *restrict y) {                       \
    const href_ctx_t *c = (const href_ctx_t *)op->ctx->data;                                     \
    const RT *d = (const RT *)c->v;                                                              \
    _Pragma("omp parallel for schedule(static)")                                                 \
    for (int64_t i = 0; i < c->n; i++) y[i] = d[i] * x[i];                                       \
  }                                                                                              \
  static void P##_mv_bdg(const OPT *op, CT *restrict x, CT *restrict y) {                        \
    /* A = [[K + s, d],[conj(d), K + s]] on [u; v], K = stencil */                               \
    const href_ctx_t *c = (const href_ctx_t *)op->ctx->data;                                     \
    const int64_t m = c->gx * c->gy * c->gz;                                                     \
    P##_stencil_core(c, x, y, (CT)(RT)c->shift);                                                 \
    P##_stencil_core(c, x + m, y + m, (CT)(RT)c->shift);                                         \
    const CT d = ISCPLX ? (CT)((RT)c->dre + (RT)c->dim * I) : (CT)(RT)c->dre;                    \
    const CT dc = ISCPLX ? (CT)((RT)c->dre - (RT)c->dim * I) : (CT)(RT)c->dre;                   \
    _Pragma("omp parallel for schedule(static)")                                                 \
    for (int64_t i = 0; i < m; i++) {                                                            \
      y[i] += d * x[m + i];                                                                      \
      y[m + i] += dc * x[i];                                                                     \
    }                                                                                            \
  }                                                                                              \
  static void *P##_mk(uint64_t n, matvec_func_##P##_t mv, href_ctx_t proto) {                    \
    href_ctx_t *c = calloc(1, sizeof(href_ctx_t));                                               \
    *c = proto;                                                                                  \
    linop_ctx_t *lc = calloc(1, sizeof(linop_ctx_t));                                            \
    lc->data = c;                                                                                \
    lc->data_size = sizeof(href_ctx_t);                                                          \
    return linop_create_##P(n, n, mv, NULL, lc);                                                 \
  }                                                                                              \
  void *ref_##P##_op_stencil(int64_t gx, int64_t gy, int64_t gz, double cdiag, double coff,      \
                             const RT *v) {                                                      \
    href_ctx_t c = {0};                                                                          \
    c.kind = 0; c.gx = gx; c.gy = gy; c.gz = gz; c.n = gx * gy * gz;                             \
    c.cdiag = cdiag; c.coff = coff; c.v = v;                                                     \
    return P##_mk((uint64_t)c.n, P##_mv_stencil, c);                                             \
  }                                                                                              \
  void *ref_##P##_op_csr(int64_t n, const int64_t *rowptr, const int32_t *col, const CT *val) {  \
    href_ctx_t c = {0};                                                                          \
    c.kind = 1; c.n = n; c.rowptr = rowptr; c.col = col; c.val = val;                            \
    return P##_mk((uint64_t)n, P##_mv_csr, c);                                                   \
  }                                                                                              \
  void *ref_##P##_op_diag(int64_t n, const RT *d) {                                              \
    href_ctx_t c = {0};                                                                          \
    c.kind = 2; c.n = n; c.v = d;                                                                \
    return P##_mk((uint64_t)n, P##_mv_diag, c);                                                  \
  }                                                                                              \
  void *ref_##P##_op_bdg(int64_t gx, int64_t gy, int64_t gz, double cdiag, double coff,          \
                         double shift, double dre, double dim) {                                 \
    href_ctx_t c = {0};                                                                          \
    c.kind = 3; c.gx = gx; c.gy = gy; c.gz = gz; c.n = 2 * gx * gy * gz;                         \
    c.cdiag = cdiag; c.coff = coff; c.shift = shift; c.dre = dre; c.dim = dim;                   \
    return P##_mk((uint64_t)c.n, P##_mv_bdg, c);                                                 \
  }                                                                                              \
  /* T = p(A): `degree` Chebyshev steps for A y = x on [lo, hi] through the inner operator's own matvec — the   \
   * reference-side twin of lb2_op_chebyshev, used as alg->T of the UNMODIFIED reference solver in parity runs */  \
  static void P##_mv_cheb(const OPT *op, CT *restrict x, CT *restrict y) {                       \
    const href_ctx_t *c = (const href_ctx_t *)op->ctx->data;                                     \
    const OPT *A = (const OPT *)c->val;                                                          \
    const int64_t n = c->n;                                                                      \
    CT *r = (CT *)c->scratch, *d = r + n, *ad = d + n;                                           \
    const double theta = 0.5 * (c->dim + c->dre), delta = 0.5 * (c->dim - c->dre);               \
    const double sigma = theta / delta;                                                          \
    double rho_old = 1.0 / sigma;                                                                \
    for (int64_t i = 0; i < n; i++) { r[i] = x[i]; d[i] = x[i] * (RT)(1.0 / theta); y[i] = d[i]; } \
    for (int64_t j = 0; j < c->gx; j++) {                                                        \
      A->matvec(A, d, ad);                                                                       \
      const double rho = 1.0 / (2.0 * sigma - rho_old);                                          \
      const RT c1 = (RT)(rho * rho_old), c2 = (RT)(2.0 * rho / delta);                           \
      for (int64_t i = 0; i < n; i++) {                                                          \
        r[i] -= ad[i];                                                                           \
        d[i] = c1 * d[i] + c2 * r[i];                                                            \
        y[i] += d[i];                                                                            \
      }                                                                                          \
      rho_old = rho;                                                                             \
    }                                                                                            \
  }                                                                                              \
  void *ref_##P##_op_cheb(void *inner, int64_t degree, double lo, double hi) {                   \
    href_ctx_t c = {0};                                                                          \
    c.kind = 4; c.n = (int64_t)((OPT *)inner)->rows; c.gx = degree; c.dre = lo; c.dim = hi;      \
    c.val = inner;                                                                               \
    c.scratch = calloc((size_t)(3 * c.n), sizeof(CT));                                           \
    return P##_mk((uint64_t)c.n, P##_mv_cheb, c);                                                \
  }                                                                                              \
  void ref_##P##_op_free(void *opv) {                                                            \
    OPT *op = (OPT *)opv;                                                                        \
    if (!op) return;                                                                             \
    free(((href_ctx_t *)op->ctx->data)->scratch);                                                \
    free(op->ctx->data);                                                                         \
    free(op->ctx);                                                                               \
    free(op);                                                                                    \
  }                                                                                              \
  void ref_##P##_op_apply(void *opv, CT *x, CT *y, uint64_t ncols) {                             \
    OPT *op = (OPT *)opv;                                                                        \
    P##_apply_block_op(op, x, y, op->rows, ncols);                                               \
  }                                                                                              \
  /* Runs the reference solver. X is n x k column-major, in: X0 (all zero => reference fills     \
   * with libc rand), out: eigenvectors.  Returns 0. */                                          \
  int ref_##P##_solve(int indefinite, uint64_t n, uint64_t nev, uint64_t k, uint64_t maxIter,    \
                      double tol, void *A, void *B, void *T, CT *X, RT *eig, RT *res,            \
                      int8_t *sig, uint64_t *iter, uint64_t *conv, int verbosity) {              \
    P##_lobpcg_t *alg = indefinite ? P##_ilobpcg_alloc(n, nev, k) : P##_lobpcg_alloc(n, nev, k); \
    alg->A = (OPT *)A; alg->B = (OPT *)B; alg->T = (OPT *)T;                                     \
    alg->maxIter = maxIter; alg->tol = (RT)tol; alg->verbosity = (int8_t)verbosity;              \
    memcpy(alg->S, X, n * k * sizeof(CT));                                                       \
    if (indefinite) P##_ilobpcg(alg); else P##_lobpcg(alg);                                      \
    memcpy(X, alg->S, n * k * sizeof(CT));                                                       \
    memcpy(eig, alg->eigVals, k * sizeof(RT));                                                   \
    memcpy(res, alg->resNorm, k * sizeof(RT));                                                   \
    if (sig && alg->signature) memcpy(sig, alg->signature, 3 * k);                               \
    *iter = alg->iter; *conv = alg->converged;                                                   \
    P##_lobpcg_free(&alg);                                                                       \
    return 0;                                                                                    \
  }

HARNESS(s, f32, f32, LinearOperator_s_t, 0)
HARNESS(d, f64, f64, LinearOperator_d_t, 0)
HARNESS(c, c32, f32, LinearOperator_c_t, 1)
HARNESS(z, c64, f64, LinearOperator_z_t, 1)

/* BLAS threading control + provenance, so the bench can state cores and BLAS build. */
extern void scipy_openblas_set_num_threads(int);
extern char *scipy_openblas_get_config(void);
void ref_set_threads(int n) {
  scipy_openblas_set_num_threads(n);
  omp_set_num_threads(n);
}
const char *ref_blas_config(void) { return scipy_openblas_get_config(); }

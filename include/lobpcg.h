/* include/lobpcg.h — drop-in front end of the B200-native LOBPCG hot path.
 *
 * ABI-compatible with the reference's public header for the solver path (reference lobpcg.h:10-92,
 * 590-686 and include/lobpcg/linop.h:7-81): same state-struct field order and types, same
 * LinearOperator layout, same entry-point names, same C11 _Generic dispatch.  A caller compiled against
 * the reference header can be relinked against liblobpcg_b200.so without source changes; a caller
 * compiled against this header gets the same thing plus the built-in device operators declared in
 * lobpcg_b200.h.
 *
 * Semantics kept from the reference (src/core/lobpcg_impl.inc:60-248):
 *   - alg->S[0 : size*sizeSub) holds the initial block X0 on entry (all zero => random start) and the
 *     eigenvectors on exit; eigVals[0:sizeSub), resNorm[0:nev), converged, iter are outputs;
 *   - A is required; B and T may be NULL (ilobpcg requires B);
 *   - errors never abort: a message goes to stderr and the call returns with outputs untouched.
 * These buffers are HOST memory (as allocated by <p>_lobpcg_alloc); the solver uploads X0 once, keeps
 * every block vector and the projected problem on the GPU, and downloads the results once.
 */
#ifndef LOBPCG_B200_COMPAT_LOBPCG_H
#define LOBPCG_B200_COMPAT_LOBPCG_H

#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#ifdef __cplusplus
#error "include/lobpcg.h is the C11 front end; C++ callers use lobpcg_b200.h"
#endif
#include <complex.h>

typedef float f32;
typedef double f64;
typedef float complex c32;
typedef double complex c64;

/* ---- linear operator: single-vector callback interface (reference linop.h:7-26) ------------------- */
typedef struct {
  void *data;
  size_t data_size;
} linop_ctx_t;

#define LB2_DECLARE_LINOP(P, CT)                                                            \
  typedef struct LinearOperator_##P##_t LinearOperator_##P##_t;                             \
  typedef void (*matvec_func_##P##_t)(const LinearOperator_##P##_t *op, CT *restrict x,     \
                                      CT *restrict y);                                      \
  typedef void (*cleanup_func_##P##_t)(linop_ctx_t *ctx);                                   \
  struct LinearOperator_##P##_t {                                                           \
    uint64_t rows, cols;                                                                    \
    matvec_func_##P##_t matvec;                                                             \
    cleanup_func_##P##_t cleanup;                                                           \
    linop_ctx_t *ctx;                                                                       \
  };                                                                                        \
  static inline LinearOperator_##P##_t *linop_create_##P(                                   \
      uint64_t rows, uint64_t cols, matvec_func_##P##_t mv, cleanup_func_##P##_t cl,        \
      linop_ctx_t *ctx) {                                                                   \
    LinearOperator_##P##_t *op = (LinearOperator_##P##_t *)calloc(1, sizeof(*op));          \
    if (!op) { fprintf(stderr, "linop_create: out of memory\n"); exit(1); }                 \
    op->rows = rows; op->cols = cols; op->matvec = mv; op->cleanup = cl; op->ctx = ctx;     \
    return op;                                                                              \
  }                                                                                         \
  static inline void linop_destroy_##P(LinearOperator_##P##_t **op) {                       \
    if (!op || !*op) return;                                                                \
    if ((*op)->cleanup && (*op)->ctx) (*op)->cleanup((*op)->ctx);                           \
    free(*op);                                                                              \
    *op = NULL;                                                                             \
  }                                                                                         \
  static inline void linop_apply_##P(const LinearOperator_##P##_t *op, CT *restrict x,      \
                                     CT *restrict y) {                                      \
    op->matvec(op, x, y);                                                                   \
  }

LB2_DECLARE_LINOP(s, f32)
LB2_DECLARE_LINOP(d, f64)
LB2_DECLARE_LINOP(c, c32)
LB2_DECLARE_LINOP(z, c64)

#define linop_create(rows, cols, mv, cl, ctx)                                               \
  _Generic((mv), matvec_func_s_t: linop_create_s, matvec_func_d_t: linop_create_d,          \
           matvec_func_c_t: linop_create_c, matvec_func_z_t: linop_create_z)(rows, cols, mv, cl, ctx)
#define linop_apply(op, x, y)                                                               \
  _Generic((op), LinearOperator_s_t *: linop_apply_s, LinearOperator_d_t *: linop_apply_d,  \
           LinearOperator_c_t *: linop_apply_c, LinearOperator_z_t *: linop_apply_z)(op, x, y)
#define linop_destroy(op)                                                                   \
  _Generic((op), LinearOperator_s_t **: linop_destroy_s, LinearOperator_d_t **: linop_destroy_d, \
           LinearOperator_c_t **: linop_destroy_c, LinearOperator_z_t **: linop_destroy_z)(op)

/* ---- solver state (field order == reference lobpcg.h:13-55; do not reorder) ------------------------ */
#define LB2_DECLARE_STATE(P, CT, RT)                                                        \
  typedef struct P##_lobpcg_t P##_lobpcg_t;                                                 \
  struct P##_lobpcg_t {                                                                     \
    CT *restrict S, *restrict Cx, *restrict Cp;      /* [X|P|W] slab, RR coefficients     */ \
    CT *restrict AX;                                 /* cached A*X                        */ \
    CT *restrict AS, *restrict BS;                   /* reserved by the reference, unused */ \
    RT *restrict eigVals, *restrict resNorm;                                                \
    int8_t *restrict signature;                      /* ilobpcg only                      */ \
    CT *restrict wrk1, *restrict wrk2, *restrict wrk3, *restrict wrk4;                      \
    RT *restrict rr_D, *restrict rr_eigvals;                                                \
    CT *restrict rr_tau, *restrict rr_VR;                                                   \
    int8_t *restrict rr_sig;                                                                \
    uint64_t *restrict rr_indices;                                                          \
    CT *restrict rr_ggev;                                                                   \
    int8_t implicit_product_update, verbosity;                                              \
    uint64_t iter, nev, converged, size, sizeSub, maxIter;                                  \
    RT tol;                                                                                 \
    LinearOperator_##P##_t *A, *B, *T;                                                      \
  };                                                                                        \
  /* exported by liblobpcg_b200.so (lobpcg_b200/csrc/capi.cu) */                            \
  void P##_lobpcg(P##_lobpcg_t *alg);                                                       \
  void P##_ilobpcg(P##_lobpcg_t *alg);                                                      \
  P##_lobpcg_t *lb2_##P##_state_alloc(uint64_t n, uint64_t nev, uint64_t sizeSub, int indefinite); \
  void lb2_##P##_state_free(P##_lobpcg_t *alg);                                             \
  /* The device solver only touches S[0:n*sizeSub), eigVals, resNorm, signature: the host */ \
  /* workspaces of the reference (12 n k scalars) are not allocated here.                 */ \
  static inline P##_lobpcg_t *P##_lobpcg_alloc(uint64_t n, uint64_t nev, uint64_t sizeSub) { \
    return lb2_##P##_state_alloc(n, nev, sizeSub, 0);                                       \
  }                                                                                         \
  static inline P##_lobpcg_t *P##_ilobpcg_alloc(uint64_t n, uint64_t nev, uint64_t sizeSub) { \
    return lb2_##P##_state_alloc(n, nev, sizeSub, 1);                                       \
  }                                                                                         \
  static inline void P##_lobpcg_free(P##_lobpcg_t **alg) {                                  \
    if (alg && *alg) { lb2_##P##_state_free(*alg); *alg = NULL; }                           \
  }

LB2_DECLARE_STATE(s, f32, f32)
LB2_DECLARE_STATE(d, f64, f64)
LB2_DECLARE_STATE(c, c32, f32)
LB2_DECLARE_STATE(z, c64, f64)

#define lobpcg(alg)                                                                         \
  _Generic((alg), s_lobpcg_t *: s_lobpcg, d_lobpcg_t *: d_lobpcg, c_lobpcg_t *: c_lobpcg,   \
           z_lobpcg_t *: z_lobpcg)(alg)
#define ilobpcg(alg)                                                                        \
  _Generic((alg), s_lobpcg_t *: s_ilobpcg, d_lobpcg_t *: d_ilobpcg, c_lobpcg_t *: c_ilobpcg, \
           z_lobpcg_t *: z_ilobpcg)(alg)
#define lobpcg_alloc(n, nev, sizeSub, prefix) prefix##_lobpcg_alloc(n, nev, sizeSub)
#define ilobpcg_alloc(n, nev, sizeSub, prefix) prefix##_ilobpcg_alloc(n, nev, sizeSub)
#define lobpcg_free(alg)                                                                    \
  _Generic((*alg), s_lobpcg_t *: s_lobpcg_free, d_lobpcg_t *: d_lobpcg_free,                \
           c_lobpcg_t *: c_lobpcg_free, z_lobpcg_t *: z_lobpcg_free)(alg)

#include "lobpcg_b200.h"

#endif /* LOBPCG_B200_COMPAT_LOBPCG_H */

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


/* ---- L2-L4 helpers of the reference (reference lobpcg.h:98-555), same names and signatures, on HOST buffers ----
 * Exported so that unit-level callers of the reference link unchanged; each call stages its operands through the
 * device (the solver itself never does).  Workspace arguments are accepted and ignored. */
#define LB2_DECLARE_HELPERS(P, CT, RT)                                                                         \
  void P##_apply_block_op(const LinearOperator_##P##_t *Op, CT *restrict X, CT *restrict Y, const uint64_t n,  \
                          const uint64_t k);                                                                   \
  void P##_gram_self(CT *restrict U, const uint64_t n, const uint64_t k, const LinearOperator_##P##_t *B,      \
                     CT *restrict G, const uint64_t ldg, CT *restrict wrk);                                    \
  void P##_gram_cross(CT *restrict V, const uint64_t nv, CT *restrict U, const uint64_t nu, const uint64_t n,  \
                      const LinearOperator_##P##_t *B, CT *restrict G, const uint64_t ldg, CT *restrict wrk);  \
  void P##_gram_self_mat(CT *restrict U, const uint64_t n, const uint64_t k, const CT *mat, CT *restrict G,    \
                         const uint64_t ldg, CT *restrict wrk);                                                \
  void P##_gram_cross_mat(CT *restrict V, const uint64_t nv, CT *restrict U, const uint64_t nu,                \
                          const uint64_t n, const CT *mat, CT *restrict G, const uint64_t ldg,                 \
                          CT *restrict wrk);                                                                   \
  void P##_get_residual(const uint64_t size, const uint64_t sizeSub, CT *restrict X, CT *restrict AX,          \
                        CT *restrict R, RT *restrict eigVal, CT *restrict wrk, LinearOperator_##P##_t *A,      \
                        LinearOperator_##P##_t *B);                                                            \
  void P##_get_residual_norm(const uint64_t size, const uint64_t nev, CT *restrict W, RT *restrict eigVals,    \
                             RT *restrict resNorm, CT *restrict wrk1, CT *restrict wrk2, CT *restrict wrk3,    \
                             const RT ANorm, const RT BNorm, LinearOperator_##P##_t *B);                       \
  uint64_t P##_svqb(const uint64_t m, const uint64_t n, const RT tau, const char drop, CT *restrict U,         \
                    CT *restrict wrk1, CT *restrict wrk2, CT *restrict wrk3, LinearOperator_##P##_t *B);       \
  uint64_t P##_svqb_mat(const uint64_t m, const uint64_t n, const RT tau, const char drop, CT *restrict U,     \
                        CT *restrict mat, CT *restrict wrk1, CT *restrict wrk2, CT *restrict wrk3);            \
  uint64_t P##_ortho_drop(const uint64_t m, const uint64_t n_u, const uint64_t n_v, const RT eps_ortho,        \
                          const RT eps_drop, CT *restrict U, CT *restrict V, CT *restrict wrk1,                \
                          CT *restrict wrk2, CT *restrict wrk3, LinearOperator_##P##_t *B);                    \
  uint64_t P##_ortho_indefinite(const uint64_t m, const uint64_t n_u, const uint64_t n_v, const RT eps_ortho,  \
                                const RT eps_drop, CT *restrict U, CT *restrict V, CT *restrict sig,           \
                                CT *restrict wrk1, CT *restrict wrk2, CT *restrict wrk3,                       \
                                LinearOperator_##P##_t *B);                                                    \
  uint64_t P##_ortho_indefinite_mat(const uint64_t m, const uint64_t n_u, const uint64_t n_v,                  \
                                    const RT eps_ortho, const RT eps_drop, CT *restrict U, CT *restrict V,     \
                                    CT *restrict mat, CT *restrict wrk1, CT *restrict wrk2, CT *restrict wrk3); \
  void P##_rayleigh_ritz(const uint64_t size, const uint64_t sizeSub, CT *restrict S, CT *restrict Cx,         \
                         RT *restrict eigVal, CT *restrict wrk1, CT *restrict wrk2, CT *restrict wrk3,         \
                         RT *restrict rr_D, LinearOperator_##P##_t *A, LinearOperator_##P##_t *B);             \
  void P##_rayleigh_ritz_modified(const uint64_t size, const uint64_t nx, const uint64_t mult,                 \
                                  const uint64_t nconv, const uint64_t ndrop, uint8_t *useOrtho,               \
                                  CT *restrict S, const CT *restrict AX, CT *restrict wrk1, CT *restrict wrk2, \
                                  CT *restrict wrk3, CT *restrict Cx, CT *restrict Cp, RT *restrict eigVal,    \
                                  RT *restrict rr_eigvals, CT *restrict rr_tau, RT *restrict rr_D,             \
                                  LinearOperator_##P##_t *A, LinearOperator_##P##_t *B);                       \
  void P##_indefinite_rayleigh_ritz(const uint64_t size, const uint64_t sizeSub, CT *restrict S,               \
                                    CT *restrict Cx, RT *restrict eigVal, int8_t *restrict signature,          \
                                    CT *restrict wrk1, CT *restrict wrk2, CT *restrict wrk3, CT *restrict wrk4, \
                                    uint64_t *restrict rr_indices, CT *restrict rr_ggev,                       \
                                    LinearOperator_##P##_t *A, LinearOperator_##P##_t *B);                     \
  void P##_indefinite_rayleigh_ritz_modified(                                                                  \
      const uint64_t size, const uint64_t nx, const uint64_t mult, const uint64_t nconv, const uint64_t ndrop, \
      CT *restrict S, const CT *restrict AX, CT *restrict wrk1, CT *restrict wrk2, CT *restrict wrk3,          \
      CT *restrict wrk4, CT *Cx, CT *restrict Cp, CT *Cx_ortho, RT *restrict eigVal, int8_t *restrict signature, \
      int *restrict quality_flag, RT *restrict rr_eigvals, int8_t *restrict rr_sig,                            \
      uint64_t *restrict rr_indices, CT *restrict rr_VR, CT *restrict rr_ggev, LinearOperator_##P##_t *A,      \
      LinearOperator_##P##_t *B);                                                                              \
  void P##_fill_random(uint64_t n, CT *x);                                                                     \
  RT P##_estimate_norm(uint64_t size, LinearOperator_##P##_t *A, CT *wrk1, CT *wrk2);

LB2_DECLARE_HELPERS(s, f32, f32)
LB2_DECLARE_HELPERS(d, f64, f64)
LB2_DECLARE_HELPERS(c, c32, f32)
LB2_DECLARE_HELPERS(z, c64, f64)

#define LB2_G4(x, name) _Generic((x), f32 *: s_##name, f64 *: d_##name, c32 *: c_##name, c64 *: z_##name)
#define get_residual(size, sizeSub, X, AX, R, eigVal, wrk, A, B) LB2_G4(X, get_residual)(size, sizeSub, X, AX, R, eigVal, wrk, A, B)
#define get_residual_norm(size, nev, W, eigVals, resNorm, w1, w2, w3, ANorm, BNorm, B) \
  LB2_G4(W, get_residual_norm)(size, nev, W, eigVals, resNorm, w1, w2, w3, ANorm, BNorm, B)
#define svqb(m, n, tau, drop, U, w1, w2, w3, B) LB2_G4(U, svqb)(m, n, tau, drop, U, w1, w2, w3, B)
#define svqb_mat(m, n, tau, drop, U, mat, w1, w2, w3) LB2_G4(U, svqb_mat)(m, n, tau, drop, U, mat, w1, w2, w3)
#define ortho_drop(m, n_u, n_v, eo, ed, U, V, w1, w2, w3, B) LB2_G4(U, ortho_drop)(m, n_u, n_v, eo, ed, U, V, w1, w2, w3, B)
#define ortho_indefinite(m, n_u, n_v, eo, ed, U, V, sig, w1, w2, w3, B) \
  LB2_G4(U, ortho_indefinite)(m, n_u, n_v, eo, ed, U, V, sig, w1, w2, w3, B)
#define ortho_indefinite_mat(m, n_u, n_v, eo, ed, U, V, mat, w1, w2, w3) \
  LB2_G4(U, ortho_indefinite_mat)(m, n_u, n_v, eo, ed, U, V, mat, w1, w2, w3)
#define rayleigh_ritz(size, sizeSub, S, Cx, eigVal, w1, w2, w3, rr_D, A, B) \
  LB2_G4(S, rayleigh_ritz)(size, sizeSub, S, Cx, eigVal, w1, w2, w3, rr_D, A, B)
#define rayleigh_ritz_modified(size, nx, mult, nconv, ndrop, useOrtho, S, AX, w1, w2, w3, Cx, Cp, eigVal, re, rt, rd, A, B) \
  LB2_G4(S, rayleigh_ritz_modified)(size, nx, mult, nconv, ndrop, useOrtho, S, AX, w1, w2, w3, Cx, Cp, eigVal, re, rt, rd, A, B)
#define apply_block_op(Op, X, Y, n, k) LB2_G4(X, apply_block_op)(Op, X, Y, n, k)
#define gram_self(U, n, k, B, G, ldg, wrk) LB2_G4(U, gram_self)(U, n, k, B, G, ldg, wrk)
#define gram_cross(V, nv, U, nu, n, B, G, ldg, wrk) LB2_G4(U, gram_cross)(V, nv, U, nu, n, B, G, ldg, wrk)
#define gram_self_mat(U, n, k, mat, G, ldg, wrk) LB2_G4(U, gram_self_mat)(U, n, k, mat, G, ldg, wrk)
#define gram_cross_mat(V, nv, U, nu, n, mat, G, ldg, wrk) LB2_G4(U, gram_cross_mat)(V, nv, U, nu, n, mat, G, ldg, wrk)
#define fill_random(n, x) LB2_G4(x, fill_random)(n, x)

#include "lobpcg_b200.h"

#endif /* LOBPCG_B200_COMPAT_LOBPCG_H */

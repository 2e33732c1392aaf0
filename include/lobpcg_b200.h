/* include/lobpcg_b200.h — C ABI of liblobpcg_b200.so (plain pointers and sizes, no C++/torch types).
 *
 * Three layers, all `extern "C"`:
 *
 *  (1) reference entry points  <p>_lobpcg / <p>_ilobpcg, p in {s,d,c,z}  — declared in lobpcg.h, replace
 *      reference src/core/lobpcg_impl.inc:60 and src/core/ilobpcg_impl.inc:54 (host state struct in, host
 *      results out).
 *  (2) built-in device operators returned as ordinary LinearOperator_<p>_t* (reference
 *      include/lobpcg/linop.h:20-26); the solver recognises them by a tag in ctx->data and applies them to
 *      whole blocks on the GPU (replaces the per-column loop of src/gram/gram_impl.inc:29-33).  Their
 *      `matvec` member also works on host vectors, so reference-style callers can still apply them.
 *      Any other operator is treated as a host callback and staged through host memory column by column.
 *  (3) kernel-level entry points on DEVICE pointers (column-major, leading dimension in elements) — what
 *      a maintainer of the reference would bind from src/gram, src/ortho, src/residual if only single
 *      calls are to be offloaded; also what tests/ and bench.py drive.
 *
 * Scalars: s=float, d=double, c=float _Complex, z=double _Complex (interleaved re,im).  Complex and
 * real data are passed as `void*` / `const void*`; real-valued side arrays (eigenvalues, norms,
 * diagonals) are float for s,c and double for d,z.  All functions return 0 on success, non-zero after
 * printing a message to stderr; nothing throws or aborts.
 */
#ifndef LOBPCG_B200_H
#define LOBPCG_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lb2_ctx lb2_ctx;       /* one device + one stream + scratch */
typedef struct lb2_solver lb2_solver; /* resumable solver instance (device resident) */

/* ---- context ------------------------------------------------------------------------------------- */
lb2_ctx *lb2_ctx_create(int device, void *cuda_stream /* cudaStream_t or NULL => own stream */);
void lb2_ctx_destroy(lb2_ctx *ctx);
int lb2_ctx_sync(lb2_ctx *ctx);
int lb2_ctx_set_option(lb2_ctx *ctx, const char *key, int value);
int lb2_ctx_trim(lb2_ctx *ctx); /* free the solver arena this context keeps for reuse by its next solve (LB2_ARENA_CACHE=0: never kept) */
unsigned long long lb2_ctx_launches(lb2_ctx *ctx); /* kernels launched so far through this context */
/* device times (ms) of the int8 tensor-path Gram since the last query: out4 = {split, MMA kernel, reduce, calls} */
int lb2_ctx_oz_stats(lb2_ctx *ctx, double *out4);
/* host-only self-check of the int8 schedules of the column-block products (no device needed): mode 0 = one tile per CTA with an
 * equal-cost cut, 1 = lock-step cohorts, 2 = 4-CTA clusters (nworkers = clusters).  0 = exact cover of the outputs, the items of every
 * (tile, level group) partition the rows.  stats[4] (may be NULL): items, busiest worker / mean, tiles written, workers used. */
int lb2_oz_plan_check(int m, int nw, int nprod, int tri_c0, int64_t n, int nworkers, int mode, double *stats);
lb2_ctx *lb2_default_ctx(void);                    /* lazily created context on the current device */

/* device memory for FFI callers that have no CUDA runtime of their own */
void *lb2_malloc(size_t bytes);
void lb2_free(void *dptr);
void *lb2_malloc_host(size_t bytes); /* pinned */
void lb2_free_host(void *hptr);
int lb2_memcpy_h2d(lb2_ctx *ctx, void *dst, const void *src, size_t bytes);
int lb2_memcpy_d2h(lb2_ctx *ctx, void *dst, const void *src, size_t bytes);
int lb2_memset(lb2_ctx *ctx, void *dst, int byte, size_t bytes);

/* ---- (2) built-in operators: host arrays in, device-resident operator out -------------------------- */
/* Dirichlet stencil on a gx*gy*gz grid (x fastest): y_i = (cdiag + potential_i) x_i + coff * sum(nbrs).
 * gy = gz = 1 gives 1-D, gz = 1 gives 2-D.  potential may be NULL. */
void *lb2_op_stencil(char prefix, int64_t gx, int64_t gy, int64_t gz, double cdiag, double coff,
                     const void *potential_host);
/* CSR with int64 row pointers and int32 column indices (host arrays are copied). */
void *lb2_op_csr(char prefix, int64_t n, const int64_t *rowptr_host, const int32_t *col_host,
                 const void *val_host);
/* real diagonal (mass matrix B, Jacobi preconditioner T) */
void *lb2_op_diag(char prefix, int64_t n, const void *diag_host);
/* BdG-style pencil block A = [[K+shift, d],[conj d, K+shift]] with K the stencil above (config C4) */
void *lb2_op_bdg(char prefix, int64_t gx, int64_t gy, int64_t gz, double cdiag, double coff, double shift,
                 double d_re, double d_im);
/* CSR with int32 row pointers; Matrix Market coordinate file -> CSR operator (symmetric / hermitian storage expanded) */
void *lb2_op_csr32(char prefix, int64_t n, const int32_t *rowptr_host, const int32_t *col_host, const void *val_host);
void *lb2_op_csr_from_mtx(char prefix, const char *path);
/* eigenpair write-out: host block (column-major, leading dimension ld, scalars of type `prefix`) -> Matrix Market dense
 * "array" file; eigenvalues are a rows x 1 block of the real type ('s' / 'd').  0 on success. */
int lb2_write_mtx(const char *path, char prefix, int64_t rows, int64_t cols, const void *host, int64_t ld);
/* dense n x n operator (column-major host matrix of the operator's scalar type, uploaded once; block apply = library GEMM) */
void *lb2_op_dense(char prefix, int64_t n, const void *A_host);
/* Caller-supplied block operator on DEVICE pointers: the extension of the reference's one-host-vector-at-a-time operator
 * interface (include/lobpcg/linop.h:15-26) for applications that already have CUDA kernels.  fn computes Y = Op X for
 * `ncols` columns (column-major, n rows, leading dimensions ldx / ldy, scalars of the operator's type), must enqueue all
 * of its work on `cuda_stream` (a cudaStream_t) without synchronising, and returns 0 on success (any other value aborts
 * the solve like a failed kernel launch).  spec_hi > 0 passes an upper bound of the spectrum for lb2_op_chebyshev.
 * The result is an ordinary LinearOperator_<p>_t*; its matvec also works on host vectors. */
typedef int (*lb2_matmat_fn)(void *user, int ncols, const void *X_dev, int64_t ldx, void *Y_dev, int64_t ldy,
                             void *cuda_stream);
void *lb2_op_device(char prefix, int64_t n, lb2_matmat_fn fn, void *user, double spec_hi);
/* Built-in preconditioner for alg->T (the reference only plans built-ins, README.md:15): T = p(A), `degree`
 * steps of the Chebyshev iteration for A y = x on the spectrum window [lo, hi], i.e. `degree` block applies of A per
 * block apply of T.  hi <= 0: Gershgorin bound of the built-in inner operator; lo <= 0: hi / 50.  The inner operator
 * must outlive the result. */
void *lb2_op_chebyshev(char prefix, const void *inner_linop, int degree, double lo, double hi);
/* same polynomial, evaluated in float / complex float inside a double / complex double solve (prefix 'd' / 'z', built-in
 * stencil inner operator; otherwise identical to lb2_op_chebyshev): half the HBM traffic per preconditioner apply */
void *lb2_op_chebyshev_mixed(char prefix, const void *inner_linop, int degree, double lo, double hi);
/* upper bound of the spectrum recorded at construction (Gershgorin); 0 = unknown, -1 = not a built-in operator */
double lb2_op_spec_hi(const void *linop);
void lb2_op_destroy(void *linop);
/* Y = Op X on device block vectors (n x nc) */
int lb2_op_apply(lb2_ctx *ctx, const void *linop, char prefix, int nc, const void *X, int64_t ldx, void *Y,
                 int64_t ldy);

/* ---- (1b) resumable solver on top of the same state struct --------------------------------------- */
/* alg is a <prefix>_lobpcg_t*.  create+init = everything before the while loop of lobpcg_impl.inc:130;
 * step(k) = at most k passes of that loop (stops early on convergence, returns passes done or <0);
 * finish = download X, eigVals, resNorm, converged, iter into alg. */
lb2_solver *lb2_solver_create(lb2_ctx *ctx, char prefix, void *alg, int indefinite);
int lb2_solver_prepare(lb2_solver *s); /* validate + allocate only (called by init if not done before) */
int lb2_solver_init(lb2_solver *s);
int lb2_solver_step(lb2_solver *s, int max_steps);
int lb2_solver_finish(lb2_solver *s);
void lb2_solver_destroy(lb2_solver *s);
/* X0 generated on the device from the portable counter-based generator instead of uploading alg->S */
int lb2_solver_set_device_x0(lb2_solver *s, uint64_t seed);
/* device-pointer fast path: X0 is read from x0_dev and/or the eigenvectors are written to x_out_dev (device blocks of this
 * rank's rows, n_local x sizeSub, column-major, ld = n_local) instead of the host buffer alg->S; NULL keeps the host path.
 * eigVals / resNorm / converged / iter are still returned through the state struct. */
int lb2_solver_set_device_io(lb2_solver *s, const void *x0_dev, void *x_out_dev);
/* statistics: per-phase device milliseconds accumulated by step(); names via lb2_solver_stat_name */
int lb2_solver_num_stats(void);
const char *lb2_solver_stat_name(int i);
double lb2_solver_stat(lb2_solver *s, int i);
/* algorithmic work of the phase: flops for gram / tall_nn, bytes for spmm / residual */
double lb2_solver_stat_work(lb2_solver *s, int i);
unsigned long long lb2_solver_stat_calls(lb2_solver *s, int i);
void lb2_solver_reset_stats(lb2_solver *s);
int lb2_solver_state(lb2_solver *s, uint64_t *iter, uint64_t *converged, int *use_ortho);
/* solver options (before lb2_solver_init unless noted): "gram_cache" 0/1 — cached Gram blocks (default 1, definite solver;
 * env LB2_GRAM_CACHE), "gram_cache_period" — passes between recomputations of the cached blocks (any time),
 * "force_ortho" 1 — every pass in the ortho branch (any time), "debug_min_conv" n — soft-lock at least n columns (timing). */
int lb2_solver_set_option(lb2_solver *s, const char *key, int value);
/* "gram_cache", "gram_cache_refreshes", "gram_cache_monitor", "gram_cache_monitor_max", "arena_bytes", "arena_columns" */
double lb2_solver_info(lb2_solver *s, const char *key);
/* status of the last <p>_lobpcg / <p>_ilobpcg call on this thread (they return void, reference lobpcg.h:63-83):
 * 0 = ran; 1 = parameters rejected with the reference's message, outputs untouched as in the reference;
 * 2 = run-time failure: alg->converged = 0, alg->iter = passes done, eigVals / resNorm = NaN */
int lb2_last_status(void);
/* Ritz values (first neig) and residual norms (first nres) of the last pass, as doubles, without downloading X */
int lb2_solver_results(lb2_solver *s, double *eig, int neig, double *res, int nres);

/* ---- (3) kernels on device pointers -------------------------------------------------------------- */
#define LB2_DECLARE_KERNELS(P)                                                                            \
  /* G(ma x mb) = A^H B; upper!=0: Hermitian product (ma==mb), upper tiles computed, result mirrored */    \
  int lb2_##P##_gram(lb2_ctx *ctx, int64_t n, int ma, int mb, const void *A, int64_t lda, const void *B,   \
                     int64_t ldb, void *G, int ldg, int upper);                                            \
  /* column-block products of the cached-Gram pass (reference design docs/plans/2026-04-08-soft-locking-merge-design.md:48-61): \
   * G0[0:m,0:nw] = S^H W0 and, when W1 != NULL, G1[0:m,0:nw] = S^H W1 in one pass over S.  tri_c0 >= 0: rows tri_c0.. of the   \
   * results form a Hermitian nw x nw block whose strictly-lower part may be left unwritten. */                \
  int lb2_##P##_gram_cols(lb2_ctx *ctx, int64_t n, int m, int nw, const void *S, int64_t lds, const void *W0,  \
                          int64_t ldw0, void *G0, int ldg0, const void *W1, int64_t ldw1, void *G1, int ldg1,  \
                          int tri_c0);                                                                     \
  /* Out(n x nb) = alpha S(n x kd) C(kd x nb) + beta Out; alpha,beta point to one scalar each */           \
  int lb2_##P##_tall_nn(lb2_ctx *ctx, int64_t n, int kd, int nb, const void *alpha, const void *S,         \
                        int64_t lds, const void *C, int ldc, const void *beta, void *Out, int64_t ldo);    \
  /* W = AX - BX diag(lambda) (W may be NULL), sumsq[j] = ||W[:,j]||^2 (may be NULL); lambda, sumsq real */ \
  int lb2_##P##_residual(lb2_ctx *ctx, int64_t n, int nc, const void *AX, int64_t ldax, const void *BX,    \
                         int64_t ldbx, const void *lambda, void *W, int64_t ldw, void *sumsq);             \
  int lb2_##P##_col_sumsq(lb2_ctx *ctx, int64_t n, int nc, const void *X, int64_t ldx, void *sumsq);       \
  /* X[i,j] = uniform[-.5,.5) from splitmix64(seed, j*n_global + row0 + i) */                              \
  int lb2_##P##_fill_uniform(lb2_ctx *ctx, int64_t n, int nc, void *X, int64_t ldx, uint64_t seed,         \
                             int64_t n_global, int64_t row0);                                              \
  int lb2_##P##_spmm_stencil(lb2_ctx *ctx, int64_t gx, int64_t gy, int64_t gz, double cdiag, double coff,  \
                             const void *potential, int nc, const void *X, int64_t ldx, void *Y,           \
                             int64_t ldy);                                                                 \
  /* same with z-halo planes (column stride halo_ld) below z=0 / above z=gz-1; NULL = Dirichlet */        \
  int lb2_##P##_spmm_stencil_halo(lb2_ctx *ctx, int64_t gx, int64_t gy, int64_t gz, double cdiag,          \
                                  double coff, const void *potential, const void *halo_lo,                 \
                                  const void *halo_hi, int64_t halo_ld, int nc, const void *X, int64_t ldx, \
                                  void *Y, int64_t ldy);                                                   \
  int lb2_##P##_spmm_csr(lb2_ctx *ctx, int64_t n, const int64_t *rowptr, const int32_t *col,               \
                         const void *val, int nc, const void *X, int64_t ldx, void *Y, int64_t ldy);       \
  int lb2_##P##_spmm_diag(lb2_ctx *ctx, int64_t n, const void *diag, int nc, const void *X, int64_t ldx,   \
                          void *Y, int64_t ldy);

LB2_DECLARE_KERNELS(s)
LB2_DECLARE_KERNELS(d)
LB2_DECLARE_KERNELS(c)
LB2_DECLARE_KERNELS(z)

/* ---- multi-GPU (one process per GPU; rows of all block vectors are partitioned) -------------------- */
/* rank-local z-slab [z0, z0+gz_local) of a gx*gy*gz_global stencil; alg->size stays the GLOBAL row count */
void *lb2_op_stencil_slab(char prefix, int64_t gx, int64_t gy, int64_t gz_local, int64_t gz_global, int64_t z0,
                          double cdiag, double coff, const void *potential_local_host);
/* the solver keeps every tall block in one device arena; neighbours map it through CUDA IPC and the stencil
 * kernel reads their boundary planes directly over NVLink */
/* row block [row0, row0 + n_local) of a CSR matrix (GLOBAL column indices; rowptr_local starts at 0): equal blocks on every
 * rank, couplings at most into the two neighbouring blocks, whose rows the kernel reads in place from the neighbours'
 * arenas (lb2_solver_set_peers).  Its spectrum bound covers the local rows only. */
void *lb2_op_csr_slab(char prefix, int64_t n_global, int64_t row0, int64_t n_local, const int64_t *rowptr_local,
                      const int32_t *col_global, const void *val_host);
/* rank-local part of the BdG operator: both fields are split by the same z-slabs, local rows = [u slab ; v slab] */
void *lb2_op_bdg_slab(char prefix, int64_t gx, int64_t gy, int64_t gz_local, int64_t gz_global, int64_t z0, double cdiag,
                      double coff, double shift, double d_re, double d_im);
/* neighbour data for a stand-alone lb2_op_apply of a row-block operator: boundary planes (stencil slab) or whole blocks
 * (CSR row block) below / above, column stride ld; NULL = no neighbour.  A solver sets these itself from the peer arenas. */
int lb2_op_set_halo(void *linop, const void *lo, const void *hi, int64_t ld);
int lb2_solver_arena(lb2_solver *s, void **ptr, size_t *bytes);
int lb2_solver_set_peers(lb2_solver *s, const void *lo_arena, const void *hi_arena);
int lb2_comm_unique_id(void *out128, const char *nccl_lib_path);   /* rank 0; broadcast by the launcher */
int lb2_ctx_attach_comm(lb2_ctx *ctx, int rank, int size, const void *unique_id128, const char *nccl_lib_path);
int lb2_ctx_detach_comm(lb2_ctx *ctx);
int lb2_comm_allreduce(lb2_ctx *ctx, void *dev_buf, size_t count, int is_double);
/* Several GPUs inside ONE call of the reference entry points (one process, one worker thread per device, peer access for
 * the halos, an NCCL communicator kept per process): n > 1 lets <p>_lobpcg / <p>_ilobpcg partition the rows of a solve over
 * up to n devices when every operator is a built-in stencil / BdG / CSR / diagonal / polynomial-preconditioner operator and
 * the grid splits into equal slabs; otherwise the call runs on one GPU.  0 = take the count from the environment
 * (LB2_GPUS = N | all; default 1).  Returns the previous setting. */
int lb2_set_num_gpus(int n);
int lb2_last_num_gpus(void); /* how many GPUs the last call really used */
int lb2_ipc_get_handle(void *dev_ptr, void *out64);
void *lb2_ipc_open_handle(const void *in64);
int lb2_ipc_close_handle(void *mapped_ptr);

/* host-only self-check of the work-list Gram schedule for a shape (no device needed): 0 = every 8x8 block of every
 * tile owned once, rows partitioned exactly, contiguous items per CTA.  stats[4] (may be NULL): items, busiest CTA /
 * mean CTA cost, issued / needed DMMA blocks, tiles. */
int lb2_gram_wl_plan_check(int ma, int mb, int upper, int64_t n, int ncta, int bk, double *stats);
/* same for the column-block schedule of lb2_d_gram_cols (nprod = 1 or 2 products): additionally every output entry that is
 * not strictly below the diagonal of the Hermitian block is covered exactly once.  stats[4]: items, busiest CTA / mean,
 * tiles, computed tile area / full rectangular area. */
int lb2_gram_wl_cols_plan_check(int m, int nw, int nprod, int tri_c0, int64_t n, int ncta, int bk, double *stats);
/* host-only model of the operand sharing of that schedule: fraction of the panel requests that are distinct (must come from
 * DRAM) when all CTAs advance at their tiles' cost rate; phase = 1: phase-aligned cyclic walk of the pieces (the default) */
int lb2_gram_wl_plan_sharing(int ma, int mb, int upper, int64_t n, int ncta, int bk, int phase, int window_chunks,
                             int samples, double *share);

const char *lb2_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LOBPCG_B200_H */

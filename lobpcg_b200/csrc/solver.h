// lobpcg_b200/csrc/solver.h — device-resident LOBPCG driver (definite and indefinite), templated on the
// scalar type; the C ABI in capi.cu instantiates it for s/d/c/z.
//
// Mirrors the reference driver's state machine (src/core/lobpcg_impl.inc:60-248,
// src/core/ilobpcg_impl.inc:54-282) but not its memory plan:
//
//   reference (host)                               here (HBM)
//   S = [X|P|W] n x 3k, wrk1..4 (2+3+1+2) n k      two slabs n x 3k used as ping-pong: projections write the
//   AX n k                  => 12 n k scalars      next [X|P|W] out of place (no wrk4 + memcpy), AS = [AX|AP|AW]
//                                                  n x 3k, two n x k scratch blocks        => 11 n k scalars
//   P, W always k columns, compacted by memmove     P, W are only ever produced for the active (unconverged)
//   at the top of the next pass (:139-145)          columns, directly at their compacted position
//
// Host round trips per pass: potrf info + condition number (1 sync), eigensolver info (1), residual norms
// (1); ortho passes add one per SVQB sweep.  Everything else is enqueued on ctx->stream.
#pragma once
#include <stdint.h>
#include <vector>
#include "common.cuh"
#include "context.h"

namespace lb2 {

// ---- raw C layouts shared with include/lobpcg.h ---------------------------------------------------------
struct LinOpCtxRaw {
  void* data;
  size_t data_size;
};
struct LinOpRaw {
  uint64_t rows, cols;
  void (*matvec)(const LinOpRaw*, void*, void*);
  void (*cleanup)(LinOpCtxRaw*);
  LinOpCtxRaw* ctx;
};

template <typename T>
struct State {  // == <p>_lobpcg_t (reference lobpcg.h:13-55)
  T *S, *Cx, *Cp, *AX, *AS, *BS;
  real_t<T>*eigVals, *resNorm;
  int8_t* signature;
  T *wrk1, *wrk2, *wrk3, *wrk4;
  real_t<T>*rr_D, *rr_eigvals;
  T *rr_tau, *rr_VR;
  int8_t* rr_sig;
  uint64_t* rr_indices;
  T* rr_ggev;
  int8_t implicit_product_update, verbosity;
  uint64_t iter, nev, converged, size, sizeSub, maxIter;
  real_t<T> tol;
  LinOpRaw *A, *B, *T_;
};

// ---- built-in operators (tag lives at the head of ctx->data) -------------------------------------------
constexpr uint64_t kOpMagic = 0x4C42324F50455221ULL;  // "LB2OPER!"
enum OpKind { OP_STENCIL = 0, OP_CSR = 1, OP_DIAG = 2, OP_BDG = 3, OP_CHEB = 4, OP_DENSE = 5, OP_DEVICE = 6 };
// OP_DEVICE: caller-supplied block operator on DEVICE pointers (lb2_matmat_fn of include/lobpcg_b200.h)
typedef int (*DeviceMatmat)(void* user, int ncols, const void* X, int64_t ldx, void* Y, int64_t ldy, void* cuda_stream);

struct BuiltinOp {
  uint64_t magic;
  int kind;
  char prefix;
  int device;
  int64_t n;        // local rows
  int64_t n_global; // == n on a single GPU
  int64_t row0;     // first global row of this rank's slab
  // stencil / bdg
  int64_t gx, gy, gz;
  double cdiag, coff, shift, dre, dim;
  void* potential;  // device, real
  const void* halo_lo;
  const void* halo_hi;
  int64_t halo_ld;
  // csr
  int64_t* rowptr;
  int32_t* col;
  void* val;
  int64_t nnz;
  int from_csr;     // stencil operator recognised from CSR input (capi.cu: detect_stencil)
  int csr_halo;     // half-width H of the near-diagonal window of the windowed CSR kernel (0: plain kernel; capi.cu: csr_window_halo)
  // diag
  void* diag;
  // dense n x n matrix (column-major, leading dimension n), device
  void* dense;
  // upper bound of the spectrum (Gershgorin), computed on the host at construction; 0 = unknown
  double spec_hi;
  // chebyshev preconditioner T = p(A): `degree` applications of `inner` per block apply, spectrum window [lo, hi]
  const LinOpRaw* inner;
  int cheb_degree;
  double cheb_lo, cheb_hi;
  int cheb_mixed;        // 1: inside the solver the polynomial is evaluated in the lower precision (float for d, c32 for z)
  void* potential_lo;    // float copy of the inner stencil's potential for that evaluation (device, may be null)
  // caller's device block operator (OP_DEVICE)
  DeviceMatmat dev_fn;
  void* dev_user;
  // back pointer for host matvec shim
  LinOpRaw* self;
};

inline const BuiltinOp* builtin_of(const LinOpRaw* op) {
  if (!op || !op->ctx || !op->ctx->data || op->ctx->data_size != sizeof(BuiltinOp)) return nullptr;
  const BuiltinOp* b = (const BuiltinOp*)op->ctx->data;
  return b->magic == kOpMagic ? b : nullptr;
}

// Y = Op X for a built-in operator (device block vectors)
template <typename T>
int apply_builtin(lb2_ctx* ctx, const BuiltinOp* b, int nc, const T* X, int64_t ldx, T* Y, int64_t ldy);

// OP_CHEB: Y = p(A) X, cheb_degree steps of the Chebyshev iteration for A y = x on [lo, hi] from y = 0
// (y_1 = x / theta; r_j = r_{j-1} - A d_{j-1}; d_j = rho_j rho_{j-1} d_{j-1} + (2 rho_j / delta) r_j; y += d_j);
// implemented in solver.cu (cheb_apply): fused into the stencil kernel's epilogue when A is a built-in stencil.

// ---- phase statistics ----------------------------------------------------------------------------------
enum Phase { PH_SPMM = 0, PH_GRAM, PH_TALLNN, PH_RESID, PH_SMALL, PH_COMM, PH_OTHER, PH_COUNT };

struct SolverBase {
  virtual ~SolverBase() {}
  virtual int prepare() = 0;  // validate + allocate (row-partitioned runs exchange arena handles between prepare and init)
  virtual int init() = 0;
  virtual void arena_info(void** p, size_t* bytes) = 0;
  virtual void set_peers(const void* lo, const void* hi) = 0;
  virtual int step(int max_steps) = 0;
  virtual int finish() = 0;
  virtual void state(uint64_t* iter, uint64_t* conv, int* use_ortho) = 0;
  virtual int results(double* eig, int neig, double* res, int nres) = 0;  // host copies of the last pass
  virtual int set_option(const char* key, int value) = 0;   // "gram_cache", "gram_cache_period", "force_ortho", "debug_min_conv", "implicit"
  virtual double info(const char* key) = 0;                  // "gram_cache_refreshes", "gram_cache_monitor", "arena_bytes", ...
  virtual void write_failure_state() = 0;                    // alg->converged = 0, iter, eigVals/resNorm = NaN (run-time failure)
  bool param_error = false;   // prepare() rejected the parameters with the reference's own message (outputs stay untouched, as there)
  uint64_t cache_refreshes = 0;
  double phase_ms[PH_COUNT] = {0};
  double phase_work[PH_COUNT] = {0};   // algorithmic flops (gram, tall_nn) or bytes (spmm, residual)
  uint64_t phase_calls[PH_COUNT] = {0};
  uint64_t device_seed = 0;
  bool use_device_x0 = false;
  // device-pointer fast path: X0 is read from / the eigenvectors are written to DEVICE blocks of this rank's
  // rows (n_local x sizeSub, column-major, leading dimension n_local) instead of alg->S
  const void* dev_x0 = nullptr;
  void* dev_xout = nullptr;
};

SolverBase* make_solver(lb2_ctx* ctx, char prefix, void* alg, int indefinite);

// multi-GPU: sum `count` reals across ranks in place on ctx->stream (no-op when ctx->comm == nullptr)
int allreduce_sum(lb2_ctx* ctx, void* buf, size_t count, bool is_double);
int comm_rank(lb2_ctx* ctx);
int comm_size(lb2_ctx* ctx);

}  // namespace lb2

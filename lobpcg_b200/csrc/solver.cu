// lobpcg_b200/csrc/solver.cu — device-resident definite LOBPCG driver (see solver.h for the memory plan).
//
// Reference being mirrored (same state machine, same constants):
//   driver            src/core/lobpcg_impl.inc:60-248
//   initial RR        src/rayleigh/rayleigh_ritz_impl.inc:37-100
//   modified RR       src/rayleigh/rayleigh_ritz_modified_impl.inc:42-273   (tol_skip = 5e-3, sticky useOrtho)
//   ortho_drop/svqb   src/ortho/ortho_drop_impl.inc:43-125, src/ortho/svqb_impl.inc:48-106 (3 x 3 sweeps)
//   residual / norms  src/residual/residual_impl.inc:32-99   (2-norm even when B != NULL)
//   estimate_norm     src/residual/estimate_norm_impl.inc:38-57  (10 power steps)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <type_traits>
#include <vector>
#include <inttypes.h>

#include "common.cuh"
#include "context.h"
#include "kernels.h"
#include "smalldense.h"
#include "solver.h"
#include "../../include/lobpcg_b200.h"

namespace lb2 {

// -------------------------------------------------------------------------------------------------------
static StencilDesc stencil_desc(const BuiltinOp* b) {
  StencilDesc d;
  d.gx = (int)b->gx; d.gy = (int)b->gy; d.gz = (int)b->gz;
  d.cdiag = b->cdiag; d.coff = b->coff; d.shift = b->shift;
  d.potential = b->potential;
  d.halo_lo = b->halo_lo; d.halo_hi = b->halo_hi; d.halo_ld = b->halo_ld;
  d.bdg = (b->kind == OP_BDG) ? 1 : 0;
  d.dre = b->dre; d.dim = b->dim;
  return d;
}

// applyA(nc, D, AD): AD = A D.  fusedStep(nc, Din, Dout, Rin, ldrin, Rout, Y, c1, c2, write_r): one whole step inside the
// operator kernel (stencil epilogue); returns -100 when the inner operator has no fused form.
template <typename T, typename ApplyA, typename FusedStep>
int cheb_apply(lb2_ctx* ctx, const BuiltinOp* b, int64_t n, int nc, const T* X, int64_t ldx, T* Y, int64_t ldy, T* Rw,
               T* D0, T* D1, T* ADw, int64_t ldw, ApplyA&& applyA, FusedStep&& fusedStep) {
  using R = real_t<T>;
  const double theta = 0.5 * (b->cheb_hi + b->cheb_lo), delta = 0.5 * (b->cheb_hi - b->cheb_lo);
  const double sigma = theta / delta;
  double rho_old = 1.0 / sigma;
  T* Dcur = D0;
  T* Dnext = D1;
  bool fused = true;
  if (int rc = cheb_init<T>(ctx, n, nc, X, ldx, Dcur, Y, ldy, ldw, (R)(1.0 / theta))) return rc;
  for (int j = 1; j <= b->cheb_degree; j++) {
    const double rho = 1.0 / (2.0 * sigma - rho_old);
    const R c1 = (R)(rho * rho_old), c2 = (R)(2.0 * rho / delta);
    const T* Rin = (j == 1) ? X : Rw;
    const int64_t ldrin = (j == 1) ? ldx : ldw;
    const bool wr = j < b->cheb_degree;
    int rc = fused ? fusedStep(nc, Dcur, Dnext, Rin, ldrin, Rw, Y, c1, c2, wr) : -100;
    if (rc == -100) {
      fused = false;
      if ((rc = applyA(nc, Dcur, ADw))) return rc;
      rc = cheb_update<T>(ctx, n, nc, ADw, Rin, ldrin, Rw, Dcur, Dnext, ldw, Y, ldy, c1, c2, wr);
    }
    if (rc) return rc;
    std::swap(Dcur, Dnext);
    rho_old = rho;
  }
  return 0;
}

template <typename T>
int apply_builtin(lb2_ctx* ctx, const BuiltinOp* b, int nc, const T* X, int64_t ldx, T* Y, int64_t ldy) {
  if (b->prefix != Sc<T>::prefix) {
    fprintf(stderr, "lobpcg_b200: operator built for type '%c' applied to type '%c'\n", b->prefix, Sc<T>::prefix);
    return -1;
  }
  switch (b->kind) {
    case OP_STENCIL:
    case OP_BDG:
      return spmm_stencil<T>(ctx, stencil_desc(b), nc, X, ldx, Y, ldy);
    case OP_CSR:
      if (b->n != b->n_global) {   // row block of a partitioned matrix: neighbour blocks set by Solver::localize
        CsrHalo h;
        h.lo = b->halo_lo; h.hi = b->halo_hi; h.ld_lo = h.ld_hi = b->halo_ld;
        return spmm_csr<T>(ctx, b->n, b->rowptr, b->col, (const T*)b->val, nc, X, ldx, Y, ldy, &h, b->nnz);
      }
      if (b->csr_halo >= 2 && ctx->csr_window > 0) {   // banded matrix: shared-memory X window (spmm.cu: csr_win_kernel), opt-in — measured slower than the plain kernel
        const int rc = spmm_csr_window<T>(ctx, b->n, b->rowptr, b->col, (const T*)b->val, nc, X, ldx, Y, ldy, b->csr_halo);
        if (rc != -100) return rc;
      }
      return spmm_csr<T>(ctx, b->n, b->rowptr, b->col, (const T*)b->val, nc, X, ldx, Y, ldy, nullptr, b->nnz);
    case OP_DIAG:
      return spmm_diag<T>(ctx, b->n, (const real_t<T>*)b->diag, nc, X, ldx, Y, ldy);
    case OP_DENSE:   // plain library GEMM (cuBLAS): Y = A X with A dense n x n
      if (b->n > INT32_MAX || ldx > INT32_MAX || ldy > INT32_MAX) return -2;
      return sd_gemm<T>(ctx, 'N', (int)b->n, nc, (int)b->n, (const T*)b->dense, (int)b->n, X, (int)ldx, Y, (int)ldy);
    case OP_DEVICE: {   // the caller's own kernels, enqueued on the solver stream (SURVEY §8b "foreign operators")
      if (!b->dev_fn) return -1;
      const int rc = b->dev_fn(b->dev_user, nc, X, ldx, Y, ldy, (void*)ctx->stream);
      if (rc) fprintf(stderr, "lobpcg_b200: device operator callback returned %d\n", rc);
      return rc;
    }
    case OP_CHEB: {   // stand-alone apply (outside a solver): temporary workspace
      const BuiltinOp* in = builtin_of(b->inner);
      if (!in) {
        fprintf(stderr, "lobpcg_b200: chebyshev preconditioner needs a built-in inner operator outside the solver\n");
        return -1;
      }
      T* w = nullptr;
      const size_t blk = (size_t)b->n * nc;
      LB2_CUDA_OK(cudaMalloc(&w, sizeof(T) * 4 * blk));
      const bool can_fuse = (in->kind == OP_STENCIL) && !getenv("LB2_NO_CHEB_FUSE");
      int rc = cheb_apply<T>(
          ctx, b, b->n, nc, X, ldx, Y, ldy, w, w + blk, w + 2 * blk, w + 3 * blk, b->n,
          [&](int c, const T* D, T* AD) { return apply_builtin<T>(ctx, in, c, D, b->n, AD, b->n); },
          [&](int c, const T* Din, T* Dout, const T* Rin, int64_t ldrin, T* Rout, T* Yacc, real_t<T> c1, real_t<T> c2,
              bool wr) {
            if (!can_fuse) return -100;
            ChebEpilogue<T> ep;
            ep.rin = Rin; ep.ldrin = ldrin; ep.rout = Rout; ep.dout = Dout; ep.ldw = b->n; ep.c1 = c1; ep.c2 = c2;
            ep.write_r = wr ? 1 : 0;
            return spmm_stencil_cheb<T>(ctx, stencil_desc(in), c, Din, b->n, Yacc, ldy, ep);
          });
      cudaStreamSynchronize(ctx->stream);
      cudaFree(w);
      return rc;
    }
  }
  return -1;
}
template int apply_builtin<float>(lb2_ctx*, const BuiltinOp*, int, const float*, int64_t, float*, int64_t);
template int apply_builtin<double>(lb2_ctx*, const BuiltinOp*, int, const double*, int64_t, double*, int64_t);
template int apply_builtin<c32>(lb2_ctx*, const BuiltinOp*, int, const c32*, int64_t, c32*, int64_t);
template int apply_builtin<c64>(lb2_ctx*, const BuiltinOp*, int, const c64*, int64_t, c64*, int64_t);

// -------------------------------------------------------------------------------------------------------
struct Timers {
  struct Rec { int ph; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  size_t used = 0;
  cudaStream_t st = nullptr;
  void begin(int ph) {
    if (used == recs.size()) {
      Rec r; r.ph = ph;
      cudaEventCreate(&r.a); cudaEventCreate(&r.b);
      recs.push_back(r);
    }
    recs[used].ph = ph;
    cudaEventRecord(recs[used].a, st);
  }
  void end() { cudaEventRecord(recs[used].b, st); used++; }
  void collect(double* ms) {  // call after a stream synchronize
    for (size_t i = 0; i < used; i++) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, recs[i].a, recs[i].b) == cudaSuccess) ms[recs[i].ph] += t;
    }
    used = 0;
  }
  ~Timers() { for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); } }
};

#define LB2_TRY(expr)                                                                    \
  do {                                                                                   \
    int _rc = (expr);                                                                    \
    if (_rc != 0) {                                                                      \
      fprintf(stderr, "lobpcg_b200: %s failed (%d) at %s:%d\n", #expr, _rc, __FILE__, __LINE__); \
      return _rc;                                                                        \
    }                                                                                    \
  } while (0)

template <typename T> struct EpsTol;  // EPS_TOL of src/core/lobpcg_{s,d,c,z}.c
template <> struct EpsTol<float>  { static constexpr double v = 1e-5; };
template <> struct EpsTol<double> { static constexpr double v = 1e-12; };
template <> struct EpsTol<c32>    { static constexpr double v = 1e-5; };
template <> struct EpsTol<c64>    { static constexpr double v = 1e-12; };

template <typename T>
class Solver : public SolverBase {
  using R = real_t<T>;
  static constexpr bool kDouble = sizeof(R) == 8;
  static constexpr int kCplx = Sc<T>::cplx ? 2 : 1;

 public:
  Solver(lb2_ctx* c, State<T>* a, bool indefinite) : ctx(c), alg(a), indef(indefinite) { tm.st = c->stream; }
  ~Solver() override { release(); }

  int init() override;
  int step(int max_steps) override;
  int finish() override;
  void arena_info(void** p, size_t* bytes) override { *p = arena; *bytes = arena_bytes; }
  void set_peers(const void* lo, const void* hi) override { peer_lo = (const char*)lo; peer_hi = (const char*)hi; }
  int prepare() override;
  // minimal set-up for the helper API: n rows, blocks of at most kmax columns, no validation of 3k <= n
  int helper_setup(int64_t rows, int kmax, const LinOpRaw* A, const LinOpRaw* B) {
    ng = n = rows; row0 = 0; k = kmax > 0 ? kmax : 1; nev = k;
    opA = A; opB = B; opT = nullptr;
    helper_mode = true;
    LB2_CUDA_OK(cudaSetDevice(ctx->device));
    if (sd_init(ctx)) return 1;
    LB2_TRY(alloc());
    LB2_CUDA_OK(cudaMemsetAsync(Scal, 0, sizeof(R) * 16, ctx->stream));
    prepared = true;
    return 0;
  }
  int up(T* dev, const T* host, int64_t rows, int cols) {
    if (rows <= 0 || cols <= 0) return 0;
    LB2_CUDA_OK(cudaMemcpyAsync(dev, host, sizeof(T) * (size_t)rows * cols, cudaMemcpyHostToDevice, ctx->stream));
    LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  int down(T* host, const T* dev, int64_t rows, int cols) {
    if (rows <= 0 || cols <= 0) return 0;
    LB2_CUDA_OK(cudaMemcpyAsync(host, dev, sizeof(T) * (size_t)rows * cols, cudaMemcpyDeviceToHost, ctx->stream));
    LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  double eps_override = -1;   // helper API: caller-supplied eps_ortho / eps_drop (the drivers always pass EPS_TOL)
  R eps_tol() const { return eps_override >= 0 ? (R)eps_override : (R)EpsTol<T>::v; }
  int results(double* eig, int neig, double* res, int nres) override {
    for (int i = 0; i < neig && i < k; i++) eig[i] = (double)hEig[i];
    for (int i = 0; i < nres && i < nev; i++) res[i] = (double)hRes[i];
    return 0;
  }
  void state(uint64_t* it, uint64_t* cv, int* uo) override {
    if (it) *it = iter;
    if (cv) *cv = conv;
    if (uo) *uo = useOrtho;
  }
  int set_option(const char* key, int value) override {
    if (!key) return -1;
    if (!strcmp(key, "gram_cache")) {
      if (prepared) return -2;   // decides the arena layout
      gram_cache = value != 0;
      gram_cache_forced = true;
    } else if (!strcmp(key, "gram_cache_period")) cache_period = value > 0 ? value : 1;
    else if (!strcmp(key, "force_ortho")) force_ortho = value;
    else if (!strcmp(key, "indef_geev")) indef_geev = value;
    else if (!strcmp(key, "debug_min_conv")) debug_min_conv = value;
    else return -1;
    return 0;
  }
  double info(const char* key) override {
    if (!key) return -1;
    if (!strcmp(key, "gram_cache")) return gram_cache ? 1 : 0;
    if (!strcmp(key, "gram_cache_refreshes")) return (double)cache_refreshes;
    if (!strcmp(key, "gram_cache_monitor")) return cache_monitor;
    if (!strcmp(key, "gram_cache_monitor_max")) return cache_monitor_max;
    if (!strcmp(key, "general_rr_calls")) return (double)general_rr_calls;
    if (!strcmp(key, "quality5_passes")) return (double)quality5_passes;
    if (!strcmp(key, "arena_bytes")) return (double)arena_bytes;
    if (!strcmp(key, "arena_columns")) return n > 0 ? (double)tall_bytes / ((double)n * sizeof(T)) : 0.0;   // block-vector columns
    return -1;
  }
  void write_failure_state() override {
    if (!alg) return;
    const R nan = std::numeric_limits<R>::quiet_NaN();
    alg->converged = 0;
    alg->iter = iter;
    if (alg->eigVals) for (uint64_t i = 0; i < alg->sizeSub; i++) alg->eigVals[i] = nan;
    if (alg->resNorm) for (uint64_t i = 0; i < alg->nev && i < alg->sizeSub; i++) alg->resNorm[i] = nan;
  }

 public:  // (the host-buffer helper API at the end of this file drives the same building blocks)
  lb2_ctx* ctx;
  State<T>* alg;
  bool indef;
  Timers tm;
  int64_t n = 0, ng = 0, row0 = 0;
  // partial sums are combined across ranks only when the rows are partitioned (an unpartitioned solve on a context that
  // has a communicator attached is a replica and must not be summed)
  bool reduce() const { return ctx->comm != nullptr && n != ng; }
  int nseg = 1;   // row-partitioned BdG operator: local rows are two runs (the u and the v field), each a z-slab of its field
  // local run s = rows [s n/nseg, (s+1) n/nseg) of this rank = global rows seg_global(s) + [0, n/nseg)
  int64_t seg_len() const { return n / nseg; }
  int64_t seg_global(int s) const { return (int64_t)s * (ng / nseg) + row0; }
  int fill_rows(T* X, int nc, uint64_t seed) {   // uniform block with the GLOBAL counters of this rank's rows
    for (int sg = 0; sg < nseg; sg++)
      LB2_TRY(fill_uniform<T>(ctx, seg_len(), nc, X + sg * seg_len(), n, seed, ng, seg_global(sg)));
    return 0;
  }
  int k = 0, nev = 0;
  void* arena = nullptr;
  size_t arena_bytes = 0, tall_bytes = 0;   // whole allocation / the n-proportional part (block vectors)
  const char *peer_lo = nullptr, *peer_hi = nullptr;  // neighbours' arena bases (CUDA-IPC mappings), or null
  T* slab[2] = {nullptr, nullptr};
  int cur = 0;
  T *AS = nullptr, *wA = nullptr, *wB = nullptr;
  T *G = nullptr, *GA = nullptr, *DinvR = nullptr, *Z = nullptr, *Tmp = nullptr, *Cx = nullptr, *Cp = nullptr,
    *Q = nullptr, *Tau = nullptr;
  R *D = nullptr, *Lam = nullptr, *Eig = nullptr, *Sums = nullptr, *Scal = nullptr;
  int* Count = nullptr;
  R* Theta = nullptr;      // indefinite RR: all m Ritz values, signature-sorted
  int8_t* dSig = nullptr;  // indefinite RR: signatures (+1/-1), sorted
  R* hbuf = nullptr;  // pinned
  size_t hbuf_bytes = 0;
  T *hX = nullptr, *hY = nullptr;  // host staging for host-callback operators
  int np = 0, nw = 0;
  int sig_len = 0;
  int useOrtho = 0;
  uint64_t iter = 0, conv = 0;
  R ANorm = 0, BNorm = 1;
  std::vector<R> hEig, hRes;
  const LinOpRaw *opA = nullptr, *opB = nullptr, *opT = nullptr;
  bool inited = false, done = false, prepared = false;
  // ---- cached Gram blocks (reference design docs/plans/2026-04-08-soft-locking-merge-design.md:48-61) ----
  // Per pass only the W columns of S^H B S and S^H A S are contracted over n (gram_cols); the [X P] blocks follow from
  // the previous pass's small matrices: [X' P'] = S [Cx | Cp_act]  =>  [X' P']^H B [X' P'] = C^H G C, same for A.
  // AS then only holds [AX | AW] (2k columns; the legacy path keeps [AX | AP | AW]).
  bool gram_cache_forced = false;
  bool gram_cache = true;        // option "gram_cache" / LB2_GRAM_CACHE=0: legacy pass (both Grams recomputed in full)
  bool helper_mode = false;      // host-buffer helper API: one-shot calls, full-size workspaces
  int as_cols = 0;               // columns of AS
  T *Graw = nullptr, *GAraw = nullptr;   // assembled S^H B S / S^H A S of the current pass (before factorisation)
  T *Gc = nullptr, *GAc = nullptr;       // cached [X P] blocks, leading dimension cache_mxp
  T *Gw = nullptr, *GAw = nullptr;       // W columns (m x nw, ld m) from gram_cols
  T* Ccat = nullptr;                     // [Cx | Cp_act] (m x (k + n_act), ld m)
  bool cache_ok = false, cache_has_b = false;
  int cache_mxp = 0, since_refresh = 0;
  int cache_period = 64;         // option "gram_cache_period": recompute the [X P] blocks from the tall vectors every .. passes
  double cache_monitor = 0;      // last drift monitor value (max |x^H B x - 1|, |x^H A x - theta| / ||A|| over the nev columns)
  double cache_monitor_max = 0;
  // general indefinite Rayleigh-Ritz (rr_indef_general)
  int indef_geev = 0;            // option "indef_geev": 1 = always take the GEEV route (testing); default: only when S^H A S is not positive definite
  int last_quality = 1;          // quality_flag of the last indefinite RR: 1 or 5 (indefinite_rr_modified_impl.inc:236-251)
  uint64_t general_rr_calls = 0, quality5_passes = 0;
  void* geev_ws = nullptr;       // complex scratch of Xgeev (lazy)
  int64_t* geev_piv = nullptr;
  int8_t* geev_sig = nullptr;
  T* CxAcc = nullptr;            // accurate Cx of a quality-5 pass (lazy)
  T* SigV = nullptr;             // indefinite solver: V^H B V of this pass's ortho_indefinite = [X P] block of S^H B S
  int sigv_cols = 0;
  // testing / measurement switches (lb2_solver_set_option)
  int force_ortho = 0;           // 1: run every pass in the ortho branch (useOrtho = 1 from the first pass)
  int debug_min_conv = 0;        // soft-lock at least this many leading columns regardless of their residuals (timing only)

  T* Xp() { return slab[cur]; }
  T* col(T* base, int64_t c) { return base + c * n; }

  void release();
  int alloc();
  int qr_info_pending = 0;   // cp_from_z: the info words of geqrf / orgqr travel with the next host synchronisation
  int sync() {
    LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    tm.collect(phase_ms);
    if (qr_info_pending) {
      qr_info_pending = 0;
      const int* qi = reinterpret_cast<const int*>(hbuf + 4 * (size_t)k + 60);
      if (qi[0] != 0 || qi[1] != 0) {
        fprintf(stderr, "rayleigh_ritz_modified: QR of the P basis failed (geqrf info=%d, orgqr info=%d)\n", qi[0], qi[1]);
        return -1;
      }
    }
    return 0;
  }
  int d2h(void* dst, const void* src, size_t bytes) {
    LB2_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
  }

  int apply(const LinOpRaw* op, int nc, const T* X, T* Y);
  int gram_ar(int ma, int mb, const T* A, const T* B, T* Gout, int upper);
  int nn(int kd, int nb, T alpha, const T* S, const T* C, int ldc, T beta, T* Out);
  int sumsq_total(int nc, const T* X, R* out_dev);  // out_dev[0] = ||X||_F^2 (all ranks)
  int resid(int nc, const T* AXp, const T* BXp, const R* lam, T* Wp, R* ss) {
    tm.begin(PH_RESID);
    int rc = residual<T>(ctx, n, nc, AXp, n, BXp, n, lam, Wp, n, ss);
    tm.end();
    phase_work[PH_RESID] += (Wp ? 3.0 : 2.0) * (double)n * nc * sizeof(T);
    phase_calls[PH_RESID]++;
    return rc;
  }
  int estimate_norm(const LinOpRaw* op, uint64_t seed, R* out);
  int gram_self_B(int m, T* S, T* Gout);             // S^H B S (mirrored)
  int chol_transform(int m, int* bad);               // G -> D, R (in G), DinvR ; bad=1 if potrf failed or rcond small
  int rr_initial();
  int rr_modified(int m);
  int update_gram_cache(int m, int nconv);
  double monitor_threshold() const { return kDouble ? 5e-12 : 2e-5; }
  int cp_from_z(int m, const T* Zm, T* VQ);           // VQ (m x k) = Z_perp Q
  int svqb(T* U, int nu, R tau, bool drop, int* nret);
  int localize(const BuiltinOp*& b, BuiltinOp& local, const T* X);
  int localize_bytes(const BuiltinOp*& b, BuiltinOp& local, const void* X, size_t elem);
  int apply_cheb_mixed(const BuiltinOp* b, const BuiltinOp* in, int nc, const T* X, T* Y);
  int ortho_drop(T* U, int nu, T* V, int nv, int* nret, bool indefinite = false);
  int rr_indef(int m, int from_col, bool initial);
  int rr_indef_general(int m, bool initial);
  int ortho_indef_mat(int m, int nu, int nv, T* U, const T* V, const T* mat);
  int svqb_mat_dev(int m, int nu, T* U, const T* mat, R tau);
  int residual_pass(bool initial, const T* Xacc = nullptr);
  const T *res_AX = nullptr, *res_BX = nullptr;   // operands of the last residual (A X, B X — or of X_accurate in a quality-5 pass)
  int step_impl(int max_steps, int* passes_out);
  void print_state(bool header);
};

template <typename T>
void Solver<T>::release() {
  if (arena) {
    cudaStreamSynchronize(ctx->stream);   // nothing may still be running on a block the next solve reuses
    arena_release(ctx, arena, arena_bytes);
    if (ctx->active_solvers > 0) ctx->active_solvers--;
    ctx->oz_reuse = false;
  }
  arena = nullptr;
  slab[0] = slab[1] = AS = wA = wB = nullptr;
  // (the small matrices are carved out of the arena: nothing to free)
  G = GA = DinvR = Z = Tmp = Cx = Cp = Q = Tau = Graw = GAraw = Gc = GAc = Gw = GAw = Ccat = SigV = nullptr;
  D = Lam = Eig = Sums = Scal = Theta = nullptr;
  Count = nullptr;
  dSig = nullptr;
  if (geev_ws) cudaFree(geev_ws); geev_ws = nullptr;
  if (geev_piv) cudaFree(geev_piv); geev_piv = nullptr;
  if (geev_sig) cudaFree(geev_sig); geev_sig = nullptr;
  if (CxAcc) cudaFree(CxAcc); CxAcc = nullptr;
  if (hbuf) pinned_give(ctx, hbuf, hbuf_bytes);
  hbuf = nullptr;
  if (hX) cudaFreeHost(hX); hX = nullptr;
  if (hY) cudaFreeHost(hY); hY = nullptr;
}

template <typename T>
int Solver<T>::alloc() {
  const size_t nk = (size_t)n * k;
  const size_t m3 = 3 * (size_t)k;
  // one arena for every tall block: a single allocation => a single CUDA-IPC handle, and identical offsets on
  // every rank of a row-partitioned run (peer halo address = peer arena base + my offset)
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t slab_b = al(sizeof(T) * 3 * nk), wrk_b = al(sizeof(T) * std::max<size_t>(nk, 2 * (size_t)n));
  // AS = [AX | AW] with the cached Gram blocks, [AX | AP | AW] for the legacy pass, the indefinite solver and the helpers
  if (const char* e = getenv("LB2_GRAM_CACHE")) { if (!gram_cache_forced) gram_cache = (atoi(e) != 0); }
  { const char* e = getenv("LB2_GRAM_I8"); ctx->gram_i8_env = e ? atoi(e) : -1; }   // f64 Gram through the int8 tensor path (gram_i8.cu)
  if (indef) gram_cache = false;
  as_cols = (gram_cache && !helper_mode) ? 2 * k : 3 * k;
  const size_t as_b = al(sizeof(T) * (size_t)as_cols * n);
  // scratch blocks only where something uses them: wA holds B-products (B != NULL), wB the input of T, X_accurate of the
  // indefinite solver and the out-of-place result of SVQB — which, without those, goes to the slab that is not current
  // (dead while W is being orthogonalised).  B = T = NULL (config C5): 2 slabs + [AX | AW] = 8 n k scalars.
  const bool need_wA = opB != nullptr || helper_mode, need_wB = opT != nullptr || indef || helper_mode;
  // The small matrices live behind the tall blocks in the SAME allocation: one cudaMalloc per solve (none at all when the
  // context still holds the arena of the previous solve), no cudaFree at the end — two dozen cudaMalloc / cudaFree pairs
  // cost 0.25-1.4 s per reference-facing call on a device with ~100 GB mapped (tools/e2e_probe2.py, r02).
  struct Req { void** p; size_t bytes; bool zero; };
  std::vector<Req> req;
  auto want = [&](auto** p, size_t bytes, bool zero = false) { req.push_back(Req{(void**)p, al(bytes), zero}); };
  for (T** p : {&G, &GA, &DinvR, &Z, &Tmp}) want(p, sizeof(T) * m3 * m3);
  if (indef && !helper_mode) {   // W columns of S^H B S + the signature matrix of the pass (rr_indef)
    want(&Gw, sizeof(T) * m3 * k, true);
    want(&SigV, sizeof(T) * 4 * (size_t)k * k);
  }
  if (gram_cache) {
    want(&Graw, sizeof(T) * m3 * m3);
    want(&GAraw, sizeof(T) * m3 * m3);
    want(&Gc, sizeof(T) * 4 * (size_t)k * k);
    want(&GAc, sizeof(T) * 4 * (size_t)k * k);
    want(&Gw, sizeof(T) * m3 * k, true);    // zeroed: the column-block kernel never writes the tiles below the diagonal of the W block
    want(&GAw, sizeof(T) * m3 * k, true);
    want(&Ccat, sizeof(T) * m3 * 2 * k);
  }
  for (T** p : {&Cx, &Cp, &Q}) want(p, sizeof(T) * m3 * k);
  want(&Tau, sizeof(T) * m3);
  for (R** p : {&D, &Lam, &Eig, &Theta}) want(p, sizeof(R) * m3);
  want(&Sums, sizeof(R) * (m3 + 16));
  want(&Scal, sizeof(R) * 16);
  want(&Count, sizeof(int) * 4);
  want(&dSig, m3);
  size_t small_b = 0;
  for (auto& r : req) small_b += r.bytes;
  const size_t tall_b = 2 * slab_b + as_b + (need_wA ? wrk_b : 0) + (need_wB ? wrk_b : 0);
  arena_bytes = tall_b + small_b;
  tall_bytes = tall_b;
  arena = arena_alloc(ctx, arena_bytes);
  if (!arena) return -1;
  ctx->active_solvers++;
  // int8 path: exponent hints never cross solves (a solve must not depend on what ran before it on this context)
  ctx->oz_hint_valid[0] = ctx->oz_hint_valid[1] = ctx->oz_hint_valid[2] = false;
  char* base = (char*)arena;
  slab[0] = (T*)base; base += slab_b;
  slab[1] = (T*)base; base += slab_b;
  AS = (T*)base; base += as_b;
  wA = wB = nullptr;
  if (need_wA) { wA = (T*)base; base += wrk_b; }
  if (need_wB) { wB = (T*)base; base += wrk_b; }
  for (auto& r : req) {
    *r.p = (void*)base;
    if (r.zero) LB2_CUDA_OK(cudaMemsetAsync(base, 0, r.bytes, ctx->stream));
    base += r.bytes;
  }
  hbuf_bytes = sizeof(R) * (4 * (size_t)k + 64);
  hbuf = (R*)pinned_take(ctx, hbuf_bytes);
  if (!hbuf) return -1;
  hEig.assign(k, R(0));
  hRes.assign(k, R(0));
  return 0;
}

// z-slab partition of a stencil operator: every rank's X must be complete before neighbours read its boundary
// planes; a one-element all-reduce on the solver stream is the barrier (all later overwrites of X are separated from
// this read by the Gram / norm all-reduces of the pass, or by the barrier of the next apply).  Fills `local` with the
// peer halo pointers for X (which must live in the arena: same offset on every rank) and redirects b to it.
template <typename T>
int Solver<T>::localize(const BuiltinOp*& b, BuiltinOp& local, const T* X) {
  return localize_bytes(b, local, X, sizeof(T));
}
template <typename T>
int Solver<T>::localize_bytes(const BuiltinOp*& b, BuiltinOp& local, const void* X, size_t elem) {
  if (!(ctx->comm && b->n != b->n_global && (b->kind == OP_STENCIL || b->kind == OP_CSR || b->kind == OP_BDG))) return 0;
  tm.begin(PH_COMM);
  int rcb = allreduce_sum(ctx, Scal + 12, 1, kDouble);
  tm.end();
  if (rcb) return rcb;
  local = *b;
  const size_t off = (size_t)((const char*)X - (const char*)arena);
  if (off >= arena_bytes) {
    fprintf(stderr, "lobpcg_b200: partitioned operator applied to a buffer outside the solver arena\n");
    return -1;
  }
  if (b->kind == OP_CSR) {   // the neighbour's whole block (column 0, row 0); the kernel picks rows by column index
    local.halo_lo = peer_lo ? (const void*)(peer_lo + off) : nullptr;
    local.halo_hi = peer_hi ? (const void*)(peer_hi + off) : nullptr;
    local.halo_ld = n;
    b = &local;
    return 0;
  }
  const int64_t plane = b->gx * b->gy;
  local.halo_lo = peer_lo ? (const void*)(peer_lo + off + (size_t)((b->gz - 1) * plane) * elem) : nullptr;
  local.halo_hi = peer_hi ? (const void*)(peer_hi + off) : nullptr;
  local.halo_ld = n;
  b = &local;
  return 0;
}

// T = p(A) evaluated in the lower precision (float / c32) for a double / c64 solve: the preconditioner only has to be a
// fixed, good approximation of A^-1 on the unwanted part of the spectrum, and its cost is pure HBM traffic, which halves.
// X is converted into the workspace, the fused stencil + Chebyshev steps run on float blocks (same arena region, so
// row-partitioned halo reads keep working), the result is converted back.
template <typename T>
int Solver<T>::apply_cheb_mixed(const BuiltinOp* b, const BuiltinOp* in, int nc, const T* X, T* Y) {
  if constexpr (sizeof(R) != 8) {
    return -1;
  } else {
    using TL = typename std::conditional<Sc<T>::cplx, c32, float>::type;
    using RL = float;
    TL* w = reinterpret_cast<TL*>(slab[1 - cur]);      // 3k columns of T = 6k columns of TL, leading dimension n
    TL* Xf = w;
    TL* Rw = w + (int64_t)k * n;
    TL* D0 = w + (int64_t)2 * k * n;
    TL* D1 = w + (int64_t)3 * k * n;
    TL* Yf = w + (int64_t)4 * k * n;
    LB2_TRY((convert_block<TL, T>(ctx, n, nc, X, n, Xf, n)));
    BuiltinOp twin = *in;                              // the inner stencil as an operator of the lower precision
    twin.prefix = Sc<TL>::prefix;
    twin.potential = b->potential_lo;
    int rc = cheb_apply<TL>(
        ctx, b, n, nc, Xf, n, Yf, n, Rw, D0, D1, (TL*)nullptr, n,
        [&](int, const TL*, TL*) { return -1; },       // never taken: the fused step below always applies
        [&](int c, const TL* Din, TL* Dout, const TL* Rin, int64_t ldrin, TL* Rout, TL* Yacc, RL c1, RL c2, bool wr) {
          BuiltinOp local;
          const BuiltinOp* bb = &twin;
          if (int rcl = localize_bytes(bb, local, Din, sizeof(TL))) return rcl;
          ChebEpilogue<TL> ep;
          ep.rin = Rin; ep.ldrin = ldrin; ep.rout = Rout; ep.dout = Dout; ep.ldw = n; ep.c1 = c1; ep.c2 = c2;
          ep.write_r = wr ? 1 : 0;
          tm.begin(PH_SPMM);
          int r2 = spmm_stencil_cheb<TL>(ctx, stencil_desc(bb), c, Din, n, Yacc, n, ep);
          tm.end();
          phase_work[PH_SPMM] += (wr ? 6.0 : 5.0) * (double)n * c * sizeof(TL);
          phase_calls[PH_SPMM]++;
          return r2;
        });
    if (rc) return rc;
    return convert_block<T, TL>(ctx, n, nc, Yf, n, Y, n);
  }
}

// Y = Op X.  Built-in operators run as block kernels; anything else is a host callback (reference
// linop.h:15-17) and is staged through pinned host memory column by column — functional, not fast.
template <typename T>
int Solver<T>::apply(const LinOpRaw* op, int nc, const T* X, T* Y) {
  if (nc <= 0) return 0;
  const BuiltinOp* b = builtin_of(op);
  if (b && b->kind == OP_CHEB) {
    // T = p(A): workspace = the slab that is not current plus AS[:, k:2k] (both dead whenever T is applied: the old
    // [X P W] has just been projected and [AP AW] is recomputed by the next Rayleigh-Ritz; step_impl / init), which also
    // keeps every vector the inner operator reads inside the arena (row-partitioned runs read halo planes from the
    // neighbour's arena at the same offset)
    if (nc > k) { fprintf(stderr, "lobpcg_b200: preconditioner applied to more than sizeSub columns\n"); return -1; }
    T* w = slab[1 - cur];
    if (X == w || Y == w) { fprintf(stderr, "lobpcg_b200: preconditioner workspace aliases its operands\n"); return -1; }
    const BuiltinOp* in = builtin_of(b->inner);
    const bool can_fuse = in && in->kind == OP_STENCIL && !getenv("LB2_NO_CHEB_FUSE");
    if constexpr (sizeof(R) == 8) {
      if (b->cheb_mixed && can_fuse) return apply_cheb_mixed(b, in, nc, X, Y);
    }
    return cheb_apply<T>(
        ctx, b, n, nc, X, n, Y, n, w, col(w, k), col(w, 2 * k), col(AS, k), n,
        [&](int c, const T* D, T* AD) { return apply(b->inner, c, D, AD); },
        [&](int c, const T* Din, T* Dout, const T* Rin, int64_t ldrin, T* Rout, T* Yacc, R c1, R c2, bool wr) {
          if (!can_fuse) return -100;
          BuiltinOp local;
          const BuiltinOp* bb = in;
          if (int rcl = localize(bb, local, Din)) return rcl;
          ChebEpilogue<T> ep;
          ep.rin = Rin; ep.ldrin = ldrin; ep.rout = Rout; ep.dout = Dout; ep.ldw = n; ep.c1 = c1; ep.c2 = c2;
          ep.write_r = wr ? 1 : 0;
          tm.begin(PH_SPMM);
          int rc = spmm_stencil_cheb<T>(ctx, stencil_desc(bb), c, Din, n, Yacc, n, ep);
          tm.end();
          phase_work[PH_SPMM] += (wr ? 6.0 : 5.0) * (double)n * c * sizeof(T);
          phase_calls[PH_SPMM]++;
          return rc;
        });
  }
  if (b) {
    BuiltinOp local;
    if (int rcl = localize(b, local, X)) return rcl;
    tm.begin(PH_SPMM);
    int rc = apply_builtin<T>(ctx, b, nc, X, n, Y, n);
    tm.end();
    double bytes = 2.0 * (double)n * nc * sizeof(T);
    if (b->kind == OP_CSR || b->from_csr) bytes += (double)b->nnz * (sizeof(T) + 4) + 8.0 * (double)(b->n + 1);
    if (b->kind == OP_DIAG) bytes += (double)n * sizeof(R);
    if (b->kind == OP_DENSE) bytes += (double)n * (double)n * sizeof(T);
    phase_work[PH_SPMM] += bytes;
    phase_calls[PH_SPMM]++;
    return rc;
  }
  if (!hX) {
    LB2_CUDA_OK(cudaMallocHost(&hX, sizeof(T) * (size_t)n * std::max(k, 2)));
    LB2_CUDA_OK(cudaMallocHost(&hY, sizeof(T) * (size_t)n * std::max(k, 2)));
  }
  const int chunk = std::max(k, 2);
  for (int c0 = 0; c0 < nc; c0 += chunk) {
    const int w = std::min(chunk, nc - c0);
    LB2_CUDA_OK(cudaMemcpyAsync(hX, X + (int64_t)c0 * n, sizeof(T) * (size_t)n * w, cudaMemcpyDeviceToHost, ctx->stream));
    LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    for (int j = 0; j < w; j++) op->matvec(op, hX + (size_t)j * n, hY + (size_t)j * n);
    LB2_CUDA_OK(cudaMemcpyAsync(Y + (int64_t)c0 * n, hY, sizeof(T) * (size_t)n * w, cudaMemcpyHostToDevice, ctx->stream));
    LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

template <typename T>
int Solver<T>::gram_ar(int ma, int mb, const T* A, const T* B, T* Gout, int upper) {
  tm.begin(PH_GRAM);
  int rc = gram<T>(ctx, n, ma, mb, A, n, B, n, Gout, ma, upper);
  tm.end();
  phase_work[PH_GRAM] += (Sc<T>::cplx ? 4.0 : 1.0) * (double)n * ma * (upper ? (double)(ma + 1) : 2.0 * mb);
  phase_calls[PH_GRAM]++;
  if (rc) return rc;
  if (reduce()) {
    tm.begin(PH_COMM);
    rc = allreduce_sum(ctx, Gout, (size_t)ma * mb * kCplx, kDouble);
    tm.end();
  }
  return rc;
}

template <typename T>
int Solver<T>::nn(int kd, int nb, T alpha, const T* S, const T* C, int ldc, T beta, T* Out) {
  tm.begin(PH_TALLNN);
  int rc = tall_nn<T>(ctx, n, kd, nb, alpha, S, n, C, ldc, beta, Out, n);
  tm.end();
  phase_work[PH_TALLNN] += (Sc<T>::cplx ? 8.0 : 2.0) * (double)n * kd * nb;
  phase_calls[PH_TALLNN]++;
  return rc;
}

template <typename T>
int Solver<T>::sumsq_total(int nc, const T* X, R* out_dev) {
  // nc <= 3k everywhere in this solver (Sums holds 3k + 16 entries)
  tm.begin(PH_RESID);
  int rc = col_sumsq<T>(ctx, n, nc, X, n, Sums);
  if (!rc) rc = sum_reals<R>(ctx, nc, Sums, out_dev);
  tm.end();
  phase_work[PH_RESID] += (double)n * nc * sizeof(T);
  phase_calls[PH_RESID]++;
  if (!rc && reduce()) rc = allreduce_sum(ctx, out_dev, 1, kDouble);
  return rc;
}

template <typename T>
int Solver<T>::estimate_norm(const LinOpRaw* op, uint64_t seed, R* out) {
  T* x = slab[1];             // two columns of the slab that is not in use yet (X0 lives in slab[0]; init / helpers only)
  T* y = slab[1] + n;
  LB2_TRY(fill_rows(x, 1, seed));
  LB2_TRY(sumsq_total(1, x, Scal));
  LB2_TRY(normalize_by<T>(ctx, n, x, Scal));
  for (int it = 0; it < 10; it++) {
    LB2_TRY(apply(op, 1, x, y));
    LB2_TRY(sumsq_total(1, y, Scal));
    LB2_TRY(normalize_by<T>(ctx, n, y, Scal));
    std::swap(x, y);
  }
  LB2_TRY(d2h(hbuf, Scal, sizeof(R)));
  LB2_TRY(sync());
  *out = std::sqrt(hbuf[0]);
  return 0;
}

// G = S^H B S, all of it valid (mirrored when B == NULL, full product otherwise; gram_impl.inc:49-67)
template <typename T>
int Solver<T>::gram_self_B(int m, T* S, T* Gout) {
  if (!opB) return gram_ar(m, m, S, S, Gout, 1);
  for (int c0 = 0; c0 < m; c0 += k) {
    const int w = std::min(k, m - c0);
    LB2_TRY(apply(opB, w, col(S, c0), wA));
    LB2_TRY(gram_ar(m, w, S, wA, Gout + (size_t)c0 * m, 0));
  }
  return 0;
}

// D-scaling, upper Cholesky, condition check, DinvR = D R^-1  (rayleigh_ritz_modified_impl.inc:148-186)
template <typename T>
int Solver<T>::chol_transform(int m, int* bad) {
  tm.begin(PH_SMALL);
  int info = 0;
  LB2_TRY(sd_dscale<T>(ctx, m, G, m, D));
  LB2_TRY(sd_potrf_upper<T>(ctx, m, G, m, &info));
  if (info != 0) { tm.end(); *bad = 1; return 0; }
  LB2_TRY(sd_set_diag<T>(ctx, m, DinvR, m, D));
  LB2_TRY(sd_trsm_run<T>(ctx, m, m, G, m, DinvR, m));
  LB2_TRY(sd_rcond<T>(ctx, m, G, m, DinvR, m, D, Scal));
  LB2_TRY(d2h(hbuf, Scal, sizeof(R)));
  tm.end();
  LB2_TRY(sync());
  *bad = (hbuf[0] < (R)5.0e-3) ? 2 : 0;
  return 0;
}

// Initial Rayleigh-Ritz on X (rayleigh_ritz_impl.inc:37-100) + X <- X Cx (lobpcg_impl.inc:99-104)
template <typename T>
int Solver<T>::rr_initial() {
  T* X = Xp();
  LB2_TRY(gram_self_B(k, X, G));
  int bad = 0;
  LB2_TRY(chol_transform(k, &bad));
  if (bad == 1) {
    fprintf(stderr, "rayleigh_ritz: Cholesky failed\n");
    return 1;
  }
  LB2_TRY(apply(opA, k, X, AS));
  LB2_TRY(gram_ar(k, k, X, AS, GA, 1));
  tm.begin(PH_SMALL);
  LB2_TRY(sd_gemm<T>(ctx, 'N', k, k, k, GA, k, DinvR, k, Tmp, k));
  LB2_TRY(sd_gemm<T>(ctx, 'H', k, k, k, DinvR, k, Tmp, k, Z, k));
  int info = 0;
  LB2_TRY(sd_syevd_upper<T>(ctx, k, Z, k, Lam, &info));
  if (info != 0) {
    tm.end();
    fprintf(stderr, "rayleigh_ritz: eigensolve failed\n");
    return 1;
  }
  LB2_TRY(sd_gemm<T>(ctx, 'N', k, k, k, DinvR, k, Z, k, Cx, k));
  LB2_CUDA_OK(cudaMemcpyAsync(Eig, Lam, sizeof(R) * k, cudaMemcpyDeviceToDevice, ctx->stream));
  tm.end();
  T* Xn = slab[1 - cur];
  LB2_TRY(nn(k, k, make<T>(1), X, Cx, k, zero<T>(), Xn));
  cur = 1 - cur;
  return 0;
}

// VQ (m x k) = Z[:, k:m] * orth(Z[0:k, k:m]^T)   (rayleigh_ritz_modified_impl.inc:98-130 / 230-264)
template <typename T>
int Solver<T>::cp_from_z(int m, const T* Zm, T* VQ) {
  const int nrem = m - k;
  LB2_TRY(sd_transpose<T>(ctx, nrem, k, Zm + (size_t)k * m, m, Q, nrem));
  LB2_TRY(sd_qr_q<T>(ctx, nrem, k, Q, nrem, Tau));
  if (ctx->dev_info && hbuf) {   // read with the next host synchronisation (Solver::sync)
    LB2_CUDA_OK(cudaMemcpyAsync(hbuf + 4 * (size_t)k + 60, ctx->dev_info + 2, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    qr_info_pending = 1;
  }
  LB2_TRY(sd_gemm<T>(ctx, 'N', m, k, nrem, Zm + (size_t)k * m, m, Q, nrem, VQ, m));
  return 0;
}

// Modified Rayleigh-Ritz on S = slab[cur][:, 0:m] = [X | P_act | W_act] (np, nw = widths of P_act, W_act).  On return
// useOrtho in {0,1,2}; for 0/1: Cx, Cp (m x k, ld m) and Eig[0:k] are set.
//
// Legacy pass (gram_cache off): S^H B S and S^H A S are contracted in full, AS[:, k:m] = A [P W] is recomputed here.
// Cached pass: only the W columns are contracted — one launch for [S^H (B)W | S^H AW] — and the [X P] blocks
// come from the cache that step_impl derives from this pass's coefficients (C^H G C); when the cache is not valid
// (first pass, every cache_period passes, after the drift monitor tripped) the [X P] blocks are recomputed from the
// tall vectors first.  AS = [AX | AW]; A P is only formed on such a refresh.
template <typename T>
int Solver<T>::rr_modified(int m) {
  T* S = Xp();
  const int nrem = m - k;
  const int mxp = k + np, nwc = m - mxp;
  const bool cached = gram_cache && nwc > 0;
  if (!cached) {
    cache_ok = false;
    if (useOrtho == 0) {
      LB2_TRY(gram_self_B(m, S, G));
      int bad = 0;
      LB2_TRY(chol_transform(m, &bad));
      if (bad) {
        if (bad == 1) fprintf(stderr, "rayleigh_ritz_modified: Cholesky failed\n");
        useOrtho = 2;
        return 0;
      }
    } else {
      useOrtho = 1;
    }
    if (as_cols < m) { fprintf(stderr, "lobpcg_b200: internal error: AS holds %d columns, need %d\n", as_cols, m); return -1; }
    LB2_TRY(apply(opA, m - k, col(S, k), col(AS, k)));
    LB2_TRY(gram_ar(m, m, S, AS, GA, 1));
  } else {
    const bool chol = (useOrtho == 0);
    if (!chol) useOrtho = 1;
    T* Wp = col(S, mxp);
    T* AW = col(AS, k);
    if (!cache_ok || cache_mxp != mxp || (chol && !cache_has_b)) {
      // refresh: [X P]^H B [X P] and [X P]^H A [X P] from the tall vectors (A P goes where A W will be written next)
      if (chol) LB2_TRY(gram_self_B(mxp, S, Gc));
      if (np > 0) LB2_TRY(apply(opA, np, col(S, k), col(AS, k)));
      LB2_TRY(gram_ar(mxp, mxp, S, AS, GAc, 1));
      cache_ok = true;
      cache_has_b = chol;
      cache_mxp = mxp;
      since_refresh = 0;
      cache_refreshes++;
    }
    LB2_TRY(apply(opA, nwc, Wp, AW));
    const T* BW = Wp;
    if (chol && opB) { LB2_TRY(apply(opB, nwc, Wp, wA)); BW = wA; }
    tm.begin(PH_GRAM);
    int rc = chol ? gram_cols<T>(ctx, n, m, nwc, S, n, BW, n, Gw, m, AW, n, GAw, m, mxp)
                  : gram_cols<T>(ctx, n, m, nwc, S, n, AW, n, GAw, m, (const T*)nullptr, 0, (T*)nullptr, 0, mxp);
    tm.end();
    ctx->oz_reuse = true;   // int8 path: S is not written again before this pass's projections X' = S Cx, P' = S Cp read its slices
    // algorithmic flops: rectangular [X P]^H W part + Hermitian W^H W part, per product
    phase_work[PH_GRAM] += (chol ? 2.0 : 1.0) * (Sc<T>::cplx ? 4.0 : 1.0) * (double)n * (2.0 * mxp * nwc + (double)nwc * (nwc + 1));
    phase_calls[PH_GRAM]++;
    if (rc) return rc;
    if (reduce()) {
      tm.begin(PH_COMM);
      if (chol) rc = allreduce_sum(ctx, Gw, (size_t)m * nwc * kCplx, kDouble);
      if (!rc) rc = allreduce_sum(ctx, GAw, (size_t)m * nwc * kCplx, kDouble);
      tm.end();
      if (rc) return rc;
    }
    tm.begin(PH_SMALL);
    if (chol) {
      LB2_TRY(sd_assemble_gram<T>(ctx, m, mxp, Gc, mxp, Gw, m, Graw, m));
      LB2_CUDA_OK(cudaMemcpyAsync(G, Graw, sizeof(T) * (size_t)m * m, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    LB2_TRY(sd_assemble_gram<T>(ctx, m, mxp, GAc, mxp, GAw, m, GAraw, m));
    LB2_CUDA_OK(cudaMemcpyAsync(GA, GAraw, sizeof(T) * (size_t)m * m, cudaMemcpyDeviceToDevice, ctx->stream));
    tm.end();
    if (chol) {
      int bad = 0;
      LB2_TRY(chol_transform(m, &bad));
      if (bad) {
        if (bad == 1) fprintf(stderr, "rayleigh_ritz_modified: Cholesky failed\n");
        useOrtho = 2;
        return 0;
      }
    }
  }
  tm.begin(PH_SMALL);
  T* Zm = Z;
  if (useOrtho == 0) {
    LB2_TRY(sd_gemm<T>(ctx, 'N', m, m, m, GA, m, DinvR, m, Tmp, m));
    LB2_TRY(sd_gemm<T>(ctx, 'H', m, m, m, DinvR, m, Tmp, m, Z, m));
  } else {
    Zm = GA;
  }
  int info = 0;
  LB2_TRY(sd_syevd_upper<T>(ctx, m, Zm, m, Lam, &info));
  if (info != 0) {
    tm.end();
    fprintf(stderr, "rayleigh_ritz_modified: eigensolve failed (info=%d)\n", info);
    return 1;
  }
  LB2_CUDA_OK(cudaMemcpyAsync(Eig, Lam, sizeof(R) * k, cudaMemcpyDeviceToDevice, ctx->stream));
  if (useOrtho == 0) {
    LB2_TRY(sd_gemm<T>(ctx, 'N', m, k, m, DinvR, m, Zm, m, Cx, m));
  } else {
    LB2_CUDA_OK(cudaMemcpyAsync(Cx, Zm, sizeof(T) * (size_t)m * k, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if (nrem < k) {
    // reference guard (:90-94, :222-226): no room for a k-column P basis
    LB2_CUDA_OK(cudaMemsetAsync(Cp, 0, sizeof(T) * (size_t)m * k, ctx->stream));
  } else if (useOrtho == 0) {
    LB2_TRY(cp_from_z(m, Zm, Tmp));
    LB2_TRY(sd_gemm<T>(ctx, 'N', m, k, m, DinvR, m, Tmp, m, Cp, m));
  } else {
    LB2_TRY(cp_from_z(m, Zm, Cp));
  }
  tm.end();
  return 0;
}

// After a pass: the cached [X P] blocks of the NEXT pass's basis [X' | P'_act] = S [Cx | Cp[:, nconv:]] from this pass's
// assembled Gram matrices: G' = C^H G C (the "analytical Gram blocks" of the reference's design notes, computed instead of
// assumed so that the rounding of the small factorisations is carried along).  Ortho branch: only the A block.
template <typename T>
int Solver<T>::update_gram_cache(int m, int nconv) {
  if (!gram_cache || !cache_ok) return 0;
  const int nact = k - nconv, mc = k + nact;
  tm.begin(PH_SMALL);
  LB2_CUDA_OK(cudaMemcpyAsync(Ccat, Cx, sizeof(T) * (size_t)m * k, cudaMemcpyDeviceToDevice, ctx->stream));
  if (nact > 0)
    LB2_CUDA_OK(cudaMemcpyAsync(Ccat + (size_t)m * k, Cp + (size_t)nconv * m, sizeof(T) * (size_t)m * nact,
                                cudaMemcpyDeviceToDevice, ctx->stream));
  if (useOrtho == 0) {
    LB2_TRY(sd_gemm<T>(ctx, 'N', m, mc, m, Graw, m, Ccat, m, Tmp, m));
    LB2_TRY(sd_gemm<T>(ctx, 'H', mc, mc, m, Ccat, m, Tmp, m, Gc, mc));
    cache_has_b = true;
  } else {
    cache_has_b = false;
  }
  LB2_TRY(sd_gemm<T>(ctx, 'N', m, mc, m, GAraw, m, Ccat, m, Tmp, m));
  LB2_TRY(sd_gemm<T>(ctx, 'H', mc, mc, m, Ccat, m, Tmp, m, GAc, mc));
  tm.end();
  cache_mxp = mc;
  since_refresh++;
  // drift monitor (residual_pass) and the periodic refresh decide whether the next pass trusts these blocks
  if (since_refresh >= cache_period || cache_monitor > monitor_threshold()) cache_ok = false;
  return 0;
}

// SVQB (svqb_impl.inc:48-106): U (n x nu) <- U T, returns retained columns.
template <typename T>
int Solver<T>::svqb(T* U, int nu, R tau, bool drop, int* nret) {
  *nret = nu;
  if (nu == 0) return 0;
  if (opB) {
    LB2_TRY(apply(opB, nu, U, wA));
    LB2_TRY(gram_ar(nu, nu, U, wA, G, 1));
  } else {
    LB2_TRY(gram_ar(nu, nu, U, U, G, 1));
  }
  const bool reuse_u = true;   // int8 path: U is unchanged until U T below has been formed — the slices of this Gram serve it
  tm.begin(PH_SMALL);
  LB2_TRY(sd_dscale<T>(ctx, nu, G, nu, D));
  int info = 0;
  LB2_TRY(sd_syevd_upper<T>(ctx, nu, G, nu, Lam, &info));
  if (info != 0) {
    tm.end();
    fprintf(stderr, "svqb: eig failed with info=%d\n", info);
    return 0;  // reference returns cols unchanged
  }
  LB2_TRY(sd_svqb_transform<T>(ctx, nu, G, nu, Lam, D, tau, drop ? 1 : 0, Tmp, nu, Count));
  LB2_CUDA_OK(cudaMemcpyAsync(hbuf, Count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  tm.end();
  LB2_TRY(sync());
  const int keep = *(int*)hbuf;
  T* out = wB ? wB : slab[1 - cur];   // (no scratch block: the other slab is dead while W is orthogonalised)
  if (out == U || (U > out && U < out + (int64_t)keep * n)) { fprintf(stderr, "lobpcg_b200: svqb workspace aliases its operand\n"); return -1; }
  ctx->oz_reuse = reuse_u;
  LB2_TRY(nn(nu, keep, make<T>(1), U, Tmp, nu, zero<T>(), out));
  ctx->oz_reuse = false;
  tm.begin(PH_OTHER);
  LB2_TRY(copy_block<T>(ctx, n, keep, out, n, U, n));
  tm.end();
  *nret = keep;
  return 0;
}

// ortho_drop (ortho_drop_impl.inc:43-125): B-orthogonalise U against V, B-orthonormalise U.
template <typename T>
int Solver<T>::ortho_drop(T* U, int nu0, T* V, int nv, int* nret, bool indefinite) {
  *nret = nu0;
  if (nu0 == 0 || nv == 0) return 0;
  if (ng < (int64_t)nu0 + nv) {
    fprintf(stderr, "ortho_drop: overdetermined shape m=%lu < n_u+n_v=%lu+%lu\n", (unsigned long)ng,
            (unsigned long)nu0, (unsigned long)nv);
    return 0;
  }
  const R eps = eps_tol();
  int nu = nu0;
  // indefinite metric (ortho_indefinite_impl.inc:98-105): signature matrix sig = V^H B V, kept in GA
  if (indefinite) {
    LB2_TRY(gram_self_B(nv, V, GA));
    if (SigV && !helper_mode) {   // the same matrix is the [X P] block of S^H B S in this pass's Rayleigh-Ritz (rr_indef)
      LB2_CUDA_OK(cudaMemcpyAsync(SigV, GA, sizeof(T) * (size_t)nv * nv, cudaMemcpyDeviceToDevice, ctx->stream));
      sigv_cols = nv;
    }
  }
  // ||B V||_F
  R BV_norm = 0;
  {
    double acc = 0;
    if (!opB) {
      LB2_TRY(sumsq_total(nv, V, Scal));
      LB2_TRY(d2h(hbuf, Scal, sizeof(R)));
      LB2_TRY(sync());
      acc = hbuf[0];
    } else {
      for (int c0 = 0; c0 < nv; c0 += k) {
        const int w = std::min(k, nv - c0);
        LB2_TRY(apply(opB, w, col(V, c0), wA));
        LB2_TRY(sumsq_total(w, wA, Scal));
        LB2_TRY(d2h(hbuf, Scal, sizeof(R)));
        LB2_TRY(sync());
        acc += hbuf[0];
      }
    }
    BV_norm = (R)std::sqrt(acc);
    if (BV_norm < eps) BV_norm = 1;
  }
  for (int outer = 0; outer < 3; outer++) {
    // C = V^H B U ; U -= V C
    const T* BU = U;
    if (opB) { LB2_TRY(apply(opB, nu, U, wA)); BU = wA; }
    LB2_TRY(gram_ar(nv, nu, V, BU, Tmp, 0));
    ctx->oz_reuse = true;    // int8 path: V is unchanged between this product and the update below — its slices serve both
    if (indefinite) {  // U -= V (sig (V^H B U))   (ortho_indefinite_impl.inc:121-133)
      LB2_TRY(sd_gemm<T>(ctx, 'N', nv, nu, nv, GA, nv, Tmp, nv, Z, nv));
      LB2_TRY(nn(nv, nu, make<T>(-1), V, Z, nv, make<T>(1), U));
    } else {
      LB2_TRY(nn(nv, nu, make<T>(-1), V, Tmp, nv, make<T>(1), U));
    }
    ctx->oz_reuse = false;
    for (int inner = 0; inner < 3; inner++) {
      int keep = nu;
      LB2_TRY(svqb(U, nu, eps, true, &keep));
      if (keep != nu) {
        fprintf(stderr, "ortho_drop: svqb dropped columns (%lu -> %lu)\n", (unsigned long)nu, (unsigned long)keep);
        nu = keep;
      }
      if (nu == 0) break;
      // ||U^H B U - I||_F / (||B U|| ||U||)
      LB2_TRY(sumsq_total(nu, U, Scal + 1));
      if (opB) {
        LB2_TRY(apply(opB, nu, U, wA));
        LB2_TRY(gram_ar(nu, nu, U, wA, G, 1));
        LB2_TRY(sumsq_total(nu, wA, Scal + 2));
      } else {
        LB2_TRY(gram_ar(nu, nu, U, U, G, 1));
      }
      LB2_TRY(sd_ortho_err_upper<T>(ctx, nu, G, nu, Scal));
      LB2_TRY(d2h(hbuf, Scal, 3 * sizeof(R)));
      LB2_TRY(sync());
      R U_norm = std::sqrt(hbuf[1]);
      if (U_norm < eps) U_norm = 1;
      const R BU_norm = opB ? std::sqrt(hbuf[2]) : U_norm;
      // ortho_indefinite uses ||U||^2 as denominator (ortho_indefinite_impl.inc:141-152)
      const R rerr = indefinite ? hbuf[0] / (U_norm * U_norm) : hbuf[0] / (BU_norm * U_norm);
      if (rerr < eps) break;
    }
    if (nu == 0) break;
    // ||V^H B U||_F / (||B V|| ||U||)
    const T* BU2 = U;
    if (opB) { LB2_TRY(apply(opB, nu, U, wA)); BU2 = wA; }
    LB2_TRY(gram_ar(nv, nu, V, BU2, Tmp, 0));
    LB2_TRY(sd_frob<T>(ctx, nv, nu, Tmp, nv, Scal));
    LB2_TRY(sumsq_total(nu, U, Scal + 1));
    LB2_TRY(d2h(hbuf, Scal, 2 * sizeof(R)));
    LB2_TRY(sync());
    R U_norm = std::sqrt(hbuf[1]);
    if (U_norm < eps) U_norm = 1;
    const R rerr = hbuf[0] / (BV_norm * U_norm);
    if (rerr < eps) break;
  }
  *nret = nu;
  return 0;
}

// svqb_mat (src/ortho/svqb_mat_impl.inc:49-100), drop = 'n': U (m x nu) <- U T with T from the eigendecomposition of
// the D-scaled Gram U^H mat U.  Scratch: DinvR blocks t1/t2/c1 and Q.
template <typename T>
int Solver<T>::svqb_mat_dev(int m, int nu, T* U, const T* mat, R tau) {
  if (nu == 0) return 0;
  const size_t blk = (size_t)3 * k * k;
  T* t1 = DinvR;
  T* t2 = DinvR + blk;
  T* c1 = DinvR + 2 * blk;
  T* c2 = Q;
  if (mat) {
    LB2_TRY(sd_gemm<T>(ctx, 'N', m, nu, m, mat, m, U, m, t1, m));
    LB2_TRY(sd_gemm<T>(ctx, 'H', nu, nu, m, U, m, t1, m, c1, nu));
  } else {
    LB2_TRY(sd_gemm<T>(ctx, 'H', nu, nu, m, U, m, U, m, c1, nu));
  }
  LB2_TRY(sd_dscale<T>(ctx, nu, c1, nu, D));
  int info = 0;
  LB2_TRY(sd_syevd_upper<T>(ctx, nu, c1, nu, Lam, &info));
  if (info != 0) {
    fprintf(stderr, "svqb_mat: eig failed with info=%d\n", info);
    return 0;
  }
  LB2_TRY(sd_svqb_transform<T>(ctx, nu, c1, nu, Lam, D, tau, 0, c2, nu, Count));
  LB2_TRY(sd_gemm<T>(ctx, 'N', m, nu, nu, U, m, c2, nu, t2, m));
  LB2_CUDA_OK(cudaMemcpyAsync(U, t2, sizeof(T) * (size_t)m * nu, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

// Coefficient-space orthogonalisation of U (m x nu) against V (m x nv) in the metric `mat` (m x m):
// ortho_indefinite_mat + svqb_mat of the reference (src/ortho/ortho_indefinite_mat_impl.inc:52-123,
// src/ortho/svqb_mat_impl.inc:49-100), everything on the device.  Scratch lives in DinvR / Q / GA.
template <typename T>
int Solver<T>::ortho_indef_mat(int m, int nu, int nv, T* U, const T* V, const T* mat) {
  if (nu == 0 || nv == 0) return 0;
  if (m < nu + nv) {
    fprintf(stderr, "ortho_indefinite_mat: overdetermined m=%lu < n_u+n_v=%lu+%lu\n", (unsigned long)m,
            (unsigned long)nu, (unsigned long)nv);
    return 0;
  }
  const R eps = eps_tol();
  const size_t blk = (size_t)3 * k * k;
  T* t1 = DinvR;            // m x max(nu,nv)
  T* t2 = DinvR + blk;      // m x nu
  T* c1 = DinvR + 2 * blk;  // nv x nu  /  nu x nu   (svqb_mat_dev keeps its nu x nu transform in Q)
  LB2_TRY(sd_gemm<T>(ctx, 'N', m, nv, m, mat, m, V, m, t1, m));
  LB2_TRY(sd_frob<T>(ctx, m, nv, t1, m, Scal));
  LB2_TRY(d2h(hbuf, Scal, sizeof(R)));
  LB2_TRY(sync());
  R MV_norm = hbuf[0];
  if (MV_norm < eps) MV_norm = 1;
  for (int outer = 0; outer < 3; outer++) {
    // U -= V V^H mat V V^H mat U, right to left (:85-101)
    LB2_TRY(sd_gemm<T>(ctx, 'N', m, nu, m, mat, m, U, m, t1, m));
    LB2_TRY(sd_gemm<T>(ctx, 'H', nv, nu, m, V, m, t1, m, c1, nv));
    LB2_TRY(sd_gemm<T>(ctx, 'N', m, nu, nv, V, m, c1, nv, t2, m));
    LB2_TRY(sd_gemm<T>(ctx, 'N', m, nu, m, mat, m, t2, m, t1, m));
    LB2_TRY(sd_gemm<T>(ctx, 'H', nv, nu, m, V, m, t1, m, c1, nv));
    LB2_TRY(sd_gemm_ab<T>(ctx, 'N', m, nu, nv, make<T>(-1), V, m, c1, nv, make<T>(1), U, m));
    for (int inner = 0; inner < 3; inner++) {
      LB2_TRY(svqb_mat_dev(m, nu, U, mat, eps));
      // ||U^H mat U - I_sig||_F / ||U||^2
      LB2_TRY(sd_gemm<T>(ctx, 'N', m, nu, m, mat, m, U, m, t1, m));
      LB2_TRY(sd_gemm<T>(ctx, 'H', nu, nu, m, U, m, t1, m, c1, nu));
      LB2_TRY(sd_ortho_err_upper<T>(ctx, nu, c1, nu, Scal));
      LB2_TRY(sd_frob<T>(ctx, m, nu, U, m, Scal + 1));
      LB2_TRY(d2h(hbuf, Scal, 2 * sizeof(R)));
      LB2_TRY(sync());
      R Un = hbuf[1];
      if (Un < eps) Un = 1;
      if (hbuf[0] / (Un * Un) < eps) break;
    }
    LB2_TRY(sd_gemm<T>(ctx, 'N', m, nu, m, mat, m, U, m, t1, m));
    LB2_TRY(sd_gemm<T>(ctx, 'H', nv, nu, m, V, m, t1, m, c1, nv));
    LB2_TRY(sd_frob<T>(ctx, nv, nu, c1, nv, Scal));
    LB2_TRY(sd_frob<T>(ctx, m, nu, U, m, Scal + 1));
    LB2_TRY(d2h(hbuf, Scal, 2 * sizeof(R)));
    LB2_TRY(sync());
    R Un = hbuf[1];
    if (Un < eps) Un = 1;
    if (hbuf[0] / (MV_norm * Un) < eps) break;
  }
  return 0;
}

// Indefinite Rayleigh-Ritz on S = slab[cur][:, 0:m] (reference src/rayleigh/indefinite_rr_impl.inc:51-149 and
// indefinite_rr_modified_impl.inc:58-255).  The reference hands (G_A, G_B) to LAPACK GGEV, B-normalises the
// eigenvectors twice and checks their B-orthogonality (quality flag); cuSOLVER has no GGEV.
//  * G_A = S^H A S positive definite (the BdG-type pencils the solver targets): the same eigenpairs come from a Hermitian
//    problem, G_A = R^H R, K = R^-H G_B R^-1 = W diag(mu) W^H, v = R^-1 w, theta = 1/mu, v^H G_B v = mu — B-orthogonal
//    by construction, i.e. always the reference's quality_flag == 1 path.
//  * otherwise (Cholesky of G_A fails; or option "indef_geev"): rr_indef_general below — the non-symmetric route with the
//    reference's own B-normalisation, quality check and quality-5 fallback.
template <typename T>
int Solver<T>::rr_indef(int m, int from_col, bool initial) {
  T* S = Xp();
  last_quality = 1;
  LB2_TRY(apply(opA, m - from_col, col(S, from_col), col(AS, from_col)));
  LB2_TRY(gram_ar(m, m, S, AS, GA, 1));
  const int mxp = k + np, nwc = m - mxp;
  if (!initial && SigV && sigv_cols == mxp && nwc > 0 && nwc <= k && !helper_mode) {
    // S^H B S = [[V^H B V, V^H B W], [., W^H B W]]: the [X P] block is the signature matrix ortho_indefinite computed for
    // this very V a moment ago (ortho_drop); only the W columns are contracted over n (6 n k^2 instead of 18 n k^2).
    // (Carrying V^H B V and V^H A V ACROSS passes as C^H G C — what the definite solver does, §4b of DESIGN.md — was tried and
    // removed: the indefinite Rayleigh-Ritz has no column scaling, ||C|| is large, and the congruence lost the positive
    // definiteness of S^H A S / stalled the negative-shift pencils.)
    T* Wp = col(S, mxp);
    LB2_TRY(apply(opB, nwc, Wp, wA));
    tm.begin(PH_GRAM);
    int rc = gram_cols<T>(ctx, n, m, nwc, S, n, wA, n, Gw, m, (const T*)nullptr, 0, (T*)nullptr, 0, mxp);
    tm.end();
    phase_work[PH_GRAM] += (Sc<T>::cplx ? 4.0 : 1.0) * (double)n * (2.0 * mxp * nwc + (double)nwc * (nwc + 1));
    phase_calls[PH_GRAM]++;
    if (rc) return rc;
    if (reduce()) {
      tm.begin(PH_COMM);
      rc = allreduce_sum(ctx, Gw, (size_t)m * nwc * kCplx, kDouble);
      tm.end();
      if (rc) return rc;
    }
    LB2_TRY(sd_assemble_gram<T>(ctx, m, mxp, SigV, mxp, Gw, m, G, m));
  } else {
    LB2_TRY(gram_self_B(m, S, G));
  }
  sigv_cols = 0;   // valid for one Rayleigh-Ritz only (V changes with the projection)
  tm.begin(PH_SMALL);
  LB2_CUDA_OK(cudaMemcpyAsync(Tmp, G, sizeof(T) * (size_t)m * m, cudaMemcpyDeviceToDevice, ctx->stream));
  LB2_CUDA_OK(cudaMemcpyAsync(DinvR, GA, sizeof(T) * (size_t)m * m, cudaMemcpyDeviceToDevice, ctx->stream));   // G_A survives a failed potrf
  int info = 0;
  if (!indef_geev) LB2_TRY(sd_potrf_upper<T>(ctx, m, GA, m, &info));
  if (indef_geev || info != 0) {
    if (int rc = rr_indef_general(m, initial)) { tm.end(); return rc; }
  } else {
    LB2_TRY(sd_trsm_upper<T>(ctx, 'R', 'N', m, m, GA, m, Tmp, m));
    LB2_TRY(sd_trsm_upper<T>(ctx, 'L', 'H', m, m, GA, m, Tmp, m));
    LB2_TRY(sd_syevd_upper<T>(ctx, m, Tmp, m, Lam, &info));
    if (info != 0) {
      tm.end();
      fprintf(stderr, "indefinite_rayleigh_ritz: eigensolve failed (info=%d)\n", info);
      return 1;
    }
    LB2_TRY(sd_trsm_upper<T>(ctx, 'L', 'N', m, m, GA, m, Tmp, m));
    LB2_TRY(sd_indef_finalize<T>(ctx, m, Lam, Tmp, m, Z, m, Theta, dSig));
  }
  LB2_CUDA_OK(cudaMemcpyAsync(Eig, Theta, sizeof(R) * k, cudaMemcpyDeviceToDevice, ctx->stream));
  LB2_CUDA_OK(cudaMemcpyAsync(Cx, Z, sizeof(T) * (size_t)m * k, cudaMemcpyDeviceToDevice, ctx->stream));
  sig_len = m;
  if (!initial) {
    LB2_TRY(sd_cp_lower<T>(ctx, m, k, Cx, Cp));
    tm.end();
    if (last_quality == 5) {
      // indefinite_rr_modified_impl.inc:236-247: keep the accurate Cx for the residual, B-orthogonalise a copy for the basis
      if (!CxAcc) LB2_CUDA_OK(cudaMalloc(&CxAcc, sizeof(T) * 3 * (size_t)k * k));
      LB2_CUDA_OK(cudaMemcpyAsync(CxAcc, Cx, sizeof(T) * (size_t)m * k, cudaMemcpyDeviceToDevice, ctx->stream));
      LB2_TRY(svqb_mat_dev(m, k, Cx, G, eps_tol()));
    }
    LB2_TRY(ortho_indef_mat(m, k, k, Cp, Cx, G));
  } else {
    tm.end();
  }
  return 0;
}

// General projected pencil (G_A in DinvR, G_B in G, copy of G_B in Tmp): M = G_B^-1 G_A -> Xgeev -> theta = Re(w), eigenvectors
// B-normalised twice, signatures, quality check, signature sort: Z (sorted eigenvectors), Theta, dSig, last_quality.
// Mirrors src/rayleigh/indefinite_rr_modified_impl.inc:104-206 (GGEV replaced by LU + GEEV, see smalldense.cu).
template <typename T>
int Solver<T>::rr_indef_general(int m, bool initial) {
  using CT = typename ComplexOf<T>::type;
  const size_t mm = (size_t)3 * k * 3 * k;
  if (!geev_ws) {
    LB2_CUDA_OK(cudaMalloc(&geev_ws, sizeof(CT) * (2 * mm + 3 * (size_t)k)));
    LB2_CUDA_OK(cudaMalloc(&geev_piv, sizeof(int64_t) * 3 * (size_t)k));
    LB2_CUDA_OK(cudaMalloc(&geev_sig, 3 * (size_t)k));
  }
  CT* Mc = (CT*)geev_ws;
  CT* Vc = Mc + mm;
  CT* Wc = Vc + mm;
  int info = 0;
  LB2_TRY(sd_lu_solve<T>(ctx, m, Tmp, m, DinvR, m, m, geev_piv, &info));      // DinvR <- G_B^-1 G_A
  if (info != 0) {
    fprintf(stderr, "indefinite_rayleigh_ritz: S^H B S is singular (LU info=%d)\n", info);
    return 1;
  }
  LB2_TRY(sd_geev<T>(ctx, m, DinvR, m, Mc, Wc, Vc, &info));
  if (info != 0) {
    fprintf(stderr, "indefinite_rayleigh_ritz_modified: GGEV failed\n");
    return 1;
  }
  T* V = GA;   // eigenvectors (unsorted); G_A itself is no longer needed
  LB2_TRY(sd_geev_extract<T>(ctx, m, Wc, Vc, Lam, V, m));
  // B-normalise twice, signatures from the second pass
  for (int pass = 0; pass < 2; pass++) {
    LB2_TRY(sd_gemm<T>(ctx, 'N', m, m, m, G, m, V, m, Tmp, m));
    LB2_TRY(sd_gemm<T>(ctx, 'H', m, m, m, V, m, Tmp, m, DinvR, m));
    LB2_TRY(sd_bnormalize<T>(ctx, m, V, m, DinvR, m, pass == 1 ? geev_sig : nullptr));
  }
  // quality: ||V^H G_B V - I_sig||_F <= 1e-12 ||V||_F ||G_B V||_F
  LB2_TRY(sd_gemm<T>(ctx, 'N', m, m, m, G, m, V, m, Tmp, m));
  LB2_TRY(sd_gemm<T>(ctx, 'H', m, m, m, V, m, Tmp, m, DinvR, m));
  LB2_TRY(sd_indef_quality<T>(ctx, m, DinvR, V, Tmp, Scal + 4));
  LB2_TRY(d2h(hbuf, Scal + 4, 3 * sizeof(R)));
  tm.end();            // (sync() collects the finished timer records; none may be open)
  LB2_TRY(sync());
  tm.begin(PH_SMALL);
  const R eerr = hbuf[0], cerr = hbuf[1], bcerr = hbuf[2];
  const bool quality_ok = (bcerr < (R)1e-30) || (eerr <= (R)1e-12 * cerr * bcerr);
  last_quality = (quality_ok || initial) ? 1 : 5;   // the initial RR has no quality flag (indefinite_rr_impl.inc:51-149)
  LB2_TRY(sd_indef_sort<T>(ctx, m, Lam, geev_sig, V, m, Z, m, Theta, dSig));
  general_rr_calls++;
  return 0;
}

template <typename T>
void Solver<T>::print_state(bool header) {
  if (alg->verbosity <= 0) return;
  if (header)
    printf("Iteration %" PRIu64 "\t Converged Eigenpairs: %" PRIu64 "/%" PRIu64 "\n", iter, conv, (uint64_t)nev);
  for (int i = 0; i < nev; i++)
    printf("Eigenvalue %" PRIu64 ": %.9e\t Residual Norm: %.5e\n", (uint64_t)i, (double)hEig[i], (double)hRes[i]);
  printf("\n");
}

// After X (slab[cur][:,0:k]) and Eig are final for this pass: AX, residual norms of the first nev columns,
// convergence count (contiguous prefix, lobpcg_impl.inc:224-228).  Leaves B X in wA when B != NULL.
template <typename T>
int Solver<T>::residual_pass(bool initial, const T* Xacc) {
  T* X = Xp();
  LB2_TRY(apply(opA, k, X, AS));
  const T* BX = X;
  res_AX = AS;
  if (Xacc) {
    // quality-5 pass of the indefinite solver (ilobpcg_impl.inc:228-256): the basis continues with X = S Cx_ortho, the residual
    // is taken from X_accurate = S Cx (accurate eigenvalue correspondence): A X_acc -> AS[:, k:2k], B X_acc -> wA
    LB2_TRY(apply(opA, k, Xacc, col(AS, k)));
    res_AX = col(AS, k);
    BX = Xacc;
    if (opB) { LB2_TRY(apply(opB, k, Xacc, wA)); BX = wA; }
  } else if (opB) {
    LB2_TRY(apply(opB, k, X, wA));
    BX = wA;
  }
  res_BX = BX;
  const T* AXr = res_AX;
  const bool monitor = gram_cache && !initial && !Xacc;
  if (monitor) {
    // same two (three with B) streams as the plain norm pass, plus the Rayleigh quotients x^H A x and x^H B x of the new
    // Ritz vectors: the drift monitor of the cached Gram blocks
    tm.begin(PH_RESID);
    int rc = residual_monitor<T>(ctx, n, nev, X, n, AXr, n, opB ? BX : (const T*)nullptr, n, Eig, Sums);
    tm.end();
    phase_work[PH_RESID] += (opB ? 3.0 : 2.0) * (double)n * nev * sizeof(T);
    phase_calls[PH_RESID]++;
    if (rc) return rc;
    if (reduce()) LB2_TRY(allreduce_sum(ctx, Sums, 3 * (size_t)nev, kDouble));
    LB2_TRY(d2h(hbuf, Sums, sizeof(R) * 3 * nev));
    LB2_TRY(d2h(hbuf + 3 * nev, Eig, sizeof(R) * k));
  } else {
    LB2_TRY(resid(nev, AXr, BX, Eig, nullptr, Sums));
    if (reduce()) LB2_TRY(allreduce_sum(ctx, Sums, nev, kDouble));
    LB2_TRY(d2h(hbuf, Sums, sizeof(R) * nev));
    LB2_TRY(d2h(hbuf + 3 * nev, Eig, sizeof(R) * k));
  }
  LB2_TRY(sync());
  for (int i = 0; i < k; i++) hEig[i] = hbuf[3 * nev + i];
  const R bn = BNorm > 0 ? BNorm : R(1);
  for (int i = 0; i < nev; i++) hRes[i] = std::sqrt(hbuf[i]) / (ANorm + std::fabs(hEig[i]) * bn);
  if (monitor) {
    double dev = 0;
    const double an = ANorm > 0 ? (double)ANorm : 1.0;
    for (int i = 0; i < nev; i++) {   // x^H A x = theta x^H B x and |x^H B x| = 1 (signature +-1 in the indefinite solver)
      const double xax = (double)hbuf[nev + i], xbx = (double)hbuf[2 * nev + i];
      dev = std::max(dev, std::fabs(xax - (double)hEig[i] * xbx) / an);
      dev = std::max(dev, std::fabs(std::fabs(xbx) - 1.0));
    }
    cache_monitor = dev;
    cache_monitor_max = std::max(cache_monitor_max, dev);
  }
  if (initial) {
    conv = 0;
  } else {
    conv = 0;
    for (int i = 0; i < nev; i++) {
      if (hRes[i] > alg->tol && i >= debug_min_conv) break;
      conv++;
    }
  }
  return 0;
}

template <typename T>
int Solver<T>::prepare() {
  if (prepared) return 0;
  param_error = true;    // cleared once the parameters have been accepted
  if (indef && !alg->B) {
    fprintf(stderr, "ilobpcg: B operator must not be NULL\n");
    return 1;
  }
  ng = (int64_t)alg->size;
  n = ng;
  row0 = 0;
  k = (int)alg->sizeSub;
  nev = (int)alg->nev;
  opA = alg->A; opB = alg->B; opT = alg->T_;
  // parameter validation, same messages as lobpcg_impl.inc:66-75
  if (!opA) { fprintf(stderr, "lobpcg: A operator must not be NULL\n"); return 1; }
  if (alg->nev > alg->sizeSub) {
    fprintf(stderr, "lobpcg: nev (%lu) > sizeSub (%lu)\n", (unsigned long)alg->nev, (unsigned long)alg->sizeSub);
    return 1;
  }
  if (3 * alg->sizeSub > alg->size) {
    fprintf(stderr, "lobpcg: 3*sizeSub (%lu) > size (%lu)\n", (unsigned long)(3 * alg->sizeSub), (unsigned long)alg->size);
    return 1;
  }
  if (alg->sizeSub > 12000) {   // svqb keeps one int per column in shared memory (smalldense.cu: svqb_transform_kernel)
    fprintf(stderr, "lobpcg: sizeSub (%lu) > 12000 is not supported\n", (unsigned long)alg->sizeSub);
    return 1;
  }
  const BuiltinOp* ba = builtin_of(opA);
  if (ba && ba->n != ba->n_global) {  // row-partitioned operator: local rows, equal slabs on every rank
    n = ba->n;
    row0 = ba->row0;
    nseg = (ba->kind == OP_BDG) ? 2 : 1;
    if (!ctx->comm) {
      fprintf(stderr, "lobpcg: row-partitioned operator needs a communicator (lb2_ctx_attach_comm)\n");
      return 1;
    }
  }
  // every operator must act on the same rows as the block vectors: the kernels size their launches from the
  // operator's own row count while the solver passes its leading dimension n (a mismatch would write past a block)
  {
    const LinOpRaw* ops[3] = {opA, opB, opT};
    const char* names[3] = {"A", "B", "T"};
    for (int q = 0; q < 3; q++) {
      const LinOpRaw* op = ops[q];
      if (!op) continue;
      const BuiltinOp* b = builtin_of(op);
      for (int depth = 0; b && depth < 2; depth++) {   // depth 1: inner operator of a polynomial preconditioner
        if (b->prefix != Sc<T>::prefix) {
          fprintf(stderr, "lobpcg: operator %s was built for type '%c', the solver runs type '%c'\n", names[q], b->prefix,
                  Sc<T>::prefix);
          return 1;
        }
        if (b->n != n || b->n_global != ng || (b->n != b->n_global && b->row0 != row0)) {
          fprintf(stderr, "lobpcg: operator %s has %lld local / %lld global rows (first row %lld), the solver %lld / %lld "
                          "(first row %lld)\n", names[q], (long long)b->n, (long long)b->n_global, (long long)b->row0,
                  (long long)n, (long long)ng, (long long)row0);
          return 1;
        }
        if (b->kind != OP_CHEB && b->kind != OP_DEVICE && b->device != ctx->device) {
          fprintf(stderr, "lobpcg: operator %s lives on device %d, the solver context on device %d\n", names[q], b->device,
                  ctx->device);
          return 1;
        }
        b = (b->kind == OP_CHEB) ? builtin_of(b->inner) : nullptr;
      }
      // foreign operators: the reference never reads rows/cols (callers may leave them 0); reject only a stated mismatch
      if (!builtin_of(op) && ((op->rows && op->rows != (uint64_t)ng) || (op->cols && op->cols != (uint64_t)ng))) {
        fprintf(stderr, "lobpcg: operator %s is %lu x %lu, size is %lu\n", names[q], (unsigned long)op->rows,
                (unsigned long)op->cols, (unsigned long)ng);
        return 1;
      }
    }
  }
  param_error = false;
  LB2_CUDA_OK(cudaSetDevice(ctx->device));
  if (sd_init(ctx)) return 1;
  LB2_TRY(alloc());
  LB2_CUDA_OK(cudaMemsetAsync(Scal, 0, sizeof(R) * 16, ctx->stream));
  prepared = true;
  return 0;
}

template <typename T>
int Solver<T>::init() {
  if (int rc = prepare()) return rc;

  // X0: device generator, or upload of alg->S[0 : n*k) (rows row0.. of every column when partitioned)
  cur = 0;
  T* X = Xp();
  if (use_device_x0) {
    LB2_TRY(fill_rows(X, k, device_seed));
  } else if (dev_x0) {
    LB2_CUDA_OK(cudaMemcpyAsync(X, dev_x0, sizeof(T) * (size_t)n * k, cudaMemcpyDeviceToDevice, ctx->stream));
  } else if (n == ng) {   // whole columns: one contiguous block, pipelined through pinned chunks (hostcopy.cu)
    LB2_TRY(host_copy(ctx, X, alg->S, sizeof(T) * (size_t)n * k, true));
  } else {   // this rank's rows of every column: strided on the host, pipelined through the pinned ring (hostcopy.cu)
    for (int sg = 0; sg < nseg; sg++)
      LB2_TRY(host_copy_2d(ctx, X + sg * seg_len(), sizeof(T) * n, alg->S + seg_global(sg), sizeof(T) * ng,
                           sizeof(T) * seg_len(), k, true));
  }
  LB2_TRY(estimate_norm(opA, 0xA5EEDULL, &ANorm));
  if (opB) LB2_TRY(estimate_norm(opB, 0xB5EEDULL, &BNorm));
  else BNorm = 1;
  if (alg->verbosity > 0) printf("F-Norm: %.5e %.5e\n", (double)ANorm, (double)BNorm);

  LB2_TRY(sumsq_total(k, X, Scal));
  LB2_TRY(d2h(hbuf, Scal, sizeof(R)));
  LB2_TRY(sync());
  if (std::sqrt(hbuf[0]) < (R)EpsTol<T>::v) LB2_TRY(fill_rows(X, k, 0xC0FFEEULL));

  if (indef) {
    // ilobpcg_impl.inc:100-113: B-orthonormalise X (svqb, no dropping), indefinite RR, X <- X Cx
    int keep = k;
    LB2_TRY(svqb(X, k, (R)EpsTol<T>::v, false, &keep));
    if (int rc = rr_indef(k, 0, true)) return rc;
    T* Xn = slab[1 - cur];
    LB2_TRY(nn(k, k, make<T>(1), Xp(), Cx, k, zero<T>(), Xn));
    cur = 1 - cur;
  } else {
    if (int rc = rr_initial()) return rc;
  }
  LB2_TRY(residual_pass(true));
  // W for all k columns (iter 0 uses sizeW = sizeSub, lobpcg_impl.inc:134), preconditioned
  np = 0;
  nw = k;
  {
    T* X2 = Xp();
    const T* BX = opB ? wA : X2;
    T* Wdst = col(X2, k);
    LB2_TRY(resid(k, AS, BX, Eig, opT ? wB : Wdst, nullptr));
    if (opT) LB2_TRY(apply(opT, k, wB, Wdst));
  }
  print_state(false);
  useOrtho = 0;
  conv = 0;
  iter = 0;
  done = false;
  inited = true;
  return 0;
}

template <typename T>
int Solver<T>::step(int max_steps) {
  if (!inited) return -1;
  int passes = 0;
  const int rc = step_impl(max_steps, &passes);
  if (rc != 0) return rc > 0 ? -rc : rc;
  return passes;
}

template <typename T>
int Solver<T>::step_impl(int max_steps, int* passes_out) {
  int& passes = *passes_out;
  while (!done && passes < max_steps && iter < alg->maxIter) {
    ctx->oz_reuse = false;   // int8 path: the slices of the previous pass's basis are stale
    T* S = Xp();
    T* V = S;
    T* W = col(S, k + np);
    if (force_ortho && !indef && useOrtho == 0) useOrtho = 1;
    // orthogonalise W against [X, P_act]
    if (useOrtho || indef) {
      int keep = nw;
      LB2_TRY(ortho_drop(W, nw, V, k + np, &keep, indef));
      nw = keep;
    }
    int m = k + np + nw;
    if (indef) {
      if (int rc = rr_indef(m, k, false)) return rc;
    } else if (int rc = rr_modified(m)) {
      return rc;
    }
    if (!indef && useOrtho == 2) {
      useOrtho = 1;
      int keep = nw;
      LB2_TRY(ortho_drop(W, nw, V, k + np, &keep));
      nw = keep;
      m = k + np + nw;
      if (int rc = rr_modified(m)) return rc;
    }
    // X_new = S Cx into the other slab
    T* Sn = slab[1 - cur];
    const T* Xacc = nullptr;
    if (indef && last_quality == 5) {   // X_accurate = S Cx for the residual; the basis continues with S Cx_ortho (now in Cx)
      LB2_TRY(nn(m, k, make<T>(1), S, CxAcc, m, zero<T>(), wB));
      Xacc = wB;
      quality5_passes++;
    }
    LB2_TRY(nn(m, k, make<T>(1), S, Cx, m, zero<T>(), Sn));
    cur = 1 - cur;
    LB2_TRY(residual_pass(false, Xacc));
    if (alg->verbosity > 0) print_state(true);
    passes++;
    if (conv == (uint64_t)nev) {
      done = true;
      break;
    }
    // P_act and W_act only for the unconverged columns, written at their compacted positions
    const int nconv = (int)conv;
    const int nact = k - nconv;
    T* Sold = slab[1 - cur];
    LB2_TRY(nn(m, nact, make<T>(1), Sold, Cp + (size_t)nconv * m, m, zero<T>(), col(Sn, k)));
    if (!indef) LB2_TRY(update_gram_cache(m, nconv));
    {
      T* Wdst = col(Sn, k + nact);
      LB2_TRY(resid(nact, res_AX + (int64_t)nconv * n, res_BX + (int64_t)nconv * n, Eig + nconv, opT ? wB : Wdst, nullptr));
      if (opT) LB2_TRY(apply(opT, nact, wB, Wdst));
    }
    np = nact;
    nw = nact;
    iter++;
  }
  LB2_TRY(sync());
  return 0;
}

template <typename T>
int Solver<T>::finish() {
  if (!inited) return 1;
  T* X = Xp();
  if (dev_xout) {
    LB2_CUDA_OK(cudaMemcpyAsync(dev_xout, X, sizeof(T) * (size_t)n * k, cudaMemcpyDeviceToDevice, ctx->stream));
  } else if (n == ng) {
    LB2_TRY(host_copy(ctx, alg->S, X, sizeof(T) * (size_t)n * k, false));
  } else {
    for (int sg = 0; sg < nseg; sg++)
      LB2_TRY(host_copy_2d(ctx, X + sg * seg_len(), sizeof(T) * n, alg->S + seg_global(sg), sizeof(T) * ng,
                           sizeof(T) * seg_len(), k, false));
  }
  LB2_TRY(sync());
  for (int i = 0; i < k; i++) alg->eigVals[i] = hEig[i];
  for (int i = 0; i < nev; i++) alg->resNorm[i] = hRes[i];
  if (indef && alg->signature && sig_len > 0) {
    LB2_CUDA_OK(cudaMemcpyAsync(alg->signature, dSig, (size_t)sig_len, cudaMemcpyDeviceToHost, ctx->stream));
    LB2_TRY(sync());
  }
  alg->converged = conv;
  alg->iter = iter;
  return 0;
}

SolverBase* make_solver(lb2_ctx* ctx, char prefix, void* alg, int indefinite) {
  switch (prefix) {
    case 's': return new Solver<float>(ctx, (State<float>*)alg, indefinite != 0);
    case 'd': return new Solver<double>(ctx, (State<double>*)alg, indefinite != 0);
    case 'c': return new Solver<c32>(ctx, (State<c32>*)alg, indefinite != 0);
    case 'z': return new Solver<c64>(ctx, (State<c64>*)alg, indefinite != 0);
  }
  return nullptr;
}

// =======================================================================================================
// Host-buffer helper API: the reference's L2-L4 helpers (reference lobpcg.h:98-555) with their original
// signatures, so that unit-level callers link unchanged.  Every call uploads its operands, runs the same device
// building blocks the solver uses, and downloads the result; the host workspace arguments (wrk*, rr_*) are
// accepted and ignored.  These are convenience entry points — the solver itself never leaves the device.
// =======================================================================================================
template <typename T>
struct HelperEnv {
  State<T> st;
  Solver<T>* s = nullptr;
  bool ok = false;
  HelperEnv(int64_t rows, int kmax, const LinOpRaw* A, const LinOpRaw* B) {
    memset(&st, 0, sizeof(st));
    lb2_ctx* ctx = lb2_default_ctx();
    if (!ctx) return;
    st.size = (uint64_t)rows;
    st.sizeSub = st.nev = (uint64_t)(kmax > 0 ? kmax : 1);
    s = new Solver<T>(ctx, &st, false);
    ok = (s->helper_setup(rows, kmax, A, B) == 0);
  }
  ~HelperEnv() { delete s; }
};

template <typename T>
static void h_apply_block_op(const LinOpRaw* op, T* X, T* Y, uint64_t n, uint64_t kc) {
  const BuiltinOp* b = builtin_of(op);
  if (!b) {  // foreign host operator: exactly the reference's loop (src/gram/gram_impl.inc:29-33)
    for (uint64_t j = 0; j < kc; j++) op->matvec(op, X + j * n, Y + j * n);
    return;
  }
  HelperEnv<T> e((int64_t)n, (int)std::max<uint64_t>(kc, 1), op, nullptr);
  if (!e.ok) return;
  Solver<T>& s = *e.s;
  for (uint64_t c0 = 0; c0 < kc; c0 += 3 * (uint64_t)s.k) {
    const int w = (int)std::min<uint64_t>(3 * (uint64_t)s.k, kc - c0);
    if (s.up(s.slab[0], X + c0 * n, n, w) || s.apply(op, w, s.slab[0], s.slab[1]) || s.down(Y + c0 * n, s.slab[1], n, w)) return;
  }
}

template <typename T>
static void h_gram(T* V, uint64_t nv, T* U, uint64_t nu, uint64_t n, const LinOpRaw* B, T* G, bool self) {
  // gram_self: G = U^H B U (B == NULL: upper triangle only, as syrk/herk); gram_cross: G = V^H B U (full)
  const int kmax = (int)std::max<uint64_t>(std::max(nu, nv), 1);
  HelperEnv<T> e((int64_t)n, kmax, nullptr, B);
  if (!e.ok) return;
  Solver<T>& s = *e.s;
  T* dU = s.slab[0];
  T* dV = self ? dU : s.slab[1];
  if (s.up(dU, U, n, (int)nu)) return;
  if (!self && s.up(dV, V, n, (int)nv)) return;
  const T* BU = dU;
  if (B) { if (s.apply(B, (int)nu, dU, s.wA)) return; BU = s.wA; }
  const int rows = self ? (int)nu : (int)nv;
  if (s.gram_ar(rows, (int)nu, dV, BU, s.G, (self && !B) ? 1 : 0)) return;
  std::vector<T> h((size_t)rows * nu);
  if (s.down(h.data(), s.G, rows, (int)nu)) return;
  for (uint64_t j = 0; j < nu; j++)
    for (int i = 0; i < rows; i++)
      if (!(self && !B) || (uint64_t)i <= j) G[i + j * rows] = h[i + j * rows];
}

template <typename T>
static void h_get_residual(uint64_t n, uint64_t kc, T* X, T* AX, T* W, real_t<T>* eig, const LinOpRaw* A, const LinOpRaw* B) {
  HelperEnv<T> e((int64_t)n, (int)std::max<uint64_t>(kc, 1), A, B);
  if (!e.ok) return;
  Solver<T>& s = *e.s;
  T* dX = s.slab[0];
  if (s.up(dX, X, n, (int)kc)) return;
  if (AX) { if (s.up(s.AS, AX, n, (int)kc)) return; }
  else if (s.apply(A, (int)kc, dX, s.AS)) return;
  const T* BX = dX;
  if (B) { if (s.apply(B, (int)kc, dX, s.wA)) return; BX = s.wA; }
  cudaMemcpyAsync(s.Eig, eig, sizeof(real_t<T>) * kc, cudaMemcpyHostToDevice, s.ctx->stream);
  if (s.resid((int)kc, s.AS, BX, s.Eig, s.slab[1], nullptr)) return;
  s.down(W, s.slab[1], n, (int)kc);
}

template <typename T>
static void h_get_residual_norm(uint64_t n, uint64_t nev, T* W, real_t<T>* eig, real_t<T>* out, real_t<T> an, real_t<T> bn) {
  using R = real_t<T>;
  HelperEnv<T> e((int64_t)n, (int)std::max<uint64_t>(nev, 1), nullptr, nullptr);
  if (!e.ok) return;
  Solver<T>& s = *e.s;
  if (s.up(s.slab[0], W, n, (int)nev)) return;
  if (col_sumsq<T>(s.ctx, n, (int)nev, s.slab[0], n, s.Sums)) return;
  std::vector<R> ss(nev);
  cudaMemcpyAsync(ss.data(), s.Sums, sizeof(R) * nev, cudaMemcpyDeviceToHost, s.ctx->stream);
  cudaStreamSynchronize(s.ctx->stream);
  const R b = bn > 0 ? bn : R(1);
  for (uint64_t i = 0; i < nev; i++) out[i] = std::sqrt(ss[i]) / (an + std::fabs(eig[i]) * b);
}

template <typename T>
static uint64_t h_svqb(uint64_t m, uint64_t nc, real_t<T> tau, char drop, T* U, const LinOpRaw* B) {
  if (nc == 0) return 0;
  HelperEnv<T> e((int64_t)m, (int)nc, nullptr, B);
  if (!e.ok) return nc;
  Solver<T>& s = *e.s;
  int keep = (int)nc;
  if (s.up(s.slab[0], U, m, (int)nc) || s.svqb(s.slab[0], (int)nc, tau, drop == 'y', &keep) || s.down(U, s.slab[0], m, keep)) return nc;
  return (uint64_t)keep;
}

template <typename T>
static uint64_t h_ortho(uint64_t m, uint64_t nu, uint64_t nv, real_t<T> eps_ortho, T* U, T* V, const LinOpRaw* B, bool indefinite) {
  if (nu == 0 || nv == 0) return nu;
  HelperEnv<T> e((int64_t)m, (int)std::max(nu, nv), nullptr, B);
  if (!e.ok) return nu;
  Solver<T>& s = *e.s;
  s.eps_override = (double)eps_ortho;
  T* dV = s.slab[0];
  T* dU = s.slab[0] + (int64_t)nv * s.n;
  int keep = (int)nu;
  if (s.up(dV, V, m, (int)nv) || s.up(dU, U, m, (int)nu) || s.ortho_drop(dU, (int)nu, dV, (int)nv, &keep, indefinite) ||
      s.down(U, dU, m, keep))
    return nu;
  return (uint64_t)keep;
}

template <typename T>
static void h_rr(uint64_t n, uint64_t kc, T* S, T* Cx, real_t<T>* eig, const LinOpRaw* A, const LinOpRaw* B) {
  HelperEnv<T> e((int64_t)n, (int)kc, A, B);
  if (!e.ok) return;
  Solver<T>& s = *e.s;
  if (s.up(s.Xp(), S, n, (int)kc) || s.rr_initial()) return;
  s.down(Cx, s.Cx, kc, (int)kc);
  cudaMemcpyAsync(eig, s.Eig, sizeof(real_t<T>) * kc, cudaMemcpyDeviceToHost, s.ctx->stream);
  cudaStreamSynchronize(s.ctx->stream);
}

template <typename T>
static void h_rr_mod(uint64_t n, uint64_t nx, uint64_t mult, uint64_t nconv, uint64_t nretain, uint8_t* useOrtho, T* S,
                     const T* AX, T* Cx, T* Cp, real_t<T>* eig, const LinOpRaw* A, const LinOpRaw* B) {
  const int m = (int)((mult - 1) * nx + nretain - (mult == 3 ? nconv : 0));
  HelperEnv<T> e((int64_t)n, (int)nx, A, B);
  if (!e.ok) return;
  Solver<T>& s = *e.s;
  if (s.up(s.Xp(), S, n, m)) return;
  if (AX) { if (s.up(s.AS, AX, n, (int)nx)) return; }
  else if (s.apply(A, (int)nx, s.Xp(), s.AS)) return;
  s.useOrtho = *useOrtho;
  s.np = (mult == 3) ? (int)(nx - nconv) : 0;   // S = [X | P_act | W]: the [X P] blocks are contracted from the tall vectors
  s.nw = m - (int)nx - s.np;
  s.cache_ok = false;
  if (s.nw < 0 || s.rr_modified(m)) return;
  *useOrtho = (uint8_t)s.useOrtho;
  if (s.useOrtho == 2) return;
  s.down(Cx, s.Cx, m, (int)nx);
  s.down(Cp, s.Cp, m, (int)nx);
  cudaMemcpyAsync(eig, s.Eig, sizeof(real_t<T>) * nx, cudaMemcpyDeviceToHost, s.ctx->stream);
  cudaStreamSynchronize(s.ctx->stream);
}

template <typename T>
static void h_irr(uint64_t n, uint64_t nx, int m, bool initial, T* S, const T* AX, T* Cx, T* Cp, T* Cx_ortho,
                  real_t<T>* eig, int8_t* sig, int* quality, const LinOpRaw* A, const LinOpRaw* B) {
  HelperEnv<T> e((int64_t)n, (int)nx, A, B);
  if (!e.ok) return;
  Solver<T>& s = *e.s;
  if (s.up(s.Xp(), S, n, m)) return;
  int from = 0;
  if (!initial) {
    from = (int)nx;
    if (AX) { if (s.up(s.AS, AX, n, (int)nx)) return; }
    else if (s.apply(A, (int)nx, s.Xp(), s.AS)) return;
  }
  if (s.rr_indef(m, from, initial)) return;
  const int ncx = initial ? m : (int)nx;
  if (initial) {
    s.down(Cx, s.Z, m, m);   // all eigenvectors, signature-sorted
    cudaMemcpyAsync(eig, s.Theta, sizeof(real_t<T>) * m, cudaMemcpyDeviceToHost, s.ctx->stream);
  } else {
    s.down(Cx, s.last_quality == 5 ? s.CxAcc : s.Cx, m, ncx);   // quality 5: Cx stays the accurate one, Cx_ortho is the B-orthogonalised copy
    s.down(Cp, s.Cp, m, ncx);
    if (Cx_ortho) s.down(Cx_ortho, s.Cx, m, ncx);
    cudaMemcpyAsync(eig, s.Eig, sizeof(real_t<T>) * nx, cudaMemcpyDeviceToHost, s.ctx->stream);
    if (quality) *quality = s.last_quality;
  }
  cudaMemcpyAsync(sig, s.dSig, (size_t)m, cudaMemcpyDeviceToHost, s.ctx->stream);
  cudaStreamSynchronize(s.ctx->stream);
}

// coefficient-space helpers: everything is small (m x m metric), rows = 3k >= m keeps the scratch large enough
template <typename T>
struct MatEnv {
  HelperEnv<T> e;
  T *dU = nullptr, *dV = nullptr, *dM = nullptr;
  MatEnv(uint64_t m, uint64_t nu, uint64_t nv) : e((int64_t)(3 * std::max<uint64_t>({nu, nv, (m + 2) / 3, 1})),
                                                     (int)std::max<uint64_t>({nu, nv, (m + 2) / 3, 1}), nullptr, nullptr) {
    if (!e.ok) return;
    Solver<T>& s = *e.s;
    dM = s.G;  dU = s.Cp;  dV = s.Cx;   // G: 3k x 3k >= m x m; Cx, Cp: 3k x k >= m x n
  }
};

template <typename T>
static uint64_t h_svqb_mat(uint64_t m, uint64_t nc, real_t<T> tau, T* U, T* mat) {
  if (nc == 0) return 0;
  MatEnv<T> me(m, nc, 0);
  if (!me.e.ok) return nc;
  Solver<T>& s = *me.e.s;
  if (s.up(me.dU, U, m, (int)nc) || (mat && s.up(me.dM, mat, m, (int)m))) return nc;
  if (s.svqb_mat_dev((int)m, (int)nc, me.dU, mat ? me.dM : nullptr, tau)) return nc;
  s.down(U, me.dU, m, (int)nc);
  return nc;
}

template <typename T>
static uint64_t h_ortho_indef_mat(uint64_t m, uint64_t nu, uint64_t nv, real_t<T> eps, T* U, T* V, T* mat) {
  if (nu == 0 || nv == 0) return nu;
  MatEnv<T> me(m, nu, nv);
  if (!me.e.ok) return 0;
  Solver<T>& s = *me.e.s;
  s.eps_override = (double)eps;
  if (s.up(me.dU, U, m, (int)nu) || s.up(me.dV, V, m, (int)nv) || s.up(me.dM, mat, m, (int)m)) return 0;
  if (s.ortho_indef_mat((int)m, (int)nu, (int)nv, me.dU, me.dV, me.dM)) return 0;
  s.down(U, me.dU, m, (int)nu);
  return 0;   // the reference returns 0 ("no column dropping implemented")
}

template <typename T>
static void h_gram_mat(T* V, uint64_t nv, T* U, uint64_t nu, uint64_t n, const T* mat, T* G, bool self) {
  MatEnv<T> me(n, nu, nv);
  if (!me.e.ok) return;
  Solver<T>& s = *me.e.s;
  T* dV = self ? me.dU : me.dV;
  if (s.up(me.dU, U, n, (int)nu) || (!self && s.up(dV, V, n, (int)nv))) return;
  const T* MU = me.dU;
  if (mat) {
    if (s.up(me.dM, mat, n, (int)n)) return;
    if (sd_gemm<T>(s.ctx, 'N', (int)n, (int)nu, (int)n, me.dM, (int)n, me.dU, (int)n, s.Tmp, (int)n)) return;
    MU = s.Tmp;
  }
  const int rows = self ? (int)nu : (int)nv;
  if (sd_gemm<T>(s.ctx, 'H', rows, (int)nu, (int)n, dV, (int)n, MU, (int)n, s.Z, rows)) return;
  std::vector<T> h((size_t)rows * nu);
  if (s.down(h.data(), s.Z, rows, (int)nu)) return;
  for (uint64_t j = 0; j < nu; j++)
    for (int i = 0; i < rows; i++)
      if (!(self && !mat) || (uint64_t)i <= j) G[i + j * rows] = h[i + j * rows];
}

template <typename T>
static void h_fill_random(uint64_t n, T* x) {
  static uint64_t calls = 0;
  HelperEnv<T> e((int64_t)n, 1, nullptr, nullptr);
  if (!e.ok) return;
  Solver<T>& s = *e.s;
  if (fill_uniform<T>(s.ctx, (int64_t)n, 1, s.slab[0], (int64_t)n, 0x5EEDULL + 7919ULL * calls++, (int64_t)n, 0)) return;
  s.down(x, s.slab[0], n, 1);
}

template <typename T>
static real_t<T> h_estimate_norm(uint64_t n, const LinOpRaw* A) {
  HelperEnv<T> e((int64_t)n, 1, A, nullptr);
  real_t<T> out = 0;
  if (!e.ok) return out;
  e.s->estimate_norm(A, 0xA5EEDULL, &out);
  return out;
}

}  // namespace lb2

using namespace lb2;
extern "C" {
#define LB2_HELPERS(P, T, RT)                                                                                         \
  void P##_apply_block_op(const void* Op, T* X, T* Y, uint64_t n, uint64_t k) {                                        \
    h_apply_block_op<T>((const LinOpRaw*)Op, X, Y, n, k);                                                              \
  }                                                                                                                    \
  void P##_gram_self(T* U, uint64_t n, uint64_t k, const void* B, T* G, uint64_t ldg, T* wrk) {                        \
    (void)ldg; (void)wrk;                                                                                              \
    h_gram<T>(nullptr, 0, U, k, n, (const LinOpRaw*)B, G, true);                                                       \
  }                                                                                                                    \
  void P##_gram_cross(T* V, uint64_t nv, T* U, uint64_t nu, uint64_t n, const void* B, T* G, uint64_t ldg, T* wrk) {   \
    (void)ldg; (void)wrk;                                                                                              \
    h_gram<T>(V, nv, U, nu, n, (const LinOpRaw*)B, G, false);                                                          \
  }                                                                                                                    \
  void P##_gram_self_mat(T* U, uint64_t n, uint64_t k, const T* mat, T* G, uint64_t ldg, T* wrk) {                     \
    (void)ldg; (void)wrk;                                                                                              \
    h_gram_mat<T>(nullptr, 0, U, k, n, mat, G, true);                                                                  \
  }                                                                                                                    \
  void P##_gram_cross_mat(T* V, uint64_t nv, T* U, uint64_t nu, uint64_t n, const T* mat, T* G, uint64_t ldg, T* wrk) { \
    (void)ldg; (void)wrk;                                                                                              \
    h_gram_mat<T>(V, nv, U, nu, n, mat, G, false);                                                                     \
  }                                                                                                                    \
  void P##_get_residual(uint64_t size, uint64_t sizeSub, T* X, T* AX, T* R, RT* eigVal, T* wrk, void* A, void* B) {    \
    (void)wrk;                                                                                                         \
    h_get_residual<T>(size, sizeSub, X, AX, R, eigVal, (const LinOpRaw*)A, (const LinOpRaw*)B);                        \
  }                                                                                                                    \
  void P##_get_residual_norm(uint64_t size, uint64_t nev, T* W, RT* eigVals, RT* resNorm, T* w1, T* w2, T* w3,         \
                             RT ANorm, RT BNorm, void* B) {                                                            \
    (void)w1; (void)w2; (void)w3; (void)B;                                                                             \
    h_get_residual_norm<T>(size, nev, W, eigVals, resNorm, ANorm, BNorm);                                              \
  }                                                                                                                    \
  uint64_t P##_svqb(uint64_t m, uint64_t n, RT tau, char drop, T* U, T* w1, T* w2, T* w3, void* B) {                   \
    (void)w1; (void)w2; (void)w3;                                                                                      \
    return h_svqb<T>(m, n, tau, drop, U, (const LinOpRaw*)B);                                                          \
  }                                                                                                                    \
  uint64_t P##_svqb_mat(uint64_t m, uint64_t n, RT tau, char drop, T* U, T* mat, T* w1, T* w2, T* w3) {                \
    (void)drop; (void)w1; (void)w2; (void)w3;                                                                          \
    return h_svqb_mat<T>(m, n, tau, U, mat);                                                                           \
  }                                                                                                                    \
  uint64_t P##_ortho_drop(uint64_t m, uint64_t n_u, uint64_t n_v, RT eps_ortho, RT eps_drop, T* U, T* V, T* w1, T* w2, \
                          T* w3, void* B) {                                                                            \
    (void)eps_drop; (void)w1; (void)w2; (void)w3;                                                                      \
    return h_ortho<T>(m, n_u, n_v, eps_ortho, U, V, (const LinOpRaw*)B, false);                                        \
  }                                                                                                                    \
  uint64_t P##_ortho_indefinite(uint64_t m, uint64_t n_u, uint64_t n_v, RT eps_ortho, RT eps_drop, T* U, T* V, T* sig, \
                                T* w1, T* w2, T* w3, void* B) {                                                        \
    (void)eps_drop; (void)sig; (void)w1; (void)w2; (void)w3; /* sig is recomputed as V^H B V */                        \
    return h_ortho<T>(m, n_u, n_v, eps_ortho, U, V, (const LinOpRaw*)B, true);                                         \
  }                                                                                                                    \
  uint64_t P##_ortho_indefinite_mat(uint64_t m, uint64_t n_u, uint64_t n_v, RT eps_ortho, RT eps_drop, T* U, T* V,     \
                                    T* mat, T* w1, T* w2, T* w3) {                                                     \
    (void)eps_drop; (void)w1; (void)w2; (void)w3;                                                                      \
    return h_ortho_indef_mat<T>(m, n_u, n_v, eps_ortho, U, V, mat);                                                    \
  }                                                                                                                    \
  void P##_rayleigh_ritz(uint64_t size, uint64_t sizeSub, T* S, T* Cx, RT* eigVal, T* w1, T* w2, T* w3, RT* rr_D,      \
                         void* A, void* B) {                                                                           \
    (void)w1; (void)w2; (void)w3; (void)rr_D;                                                                          \
    h_rr<T>(size, sizeSub, S, Cx, eigVal, (const LinOpRaw*)A, (const LinOpRaw*)B);                                     \
  }                                                                                                                    \
  void P##_rayleigh_ritz_modified(uint64_t size, uint64_t nx, uint64_t mult, uint64_t nconv, uint64_t nretain,         \
                                  uint8_t* useOrtho, T* S, const T* AX, T* w1, T* w2, T* w3, T* Cx, T* Cp, RT* eigVal, \
                                  RT* rr_eigvals, T* rr_tau, RT* rr_D, void* A, void* B) {                             \
    (void)w1; (void)w2; (void)w3; (void)rr_eigvals; (void)rr_tau; (void)rr_D;                                          \
    h_rr_mod<T>(size, nx, mult, nconv, nretain, useOrtho, S, AX, Cx, Cp, eigVal, (const LinOpRaw*)A,                   \
                (const LinOpRaw*)B);                                                                                   \
  }                                                                                                                    \
  void P##_indefinite_rayleigh_ritz(uint64_t size, uint64_t sizeSub, T* S, T* Cx, RT* eigVal, int8_t* signature,       \
                                    T* w1, T* w2, T* w3, T* w4, uint64_t* rr_indices, T* rr_ggev, void* A, void* B) {  \
    (void)w1; (void)w2; (void)w3; (void)w4; (void)rr_indices; (void)rr_ggev;                                           \
    h_irr<T>(size, sizeSub, (int)sizeSub, true, S, nullptr, Cx, nullptr, nullptr, eigVal, signature, nullptr,          \
             (const LinOpRaw*)A, (const LinOpRaw*)B);                                                                  \
  }                                                                                                                    \
  void P##_indefinite_rayleigh_ritz_modified(uint64_t size, uint64_t nx, uint64_t mult, uint64_t nconv,                \
                                             uint64_t nretain, T* S, const T* AX, T* w1, T* w2, T* w3, T* w4, T* Cx,   \
                                             T* Cp, T* Cx_ortho, RT* eigVal, int8_t* signature, int* quality_flag,     \
                                             RT* rr_eigvals, int8_t* rr_sig, uint64_t* rr_indices, T* rr_VR,           \
                                             T* rr_ggev, void* A, void* B) {                                           \
    (void)w1; (void)w2; (void)w3; (void)w4; (void)rr_eigvals; (void)rr_sig; (void)rr_indices; (void)rr_VR;             \
    (void)rr_ggev;                                                                                                     \
    const int m = (int)((mult - 1) * nx + nretain - (mult == 3 ? nconv : 0));                                          \
    h_irr<T>(size, nx, m, false, S, AX, Cx, Cp, Cx_ortho, eigVal, signature, quality_flag, (const LinOpRaw*)A,         \
             (const LinOpRaw*)B);                                                                                      \
  }                                                                                                                    \
  void P##_fill_random(uint64_t n, T* x) { h_fill_random<T>(n, x); }                                                   \
  RT P##_estimate_norm(uint64_t size, void* A, T* w1, T* w2) {                                                         \
    (void)w1; (void)w2;                                                                                                \
    return h_estimate_norm<T>(size, (const LinOpRaw*)A);                                                               \
  }

LB2_HELPERS(s, float, float)
LB2_HELPERS(d, double, double)
LB2_HELPERS(c, c32, float)
LB2_HELPERS(z, c64, double)
}  // extern "C"

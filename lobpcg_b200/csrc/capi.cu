// lobpcg_b200/csrc/capi.cu — the C ABI declared in include/lobpcg_b200.h and include/lobpcg.h.
#include <chrono>
#include <climits>
#include <cmath>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "context.h"
#include "kernels.h"
#include "solver.h"
#include "../../include/lobpcg_b200.h"

using namespace lb2;

namespace lb2 {
int mg_requested_gpus();                                                   // multigpu.cu
int run_solver_multi(char prefix, void* alg, int indefinite, int want);
}
extern "C" void lb2_note_single_gpu_call(void);

namespace lb2 {
static std::mutex g_ctx_mu;
static std::vector<lb2_ctx*> g_all_ctx;   // every live context (their cached arenas are freed when an allocation fails)

static void trim_ctx(lb2_ctx* c) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (c->arena_cache) {
    cudaSetDevice(c->device);
    cudaFree(c->arena_cache);
  }
  c->arena_cache = nullptr;
  c->arena_cache_bytes = 0;
  if (c->oz_buf && c->active_solvers == 0) {   // int8 slice buffer of a context with no solve in flight (gram_i8.cu)
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->oz_buf);
    c->oz_buf = nullptr;
    c->oz_bytes = 0;
    c->oz_tag_ptr = nullptr;
  }
  cudaSetDevice(dev);
}

void* oz_malloc(lb2_ctx* ctx, size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) == cudaSuccess) return p;
  cudaGetLastError();
  {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (lb2_ctx* c : g_all_ctx)
      if (c != ctx) trim_ctx(c);
    if (ctx->arena_cache) {   // this context's own cached arena (no solve in flight on it uses the cache)
      cudaFree(ctx->arena_cache);
      ctx->arena_cache = nullptr;
      ctx->arena_cache_bytes = 0;
    }
  }
  if (cudaMalloc(&p, bytes) == cudaSuccess) return p;
  cudaGetLastError();
  return nullptr;
}

void* arena_alloc(lb2_ctx* ctx, size_t bytes) {
  {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (ctx->arena_cache && ctx->arena_cache_bytes >= bytes && ctx->arena_cache_bytes <= bytes + bytes / 4) {
      void* p = ctx->arena_cache;
      ctx->arena_cache = nullptr;
      ctx->arena_cache_bytes = 0;
      return p;
    }
    trim_ctx(ctx);
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {   // another context may be sitting on a cached arena: free them all and retry once
    cudaGetLastError();
    {
      std::lock_guard<std::mutex> lk(g_ctx_mu);
      for (lb2_ctx* c : g_all_ctx) trim_ctx(c);
    }
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) {
    fprintf(stderr, "lobpcg_b200: cannot allocate the %.2f GB solver arena (%s)\n", bytes / 1e9, cudaGetErrorString(e));
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void arena_release(lb2_ctx* ctx, void* p, size_t bytes) {
  if (!p) return;
  static const bool keep = [] { const char* e = getenv("LB2_ARENA_CACHE"); return !(e && atoi(e) == 0); }();
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  if (keep && !ctx->arena_cache) {
    ctx->arena_cache = p;
    ctx->arena_cache_bytes = bytes;
    return;
  }
  cudaFree(p);
}

void* pinned_take(lb2_ctx* ctx, size_t bytes) {
  {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (ctx->pinned_cache && ctx->pinned_cache_bytes >= bytes) {
      void* p = ctx->pinned_cache;
      ctx->pinned_cache = nullptr;
      ctx->pinned_cache_bytes = 0;
      return p;
    }
  }
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void pinned_give(lb2_ctx* ctx, void* p, size_t bytes) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (!ctx->pinned_cache) { ctx->pinned_cache = p; ctx->pinned_cache_bytes = bytes; return; }
  }
  cudaFreeHost(p);
}

void* ctx_scratch(lb2_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return ctx->ws;
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return nullptr;
  if (ctx->ws) cudaFree(ctx->ws);
  ctx->ws = nullptr;
  ctx->ws_bytes = 0;
  const size_t want = bytes + (bytes >> 1) + (1 << 20);
  cudaError_t e = cudaMalloc(&ctx->ws, want);
  if (e != cudaSuccess) {
    fprintf(stderr, "lobpcg_b200: cannot allocate %zu bytes of scratch (%s)\n", want, cudaGetErrorString(e));
    return nullptr;
  }
  ctx->ws_bytes = want;
  return ctx->ws;
}
}  // namespace lb2

// Structure detection for CSR input: if the matrix is exactly a Dirichlet 3/5/7-point stencil in natural ordering
// (offsets {0,+-1,+-gx,+-gx*gy}, neighbour present iff inside the grid, one constant off-diagonal value, arbitrary
// real diagonal) it is applied by the matrix-free stencil kernel with the diagonal as `potential`: per column that
// reads 8 n bytes of diagonal (L2-resident across columns) instead of gathering 7 X values per row through L2.
// Anything else (or LB2_CSR_NO_STENCIL_DETECT=1) takes the general CSR kernel.
template <typename T>
static bool detect_stencil(int64_t n, const int64_t* rp, const int32_t* col, const T* val, int64_t& gx, int64_t& gy,
                           int64_t& gz, double& coff, std::vector<real_t<T>>& diag) {
  if (getenv("LB2_CSR_NO_STENCIL_DETECT")) return false;
  if (n < 2) return false;
  int64_t a = 0, b = 0;   // second and third positive offsets
  for (int64_t p = rp[0]; p < rp[1]; p++) {   // row 0 of a stencil: neighbours +1, +gx, +gx*gy
    const int64_t o = col[p];
    if (o == 0 || o == 1) continue;
    if (!a) a = o; else if (!b) b = o; else return false;
  }
  gx = a ? a : n;
  if (n % gx) return false;
  gy = b ? b / gx : (a ? n / gx : 1);
  if (b && (b % gx)) return false;
  if (gy < 1 || (n / gx) % gy) return false;
  gz = n / (gx * gy);
  if (gx < 2 || gx * gy * gz != n) return false;
  diag.assign((size_t)n, real_t<T>(0));
  bool have_c = false;
  double c = 0;
  for (int64_t i = 0; i < n; i++) {
    const int64_t x = i % gx, y = (i / gx) % gy, z = i / (gx * gy);
    const int64_t exp_off[6] = {-gx * gy, -gx, -1, 1, gx, gx * gy};
    const bool exp_ok[6] = {z > 0, y > 0, x > 0, x + 1 < gx, y + 1 < gy, z + 1 < gz};
    int64_t p = rp[i];
    bool saw_diag = false;
    for (int e = 0; e < 6; e++) {
      // entries are ascending in column: the diagonal sits between offset -1 and +1
      if (e == 3) {
        if (p < rp[i + 1] && col[p] == i) {
          if (Sc<T>::cplx && ((const real_t<T>*)&val[p])[1] != 0) return false;
          diag[(size_t)i] = real_(val[p]);
          saw_diag = true;
          p++;
        }
      }
      if (!exp_ok[e]) continue;
      if (p >= rp[i + 1] || col[p] != i + exp_off[e]) return false;
      if (Sc<T>::cplx && ((const real_t<T>*)&val[p])[1] != 0) return false;
      const double v = (double)real_(val[p]);
      if (!have_c) { c = v; have_c = true; }
      else if (v != c) return false;
      p++;
    }
    if (p != rp[i + 1]) return false;
    (void)saw_diag;
  }
  if (!have_c) return false;
  coff = c;
  return true;
}

extern "C" {

int lb2_gram_wl_plan_sharing(int ma, int mb, int upper, int64_t n, int ncta, int bk, int phase, int window, int samples,
                             double* share) {
  return lb2::gram_wl_plan_sharing(ma, mb, upper, n, ncta, bk, phase, window, samples, share);
}
int lb2_gram_wl_cols_plan_check(int m, int nw, int nprod, int tri_c0, int64_t n, int ncta, int bk, double* stats) {
  return lb2::gram_wl_cols_plan_check(m, nw, nprod, tri_c0, n, ncta, bk, stats);
}
int lb2_gram_wl_plan_check(int ma, int mb, int upper, int64_t n, int ncta, int bk, double* stats) {
  return lb2::gram_wl_plan_check(ma, mb, upper, n, ncta, bk, stats);
}

const char* lb2_version(void) { return "lobpcg_b200 0.1 (sm_100a)"; }

lb2_ctx* lb2_ctx_create(int device, void* cuda_stream) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "lobpcg_b200: no CUDA device available — this library has no CPU fallback\n");
    return nullptr;
  }
  if (device < 0) cudaGetDevice(&device);
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  lb2_ctx* c = new lb2_ctx();
  c->device = device;
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (cuda_stream) {
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete c;
      return nullptr;
    }
    c->own_stream = true;
  }
  {
    std::lock_guard<std::mutex> lk(lb2::g_ctx_mu);
    lb2::g_all_ctx.push_back(c);
  }
  return c;
}

// release the memory this context keeps for reuse (the solver arena of the last solve)
int lb2_ctx_trim(lb2_ctx* c) {
  if (!c) return -1;
  std::lock_guard<std::mutex> lk(lb2::g_ctx_mu);
  lb2::trim_ctx(c);
  return 0;
}

void lb2_ctx_destroy(lb2_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  {
    std::lock_guard<std::mutex> lk(lb2::g_ctx_mu);
    lb2::trim_ctx(c);
    if (c->pinned_cache) cudaFreeHost(c->pinned_cache);
    c->pinned_cache = nullptr;
    auto& v = lb2::g_all_ctx;
    v.erase(std::remove(v.begin(), v.end(), c), v.end());
  }
  if (c->cublas) cublasDestroy(c->cublas);
  if (c->cusolver) cusolverDnDestroy(c->cusolver);
  lb2::gram_wl_cache_free(c);
  lb2::hostcopy_free(c);
  if (c->ws) cudaFree(c->ws);
  if (c->oz_buf) cudaFree(c->oz_buf);
  if (c->oz_hint) cudaFree(c->oz_hint);
  for (auto& e : c->oz_ev) if (e) cudaEventDestroy(e);
  if (c->solver_ws) cudaFree(c->solver_ws);
  if (c->solver_hws) free(c->solver_hws);
  if (c->dev_info) cudaFree(c->dev_info);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

int lb2_ctx_sync(lb2_ctx* c) {
  LB2_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

int lb2_ctx_set_option(lb2_ctx* c, const char* key, int value) {
  if (!c || !key) return -1;
  if (!strcmp(key, "gram_tile")) c->gram_tile = value;
  else if (!strcmp(key, "nn_tile")) c->nn_tile = value;
  else if (!strcmp(key, "nn_bk")) c->nn_bk = value;
  else if (!strcmp(key, "nn_persist")) c->nn_persist = value;
  else if (!strcmp(key, "nn_stagger")) c->nn_stagger = value;
  else if (!strcmp(key, "hostcopy_threads")) lb2::hostcopy_set_threads(c, value);
  else if (!strcmp(key, "gram_merge")) c->gram_merge = value;
  else if (!strcmp(key, "nn_warps")) c->nn_warps = value;
  else if (!strcmp(key, "force_simt")) c->force_simt = value;
  else if (!strcmp(key, "gram_wl")) c->gram_wl = value;
  else if (!strcmp(key, "gram_bk")) c->gram_bk = value;
  else if (!strcmp(key, "gram_tc5")) c->gram_tc5 = value;
  else if (!strcmp(key, "gram_i8")) c->gram_i8 = value;
  else if (!strcmp(key, "oz_load_pct")) c->oz_load_pct = value;
  else if (!strcmp(key, "oz_cluster")) c->oz_cluster = value;
  else if (!strcmp(key, "oz_lockstep")) c->oz_lockstep = value;
  else if (!strcmp(key, "oz_prefetch")) c->oz_prefetch = value;
  else if (!strcmp(key, "oz_hints")) c->oz_hints = value;
  else if (!strcmp(key, "oz_ring")) c->oz_ring = value;
  else if (!strcmp(key, "oz_nn_ring")) c->oz_nn_ring = value;
  else if (!strcmp(key, "oz_reuse")) c->oz_reuse = value != 0;   // what the solver sets between a Gram and the projections from the same block (tests)
  else if (!strcmp(key, "nn_i8")) c->nn_i8 = value;
  else if (!strcmp(key, "oz_clusters")) c->oz_clusters = value;
  else if (!strcmp(key, "gram_tma")) c->gram_tma = value;
  else if (!strcmp(key, "gram_load_pct")) c->gram_load_pct = value;
  else if (!strcmp(key, "gram_phase")) c->gram_phase = value;
  else if (!strcmp(key, "gram_strip_max")) c->gram_strip_max = value;
  else if (!strcmp(key, "gram_strip_fma")) c->gram_strip_fma = value;
  else if (!strcmp(key, "spmm_cols")) c->spmm_cols = value;
  else if (!strcmp(key, "csr_window")) c->csr_window = value;
  else if (!strcmp(key, "csr_staged")) c->csr_staged = value;
  else if (!strcmp(key, "csr_order")) c->csr_order = value;
  else if (!strcmp(key, "csr_pipe")) c->csr_pipe = value;
  else if (!strcmp(key, "csr_lpr")) c->csr_lpr = value;
  else return -1;
  return 0;
}

unsigned long long lb2_ctx_launches(lb2_ctx* c) { return c ? c->launches : 0ULL; }
int lb2_oz_plan_check(int m, int nw, int nprod, int tri_c0, int64_t n, int nworkers, int mode, double* stats) {
  return lb2::oz_plan_check(m, nw, nprod, tri_c0, n, nworkers, mode, stats);
}
int lb2_ctx_oz_stats(lb2_ctx* c, double* out4) { return (c && out4) ? lb2::oz_stats_query(c, out4) : -1; }

static std::mutex g_mu;
static std::map<int, lb2_ctx*> g_default;
lb2_ctx* lb2_default_ctx(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    fprintf(stderr, "lobpcg_b200: no CUDA device available — this library has no CPU fallback\n");
    return nullptr;
  }
  auto it = g_default.find(dev);
  if (it != g_default.end()) return it->second;
  lb2_ctx* c = lb2_ctx_create(dev, nullptr);
  g_default[dev] = c;
  return c;
}

void* lb2_malloc(size_t bytes) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
  if (e != cudaSuccess) {   // cached solver arenas (arena_release) give way to any other allocation of the library
    cudaGetLastError();
    {
      std::lock_guard<std::mutex> lk(lb2::g_ctx_mu);
      for (lb2_ctx* c : lb2::g_all_ctx) lb2::trim_ctx(c);
    }
    e = cudaMalloc(&p, bytes ? bytes : 1);
  }
  if (e != cudaSuccess) {
    fprintf(stderr, "lobpcg_b200: cudaMalloc(%zu) failed: %s\n", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  return p;
}
void lb2_free(void* p) { if (p) cudaFree(p); }
void* lb2_malloc_host(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
  return p;
}
void lb2_free_host(void* p) { if (p) cudaFreeHost(p); }
// pageable host memory <-> device; large blocks are pipelined through a pinned ring by several host threads (hostcopy.cu)
int lb2_memcpy_h2d(lb2_ctx* c, void* dst, const void* src, size_t bytes) { return lb2::host_copy(c, dst, src, bytes, true); }
int lb2_memcpy_d2h(lb2_ctx* c, void* dst, const void* src, size_t bytes) { return lb2::host_copy(c, dst, src, bytes, false); }
int lb2_memset(lb2_ctx* c, void* dst, int byte, size_t bytes) {
  LB2_CUDA_OK(cudaMemsetAsync(dst, byte, bytes, c->stream));
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// built-in operators
// ---------------------------------------------------------------------------------------------------
static size_t scalar_size(char p) { return p == 's' ? 4 : (p == 'd' || p == 'c') ? 8 : 16; }
static size_t real_size(char p) { return (p == 's' || p == 'c') ? 4 : 8; }

// host-pointer matvec so that reference-style callers (linop_apply on host vectors) keep working
static void builtin_host_matvec(const LinOpRaw* op, void* x, void* y) {
  const BuiltinOp* b = builtin_of(op);
  if (!b) return;
  lb2_ctx* c = lb2_default_ctx();
  if (!c) return;
  const size_t bytes = scalar_size(b->prefix) * (size_t)b->n;
  void *dx = lb2_malloc(bytes), *dy = lb2_malloc(bytes);
  if (!dx || !dy) { lb2_free(dx); lb2_free(dy); return; }
  lb2_memcpy_h2d(c, dx, x, bytes);
  lb2_op_apply(c, op, b->prefix, 1, dx, b->n, dy, b->n);
  lb2_memcpy_d2h(c, y, dy, bytes);
  lb2_free(dx);
  lb2_free(dy);
}

static void builtin_cleanup(LinOpCtxRaw* ctx) {
  if (!ctx) return;
  BuiltinOp* b = (BuiltinOp*)ctx->data;
  if (b && b->magic == kOpMagic) {
    if (b->potential) cudaFree(b->potential);
    if (b->rowptr) cudaFree(b->rowptr);
    if (b->col) cudaFree(b->col);
    if (b->val) cudaFree(b->val);
    if (b->diag) cudaFree(b->diag);
    if (b->dense) cudaFree(b->dense);
    if (b->potential_lo) cudaFree(b->potential_lo);
    b->magic = 0;
    free(b);
  }
  free(ctx);
}

static LinOpRaw* wrap_builtin(BuiltinOp* b) {
  LinOpCtxRaw* lc = (LinOpCtxRaw*)calloc(1, sizeof(LinOpCtxRaw));
  LinOpRaw* op = (LinOpRaw*)calloc(1, sizeof(LinOpRaw));
  lc->data = b;
  lc->data_size = sizeof(BuiltinOp);
  op->rows = op->cols = (uint64_t)b->n_global;
  op->matvec = builtin_host_matvec;
  op->cleanup = builtin_cleanup;
  op->ctx = lc;
  b->self = op;
  return op;
}

static BuiltinOp* new_builtin(int kind, char prefix, int64_t n) {
  BuiltinOp* b = (BuiltinOp*)calloc(1, sizeof(BuiltinOp));
  b->magic = kOpMagic;
  b->kind = kind;
  b->prefix = prefix;
  cudaGetDevice(&b->device);
  b->n = b->n_global = n;
  return b;
}

static void* upload(const void* host, size_t bytes) {
  if (!host || !bytes) return nullptr;
  void* d = lb2_malloc(bytes);
  if (!d) return nullptr;
  if (cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(d);
    return nullptr;
  }
  return d;
}

static bool valid_prefix(char p) { return p == 's' || p == 'd' || p == 'c' || p == 'z'; }

void* lb2_op_stencil(char prefix, int64_t gx, int64_t gy, int64_t gz, double cdiag, double coff,
                     const void* potential_host) {
  if (!valid_prefix(prefix) || gx < 1 || gy < 1 || gz < 1) return nullptr;
  if (!lb2_default_ctx()) return nullptr;
  BuiltinOp* b = new_builtin(OP_STENCIL, prefix, gx * gy * gz);
  b->gx = gx; b->gy = gy; b->gz = gz; b->cdiag = cdiag; b->coff = coff;
  double vmax = 0;
  if (potential_host) {
    b->potential = upload(potential_host, real_size(prefix) * (size_t)b->n);
    if (!b->potential) { free(b); return nullptr; }
    for (int64_t i = 0; i < b->n; i++)
      vmax = std::max(vmax, real_size(prefix) == 4 ? (double)((const float*)potential_host)[i] : ((const double*)potential_host)[i]);
  }
  // Gershgorin bound of the spectrum: diagonal + |coff| * number of neighbours
  b->spec_hi = cdiag + vmax + std::fabs(coff) * 2.0 * ((gx > 1) + (gy > 1) + (gz > 1));
  return wrap_builtin(b);
}

// Built-in preconditioner T = p(A): `degree` steps of the Chebyshev iteration for A y = x on the spectrum
// window [lo, hi] — a polynomial approximation of A^-1 that damps every component above `lo` by 1 / T_degree(sigma).
// hi <= 0: Gershgorin bound of the inner built-in operator; lo <= 0: hi / 50.  The result is an ordinary
// LinearOperator_<p>_t* for alg->T (the inner operator must outlive it).
void* lb2_op_chebyshev(char prefix, const void* inner_linop, int degree, double lo, double hi) {
  const LinOpRaw* in = (const LinOpRaw*)inner_linop;
  if (!valid_prefix(prefix) || !in || degree < 0) return nullptr;
  const BuiltinOp* bi = builtin_of(in);
  if (hi <= 0 && bi && bi->kind == OP_CSR && bi->n != bi->n_global) {
    // a row block only knows the Gershgorin bound of its own rows; different windows on different ranks would apply
    // different polynomials
    fprintf(stderr, "lobpcg_b200: lb2_op_chebyshev over a row-partitioned CSR operator needs an explicit upper bound (hi), "
                    "the same on every rank (maximum over ranks of lb2_op_spec_hi)\n");
    return nullptr;
  }
  if (hi <= 0) hi = bi ? bi->spec_hi : 0;
  if (!(hi > 0)) {
    fprintf(stderr, "lobpcg_b200: lb2_op_chebyshev needs an upper spectrum bound (hi) for this operator\n");
    return nullptr;
  }
  if (lo <= 0) lo = hi / 50.0;
  if (!(lo < hi)) return nullptr;
  BuiltinOp* b = new_builtin(OP_CHEB, prefix, bi ? bi->n : (int64_t)in->rows);
  if (bi) { b->n_global = bi->n_global; b->row0 = bi->row0; }
  b->inner = in;
  b->cheb_degree = degree;
  b->cheb_lo = lo;
  b->cheb_hi = hi;
  return wrap_builtin(b);
}

// rank-local z-slab [z0, z0+gz_local) of a gx*gy*gz_global stencil (row-partitioned multi-GPU runs)
void* lb2_op_stencil_slab(char prefix, int64_t gx, int64_t gy, int64_t gz_local, int64_t gz_global, int64_t z0,
                          double cdiag, double coff, const void* potential_local_host) {
  LinOpRaw* op = (LinOpRaw*)lb2_op_stencil(prefix, gx, gy, gz_local, cdiag, coff, potential_local_host);
  if (!op) return nullptr;
  BuiltinOp* b = (BuiltinOp*)op->ctx->data;
  b->n_global = gx * gy * gz_global;
  b->row0 = gx * gy * z0;
  op->rows = op->cols = (uint64_t)b->n_global;
  return op;
}

void* lb2_op_bdg(char prefix, int64_t gx, int64_t gy, int64_t gz, double cdiag, double coff, double shift,
                 double d_re, double d_im) {
  if (!valid_prefix(prefix) || gx < 1 || gy < 1 || gz < 1) return nullptr;
  if (!lb2_default_ctx()) return nullptr;
  BuiltinOp* b = new_builtin(OP_BDG, prefix, 2 * gx * gy * gz);
  b->gx = gx; b->gy = gy; b->gz = gz; b->cdiag = cdiag; b->coff = coff; b->shift = shift;
  b->dre = d_re; b->dim = d_im;
  b->spec_hi = cdiag + shift + std::fabs(coff) * 2.0 * ((gx > 1) + (gy > 1) + (gz > 1)) + std::hypot(d_re, d_im);
  return wrap_builtin(b);
}

// rank-local part of the BdG operator for a row-partitioned run ("partition each half identically so the block
// coupling is rank-local"): local rows = [u(z-slab) ; v(z-slab)], 2 gx gy gz_local rows
void* lb2_op_bdg_slab(char prefix, int64_t gx, int64_t gy, int64_t gz_local, int64_t gz_global, int64_t z0, double cdiag,
                      double coff, double shift, double d_re, double d_im) {
  if (gz_local < 1 || gz_global < gz_local || z0 < 0 || z0 + gz_local > gz_global) return nullptr;
  LinOpRaw* op = (LinOpRaw*)lb2_op_bdg(prefix, gx, gy, gz_local, cdiag, coff, shift, d_re, d_im);
  if (!op) return nullptr;
  BuiltinOp* b = (BuiltinOp*)op->ctx->data;
  b->n_global = 2 * gx * gy * gz_global;
  b->row0 = gx * gy * z0;   // first row of the slab inside EACH field
  op->rows = op->cols = (uint64_t)b->n_global;
  return op;
}

static double csr_gershgorin(char prefix, int64_t n, const int64_t* rp, const void* val) {
  double hi = 0;
  for (int64_t i = 0; i < n; i++) {
    double s = 0;
    for (int64_t p = rp[i]; p < rp[i + 1]; p++) {
      switch (prefix) {
        case 's': s += std::fabs((double)((const float*)val)[p]); break;
        case 'd': s += std::fabs(((const double*)val)[p]); break;
        case 'c': s += std::hypot((double)((const float*)val)[2 * p], (double)((const float*)val)[2 * p + 1]); break;
        default: s += std::hypot(((const double*)val)[2 * p], ((const double*)val)[2 * p + 1]); break;
      }
    }
    hi = std::max(hi, s);
  }
  return hi;
}

// Half-width H of the near-diagonal window for the windowed CSR kernel (spmm.cu: csr_win_kernel): the largest offset
// |col - row| <= 256 that at least 2 % of the rows use, provided the window then holds >= 40 % of the nonzeros and reaches
// beyond the immediate neighbours (which L1 serves anyway).  0 = use the plain kernel.
static int csr_window_halo(int64_t n, const int64_t* rp, const int32_t* col, int64_t col_shift) {
  constexpr int HMAX = 256;
  std::vector<int64_t> cnt(HMAX + 1, 0);
  const int64_t nnz = rp[n];
  for (int64_t i = 0; i < n; i++)
    for (int64_t p = rp[i]; p < rp[i + 1]; p++) {
      const int64_t d = std::llabs((int64_t)col[p] + col_shift - i);
      if (d <= HMAX) cnt[(size_t)d]++;
    }
  int H = 0;
  for (int d = HMAX; d >= 2; d--)
    if (cnt[(size_t)d] >= n / 50 + 1) { H = d; break; }
  if (H < 2) return 0;
  int64_t in = 0;
  for (int d = 0; d <= H; d++) in += cnt[(size_t)d];
  return (in * 10 >= nnz * 4) ? H : 0;
}

void* lb2_op_csr(char prefix, int64_t n, const int64_t* rowptr_host, const int32_t* col_host, const void* val_host) {
  if (!valid_prefix(prefix) || n < 1 || !rowptr_host || !col_host || !val_host) return nullptr;
  if (!lb2_default_ctx()) return nullptr;
  {
    int64_t gx = 0, gy = 0, gz = 0;
    double coff = 0;
    bool st = false, const_diag = false;
    double diag0 = 0;
    void* pot = nullptr;
    auto finish = [&](auto& dg) {
      if (!st) return;
      const_diag = true;
      diag0 = (double)dg[0];
      for (size_t i = 1; i < dg.size() && const_diag; i++) const_diag = (dg[i] == dg[0]);
      pot = upload(dg.data(), sizeof(dg[0]) * dg.size());
    };
    if (prefix == 'd') { std::vector<double> dg; st = detect_stencil<double>(n, rowptr_host, col_host, (const double*)val_host, gx, gy, gz, coff, dg); finish(dg); }
    else if (prefix == 's') { std::vector<float> dg; st = detect_stencil<float>(n, rowptr_host, col_host, (const float*)val_host, gx, gy, gz, coff, dg); finish(dg); }
    else if (prefix == 'z') { std::vector<double> dg; st = detect_stencil<c64>(n, rowptr_host, col_host, (const c64*)val_host, gx, gy, gz, coff, dg); finish(dg); }
    else { std::vector<float> dg; st = detect_stencil<c32>(n, rowptr_host, col_host, (const c32*)val_host, gx, gy, gz, coff, dg); finish(dg); }
    if (st && pot) {
      BuiltinOp* b = new_builtin(OP_STENCIL, prefix, n);
      b->gx = gx; b->gy = gy; b->gz = gz; b->cdiag = 0.0; b->coff = coff;
      b->potential = pot;
      if (const_diag) {   // constant diagonal: no per-point array at all
        cudaFree(pot);
        b->potential = nullptr;
        b->cdiag = diag0;
      }
      b->nnz = rowptr_host[n];   // kept for the CSR traffic accounting of the solver statistics
      b->from_csr = 1;
      b->spec_hi = csr_gershgorin(prefix, n, rowptr_host, val_host);
      return wrap_builtin(b);
    }
  }
  BuiltinOp* b = new_builtin(OP_CSR, prefix, n);
  b->nnz = rowptr_host[n];
  b->rowptr = (int64_t*)upload(rowptr_host, sizeof(int64_t) * (size_t)(n + 1));
  b->col = (int32_t*)upload(col_host, sizeof(int32_t) * (size_t)b->nnz);
  b->val = upload(val_host, scalar_size(prefix) * (size_t)b->nnz);
  if (!b->rowptr || !b->col || !b->val) {
    LinOpCtxRaw* lc = (LinOpCtxRaw*)calloc(1, sizeof(LinOpCtxRaw));
    lc->data = b;
    builtin_cleanup(lc);
    return nullptr;
  }
  b->spec_hi = csr_gershgorin(prefix, n, rowptr_host, val_host);
  b->csr_halo = csr_window_halo(n, rowptr_host, col_host, 0);
  return wrap_builtin(b);
}

void* lb2_op_diag(char prefix, int64_t n, const void* diag_host) {
  if (!valid_prefix(prefix) || n < 1 || !diag_host) return nullptr;
  if (!lb2_default_ctx()) return nullptr;
  BuiltinOp* b = new_builtin(OP_DIAG, prefix, n);
  b->diag = upload(diag_host, real_size(prefix) * (size_t)n);
  if (!b->diag) { free(b); return nullptr; }
  return wrap_builtin(b);
}

// Same preconditioner, evaluated in the LOWER precision inside a double / complex-double solve (float / complex float):
// its cost is HBM traffic and a preconditioner only has to be a fixed approximation of A^-1.  Needs a built-in stencil
// as the inner operator and prefix 'd' or 'z'; anything else falls back to lb2_op_chebyshev.
void* lb2_op_chebyshev_mixed(char prefix, const void* inner_linop, int degree, double lo, double hi) {
  LinOpRaw* op = (LinOpRaw*)lb2_op_chebyshev(prefix, inner_linop, degree, lo, hi);
  if (!op) return nullptr;
  const BuiltinOp* bi = builtin_of((const LinOpRaw*)inner_linop);
  if (!(prefix == 'd' || prefix == 'z') || !bi || bi->kind != OP_STENCIL) return op;
  BuiltinOp* b = (BuiltinOp*)op->ctx->data;
  if (bi->potential) {   // float copy of the (double) potential
    std::vector<double> hd((size_t)bi->n);
    std::vector<float> hf((size_t)bi->n);
    if (cudaMemcpy(hd.data(), bi->potential, sizeof(double) * hd.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return op;
    for (size_t i = 0; i < hd.size(); i++) hf[i] = (float)hd[i];
    b->potential_lo = upload(hf.data(), sizeof(float) * hf.size());
    if (!b->potential_lo) return op;
  }
  b->cheb_mixed = 1;
  return op;
}

// CSR with 32-bit row pointers (the other common host layout)
// Row block [row0, row0 + n_local) of an n_global x n_global CSR matrix for a row-partitioned run ("same row
// blocks, column indices remapped to local+halo").  rowptr_local starts at 0; col_global holds GLOBAL column indices.
// Equal blocks on every rank (n_local the same everywhere: the neighbour's block is addressed at the same arena offset)
// and couplings that reach at most into the two neighbouring blocks.  Columns are stored relative to row0; the kernel reads
// rows of the neighbouring blocks in place from the neighbours' arenas.  spec_hi is the Gershgorin bound of the LOCAL
// rows: pass the maximum over ranks explicitly to lb2_op_chebyshev.
void* lb2_op_csr_slab(char prefix, int64_t n_global, int64_t row0, int64_t n_local, const int64_t* rowptr_local,
                      const int32_t* col_global, const void* val) {
  if (!valid_prefix(prefix) || n_local < 1 || n_global < n_local || row0 < 0 || row0 + n_local > n_global ||
      n_global % n_local != 0 || row0 % n_local != 0 || !rowptr_local || !col_global || !val || rowptr_local[0] != 0)
    return nullptr;
  if (!lb2_default_ctx()) return nullptr;
  const int64_t nnz = rowptr_local[n_local];
  std::vector<int32_t> col((size_t)nnz);
  const int64_t lo = std::max<int64_t>(row0 - n_local, 0), hi = std::min<int64_t>(row0 + 2 * n_local, n_global);
  for (int64_t q = 0; q < nnz; q++) {
    const int64_t c = col_global[q];
    if (c < lo || c >= hi) {
      fprintf(stderr, "lobpcg_b200: lb2_op_csr_slab: column %lld of the row block at %lld reaches beyond the neighbouring "
                      "blocks [%lld, %lld)\n", (long long)c, (long long)row0, (long long)lo, (long long)hi);
      return nullptr;
    }
    col[(size_t)q] = (int32_t)(c - row0);
  }
  BuiltinOp* b = new_builtin(OP_CSR, prefix, n_local);
  b->n_global = n_global;
  b->row0 = row0;
  b->nnz = nnz;
  b->rowptr = (int64_t*)upload(rowptr_local, sizeof(int64_t) * (size_t)(n_local + 1));
  b->col = (int32_t*)upload(col.data(), sizeof(int32_t) * (size_t)nnz);
  b->val = upload(val, scalar_size(prefix) * (size_t)nnz);
  if (!b->rowptr || !b->col || !b->val) {
    LinOpCtxRaw* lc = (LinOpCtxRaw*)calloc(1, sizeof(LinOpCtxRaw));
    lc->data = b;
    builtin_cleanup(lc);
    return nullptr;
  }
  b->spec_hi = csr_gershgorin(prefix, n_local, rowptr_local, val);
  LinOpRaw* op = wrap_builtin(b);
  op->rows = op->cols = (uint64_t)n_global;
  return op;
}

void* lb2_op_csr32(char prefix, int64_t n, const int32_t* rowptr_host, const int32_t* col_host, const void* val_host) {
  if (n < 1 || !rowptr_host) return nullptr;
  std::vector<int64_t> rp((size_t)n + 1);
  for (int64_t i = 0; i <= n; i++) rp[(size_t)i] = rowptr_host[i];
  return lb2_op_csr(prefix, n, rp.data(), col_host, val_host);
}

// Matrix Market reader (coordinate format; real / integer / complex; general / symmetric / hermitian / skew-symmetric):
// builds a CSR operator of the requested scalar type with ascending column indices, duplicate entries summed.
void* lb2_op_csr_from_mtx(char prefix, const char* path) {
  if (!valid_prefix(prefix) || !path) return nullptr;
  FILE* f = fopen(path, "r");
  if (!f) { fprintf(stderr, "lobpcg_b200: cannot open %s\n", path); return nullptr; }
  char line[1024];
  if (!fgets(line, sizeof line, f)) { fclose(f); return nullptr; }
  char obj[64], fmt[64], field[64], sym[64];
  if (sscanf(line, "%%%%MatrixMarket %63s %63s %63s %63s", obj, fmt, field, sym) != 4 || strcmp(obj, "matrix") ||
      strcmp(fmt, "coordinate")) {
    fprintf(stderr, "lobpcg_b200: %s is not a MatrixMarket coordinate matrix\n", path);
    fclose(f);
    return nullptr;
  }
  const bool cplx = !strcmp(field, "complex"), pattern = !strcmp(field, "pattern");
  const bool symm = !strcmp(sym, "symmetric"), herm = !strcmp(sym, "hermitian"), skew = !strcmp(sym, "skew-symmetric");
  do { if (!fgets(line, sizeof line, f)) { fclose(f); return nullptr; } } while (line[0] == '%');
  long long nr = 0, ncl = 0, nz = 0;
  if (sscanf(line, "%lld %lld %lld", &nr, &ncl, &nz) != 3 || nr != ncl || nr < 1 || nr > INT32_MAX) {
    fprintf(stderr, "lobpcg_b200: %s: need a square matrix\n", path);
    fclose(f);
    return nullptr;
  }
  struct Ent { int32_t r, c; double re, im; };
  std::vector<Ent> e;
  e.reserve((size_t)nz * ((symm || herm || skew) ? 2 : 1));
  for (long long q = 0; q < nz; q++) {
    long long i, j;
    double re = 1.0, im = 0.0;
    if (!fgets(line, sizeof line, f)) { fclose(f); return nullptr; }
    const int got = pattern ? sscanf(line, "%lld %lld", &i, &j)
                            : (cplx ? sscanf(line, "%lld %lld %lf %lf", &i, &j, &re, &im) : sscanf(line, "%lld %lld %lf", &i, &j, &re));
    if (got < (pattern ? 2 : (cplx ? 4 : 3)) || i < 1 || j < 1 || i > nr || j > nr) { fclose(f); return nullptr; }
    e.push_back({(int32_t)(i - 1), (int32_t)(j - 1), re, im});
    if (i != j) {
      if (symm) e.push_back({(int32_t)(j - 1), (int32_t)(i - 1), re, im});
      else if (herm) e.push_back({(int32_t)(j - 1), (int32_t)(i - 1), re, -im});
      else if (skew) e.push_back({(int32_t)(j - 1), (int32_t)(i - 1), -re, -im});
    }
  }
  fclose(f);
  std::sort(e.begin(), e.end(), [](const Ent& a, const Ent& b) { return a.r != b.r ? a.r < b.r : a.c < b.c; });
  std::vector<int64_t> rp((size_t)nr + 1, 0);
  std::vector<int32_t> col;
  std::vector<double> vre, vim;
  for (size_t q = 0; q < e.size(); q++) {
    if (q > 0 && e[q].r == e[q - 1].r && e[q].c == e[q - 1].c) { vre.back() += e[q].re; vim.back() += e[q].im; continue; }
    col.push_back(e[q].c); vre.push_back(e[q].re); vim.push_back(e[q].im);
    rp[(size_t)e[q].r + 1]++;
  }
  for (long long i = 0; i < nr; i++) rp[(size_t)i + 1] += rp[(size_t)i];
  const size_t nnz = col.size();
  std::vector<char> val(scalar_size(prefix) * nnz);
  for (size_t q = 0; q < nnz; q++) {
    switch (prefix) {
      case 's': ((float*)val.data())[q] = (float)vre[q]; break;
      case 'd': ((double*)val.data())[q] = vre[q]; break;
      case 'c': ((float*)val.data())[2 * q] = (float)vre[q]; ((float*)val.data())[2 * q + 1] = (float)vim[q]; break;
      default: ((double*)val.data())[2 * q] = vre[q]; ((double*)val.data())[2 * q + 1] = vim[q]; break;
    }
  }
  return lb2_op_csr(prefix, nr, rp.data(), col.data(), val.data());
}

// Eigenpair write-out: a host block (column-major, leading dimension ld) as a Matrix Market dense
// "array" file, real or complex general, 17 significant digits (9 for the single-precision types).  Host-only.
int lb2_write_mtx(const char* path, char prefix, int64_t rows, int64_t cols, const void* host, int64_t ld) {
  if (!path || !valid_prefix(prefix) || rows < 0 || cols < 0 || ld < rows || (!host && rows * cols > 0)) return -1;
  FILE* f = fopen(path, "w");
  if (!f) {
    fprintf(stderr, "lobpcg_b200: cannot open %s for writing\n", path);
    return -1;
  }
  const bool cplx = (prefix == 'c' || prefix == 'z');
  const bool dbl = (prefix == 'd' || prefix == 'z');
  fprintf(f, "%%%%MatrixMarket matrix array %s general\n%lld %lld\n", cplx ? "complex" : "real", (long long)rows,
          (long long)cols);
  const int prec = dbl ? 17 : 9;
  for (int64_t j = 0; j < cols; j++)
    for (int64_t i = 0; i < rows; i++) {
      const size_t q = (size_t)i + (size_t)j * (size_t)ld;
      if (cplx) {
        const double re = dbl ? ((const double*)host)[2 * q] : (double)((const float*)host)[2 * q];
        const double im = dbl ? ((const double*)host)[2 * q + 1] : (double)((const float*)host)[2 * q + 1];
        fprintf(f, "%.*g %.*g\n", prec, re, prec, im);
      } else {
        fprintf(f, "%.*g\n", prec, dbl ? ((const double*)host)[q] : (double)((const float*)host)[q]);
      }
    }
  const bool ok = (ferror(f) == 0);
  return (fclose(f) == 0 && ok) ? 0 : -1;
}

// Dense operator: A is n x n, column-major, leading dimension n, of the operator's scalar type; applied
// to whole blocks by a library GEMM.  The reference's dense examples are host callbacks (tests/test_lobpcg.c:29-42),
// which keep working through the staged path; this constructor is the device-resident version.
void* lb2_op_dense(char prefix, int64_t n, const void* A_host) {
  if (!valid_prefix(prefix) || n < 1 || n > 46340 || !A_host) return nullptr;
  if (!lb2_default_ctx()) return nullptr;
  BuiltinOp* b = new_builtin(OP_DENSE, prefix, n);
  b->dense = upload(A_host, scalar_size(prefix) * (size_t)n * (size_t)n);
  if (!b->dense) { free(b); return nullptr; }
  double hi = 0;   // Gershgorin bound from the absolute row sums
  std::vector<double> rs((size_t)n, 0.0);
  for (int64_t j = 0; j < n; j++)
    for (int64_t i = 0; i < n; i++) {
      const size_t p = (size_t)i + (size_t)j * (size_t)n;
      double v;
      switch (prefix) {
        case 's': v = std::fabs((double)((const float*)A_host)[p]); break;
        case 'd': v = std::fabs(((const double*)A_host)[p]); break;
        case 'c': v = std::hypot((double)((const float*)A_host)[2 * p], (double)((const float*)A_host)[2 * p + 1]); break;
        default: v = std::hypot(((const double*)A_host)[2 * p], ((const double*)A_host)[2 * p + 1]); break;
      }
      rs[(size_t)i] += v;
    }
  for (double v : rs) hi = std::max(hi, v);
  b->spec_hi = hi;
  return wrap_builtin(b);
}

// Caller-supplied block operator on device pointers (the documented extension for foreign operators).  The
// reference's operator interface is one host vector at a time (include/lobpcg/linop.h:15-26); this constructor keeps the
// LinearOperator_<p>_t shape but lets a CUDA application hand its own kernels to the solver: `fn` is called with device
// block vectors (column-major, ncols columns of n rows) and must enqueue its work on `cuda_stream` without synchronising.
void* lb2_op_device(char prefix, int64_t n, lb2_matmat_fn fn, void* user, double spec_hi) {
  if (!valid_prefix(prefix) || n < 1 || !fn) return nullptr;
  if (!lb2_default_ctx()) return nullptr;
  BuiltinOp* b = new_builtin(OP_DEVICE, prefix, n);
  b->dev_fn = (DeviceMatmat)fn;
  b->dev_user = user;
  b->spec_hi = spec_hi > 0 ? spec_hi : 0.0;
  return wrap_builtin(b);
}

// neighbour blocks for a stand-alone lb2_op_apply of a row-block operator (inside a solver they are set per apply from
// the peer arenas): stencil slabs take the boundary PLANES below / above (column stride ld), CSR row blocks take the
// neighbours' whole BLOCKS (column stride ld = their row count)
int lb2_op_set_halo(void* linop, const void* lo, const void* hi, int64_t ld) {
  const BuiltinOp* cb = builtin_of((const LinOpRaw*)linop);
  if (!cb || !(cb->kind == OP_STENCIL || cb->kind == OP_CSR || cb->kind == OP_BDG)) return -1;
  BuiltinOp* b = const_cast<BuiltinOp*>(cb);
  b->halo_lo = lo; b->halo_hi = hi; b->halo_ld = ld;
  return 0;
}

// upper bound of the spectrum recorded at construction (Gershgorin; 0 = unknown, -1 = not a built-in operator)
double lb2_op_spec_hi(const void* linop) {
  const BuiltinOp* b = builtin_of((const LinOpRaw*)linop);
  return b ? b->spec_hi : -1.0;
}

void lb2_op_destroy(void* linop) {
  LinOpRaw* op = (LinOpRaw*)linop;
  if (!op) return;
  if (op->cleanup && op->ctx) op->cleanup(op->ctx);
  free(op);
}

int lb2_op_apply(lb2_ctx* ctx, const void* linop, char prefix, int nc, const void* X, int64_t ldx, void* Y, int64_t ldy) {
  const BuiltinOp* b = builtin_of((const LinOpRaw*)linop);
  if (!b) {
    fprintf(stderr, "lobpcg_b200: lb2_op_apply needs a built-in device operator\n");
    return -1;
  }
  switch (prefix) {
    case 's': return apply_builtin<float>(ctx, b, nc, (const float*)X, ldx, (float*)Y, ldy);
    case 'd': return apply_builtin<double>(ctx, b, nc, (const double*)X, ldx, (double*)Y, ldy);
    case 'c': return apply_builtin<c32>(ctx, b, nc, (const c32*)X, ldx, (c32*)Y, ldy);
    case 'z': return apply_builtin<c64>(ctx, b, nc, (const c64*)X, ldx, (c64*)Y, ldy);
  }
  return -1;
}

// ---------------------------------------------------------------------------------------------------
// solver handle + reference entry points
// ---------------------------------------------------------------------------------------------------
struct lb2_solver {
  SolverBase* impl;
};

lb2_solver* lb2_solver_create(lb2_ctx* ctx, char prefix, void* alg, int indefinite) {
  if (!ctx || !alg || !valid_prefix(prefix)) return nullptr;
  SolverBase* s = make_solver(ctx, prefix, alg, indefinite);
  if (!s) return nullptr;
  lb2_solver* h = new lb2_solver();
  h->impl = s;
  return h;
}
int lb2_solver_init(lb2_solver* s) { return s ? s->impl->init() : -1; }
int lb2_solver_prepare(lb2_solver* s) { return s ? s->impl->prepare() : -1; }
int lb2_solver_arena(lb2_solver* s, void** ptr, size_t* bytes) {
  if (!s) return -1;
  s->impl->arena_info(ptr, bytes);
  return 0;
}
int lb2_solver_set_peers(lb2_solver* s, const void* lo_arena, const void* hi_arena) {
  if (!s) return -1;
  s->impl->set_peers(lo_arena, hi_arena);
  return 0;
}
int lb2_solver_step(lb2_solver* s, int max_steps) { return s ? s->impl->step(max_steps) : -1; }
int lb2_solver_finish(lb2_solver* s) { return s ? s->impl->finish() : -1; }
void lb2_solver_destroy(lb2_solver* s) {
  if (!s) return;
  delete s->impl;
  delete s;
}
int lb2_solver_set_device_x0(lb2_solver* s, uint64_t seed) {
  if (!s) return -1;
  s->impl->use_device_x0 = true;
  s->impl->device_seed = seed;
  return 0;
}
int lb2_solver_set_device_io(lb2_solver* s, const void* x0_dev, void* x_out_dev) {
  if (!s) return -1;
  s->impl->dev_x0 = x0_dev;
  s->impl->dev_xout = x_out_dev;
  return 0;
}
int lb2_solver_num_stats(void) { return PH_COUNT; }
const char* lb2_solver_stat_name(int i) {
  static const char* names[PH_COUNT] = {"spmm_ms", "gram_ms", "tall_nn_ms", "residual_ms", "small_dense_ms", "comm_ms", "other_ms"};
  return (i >= 0 && i < PH_COUNT) ? names[i] : "";
}
double lb2_solver_stat(lb2_solver* s, int i) { return (s && i >= 0 && i < PH_COUNT) ? s->impl->phase_ms[i] : 0.0; }
double lb2_solver_stat_work(lb2_solver* s, int i) { return (s && i >= 0 && i < PH_COUNT) ? s->impl->phase_work[i] : 0.0; }
unsigned long long lb2_solver_stat_calls(lb2_solver* s, int i) { return (s && i >= 0 && i < PH_COUNT) ? s->impl->phase_calls[i] : 0ULL; }
void lb2_solver_reset_stats(lb2_solver* s) {
  if (!s) return;
  for (int i = 0; i < PH_COUNT; i++) { s->impl->phase_ms[i] = 0; s->impl->phase_work[i] = 0; s->impl->phase_calls[i] = 0; }
}
int lb2_solver_results(lb2_solver* s, double* eig, int neig, double* res, int nres) {
  return s ? s->impl->results(eig, neig, res, nres) : -1;
}
int lb2_solver_state(lb2_solver* s, uint64_t* iter, uint64_t* converged, int* use_ortho) {
  if (!s) return -1;
  s->impl->state(iter, converged, use_ortho);
  return 0;
}

// status of the last reference-facing call on this thread (the entry points return void, reference lobpcg.h:63-83):
// 0 = ran (converged or maxIter reached), 1 = parameters rejected with the reference's own message — outputs untouched,
// exactly as the reference leaves them (src/core/lobpcg_impl.inc:66-75), 2 = run-time failure (device error, failed
// factorisation, operator callback error): alg->converged = 0, alg->iter = passes done, eigVals / resNorm = NaN.
static thread_local int g_last_status = 0;
int lb2_last_status(void) { return g_last_status; }

int lb2_solver_set_option(lb2_solver* s, const char* key, int value) { return s ? s->impl->set_option(key, value) : -1; }
double lb2_solver_info(lb2_solver* s, const char* key) { return s ? s->impl->info(key) : -1.0; }

static void run_solver(char prefix, void* alg, int indefinite) {
  g_last_status = 2;
  // several GPUs in this one call (multigpu.cu): LB2_GPUS / lb2_set_num_gpus; -100 = not applicable, single-GPU path below
  const int want = lb2::mg_requested_gpus();
  if (want > 1) {
    const int st = lb2::run_solver_multi(prefix, alg, indefinite, want);
    if (st != -100) {
      g_last_status = st;
      if (st == 2) fprintf(stderr, "lobpcg_b200: multi-GPU solve failed; alg->converged = 0, eigVals = NaN (lb2_last_status() = 2)\n");
      return;
    }
  }
  lb2_note_single_gpu_call();
  lb2_ctx* ctx = lb2_default_ctx();
  if (!ctx) return;
  // LB2_TIMING=1: wall-clock split of the call on stderr (set-up + X0 upload + initial RR | passes | download | tear-down)
  const bool timing = getenv("LB2_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  lb2_solver* s = lb2_solver_create(ctx, prefix, alg, indefinite);
  if (!s) return;
  double t1 = t0, t2 = t0, t3 = t0;
  int rc = lb2_solver_init(s);
  if (rc == 0) {
    t1 = now();
    rc = lb2_solver_step(s, INT_MAX);
    t2 = now();
    rc = (rc >= 0) ? lb2_solver_finish(s) : rc;
    t3 = now();
  }
  if (rc == 0) g_last_status = 0;
  else if (s->impl->param_error) g_last_status = 1;
  else {
    g_last_status = 2;
    s->impl->write_failure_state();
    fprintf(stderr, "lobpcg_b200: solver failed (code %d); alg->converged = 0, eigVals = NaN (lb2_last_status() = 2)\n", rc);
  }
  lb2_solver_destroy(s);
  if (timing)
    fprintf(stderr, "lobpcg_b200 timing: create+alloc+upload+init %.3f s, passes %.3f s, download %.3f s, destroy %.3f s\n",
            t1 - t0, t2 - t1, t3 - t2, now() - t3);
}

#define LB2_ENTRY(P, T)                                                                               \
  void P##_lobpcg(void* alg) { run_solver(#P[0], alg, 0); }                                            \
  void P##_ilobpcg(void* alg) { run_solver(#P[0], alg, 1); }                                           \
  void* lb2_##P##_state_alloc(uint64_t n, uint64_t nev, uint64_t sizeSub, int indefinite) {            \
    State<T>* a = (State<T>*)calloc(1, sizeof(State<T>));                                              \
    if (!a) return nullptr;                                                                            \
    a->size = n; a->nev = nev; a->sizeSub = sizeSub;                                                   \
    a->S = (T*)calloc((size_t)3 * n * sizeSub, sizeof(T)); /* lazily backed; only X is touched */      \
    a->eigVals = (real_t<T>*)calloc(sizeSub, sizeof(real_t<T>));                                       \
    a->resNorm = (real_t<T>*)calloc(sizeSub, sizeof(real_t<T>));                                       \
    if (indefinite) a->signature = (int8_t*)calloc(3 * sizeSub, 1);                                    \
    if (!a->S || !a->eigVals || !a->resNorm) {                                                         \
      fprintf(stderr, "lobpcg_b200: state allocation failed\n");                                       \
      free(a->S); free(a->eigVals); free(a->resNorm); free(a->signature); free(a);                     \
      return nullptr;                                                                                  \
    }                                                                                                  \
    return a;                                                                                          \
  }                                                                                                    \
  void lb2_##P##_state_free(void* alg) {                                                               \
    State<T>* a = (State<T>*)alg;                                                                      \
    if (!a) return;                                                                                    \
    free(a->S); free(a->eigVals); free(a->resNorm); free(a->signature);                                \
    free(a);                                                                                           \
  }                                                                                                    \
  int lb2_##P##_gram(lb2_ctx* ctx, int64_t n, int ma, int mb, const void* A, int64_t lda, const void* B, \
                     int64_t ldb, void* G, int ldg, int upper) {                                       \
    return gram<T>(ctx, n, ma, mb, (const T*)A, lda, (const T*)B, ldb, (T*)G, ldg, upper);             \
  }                                                                                                    \
  int lb2_##P##_gram_cols(lb2_ctx* ctx, int64_t n, int m, int nw, const void* S, int64_t lds, const void* W0, \
                          int64_t ldw0, void* G0, int ldg0, const void* W1, int64_t ldw1, void* G1, int ldg1, \
                          int tri_c0) {                                                                \
    return gram_cols<T>(ctx, n, m, nw, (const T*)S, lds, (const T*)W0, ldw0, (T*)G0, ldg0, (const T*)W1, ldw1, \
                        (T*)G1, ldg1, tri_c0);                                                         \
  }                                                                                                    \
  int lb2_##P##_tall_nn(lb2_ctx* ctx, int64_t n, int kd, int nb, const void* alpha, const void* S,     \
                        int64_t lds, const void* C, int ldc, const void* beta, void* Out, int64_t ldo) { \
    return tall_nn<T>(ctx, n, kd, nb, *(const T*)alpha, (const T*)S, lds, (const T*)C, ldc, *(const T*)beta, \
                      (T*)Out, ldo);                                                                   \
  }                                                                                                    \
  int lb2_##P##_residual(lb2_ctx* ctx, int64_t n, int nc, const void* AX, int64_t ldax, const void* BX, \
                         int64_t ldbx, const void* lambda, void* W, int64_t ldw, void* sumsq) {        \
    return residual<T>(ctx, n, nc, (const T*)AX, ldax, (const T*)BX, ldbx, (const real_t<T>*)lambda,   \
                       (T*)W, ldw, (real_t<T>*)sumsq);                                                 \
  }                                                                                                    \
  int lb2_##P##_col_sumsq(lb2_ctx* ctx, int64_t n, int nc, const void* X, int64_t ldx, void* sumsq) {  \
    return col_sumsq<T>(ctx, n, nc, (const T*)X, ldx, (real_t<T>*)sumsq);                              \
  }                                                                                                    \
  int lb2_##P##_fill_uniform(lb2_ctx* ctx, int64_t n, int nc, void* X, int64_t ldx, uint64_t seed,     \
                             int64_t n_global, int64_t row0) {                                         \
    return fill_uniform<T>(ctx, n, nc, (T*)X, ldx, seed, n_global, row0);                              \
  }                                                                                                    \
  int lb2_##P##_spmm_stencil(lb2_ctx* ctx, int64_t gx, int64_t gy, int64_t gz, double cdiag, double coff, \
                             const void* potential, int nc, const void* X, int64_t ldx, void* Y,       \
                             int64_t ldy) {                                                            \
    StencilDesc d;                                                                                     \
    memset(&d, 0, sizeof(d));                                                                          \
    d.gx = (int)gx; d.gy = (int)gy; d.gz = (int)gz; d.cdiag = cdiag; d.coff = coff;                    \
    d.potential = potential;                                                                           \
    return spmm_stencil<T>(ctx, d, nc, (const T*)X, ldx, (T*)Y, ldy);                                  \
  }                                                                                                    \
  int lb2_##P##_spmm_stencil_halo(lb2_ctx* ctx, int64_t gx, int64_t gy, int64_t gz, double cdiag, double coff, \
                                  const void* potential, const void* halo_lo, const void* halo_hi,     \
                                  int64_t halo_ld, int nc, const void* X, int64_t ldx, void* Y, int64_t ldy) { \
    StencilDesc d;                                                                                     \
    memset(&d, 0, sizeof(d));                                                                          \
    d.gx = (int)gx; d.gy = (int)gy; d.gz = (int)gz; d.cdiag = cdiag; d.coff = coff;                    \
    d.potential = potential; d.halo_lo = halo_lo; d.halo_hi = halo_hi; d.halo_ld = halo_ld;            \
    return spmm_stencil<T>(ctx, d, nc, (const T*)X, ldx, (T*)Y, ldy);                                  \
  }                                                                                                    \
  int lb2_##P##_spmm_csr(lb2_ctx* ctx, int64_t n, const int64_t* rowptr, const int32_t* col,           \
                         const void* val, int nc, const void* X, int64_t ldx, void* Y, int64_t ldy) {  \
    return spmm_csr<T>(ctx, n, rowptr, col, (const T*)val, nc, (const T*)X, ldx, (T*)Y, ldy);          \
  }                                                                                                    \
  int lb2_##P##_spmm_diag(lb2_ctx* ctx, int64_t n, const void* diag, int nc, const void* X, int64_t ldx, \
                          void* Y, int64_t ldy) {                                                      \
    return spmm_diag<T>(ctx, n, (const real_t<T>*)diag, nc, (const T*)X, ldx, (T*)Y, ldy);             \
  }

LB2_ENTRY(s, float)
LB2_ENTRY(d, double)
LB2_ENTRY(c, c32)
LB2_ENTRY(z, c64)

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// multi-GPU hooks (comm.cu provides the NCCL-backed implementation when a communicator is attached)
// ---------------------------------------------------------------------------------------------------
namespace lb2 {
int comm_allreduce_impl(void* comm, void* buf, size_t count, bool is_double, cudaStream_t st);
int comm_rank_impl(void* comm);
int comm_size_impl(void* comm);
int allreduce_sum(lb2_ctx* ctx, void* buf, size_t count, bool is_double) {
  if (!ctx->comm) return 0;
  ctx->launches++;
  return comm_allreduce_impl(ctx->comm, buf, count, is_double, ctx->stream);
}
int comm_rank(lb2_ctx* ctx) { return ctx->comm ? comm_rank_impl(ctx->comm) : 0; }
int comm_size(lb2_ctx* ctx) { return ctx->comm ? comm_size_impl(ctx->comm) : 1; }
}  // namespace lb2

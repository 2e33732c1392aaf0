// lobpcg_b200/csrc/context.h — per-device execution context behind the opaque `lb2_ctx` of the C ABI
// (include/lobpcg_b200.h).  One stream, one growable scratch buffer for split-reduction partials,
// cuBLAS/cuSOLVER handles for the small (<= 3k x 3k) dense factorizations only.
#pragma once
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <stdint.h>
#include <stddef.h>

struct lb2_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // scratch for deterministic split-n partial sums (Gram) and per-CTA norm partials
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // int8 slices, exponents and partial sums of the Ozaki-split f64 Gram (gram_i8.cu); sized exactly, kept between calls
  void* oz_buf = nullptr;
  size_t oz_bytes = 0;
  // the block whose slices sit at the start of oz_buf (left by the last column-block Gram; reused by the projections of the pass)
  const void* oz_tag_ptr = nullptr;
  int64_t oz_tag_n = 0, oz_tag_ld = 0;
  int oz_tag_m = 0;
  size_t oz_tag_e_off = 0;
  // device times of the int8 Gram phases (split / MMA kernel / reduce), accumulated over calls (lb2_ctx_oz_stats)
  cudaEvent_t oz_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool oz_ev_pending = false;
  double oz_ms[3] = {0.0, 0.0, 0.0};
  long long oz_calls = 0;
  // exponent hints of the three operand roles of the solver's column-block Gram (device: 3 x 4096 exponents + 3 redo flags)
  void* oz_hint = nullptr;
  bool oz_hint_valid[3] = {false, false, false};
  int64_t oz_hint_n[3] = {0, 0, 0};
  int oz_hint_m[3] = {0, 0, 0};
  long long oz_hint_redos = 0;
  int oz_hints = 1;          // option: 0 = every split takes its column maxima in a separate pass
  bool oz_reuse = false;     // set by the solver between the Gram of a pass and its projections: only then are the cached slices trusted
  int active_solvers = 0;    // solvers holding an arena on this context (their slice buffer must not be taken away under memory pressure)
  // small-dense library handles (created lazily)
  cublasHandle_t cublas = nullptr;
  cusolverDnHandle_t cusolver = nullptr;
  void* solver_ws = nullptr;   // cuSOLVER device workspace
  size_t solver_ws_bytes = 0;
  void* solver_hws = nullptr;  // cuSOLVER host workspace (X* API)
  size_t solver_hws_bytes = 0;
  int* dev_info = nullptr;
  // tuning knobs (lb2_ctx_set_option)
  int gram_tile = 0;     // 0 = heuristic, 64 or 128
  int nn_tile = 0;       // 0 = heuristic
  int nn_bk = 0;         // K chunk of the 128x128 tall_nn tile: 0/16 or 32 (tuning)
  int nn_warps = 0;      // f64 tall_nn experiment: 16 = 16-warp CTAs with 32x32 warp tiles (one tile per CTA)
  int gram_merge = 1;    // column-block Gram: the two products' ragged remainder column tiles computed as one merged tile (0 = separately)
  int nn_stagger = 1;    // f64 tall_nn / work-list Gram: the two warps of a scheduler issue their chunk copies half a chunk apart (0 = together)
  int nn_persist = -1;   // f64 tall_nn: persistent 128x128 tiles with the copy ring running across tiles (-1 = on, 0 = off)
  int gram_wl = -1;      // f64 Gram through the work-list kernel (gram_wl.cu): -1 = auto, 0 = never, 1 = always
  int gram_bk = 0;       // K chunk of the work-list Gram: 0/16 or 32 (tuning)
  int gram_strip_max = 0; // work-list Gram: ragged last tile columns narrower than this go to the lock-step kernel (0 = always, -1 = never)
  int gram_strip_fma = 0; // work-list Gram: -1 = never use the streaming FMA kernel for narrow strips (testing)
  int gram_load_pct = 0; // work-list Gram: staging-traffic cost of a tile with 256 columns in % of its DMMA time (0 = default)
  int gram_phase = -1;      // work-list Gram: phase-aligned cyclic walk of the pieces (gram_wl.cu); 0 = off
  void* hostcopy = nullptr;        // lb2::HostCopyState* (pinned ring for pageable host <-> device block transfers)
  void* gram_wl_cache = nullptr;   // lb2::WlCache* (schedules per Gram shape)
  void* gram_wl_cols_cache = nullptr;   // lb2::WlColsCache* (schedules of the column-block products, gram_wl_cols_f64)
  int gram_i8 = -1;      // f64 Gram / projection on tcgen05 kind::i8 through an Ozaki split (gram_i8.cu): -1 = auto (on for n >= 2^18 rows and products of at least 200 x 200), 1 = on (n >= 4096), 0 = DMMA kernels, 2 = on + phase times on stderr
  int gram_i8_env = -1;  // LB2_GRAM_I8 as seen by the last solver set-up (-1 = unset); overrides gram_i8 for the drop-in entry points
  int nn_i8 = 1;         // with gram_i8 on: projections Out = S C (alpha 1, beta 0) on the int8 tensor path too (0 = DMMA kernel)
  int oz_ring = 0;       // cluster Gram kernel: A-ring slots (3 / 4 / 5 / 6 of 12; 0 = 4)
  int oz_nn_ring = 0;    // int8 projection kernel: A-ring slots (5 / 6 / 7 / 8; 0 = 7)
  int oz_prefetch = 0;   // int8 kernels: L2 prefetch distance of the slice tiles in chunks (0 = off: measured slower at every distance, r02)
  int oz_lockstep = 1;   // gram_i8 one-tile-per-CTA kernel: 1 = lock-step cohorts (every (tile, level group) has its own CTAs), 0 = equal-cost cut
  int oz_cluster = 1;    // gram_i8 column-block products: 1 = 4-CTA clusters with multicast slice tiles (r02, C5 shape: 71 ms against 78 ms of the one-tile-per-CTA kernel), 0 = one tile per CTA
  int oz_clusters = -1;  // resident clusters of the cluster kernel (queried once; option: force a count)
  int oz_load_pct = 0;   // gram_i8 schedule: cost of one 16 KB slice-tile load relative to one full-width slice product, in % (0 = 100)
  int gram_tc5 = -1;     // float Gram through tcgen05 / TMEM (gram_tc5.cu): -1 = auto (on), 0 = off, 1 = on
  int gram_tma = -1;     // float Gram: TMA-fed tcgen05 kernel (gram_tc5.cu: gram_tc5_tma_kernel): -1 / 1 = on, 0 = cp.async-fed kernel
  int force_simt = 0;    // 1 = use the generic SIMT kernels even for f64 (testing)
  int spmm_cols = 0;     // CSR SpMM columns per thread (0 = heuristic)
  int csr_order = 512;   // plain CSR kernel: row blocks per chunk of the chunked 1-D launch order (spmm.cu: csr_kernel), 0 = column-group-major 2-D grid
  int csr_pipe = 1;      // plain CSR kernel: 1 = next (col, val) pair requested ahead of the current gathers, 2 = two couplings per step, 0 = plain loop
  int csr_staged = 0;    // 1: CSR kernel with the (col, val) stream staged in shared memory (r02: 1.71 ms against 1.37 ms of the plain kernel at 128^3 x 128)
  int csr_lpr = 0;       // lanes per row of the staged CSR kernel: 0 = from the mean row length, else 1 / 4 / 16
  int csr_window = -1;   // windowed CSR kernel for banded matrices: 1 = on; -1 / 0 = off (r02: 2.09 ms against 1.38 ms of the plain kernel at 128^3 x 128)
  // one released solver arena kept for the next solve on this context (arena_alloc / arena_release in capi.cu): a 100 GB
  // cudaMalloc + cudaFree pair costs 0.1-0.5 s per reference-facing call; LB2_ARENA_CACHE=0 disables, lb2_ctx_trim frees
  void* arena_cache = nullptr;
  size_t arena_cache_bytes = 0;
  void* pinned_cache = nullptr;      // small pinned host buffer of the last solver (same idea)
  size_t pinned_cache_bytes = 0;
  // launch counter (bench.py "gpu_launches")
  unsigned long long launches = 0;
  // multi-GPU (row-partitioned) state; comm == nullptr => single GPU
  void* comm = nullptr;  // lb2::Comm*
};

namespace lb2 {
// returns a scratch pointer of at least `bytes` (grows, stream-ordered-safe because every user is
// enqueued on ctx->stream and growth synchronizes first).
void* ctx_scratch(lb2_ctx* ctx, size_t bytes);
void gram_wl_cache_free(lb2_ctx* ctx);   // gram_wl.cu
// solver arena with reuse across solves (capi.cu); arena_alloc frees every context's cached arena before giving up
void* arena_alloc(lb2_ctx* ctx, size_t bytes);
void arena_release(lb2_ctx* ctx, void* p, size_t bytes);
void* pinned_take(lb2_ctx* ctx, size_t bytes);
void pinned_give(lb2_ctx* ctx, void* p, size_t bytes);
// hostcopy.cu: pipelined copy between PAGEABLE host memory and the device (host-synchronous, ordered after ctx->stream)
int host_copy(lb2_ctx* ctx, void* dst, const void* src, size_t bytes, bool to_device);
void hostcopy_free(lb2_ctx* ctx);
// strided pageable host <-> device transfer (cols segments of rows_bytes; column strides in bytes)
int host_copy_2d(lb2_ctx* ctx, void* dev, size_t ld_dev_bytes, void* host, size_t ld_host_bytes, size_t rows_bytes, int cols,
                 bool to_device);
void hostcopy_set_threads(lb2_ctx* ctx, int nthreads);
}  // namespace lb2

// effective setting of the int8 Gram path (environment for the reference entry points, else the context option)
inline int lb2_gram_i8_mode(const lb2_ctx* c) { return c->gram_i8_env >= 0 ? c->gram_i8_env : c->gram_i8; }
// does a product over n rows take the int8 path?  auto: only where the tall kernels dominate (the path has more launches and a
// host-side schedule per call)
// (ma x mb = shape of the small side of the product: the path pays where the product is bound by the tensor pipe; narrow blocks are
// bound by HBM, and there the split's extra pass over the operands costs more than the faster MMAs return)
inline bool lb2_gram_i8_on(const lb2_ctx* c, int64_t n, int64_t ma = 1 << 20, int64_t mb = 1) {
  const int mode = lb2_gram_i8_mode(c);
  return mode > 0 ? n >= 4096 : (mode < 0 && n >= ((int64_t)1 << 18) && ma * mb >= 40000);
}
namespace lb2 {
// device memory for the int8 slices; under pressure frees the cached arenas and the idle slice buffers of other contexts (capi.cu)
void* oz_malloc(lb2_ctx* ctx, size_t bytes);
}

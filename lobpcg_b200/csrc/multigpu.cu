// lobpcg_b200/csrc/multigpu.cu — the reference-facing entry points on SEVERAL GPUs of one node, in ONE process.
//
// <p>_lobpcg(alg) / <p>_ilobpcg(alg) (reference lobpcg.h:63-83) are single calls on host buffers; a drop-in caller cannot
// be asked to launch one process per GPU.  With LB2_GPUS=N (or lb2_set_num_gpus) the call itself spreads over N devices:
//   * one worker thread per device, each with that device's default context and an ordinary row-partitioned Solver
//     (solver.cu) — the same code path the torchrun launcher drives with one process per GPU;
//   * the caller's built-in operators are re-created as row slabs on every device (stencil / BdG z-slabs, CSR row
//     blocks, diagonal slices, the polynomial preconditioner over the local inner operator);
//   * neighbour halos are read in place through CUDA peer access (same process: no IPC handles), partial Gram sums go
//     through an NCCL communicator created once per process (ncclCommInitRank from the worker threads);
//   * every device uploads its own rows of X0 and downloads its own rows of the eigenvectors, so the host <-> device
//     traffic of the call is split over N PCIe links (hostcopy.cu: host_copy_2d).
// Falls back to the single-GPU path (return -100) when an operator cannot be partitioned (host callbacks, dense and
// caller-supplied device operators) or the grid does not split into equal slabs.
#include <algorithm>
#include <chrono>
#include <climits>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"
#include "context.h"
#include "solver.h"
#include "../../include/lobpcg_b200.h"

namespace lb2 {

namespace {

struct Barrier {
  std::mutex mu;
  std::condition_variable cv;
  int count = 0, gen = 0, n = 0;
  explicit Barrier(int n_) : n(n_) {}
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    const int g = gen;
    if (++count == n) { count = 0; gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};

size_t scalar_bytes(char p) { return p == 's' ? 4 : (p == 'd' || p == 'c') ? 8 : 16; }
size_t real_bytes(char p) { return (p == 's' || p == 'c') ? 4 : 8; }

// host copies of the device arrays of the caller's operators (made once, on the operators' own device)
struct HostOp {
  const BuiltinOp* b = nullptr;
  std::vector<char> potential, diag, val, potential_inner;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> col;
  HostOp* inner = nullptr;
};

int download(std::vector<char>& dst, const void* dev, size_t bytes) {
  dst.resize(bytes);
  if (!bytes) return 0;
  LB2_CUDA_OK(cudaMemcpy(dst.data(), dev, bytes, cudaMemcpyDeviceToHost));
  return 0;
}

bool partitionable(const BuiltinOp* b) {
  if (!b || b->n != b->n_global) return false;
  switch (b->kind) {
    case OP_STENCIL: case OP_BDG: case OP_CSR: case OP_DIAG: return true;
    case OP_CHEB: return partitionable(builtin_of(b->inner));
    default: return false;
  }
}

int snapshot(const BuiltinOp* b, HostOp& h, std::vector<HostOp*>& pool) {
  h.b = b;
  LB2_CUDA_OK(cudaSetDevice(b->device));
  const size_t rs = real_bytes(b->prefix), ss = scalar_bytes(b->prefix);
  if ((b->kind == OP_STENCIL || b->kind == OP_BDG) && b->potential) {
    if (download(h.potential, b->potential, rs * (size_t)(b->kind == OP_BDG ? b->n / 2 : b->n))) return -1;
  }
  if (b->kind == OP_DIAG && download(h.diag, b->diag, rs * (size_t)b->n)) return -1;
  if (b->kind == OP_CSR) {
    h.rowptr.resize((size_t)b->n + 1);
    LB2_CUDA_OK(cudaMemcpy(h.rowptr.data(), b->rowptr, sizeof(int64_t) * h.rowptr.size(), cudaMemcpyDeviceToHost));
    h.col.resize((size_t)b->nnz);
    LB2_CUDA_OK(cudaMemcpy(h.col.data(), b->col, sizeof(int32_t) * h.col.size(), cudaMemcpyDeviceToHost));
    if (download(h.val, b->val, ss * (size_t)b->nnz)) return -1;
  }
  if (b->kind == OP_CHEB) {
    h.inner = new HostOp();
    pool.push_back(h.inner);
    if (snapshot(builtin_of(b->inner), *h.inner, pool)) return -1;
  }
  return 0;
}

// row map of the partition, taken from A: nseg runs per rank (2 for the BdG operator: u slab and v slab)
struct RowMap {
  int R = 1, nseg = 1;
  int64_t ng = 0, seg_len = 0;     // rows of one run on one rank
  int64_t plane = 0, gz = 0;       // stencil-type A: z-slabs
  int64_t run_global(int rank, int s) const { return (int64_t)s * (ng / nseg) + (int64_t)rank * seg_len; }
  int64_t n_local() const { return seg_len * nseg; }
};

// this rank's slab of an operator, created on the CURRENT device; nullptr on failure
LinOpRaw* make_slab(const HostOp& h, const RowMap& rm, int rank, std::vector<LinOpRaw*>& owned) {
  const BuiltinOp* b = h.b;
  const char p = b->prefix;
  const size_t rs = real_bytes(p), ss = scalar_bytes(p);
  LinOpRaw* op = nullptr;
  switch (b->kind) {
    case OP_STENCIL: {
      const int64_t gzl = b->gz / rm.R, z0 = gzl * rank, pl = b->gx * b->gy;
      const void* pot = h.potential.empty() ? nullptr : (const void*)(h.potential.data() + rs * (size_t)(pl * z0));
      op = (LinOpRaw*)lb2_op_stencil_slab(p, b->gx, b->gy, gzl, b->gz, z0, b->cdiag, b->coff, pot);
      if (op) {
        BuiltinOp* nb = (BuiltinOp*)op->ctx->data;
        nb->from_csr = b->from_csr;
        nb->nnz = b->nnz / rm.R;
      }
      break;
    }
    case OP_BDG: {
      const int64_t gzl = b->gz / rm.R, z0 = gzl * rank;
      op = (LinOpRaw*)lb2_op_bdg_slab(p, b->gx, b->gy, gzl, b->gz, z0, b->cdiag, b->coff, b->shift, b->dre, b->dim);
      break;
    }
    case OP_CSR: {
      const int64_t nl = rm.n_local(), r0 = rm.run_global(rank, 0);
      if (rm.nseg != 1) return nullptr;
      const int64_t p0 = h.rowptr[(size_t)r0], p1 = h.rowptr[(size_t)(r0 + nl)];
      std::vector<int64_t> rp((size_t)nl + 1);
      for (int64_t i = 0; i <= nl; i++) rp[(size_t)i] = h.rowptr[(size_t)(r0 + i)] - p0;
      (void)p1;
      op = (LinOpRaw*)lb2_op_csr_slab(p, b->n, r0, nl, rp.data(), h.col.data() + p0, h.val.data() + ss * (size_t)p0);
      break;
    }
    case OP_DIAG: {
      std::vector<char> loc(rs * (size_t)rm.n_local());
      for (int s = 0; s < rm.nseg; s++)
        memcpy(loc.data() + rs * (size_t)(s * rm.seg_len), h.diag.data() + rs * (size_t)rm.run_global(rank, s),
               rs * (size_t)rm.seg_len);
      op = (LinOpRaw*)lb2_op_diag(p, rm.n_local(), loc.data());
      if (op) {
        BuiltinOp* nb = (BuiltinOp*)op->ctx->data;
        nb->n_global = rm.ng;
        nb->row0 = rm.run_global(rank, 0);
        op->rows = op->cols = (uint64_t)rm.ng;
      }
      break;
    }
    case OP_CHEB: {
      LinOpRaw* in = make_slab(*h.inner, rm, rank, owned);
      if (!in) return nullptr;
      op = (LinOpRaw*)(b->cheb_mixed ? lb2_op_chebyshev_mixed(p, in, b->cheb_degree, b->cheb_lo, b->cheb_hi)
                                     : lb2_op_chebyshev(p, in, b->cheb_degree, b->cheb_lo, b->cheb_hi));
      break;
    }
    default: return nullptr;
  }
  if (!op) return nullptr;
  BuiltinOp* nb = (BuiltinOp*)op->ctx->data;
  if (b->spec_hi > 0) nb->spec_hi = b->spec_hi;   // the bound of the WHOLE operator (a slab only knows its own rows)
  owned.push_back(op);
  return op;
}

std::mutex g_mg_mu;
int g_num_gpus = 0;       // lb2_set_num_gpus (0 = environment)
int g_last_gpus = 1;
unsigned char g_nccl_id[128];

template <typename T>
int run_multi_typed(char prefix, State<T>* alg, int indefinite, int want) {
  const LinOpRaw *A = alg->A, *B = alg->B, *Tp = alg->T_;
  const BuiltinOp* ba = builtin_of(A);
  if (!A || !partitionable(ba) || ba->kind == OP_DIAG || ba->kind == OP_CHEB) return -100;
  if (B && !partitionable(builtin_of(B))) return -100;
  if (Tp && !partitionable(builtin_of(Tp))) return -100;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 2) return -100;
  want = std::min(want, ndev);
  RowMap rm;
  rm.ng = (int64_t)alg->size;
  if (ba->n_global != rm.ng) return -100;
  rm.nseg = (ba->kind == OP_BDG) ? 2 : 1;
  int R = want;
  if (ba->kind == OP_STENCIL || ba->kind == OP_BDG) {
    while (R > 1 && ba->gz % R) R--;
    rm.plane = ba->gx * ba->gy;
    rm.gz = ba->gz;
  } else {
    while (R > 1 && rm.ng % R) R--;
  }
  if (R < 2) return -100;
  // every device must reach its neighbours' memory
  for (int r = 0; r + 1 < R; r++) {
    int ok1 = 0, ok2 = 0;
    cudaDeviceCanAccessPeer(&ok1, r, r + 1);
    cudaDeviceCanAccessPeer(&ok2, r + 1, r);
    if (!ok1 || !ok2) return -100;
  }
  rm.R = R;
  rm.seg_len = rm.ng / rm.nseg / R;
  if (3 * alg->sizeSub > alg->size || alg->nev > alg->sizeSub) return -100;   // let the single path print the reference's messages

  const bool timing = getenv("LB2_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  int dev0 = 0;
  cudaGetDevice(&dev0);
  std::vector<HostOp*> pool;
  HostOp hA, hB, hT;
  int rc0 = snapshot(ba, hA, pool);
  if (!rc0 && B) rc0 = snapshot(builtin_of(B), hB, pool);
  if (!rc0 && Tp) {
    const BuiltinOp* bt = builtin_of(Tp);
    // a preconditioner over A itself shares A's snapshot
    rc0 = snapshot(bt, hT, pool);
  }
  cudaSetDevice(dev0);
  if (rc0) { for (auto* q : pool) delete q; return 2; }

  // NCCL communicator over the R devices, created once and kept on the devices' default contexts
  bool need_comm = false;
  {
    std::lock_guard<std::mutex> lk(g_mg_mu);
    for (int r = 0; r < R; r++) {
      cudaSetDevice(r);
      lb2_ctx* c = lb2_default_ctx();
      if (!c) { cudaSetDevice(dev0); for (auto* q : pool) delete q; return 2; }
      if (!c->comm || comm_size(c) != R || comm_rank(c) != r) need_comm = true;
    }
    cudaSetDevice(dev0);
    if (need_comm && lb2_comm_unique_id(g_nccl_id, getenv("LB2_NCCL_LIB"))) { for (auto* q : pool) delete q; return 2; }
  }

  std::vector<int> rcs(R, 0);
  std::vector<void*> arenas(R, nullptr);
  std::vector<State<T>> locals(R);
  std::vector<std::vector<real_t<T>>> eigs(R), ress(R);
  std::vector<std::vector<int8_t>> sigs(R);
  std::vector<char> perr(R, 0);
  Barrier bar(R);
  double t_init = 0, t_step = 0, t_fin = 0;
  const unsigned hw = std::thread::hardware_concurrency();

  auto worker = [&](int r) {
    int rc = 0;
    std::vector<LinOpRaw*> owned;
    SolverBase* s = nullptr;
    bool alive = true;
    auto fail = [&](int code) { rc = code; alive = false; };
    if (cudaSetDevice(r) != cudaSuccess) fail(2);
    lb2_ctx* ctx = alive ? lb2_default_ctx() : nullptr;
    if (alive && !ctx) fail(2);
    if (alive) {
      for (int nb : {r - 1, r + 1})
        if (nb >= 0 && nb < R) {
          cudaError_t e = cudaDeviceEnablePeerAccess(nb, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) fail(2);
          cudaGetLastError();
        }
      hostcopy_set_threads(ctx, (int)std::max(1u, std::min(8u, (hw ? hw : 8u) / (2u * (unsigned)R) + 1u)));
    }
    if (need_comm) {   // collective: every worker must take part, also one that already failed locally is fatal for all
      if (alive) {
        lb2_ctx_detach_comm(ctx);
        if (lb2_ctx_attach_comm(ctx, r, R, g_nccl_id, getenv("LB2_NCCL_LIB"))) fail(2);
      }
    }
    LinOpRaw *lA = nullptr, *lB = nullptr, *lT = nullptr;
    if (alive) {
      lA = make_slab(hA, rm, r, owned);
      if (B) lB = make_slab(hB, rm, r, owned);
      if (Tp) lT = make_slab(hT, rm, r, owned);
      if (!lA || (B && !lB) || (Tp && !lT)) fail(2);
    }
    if (alive) {
      State<T>& loc = locals[r];
      loc = *alg;
      loc.A = lA; loc.B = lB; loc.T_ = lT;
      if (r != 0) loc.verbosity = 0;
      eigs[r].assign((size_t)alg->sizeSub, real_t<T>(0));
      ress[r].assign((size_t)alg->sizeSub, real_t<T>(0));
      sigs[r].assign(3 * (size_t)alg->sizeSub, 0);
      if (r != 0) {   // all ranks hold identical scalars; only rank 0 writes the caller's arrays
        loc.eigVals = eigs[r].data();
        loc.resNorm = ress[r].data();
        loc.signature = alg->signature ? sigs[r].data() : nullptr;
      }
      s = make_solver(ctx, prefix, &loc, indefinite);
      if (!s) fail(2);
    }
    if (alive) {
      const int prc = s->prepare();
      if (prc) { perr[r] = s->param_error ? 1 : 0; fail(prc); }
    }
    size_t bytes = 0;
    if (alive) s->arena_info(&arenas[r], &bytes);
    rcs[r] = rc;
    bar.wait();                                   // every arena is allocated (or a rank has failed)
    bool all_ok = true;
    for (int q = 0; q < R; q++) all_ok = all_ok && rcs[q] == 0;
    if (all_ok) {
      s->set_peers(r > 0 ? arenas[r - 1] : nullptr, r + 1 < R ? arenas[r + 1] : nullptr);
      const double ta = now();
      rc = s->init();
      const double tb = now();
      if (rc == 0) {
        const int st = s->step(INT_MAX);
        rc = st < 0 ? st : 0;
      }
      const double tc = now();
      if (rc == 0) rc = s->finish();
      const double td = now();
      if (r == 0) { t_init = tb - ta; t_step = tc - tb; t_fin = td - tc; }
      if (rc != 0 && r == 0) s->write_failure_state();
      rcs[r] = rc;
    }
    bar.wait();                                   // nobody frees an arena a neighbour may still read
    if (s) delete s;
    for (auto it = owned.rbegin(); it != owned.rend(); ++it) lb2_op_destroy(*it);
  };

  std::vector<std::thread> th;
  for (int r = 1; r < R; r++) th.emplace_back(worker, r);
  worker(0);
  for (auto& x : th) x.join();
  cudaSetDevice(dev0);
  for (auto* q : pool) delete q;
  int status = 0;
  for (int r = 0; r < R; r++) if (rcs[r] != 0) status = perr[r] ? 1 : 2;
  if (status == 0) {
    alg->iter = locals[0].iter;
    alg->converged = locals[0].converged;
  }
  g_last_gpus = R;
  if (timing)
    fprintf(stderr, "lobpcg_b200 timing (%d GPUs in one process): set-up %.3f s, upload+init %.3f s, passes %.3f s, download %.3f s\n",
            R, now() - t0 - t_init - t_step - t_fin, t_init, t_step, t_fin);
  return status;
}

}  // namespace

int mg_requested_gpus() {
  if (g_num_gpus > 0) return g_num_gpus;
  const char* e = getenv("LB2_GPUS");
  if (!e || !*e) return 1;
  if (!strcmp(e, "all")) {
    int nd = 1;
    if (cudaGetDeviceCount(&nd) != cudaSuccess) nd = 1;
    return nd;
  }
  const int v = atoi(e);
  return v > 0 ? v : 1;
}

// returns the status of the call (0 / 1 / 2 as lb2_last_status), or -100 when the call has to take the single-GPU path
int run_solver_multi(char prefix, void* alg, int indefinite, int want) {
  if (want < 2 || !alg) return -100;
  switch (prefix) {
    case 's': return run_multi_typed<float>(prefix, (State<float>*)alg, indefinite, want);
    case 'd': return run_multi_typed<double>(prefix, (State<double>*)alg, indefinite, want);
    case 'c': return run_multi_typed<c32>(prefix, (State<c32>*)alg, indefinite, want);
    case 'z': return run_multi_typed<c64>(prefix, (State<c64>*)alg, indefinite, want);
  }
  return -100;
}

}  // namespace lb2

extern "C" {
// number of GPUs the reference-facing entry points may use (one process, one worker thread per device); 0 = take it from
// the environment (LB2_GPUS = N | all; default 1).  Returns the previous setting.
int lb2_set_num_gpus(int n) {
  const int old = lb2::g_num_gpus;
  lb2::g_num_gpus = n > 0 ? n : 0;
  return old;
}
// how many GPUs the last <p>_lobpcg / <p>_ilobpcg call really ran on
int lb2_last_num_gpus(void) { return lb2::g_last_gpus; }
void lb2_note_single_gpu_call(void) { lb2::g_last_gpus = 1; }
}

// lobpcg_b200/csrc/comm.cu — row-partitioned multi-GPU plumbing (one process per GPU).
//
// The reference has no distributed path at all; this is new.  Rows of every block vector
// are split into contiguous z-slabs, one per rank.  Two exchange points exist on the hot path:
//   (1) k x k partial Gram sums and column norms  -> ncclAllReduce(sum) on the solver stream;
//   (2) stencil halo planes                        -> read directly from the neighbour's memory by the
//       stencil kernel through CUDA-IPC peer mappings over NVLink (spmm.cu, StencilDesc::halo_*).
// NCCL is loaded with dlopen so that single-GPU use has no NCCL dependency; the unique id is created by
// rank 0 and distributed by the launcher (torch.distributed in bench.py / lobpcg_b200/dist.py).
#include <dlfcn.h>
#include <cstring>

#include "common.cuh"
#include "context.h"
#include "solver.h"

namespace lb2 {

// minimal NCCL surface (binary-compatible with nccl.h of NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSum_ = 0 };
enum { ncclFloat32_ = 7, ncclFloat64_ = 8 };

struct Comm {
  void* lib = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, size = 1;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static Comm* load_nccl(const char* path) {
  const char* cands[] = {path, "libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* c : cands) {
    if (!c || !*c) continue;
    lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    fprintf(stderr, "lobpcg_b200: cannot load NCCL (%s)\n", dlerror());
    return nullptr;
  }
  Comm* c = new Comm();
  c->lib = lib;
  c->GetUniqueId = (decltype(c->GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  c->CommInitRank = (decltype(c->CommInitRank))dlsym(lib, "ncclCommInitRank");
  c->AllReduce = (decltype(c->AllReduce))dlsym(lib, "ncclAllReduce");
  c->CommDestroy = (decltype(c->CommDestroy))dlsym(lib, "ncclCommDestroy");
  c->GetErrorString = (decltype(c->GetErrorString))dlsym(lib, "ncclGetErrorString");
  if (!c->GetUniqueId || !c->CommInitRank || !c->AllReduce || !c->CommDestroy) {
    fprintf(stderr, "lobpcg_b200: NCCL library lacks required symbols\n");
    delete c;
    return nullptr;
  }
  return c;
}

int comm_allreduce_impl(void* comm, void* buf, size_t count, bool is_double, cudaStream_t st) {
  Comm* c = (Comm*)comm;
  ncclResult_t r = c->AllReduce(buf, buf, count, is_double ? ncclFloat64_ : ncclFloat32_, ncclSum_, c->comm, st);
  if (r != 0) {
    fprintf(stderr, "lobpcg_b200: ncclAllReduce failed: %s\n", c->GetErrorString ? c->GetErrorString(r) : "?");
    return -1;
  }
  return 0;
}
int comm_rank_impl(void* comm) { return ((Comm*)comm)->rank; }
int comm_size_impl(void* comm) { return ((Comm*)comm)->size; }

}  // namespace lb2

using namespace lb2;

extern "C" {

// rank 0: fill 128 bytes with a fresh NCCL unique id
int lb2_comm_unique_id(void* out128, const char* nccl_lib_path) {
  Comm* c = load_nccl(nccl_lib_path);
  if (!c) return -1;
  ncclUniqueId id;
  ncclResult_t r = c->GetUniqueId(&id);
  if (r == 0) memcpy(out128, &id, 128);
  delete c;  // library handle stays loaded
  return r == 0 ? 0 : -1;
}

// every rank: attach an NCCL communicator to the context (collective call)
int lb2_ctx_attach_comm(lb2_ctx* ctx, int rank, int size, const void* unique_id128, const char* nccl_lib_path) {
  if (!ctx || size < 1 || rank < 0 || rank >= size) return -1;
  if (size == 1) return 0;
  Comm* c = load_nccl(nccl_lib_path);
  if (!c) return -1;
  ncclUniqueId id;
  memcpy(&id, unique_id128, 128);
  LB2_CUDA_OK(cudaSetDevice(ctx->device));
  ncclResult_t r = c->CommInitRank(&c->comm, size, id, rank);
  if (r != 0) {
    fprintf(stderr, "lobpcg_b200: ncclCommInitRank failed: %s\n", c->GetErrorString ? c->GetErrorString(r) : "?");
    delete c;
    return -1;
  }
  c->rank = rank;
  c->size = size;
  ctx->comm = c;
  return 0;
}

int lb2_ctx_detach_comm(lb2_ctx* ctx) {
  if (!ctx || !ctx->comm) return 0;
  Comm* c = (Comm*)ctx->comm;
  cudaStreamSynchronize(ctx->stream);
  if (c->comm) c->CommDestroy(c->comm);
  delete c;
  ctx->comm = nullptr;
  return 0;
}

// sum `count` reals in place across ranks (exposed for tests)
int lb2_comm_allreduce(lb2_ctx* ctx, void* dev_buf, size_t count, int is_double) {
  return allreduce_sum(ctx, dev_buf, count, is_double != 0);
}

// ---- CUDA IPC: share a device allocation with the other ranks of the node (halo planes over NVLink) ----
int lb2_ipc_get_handle(void* dev_ptr, void* out64) {
  cudaIpcMemHandle_t h;
  LB2_CUDA_OK(cudaIpcGetMemHandle(&h, dev_ptr));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out64, &h, 64);
  return 0;
}
void* lb2_ipc_open_handle(const void* in64) {
  cudaIpcMemHandle_t h;
  memcpy(&h, in64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    fprintf(stderr, "lobpcg_b200: cudaIpcOpenMemHandle failed: %s\n", cudaGetErrorString(e));
    return nullptr;
  }
  return p;
}
int lb2_ipc_close_handle(void* p) {
  LB2_CUDA_OK(cudaIpcCloseMemHandle(p));
  return 0;
}

}  // extern "C"

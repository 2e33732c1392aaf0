// lobpcg_b200/csrc/hostcopy.cu — host <-> device transfer of the caller's PAGEABLE block vectors.
//
// The reference's state struct hands the solver calloc'ed host buffers (<p>_lobpcg_alloc, reference lobpcg.h:590-614):
// X0 goes up once, the eigenvectors come down once (n x sizeSub scalars each way; 9.8 GB at config C5).  A plain
// cudaMemcpy from pageable memory is staged by the driver through one thread (~10 GB/s); here several host threads
// copy slices of the block into / out of a ring of pinned chunks while the DMA engine moves the neighbouring chunk,
// which brings the transfer close to what PCIe gives.  Falls back to cudaMemcpyAsync for small transfers.
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"
#include "context.h"

namespace lb2 {

namespace {
constexpr size_t kChunk = (size_t)64 << 20;   // bytes per pinned chunk (copy threads are spawned per chunk: ~0.15 ms against ~1.6 ms of copying)
constexpr int kRing = 3;

struct HostCopyState {
  char* pinned[kRing] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev[kRing] = {nullptr, nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;
  int nthreads = 4;
};

void par_memcpy(char* dst, const char* src, size_t bytes, int nthreads) {
  if (bytes < ((size_t)4 << 20) || nthreads <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  std::vector<std::thread> th;
  const size_t per = ((bytes + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
  for (int t = 1; t < nthreads; t++) {
    const size_t o = per * t;
    if (o >= bytes) break;
    th.emplace_back([=] { memcpy(dst + o, src + o, std::min(per, bytes - o)); });
  }
  memcpy(dst, src, std::min(per, bytes));
  for (auto& x : th) x.join();
}

HostCopyState* state_of(lb2_ctx* ctx) {
  if (ctx->hostcopy) return (HostCopyState*)ctx->hostcopy;
  HostCopyState* s = new HostCopyState();
  for (int i = 0; i < kRing; i++) {
    if (cudaMallocHost((void**)&s->pinned[i], kChunk) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      for (int j = 0; j <= i; j++) {
        if (s->pinned[j]) cudaFreeHost(s->pinned[j]);
        if (s->ev[j]) cudaEventDestroy(s->ev[j]);
      }
      delete s;
      return nullptr;
    }
  }
  if (cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) != cudaSuccess) s->copy_stream = nullptr;
  const unsigned hw = std::thread::hardware_concurrency();
  // r02 probe on the pool's 16-CPU hosts (tools/hostcopy_probe.py): 4 threads 20 / 25 GB/s (h2d / d2h), 8: 38 / 29, 16: 40 / 36
  s->nthreads = (int)std::max(1u, std::min(16u, hw ? hw : 4u));
  ctx->hostcopy = s;
  return s;
}
}  // namespace

void hostcopy_free(lb2_ctx* ctx) {
  HostCopyState* s = (HostCopyState*)ctx->hostcopy;
  if (!s) return;
  for (int i = 0; i < kRing; i++) {
    if (s->pinned[i]) cudaFreeHost(s->pinned[i]);
    if (s->ev[i]) cudaEventDestroy(s->ev[i]);
  }
  if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
  delete s;
  ctx->hostcopy = nullptr;
}

// dst <- src, `bytes` contiguous; to_device: src is pageable host memory, dst device memory (and vice versa).
// Synchronous with respect to the host AND ordered after everything already enqueued on ctx->stream.
int host_copy(lb2_ctx* ctx, void* dst, const void* src, size_t bytes, bool to_device) {
  if (bytes == 0) return 0;
  HostCopyState* s = bytes >= 4 * kChunk ? state_of(ctx) : nullptr;
  if (!s || !s->copy_stream) {
    LB2_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));   // producers / consumers of the device block are done
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  auto len = [&](size_t c) { return std::min(kChunk, bytes - c * kChunk); };
  if (to_device) {
    for (size_t c = 0; c < nchunks; c++) {
      const int b = (int)(c % kRing);
      if (c >= (size_t)kRing) LB2_CUDA_OK(cudaEventSynchronize(s->ev[b]));   // DMA out of this chunk finished
      par_memcpy(s->pinned[b], (const char*)src + c * kChunk, len(c), s->nthreads);
      LB2_CUDA_OK(cudaMemcpyAsync((char*)dst + c * kChunk, s->pinned[b], len(c), cudaMemcpyHostToDevice, s->copy_stream));
      LB2_CUDA_OK(cudaEventRecord(s->ev[b], s->copy_stream));
    }
    LB2_CUDA_OK(cudaStreamSynchronize(s->copy_stream));
  } else {
    const size_t ahead = kRing - 1;
    for (size_t c = 0; c < std::min(ahead, nchunks); c++) {
      const int b = (int)(c % kRing);
      LB2_CUDA_OK(cudaMemcpyAsync(s->pinned[b], (const char*)src + c * kChunk, len(c), cudaMemcpyDeviceToHost, s->copy_stream));
      LB2_CUDA_OK(cudaEventRecord(s->ev[b], s->copy_stream));
    }
    for (size_t c = 0; c < nchunks; c++) {
      const int b = (int)(c % kRing);
      const size_t nx = c + ahead;
      if (nx < nchunks) {   // its buffer was drained by the host copy of chunk nx - kRing (previous iteration)
        const int bn = (int)(nx % kRing);
        LB2_CUDA_OK(cudaMemcpyAsync(s->pinned[bn], (const char*)src + nx * kChunk, len(nx), cudaMemcpyDeviceToHost, s->copy_stream));
        LB2_CUDA_OK(cudaEventRecord(s->ev[bn], s->copy_stream));
      }
      LB2_CUDA_OK(cudaEventSynchronize(s->ev[b]));
      par_memcpy((char*)dst + c * kChunk, s->pinned[b], len(c), s->nthreads);
    }
  }
  return 0;
}


// Strided version for row-partitioned runs: `cols` column segments of `rows_bytes` contiguous bytes each, host column stride
// ld_host_bytes, device column stride ld_dev_bytes.  Segments are packed back to back into the pinned ring by the copy threads
// and moved by one DMA per segment (one per chunk when the device block is contiguous, ld_dev_bytes == rows_bytes).
int host_copy_2d(lb2_ctx* ctx, void* dev, size_t ld_dev_bytes, void* host, size_t ld_host_bytes, size_t rows_bytes, int cols,
                 bool to_device) {
  if (rows_bytes == 0 || cols <= 0) return 0;
  if (ld_dev_bytes == rows_bytes && ld_host_bytes == rows_bytes) return host_copy(ctx, to_device ? dev : host, to_device ? host : dev, rows_bytes * (size_t)cols, to_device);
  HostCopyState* s = rows_bytes * (size_t)cols >= 4 * kChunk ? state_of(ctx) : nullptr;
  if (!s || !s->copy_stream) {
    LB2_CUDA_OK(cudaMemcpy2DAsync(to_device ? dev : host, to_device ? ld_dev_bytes : ld_host_bytes, to_device ? host : dev,
                                  to_device ? ld_host_bytes : ld_dev_bytes, rows_bytes, cols,
                                  to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  // pieces: (column, byte range inside the segment), each at most kChunk bytes; a ring slot holds consecutive pieces
  struct Piece { int col; size_t off, len; };
  std::vector<std::vector<Piece>> groups;
  {
    std::vector<Piece> cur;
    size_t fill = 0;
    for (int j = 0; j < cols; j++)
      for (size_t o = 0; o < rows_bytes; o += kChunk) {
        const size_t l = std::min(kChunk, rows_bytes - o);
        if (fill + l > kChunk && !cur.empty()) { groups.push_back(cur); cur.clear(); fill = 0; }
        cur.push_back(Piece{j, o, l});
        fill += l;
      }
    if (!cur.empty()) groups.push_back(cur);
  }
  auto host_side = [&](const std::vector<Piece>& g, char* pin, bool gather) {   // pageable <-> pinned, several threads
    std::vector<size_t> start(g.size());
    size_t acc = 0;
    for (size_t i = 0; i < g.size(); i++) { start[i] = acc; acc += g[i].len; }
    const int nt = std::max(1, std::min<int>(s->nthreads, (int)g.size()));
    auto work = [&](int t) {
      for (size_t i = t; i < g.size(); i += nt) {
        char* h = (char*)host + (size_t)g[i].col * ld_host_bytes + g[i].off;
        if (gather) memcpy(pin + start[i], h, g[i].len);
        else memcpy(h, pin + start[i], g[i].len);
      }
    };
    if (nt == 1) {
      if (g.size() == 1) {
        char* h = (char*)host + (size_t)g[0].col * ld_host_bytes + g[0].off;
        if (gather) par_memcpy(pin, h, g[0].len, s->nthreads); else par_memcpy(h, pin, g[0].len, s->nthreads);
      } else work(0);
      return;
    }
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
  };
  auto dma = [&](const std::vector<Piece>& g, char* pin) -> int {
    size_t acc = 0;
    for (const Piece& pc : g) {
      char* d = (char*)dev + (size_t)pc.col * ld_dev_bytes + pc.off;
      if (to_device) LB2_CUDA_OK(cudaMemcpyAsync(d, pin + acc, pc.len, cudaMemcpyHostToDevice, s->copy_stream));
      else LB2_CUDA_OK(cudaMemcpyAsync(pin + acc, d, pc.len, cudaMemcpyDeviceToHost, s->copy_stream));
      acc += pc.len;
    }
    return 0;
  };
  const size_t ng = groups.size();
  if (to_device) {
    for (size_t c = 0; c < ng; c++) {
      const int b = (int)(c % kRing);
      if (c >= (size_t)kRing) LB2_CUDA_OK(cudaEventSynchronize(s->ev[b]));
      host_side(groups[c], s->pinned[b], true);
      if (dma(groups[c], s->pinned[b])) return -1;
      LB2_CUDA_OK(cudaEventRecord(s->ev[b], s->copy_stream));
    }
    LB2_CUDA_OK(cudaStreamSynchronize(s->copy_stream));
  } else {
    const size_t ahead = kRing - 1;
    for (size_t c = 0; c < std::min(ahead, ng); c++) {
      const int b = (int)(c % kRing);
      if (dma(groups[c], s->pinned[b])) return -1;
      LB2_CUDA_OK(cudaEventRecord(s->ev[b], s->copy_stream));
    }
    for (size_t c = 0; c < ng; c++) {
      const int b = (int)(c % kRing);
      const size_t nx = c + ahead;
      if (nx < ng) {
        const int bn = (int)(nx % kRing);
        if (dma(groups[nx], s->pinned[bn])) return -1;
        LB2_CUDA_OK(cudaEventRecord(s->ev[bn], s->copy_stream));
      }
      LB2_CUDA_OK(cudaEventSynchronize(s->ev[b]));
      host_side(groups[c], s->pinned[b], false);
    }
  }
  return 0;
}

// cap on the copy threads of this context (multi-GPU runs in one process share the host cores)
void hostcopy_set_threads(lb2_ctx* ctx, int nthreads) {
  HostCopyState* s = state_of(ctx);
  if (s && nthreads >= 1) s->nthreads = nthreads;
}

}  // namespace lb2

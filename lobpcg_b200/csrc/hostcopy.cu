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
constexpr size_t kChunk = (size_t)32 << 20;   // bytes per pinned chunk
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
  s->nthreads = (int)std::max(1u, std::min(8u, hw ? hw / 2 : 4u));
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

}  // namespace lb2

// lobpcg_b200/csrc/tile_loader.cuh — per-thread cp.async tile loader shared by the f64 DMMA kernels
// (dense.cu: gram_dmma_kernel / tall_nn_dmma_kernel, gram_wl.cu: gram_wl_kernel).
#pragma once
#include "common.cuh"

namespace lb2 {

// Per-thread view of the same copy: the (column, chunk) slots of a thread are the same for every K chunk, so the
// source pointer of slot 0 is kept in a register pair and advanced by a uniform step; slot s is slot 0 plus
// s * slot_stride in global memory and a compile-time offset in shared memory.  This removes the per-copy index
// arithmetic (~25 integer instructions per LDGSTS in the first version, which kept the two warps of a scheduler
// away from the DMMA pipe: ncu showed 80 % pipe utilisation with `wait` as the second stall reason).
template <int NCOLS, int RUN, int LDS, int NT, bool VEC>
struct TileLoaderF64 {
  static constexpr int EPC = VEC ? 2 : 1;             // elements per copy
  static constexpr int CPC = RUN / EPC;               // copies per column
  static constexpr int TOTAL = NCOLS * CPC;
  static constexpr int NSLOT = TOTAL / NT;
  static constexpr int CSTEP = NT / CPC;              // columns between consecutive slots of a thread
  static_assert(TOTAL % NT == 0 && NT % CPC == 0, "tile / thread-count mismatch");
  const double* p;        // slot-0 source for the current chunk
  int64_t slot_stride;    // CSTEP * ld
  int soff;               // slot-0 offset inside a stage
  int roff;               // offset of this thread's copy inside the run
  unsigned colmask;       // bit s: column of slot s exists
  __device__ __forceinline__ void init(const double* base, int64_t ld, int64_t run0, int col0, int col_end, int tid) {
    const int c = tid / CPC, ch = tid % CPC;
    roff = ch * EPC;
    soff = c * LDS + roff;
    slot_stride = (int64_t)CSTEP * ld;
    colmask = 0;
#pragma unroll
    for (int s = 0; s < NSLOT; s++)
      if (col0 + c + s * CSTEP < col_end) colmask |= 1u << s;
    // clamp the pointer of non-existing columns to a valid address (never dereferenced: src-size 0)
    p = base + (int64_t)(col0 + c) * ld + run0 + roff;
  }
  // copy one chunk into `stage`; `valid` = number of run elements that exist from the start of this chunk
  __device__ __forceinline__ void issue(double* stage, const double* safe, int64_t valid) const {
    if (valid >= RUN) {
#pragma unroll
      for (int s = 0; s < NSLOT; s++) {
        const bool ok = (colmask >> s) & 1u;
        cp_async_zfill<EPC * 8>(stage + soff + s * CSTEP * LDS, ok ? p + s * slot_stride : safe, ok ? EPC * 8 : 0);
      }
    } else {
      const int64_t left = valid - roff;
      const int bytes = left >= EPC ? EPC * 8 : (left > 0 ? (int)left * 8 : 0);
#pragma unroll
      for (int s = 0; s < NSLOT; s++) {
        const bool ok = ((colmask >> s) & 1u) && bytes > 0;
        cp_async_zfill<EPC * 8>(stage + soff + s * CSTEP * LDS, ok ? p + s * slot_stride : safe, ok ? bytes : 0);
      }
    }
  }
  __device__ __forceinline__ void advance(int64_t step) { p += step; }
};

}  // namespace lb2

// lobpcg_b200/csrc/gram_i8.cu — f64 Gram products on the 5th-generation tensor cores through an Ozaki-style split
// (option gram_i8; the DMMA work-list kernel of gram_wl.cu stays the default).
//
// tcgen05.mma has no f64 kind: FP64 runs on mma.sync DMMA at 37 TFLOP/s, the hard ceiling of gram_wl.cu / dense.cu.  The only
// Blackwell-native way past it is integer arithmetic: tcgen05.mma kind::i8 (SASS UTCIMMA) multiplies signed bytes EXACTLY
// into 32-bit integer accumulators in TMEM.  A product G = A^H B (contraction over the n rows) becomes
//
//   A[r, a] = 2^(eA[a] - 55) * N_A[r, a],  N = rint(x * 2^(55 - e)),  |N| < 2^54      (e from the column's largest entry)
//   N = sum_i d_i 256^i,  i = 0..6,  d_i in [-128, 127]                               (balanced base-256 digits = 7 int8 "slices")
//   G[a, b] = 2^(eA[a] + eB[b] - 110) * sum_L 256^L * sum_{i + j = L} sum_r dA_i[r, a] dB_j[r, b]
//
// with the 28 slice products of the levels L = 6..12 (the 21 products below contribute < 2^-53 of the full scale and are
// dropped).  Every slice product is an exact integer GEMM; the per-level sums are exact in int32 for 16 384 rows
// (7 * 2^14 * 2^14 < 2^31), are drained from TMEM into exact int64 partial sums, and the only roundings are the input
// quantisation (2^-55 of the column maximum per element, i.e. finer than the f64 ulp of the large entries) and ONE rounding
// per output element when the seven level sums are combined in f64.  Deterministic by construction (integer sums).
//
// Kernels:
//   oz_absmax_kernel / oz_exp_kernel   column maxima -> exponents e[c]
//   oz_split_kernel                    f64 block -> 7 int8 slices in a TILED layout: [row chunk of 128][slice][column][128 B].
//                                      One (chunk, slice, 128-column panel) tile is 16 KB CONTIGUOUS — it is exactly one TMA box
//                                      and one canonical K-major SWIZZLE_128B UMMA operand (a column of S = one 128-byte row).
//   oz_gram_kernel                     persistent, warp-specialised: warp 0 = TMA producer (one cp.async.bulk.tensor.4d per slice
//                                      tile into a ring of twelve 16 KB slots, full/empty mbarriers per slot), warp 1 = MMA issuer
//                                      (one thread: 4 UTCIMMA of K = 32 per slice product, tcgen05.commit releases slots as their last
//                                      use retires), warps 2-5 = drain (tcgen05.ld of the level accumulators every 128 chunks).
//                                      TMEM holds 4 level accumulators of 128 x 128 int32 (all 512 columns), so the 7 levels are two
//                                      work items per output tile: "lo" = levels 9..12 (10 products, slices 3..6) and "hi" = levels
//                                      6..8 (18 products, all slices).  The slice tiles of a chunk arrive in the order B6 A0 B5 A1 ...
//                                      so that every A slice meets the (at most three) B slices it multiplies as a sliding window.
//   oz_reduce_kernel                   int64 partials of all items of a tile -> f64 G (+ Hermitian mirror).
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cmath>
#include <climits>
#include <cuda.h>

#include "common.cuh"
#include "context.h"
#include "kernels.h"

namespace lb2 {

namespace {

constexpr int OZ_S = 7;                    // slices (base-256 digits)
constexpr int OZ_CH = 128;                 // rows per chunk = bytes per tile row
constexpr int OZ_T = 128;                  // tile edge
constexpr uint32_t OZ_TILE = OZ_T * OZ_CH; // 16 KB
constexpr int OZ_NS = 12;                  // ring slots
constexpr int OZ_SEG = 128;                // chunks per TMEM accumulation segment (16 384 rows)
constexpr int OZ_NT = 192;                 // warp 0 producer, warp 1 MMA, warps 2..5 drain
constexpr uint32_t OZ_BAR = OZ_NS * OZ_TILE;
constexpr uint32_t OZ_SMEM = OZ_BAR + 1024 + 1024;   // + barriers, + 1 KB alignment slack
constexpr int OZ_SHIFT = 55;               // N = rint(x * 2^(55 - e))

struct OzItem {
  int32_t chunk_begin, chunk_end;   // row chunks of this piece
  int32_t a_col0, b_col0;           // first column of the A / B panel (in the slice arrays' column numbering)
  int32_t b_sel;                    // 0: B panel from the first B slice array, 1: from the second
  int32_t n16;                      // UMMA N: B-panel columns rounded up to 16
  int32_t group;                    // 0: levels 9..12 (slices 3..6), 1: levels 6..8 (all slices)
  int32_t tile;                     // output tile index
};
struct OzTile {
  int32_t a_col0, a_cols, b_col0, b_cols, b_sel;
  int32_t g_row0, g_col0;           // where the tile goes in G
  int32_t first, last;              // items first, first + stride, ... below last
  int32_t stride;                   // item index stride (0 = 1)
  int32_t rank;                     // CTA rank inside the cluster that computes this tile (cluster kernel), else 0
  int32_t diag;                     // 1: diagonal tile of a Hermitian product (lower part not written, mirrored instead)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
// bounded wait: a lost arrival traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 28); spin++) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
// The producer and MMA warps run WARP-UNIFORM code (all 32 lanes, identical values) and every single-thread instruction is
// predicated on elect.sync inside its asm block.  Issuing from inside an `if (lane == 0)` region instead makes the operands
// thread-private in the compiler's eyes: ptxas then wraps every UTCIMMA / UTMALDG / UTCBAR (which take uniform registers) in a
// "waterfall" loop — ELECT, six R2UR.BROADCAST, branch — about 130 clocks per MMA against 64 clocks of MMA execution
// (r02 SASS; the issue thread, not memory, was what held the tensor pipe at 48 %).
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}\n" ::"r"(bar),
      "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int32_t c0, int32_t c1, int32_t c2,
                                            int32_t c3) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}\n" ::"r"(
          dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// L2 prefetch of a box (UTMAPF.L2): no shared memory, no barrier — the later load of the same box finds it in L2
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* tm, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];\n\t}\n" ::"l"(tm),
      "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (same as gram_tc5.cu): 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((1024u >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = S32 (2 << 4), A = B = signed 8-bit (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t oz_idesc(int n16) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n16 >> 3) << 17) | ((uint32_t)(OZ_T >> 4) << 24);
}
// both operand descriptors share the constant upper word (SBO = 1024 B, version 1, SWIZZLE_128B); the lower word is the
// shared-memory address >> 4 — one integer add per MMA in the issuing thread
constexpr uint32_t OZ_DESC_HI = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(OZ_DESC_HI)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(bar)
      : "memory");
}

// ---------------------------------------------------------------------------------------------------- exponents
__global__ void __launch_bounds__(256)
    oz_absmax_kernel(const double* __restrict__ X, int64_t ld, int64_t n, int64_t rows_per_cta, unsigned long long* __restrict__ mx) {
  const int c = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
  const double* x = X + (int64_t)c * ld;
  unsigned long long m = 0;
  for (int64_t r = r0 + threadIdx.x; r < r1; r += 256) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(x[r]));   // non-negative doubles order like integers
    m = b > m ? b : m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t > m ? t : m;
  }
  if ((threadIdx.x & 31) == 0 && m) atomicMax(mx + c, m);
}
// e[c]: every |x| of the column is < 2^(e - 1), so |N| = |rint(x 2^(55 - e))| <= 2^54
__global__ void oz_exp_kernel(const unsigned long long* __restrict__ mx, int m, int* __restrict__ e, int* __restrict__ nonfinite) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  const unsigned long long b = mx[c];
  if (((b >> 52) & 0x7FF) == 0x7FF) atomicOr(nonfinite, 1);   // Inf / NaN in the block: the outputs become NaN (flag read by the reduce / drain)
  int ex = 0;
  if (b) {
    ex = (int)((b >> 52) & 0x7FF) - 1023 + 2;
    ex = max(-900, min(900, ex));
  }
  e[c] = ex;
}

// ---------------------------------------------------------------------------------------------------- split
// out[((rc * 7 + s) * m + c) * 128 + rr]: warp = one column of one chunk, lane = 4 consecutive rows (32 bytes in, 7 x 4 bytes out)
template <bool TRACK>
__global__ void __launch_bounds__(256)
    oz_split_kernel(const double* __restrict__ X, int64_t ld, int64_t n, int m, const int* __restrict__ e, int8_t* __restrict__ out,
                    int chunks_per_cta, unsigned long long* __restrict__ mx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.y * 8 + warp;
  if (c >= m) return;
  unsigned long long vmax = 0;   // TRACK: largest |x| of this warp's part of the column (the exponents came from a hint)
  const double scale = __longlong_as_double((long long)(1023 + OZ_SHIFT - e[c]) << 52);
  const double* x = X + (int64_t)c * ld;
  const bool vec = ((ld & 1) == 0) && (((uintptr_t)X & 15) == 0);
  for (int q = 0; q < chunks_per_cta; q++) {
    const int64_t rc = (int64_t)blockIdx.x * chunks_per_cta + q;
    const int64_t r = rc * OZ_CH + lane * 4;
    if (rc * OZ_CH >= n) break;
    double v[4];
    if (vec && r + 3 < n) {
      const double2 a = *reinterpret_cast<const double2*>(x + r), b = *reinterpret_cast<const double2*>(x + r + 2);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++) v[k] = (r + k < n) ? x[r + k] : 0.0;
    }
    long long N[4];
#pragma unroll
    for (int k = 0; k < 4; k++) N[k] = __double2ll_rn(v[k] * scale);
    if constexpr (TRACK) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(v[k]));
        vmax = b > vmax ? b : vmax;
      }
    }
    uint32_t* o = reinterpret_cast<uint32_t*>(out + ((rc * OZ_S) * (int64_t)m + c) * OZ_CH + lane * 4);
#pragma unroll
    for (int i = 0; i < OZ_S; i++) {
      uint32_t w = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int d = (int)(int8_t)(N[k] & 0xFF);   // balanced digit in [-128, 127]
        N[k] = (N[k] - d) >> 8;
        w |= (uint32_t)(d & 0xFF) << (8 * k);
      }
      o[(int64_t)i * m * (OZ_CH / 4)] = w;
    }
  }
  if constexpr (TRACK) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long t = __shfl_xor_sync(0xffffffffu, vmax, o);
      vmax = t > vmax ? t : vmax;
    }
    if (lane == 0 && vmax) atomicMax(mx + c, vmax);
  }
}

// After a split with HINTED exponents: were they right?  e_used[c] must cover the column's actual maximum (else digits
// overflowed) and must not exceed what the maximum needs by more than 4 bits (else precision was given away); either sets
// *redo.  The hint for the next split of this operand role is the exponent the fresh maximum needs plus one guard bit.
__global__ void oz_hint_check_kernel(const unsigned long long* __restrict__ mx, int m, const int* __restrict__ e_used,
                                     int* __restrict__ hint, int* __restrict__ redo, int* __restrict__ nonfinite) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  const unsigned long long b = mx[c];
  if (((b >> 52) & 0x7FF) == 0x7FF) atomicOr(nonfinite, 1);
  int need = 0;
  if (b) need = max(-900, min(900, (int)((b >> 52) & 0x7FF) - 1023 + 2));
  if (b && (need > e_used[c] || e_used[c] - need > 4)) atomicOr(redo, 1);
  if (!b && e_used[c] != 1) { /* an all-zero column is exact with any exponent */ }
  hint[c] = need + 1;
}
// fresh exponents -> hint (+ 1 guard bit) after a full (absmax) split
__global__ void oz_hint_store_kernel(const int* __restrict__ e, int m, int* __restrict__ hint) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < m) hint[c] = e[c] + 1;
}

// ---------------------------------------------------------------------------------------------------- main kernel
struct OzGroup { int amin, smin, lmin, lmax; };
__device__ __forceinline__ OzGroup oz_group(int g) { return g ? OzGroup{0, 0, 6, 8} : OzGroup{3, 3, 9, 12}; }

// MMAs of one row chunk of a level group, issued by ONE thread.  Slice tiles arrive in ring order B6 A_amin B5 A_amin+1 ...:
// before A_a come the B slices first needed at a (j = lmin - a); A_a multiplies B_j for j in [max(smin, lmin - a), min(6, lmax - a)]
// (level a + j -> accumulator a + j - lmin), then A_a and every B slice whose last use this was are released.  Everything is
// unrolled at compile time: ring slots and parities live in registers, a descriptor costs one add.
template <int AMIN, int SMIN, int LMIN, int LMAX>
__device__ __forceinline__ void oz_mma_chunk(uint32_t sbase, uint32_t bar_full, uint32_t bar_empty, uint32_t tmem, uint32_t idesc,
                                             uint32_t& slot, uint32_t& par, uint32_t& level_init) {
  uint32_t sB[OZ_S], pB[OZ_S];
  auto advance = [&]() {
    if (++slot == OZ_NS) { slot = 0; par ^= 1u; }
  };
#pragma unroll
  for (int a = AMIN; a < OZ_S; ++a) {
    const int jlo = (LMIN - a) > SMIN ? (LMIN - a) : SMIN;
    const int jhi = (LMAX - a) < (OZ_S - 1) ? (LMAX - a) : (OZ_S - 1);
    const int jprev = (a == AMIN) ? OZ_S : ((LMIN - (a - 1)) > SMIN ? (LMIN - (a - 1)) : SMIN);   // B slices [jlo, jprev) arrive now
#pragma unroll
    for (int j = OZ_S - 1; j >= 0; --j)
      if (j < jprev && j >= jlo) { sB[j] = slot; pB[j] = par; advance(); }
    const uint32_t sA = slot, pA = par;
    advance();
#pragma unroll
    for (int j = OZ_S - 1; j >= 0; --j)
      if (j < jprev && j >= jlo) mbar_wait(bar_full + 8 * sB[j], pB[j]);
    mbar_wait(bar_full + 8 * sA, pA);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t a_lo = (sbase + sA * OZ_TILE) >> 4;
#pragma unroll
    for (int j = 0; j < OZ_S; ++j)
      if (j >= jlo && j <= jhi) {
        const int L = a + j;
        const uint32_t acc = tmem + (uint32_t)((L - LMIN) * OZ_T);
        const uint32_t b_lo = (sbase + sB[j] * OZ_TILE) >> 4;
        const uint32_t init = (level_init >> L) & 1u;
#pragma unroll
        for (int ks = 0; ks < OZ_CH / 32; ks++) umma_i8(acc, a_lo + 2u * ks, b_lo + 2u * ks, idesc, ks > 0 ? 1u : init);
        level_init |= 1u << L;
      }
    umma_commit(bar_empty + 8 * sA);
#pragma unroll
    for (int j = 0; j < OZ_S; ++j)
      if (j >= jlo && j <= jhi && ((a < OZ_S - 1) ? (j == LMAX - a) : true)) umma_commit(bar_empty + 8 * sB[j]);
  }
}

__global__ void __launch_bounds__(OZ_NT, 1)
    oz_gram_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                   const __grid_constant__ CUtensorMap tmB1, const OzItem* __restrict__ items, const int* __restrict__ cta_first,
                   long long* __restrict__ part, int pf) {
  extern __shared__ __align__(1024) unsigned char smem_oz[];
  const uint32_t sbase = (smem_u32(smem_oz) + 1023u) & ~1023u;
  unsigned char* gbase = smem_oz + (sbase - smem_u32(smem_oz));
  const uint32_t bar_full = sbase + OZ_BAR, bar_empty = bar_full + 8 * OZ_NS, bar_accf = bar_empty + 8 * OZ_NS, bar_acce = bar_accf + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + OZ_BAR + 8 * (2 * OZ_NS + 2) + 8);
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
  const int it0 = cta_first[blockIdx.x], it1 = cta_first[blockIdx.x + 1];

  if (tid == 0) {
    for (int s = 0; s < OZ_NS; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_accf, 1);
    mbar_init(bar_acce, 128);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB0) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB1) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32((const void*)tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer
    {   // all 32 lanes: warp-uniform control flow, single-thread instructions elected inside their asm blocks
      uint32_t q = 0;
      auto load = [&](const CUtensorMap* tm, int col0, int slice, int chunk) {
        const uint32_t s = q % OZ_NS;
        if (q >= OZ_NS) mbar_wait(bar_empty + 8 * s, ((q / OZ_NS) - 1) & 1);
        mbar_expect_tx(bar_full + 8 * s, OZ_TILE);
        tma_load_4d(sbase + s * OZ_TILE, tm, bar_full + 8 * s, 0, col0, slice, chunk);
        q++;
      };
      for (int it = it0; it < it1; ++it) {
        const OzItem im = items[it];
        const OzGroup g = oz_group(im.group);
        const CUtensorMap* tb = im.b_sel ? &tmB1 : &tmB0;
        for (int chunk = im.chunk_begin; chunk < im.chunk_end; ++chunk) {
          // the ring (192 KB) divided by the slot round trip bounds the fill rate, and most of the round trip is DRAM latency:
          // the boxes of chunk + pf are requested into L2 now, so that their loads a few microseconds later are L2 hits
          const int pc = chunk + pf;
          if (pf > 0 && pc < im.chunk_end)
            for (int sl = g.smin; sl < OZ_S; ++sl) {
              tma_prefetch_4d(tb, 0, im.b_col0, sl, pc);
              tma_prefetch_4d(&tmA, 0, im.a_col0, sl, pc);
            }
          int jnext = OZ_S - 1;
          for (int a = g.amin; a < OZ_S; ++a) {
            const int jlo = max(g.smin, g.lmin - a);
            while (jnext >= jlo) { load(tb, im.b_col0, jnext, chunk); --jnext; }
            load(&tmA, im.a_col0, a, chunk);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    {   // all 32 lanes: warp-uniform control flow, single-thread instructions elected inside their asm blocks
      uint32_t slot = 0, par = 0, segs = 0;
      for (int it = it0; it < it1; ++it) {
        const OzItem im = items[it];
        const uint32_t idesc = oz_idesc(im.n16);
        for (int seg0 = im.chunk_begin; seg0 < im.chunk_end; seg0 += OZ_SEG, ++segs) {
          const int seg1 = min(im.chunk_end, seg0 + OZ_SEG);
          if (segs > 0) mbar_wait(bar_acce, (segs - 1) & 1);   // the drain warps have read the previous segment's accumulators
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          uint32_t level_init = 0;
          if (im.group) {
            for (int chunk = seg0; chunk < seg1; ++chunk)
              oz_mma_chunk<0, 0, 6, 8>(sbase, bar_full, bar_empty, tmem, idesc, slot, par, level_init);
          } else {
            for (int chunk = seg0; chunk < seg1; ++chunk)
              oz_mma_chunk<3, 3, 9, 12>(sbase, bar_full, bar_empty, tmem, idesc, slot, par, level_init);
          }
          umma_commit(bar_accf);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ drain (warps 2..5)
    const int lq = warp & 3;            // TMEM lane quarter this warp may read
    const int row = lq * 32 + lane;     // tile row = column of the A panel
    uint32_t segs = 0;
    for (int it = it0; it < it1; ++it) {
      const OzItem im = items[it];
      const int nlev = im.group ? 3 : 4;
      long long* p0 = part + (int64_t)it * (4 * OZ_T * OZ_T) + row;
      for (int seg0 = im.chunk_begin; seg0 < im.chunk_end; seg0 += OZ_SEG, ++segs) {
        mbar_wait(bar_accf, segs & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const bool first = seg0 == im.chunk_begin;
        for (int lev = 0; lev < nlev; ++lev)
          for (int cq = 0; cq * 32 < im.n16; ++cq) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(lev * OZ_T + cq * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            long long* p = p0 + (int64_t)(lev * OZ_T + cq * 32) * OZ_T;
#pragma unroll
            for (int i = 0; i < 32; i++) {
              const long long add = (long long)(int32_t)v[i];
              p[(int64_t)i * OZ_T] = first ? add : p[(int64_t)i * OZ_T] + add;
            }
          }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        mbar_arrive(bar_acce);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------------- cluster kernel
// The kernel above is bound by operand traffic, not by the tensor pipe (ncu, 640 x 256: 110 GB of DRAM reads for 26 GB of
// slices, tensor pipe 48 % busy): an int8 slice set is as many bytes as the f64 block, but the MMAs retire 2.5 x faster than
// DMMA.  Here four CTAs of a thread-block cluster compute a 2 x 2 SUPER-TILE (two A panels x two B panels, same row chunks, same
// level group) and share their operands in hardware: every slice tile is needed by two CTAs, each of them fetches HALF of it
// (a 64-column box) and the TMA unit MULTICASTS the box into both CTAs' shared memory (UTMALDG.MULTICAST), so a CTA pulls
// half the bytes through L2 for the same MMAs.  A ring slot is refilled only when BOTH consumers have released it: the slot's
// "empty" mbarrier counts two arrivals and tcgen05.commit.multicast delivers each consumer's arrival to both CTAs.
// ring slots of the A / B slice tiles (B slots first in shared memory): template parameters NA + NB = 12 (option oz_ring)

struct OzCItem {
  int32_t chunk_begin, chunk_end, group, pad;
  int32_t a_col0[2], b_col0[2], b_sel[2], n16[2];
};

__device__ __forceinline__ void tma_load_4d_mc(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int32_t c0, int32_t c1, int32_t c2,
                                               int32_t c3, uint16_t mask) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], "
      "[%2], %7;\n\t}\n" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}\n" ::"r"(bar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

struct OzRing { uint32_t slot, par; };

template <int AMIN, int SMIN, int LMIN, int LMAX, int OZ_NA, int OZ_NB>
__device__ __forceinline__ void oz_mma_chunk_c(uint32_t sbase, uint32_t fullA, uint32_t emptyA, uint32_t fullB, uint32_t emptyB,
                                               uint32_t tmem, uint32_t idesc, OzRing& ra, OzRing& rb, uint32_t& level_init,
                                               uint16_t maskA, uint16_t maskB) {
  uint32_t sB[OZ_S], pB[OZ_S];
#pragma unroll
  for (int a = AMIN; a < OZ_S; ++a) {
    const int jlo = (LMIN - a) > SMIN ? (LMIN - a) : SMIN;
    const int jhi = (LMAX - a) < (OZ_S - 1) ? (LMAX - a) : (OZ_S - 1);
    const int jprev = (a == AMIN) ? OZ_S : ((LMIN - (a - 1)) > SMIN ? (LMIN - (a - 1)) : SMIN);
#pragma unroll
    for (int j = OZ_S - 1; j >= 0; --j)
      if (j < jprev && j >= jlo) {
        sB[j] = rb.slot; pB[j] = rb.par;
        if (++rb.slot == OZ_NB) { rb.slot = 0; rb.par ^= 1u; }
      }
    const uint32_t sA = ra.slot, pA = ra.par;
    if (++ra.slot == OZ_NA) { ra.slot = 0; ra.par ^= 1u; }
#pragma unroll
    for (int j = OZ_S - 1; j >= 0; --j)
      if (j < jprev && j >= jlo) mbar_wait(fullB + 8 * sB[j], pB[j]);
    mbar_wait(fullA + 8 * sA, pA);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t a_lo = (sbase + (OZ_NB + sA) * OZ_TILE) >> 4;
#pragma unroll
    for (int j = 0; j < OZ_S; ++j)
      if (j >= jlo && j <= jhi) {
        const int L = a + j;
        const uint32_t acc = tmem + (uint32_t)((L - LMIN) * OZ_T);
        const uint32_t b_lo = (sbase + sB[j] * OZ_TILE) >> 4;
        const uint32_t init = (level_init >> L) & 1u;
#pragma unroll
        for (int ks = 0; ks < OZ_CH / 32; ks++) umma_i8(acc, a_lo + 2u * ks, b_lo + 2u * ks, idesc, ks > 0 ? 1u : init);
        level_init |= 1u << L;
      }
    umma_commit_mc(emptyA + 8 * sA, maskA);
#pragma unroll
    for (int j = 0; j < OZ_S; ++j)
      if (j >= jlo && j <= jhi && ((a < OZ_S - 1) ? (j == LMAX - a) : true)) umma_commit_mc(emptyB + 8 * sB[j], maskB);
  }
}

template <int OZ_NA, int OZ_NB>
__global__ void __launch_bounds__(OZ_NT, 1)
    oz_gram_cluster_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                           const __grid_constant__ CUtensorMap tmB1, const OzCItem* __restrict__ items,
                           const int* __restrict__ cl_first, long long* __restrict__ part) {
  extern __shared__ __align__(1024) unsigned char smem_oz[];
  const uint32_t sbase = (smem_u32(smem_oz) + 1023u) & ~1023u;
  unsigned char* gbase = smem_oz + (sbase - smem_u32(smem_oz));
  const uint32_t fullB = sbase + OZ_BAR, emptyB = fullB + 8 * OZ_NB, fullA = emptyB + 8 * OZ_NB, emptyA = fullA + 8 * OZ_NA,
                 bar_accf = emptyA + 8 * OZ_NA, bar_acce = bar_accf + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + OZ_BAR + 8 * (2 * OZ_NS + 2) + 8);
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  const int ra = (int)(rank >> 1), rb = (int)(rank & 1);
  const uint16_t maskA = (uint16_t)(3u << (2 * ra)), maskB = (uint16_t)((1u << rb) | (1u << (rb + 2)));
  const int cl = blockIdx.x >> 2;
  const int it0 = cl_first[cl], it1 = cl_first[cl + 1];

  if (tid == 0) {
    for (int s = 0; s < OZ_NB; s++) { mbar_init(fullB + 8 * s, 1); mbar_init(emptyB + 8 * s, 2); }
    for (int s = 0; s < OZ_NA; s++) { mbar_init(fullA + 8 * s, 1); mbar_init(emptyA + 8 * s, 2); }
    mbar_init(bar_accf, 1);
    mbar_init(bar_acce, 128);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB0) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB1) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32((const void*)tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  cluster_sync_all();   // every CTA's barriers exist before a partner's TMA or commit can signal them
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    {   // all 32 lanes: warp-uniform control flow, single-thread instructions elected inside their asm blocks
      uint32_t qa = 0, qb = 0;
      for (int it = it0; it < it1; ++it) {
        const OzCItem im = items[it];
        const OzGroup g = oz_group(im.group);
        const CUtensorMap* tb = im.b_sel[rb] ? &tmB1 : &tmB0;
        const int acol = im.a_col0[ra] + rb * (OZ_T / 2), bcol = im.b_col0[rb] + ra * (OZ_T / 2);   // my half of each shared tile
        for (int chunk = im.chunk_begin; chunk < im.chunk_end; ++chunk) {
          int jnext = OZ_S - 1;
          for (int a = g.amin; a < OZ_S; ++a) {
            const int jlo = max(g.smin, g.lmin - a);
            while (jnext >= jlo) {
              const uint32_t s = qb % OZ_NB;
              if (qb >= OZ_NB) mbar_wait(emptyB + 8 * s, ((qb / OZ_NB) - 1) & 1);
              mbar_expect_tx(fullB + 8 * s, OZ_TILE);
              tma_load_4d_mc(sbase + s * OZ_TILE + ra * (OZ_TILE / 2), tb, fullB + 8 * s, 0, bcol, jnext, chunk, maskB);
              ++qb; --jnext;
            }
            const uint32_t s = qa % OZ_NA;
            if (qa >= OZ_NA) mbar_wait(emptyA + 8 * s, ((qa / OZ_NA) - 1) & 1);
            mbar_expect_tx(fullA + 8 * s, OZ_TILE);
            tma_load_4d_mc(sbase + (OZ_NB + s) * OZ_TILE + rb * (OZ_TILE / 2), &tmA, fullA + 8 * s, 0, acol, a, chunk, maskA);
            ++qa;
          }
        }
      }
    }
  } else if (warp == 1) {
    {   // all 32 lanes: warp-uniform control flow, single-thread instructions elected inside their asm blocks
      OzRing rga{0, 0}, rgb{0, 0};
      uint32_t segs = 0;
      for (int it = it0; it < it1; ++it) {
        const OzCItem im = items[it];
        const uint32_t idesc = oz_idesc(im.n16[rb]);
        for (int seg0 = im.chunk_begin; seg0 < im.chunk_end; seg0 += OZ_SEG, ++segs) {
          const int seg1 = min(im.chunk_end, seg0 + OZ_SEG);
          if (segs > 0) mbar_wait(bar_acce, (segs - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          uint32_t level_init = 0;
          if (im.group) {
            for (int chunk = seg0; chunk < seg1; ++chunk)
              oz_mma_chunk_c<0, 0, 6, 8, OZ_NA, OZ_NB>(sbase, fullA, emptyA, fullB, emptyB, tmem, idesc, rga, rgb, level_init, maskA, maskB);
          } else {
            for (int chunk = seg0; chunk < seg1; ++chunk)
              oz_mma_chunk_c<3, 3, 9, 12, OZ_NA, OZ_NB>(sbase, fullA, emptyA, fullB, emptyB, tmem, idesc, rga, rgb, level_init, maskA, maskB);
          }
          umma_commit(bar_accf);
        }
      }
    }
  } else {
    const int lq = warp & 3;
    const int row = lq * 32 + lane;
    uint32_t segs = 0;
    for (int it = it0; it < it1; ++it) {
      const OzCItem im = items[it];
      const int nlev = im.group ? 3 : 4;
      const int n16 = im.n16[rb];
      long long* p0 = part + ((int64_t)it * 4 + rank) * (4 * OZ_T * OZ_T) + row;
      for (int seg0 = im.chunk_begin; seg0 < im.chunk_end; seg0 += OZ_SEG, ++segs) {
        mbar_wait(bar_accf, segs & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const bool first = seg0 == im.chunk_begin;
        for (int lev = 0; lev < nlev; ++lev)
          for (int cq = 0; cq * 32 < n16; ++cq) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(lev * OZ_T + cq * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            long long* p = p0 + (int64_t)(lev * OZ_T + cq * 32) * OZ_T;
#pragma unroll
            for (int i = 0; i < 32; i++) {
              const long long add = (long long)(int32_t)v[i];
              p[(int64_t)i * OZ_T] = first ? add : p[(int64_t)i * OZ_T] + add;
            }
          }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        mbar_arrive(bar_acce);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while a partner may still multicast into it or signal its barriers
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------------- projection (tall NN)
// Out (n x nb) = S C on the same slices: the contraction now runs over the COLUMNS of S, so a slice tile of S — the very box
// the Gram kernels load, (128 rows of a chunk) x (128 columns) — is read as an MN-major operand (M = rows contiguous, K =
// columns; instruction-descriptor bit 15).  The column exponents of S are folded into the small matrix: C'[c, j] = 2^eS[c] C[c, j]
// is split into 7 slices with one exponent f[j] per output column (oz_cprep_kernel), laid out K-major like the Gram operands.
// A CTA owns row chunks and a 64-column output tile: 7 level accumulators x 64 columns fit TMEM (448 of 512 columns), so all
// 28 slice products of a K chunk run in one pass (tile order B0 .. B6 A6 A5 ... A0, A_a x B_j for j >= 6 - a).  The CTAs that
// share a row chunk (one per output tile) form a lock-step cohort: the slices of S come from DRAM once and from L2 for the rest.
// ring slots: A tiles 16 KB, B tiles 8 KB (B slots first): template parameters ON_NA, ON_NB with 16 NA + 8 NB = 192 KB, NB >= 8
constexpr uint32_t ON_BT = OZ_TILE / 2;                 // B tile: 64 output columns x 128 B of k
constexpr uint32_t ON_BAR = 12 * OZ_TILE;                 // barriers behind the 192 KB of ring slots
constexpr uint32_t ON_SMEM = ON_BAR + 1024 + 1024;
constexpr int ON_TN = 64;

// slices of C' = diag(2^eS) C: one CTA per output column j.  out[((kc * 7 + s) * nb + j) * 128 + (c % 128)], zero beyond kd.
__global__ void __launch_bounds__(128)
    oz_cprep_kernel(const double* __restrict__ C, int ldc, int kd, int nb, const int* __restrict__ eS, int8_t* __restrict__ out,
                    int* __restrict__ f, int nkc) {
  const int j = blockIdx.x;
  __shared__ int smax[4];
  // largest binary exponent of |C[c, j]| 2^eS[c]: frexp exponent + eS (a zero entry contributes nothing)
  int mx = INT_MIN;
  for (int c = threadIdx.x; c < kd; c += 128) {
    const double v = C[c + (int64_t)j * ldc];
    if (v != 0.0) {
      int ex;
      frexp(v, &ex);                       // |v| = m 2^ex, m in [0.5, 1)
      mx = max(mx, ex + eS[c]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = max(max(smax[0], smax[1]), max(smax[2], smax[3]));
  const int fj = (mx == INT_MIN) ? 0 : max(-900, min(900, mx + 1));      // every |C'| < 2^(fj - 1)
  if (threadIdx.x == 0) f[j] = fj;
  for (int c = threadIdx.x; c < nkc * OZ_CH; c += 128) {
    long long N = 0;
    if (c < kd) N = __double2ll_rn(ldexp(C[c + (int64_t)j * ldc], eS[c] + OZ_SHIFT - fj));
    int8_t* o = out + (((int64_t)(c / OZ_CH) * OZ_S) * nb + j) * OZ_CH + (c % OZ_CH);
#pragma unroll
    for (int i = 0; i < OZ_S; i++) {
      const int d = (int)(int8_t)(N & 0xFF);
      N = (N - d) >> 8;
      o[(int64_t)i * nb * OZ_CH] = (int8_t)d;
    }
  }
}

__device__ __forceinline__ uint32_t on_idesc(int n16) {   // as oz_idesc, A operand MN-major
  return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(n16 >> 3) << 17) | ((uint32_t)(OZ_T >> 4) << 24);
}

template <int ON_NA, int ON_NB>
__global__ void __launch_bounds__(OZ_NT, 1)
    oz_nn_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmC, const int* __restrict__ f,
                 double* __restrict__ Out, int64_t ldo, int64_t n, int nb, int nkc, int64_t nch, int njt, int ncoh,
                 const int* __restrict__ nonfinite, int pf, double alpha, double beta) {
  extern __shared__ __align__(1024) unsigned char smem_oz[];
  const uint32_t sbase = (smem_u32(smem_oz) + 1023u) & ~1023u;
  unsigned char* gbase = smem_oz + (sbase - smem_u32(smem_oz));
  const uint32_t abase0 = sbase + ON_NB * ON_BT;
  const uint32_t fullB = sbase + ON_BAR, emptyB = fullB + 8 * ON_NB, fullA = emptyB + 8 * ON_NB, emptyA = fullA + 8 * ON_NA,
                 bar_accf = emptyA + 8 * ON_NA, bar_acce = bar_accf + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + ON_BAR + 8 * (2 * ON_NA + 2 * ON_NB + 2) + 8);
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // CTA = (cohort, output tile): cohort c takes row chunks c, c + ncoh, ...; launch order puts the members of a cohort side by side
  const int jt = blockIdx.x % njt, coh = blockIdx.x / njt;
  const int j0 = jt * ON_TN;
  const int ncols = min(ON_TN, nb - j0);
  const int n16 = (ncols + 15) / 16 * 16;

  if (tid == 0) {
    for (int s = 0; s < ON_NB; s++) { mbar_init(fullB + 8 * s, 1); mbar_init(emptyB + 8 * s, 1); }
    for (int s = 0; s < ON_NA; s++) { mbar_init(fullA + 8 * s, 1); mbar_init(emptyA + 8 * s, 1); }
    mbar_init(bar_accf, 1);
    mbar_init(bar_acce, 128);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmS) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmC) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32((const void*)tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    {   // all 32 lanes: warp-uniform control flow, single-thread instructions elected inside their asm blocks
      uint32_t qa = 0, qb = 0;
      for (int64_t rc = coh; rc < nch; rc += ncoh)
        for (int kc = 0; kc < nkc; ++kc) {
          if (pf > 0) {   // S slices of a later K chunk into L2 (see oz_gram_kernel)
            int pk = kc + pf;
            int64_t prc = rc;
            while (pk >= nkc) { pk -= nkc; prc += ncoh; }
            if (prc < nch)
              for (int sl = 0; sl < OZ_S; ++sl) tma_prefetch_4d(&tmS, 0, pk * OZ_T, sl, (int32_t)prc);
          }
          // all seven B slices first (the first step, A slice 6, multiplies every one of them), then A slices 6 .. 0
          for (int j = 0; j < OZ_S; ++j) {
            const uint32_t s = qb % ON_NB;
            if (qb >= ON_NB) mbar_wait(emptyB + 8 * s, ((qb / ON_NB) - 1) & 1);
            mbar_expect_tx(fullB + 8 * s, ON_BT);
            tma_load_4d(sbase + s * ON_BT, &tmC, fullB + 8 * s, 0, j0, j, kc);
            ++qb;
          }
          for (int a = OZ_S - 1; a >= 0; --a) {
            const uint32_t s = qa % ON_NA;
            if (qa >= ON_NA) mbar_wait(emptyA + 8 * s, ((qa / ON_NA) - 1) & 1);
            mbar_expect_tx(fullA + 8 * s, OZ_TILE);
            tma_load_4d(abase0 + s * OZ_TILE, &tmS, fullA + 8 * s, 0, kc * OZ_T, a, (int32_t)rc);
            ++qa;
          }
        }
    }
  } else if (warp == 1) {
    {   // all 32 lanes: warp-uniform control flow, single-thread instructions elected inside their asm blocks
      const uint32_t idesc = on_idesc(n16);
      OzRing rga{0, 0}, rgb{0, 0};
      uint32_t segs = 0;
      for (int64_t rc = coh; rc < nch; rc += ncoh, ++segs) {
        if (segs > 0) mbar_wait(bar_acce, (segs - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        for (int kc = 0; kc < nkc; ++kc) {
          // A slices in DESCENDING order: step a multiplies B slices 6 - a .. 6 (level a + j - 6) and is the last use of B slice
          // 6 - a, so the B slots drain one by one during the chunk and the next chunk's B slices stream in behind them
          uint32_t sB[OZ_S];
#pragma unroll
          for (int j = 0; j < OZ_S; ++j) {
            sB[j] = rgb.slot;
            mbar_wait(fullB + 8 * rgb.slot, rgb.par);
            if (++rgb.slot == ON_NB) { rgb.slot = 0; rgb.par ^= 1u; }
          }
#pragma unroll
          for (int a = OZ_S - 1; a >= 0; --a) {
            const uint32_t sA = rga.slot;
            mbar_wait(fullA + 8 * sA, rga.par);
            if (++rga.slot == ON_NA) { rga.slot = 0; rga.par ^= 1u; }
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t a_lo = (abase0 + sA * OZ_TILE) >> 4;
            const uint32_t init = (kc == 0 && a == OZ_S - 1) ? 0u : 1u;   // the first step touches every level once
#pragma unroll
            for (int j = 0; j < OZ_S; ++j)
              if (j >= OZ_S - 1 - a) {
                const uint32_t acc = tmem + (uint32_t)((a + j - (OZ_S - 1)) * ON_TN);
                const uint32_t b_lo = (sbase + sB[j] * ON_BT) >> 4;
#pragma unroll
                for (int ks = 0; ks < OZ_CH / 32; ks++) umma_i8(acc, a_lo + 256u * ks, b_lo + 2u * ks, idesc, ks > 0 ? 1u : init);
              }
            umma_commit(emptyA + 8 * sA);
            umma_commit(emptyB + 8 * sB[OZ_S - 1 - a]);
          }
        }
        umma_commit(bar_accf);
      }
    }
  } else {
    const int lq = warp & 3;
    const int rr = lq * 32 + lane;
    const double poison = *nonfinite ? __longlong_as_double(0x7FF8000000000000LL) : 0.0;   // Inf / NaN in S: NaN out
    uint32_t segs = 0;
    for (int64_t rc = coh; rc < nch; rc += ncoh, ++segs) {
      mbar_wait(bar_accf, segs & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const int64_t row = rc * OZ_CH + rr;
      for (int cq = 0; cq * 32 < n16; ++cq) {
        double sum[32];
#pragma unroll
        for (int i = 0; i < 32; i++) sum[i] = 0.0;
#pragma unroll
        for (int lev = 0; lev < OZ_S; ++lev) {       // smallest level first
          uint32_t v[32];
          const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(lev * ON_TN + cq * 32);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
              : "r"(taddr)
              : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          const double w = __longlong_as_double((long long)(1023 + 8 * lev) << 52);   // 2^(8 lev)
#pragma unroll
          for (int i = 0; i < 32; i++) sum[i] = fma((double)(int32_t)v[i], w, sum[i]);
        }
        if (row < n) {
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const int col = j0 + cq * 32 + i;
            // 2^(f - 62) as an exact double (f is clamped far inside the exponent range by the split)
            if (col < nb) {
              double* o = Out + row + (int64_t)col * ldo;
              const double v = alpha * (sum[i] * __longlong_as_double((long long)(1023 + f[col] - 2 * OZ_SHIFT + 48) << 52)) + poison;
              *o = (beta != 0.0) ? v + beta * *o : v;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      mbar_arrive(bar_acce);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}

// G[g_row0 + r, g_col0 + c] = 2^(eA + eB - 110) * sum_L 256^L * (sum over the items of the tile of level L)
__global__ void __launch_bounds__(256)
    oz_reduce_kernel(const long long* __restrict__ part, const OzTile* __restrict__ tiles, const int* __restrict__ item_group,
                     int slot_mul, const int* __restrict__ eA, const int* __restrict__ eB0, const int* __restrict__ eB1,
                     double* __restrict__ G0, int ldg0, double* __restrict__ G1, int ldg1, int mirror, const int* __restrict__ nonfinite) {
  const OzTile t = tiles[blockIdx.x];
  const bool bad = *nonfinite != 0;
  double* __restrict__ G = t.b_sel ? G1 : G0;
  const int ldg = t.b_sel ? ldg1 : ldg0;
  const int* eB = t.b_sel ? eB1 : eB0;
  const int tot = t.a_cols * t.b_cols;
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < tot; idx += gridDim.y * blockDim.x) {
    const int r = idx % t.a_cols, c = idx / t.a_cols;
    if (t.diag && r > c) continue;
    long long lev[7] = {0, 0, 0, 0, 0, 0, 0};   // levels 6..12
    for (int k = t.first; k < t.last; k += (t.stride ? t.stride : 1)) {
      const int grp = item_group[k];
      const long long* p = part + ((int64_t)k * slot_mul + t.rank) * (4 * OZ_T * OZ_T) + r + (int64_t)c * OZ_T;
      const int nl = grp ? 3 : 4, l0 = grp ? 0 : 3;
      for (int l = 0; l < nl; l++) lev[l0 + l] += p[(int64_t)l * OZ_T * OZ_T];
    }
    double s = 0.0;
#pragma unroll
    for (int l = 0; l < 7; l++) s += ldexp((double)lev[l], 8 * l);   // smallest level first; 2^48 of the level-6 weight is in the scale
    const double val = bad ? __longlong_as_double(0x7FF8000000000000LL) : ldexp(s, eA[t.a_col0 + r] + eB[t.b_col0 + c] - 2 * OZ_SHIFT + 48);
    G[(t.g_row0 + r) + (int64_t)(t.g_col0 + c) * ldg] = val;
    if (mirror && (t.g_row0 + r) != (t.g_col0 + c)) G[(t.g_col0 + c) + (int64_t)(t.g_row0 + r) * ldg] = val;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn oz_encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return (EncodeTiledFn)p;
  }();
  return fn;
}
// 4-D map of a slice array: (byte in chunk row: 128, column: m, slice: 7, chunk: nch); box = one 128-column slice tile
int oz_make_map(CUtensorMap* tm, const int8_t* base, int m, int64_t nch) {
  EncodeTiledFn enc = oz_encode_tiled();
  if (!enc) return -100;
  const cuuint64_t dims[4] = {(cuuint64_t)OZ_CH, (cuuint64_t)m, (cuuint64_t)OZ_S, (cuuint64_t)nch};
  const cuuint64_t strides[3] = {(cuuint64_t)OZ_CH, (cuuint64_t)OZ_CH * m, (cuuint64_t)OZ_CH * m * OZ_S};
  const cuuint32_t box[4] = {(cuuint32_t)OZ_CH, (cuuint32_t)OZ_T, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -100;
}

// phase times of the int8 Gram: four events per call, resolved at the start of the next call (after the stream synchronisation that
// call needs anyway) or by lb2_ctx_oz_stats
void oz_stats_resolve(lb2_ctx* ctx) {
  if (!ctx->oz_ev_pending) return;
  if (cudaEventSynchronize(ctx->oz_ev[3]) == cudaSuccess) {
    for (int i = 0; i < 3; i++) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, ctx->oz_ev[i], ctx->oz_ev[i + 1]) == cudaSuccess) ctx->oz_ms[i] += t;
    }
    ctx->oz_calls++;
  }
  cudaGetLastError();
  ctx->oz_ev_pending = false;
}
void oz_stats_mark(lb2_ctx* ctx, int i) {
  if (!ctx->oz_ev[i] && cudaEventCreate(&ctx->oz_ev[i]) != cudaSuccess) { cudaGetLastError(); return; }
  cudaEventRecord(ctx->oz_ev[i], ctx->stream);
  if (i == 3) ctx->oz_ev_pending = true;
}

// non-finite flag: the last int of the slice buffer (stable while the buffer lives; cleared by every call that splits anew)
int* oz_flag(lb2_ctx* ctx) { return (int*)((char*)ctx->oz_buf + ctx->oz_bytes - 64); }

// exponents + slices of an n x m f64 block
constexpr int OZ_HINT_MAX = 4096;   // columns per operand role that can carry exponent hints

// exponents + slices of an n x m f64 block.  role >= 0 (operand of the solver's column-block Gram: 0 = [X P W], 1 = B W, 2 = A W)
// with use_hint: the exponents are the ones the PREVIOUS split of this role suggested (that split's column maxima + 1 guard bit),
// the column maxima are collected by the split kernel itself and oz_hint_check_kernel decides whether the split stands
// (ctx->oz_redo[role], read by the caller) — the separate pass over the block for its maxima (a third of the split's time) is gone.
int oz_split(lb2_ctx* ctx, int64_t n, int m, const double* X, int64_t ld, int8_t* slices, int* e, unsigned long long* mx,
             int role = -1, bool use_hint = false) {
  LB2_CUDA_OK(cudaMemsetAsync(mx, 0, sizeof(unsigned long long) * m, ctx->stream));
  const int64_t nch = (n + OZ_CH - 1) / OZ_CH;
  const int cpc = 16;
  const dim3 sgrid((unsigned)((nch + cpc - 1) / cpc), (m + 7) / 8);
  int* hint = (role >= 0 && ctx->oz_hint && m <= OZ_HINT_MAX) ? (int*)ctx->oz_hint + (size_t)role * OZ_HINT_MAX : nullptr;
  if (use_hint && hint) {
    int* redo = (int*)ctx->oz_hint + 3 * OZ_HINT_MAX + role;
    LB2_CUDA_OK(cudaMemcpyAsync(e, hint, sizeof(int) * m, cudaMemcpyDeviceToDevice, ctx->stream));
    LB2_CUDA_OK(cudaMemsetAsync(redo, 0, sizeof(int), ctx->stream));
    oz_split_kernel<true><<<sgrid, 256, 0, ctx->stream>>>(X, ld, n, m, e, slices, cpc, mx);
    oz_hint_check_kernel<<<(m + 127) / 128, 128, 0, ctx->stream>>>(mx, m, e, hint, redo, oz_flag(ctx));
    ctx->launches += 2;
    LB2_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const int64_t rows_per_cta = 65536;
  oz_absmax_kernel<<<dim3((unsigned)((n + rows_per_cta - 1) / rows_per_cta), m), 256, 0, ctx->stream>>>(X, ld, n, rows_per_cta, mx);
  oz_exp_kernel<<<(m + 127) / 128, 128, 0, ctx->stream>>>(mx, m, e, oz_flag(ctx));
  oz_split_kernel<false><<<sgrid, 256, 0, ctx->stream>>>(X, ld, n, m, e, slices, cpc, mx);
  ctx->launches += 3;
  if (hint) {
    oz_hint_store_kernel<<<(m + 127) / 128, 128, 0, ctx->stream>>>(e, m, hint);
    ctx->launches++;
  }
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

struct OzPlanTile { OzTile t; double w; };   // w: relative MMA cost of one chunk of the whole tile (both groups)

// tiles x {lo, hi} laid end to end weighted by their MMA count, cut into ncta equal pieces (chunk granularity)
void oz_schedule(std::vector<OzPlanTile>& tiles, int64_t nch, int ncta, double load_w, int lockstep, std::vector<OzItem>& items,
                 std::vector<int>& cta_first) {
  struct Unit { int tile, group; double cost; };
  std::vector<Unit> units;
  for (size_t i = 0; i < tiles.size(); i++) {
    // a unit costs what is larger: its MMAs (10 / 18 slice products, proportional to the UMMA N) or the slice tiles it
    // pulls through the ring (8 / 14 boxes of 16 KB whatever the panel width; load_w = relative cost of a box)
    const double nfrac = ((tiles[i].t.b_cols + 15) / 16 * 16) / 128.0;
    units.push_back({(int)i, 0, std::max(10.0 * nfrac, 8.0 * load_w)});
    units.push_back({(int)i, 1, std::max(18.0 * nfrac, 14.0 * load_w)});
  }
  // LOCK-STEP COHORTS (lockstep != 0): every unit gets its own CTAs — P_lo per lo unit, P_hi per hi unit, each covering an equal
  // share of the rows from its start.  All CTAs of a cohort (same group, same row share) run the same instruction stream over
  // the same rows at the same time, so a slice tile that several output tiles need is fetched from DRAM by whichever CTA gets
  // there first and comes out of L2 for the others (a CTA that runs ahead pays the DRAM latency and is caught up).  The equal-cost
  // cut below balances the work better but lets every CTA stream its own pieces from DRAM (ncu: 110 GB read for 26 GB of slices).
  if (lockstep) {
    const int nt = (int)tiles.size();
    int best_lo = 0, best_hi = 0;
    double best = 1e300;
    for (int plo = 1; plo <= 8; plo++)
      for (int phi = 1; phi <= 16; phi++) {
        if ((int64_t)nt * (plo + phi) > ncta) continue;
        const double tmax = std::max(10.0 / plo, 18.0 / phi);
        if (tmax < best - 1e-12) { best = tmax; best_lo = plo; best_hi = phi; }
      }
    if (best_lo > 0 && nch >= 8 * best_hi) {
      cta_first.assign(ncta + 1, 0);
      int cta = 0;
      for (int grp = 0; grp < 2; grp++) {          // cohorts are laid out group by group: neighbours in launch order share rows
        const int P = grp ? best_hi : best_lo;
        for (int pc = 0; pc < P; pc++)
          for (int i = 0; i < nt; i++) {
            OzTile& t = tiles[i].t;
            OzItem im{};
            im.chunk_begin = (int32_t)(nch * pc / P); im.chunk_end = (int32_t)(nch * (pc + 1) / P);
            im.a_col0 = t.a_col0; im.b_col0 = t.b_col0; im.b_sel = t.b_sel;
            im.n16 = (t.b_cols + 15) / 16 * 16;
            im.group = grp;
            im.tile = i;
            cta_first[cta++] = (int)items.size();
            items.push_back(im);
          }
      }
      for (int b = cta; b <= ncta; b++) cta_first[b] = (int)items.size();
      // the items of tile i are i, i + nt, i + 2 nt, ...
      for (int i = 0; i < nt; i++) { tiles[i].t.first = i; tiles[i].t.stride = nt; tiles[i].t.last = (int)items.size(); }
      return;
    }
  }
  double total = 0;
  for (auto& u : units) total += u.cost * (double)nch;
  const double L = total / ncta;
  cta_first.assign(ncta + 1, 0);
  std::vector<int> item_cta;
  double U = 0;
  for (auto& u : units) {
    OzTile& t = tiles[u.tile].t;
    if (u.group == 0) t.first = (int)items.size();
    const double span = u.cost * (double)nch;
    int b_lo = std::min(std::max((int)std::floor(U / L), 0), ncta - 1);
    int b_hi = std::min(std::max((int)std::floor((U + span) / L), 0), ncta - 1);
    auto boundary = [&](int b) -> int64_t {
      if (b <= b_lo) return 0;
      if (b > b_hi) return nch;
      int64_t r = (int64_t)std::llround(((double)b * L - U) / u.cost);
      if (r < 4) r = 0;
      if (nch - r < 4) r = nch;
      return std::min<int64_t>(std::max<int64_t>(r, 0), nch);
    };
    for (int b = b_lo; b <= b_hi; b++) {
      const int64_t c0 = boundary(b), c1 = boundary(b + 1);
      if (c1 <= c0) continue;
      OzItem im{};
      im.chunk_begin = (int32_t)c0; im.chunk_end = (int32_t)c1;
      im.a_col0 = t.a_col0; im.b_col0 = t.b_col0; im.b_sel = t.b_sel;
      im.n16 = (t.b_cols + 15) / 16 * 16;
      im.group = u.group;
      im.tile = u.tile;
      items.push_back(im);
      item_cta.push_back(b);
    }
    if (u.group == 1) t.last = (int)items.size();
    U += span;
  }
  size_t it = 0;
  for (int b = 0; b < ncta; b++) {
    while (it < items.size() && item_cta[it] < b) it++;
    cta_first[b] = (int)it;
  }
  cta_first[ncta] = (int)items.size();
}

int8_t* oz_buffer(lb2_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->oz_bytes) return (int8_t*)ctx->oz_buf;
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return nullptr;
  if (ctx->oz_buf) cudaFree(ctx->oz_buf);
  ctx->oz_buf = nullptr;
  ctx->oz_bytes = 0;
  ctx->oz_tag_ptr = nullptr;
  ctx->oz_buf = oz_malloc(ctx, bytes);
  if (!ctx->oz_buf) return nullptr;   // not enough device memory for the slices: the callers fall back to the DMMA kernels (-100)
  ctx->oz_bytes = bytes;
  return (int8_t*)ctx->oz_buf;
}

}  // namespace

namespace {

struct OzOperand {          // one f64 block and where its slices / exponents live
  const double* X; int64_t ld; int m;
  int8_t* slices; int* e;
  int role = -1;            // >= 0: operand role of the solver's column-block Gram (exponent hints carried from pass to pass)
};

// splits of the operands of a run; hinted splits are verified (one small device-to-host read) and redone in full when the hint was off
int oz_split_operands(lb2_ctx* ctx, int64_t n, OzOperand (&op)[3], int nop, unsigned long long* mx) {
  if (!ctx->oz_hint && ctx->oz_hints != 0) {
    if (cudaMalloc(&ctx->oz_hint, sizeof(int) * (3 * OZ_HINT_MAX + 4)) != cudaSuccess) { cudaGetLastError(); ctx->oz_hint = nullptr; }
  }
  bool hinted[3] = {false, false, false};
  for (int q = 0; q < nop; q++) {
    if (!op[q].X) continue;
    const int r = op[q].role;
    const bool ok = ctx->oz_hints != 0 && ctx->oz_hint && r >= 0 && r < 3 && op[q].m <= OZ_HINT_MAX && ctx->oz_hint_valid[r] &&
                    ctx->oz_hint_n[r] == n && ctx->oz_hint_m[r] == op[q].m;
    if (int rc = oz_split(ctx, n, op[q].m, op[q].X, op[q].ld, op[q].slices, op[q].e, mx, r, ok)) return rc;
    hinted[q] = ok;
    if (r >= 0 && r < 3) { ctx->oz_hint_valid[r] = ctx->oz_hint != nullptr; ctx->oz_hint_n[r] = n; ctx->oz_hint_m[r] = op[q].m; }
  }
  if (hinted[0] || hinted[1] || hinted[2]) {
    int redo[3] = {0, 0, 0};
    LB2_CUDA_OK(cudaMemcpyAsync(redo, (int*)ctx->oz_hint + 3 * OZ_HINT_MAX, sizeof(redo), cudaMemcpyDeviceToHost, ctx->stream));
    LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < nop; q++)
      if (hinted[q] && redo[op[q].role]) {   // a column grew past its guard bit or shrank by more than 4 bits: split again from fresh maxima
        ctx->oz_hint_redos++;
        if (int rc = oz_split(ctx, n, op[q].m, op[q].X, op[q].ld, op[q].slices, op[q].e, mx, op[q].role, false)) return rc;
      }
  }
  return 0;
}

// split the operands, run the tile list, reduce.  B panels come from operand 1 (b_sel = 0) or operand 2 (b_sel = 1); an
// operand with X == nullptr aliases the slices of operand 0 (its slices / e pointers are then set by the caller).
int oz_run(lb2_ctx* ctx, int64_t n, std::vector<OzPlanTile>& tiles, OzOperand (&op)[3], int nop, double* G0, int ldg0, double* G1,
           int ldg1, int mirror, int8_t* buf, size_t o_rest, unsigned long long* mx) {
  const int64_t nch = (n + OZ_CH - 1) / OZ_CH;
  auto al = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  std::vector<OzItem> items;
  std::vector<int> cta_first;
  const int ncta = (int)std::min<int64_t>(ctx->sm_count, std::max<int64_t>(1, (int64_t)tiles.size() * 2 * nch / 8));
  oz_schedule(tiles, nch, ncta, 0.01 * (ctx->oz_load_pct > 0 ? ctx->oz_load_pct : 100), ctx->oz_lockstep, items, cta_first);
  const size_t nitems = items.size();
  const size_t o_part = o_rest, o_items = o_part + al(sizeof(long long) * nitems * 4 * OZ_T * OZ_T),
               o_cta = o_items + al(sizeof(OzItem) * nitems), o_tiles = o_cta + al(sizeof(int) * (ncta + 1)),
               o_grp = o_tiles + al(sizeof(OzTile) * tiles.size()), total = o_grp + al(sizeof(int) * nitems);
  std::vector<int> grp;
  for (auto& im : items) grp.push_back(im.group);
  if (total > ctx->oz_bytes) return -3;   // the caller sized the buffer with oz_rest_bytes()
  long long* part = (long long*)(buf + o_part);
  std::vector<OzTile> tl;
  for (auto& p : tiles) tl.push_back(p.t);
  LB2_CUDA_OK(cudaMemcpyAsync(buf + o_items, items.data(), sizeof(OzItem) * nitems, cudaMemcpyHostToDevice, ctx->stream));
  LB2_CUDA_OK(cudaMemcpyAsync(buf + o_cta, cta_first.data(), sizeof(int) * (ncta + 1), cudaMemcpyHostToDevice, ctx->stream));
  LB2_CUDA_OK(cudaMemcpyAsync(buf + o_tiles, tl.data(), sizeof(OzTile) * tl.size(), cudaMemcpyHostToDevice, ctx->stream));
  LB2_CUDA_OK(cudaMemcpyAsync(buf + o_grp, grp.data(), sizeof(int) * nitems, cudaMemcpyHostToDevice, ctx->stream));
  LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));   // the host vectors go out of scope
  const bool timing = lb2_gram_i8_mode(ctx) == 2;   // debug: per-phase device times on stderr
  oz_stats_resolve(ctx);
  oz_stats_mark(ctx, 0);
  LB2_CUDA_OK(cudaMemsetAsync(oz_flag(ctx), 0, sizeof(int), ctx->stream));
  if (int rc = oz_split_operands(ctx, n, op, nop, mx)) return rc;
  alignas(64) CUtensorMap tm[3];
  for (int q = 0; q < 3; q++) {
    const OzOperand& o = op[q < nop ? q : 0];
    // an aliased operand (X == nullptr) is addressed through operand 0's map: its column offset is already in the tiles
    const OzOperand& src = o.X ? o : op[0];
    if (oz_make_map(&tm[q], src.slices, src.m, nch)) return -100;
  }
  oz_stats_mark(ctx, 1);
  LB2_CUDA_OK(cudaFuncSetAttribute(oz_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM));
  oz_gram_kernel<<<ncta, OZ_NT, OZ_SMEM, ctx->stream>>>(tm[0], tm[1], tm[2], (const OzItem*)(buf + o_items), (const int*)(buf + o_cta), part,
                                                        ctx->oz_prefetch > 0 ? ctx->oz_prefetch : 0);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  oz_stats_mark(ctx, 2);
  oz_reduce_kernel<<<dim3((unsigned)tl.size(), 8), 256, 0, ctx->stream>>>(part, (const OzTile*)(buf + o_tiles), (const int*)(buf + o_grp), 1,
                                                                          op[0].e, op[1].e, op[nop > 2 ? 2 : 1].e, G0, ldg0, G1, ldg1, mirror, oz_flag(ctx));
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  oz_stats_mark(ctx, 3);
  if (timing) {
    cudaEventSynchronize(ctx->oz_ev[3]);
    float t01 = 0, t12 = 0, t23 = 0;
    cudaEventElapsedTime(&t01, ctx->oz_ev[0], ctx->oz_ev[1]); cudaEventElapsedTime(&t12, ctx->oz_ev[1], ctx->oz_ev[2]);
    cudaEventElapsedTime(&t23, ctx->oz_ev[2], ctx->oz_ev[3]);
    fprintf(stderr, "gram_i8 n=%lld, %zu tiles: split %.2f ms, int8 MMA kernel %.2f ms (%zu items, %d CTAs), reduce %.2f ms\n",
            (long long)n, tiles.size(), t01, t12, nitems, ncta, t23);
  }
  return 0;
}
// 64-column boxes for the cluster kernel (each CTA fetches half of a shared slice tile)
int oz_make_map64(CUtensorMap* tm, const int8_t* base, int m, int64_t nch) {
  EncodeTiledFn enc = oz_encode_tiled();
  if (!enc) return -100;
  const cuuint64_t dims[4] = {(cuuint64_t)OZ_CH, (cuuint64_t)m, (cuuint64_t)OZ_S, (cuuint64_t)nch};
  const cuuint64_t strides[3] = {(cuuint64_t)OZ_CH, (cuuint64_t)OZ_CH * m, (cuuint64_t)OZ_CH * m * OZ_S};
  const cuuint32_t box[4] = {(cuuint32_t)OZ_CH, (cuuint32_t)(OZ_T / 2), 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -100;
}

struct OzPanel { int col0, cols, sel, valid; };     // a 128-column panel of an operand (sel: which B slice array)
struct OzSuper {                                    // 2 x 2 super-tile: A panels pa[0..1] x B panels pb[0..1]
  OzPanel pa[2], pb[2];
  int g_row0[2], g_col0[2];
  int valid[2][2];                                  // which of the four tiles are wanted
};

// number of 4-CTA clusters of the cluster kernel that can be resident at once (0: clusters not available)
int oz_max_clusters(lb2_ctx* ctx) {
  if (ctx->oz_clusters >= 0) return ctx->oz_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(4 * 64); cfg.blockDim = dim3(OZ_NT); cfg.dynamicSmemBytes = OZ_SMEM; cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int ncl = 0;
  if (cudaFuncSetAttribute(oz_gram_cluster_kernel<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM) != cudaSuccess ||
      cudaOccupancyMaxActiveClusters(&ncl, oz_gram_cluster_kernel<4, 8>, &cfg) != cudaSuccess)
    ncl = 0;
  cudaGetLastError();
  ctx->oz_clusters = ncl;
  return ncl;
}

// Output tiles (one-tile-per-CTA kernel) and 2 x 2 super-tiles (cluster kernel) of the column-block products G_q = S^H W_q:
// 128-column row panels over [0, tri_c0) and, separately, over the Hermitian block [tri_c0, m); tiles strictly below the diagonal
// of that block are left out.  Super-tiles: A panels paired in order, B panels paired as (W0 tile j, W1 tile j) (both products) or
// as neighbouring column tiles (one product); a missing partner repeats the panel and is not written out.  Host-only.
void oz_cols_tiles(int m, int nw, int nprod, int tri_c0, bool w0_in_s, bool want_supers, std::vector<OzPlanTile>& tiles,
                   std::vector<OzSuper>& supers) {
  const bool herm = tri_c0 >= 0 && tri_c0 + nw == m;
  const int split = herm ? tri_c0 : m;
  std::vector<std::pair<int, int>> rows;
  for (int c = 0; c < split; c += OZ_T) rows.emplace_back(c, std::min(OZ_T, split - c));
  for (int c = split; c < m; c += OZ_T) rows.emplace_back(c, std::min(OZ_T, m - c));
  const int ntn = (nw + OZ_T - 1) / OZ_T;
  for (auto& rw : rows)
    for (int q = 0; q < nprod; q++)
      for (int tj = 0; tj < ntn; tj++) {
        if (rw.first >= split && (rw.first - split) / OZ_T > tj) continue;   // below the diagonal of the Hermitian block
        OzPlanTile pt{};
        pt.t.a_col0 = rw.first; pt.t.a_cols = rw.second;
        pt.t.b_cols = std::min(OZ_T, nw - tj * OZ_T);
        pt.t.b_sel = q;
        pt.t.b_col0 = tj * OZ_T + ((q == 0 && w0_in_s) ? tri_c0 : 0);   // in the column numbering of the slice array it is read from
        pt.t.g_row0 = rw.first; pt.t.g_col0 = tj * OZ_T;
        pt.t.diag = 0;
        tiles.push_back(pt);
      }
  if (!want_supers || tiles.empty()) return;
  std::vector<OzPanel> pa, pb;
  std::vector<int> pb_gcol, pb_tj;
  for (auto& rw : rows) pa.push_back({rw.first, rw.second, 0, 1});
  if (pa.size() % 2) { OzPanel d = pa.back(); d.valid = 0; pa.push_back(d); }
  auto bpanel = [&](int q, int tj) { return OzPanel{tj * OZ_T + ((q == 0 && w0_in_s) ? tri_c0 : 0), std::min(OZ_T, nw - tj * OZ_T), q, 1}; };
  if (nprod == 2) {
    for (int tj = 0; tj < ntn; tj++)
      for (int q = 0; q < 2; q++) { pb.push_back(bpanel(q, tj)); pb_gcol.push_back(tj * OZ_T); pb_tj.push_back(tj); }
  } else {
    for (int tj = 0; tj < ntn; tj++) { pb.push_back(bpanel(0, tj)); pb_gcol.push_back(tj * OZ_T); pb_tj.push_back(tj); }
    if (pb.size() % 2) { OzPanel d = pb.back(); d.valid = 0; pb.push_back(d); pb_gcol.push_back(pb_gcol.back()); pb_tj.push_back(pb_tj.back()); }
  }
  for (size_t ia = 0; ia < pa.size(); ia += 2)
    for (size_t ib = 0; ib < pb.size(); ib += 2) {
      OzSuper st{};
      bool any = false;
      for (int x = 0; x < 2; x++) { st.pa[x] = pa[ia + x]; st.g_row0[x] = pa[ia + x].col0; }
      for (int y = 0; y < 2; y++) { st.pb[y] = pb[ib + y]; st.g_col0[y] = pb_gcol[ib + y]; }
      for (int x = 0; x < 2; x++)
        for (int y = 0; y < 2; y++) {
          const bool below = st.pa[x].col0 >= split && (st.pa[x].col0 - split) / OZ_T > pb_tj[ib + y];
          st.valid[x][y] = (st.pa[x].valid && st.pb[y].valid && !below) ? 1 : 0;
          any = any || st.valid[x][y];
        }
      if (any) supers.push_back(st);
    }
}

// Cluster schedule, host-only: (super-tile, level group) units laid end to end by cost and cut into equal pieces for ncl clusters
// (a cluster advances at the pace of its slowest CTA; each CTA pulls HALF boxes); out: items, first item of every cluster, and
// the output tiles (the wanted members of every super-tile; items of both groups of a super-tile are adjacent).
void oz_schedule_cluster(const std::vector<OzSuper>& supers, int64_t nch, int ncl_max, double load_w, std::vector<OzCItem>& items,
                         std::vector<int>& grp, std::vector<int>& cl_first, std::vector<OzTile>& tl, int* ncl_out) {
  struct Unit { int st, group; double cost; };
  std::vector<Unit> units;
  for (size_t i = 0; i < supers.size(); i++) {
    const int nmax = std::max(supers[i].pb[0].cols, supers[i].pb[1].cols);
    const double nfrac = ((nmax + 15) / 16 * 16) / 128.0;
    units.push_back({(int)i, 0, std::max(10.0 * nfrac, 4.0 * load_w)});
    units.push_back({(int)i, 1, std::max(18.0 * nfrac, 7.0 * load_w)});
  }
  const int ncl = (int)std::min<int64_t>(ncl_max, std::max<int64_t>(1, (int64_t)units.size() * nch / 8));
  *ncl_out = ncl;
  double total = 0;
  for (auto& u : units) total += u.cost * (double)nch;
  const double L = total / ncl;
  std::vector<int> item_cl;
  cl_first.assign(ncl + 1, 0);
  std::vector<std::pair<int, int>> unit_items(units.size());
  double U = 0;
  for (size_t ui = 0; ui < units.size(); ui++) {
    const Unit& u = units[ui];
    const OzSuper& st = supers[u.st];
    unit_items[ui].first = (int)items.size();
    const double span = u.cost * (double)nch;
    const int b_lo = std::min(std::max((int)std::floor(U / L), 0), ncl - 1), b_hi = std::min(std::max((int)std::floor((U + span) / L), 0), ncl - 1);
    auto boundary = [&](int b) -> int64_t {
      if (b <= b_lo) return 0;
      if (b > b_hi) return nch;
      int64_t r = (int64_t)std::llround(((double)b * L - U) / u.cost);
      if (r < 4) r = 0;
      if (nch - r < 4) r = nch;
      return std::min<int64_t>(std::max<int64_t>(r, 0), nch);
    };
    for (int b = b_lo; b <= b_hi; b++) {
      const int64_t c0 = boundary(b), c1 = boundary(b + 1);
      if (c1 <= c0) continue;
      OzCItem im{};
      im.chunk_begin = (int32_t)c0; im.chunk_end = (int32_t)c1; im.group = u.group;
      for (int h = 0; h < 2; h++) {
        im.a_col0[h] = st.pa[h].col0;
        im.b_col0[h] = st.pb[h].col0; im.b_sel[h] = st.pb[h].sel; im.n16[h] = (st.pb[h].cols + 15) / 16 * 16;
      }
      items.push_back(im);
      item_cl.push_back(b);
      grp.push_back(u.group);
    }
    unit_items[ui].second = (int)items.size();
    U += span;
  }
  size_t it = 0;
  for (int b = 0; b < ncl; b++) {
    while (it < items.size() && item_cl[it] < b) it++;
    cl_first[b] = (int)it;
  }
  cl_first[ncl] = (int)items.size();
  for (size_t i = 0; i < supers.size(); i++)
    for (int ia = 0; ia < 2; ia++)
      for (int ib = 0; ib < 2; ib++) {
        if (!supers[i].valid[ia][ib]) continue;
        OzTile t{};
        t.a_col0 = supers[i].pa[ia].col0; t.a_cols = supers[i].pa[ia].cols;
        t.b_col0 = supers[i].pb[ib].col0; t.b_cols = supers[i].pb[ib].cols; t.b_sel = supers[i].pb[ib].sel;
        t.g_row0 = supers[i].g_row0[ia]; t.g_col0 = supers[i].g_col0[ib];
        t.first = unit_items[2 * i].first; t.last = unit_items[2 * i + 1].second;
        t.rank = ia * 2 + ib;
        t.diag = 0;
        tl.push_back(t);
      }
}

int oz_run_cluster(lb2_ctx* ctx, int64_t n, std::vector<OzSuper>& supers, OzOperand (&op)[3], int nop, double* G0, int ldg0, double* G1,
                   int ldg1, int8_t* buf, size_t o_rest, unsigned long long* mx, int ncl_max) {
  const int64_t nch = (n + OZ_CH - 1) / OZ_CH;
  auto al = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  // r02 sweep at the C5 shape (one half-box = 8 KB per CTA and slice tile): 130 % 76.6 ms, 160 % 72.1, 200 % 63.1, 250 % 61.2, 300 % 63.7
  const double load_w = 0.01 * (ctx->oz_load_pct > 0 ? ctx->oz_load_pct : 250);
  std::vector<OzCItem> items;
  std::vector<int> grp, cl_first;
  std::vector<OzTile> tl;
  int ncl = 1;
  oz_schedule_cluster(supers, nch, ncl_max, load_w, items, grp, cl_first, tl, &ncl);
  const size_t nitems = items.size();
  const size_t o_part = o_rest, o_items = o_part + al(sizeof(long long) * nitems * 4 * 4 * OZ_T * OZ_T),
               o_cl = o_items + al(sizeof(OzCItem) * nitems), o_tiles = o_cl + al(sizeof(int) * (ncl + 1)),
               o_grp = o_tiles + al(sizeof(OzTile) * tl.size()), totalb = o_grp + al(sizeof(int) * nitems);
  if (totalb > ctx->oz_bytes) return -3;
  long long* part = (long long*)(buf + o_part);
  LB2_CUDA_OK(cudaMemcpyAsync(buf + o_items, items.data(), sizeof(OzCItem) * nitems, cudaMemcpyHostToDevice, ctx->stream));
  LB2_CUDA_OK(cudaMemcpyAsync(buf + o_cl, cl_first.data(), sizeof(int) * (ncl + 1), cudaMemcpyHostToDevice, ctx->stream));
  LB2_CUDA_OK(cudaMemcpyAsync(buf + o_tiles, tl.data(), sizeof(OzTile) * tl.size(), cudaMemcpyHostToDevice, ctx->stream));
  LB2_CUDA_OK(cudaMemcpyAsync(buf + o_grp, grp.data(), sizeof(int) * nitems, cudaMemcpyHostToDevice, ctx->stream));
  LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  const bool timing = lb2_gram_i8_mode(ctx) == 2;
  oz_stats_resolve(ctx);
  oz_stats_mark(ctx, 0);
  LB2_CUDA_OK(cudaMemsetAsync(oz_flag(ctx), 0, sizeof(int), ctx->stream));
  if (int rc = oz_split_operands(ctx, n, op, nop, mx)) return rc;
  alignas(64) CUtensorMap tm[3];
  for (int q = 0; q < 3; q++) {
    const OzOperand& o = op[q < nop ? q : 0];
    const OzOperand& src = o.X ? o : op[0];
    if (oz_make_map64(&tm[q], src.slices, src.m, nch)) return -100;
  }
  oz_stats_mark(ctx, 1);
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(4 * ncl); cfg.blockDim = dim3(OZ_NT); cfg.dynamicSmemBytes = OZ_SMEM; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
#define LB2_OZC(NA, NB)                                                                                                        \
  {                                                                                                                            \
    LB2_CUDA_OK(cudaFuncSetAttribute(oz_gram_cluster_kernel<NA, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM)); \
    if (cudaLaunchKernelEx(&cfg, oz_gram_cluster_kernel<NA, NB>, tm[0], tm[1], tm[2], (const OzCItem*)(buf + o_items),            \
                           (const int*)(buf + o_cl), part) != cudaSuccess) {                                                      \
      cudaGetLastError();                                                                                                       \
      ctx->oz_clusters = 0;   /* cluster launches do not work here (e.g. a restricted device): one tile per CTA from now on */    \
      return -101;                                                                                                              \
    }                                                                                                                           \
  }
    switch (ctx->oz_ring) {   // A-ring slots (tuning option); default 4 A + 8 B
      case 3: LB2_OZC(3, 9) break;
      case 5: LB2_OZC(5, 7) break;
      case 6: LB2_OZC(6, 6) break;
      default: LB2_OZC(4, 8) break;
    }
#undef LB2_OZC
  }
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  oz_stats_mark(ctx, 2);
  oz_reduce_kernel<<<dim3((unsigned)tl.size(), 8), 256, 0, ctx->stream>>>(part, (const OzTile*)(buf + o_tiles), (const int*)(buf + o_grp), 4,
                                                                          op[0].e, op[1].e, op[nop > 2 ? 2 : 1].e, G0, ldg0, G1, ldg1, 0, oz_flag(ctx));
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  oz_stats_mark(ctx, 3);
  if (timing) {
    cudaEventSynchronize(ctx->oz_ev[3]);
    float t01 = 0, t12 = 0, t23 = 0;
    cudaEventElapsedTime(&t01, ctx->oz_ev[0], ctx->oz_ev[1]); cudaEventElapsedTime(&t12, ctx->oz_ev[1], ctx->oz_ev[2]);
    cudaEventElapsedTime(&t23, ctx->oz_ev[2], ctx->oz_ev[3]);
    fprintf(stderr, "gram_i8 n=%lld, %zu super-tiles (%zu tiles): split %.2f ms, int8 cluster kernel %.2f ms (%zu items, %d clusters), reduce %.2f ms\n",
            (long long)n, supers.size(), tl.size(), t01, t12, nitems, ncl, t23);
  }
  return 0;
}

// 64-output-column boxes of the C' slices for the projection kernel
int oz_make_map_c(CUtensorMap* tm, const int8_t* base, int nb, int nkc) {
  EncodeTiledFn enc = oz_encode_tiled();
  if (!enc) return -100;
  const cuuint64_t dims[4] = {(cuuint64_t)OZ_CH, (cuuint64_t)nb, (cuuint64_t)OZ_S, (cuuint64_t)nkc};
  const cuuint64_t strides[3] = {(cuuint64_t)OZ_CH, (cuuint64_t)OZ_CH * nb, (cuuint64_t)OZ_CH * nb * OZ_S};
  const cuuint32_t box[4] = {(cuuint32_t)OZ_CH, (cuuint32_t)ON_TN, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -100;
}

constexpr size_t OZ_TAIL = (size_t)32 << 20;   // room behind every slice buffer for the projection's C' slices and exponents

// bytes behind the slices: partial sums (one 4-level slot per item, items <= 2 tiles + CTAs) and the schedule
size_t oz_rest_bytes(lb2_ctx* ctx, size_t ntiles) {
  const size_t nitems = 2 * ntiles + (size_t)ctx->sm_count + 8;
  return sizeof(long long) * nitems * 4 * OZ_T * OZ_T + (sizeof(OzCItem) + sizeof(OzTile) + 64) * nitems + 65536 + OZ_TAIL;
}

}  // namespace

// G (ma x mb) = A^H B through the int8 split; upper != 0: Hermitian product (A and B span the same columns), the lower triangle
// is mirrored.  -100 = not available.
int gram_i8_f64(lb2_ctx* ctx, int64_t n, int ma, int mb, const double* A, int64_t lda, const double* B, int64_t ldb, double* G,
                int ldg, int upper) {
  if (n >= ((int64_t)1 << 31) * OZ_CH || !oz_encode_tiled()) return -100;
  const bool same = (A == B && lda == ldb && ma == mb);
  const int64_t nch = (n + OZ_CH - 1) / OZ_CH;
  auto al = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  std::vector<OzPlanTile> tiles;
  const int ntm = (ma + OZ_T - 1) / OZ_T, ntn = (mb + OZ_T - 1) / OZ_T;
  for (int tj = 0; tj < ntn; tj++)
    for (int ti = 0; ti < (upper ? tj + 1 : ntm); ti++) {
      OzPlanTile pt{};
      pt.t.a_col0 = ti * OZ_T; pt.t.a_cols = std::min(OZ_T, ma - ti * OZ_T);
      pt.t.b_col0 = tj * OZ_T; pt.t.b_cols = std::min(OZ_T, mb - tj * OZ_T);
      pt.t.b_sel = 0;
      pt.t.g_row0 = pt.t.a_col0; pt.t.g_col0 = pt.t.b_col0;
      pt.t.diag = (upper && ti == tj) ? 1 : 0;
      tiles.push_back(pt);
    }
  const size_t bytesA = al((size_t)nch * OZ_S * ma * OZ_CH), bytesB = same ? 0 : al((size_t)nch * OZ_S * mb * OZ_CH);
  const size_t o_eA = bytesA + bytesB, o_eB = o_eA + al(sizeof(int) * ma), o_mx = o_eB + al(sizeof(int) * mb),
               o_rest = o_mx + al(sizeof(unsigned long long) * std::max(ma, mb));
  int8_t* buf = oz_buffer(ctx, o_rest + oz_rest_bytes(ctx, tiles.size()));
  if (!buf) return -100;
  // the slices of A stay behind for a projection from the same block (armed by the solver: ortho_drop's U -= V (V^H B U), SVQB's U T)
  ctx->oz_tag_ptr = A; ctx->oz_tag_n = n; ctx->oz_tag_m = ma; ctx->oz_tag_ld = lda; ctx->oz_tag_e_off = o_eA;
  OzOperand op[3] = {{A, lda, ma, buf, (int*)(buf + o_eA)},
                     {same ? nullptr : B, ldb, mb, same ? buf : buf + bytesA, same ? (int*)(buf + o_eA) : (int*)(buf + o_eB)},
                     {nullptr, 0, 0, nullptr, nullptr}};
  return oz_run(ctx, n, tiles, op, 2, G, ldg, G, ldg, upper ? 1 : 0, buf, o_rest, (unsigned long long*)(buf + o_mx));
}

// Column-block products of the cached-Gram pass through the int8 split (same contract as gram_wl_cols_f64): G0 = S^H W0 and, with
// W1, G1 = S^H W1.  The slices of S are made once and serve as the A panels of both products and — when W0 is the column block
// tri_c0.. of S itself (B = I) — as the B panels of the first; tiles strictly below the diagonal of the Hermitian block are skipped.
int gram_cols_i8_f64(lb2_ctx* ctx, int64_t n, int m, int nw, const double* S, int64_t lds, const double* W0, int64_t ldw0, double* G0,
                     int ldg0, const double* W1, int64_t ldw1, double* G1, int ldg1, int tri_c0) {
  if (n >= ((int64_t)1 << 31) * OZ_CH || !oz_encode_tiled()) return -100;
  const int64_t nch = (n + OZ_CH - 1) / OZ_CH;
  auto al = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  const int nprod = (W1 && G1) ? 2 : 1;
  const bool herm = tri_c0 >= 0 && tri_c0 + nw == m;
  const bool w0_in_s = herm && W0 == S + (int64_t)tri_c0 * lds && ldw0 == lds;
  std::vector<OzPlanTile> tiles;
  std::vector<OzSuper> supers;
  const int ncl_max = (ctx->oz_cluster != 0) ? oz_max_clusters(ctx) : 0;
  oz_cols_tiles(m, nw, nprod, tri_c0, w0_in_s, ncl_max > 0, tiles, supers);
  if (tiles.empty()) return 0;
  const size_t bS = al((size_t)nch * OZ_S * m * OZ_CH), bW0 = w0_in_s ? 0 : al((size_t)nch * OZ_S * nw * OZ_CH),
               bW1 = nprod == 2 ? al((size_t)nch * OZ_S * nw * OZ_CH) : 0;
  const size_t o_eS = bS + bW0 + bW1, o_e0 = o_eS + al(sizeof(int) * m), o_e1 = o_e0 + al(sizeof(int) * nw),
               o_mx = o_e1 + al(sizeof(int) * nw), o_rest = o_mx + al(sizeof(unsigned long long) * m);
  int8_t* buf = oz_buffer(ctx, o_rest + std::max(oz_rest_bytes(ctx, tiles.size()), 4 * oz_rest_bytes(ctx, supers.size())));
  if (!buf) return -100;
  int* eS = (int*)(buf + o_eS);
  // the slices of S stay valid for the projections of this pass (tall_nn_i8_f64): same block, not written in between
  ctx->oz_tag_ptr = S; ctx->oz_tag_n = n; ctx->oz_tag_m = m; ctx->oz_tag_ld = lds; ctx->oz_tag_e_off = o_eS;
  OzOperand op[3] = {{S, lds, m, buf, eS, 0},
                     {w0_in_s ? nullptr : W0, ldw0, nw, w0_in_s ? buf : buf + bS, w0_in_s ? eS : (int*)(buf + o_e0), 1},
                     {nprod == 2 ? W1 : nullptr, ldw1, nw, buf + bS + bW0, (int*)(buf + o_e1), 2}};
  // the reduce kernel indexes the exponents with the tile's b_col0, which already carries tri_c0 for an aliased W0
  if (!supers.empty()) {
    const int rc = oz_run_cluster(ctx, n, supers, op, nprod == 2 ? 3 : 2, G0, ldg0, G1, ldg1, buf, o_rest, (unsigned long long*)(buf + o_mx), ncl_max);
    if (rc != -101) return rc;   // -101: the cluster launch was refused — same product with the one-tile-per-CTA kernel
  }
  return oz_run(ctx, n, tiles, op, nprod == 2 ? 3 : 2, G0, ldg0, G1, ldg1, 0, buf, o_rest, (unsigned long long*)(buf + o_mx));
}

// Out (n x nb) = alpha S C + beta Out on the int8 tensor path.  Uses the slices of S left by the last gram_cols_i8_f64 on the
// same block (the solver's pass: Gram, then X = S Cx and P = S Cp); splits S itself otherwise.  -100 = not available.
int tall_nn_i8_f64(lb2_ctx* ctx, int64_t n, int kd, int nb, double alpha, const double* S, int64_t lds, const double* C, int ldc,
                   double beta, double* Out, int64_t ldo) {
  if (n >= ((int64_t)1 << 31) * OZ_CH || !oz_encode_tiled()) return -100;
  const int64_t nch = (n + OZ_CH - 1) / OZ_CH;
  auto al = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  const int nkc = (kd + OZ_CH - 1) / OZ_CH;
  const size_t c_bytes = al((size_t)nkc * OZ_S * nb * OZ_CH), need_tail = c_bytes + al(sizeof(int) * nb);
  if (need_tail + 1024 > OZ_TAIL) return -100;
  const bool cached = ctx->oz_reuse && ctx->oz_buf && ctx->oz_tag_ptr == S && ctx->oz_tag_n == n && ctx->oz_tag_m == kd && ctx->oz_tag_ld == lds;
  int8_t* buf;
  int* eS;
  if (cached) {
    buf = (int8_t*)ctx->oz_buf;
    eS = (int*)(buf + ctx->oz_tag_e_off);
  } else {
    const size_t bS = al((size_t)nch * OZ_S * kd * OZ_CH), o_eS = bS, o_mx = o_eS + al(sizeof(int) * kd),
                 tot = o_mx + al(sizeof(unsigned long long) * kd) + OZ_TAIL;
    buf = oz_buffer(ctx, tot);
    if (!buf) return -100;
    eS = (int*)(buf + o_eS);
    LB2_CUDA_OK(cudaMemsetAsync(oz_flag(ctx), 0, sizeof(int), ctx->stream));
    if (int rc = oz_split(ctx, n, kd, S, lds, buf, eS, (unsigned long long*)(buf + o_mx))) return rc;
    ctx->oz_tag_ptr = S; ctx->oz_tag_n = n; ctx->oz_tag_m = kd; ctx->oz_tag_ld = lds; ctx->oz_tag_e_off = o_eS;
  }
  int8_t* cs = buf + (ctx->oz_bytes - OZ_TAIL);
  int* f = (int*)(cs + c_bytes);
  oz_cprep_kernel<<<nb, 128, 0, ctx->stream>>>(C, ldc, kd, nb, eS, cs, f, nkc);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  alignas(64) CUtensorMap tmS, tmC;
  if (oz_make_map(&tmS, buf, kd, nch) || oz_make_map_c(&tmC, cs, nb, nkc)) return -100;
  const int njt = (nb + ON_TN - 1) / ON_TN;
  const int ncoh = (int)std::max<int64_t>(1, std::min<int64_t>(nch, ctx->sm_count / njt));
#define LB2_OZN(NA, NB)                                                                                                   \
  {                                                                                                                       \
    LB2_CUDA_OK(cudaFuncSetAttribute(oz_nn_kernel<NA, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ON_SMEM));   \
    oz_nn_kernel<NA, NB><<<ncoh * njt, OZ_NT, ON_SMEM, ctx->stream>>>(tmS, tmC, f, Out, ldo, n, nb, nkc, nch, njt, ncoh,   \
                                                                      oz_flag(ctx), ctx->oz_prefetch > 0 ? ctx->oz_prefetch : 0, \
                                                                      alpha, beta);                                        \
  }
  switch (ctx->oz_nn_ring) {   // A-ring slots of the projection kernel (tuning option); default 7 A + 10 B
    case 5: LB2_OZN(5, 14) break;
    case 6: LB2_OZN(6, 12) break;
    case 8: LB2_OZN(8, 8) break;
    default: LB2_OZN(7, 10) break;
  }
#undef LB2_OZN
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// Host-only self-check of the int8 schedules of the column-block products (no device needed; tests/test_oz_plan.py).
// mode 0: one tile per CTA, equal-cost cut; 1: lock-step cohorts; 2: 4-CTA clusters (ncta = resident clusters).  Verifies that every
// output entry that is not in a tile strictly below the diagonal of the Hermitian block is produced by exactly one tile, that the
// items of every (tile or super-tile, level group) partition the row chunks, that a worker's items are contiguous, and (lock-step)
// that no CTA has more than one item.  stats[4]: items, busiest worker / mean (cost model), tiles written, workers used.
int oz_plan_check(int m, int nw, int nprod, int tri_c0, int64_t n, int nworkers, int mode, double* stats) {
  if (m <= 0 || nw <= 0 || nprod < 1 || nprod > 2 || n <= 0 || nworkers < 1) return 1;
  const int64_t nch = (n + OZ_CH - 1) / OZ_CH;
  std::vector<OzPlanTile> tiles;
  std::vector<OzSuper> supers;
  oz_cols_tiles(m, nw, nprod, tri_c0, false, mode == 2, tiles, supers);
  if (tiles.empty()) return 2;
  std::vector<OzTile> tl;
  std::vector<int> grp, first;
  std::vector<std::pair<int64_t, int64_t>> ranges;   // chunk range of every item
  std::vector<double> wcost;
  int nused = 0;
  if (mode == 2) {
    if (supers.empty()) return 3;
    std::vector<OzCItem> items;
    int ncl = 1;
    oz_schedule_cluster(supers, nch, nworkers, 2.5, items, grp, first, tl, &ncl);
    nused = ncl;
    wcost.assign(ncl, 0.0);
    for (auto& im : items) ranges.emplace_back(im.chunk_begin, im.chunk_end);
    for (int b = 0; b < ncl; b++)
      for (int i = first[b]; i < first[b + 1]; i++) {   // the scheduler's own cost model (oz_schedule_cluster)
        const double nfrac = std::max(items[i].n16[0], items[i].n16[1]) / 128.0;
        const double c = grp[i] ? std::max(18.0 * nfrac, 7.0 * 2.5) : std::max(10.0 * nfrac, 4.0 * 2.5);
        wcost[b] += c * (double)(ranges[i].second - ranges[i].first);
      }
    if (first[ncl] != (int)items.size()) return 4;
    // a super-tile's items: both groups adjacent, each group a partition of [0, nch)
    for (auto& t : tl) {
      for (int g = 0; g < 2; g++) {
        int64_t r = 0;
        for (int i = t.first; i < t.last; i++)
          if (grp[i] == g) { if (ranges[i].first != r) return 5; r = ranges[i].second; }
        if (r != nch) return 6;
      }
      if (t.rank < 0 || t.rank > 3) return 7;
    }
  } else {
    std::vector<OzItem> items;
    const int ncta = (int)std::min<int64_t>(nworkers, std::max<int64_t>(1, (int64_t)tiles.size() * 2 * nch / 8));
    oz_schedule(tiles, nch, ncta, 1.0, mode == 1, items, first);
    nused = ncta;
    wcost.assign(ncta, 0.0);
    for (auto& im : items) { ranges.emplace_back(im.chunk_begin, im.chunk_end); grp.push_back(im.group); }
    if (first[ncta] != (int)items.size()) return 4;
    for (int b = 0; b < ncta; b++) {
      if (mode == 1 && first[b + 1] - first[b] > 1 && (int64_t)tiles.size() * 3 <= ncta) return 8;
      for (int i = first[b]; i < first[b + 1]; i++) {   // the scheduler's own cost model (oz_schedule, load_w = 1)
        const double nfrac = items[i].n16 / 128.0;
        const double c = grp[i] ? std::max(18.0 * nfrac, 14.0) : std::max(10.0 * nfrac, 8.0);
        wcost[b] += c * (double)(ranges[i].second - ranges[i].first);
      }
    }
    for (size_t ti = 0; ti < tiles.size(); ti++) {
      const OzTile& t = tiles[ti].t;
      const int st = t.stride ? t.stride : 1;
      for (int g = 0; g < 2; g++) {
        std::vector<std::pair<int64_t, int64_t>> rs;
        for (int i = t.first; i < t.last; i += st)
          if (items[i].tile == (int)ti && grp[i] == g) rs.push_back(ranges[i]);
          else if (items[i].tile != (int)ti) return 9;
        std::sort(rs.begin(), rs.end());
        int64_t r = 0;
        for (auto& x : rs) { if (x.first != r) return 5; r = x.second; }
        if (r != nch) return 6;
      }
      tl.push_back(t);
    }
  }
  // coverage of the outputs
  const bool herm = tri_c0 >= 0 && tri_c0 + nw == m;
  std::vector<int> cover((size_t)nprod * m * nw, 0);
  for (auto& t : tl) {
    if (t.a_col0 < 0 || t.a_col0 + t.a_cols > m || t.g_col0 < 0 || t.g_col0 + t.b_cols > nw || t.b_sel < 0 || t.b_sel >= nprod) return 10;
    for (int i = 0; i < t.a_cols; i++)
      for (int j = 0; j < t.b_cols; j++) cover[((size_t)t.b_sel * m + t.g_row0 + i) * nw + t.g_col0 + j]++;
  }
  for (int q = 0; q < nprod; q++)
    for (int i = 0; i < m; i++)
      for (int j = 0; j < nw; j++) {
        const int c = cover[((size_t)q * m + i) * nw + j];
        const bool below_tile = herm && i >= tri_c0 && (i - tri_c0) / OZ_T > j / OZ_T;
        if (c > 1 || (c == 0 && !below_tile)) return 11;
      }
  double mx = 0, sum = 0;
  for (double c : wcost) { mx = std::max(mx, c); sum += c; }
  if (stats) {
    stats[0] = (double)ranges.size();
    stats[1] = sum > 0 ? mx / (sum / nused) : 1.0;
    stats[2] = (double)tl.size();
    stats[3] = (double)nused;
  }
  return 0;
}

// are the slices of this block the ones the last Gram left behind (and has the solver vouched that the block is unchanged)?
bool oz_slices_cached(const lb2_ctx* ctx, const void* S, int64_t n, int kd, int64_t lds) {
  return ctx->oz_reuse && ctx->oz_buf && ctx->oz_tag_ptr == S && ctx->oz_tag_n == n && ctx->oz_tag_m == kd && ctx->oz_tag_ld == lds;
}

// accumulated device times (ms) of the int8 Gram phases since the last query: out = {split, MMA kernel, reduce, calls}
int oz_stats_query(lb2_ctx* ctx, double* out) {
  cudaStreamSynchronize(ctx->stream);
  oz_stats_resolve(ctx);
  for (int i = 0; i < 3; i++) { out[i] = ctx->oz_ms[i]; ctx->oz_ms[i] = 0.0; }
  out[3] = (double)ctx->oz_calls;
  ctx->oz_calls = 0;
  return 0;
}

}  // namespace lb2

// lobpcg_b200/csrc/gram_tc5.cu — K2/K3 for float on the 5th-generation tensor cores: G = A^H B through
// tcgen05.mma kind::tf32 with the accumulator in TMEM (SASS: UTCMMA / tensor-memory loads).
//
// The algorithm needs fp32-accurate Grams (EPS_TOL = 1e-5 for s/c, reference src/core/lobpcg_s.c:10; SURVEY §7 hard
// part 2), plain TF32 (10-bit mantissa, and the tensor core TRUNCATES fp32 inputs) is not enough.  So every operand
// element x is split into hi = trunc_tf32(x) and lo = rna_tf32(x - hi) and a product is three MMAs, lo*hi + hi*lo + hi*hi,
// accumulated in fp32 in TMEM ("3xTF32").  The first version of this path (dense.cu: gram_tf32_kernel) does the same
// with mma.sync and splits fragments in registers on every use; here the split happens ONCE per staged element.
//
// Per CTA: one 128 x 128 output tile and one contiguous row range (same tile x split grid and deterministic split
// reduction as gram_dmma_kernel).  Per K chunk of 32 rows:
//   1. cp.async (16-byte, zero-filling) lands the raw fp32 panels A[:, 128 cols] and B[:, 128 cols] directly in the
//      canonical K-major no-swizzle UMMA layout: 8 x 16-byte core matrices, 128 B each, k-units 128 B apart, column
//      groups 1 KB apart.  A warp instruction covers 8 columns x 4 k-units: 64 contiguous bytes per column in global
//      memory (full sectors) and all 8 sixteen-byte bank groups four times in shared memory (no excess wavefronts).
//   2. every thread converts the units it copied itself (no barrier needed for that): the raw tile is the hi operand
//      (the tensor core reads the upper 19 bits = exact truncation), lo = rna_tf32(x - trunc(x)) goes to a second
//      tile; fence.proxy.async makes the generic-proxy writes visible to the tensor core.
//   3. one thread issues 4 k-steps x 3 tcgen05.mma (M = N = 128, K = 8) and a tcgen05.commit on the stage's mbarrier;
//      the MMAs of chunk c run while the CTA converts chunk c + 1.
// Epilogue: the fp64 register accumulators -> partial tile in global scratch (deterministic split reduction follows).
//
// Measured (B200, n = 4.096 M): 74 TFLOP/s (3xTF32 counted once) at 896 x 896, 50 at m = 600 upper, against 55 / 37 for
// the mma.sync kernel; the same rate at n = 65 536 (TLB- and L2-friendly), so HBM / L2 / TLB are not the limit.
// Probe with parts of the chunk loop switched off (n = 2.1 M, 896 x 896, 44.9 ms in full): no copies 26.3 ms, no split
// 39.1, no proxy fence 43.6, no MMAs 39.5, empty skeleton (barrier per chunk, mbarrier waits, commit, fp64 drain) 10.4 —
// i.e. per 32-row chunk 0.47 us skeleton + 0.85 us issuing/landing the 64 warp-level cp.async + 0.27 us split + 0.25 us
// of MMA time that is not hidden + 0.06 us fence = 2.06 us, where the 12 MMAs alone need 0.73 us (119 clocks per
// 128x128x8 MMA with no-swizzle operands): every warp does every phase, so the phases add up instead of overlapping.
// A first warp-specialised variant (ninth warp issuing the MMAs, full/done mbarriers, producers still copying, splitting
// and draining in sequence) was slower (112 ms vs 95 ms).  Next step: TMA (SWIZZLE_128B tiles, one instruction per
// panel and chunk instead of 64 warp-level copies) feeding separate split / MMA / drain warps.
#include <cstdint>
#include <cuda.h>   // CUtensorMap + enums only; cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint

#include "common.cuh"
#include "context.h"
#include "kernels.h"

namespace lb2 {

namespace {

constexpr int TC_T = 128;            // tile edge = UMMA M = UMMA N
constexpr int TC_BK = 32;            // rows (K) per chunk
constexpr int TC_RAW = 4;            // raw (= hi operand) stages: chunk c + 2 is requested while chunk c is converted
constexpr int TC_LO = 2;             // lo-operand buffers (written by the CTA, read by the MMAs of one chunk)
constexpr int TC_AHEAD = TC_RAW - 2; // prefetch distance in chunks
constexpr int TC_FLUSH = 2;          // chunks per TMEM accumulation group (see "accuracy" below)
constexpr int TC_NT = 256;           // 8 warps: warp w owns TMEM lanes 32 (w % 4) .. and column half w / 4 when draining
constexpr int TC_UNITS = (TC_T * TC_BK / 4) / TC_NT;   // 16-byte units per thread, operand tile and chunk
constexpr uint32_t TC_LBO = 144;     // bytes between consecutive 16-byte k-units (core matrices along K): 128 + 16 so that the
                                     // eight k-units of a column fall into eight different 16-byte bank groups
constexpr uint32_t TC_SBO = (TC_BK / 4) * TC_LBO;          // bytes between 8-column groups (core matrices along M/N)
constexpr uint32_t TC_TILE = (TC_T / 8) * TC_SBO;          // 16 KB per operand tile
constexpr uint32_t TC_STAGE = 2 * TC_TILE;                 // raw stage: A, B;  lo buffer: A_lo, B_lo
constexpr uint32_t TC_BAR = (TC_RAW + TC_LO) * TC_STAGE;   // offset of the mbarriers / TMEM pointer
constexpr uint32_t TC_SMEM = TC_BAR + 2048;                // + barriers and 1 KB alignment slack (226 KB of 227)

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

// K-major, no-swizzle shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((TC_LBO >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((TC_SBO >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128 (InstrDescriptor bit fields)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_T >> 3) << 17) | ((uint32_t)(TC_T >> 4) << 24);

// Single-thread instructions are issued from WARP-UNIFORM code and predicated on elect.sync: inside an `if (tid == 0)` region the
// operands look thread-private to ptxas, which then wraps every UTCHMMA / UTMALDG / UTCBAR in an ELECT + R2UR.BROADCAST waterfall
// loop (~130 clocks per MMA; found on the int8 kernels, gram_i8.cu).
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(bar)
      : "memory");
}

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return r;
}

__global__ void __launch_bounds__(TC_NT, 1)
    gram_tc5_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, int ma, int mb,
                    int64_t n, int64_t rows_per_split, int upper, int ntm, float* __restrict__ out,
                    int64_t split_stride, int ldo) {
  extern __shared__ __align__(1024) unsigned char smem_tc[];
  const uint32_t sbase = (smem_u32(smem_tc) + 1023u) & ~1023u;        // tiles 1 KB aligned
  unsigned char* gbase = smem_tc + (sbase - smem_u32(smem_tc));
  const uint32_t lo_base = sbase + TC_RAW * TC_STAGE;                  // lo buffers behind the raw stages
  const uint32_t bar0 = sbase + TC_BAR;                                // TC_RAW mbarriers, then the TMEM pointer
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + TC_BAR + 64);

  int ti, tj;
  if (upper) {   // t -> (i <= j)
    int t = blockIdx.x;
    tj = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
    while ((tj + 1) * (tj + 2) / 2 <= t) ++tj;
    while (tj * (tj + 1) / 2 > t) --tj;
    ti = t - tj * (tj + 1) / 2;
  } else {
    ti = blockIdx.x % ntm;
    tj = blockIdx.x / ntm;
  }
  const int m0 = ti * TC_T, c0 = tj * TC_T;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(n, r_begin + rows_per_split);
  const int nchunks = (r_end > r_begin) ? (int)((r_end - r_begin + TC_BK - 1) / TC_BK) : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp index the compiler knows to be uniform

  if (tid == 0) {
    for (int s = 0; s < TC_RAW; s++) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {   // two accumulators of 128 columns (ping-pong between accumulation groups)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32((const void*)tmem_slot)),
                 "n"(2 * TC_T)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // this thread's copy/convert units: TC_UNITS per operand tile; unit = (column c, k-unit u) -> 16 bytes = 4 rows of a column
  // a warp instruction covers 4 columns x all 8 k-units of the chunk: one full 128-byte line per column
  const int uu = lane & 7, cc = lane >> 3;
  uint32_t uoff[TC_UNITS];                   // byte offset of the unit inside an operand tile
  int ucol[TC_UNITS], urow[TC_UNITS];        // column inside the panel, first row inside the chunk
#pragma unroll
  for (int i = 0; i < TC_UNITS; i++) {
    const int blk = i * (TC_NT / 32) + warp;          // 0..31: block of 4 columns
    const int col = blk * 4 + cc;
    ucol[i] = col;
    urow[i] = uu * 4;
    uoff[i] = (uint32_t)(col >> 3) * TC_SBO + (uint32_t)uu * TC_LBO + (uint32_t)(col & 7) * 16u;
  }

  auto issue_loads = [&](int chunk) {
    if (chunk < nchunks) {
      const uint32_t st = sbase + (uint32_t)(chunk % TC_RAW) * TC_STAGE;
      const int64_t r0 = r_begin + (int64_t)chunk * TC_BK;
#pragma unroll
      for (int i = 0; i < TC_UNITS; i++) {
        const int64_t row = r0 + urow[i];
        const int64_t left = r_end - row;
        const int rb = left >= 4 ? 16 : (left > 0 ? (int)left * 4 : 0);
        {
          const bool ok = (m0 + ucol[i] < ma) && rb > 0;
          const float* src = ok ? A + (int64_t)(m0 + ucol[i]) * lda + row : A;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(st + uoff[i]), "l"(src), "r"(ok ? rb : 0));
        }
        {
          const bool ok = (c0 + ucol[i] < mb) && rb > 0;
          const float* src = ok ? B + (int64_t)(c0 + ucol[i]) * ldb + row : B;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(st + TC_TILE + uoff[i]), "l"(src),
                       "r"(ok ? rb : 0));
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };

  // Accuracy: the tensor core TRUNCATES when it adds into its fp32 accumulator, a relative bias of ~2^-25 per MMA that
  // grows linearly with the row count (measured -4.8e-4 on the Gram diagonal at n = 4 M with one long accumulation,
  // -2.8e-4 for mma.sync).  So TMEM only ever accumulates TC_FLUSH chunks (24 MMAs); finished groups are drained into
  // fp64 registers (64 entries per thread) while the MMAs of the next group run on the other TMEM accumulator.
  const int lq = warp & 3, ch = warp >> 2;      // TMEM lane quarter of this warp, column half
  double accd[TC_T / 2];
#pragma unroll
  for (int q = 0; q < TC_T / 2; q++) accd[q] = 0.0;
  auto drain = [&](int group) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
    for (int h = 0; h < 2; h++) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)((group & 1) * TC_T + ch * (TC_T / 2) + h * 32);
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
#pragma unroll
      for (int q = 0; q < 32; q++) accd[h * 32 + q] += (double)__uint_as_float(v[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  };
  auto done_bar = [&](int chunk) { return bar0 + 8u * (uint32_t)(chunk % TC_RAW); };
  auto done_par = [&](int chunk) { return (uint32_t)((chunk / TC_RAW) & 1); };

#pragma unroll
  for (int c = 0; c < TC_AHEAD; c++) issue_loads(c);

  int drained = 0;   // groups [0, drained) are in accd
  for (int chunk = 0; chunk < nchunks; chunk++) {
    const unsigned char* raw = gbase + (size_t)(chunk % TC_RAW) * TC_STAGE;
    unsigned char* lob = gbase + (size_t)TC_RAW * TC_STAGE + (size_t)(chunk & 1) * TC_STAGE;
    asm volatile("cp.async.wait_group %0;\n" ::"n"(TC_AHEAD - 1) : "memory");   // my copies of this chunk have landed
    if (chunk >= 2) {
      // MMAs of chunk - 2 are complete: its lo buffer (this chunk's) and its raw stage (chunk + TC_AHEAD's) are free
      mbar_wait(done_bar(chunk - 2), done_par(chunk - 2));
      const int gdone = (chunk - 1) / TC_FLUSH;        // groups whose last chunk is <= chunk - 2
      if (drained < gdone) { drain(drained); drained++; }
    }
    issue_loads(chunk + TC_AHEAD);
    // lo part of my own units.  The raw fp32 tile itself is the "hi" operand: the tensor core reads the upper 19 bits,
    // i.e. hi = trunc_tf32(x) exactly; lo = rna_tf32(x - hi) is exact in fp32 before rounding, so x = hi + lo to
    // 2^-22 |x| and the dropped lo*lo term is < 2^-20 of the product.
    float4 rv[2 * TC_UNITS];
#pragma unroll
    for (int i = 0; i < TC_UNITS; i++) {
      rv[2 * i] = *reinterpret_cast<const float4*>(raw + uoff[i]);
      rv[2 * i + 1] = *reinterpret_cast<const float4*>(raw + TC_TILE + uoff[i]);
    }
#pragma unroll
    for (int i = 0; i < TC_UNITS; i++) {
#pragma unroll
      for (int op = 0; op < 2; op++) {
        const float4 v = rv[2 * i + op];
        uint4 l;
        l.x = tf32_rna(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
        l.y = tf32_rna(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
        l.z = tf32_rna(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
        l.w = tf32_rna(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
        *reinterpret_cast<uint4*>(lob + (size_t)op * TC_TILE + uoff[i]) = l;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> tensor core (async proxy)
    __syncthreads();
    if (warp_u == 0) {   // all 32 lanes of warp 0; the single-thread instructions are elected inside their asm blocks
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t hi = sbase + (uint32_t)(chunk % TC_RAW) * TC_STAGE;
      const uint32_t lo = lo_base + (uint32_t)(chunk & 1) * TC_STAGE;
      const int group = chunk / TC_FLUSH;
      const uint32_t acc = tmem + (uint32_t)((group & 1) * TC_T);
#pragma unroll
      for (int ks = 0; ks < TC_BK / 8; ks++) {
        const uint32_t ko = (uint32_t)ks * 2u * TC_LBO;       // one MMA consumes K = 8 floats = 2 k-units
        const uint64_t ah = umma_desc(hi + ko), al = umma_desc(lo + ko);
        const uint64_t bh = umma_desc(hi + TC_TILE + ko), bl = umma_desc(lo + TC_TILE + ko);
        umma_tf32(acc, al, bh, (chunk % TC_FLUSH != 0 || ks > 0) ? 1u : 0u);
        umma_tf32(acc, ah, bl, 1u);
        umma_tf32(acc, ah, bh, 1u);
      }
      umma_commit(done_bar(chunk));   // arrives when every MMA issued so far is complete
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");

  if (nchunks > 0) {
    if (nchunks >= 2) mbar_wait(done_bar(nchunks - 2), done_par(nchunks - 2));
    mbar_wait(done_bar(nchunks - 1), done_par(nchunks - 1));
    const int ngroups = (nchunks + TC_FLUSH - 1) / TC_FLUSH;
    for (; drained < ngroups; drained++) drain(drained);
  }
  float* o = out + (int64_t)blockIdx.y * split_stride;
  const int row = m0 + lq * 32 + lane;          // G row = column of the A panel = TMEM lane
  if (row < ma) {
#pragma unroll
    for (int q = 0; q < TC_T / 2; q++) {
      const int col = c0 + ch * (TC_T / 2) + q;
      if (col < mb) o[row + (int64_t)col * ldo] = (float)accd[q];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(2 * TC_T) : "memory");
}

// ------------------------------------------------------------------------------------------------- TMA-fed variant
// Same tile / split grid, same 3xTF32 split, same short TMEM accumulation groups drained into fp64 — but the operand panels
// are brought in by the TMA unit: ONE elected thread issues two cp.async.bulk.tensor.2d per K chunk (a 32-row x 128-column
// box of A and of B, SASS UTMALDG) instead of 64 warp-level cp.async per chunk and operand, the boxes land in the canonical
// K-major SWIZZLE_128B layout (a column of S = one 128-byte row of the tile, 16-byte units XOR-swizzled by the row index),
// completion is signalled on an mbarrier (complete_tx), out-of-range rows / columns are zero-filled by the unit.  The
// profile of the cp.async version (header of this file) charged 0.85 us of the 2.06 us per chunk to issuing and landing
// those copies.  The lo tile is computed element-wise from the raw tile, so it inherits the swizzled layout.
constexpr uint32_t T2_TILE = TC_T * TC_BK * 4;            // 16 KB: 128 rows (columns of S) x 128 B (32 k-values)
constexpr uint32_t T2_STAGE = 2 * T2_TILE;                // A, B
constexpr uint32_t T2_BAR = (TC_RAW + TC_LO) * T2_STAGE;  // done[TC_RAW] | full[TC_RAW] | TMEM pointer
constexpr uint32_t T2_SMEM = T2_BAR + 2048;
constexpr int T2_UNITS = (int)(T2_TILE / 16) / TC_NT;     // 16-byte units per thread and operand tile

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  // K-major, SWIZZLE_128B: stride between 8-row groups 1024 B, leading-dimension offset unused (K extent = swizzle span)
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((1024u >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}\n" ::"r"(bar),
      "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}\n" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__global__ void __launch_bounds__(TC_NT, 1)
    gram_tc5_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int ma, int mb,
                        int64_t n, int64_t rows_per_split, int upper, int ntm, float* __restrict__ out,
                        int64_t split_stride, int ldo) {
  extern __shared__ __align__(1024) unsigned char smem_tc[];
  const uint32_t sbase = (smem_u32(smem_tc) + 1023u) & ~1023u;
  unsigned char* gbase = smem_tc + (sbase - smem_u32(smem_tc));
  const uint32_t lo_base = sbase + TC_RAW * T2_STAGE;
  const uint32_t bar_done = sbase + T2_BAR, bar_full = sbase + T2_BAR + 8 * TC_RAW;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + T2_BAR + 128);

  int ti, tj;
  if (upper) {
    int t = blockIdx.x;
    tj = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
    while ((tj + 1) * (tj + 2) / 2 <= t) ++tj;
    while (tj * (tj + 1) / 2 > t) --tj;
    ti = t - tj * (tj + 1) / 2;
  } else {
    ti = blockIdx.x % ntm;
    tj = blockIdx.x / ntm;
  }
  const int m0 = ti * TC_T, c0 = tj * TC_T;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(n, r_begin + rows_per_split);
  const int nchunks = (r_end > r_begin) ? (int)((r_end - r_begin + TC_BK - 1) / TC_BK) : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp index the compiler knows to be uniform

  if (tid == 0) {
    for (int s = 0; s < TC_RAW; s++) { mbar_init(bar_done + 8 * s, 1); mbar_init(bar_full + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32((const void*)tmem_slot)),
                 "n"(2 * TC_T)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  auto issue_tma = [&](int chunk) {   // warp 0, warp-uniform
    if (chunk < nchunks) {
      const uint32_t s = (uint32_t)(chunk % TC_RAW);
      const uint32_t st = sbase + s * T2_STAGE;
      const int32_t row = (int32_t)(r_begin + (int64_t)chunk * TC_BK);
      mbar_expect_tx(bar_full + 8 * s, T2_STAGE);
      tma_load_2d(st, &tmA, bar_full + 8 * s, row, m0);
      tma_load_2d(st + T2_TILE, &tmB, bar_full + 8 * s, row, c0);
    }
  };

  const int lq = warp & 3, ch = warp >> 2;
  double accd[TC_T / 2];
#pragma unroll
  for (int q = 0; q < TC_T / 2; q++) accd[q] = 0.0;
  auto drain = [&](int group) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
    for (int h = 0; h < 2; h++) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)((group & 1) * TC_T + ch * (TC_T / 2) + h * 32);
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
#pragma unroll
      for (int q = 0; q < 32; q++) accd[h * 32 + q] += (double)__uint_as_float(v[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  };
  auto done_bar = [&](int chunk) { return bar_done + 8u * (uint32_t)(chunk % TC_RAW); };
  auto ring_par = [&](int chunk) { return (uint32_t)((chunk / TC_RAW) & 1); };

  if (warp_u == 0)
    for (int c = 0; c < TC_AHEAD; c++) issue_tma(c);

  int drained = 0;
  for (int chunk = 0; chunk < nchunks; chunk++) {
    const uint32_t stage = (uint32_t)(chunk % TC_RAW);
    const unsigned char* raw = gbase + (size_t)stage * T2_STAGE;
    unsigned char* lob = gbase + (size_t)TC_RAW * T2_STAGE + (size_t)(chunk & 1) * T2_STAGE;
    if (chunk >= 2) {
      // MMAs of chunk - 2 are complete: its lo buffer (this chunk's) and its raw stage (chunk + TC_AHEAD's) are free
      mbar_wait(done_bar(chunk - 2), ring_par(chunk - 2));
      const int gdone = (chunk - 1) / TC_FLUSH;
      if (drained < gdone) { drain(drained); drained++; }
    }
    if (warp_u == 0) issue_tma(chunk + TC_AHEAD);
    mbar_wait(bar_full + 8 * stage, ring_par(chunk));     // both boxes of this chunk have landed (async proxy -> visible)
    // lo = rna_tf32(x - trunc_tf32(x)), unit by unit (layout-agnostic: the lo tile mirrors the raw tile byte for byte)
#pragma unroll
    for (int i = 0; i < 2 * T2_UNITS; i++) {
      const uint32_t off = (uint32_t)(i * TC_NT + tid) * 16u;     // over [A tile | B tile] = one stage
      const float4 v = *reinterpret_cast<const float4*>(raw + off);
      uint4 l;
      l.x = tf32_rna(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
      l.y = tf32_rna(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
      l.z = tf32_rna(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
      l.w = tf32_rna(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
      *reinterpret_cast<uint4*>(lob + off) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (warp_u == 0) {   // all 32 lanes of warp 0; the single-thread instructions are elected inside their asm blocks
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t hi = sbase + stage * T2_STAGE;
      const uint32_t lo = lo_base + (uint32_t)(chunk & 1) * T2_STAGE;
      const int group = chunk / TC_FLUSH;
      const uint32_t acc = tmem + (uint32_t)((group & 1) * TC_T);
#pragma unroll
      for (int ks = 0; ks < TC_BK / 8; ks++) {
        const uint32_t ko = (uint32_t)ks * 32u;                 // K = 8 floats = 32 bytes inside the 128-byte swizzle span
        const uint64_t ah = umma_desc_sw128(hi + ko), al = umma_desc_sw128(lo + ko);
        const uint64_t bh = umma_desc_sw128(hi + T2_TILE + ko), bl = umma_desc_sw128(lo + T2_TILE + ko);
        umma_tf32(acc, al, bh, (chunk % TC_FLUSH != 0 || ks > 0) ? 1u : 0u);
        umma_tf32(acc, ah, bl, 1u);
        umma_tf32(acc, ah, bh, 1u);
      }
      umma_commit(done_bar(chunk));
    }
  }

  if (nchunks > 0) {
    if (nchunks >= 2) mbar_wait(done_bar(nchunks - 2), ring_par(nchunks - 2));
    mbar_wait(done_bar(nchunks - 1), ring_par(nchunks - 1));
    const int ngroups = (nchunks + TC_FLUSH - 1) / TC_FLUSH;
    for (; drained < ngroups; drained++) drain(drained);
  }
  float* o = out + (int64_t)blockIdx.y * split_stride;
  const int row = m0 + lq * 32 + lane;
  if (row < ma) {
#pragma unroll
    for (int q = 0; q < TC_T / 2; q++) {
      const int col = c0 + ch * (TC_T / 2) + q;
      if (col < mb) o[row + (int64_t)col * ldo] = (float)accd[q];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(2 * TC_T) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
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
// 2-D map of a column-major n x m float block: dimension 0 = rows (contiguous), dimension 1 = columns; box 32 rows x 128 columns
int make_panel_map(CUtensorMap* tm, const float* base, int64_t n, int m, int64_t ld) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return -100;
  const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)m};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_T};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -100;
}

}  // namespace

// float Gram through tcgen05 (3xTF32).  Returns -100 when the operands do not meet the 16-byte alignment the
// cp.async staging needs (the caller then falls back to the mma.sync kernel).
int gram_tc5_f32(lb2_ctx* ctx, int64_t n, int ma, int mb, const float* A, int64_t lda, const float* B, int64_t ldb,
                 float* G, int ldg, int upper) {
  if ((lda % 4) || (ldb % 4) || ((uintptr_t)A % 16) || ((uintptr_t)B % 16)) return -100;
  const int ntm = (ma + TC_T - 1) / TC_T, ntn = (mb + TC_T - 1) / TC_T;
  const int ntiles = upper ? ntm * (ntm + 1) / 2 : ntm * ntn;
  int nsplit = ctx->sm_count / ntiles;
  if (nsplit < 1) nsplit = 1;
  const int64_t min_rows = 8 * TC_BK;
  if ((int64_t)nsplit * min_rows > n) nsplit = (int)((n + min_rows - 1) / min_rows);
  if (nsplit < 1) nsplit = 1;
  int64_t rps = (n + nsplit - 1) / nsplit;
  rps = (rps + TC_BK - 1) / TC_BK * TC_BK;
  nsplit = (int)((n + rps - 1) / rps);
  const int64_t split_stride = (int64_t)ma * mb;
  float* part = (float*)ctx_scratch(ctx, sizeof(float) * split_stride * nsplit);
  if (!part) return -1;
  bool launched = false;
  if (ctx->gram_tma != 0 && n < (int64_t)1 << 31) {   // TMA-fed variant (box coordinates are 32-bit)
    alignas(64) CUtensorMap tmA, tmB;
    if (make_panel_map(&tmA, A, n, ma, lda) == 0 && make_panel_map(&tmB, B, n, mb, ldb) == 0) {
      LB2_CUDA_OK(cudaFuncSetAttribute(gram_tc5_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM));
      gram_tc5_tma_kernel<<<dim3(ntiles, nsplit), TC_NT, T2_SMEM, ctx->stream>>>(tmA, tmB, ma, mb, n, rps, upper, ntm, part,
                                                                                 split_stride, ma);
      launched = true;
    }
  }
  if (!launched) {
    LB2_CUDA_OK(cudaFuncSetAttribute(gram_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    gram_tc5_kernel<<<dim3(ntiles, nsplit), TC_NT, TC_SMEM, ctx->stream>>>(A, lda, B, ldb, ma, mb, n, rps, upper, ntm, part,
                                                                          split_stride, ma);
  }
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return gram_reduce_f32(ctx, part, split_stride, nsplit, ma, mb, upper, G, ldg);
}

}  // namespace lb2

// lobpcg_b200/csrc/nn_tc5.cu — K4-K6 for float on the 5th-generation tensor cores:
//   Out (n x nb) = alpha * S C + beta * Out,   S: n x kd (tall), C: kd x nb (small)
// replaces sgemm_nn at src/core/lobpcg_impl.inc:36,207 (projection), src/ortho/svqb_impl.inc:101, ortho_drop_impl.inc:79
// through tcgen05.mma kind::tf32 with the accumulator in TMEM, 3xTF32 (hi*hi + hi*lo + lo*hi) like gram_tc5.cu.
//
// Per CTA: 128 rows of S (UMMA M, TMEM lanes) x up to 128 columns of C (UMMA N, chosen at run time in steps of 16).
// Per K chunk of 32 columns of S:
//   * A = S tile (128 rows x 32 k-columns).  In memory the rows are contiguous (MN-major); an MN-major no-swizzle tf32
//     operand produced zeros on B200 in this round's attempt (descriptor as CUTLASS builds it; the staged tiles were
//     verified to be right), so the tile is TRANSPOSED while it is staged: 4-byte cp.async, one warp instruction =
//     32 consecutive rows of one k-column (128 contiguous bytes of global memory), written into the proven K-major
//     no-swizzle layout (8 rows x 16 B core matrices).  The raw tile is the hi operand (the tensor core truncates), every
//     thread computes lo = rna_tf32(x - trunc(x)) for the elements it copied itself.
//   * B = C tile, K-major; C is split ONCE per call into hi / lo copies in scratch (a kd x nb matrix), both are staged
//     with cp.async, no conversion in the loop.
//   * one thread issues 4 k-steps x 3 MMAs and commits on the chunk's mbarrier.
// TMEM accumulates at most 2 chunks (24 MMAs); finished groups are drained into fp64 registers while the next group runs
// on the second TMEM accumulator — the tensor core truncates when adding into its fp32 accumulator (see gram_tc5.cu).
// Measured (n = 4.096 M): 54 TFLOP/s at 900 -> 600, 46 at 600 -> 200 (mma.sync kernel: 49 / 40).  Like gram_tc5.cu the
// loop spends ~2.5 us per 32-column chunk where the 12 MMAs need 0.7 us: the single-role structure (every warp copies,
// splits, fences and meets at one barrier per chunk) is the limit, not the tensor core.
#include <cstdint>

#include "common.cuh"
#include "context.h"
#include "kernels.h"

namespace lb2 {

namespace {

constexpr int NT_T = 128;            // rows per CTA = UMMA M; max columns per CTA = UMMA N
constexpr int NT_BK = 32;            // k-columns of S per chunk
constexpr int NT_RAW = 4;            // stages of {A raw, B hi, B lo}
constexpr int NT_AHEAD = NT_RAW - 2;
constexpr int NT_FLUSH = 2;
constexpr int NT_NT = 256;
constexpr int NT_UNITS = (NT_T * NT_BK / 4) / NT_NT;       // 16-byte units per thread and B tile (4)
constexpr int NT_AEL = (NT_T * NT_BK) / NT_NT;             // 4-byte elements per thread and A tile (16)
constexpr uint32_t NT_TILE = NT_T * NT_BK * 4;              // 16 KB
constexpr uint32_t NT_STAGE = 3 * NT_TILE;                  // A raw | B hi | B lo
constexpr uint32_t NT_LO0 = NT_RAW * NT_STAGE;              // two A-lo buffers behind the stages
constexpr uint32_t NT_BAR = NT_LO0 + 2 * NT_TILE;
constexpr uint32_t NT_SMEM = NT_BAR + 2048;                 // 226 KB
// A and B (K-major): k-units (4 k, 16 B) 128 B apart, groups of 8 rows (A) / columns (B) 1 KB apart
constexpr uint32_t B_LBO = 128, B_SBO = (NT_BK / 4) * 128;
constexpr uint32_t A_LBO = B_LBO, A_SBO = B_SBO;

__device__ __forceinline__ uint32_t smem_u32n(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_n(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait_n(uint32_t bar, uint32_t parity) {   // bounded: a lost arrival traps
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
__device__ __forceinline__ uint64_t desc_n(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_n(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  // issued from warp-uniform code, predicated on elect.sync (see gram_tc5.cu: avoids ptxas' R2UR waterfall per MMA)
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t rna_n(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return r;
}

// C (kd x nb, ld ldc) -> hi = C (copied, ld ldp), lo = rna_tf32(C - trunc_tf32(C)); rows kd..ldp-1 zero
__global__ void split_c_kernel(const float* __restrict__ C, int ldc, int kd, int nb, float* __restrict__ hi,
                               float* __restrict__ lo, int ldp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ldp * nb) return;
  const int k = idx % ldp, j = idx / ldp;
  const float v = (k < kd) ? C[k + (int64_t)j * ldc] : 0.f;
  hi[idx] = v;
  lo[idx] = __uint_as_float(rna_n(v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u)));
}

__global__ void __launch_bounds__(NT_NT, 1)
    nn_tc5_kernel(const float* __restrict__ S, int64_t lds, const float* __restrict__ Chi, const float* __restrict__ Clo,
                  int ldp, float* __restrict__ Out, int64_t ldo, int64_t n, int kd, int nb, int nct, float alpha,
                  float beta) {
  extern __shared__ __align__(1024) unsigned char smem_nt[];
  const uint32_t sbase = (smem_u32n(smem_nt) + 1023u) & ~1023u;
  unsigned char* gbase = smem_nt + (sbase - smem_u32n(smem_nt));
  const uint32_t bar0 = sbase + NT_BAR;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + NT_BAR + 64);

  const int ct = blockIdx.x % nct;
  const int64_t r0 = (int64_t)(blockIdx.x / nct) * NT_T;
  const int c0 = ct * NT_T;
  const int tn = min(NT_T, (nb - c0 + 15) / 16 * 16);       // UMMA N of this CTA
  const int nchunks = (kd + NT_BK - 1) / NT_BK;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp index the compiler knows to be uniform
  // D = F32, A = B = TF32, both K-major, N = tn, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(tn >> 3) << 17) | ((uint32_t)(NT_T >> 4) << 24);

  if (tid == 0) {
    for (int s = 0; s < NT_RAW; s++) mbar_init_n(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32n((const void*)tmem_slot)),
                 "n"(2 * NT_T)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // copy units.  A: 4-byte elements, a warp instruction = 32 consecutive rows of one k-column; B: (column c, k-unit u),
  // 8 columns x 4 units per warp instruction
  const int cc = lane & 7, uu = lane >> 3;
  uint32_t aoff[NT_AEL], boff[NT_UNITS];
  int bcol[NT_UNITS], bk[NT_UNITS];
#pragma unroll
  for (int i = 0; i < NT_AEL; i++) {
    const int idx = i * (NT_NT / 32) + warp;     // 0..127: k-column idx & 31, block of 32 rows idx >> 5
    const int kc = idx & 31, m = (idx >> 5) * 32 + lane;
    aoff[i] = (uint32_t)(m & 7) * 16u + (uint32_t)(m >> 3) * A_SBO + (uint32_t)(kc >> 2) * A_LBO + (uint32_t)(kc & 3) * 4u;
  }
#pragma unroll
  for (int i = 0; i < NT_UNITS; i++) {
    const int blk = i * (NT_NT / 32) + warp;     // 0..31: 16 column groups x 2 halves of the 8 k-units
    const int cg = blk & 15, uh = blk >> 4;
    const int u = uh * 4 + uu;
    bcol[i] = cg * 8 + cc;
    bk[i] = u * 4;
    boff[i] = (uint32_t)cg * B_SBO + (uint32_t)u * B_LBO + (uint32_t)cc * 16u;
  }
  const int64_t rows_valid = n - r0;      // >= 1

  auto issue_loads = [&](int chunk) {
    if (chunk < nchunks) {
      const uint32_t st = sbase + (uint32_t)(chunk % NT_RAW) * NT_STAGE;
      const int k0 = chunk * NT_BK;
#pragma unroll
      for (int i = 0; i < NT_AEL; i++) {
        const int idx = i * (NT_NT / 32) + warp;
        const int kc = idx & 31, m = (idx >> 5) * 32 + lane;
        const bool ok = (k0 + kc < kd) && (m < rows_valid);
        const float* src = ok ? S + (int64_t)(k0 + kc) * lds + r0 + m : S;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(st + aoff[i]), "l"(src), "r"(ok ? 4 : 0));
      }
#pragma unroll
      for (int i = 0; i < NT_UNITS; i++) {
        // the split copies are zero-padded to ldp >= round_up(kd, 32) rows, so a k-unit is always whole
        const bool ok = (c0 + bcol[i] < nb);
        const int64_t o = (int64_t)(c0 + bcol[i]) * ldp + k0 + bk[i];
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(st + NT_TILE + boff[i]),
                     "l"(ok ? Chi + o : Chi), "r"(ok ? 16 : 0));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(st + 2 * NT_TILE + boff[i]),
                     "l"(ok ? Clo + o : Clo), "r"(ok ? 16 : 0));
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };

  const int lq = warp & 3, ch = warp >> 2;      // TMEM lane quarter of this warp, column half
  double accd[NT_T / 2];
#pragma unroll
  for (int q = 0; q < NT_T / 2; q++) accd[q] = 0.0;
  auto drain = [&](int group) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
    for (int h = 0; h < 2; h++) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)((group & 1) * NT_T + ch * (NT_T / 2) + h * 32);
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
  auto done_bar = [&](int chunk) { return bar0 + 8u * (uint32_t)(chunk % NT_RAW); };
  auto done_par = [&](int chunk) { return (uint32_t)((chunk / NT_RAW) & 1); };

#pragma unroll
  for (int c = 0; c < NT_AHEAD; c++) issue_loads(c);

  int drained = 0;
  for (int chunk = 0; chunk < nchunks; chunk++) {
    const unsigned char* raw = gbase + (size_t)(chunk % NT_RAW) * NT_STAGE;
    unsigned char* lob = gbase + NT_LO0 + (size_t)(chunk & 1) * NT_TILE;
    asm volatile("cp.async.wait_group %0;\n" ::"n"(NT_AHEAD - 1) : "memory");
    if (chunk >= 2) {
      mbar_wait_n(done_bar(chunk - 2), done_par(chunk - 2));   // frees this chunk's lo buffer and the stage of chunk + AHEAD
      const int gdone = (chunk - 1) / NT_FLUSH;
      if (drained < gdone) { drain(drained); drained++; }
    }
    issue_loads(chunk + NT_AHEAD);
    float rv[NT_AEL];
#pragma unroll
    for (int i = 0; i < NT_AEL; i++) rv[i] = *reinterpret_cast<const float*>(raw + aoff[i]);
#pragma unroll
    for (int i = 0; i < NT_AEL; i++)
      *reinterpret_cast<uint32_t*>(lob + aoff[i]) = rna_n(rv[i] - __uint_as_float(__float_as_uint(rv[i]) & 0xFFFFE000u));
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (warp_u == 0) {   // all 32 lanes of warp 0
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t st = sbase + (uint32_t)(chunk % NT_RAW) * NT_STAGE;
      const uint32_t alo = sbase + NT_LO0 + (uint32_t)(chunk & 1) * NT_TILE;
      const uint32_t acc = tmem + (uint32_t)(((chunk / NT_FLUSH) & 1) * NT_T);
#pragma unroll
      for (int ks = 0; ks < NT_BK / 8; ks++) {
        const uint64_t ah = desc_n(st + (uint32_t)ks * 2u * A_LBO, A_LBO, A_SBO), al = desc_n(alo + (uint32_t)ks * 2u * A_LBO, A_LBO, A_SBO);
        const uint64_t bh = desc_n(st + NT_TILE + (uint32_t)ks * 2u * B_LBO, B_LBO, B_SBO);
        const uint64_t bl = desc_n(st + 2 * NT_TILE + (uint32_t)ks * 2u * B_LBO, B_LBO, B_SBO);
        umma_n(acc, al, bh, idesc, (chunk % NT_FLUSH != 0 || ks > 0) ? 1u : 0u);
        umma_n(acc, ah, bl, idesc, 1u);
        umma_n(acc, ah, bh, idesc, 1u);
      }
      asm volatile(
          "{\n\t.reg .pred q;\n\t"
          "elect.sync _|q, 0xffffffff;\n\t"
          "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(done_bar(chunk))
          : "memory");
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");

  if (nchunks > 0) {
    if (nchunks >= 2) mbar_wait_n(done_bar(nchunks - 2), done_par(nchunks - 2));
    mbar_wait_n(done_bar(nchunks - 1), done_par(nchunks - 1));
    const int ngroups = (nchunks + NT_FLUSH - 1) / NT_FLUSH;
    for (; drained < ngroups; drained++) drain(drained);
  }
  const int64_t row = r0 + lq * 32 + lane;      // TMEM lane = row of the tile
  if (row < n) {
#pragma unroll
    for (int q = 0; q < NT_T / 2; q++) {
      const int col = c0 + ch * (NT_T / 2) + q;
      if (col < nb) {
        float* p = Out + row + (int64_t)col * ldo;
        float v = alpha * (float)accd[q];
        if (beta != 0.f) v += beta * (*p);
        *p = v;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(2 * NT_T) : "memory");
}

}  // namespace

// float projection through tcgen05 (3xTF32).
int nn_tc5_f32(lb2_ctx* ctx, int64_t n, int kd, int nb, float alpha, const float* S, int64_t lds, const float* C, int ldc,
               float beta, float* Out, int64_t ldo) {
  const int ldp = (kd + NT_BK - 1) / NT_BK * NT_BK;
  // scratch: hi | lo copies of C (the Gram kernels use the same scratch, but never concurrently: one stream)
  float* sc = (float*)ctx_scratch(ctx, sizeof(float) * 2 * (size_t)ldp * nb);
  if (!sc) return -1;
  float* Chi = sc;
  float* Clo = sc + (size_t)ldp * nb;
  split_c_kernel<<<(ldp * nb + 255) / 256, 256, 0, ctx->stream>>>(C, ldc, kd, nb, Chi, Clo, ldp);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  const int nct = (nb + NT_T - 1) / NT_T;
  const int64_t nrt = (n + NT_T - 1) / NT_T;
  LB2_CUDA_OK(cudaFuncSetAttribute(nn_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NT_SMEM));
  nn_tc5_kernel<<<(unsigned)(nrt * nct), NT_NT, NT_SMEM, ctx->stream>>>(S, lds, Chi, Clo, ldp, Out, ldo, n, kd, nb, nct, alpha,
                                                                       beta);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace lb2

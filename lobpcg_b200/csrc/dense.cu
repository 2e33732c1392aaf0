// lobpcg_b200/csrc/dense.cu — the two tall-skinny contraction families of the LOBPCG hot path.
//
//   K2/K3  Gram      G (ma x mb)  = A^H B          A: n x ma, B: n x mb, column-major, n >> ma,mb
//          replaces syrk/herk + gemm_tn/hn at  src/gram/gram_impl.inc:54-63,92-101,
//          src/rayleigh/rayleigh_ritz_modified_impl.inc:75-77,193-195 (reference).
//   K4-K6  tall NN   Out (n x nb) = alpha * S C + beta * Out     S: n x kd, C: kd x nb
//          replaces gemm_nn at src/core/lobpcg_impl.inc:36,207 (projection), src/ortho/svqb_impl.inc:101
//          (U <- U T), src/ortho/ortho_drop_impl.inc:79 (U -= V C).
//
// f64 runs on the FP64 tensor pipe (DMMA.8x8x4 via mma.sync.m8n8k4 — tcgen05 has no f64 kind), operands
// staged global->shared with cp.async (LDGSTS) in a multi-stage ring, fragments read conflict-free from
// padded tiles (row stride == 4 mod 16 doubles).  The Gram contraction is a reduction over n: each CTA
// owns one output tile and one contiguous n-range ("split"), partial tiles go to scratch and a second
// kernel sums them in a fixed order (deterministic, no atomics) and mirrors the Hermitian half.
// c64 uses four real DMMAs per complex MAC; f32 runs 3xTF32 on mma.sync here (Gram fallback, projection) — the default
// f32 Gram is the tcgen05 kernel of gram_tc5.cu, the default f64 Hermitian Gram the work-list kernel of gram_wl.cu;
// c32 uses the generic register-blocked SIMT kernels at the bottom.
#include "common.cuh"
#include "context.h"
#include <algorithm>
#include <cmath>
#include "kernels.h"
#include "tile_loader.cuh"

namespace lb2 {

// =====================================================================================================
// cp.async tile loader: NCOLS columns x RUN contiguous doubles -> smem[c*LDS + e], zero-filled outside
// [run0,run_end) x [col0,col_end).
// =====================================================================================================
template <int NCOLS, int RUN, int LDS, int NT, bool VEC>
__device__ __forceinline__ void load_tile_f64(double* sm, const double* __restrict__ base, int64_t ld,
                                              int64_t run0, int64_t run_end, int col0, int col_end,
                                              int tid) {
  if constexpr (VEC) {
    constexpr int CPC = RUN / 2;
    constexpr int TOTAL = NCOLS * CPC;
#pragma unroll
    for (int id0 = 0; id0 < TOTAL; id0 += NT) {
      const int id = id0 + tid;
      if (TOTAL % NT != 0 && id >= TOTAL) break;
      const int c = id / CPC, ch = id % CPC;
      const int64_t e = run0 + 2 * ch;
      int bytes = 0;
      const double* src = base;
      if (col0 + c < col_end && e < run_end) {
        bytes = (run_end - e >= 2) ? 16 : 8;
        src = base + (int64_t)(col0 + c) * ld + e;
      }
      cp_async_zfill<16>(sm + c * LDS + 2 * ch, src, bytes);
    }
  } else {
    constexpr int TOTAL = NCOLS * RUN;
#pragma unroll
    for (int id0 = 0; id0 < TOTAL; id0 += NT) {
      const int id = id0 + tid;
      if (TOTAL % NT != 0 && id >= TOTAL) break;
      const int c = id / RUN, el = id % RUN;
      const int64_t e = run0 + el;
      int bytes = 0;
      const double* src = base;
      if (col0 + c < col_end && e < run_end) {
        bytes = 8;
        src = base + (int64_t)(col0 + c) * ld + e;
      }
      cp_async_zfill<8>(sm + c * LDS + el, src, bytes);
    }
  }
}

// upper-triangular tile enumeration: t -> (i <= j)
__device__ __forceinline__ void upper_tile(int t, int& i, int& j) {
  j = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
  while ((j + 1) * (j + 2) / 2 <= t) ++j;
  while (j * (j + 1) / 2 > t) --j;
  i = t - j * (j + 1) / 2;
}

// =====================================================================================================
// K2/K3 f64: DMMA Gram.  grid = (ntiles, nsplit).
// =====================================================================================================
template <int TM, int TN, int WM, int WN, int BK, int STAGES, int OCC, bool VEC>
__global__ void __launch_bounds__(WM* WN * 32, OCC)
    gram_dmma_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                     int ma, int mb, int64_t n, int64_t rows_per_split, int upper, int ntm,
                     double* __restrict__ out, int64_t split_stride, int ldo) {
  constexpr int NT = WM * WN * 32;
  constexpr int LDS = BK + 4;
  constexpr int MB = TM / WM / 8;
  constexpr int NB = TN / WN / 8;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + (size_t)STAGES * TM * LDS;

  int ti, tj;
  if (upper) upper_tile(blockIdx.x, ti, tj);
  else { ti = blockIdx.x % ntm; tj = blockIdx.x / ntm; }
  const int m0 = ti * TM, c0 = tj * TN;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(n, r_begin + rows_per_split);
  const int nchunks = (r_end > r_begin) ? (int)((r_end - r_begin + BK - 1) / BK) : 0;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int g = lane >> 2, t = lane & 3;

  double acc[MB][NB][2];
#pragma unroll
  for (int i = 0; i < MB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  TileLoaderF64<TM, BK, LDS, NT, VEC> la;
  TileLoaderF64<TN, BK, LDS, NT, VEC> lb;
  la.init(A, lda, r_begin, m0, ma, tid);
  lb.init(B, ldb, r_begin, c0, mb, tid);
  int issued = 0, wstage = 0;
  auto issue = [&]() {
    if (issued < nchunks) {
      const int64_t valid = (r_end - r_begin) - (int64_t)issued * BK;
      la.issue(As + wstage * (TM * LDS), A, valid);
      lb.issue(Bs + wstage * (TN * LDS), B, valid);
      la.advance(BK);
      lb.advance(BK);
    }
    issued++;
    wstage = (wstage + 1 == STAGES) ? 0 : wstage + 1;
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) issue();

  int rstage = 0;
  for (int chunk = 0; chunk < nchunks; chunk++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue();
    const double* as = As + rstage * (TM * LDS) + (wm * MB * 8 + g) * LDS + t;
    const double* bs = Bs + rstage * (TN * LDS) + (wn * NB * 8 + g) * LDS + t;
    rstage = (rstage + 1 == STAGES) ? 0 : rstage + 1;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ks++) {
      double a[MB], b[NB];
#pragma unroll
      for (int i = 0; i < MB; i++) a[i] = as[i * 8 * LDS + ks * 4];
#pragma unroll
      for (int j = 0; j < NB; j++) b[j] = bs[j * 8 * LDS + ks * 4];
#pragma unroll
      for (int i = 0; i < MB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) dmma884(acc[i][j], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  double* o = out + (int64_t)blockIdx.y * split_stride;
#pragma unroll
  for (int i = 0; i < MB; i++) {
    const int row = m0 + wm * MB * 8 + i * 8 + g;
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const int col = c0 + wn * NB * 8 + j * 8 + 2 * t;
      if (row < ma) {
        if (col < mb) o[row + (int64_t)col * ldo] = acc[i][j][0];
        if (col + 1 < mb) o[row + (int64_t)(col + 1) * ldo] = acc[i][j][1];
      }
    }
  }
}

// =====================================================================================================
// c64 (double complex) on the FP64 tensor pipe: a complex MAC is four real DMMAs on the (re, im) parts of the
// fragments.  Tiles hold interleaved complex elements (16 bytes, one cp.async each, always aligned); fragments are
// read with 128-bit shared loads, conflict-free with a row stride == 4 (mod 8) complex elements ([col][k] tiles)
// or == 2 (mod 8) ([k][row] tiles).
// =====================================================================================================
template <int NCOLS, int RUN, int LDS, int NT>
__device__ __forceinline__ void load_tile_c64(c64* sm, const c64* __restrict__ base, int64_t ld, int64_t run0,
                                              int64_t run_end, int col0, int col_end, int tid) {
  constexpr int TOTAL = NCOLS * RUN;
#pragma unroll
  for (int id0 = 0; id0 < TOTAL; id0 += NT) {
    const int id = id0 + tid;
    if (TOTAL % NT != 0 && id >= TOTAL) break;
    const int c = id / RUN, el = id % RUN;
    const int64_t e = run0 + el;
    const bool ok = (col0 + c < col_end) && (e < run_end);
    const c64* src = ok ? base + (int64_t)(col0 + c) * ld + e : base;
    cp_async_zfill<16>(sm + c * LDS + el, src, ok ? 16 : 0);
  }
}

// G = A^H B (c64).  grid = (ntiles, nsplit), same split-n / deterministic reduction scheme as the real kernel.
template <int TM, int TN, int WM, int WN, int BK, int STAGES, int OCC>
__global__ void __launch_bounds__(WM* WN * 32, OCC)
    gram_zmma_kernel(const c64* __restrict__ A, int64_t lda, const c64* __restrict__ B, int64_t ldb, int ma, int mb,
                     int64_t n, int64_t rows_per_split, int upper, int ntm, c64* __restrict__ out,
                     int64_t split_stride, int ldo) {
  constexpr int NT = WM * WN * 32;
  constexpr int LDS = BK + 4;
  constexpr int MB = TM / WM / 8;
  constexpr int NB = TN / WN / 8;
  extern __shared__ __align__(16) unsigned char smem_raw_z[];
  c64* As = reinterpret_cast<c64*>(smem_raw_z);
  c64* Bs = As + (size_t)STAGES * TM * LDS;
  int ti, tj;
  if (upper) upper_tile(blockIdx.x, ti, tj);
  else { ti = blockIdx.x % ntm; tj = blockIdx.x / ntm; }
  const int m0 = ti * TM, c0 = tj * TN;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(n, r_begin + rows_per_split);
  const int nchunks = (r_end > r_begin) ? (int)((r_end - r_begin + BK - 1) / BK) : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int g = lane >> 2, t = lane & 3;
  double are[MB][NB][2], aim[MB][NB][2];
#pragma unroll
  for (int i = 0; i < MB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) are[i][j][0] = are[i][j][1] = aim[i][j][0] = aim[i][j][1] = 0.0;
  auto issue = [&](int chunk) {
    if (chunk < nchunks) {
      const int s = chunk % STAGES;
      const int64_t r = r_begin + (int64_t)chunk * BK;
      load_tile_c64<TM, BK, LDS, NT>(As + (size_t)s * TM * LDS, A, lda, r, r_end, m0, ma, tid);
      load_tile_c64<TN, BK, LDS, NT>(Bs + (size_t)s * TN * LDS, B, ldb, r, r_end, c0, mb, tid);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) issue(s);
  for (int chunk = 0; chunk < nchunks; chunk++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue(chunk + STAGES - 1);
    const c64* as = As + (size_t)(chunk % STAGES) * TM * LDS + (wm * MB * 8 + g) * LDS + t;
    const c64* bs = Bs + (size_t)(chunk % STAGES) * TN * LDS + (wn * NB * 8 + g) * LDS + t;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ks++) {
      c64 a[MB], b[NB];
#pragma unroll
      for (int i = 0; i < MB; i++) a[i] = as[i * 8 * LDS + ks * 4];
#pragma unroll
      for (int j = 0; j < NB; j++) b[j] = bs[j * 8 * LDS + ks * 4];
#pragma unroll
      for (int i = 0; i < MB; i++) {
        const double nim = -a[i].im;
#pragma unroll
        for (int j = 0; j < NB; j++) {   // conj(a) * b
          dmma884(are[i][j], a[i].re, b[j].re);
          dmma884(are[i][j], a[i].im, b[j].im);
          dmma884(aim[i][j], a[i].re, b[j].im);
          dmma884(aim[i][j], nim, b[j].re);
        }
      }
    }
  }
  cp_async_wait<0>();
  c64* o = out + (int64_t)blockIdx.y * split_stride;
#pragma unroll
  for (int i = 0; i < MB; i++) {
    const int row = m0 + wm * MB * 8 + i * 8 + g;
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const int col = c0 + wn * NB * 8 + j * 8 + 2 * t;
      if (row < ma) {
        if (col < mb) o[row + (int64_t)col * ldo] = c64{are[i][j][0], aim[i][j][0]};
        if (col + 1 < mb) o[row + (int64_t)(col + 1) * ldo] = c64{are[i][j][1], aim[i][j][1]};
      }
    }
  }
}

// Out = alpha S C + beta Out (c64)
template <int TM, int TN, int WM, int WN, int BK, int STAGES, int OCC>
__global__ void __launch_bounds__(WM* WN * 32, OCC)
    tall_nn_zmma_kernel(const c64* __restrict__ S, int64_t lds, const c64* __restrict__ C, int ldc,
                        c64* __restrict__ Out, int64_t ldo, int64_t n, int kd, int nb, int nct, c64 alpha, c64 beta,
                        int beta_zero) {
  constexpr int NT = WM * WN * 32;
  constexpr int LDA = TM + 2;
  constexpr int LDB = BK + 4;
  constexpr int MB = TM / WM / 8;
  constexpr int NB = TN / WN / 8;
  extern __shared__ __align__(16) unsigned char smem_raw_z[];
  c64* Ss = reinterpret_cast<c64*>(smem_raw_z);       // [STAGES][BK][LDA]
  c64* Cs = Ss + (size_t)STAGES * BK * LDA;            // [STAGES][TN][LDB]
  const int ct = blockIdx.x % nct;
  const int64_t rt = blockIdx.x / nct;
  const int64_t r0 = rt * TM;
  const int c0 = ct * TN;
  const int nchunks = (kd + BK - 1) / BK;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int g = lane >> 2, t = lane & 3;
  double are[MB][NB][2], aim[MB][NB][2];
#pragma unroll
  for (int i = 0; i < MB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) are[i][j][0] = are[i][j][1] = aim[i][j][0] = aim[i][j][1] = 0.0;
  auto issue = [&](int chunk) {
    if (chunk < nchunks) {
      const int s = chunk % STAGES;
      const int k0 = chunk * BK;
      load_tile_c64<BK, TM, LDA, NT>(Ss + (size_t)s * BK * LDA, S, lds, r0, n, k0, kd, tid);
      load_tile_c64<TN, BK, LDB, NT>(Cs + (size_t)s * TN * LDB, C, ldc, k0, kd, c0, nb, tid);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) issue(s);
  for (int chunk = 0; chunk < nchunks; chunk++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue(chunk + STAGES - 1);
    const c64* as = Ss + (size_t)(chunk % STAGES) * BK * LDA + t * LDA + wm * MB * 8 + g;
    const c64* bs = Cs + (size_t)(chunk % STAGES) * TN * LDB + (wn * NB * 8 + g) * LDB + t;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ks++) {
      c64 a[MB], b[NB];
#pragma unroll
      for (int i = 0; i < MB; i++) a[i] = as[ks * 4 * LDA + i * 8];
#pragma unroll
      for (int j = 0; j < NB; j++) b[j] = bs[j * 8 * LDB + ks * 4];
#pragma unroll
      for (int i = 0; i < MB; i++) {
        const double nim = -a[i].im;
#pragma unroll
        for (int j = 0; j < NB; j++) {   // a * b
          dmma884(are[i][j], a[i].re, b[j].re);
          dmma884(are[i][j], nim, b[j].im);
          dmma884(aim[i][j], a[i].re, b[j].im);
          dmma884(aim[i][j], a[i].im, b[j].re);
        }
      }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < MB; i++) {
    const int64_t row = r0 + wm * MB * 8 + i * 8 + g;
    if (row >= n) continue;
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const int col = c0 + wn * NB * 8 + j * 8 + 2 * t;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        if (col + q < nb) {
          c64* p = Out + row + (int64_t)(col + q) * ldo;
          c64 v = mul_(alpha, c64{are[i][j][q], aim[i][j][q]});
          if (!beta_zero) v = add_(v, mul_(beta, *p));
          *p = v;
        }
      }
    }
  }
}

// =====================================================================================================
// f32 on the tensor cores: 3xTF32 (hi*hi + hi*lo + lo*hi, fp32 accumulate) through mma.sync.m16n8k8.tf32 — the
// algorithm's own tolerances (EPS_TOL = 1e-5 for s/c, src/core/lobpcg_s.c:10) need fp32-accurate Grams, plain TF32
// (10-bit mantissa) is not enough (SURVEY §7 hard part 2).  Same cp.async ring / split-n structure as the f64
// kernels; tiles are [col][BK+4] floats with BK = 32 (row stride == 4 mod 32 => conflict-free fragment loads).
// =====================================================================================================
template <int NCOLS, int RUN, int LDS, int NT, bool VEC>
__device__ __forceinline__ void load_tile_f32(float* sm, const float* __restrict__ base, int64_t ld, int64_t run0,
                                              int64_t run_end, int col0, int col_end, int tid) {
  if constexpr (VEC) {
    constexpr int CPC = RUN / 4;
    constexpr int TOTAL = NCOLS * CPC;
#pragma unroll
    for (int id0 = 0; id0 < TOTAL; id0 += NT) {
      const int id = id0 + tid;
      if (TOTAL % NT != 0 && id >= TOTAL) break;
      const int c = id / CPC, ch = id % CPC;
      const int64_t e = run0 + 4 * ch;
      int bytes = 0;
      const float* src = base;
      if (col0 + c < col_end && e < run_end) {
        const int64_t left = run_end - e;
        bytes = left >= 4 ? 16 : (int)left * 4;
        src = base + (int64_t)(col0 + c) * ld + e;
      }
      cp_async_zfill<16>(sm + c * LDS + 4 * ch, src, bytes);
    }
  } else {
    constexpr int TOTAL = NCOLS * RUN;
#pragma unroll
    for (int id0 = 0; id0 < TOTAL; id0 += NT) {
      const int id = id0 + tid;
      if (TOTAL % NT != 0 && id >= TOTAL) break;
      const int c = id / RUN, el = id % RUN;
      const int64_t e = run0 + el;
      const bool ok = (col0 + c < col_end) && (e < run_end);
      cp_async_zfill<4>(sm + c * LDS + el, ok ? base + (int64_t)(col0 + c) * ld + e : base, ok ? 4 : 0);
    }
  }
}

template <int TM, int TN, int WM, int WN, int BK, int STAGES, bool VEC>
__global__ void __launch_bounds__(WM* WN * 32)
    gram_tf32_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, int ma,
                     int mb, int64_t n, int64_t rows_per_split, int upper, int ntm, float* __restrict__ out,
                     int64_t split_stride, int ldo) {
  constexpr int NT = WM * WN * 32;
  constexpr int LDS = BK + 4;
  constexpr int MB = TM / WM / 16;   // 16-row MMA blocks per warp
  constexpr int NB = TN / WN / 8;    // 8-col MMA blocks per warp
  extern __shared__ __align__(16) unsigned char smem_raw_f[];
  float* As = reinterpret_cast<float*>(smem_raw_f);
  float* Bs = As + (size_t)STAGES * TM * LDS;
  int ti, tj;
  if (upper) upper_tile(blockIdx.x, ti, tj);
  else { ti = blockIdx.x % ntm; tj = blockIdx.x / ntm; }
  const int m0 = ti * TM, c0 = tj * TN;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(n, r_begin + rows_per_split);
  const int nchunks = (r_end > r_begin) ? (int)((r_end - r_begin + BK - 1) / BK) : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int g = lane >> 2, t = lane & 3;
  float acc[MB][NB][4], accs[MB][NB][4];
#pragma unroll
  for (int i = 0; i < MB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) acc[i][j][q] = accs[i][j][q] = 0.f;
  auto issue = [&](int chunk) {
    if (chunk < nchunks) {
      const int s = chunk % STAGES;
      const int64_t r = r_begin + (int64_t)chunk * BK;
      load_tile_f32<TM, BK, LDS, NT, VEC>(As + (size_t)s * TM * LDS, A, lda, r, r_end, m0, ma, tid);
      load_tile_f32<TN, BK, LDS, NT, VEC>(Bs + (size_t)s * TN * LDS, B, ldb, r, r_end, c0, mb, tid);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) issue(s);
  for (int chunk = 0; chunk < nchunks; chunk++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue(chunk + STAGES - 1);
    const float* as = As + (size_t)(chunk % STAGES) * TM * LDS + (wm * MB * 16 + g) * LDS + t;
    const float* bs = Bs + (size_t)(chunk % STAGES) * TN * LDS + (wn * NB * 8 + g) * LDS + t;
#pragma unroll
    for (int ks = 0; ks < BK / 8; ks++) {
      uint32_t ah[MB][4], al[MB][4], bh[NB][2], bl[NB][2];
#pragma unroll
      for (int i = 0; i < MB; i++) {
        const float* p = as + i * 16 * LDS + ks * 8;
        split_tf32(p[0], ah[i][0], al[i][0]);
        split_tf32(p[8 * LDS], ah[i][1], al[i][1]);
        split_tf32(p[4], ah[i][2], al[i][2]);
        split_tf32(p[8 * LDS + 4], ah[i][3], al[i][3]);
      }
#pragma unroll
      for (int j = 0; j < NB; j++) {
        const float* p = bs + j * 8 * LDS + ks * 8;
        split_tf32(p[0], bh[j][0], bl[j][0]);
        split_tf32(p[4], bh[j][1], bl[j][1]);
      }
#pragma unroll
      for (int i = 0; i < MB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) {
          mma_tf32_1688(acc[i][j], al[i], bh[j]);
          mma_tf32_1688(acc[i][j], ah[i], bl[j]);
          mma_tf32_1688(acc[i][j], ah[i], bh[j]);
        }
    }
    // the tensor core truncates when it adds into its fp32 accumulator (a bias of ~2^-25 per MMA that grows with the
    // length of the sum, tools/f32_gram_accuracy.py): keep the hardware accumulation to one chunk and do the long sum
    // with round-to-nearest adds
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
      for (int j = 0; j < NB; j++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
          accs[i][j][q] += acc[i][j][q];
          acc[i][j][q] = 0.f;
        }
  }
  cp_async_wait<0>();
  float* o = out + (int64_t)blockIdx.y * split_stride;
#pragma unroll
  for (int i = 0; i < MB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int row = m0 + wm * MB * 16 + i * 16 + g + ((q & 2) ? 8 : 0);
        const int col = c0 + wn * NB * 8 + j * 8 + 2 * t + (q & 1);
        if (row < ma && col < mb) o[row + (int64_t)col * ldo] = accs[i][j][q];
      }
}

template <int TM, int TN, int WM, int WN, int BK, int STAGES, bool VECA, bool VECB>
__global__ void __launch_bounds__(WM* WN * 32)
    tall_nn_tf32_kernel(const float* __restrict__ S, int64_t lds, const float* __restrict__ C, int ldc,
                        float* __restrict__ Out, int64_t ldo, int64_t n, int kd, int nb, int nct, float alpha,
                        float beta) {
  constexpr int NT = WM * WN * 32;
  constexpr int LDA = TM + 8;        // [k][row] tile, row stride == 8 mod 32
  constexpr int LDB = BK + 4;        // [col][k] tile
  constexpr int MB = TM / WM / 16;
  constexpr int NB = TN / WN / 8;
  extern __shared__ __align__(16) unsigned char smem_raw_f[];
  float* Ss = reinterpret_cast<float*>(smem_raw_f);   // [STAGES][BK][LDA]
  float* Cs = Ss + (size_t)STAGES * BK * LDA;          // [STAGES][TN][LDB]
  const int ct = blockIdx.x % nct;
  const int64_t rt = blockIdx.x / nct;
  const int64_t r0 = rt * TM;
  const int c0 = ct * TN;
  const int nchunks = (kd + BK - 1) / BK;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int g = lane >> 2, t = lane & 3;
  float acc[MB][NB][4], accs[MB][NB][4];
#pragma unroll
  for (int i = 0; i < MB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) acc[i][j][q] = accs[i][j][q] = 0.f;
  auto issue = [&](int chunk) {
    if (chunk < nchunks) {
      const int s = chunk % STAGES;
      const int k0 = chunk * BK;
      load_tile_f32<BK, TM, LDA, NT, VECA>(Ss + (size_t)s * BK * LDA, S, lds, r0, n, k0, kd, tid);
      load_tile_f32<TN, BK, LDB, NT, VECB>(Cs + (size_t)s * TN * LDB, C, ldc, k0, kd, c0, nb, tid);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) issue(s);
  for (int chunk = 0; chunk < nchunks; chunk++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue(chunk + STAGES - 1);
    const float* as = Ss + (size_t)(chunk % STAGES) * BK * LDA + t * LDA + wm * MB * 16 + g;
    const float* bs = Cs + (size_t)(chunk % STAGES) * TN * LDB + (wn * NB * 8 + g) * LDB + t;
#pragma unroll
    for (int ks = 0; ks < BK / 8; ks++) {
      uint32_t ah[MB][4], al[MB][4], bh[NB][2], bl[NB][2];
#pragma unroll
      for (int i = 0; i < MB; i++) {
        const float* p = as + ks * 8 * LDA + i * 16;
        split_tf32(p[0], ah[i][0], al[i][0]);
        split_tf32(p[8], ah[i][1], al[i][1]);
        split_tf32(p[4 * LDA], ah[i][2], al[i][2]);
        split_tf32(p[4 * LDA + 8], ah[i][3], al[i][3]);
      }
#pragma unroll
      for (int j = 0; j < NB; j++) {
        const float* p = bs + j * 8 * LDB + ks * 8;
        split_tf32(p[0], bh[j][0], bl[j][0]);
        split_tf32(p[4], bh[j][1], bl[j][1]);
      }
#pragma unroll
      for (int i = 0; i < MB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) {
          mma_tf32_1688(acc[i][j], al[i], bh[j]);
          mma_tf32_1688(acc[i][j], ah[i], bl[j]);
          mma_tf32_1688(acc[i][j], ah[i], bh[j]);
        }
    }
    // the tensor core truncates when it adds into its fp32 accumulator (a bias of ~2^-25 per MMA that grows with the
    // length of the sum, tools/f32_gram_accuracy.py): keep the hardware accumulation to one chunk and do the long sum
    // with round-to-nearest adds
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
      for (int j = 0; j < NB; j++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
          accs[i][j][q] += acc[i][j][q];
          acc[i][j][q] = 0.f;
        }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < MB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int64_t row = r0 + wm * MB * 16 + i * 16 + g + ((q & 2) ? 8 : 0);
        const int col = c0 + wn * NB * 8 + j * 8 + 2 * t + (q & 1);
        if (row < n && col < nb) {
          float* p = Out + row + (int64_t)col * ldo;
          float v = alpha * accs[i][j][q];
          if (beta != 0.f) v += beta * (*p);
          *p = v;
        }
      }
}

// =====================================================================================================
// Generic SIMT Gram (all scalar types): 64x64 tile, BK=16, 256 threads x (4x4) outputs.
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
    gram_simt_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb, int ma,
                     int mb, int64_t n, int64_t rows_per_split, int upper, int ntm, T* __restrict__ out,
                     int64_t split_stride, int ldo) {
  constexpr int TS = 64, BK = 16;
  __shared__ T As[BK][TS + 1];
  __shared__ T Bs[BK][TS + 1];
  int ti, tj;
  if (upper) upper_tile(blockIdx.x, ti, tj);
  else { ti = blockIdx.x % ntm; tj = blockIdx.x / ntm; }
  const int m0 = ti * TS, c0 = tj * TS;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(n, r_begin + rows_per_split);
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = zero<T>();
  const int lc = tid / 4, lk = (tid % 4) * 4;
  for (int64_t r = r_begin; r < r_end; r += BK) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int64_t rr = r + lk + q;
      T va = zero<T>(), vb = zero<T>();
      if (rr < r_end) {
        if (m0 + lc < ma) va = A[(int64_t)(m0 + lc) * lda + rr];
        if (c0 + lc < mb) vb = B[(int64_t)(c0 + lc) * ldb + rr];
      }
      As[lk + q][lc] = va;
      Bs[lk + q][lc] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; k++) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) fmac_(acc[i][j], a[i], b[j]);
    }
    __syncthreads();
  }
  T* o = out + (int64_t)blockIdx.y * split_stride;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int row = m0 + ty * 4 + i, col = c0 + tx * 4 + j;
      if (row < ma && col < mb) o[row + (int64_t)col * ldo] = acc[i][j];
    }
}

// Sum the split partials in fixed order; optionally mirror the upper triangle (conjugated) into the
// lower one so consumers may read either half.
template <typename T>
__global__ void gram_reduce_kernel(const T* __restrict__ part, int64_t split_stride, int nsplit, int ma,
                                   int mb, int mirror, T* __restrict__ G, int ldg) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)ma * mb) return;
  const int r = (int)(idx % ma), c = (int)(idx / ma);
  const bool flip = mirror && r > c;
  const int64_t src = flip ? ((int64_t)c + (int64_t)r * ma) : ((int64_t)r + (int64_t)c * ma);
  T s = zero<T>();
  for (int k = 0; k < nsplit; k++) s = add_(s, part[(int64_t)k * split_stride + src]);
  G[r + (int64_t)c * ldg] = flip ? conj_(s) : s;
}

// =====================================================================================================
// K4-K6 f64: DMMA tall NN.  grid.x = row_tiles * col_tiles (col tile fastest => CTAs sharing an S row
// block are co-resident and hit L2).
// =====================================================================================================
template <int TM, int TN, int WM, int WN, int BK, int STAGES, bool VECA, bool VECB>
__global__ void __launch_bounds__(WM* WN * 32)
    tall_nn_dmma_kernel(const double* __restrict__ S, int64_t lds, const double* __restrict__ C, int ldc,
                        double* __restrict__ Out, int64_t ldo, int64_t n, int kd, int nb, int nct,
                        double alpha, double beta) {
  constexpr int NT = WM * WN * 32;
  constexpr int LDA = TM + 4;
  constexpr int LDB = BK + 4;
  constexpr int MB = TM / WM / 8;
  constexpr int NB = TN / WN / 8;
  extern __shared__ __align__(16) double smem[];
  double* Ss = smem;                                   // [STAGES][BK][LDA]
  double* Cs = smem + (size_t)STAGES * BK * LDA;       // [STAGES][TN][LDB]

  const int ct = blockIdx.x % nct;
  const int64_t rt = blockIdx.x / nct;
  const int64_t r0 = rt * TM;
  const int c0 = ct * TN;
  const int nchunks = (kd + BK - 1) / BK;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int g = lane >> 2, t = lane & 3;

  double acc[MB][NB][2];
#pragma unroll
  for (int i = 0; i < MB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  // S tile: BK columns (k) x TM contiguous rows; rows beyond n are fixed per thread, k beyond kd only in the tail
  TileLoaderF64<BK, TM, LDA, NT, VECA> la;
  TileLoaderF64<TN, BK, LDB, NT, VECB> lb;
  la.init(S, lds, r0, 0, BK, tid);
  lb.init(C, ldc, 0, c0, nb, tid);
  const int64_t rows_valid = n - r0;
  int issued = 0, wstage = 0;
  auto issue = [&]() {
    if (issued < nchunks) {
      const int kleft = kd - issued * BK;
      if (kleft >= BK) {
        la.issue(Ss + wstage * (BK * LDA), S, rows_valid);
      } else {   // tail chunk: only the first kleft k-columns exist
        TileLoaderF64<BK, TM, LDA, NT, VECA> lt = la;
        lt.colmask = 0;
#pragma unroll
        for (int s = 0; s < lt.NSLOT; s++)
          if (tid / lt.CPC + s * lt.CSTEP < kleft) lt.colmask |= 1u << s;
        lt.issue(Ss + wstage * (BK * LDA), S, rows_valid);
      }
      lb.issue(Cs + wstage * (TN * LDB), C, (int64_t)kleft);
      la.advance((int64_t)BK * lds);
      lb.advance(BK);
    }
    issued++;
    wstage = (wstage + 1 == STAGES) ? 0 : wstage + 1;
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) issue();

  int rstage = 0;
  for (int chunk = 0; chunk < nchunks; chunk++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue();
    const double* as = Ss + rstage * (BK * LDA) + t * LDA + wm * MB * 8 + g;
    const double* bs = Cs + rstage * (TN * LDB) + (wn * NB * 8 + g) * LDB + t;
    rstage = (rstage + 1 == STAGES) ? 0 : rstage + 1;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ks++) {
      double a[MB], b[NB];
#pragma unroll
      for (int i = 0; i < MB; i++) a[i] = as[ks * 4 * LDA + i * 8];
#pragma unroll
      for (int j = 0; j < NB; j++) b[j] = bs[j * 8 * LDB + ks * 4];
#pragma unroll
      for (int i = 0; i < MB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) dmma884(acc[i][j], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < MB; i++) {
    const int64_t row = r0 + wm * MB * 8 + i * 8 + g;
    if (row >= n) continue;
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const int col = c0 + wn * NB * 8 + j * 8 + 2 * t;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        if (col + q < nb) {
          double* p = Out + row + (int64_t)(col + q) * ldo;
          double v = alpha * acc[i][j][q];
          if (beta != 0.0) v += beta * (*p);
          *p = v;
        }
      }
    }
  }
}

// k-steps [KS0, KS1) of one staged chunk of the tall NN kernels (4 columns of S per DMMA step)
template <int MB, int NB, int LDA, int LDB, int KS0, int KS1>
__device__ __forceinline__ void nn_ksteps(double (&acc)[MB][NB][2], const double* __restrict__ as,
                                          const double* __restrict__ bs) {
#pragma unroll
  for (int ks = KS0; ks < KS1; ks++) {
    double a[MB], b[NB];
#pragma unroll
    for (int i = 0; i < MB; i++) a[i] = as[ks * 4 * LDA + i * 8];
#pragma unroll
    for (int j = 0; j < NB; j++) b[j] = bs[j * 8 * LDB + ks * 4];
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
      for (int j = 0; j < NB; j++) dmma884(acc[i][j], a[i], b[j]);
  }
}
template <int MB, int NB, int LDA, int LDB>
__device__ __forceinline__ void nn_kstep_rt(double (&acc)[MB][NB][2], const double* __restrict__ as,
                                            const double* __restrict__ bs, int ks) {
  double a[MB], b[NB];
#pragma unroll
  for (int i = 0; i < MB; i++) a[i] = as[ks * 4 * LDA + i * 8];
#pragma unroll
  for (int j = 0; j < NB; j++) b[j] = bs[j * 8 * LDB + ks * 4];
#pragma unroll
  for (int i = 0; i < MB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) dmma884(acc[i][j], a[i], b[j]);
}

// Persistent version of the 128 x 128 tall NN tile (option nn_persist, default on): one CTA per SM walks the (row tile,
// column tile) items b, b + gridDim.x, ... (column tile fastest, so the CTAs that run together share an S row tile through
// L2) and the cp.async ring runs ACROSS items: while the last chunks of an item are multiplied the first chunks of the
// next item are already in flight, and its epilogue stores overlap those loads.  The one-tile-per-CTA kernel above pays
// the pipeline fill (two HBM round trips) and the drain once per 118 us of MMA work, with a single resident CTA per SM
// (registers) nothing hides them.
template <int BK, int STAGES, bool VECA, bool VECB>
__global__ void __launch_bounds__(256, 1)
    tall_nn_persist_kernel(const double* __restrict__ S, int64_t lds, const double* __restrict__ C, int ldc,
                           double* __restrict__ Out, int64_t ldo, int64_t n, int kd, int nb, int nct, int64_t nitems,
                           double alpha, double beta, int split) {
  constexpr int TM = 128, TN = 128, WM = 2, WN = 4, NT = 256;
  constexpr int LDA = TM + 4;
  constexpr int LDB = BK + 4;
  constexpr int MB = TM / WM / 8;
  constexpr int NB = TN / WN / 8;
  extern __shared__ __align__(16) double smem[];
  double* Ss = smem;                                   // [STAGES][BK][LDA]
  double* Cs = smem + (size_t)STAGES * BK * LDA;       // [STAGES][TN][LDB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int g = lane >> 2, t = lane & 3;
  const int nchunks = (kd + BK - 1) / BK;
  const int tail_ks = (kd - (nchunks - 1) * BK + 3) / 4;   // 4-column MMA steps of the last chunk

  // load cursor: item / chunk of the next copy
  int64_t l_item = blockIdx.x;
  int l_chunk = 0;
  TileLoaderF64<BK, TM, LDA, NT, VECA> la;
  TileLoaderF64<TN, BK, LDB, NT, VECB> lb;
  int64_t l_rows_valid = 0;
  auto start_item = [&](int64_t item) {
    const int ct = (int)(item % nct);
    const int64_t r0 = (item / nct) * TM;
    la.init(S, lds, r0, 0, BK, tid);
    lb.init(C, ldc, 0, ct * TN, nb, tid);
    l_rows_valid = n - r0;
  };
  if (l_item < nitems) start_item(l_item);
  int wstage = 0;
  auto issue = [&]() {
    if (l_item < nitems) {
      const int kleft = kd - l_chunk * BK;
      if (kleft >= BK) {
        la.issue(Ss + wstage * (BK * LDA), S, l_rows_valid);
      } else {   // tail chunk: only the first kleft k-columns exist
        TileLoaderF64<BK, TM, LDA, NT, VECA> lt = la;
        lt.colmask = 0;
#pragma unroll
        for (int s = 0; s < lt.NSLOT; s++)
          if (tid / lt.CPC + s * lt.CSTEP < kleft) lt.colmask |= 1u << s;
        lt.issue(Ss + wstage * (BK * LDA), S, l_rows_valid);
      }
      lb.issue(Cs + wstage * (TN * LDB), C, (int64_t)kleft);
      la.advance((int64_t)BK * lds);
      lb.advance(BK);
      if (++l_chunk == nchunks) {
        l_chunk = 0;
        l_item += gridDim.x;
        if (l_item < nitems) start_item(l_item);
      }
    }
    wstage = (wstage + 1 == STAGES) ? 0 : wstage + 1;
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) issue();

  int rstage = 0;
  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    double acc[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
      for (int j = 0; j < NB; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int chunk = 0; chunk < nchunks; chunk++) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      // the copy instructions of a chunk (address arithmetic + 16 LDGSTS per thread) keep a warp away from the DMMA
      // pipe for several hundred clocks; the two warps of a scheduler (w, w + 4) take that detour at different times
      if (warp < split) issue();
      const double* as = Ss + rstage * (BK * LDA) + t * LDA + wm * MB * 8 + g;
      const double* bs = Cs + rstage * (TN * LDB) + (wn * NB * 8 + g) * LDB + t;
      rstage = (rstage + 1 == STAGES) ? 0 : rstage + 1;
      if (chunk + 1 < nchunks || tail_ks == BK / 4) {
        nn_ksteps<MB, NB, LDA, LDB, 0, BK / 8>(acc, as, bs);
        if (warp >= split) issue();
        nn_ksteps<MB, NB, LDA, LDB, BK / 8, BK / 4>(acc, as, bs);
      } else {   // last chunk of a k extent that is not a multiple of BK: only the 4-column steps that hold data
        if (warp >= split) issue();
#pragma unroll 1
        for (int ks = 0; ks < tail_ks; ks++) nn_kstep_rt<MB, NB, LDA, LDB>(acc, as, bs, ks);
      }
    }
    const int c0 = (int)(item % nct) * TN;
    const int64_t r0 = (item / nct) * TM;
#pragma unroll
    for (int i = 0; i < MB; i++) {
      const int64_t row = r0 + wm * MB * 8 + i * 8 + g;
      if (row >= n) continue;
#pragma unroll
      for (int j = 0; j < NB; j++) {
        const int col = c0 + wn * NB * 8 + j * 8 + 2 * t;
#pragma unroll
        for (int q = 0; q < 2; q++) {
          if (col + q < nb) {
            double* p = Out + row + (int64_t)(col + q) * ldo;
            double v = alpha * acc[i][j][q];
            if (beta != 0.0) v += beta * (*p);
            *p = v;
          }
        }
      }
    }
  }
  cp_async_wait<0>();
}

// Generic SIMT tall NN (all scalar types): 64 rows x 64 cols tile, BK = 16.
template <typename T>
__global__ void __launch_bounds__(256)
    tall_nn_simt_kernel(const T* __restrict__ S, int64_t lds, const T* __restrict__ C, int ldc,
                        T* __restrict__ Out, int64_t ldo, int64_t n, int kd, int nb, int nct, T alpha,
                        T beta, int beta_zero) {
  constexpr int TS = 64, BK = 16;
  __shared__ T Ss[BK][TS + 1];
  __shared__ T Cs[BK][TS + 1];
  const int ct = blockIdx.x % nct;
  const int64_t rt = blockIdx.x / nct;
  const int64_t r0 = rt * TS;
  const int c0 = ct * TS;
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;  // tx -> rows (contiguous), ty -> cols
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = zero<T>();
  for (int k0 = 0; k0 < kd; k0 += BK) {
    // S tile: BK columns x 64 rows, rows contiguous
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int id = tid + q * 256;
      const int kk = id / TS, rr = id % TS;
      T v = zero<T>();
      if (k0 + kk < kd && r0 + rr < n) v = S[(int64_t)(k0 + kk) * lds + r0 + rr];
      Ss[kk][rr] = v;
    }
    // C tile: 64 columns x BK (k contiguous)
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int id = tid + q * 256;
      const int cc = id / BK, kk = id % BK;
      T v = zero<T>();
      if (k0 + kk < kd && c0 + cc < nb) v = C[(int64_t)(c0 + cc) * ldc + k0 + kk];
      Cs[kk][cc] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; k++) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = Ss[k][tx + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Cs[k][ty * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) fma_(acc[i][j], a[i], b[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int64_t row = r0 + tx + 16 * i;
      const int col = c0 + ty * 4 + j;
      if (row < n && col < nb) {
        T* p = Out + row + (int64_t)col * ldo;
        T v = mul_(alpha, acc[i][j]);
        if (!beta_zero) v = add_(v, mul_(beta, *p));
        *p = v;
      }
    }
}

// =====================================================================================================
// Host launchers
// =====================================================================================================
template <int TM, int TN, int WM, int WN, int BK, int STAGES, int OCC>
static int launch_gram_dmma(lb2_ctx* ctx, int64_t n, int ma, int mb, const double* A, int64_t lda,
                            const double* B, int64_t ldb, double* G, int ldg, int upper) {
  const int ntm = (ma + TM - 1) / TM, ntn = (mb + TN - 1) / TN;
  const int ntiles = upper ? ntm * (ntm + 1) / 2 : ntm * ntn;
  // OCC resident CTAs per SM; fill the machine with splits of the n range (one wave, deterministic order)
  int nsplit = (ctx->sm_count * OCC) / ntiles;
  if (nsplit < 1) nsplit = 1;
  const int64_t min_rows = 8 * BK;
  if ((int64_t)nsplit * min_rows > n) nsplit = (int)((n + min_rows - 1) / min_rows);
  if (nsplit < 1) nsplit = 1;
  int64_t rps = (n + nsplit - 1) / nsplit;
  rps = (rps + BK - 1) / BK * BK;
  nsplit = (int)((n + rps - 1) / rps);
  const bool direct = (nsplit == 1 && !upper && ldg == ma);
  const int64_t split_stride = (int64_t)ma * mb;
  double* part = direct ? G : (double*)ctx_scratch(ctx, sizeof(double) * split_stride * nsplit);
  if (!part) return -1;
  const bool vec = (lda % 2 == 0) && (ldb % 2 == 0) && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0);
  constexpr size_t smem = sizeof(double) * (size_t)STAGES * (TM + TN) * (BK + 4);
  dim3 grid(ntiles, nsplit), block(WM * WN * 32);
  if (vec) {
    auto k = gram_dmma_kernel<TM, TN, WM, WN, BK, STAGES, OCC, true>;
    LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, block, smem, ctx->stream>>>(A, lda, B, ldb, ma, mb, n, rps, upper, ntm, part, split_stride, ma);
  } else {
    auto k = gram_dmma_kernel<TM, TN, WM, WN, BK, STAGES, OCC, false>;
    LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, block, smem, ctx->stream>>>(A, lda, B, ldb, ma, mb, n, rps, upper, ntm, part, split_stride, ma);
  }
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  if (!direct) {
    const int64_t tot = (int64_t)ma * mb;
    gram_reduce_kernel<double><<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(
        part, split_stride, nsplit, ma, mb, upper, G, ldg);
    ctx->launches++;
    LB2_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

template <typename T>
static int launch_gram_simt(lb2_ctx* ctx, int64_t n, int ma, int mb, const T* A, int64_t lda, const T* B,
                            int64_t ldb, T* G, int ldg, int upper) {
  constexpr int TS = 64, BK = 16;
  const int ntm = (ma + TS - 1) / TS, ntn = (mb + TS - 1) / TS;
  const int ntiles = upper ? ntm * (ntm + 1) / 2 : ntm * ntn;
  int nsplit = (ctx->sm_count * 4) / ntiles;
  if (nsplit < 1) nsplit = 1;
  const int64_t min_rows = 16 * BK;
  if ((int64_t)nsplit * min_rows > n) nsplit = (int)((n + min_rows - 1) / min_rows);
  if (nsplit < 1) nsplit = 1;
  int64_t rps = (n + nsplit - 1) / nsplit;
  rps = (rps + BK - 1) / BK * BK;
  nsplit = (int)((n + rps - 1) / rps);
  const int64_t split_stride = (int64_t)ma * mb;
  T* part = (T*)ctx_scratch(ctx, sizeof(T) * split_stride * nsplit);
  if (!part) return -1;
  dim3 grid(ntiles, nsplit);
  gram_simt_kernel<T><<<grid, 256, 0, ctx->stream>>>(A, lda, B, ldb, ma, mb, n, rps, upper, ntm, part,
                                                     split_stride, ma);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  const int64_t tot = (int64_t)ma * mb;
  gram_reduce_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(part, split_stride, nsplit,
                                                                               ma, mb, upper, G, ldg);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}


static int launch_gram_tf32(lb2_ctx* ctx, int64_t n, int ma, int mb, const float* A, int64_t lda, const float* B,
                            int64_t ldb, float* G, int ldg, int upper) {
  constexpr int TM = 128, TN = 128, WM = 2, WN = 4, BK = 32, STAGES = 3;
  const int ntm = (ma + TM - 1) / TM, ntn = (mb + TN - 1) / TN;
  const int ntiles = upper ? ntm * (ntm + 1) / 2 : ntm * ntn;
  int nsplit = ctx->sm_count / ntiles;
  if (nsplit < 1) nsplit = 1;
  const int64_t min_rows = 8 * BK;
  if ((int64_t)nsplit * min_rows > n) nsplit = (int)((n + min_rows - 1) / min_rows);
  if (nsplit < 1) nsplit = 1;
  int64_t rps = (n + nsplit - 1) / nsplit;
  rps = (rps + BK - 1) / BK * BK;
  nsplit = (int)((n + rps - 1) / rps);
  const int64_t split_stride = (int64_t)ma * mb;
  float* part = (float*)ctx_scratch(ctx, sizeof(float) * split_stride * nsplit);
  if (!part) return -1;
  const bool vec = (lda % 4 == 0) && (ldb % 4 == 0) && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0);
  constexpr size_t smem = sizeof(float) * (size_t)STAGES * (TM + TN) * (BK + 4);
  dim3 grid(ntiles, nsplit), block(WM * WN * 32);
  if (vec) {
    auto k = gram_tf32_kernel<TM, TN, WM, WN, BK, STAGES, true>;
    LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, block, smem, ctx->stream>>>(A, lda, B, ldb, ma, mb, n, rps, upper, ntm, part, split_stride, ma);
  } else {
    auto k = gram_tf32_kernel<TM, TN, WM, WN, BK, STAGES, false>;
    LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, block, smem, ctx->stream>>>(A, lda, B, ldb, ma, mb, n, rps, upper, ntm, part, split_stride, ma);
  }
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  const int64_t tot = (int64_t)ma * mb;
  gram_reduce_kernel<float><<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(part, split_stride, nsplit, ma,
                                                                                   mb, upper, G, ldg);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

static int launch_nn_tf32(lb2_ctx* ctx, int64_t n, int kd, int nb, float alpha, const float* S, int64_t lds,
                          const float* C, int ldc, float beta, float* Out, int64_t ldo) {
  constexpr int TM = 128, TN = 128, WM = 2, WN = 4, BK = 32, STAGES = 3;
  const int nct = (nb + TN - 1) / TN;
  const int64_t nrt = (n + TM - 1) / TM;
  const bool veca = (lds % 4 == 0) && ((uintptr_t)S % 16 == 0);
  const bool vecb = (ldc % 4 == 0) && ((uintptr_t)C % 16 == 0);
  constexpr size_t smem = sizeof(float) * (size_t)STAGES * (BK * (TM + 8) + TN * (BK + 4));
  const unsigned grid = (unsigned)(nrt * nct);
  dim3 block(WM * WN * 32);
#define LB2_NNF_LAUNCH(VA, VB)                                                                        \
  {                                                                                                   \
    auto k = tall_nn_tf32_kernel<TM, TN, WM, WN, BK, STAGES, VA, VB>;                                 \
    LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    k<<<grid, block, smem, ctx->stream>>>(S, lds, C, ldc, Out, ldo, n, kd, nb, nct, alpha, beta);     \
  }
  if (veca && vecb) LB2_NNF_LAUNCH(true, true)
  else if (veca) LB2_NNF_LAUNCH(true, false)
  else if (vecb) LB2_NNF_LAUNCH(false, true)
  else LB2_NNF_LAUNCH(false, false)
#undef LB2_NNF_LAUNCH
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

static int launch_gram_zmma(lb2_ctx* ctx, int64_t n, int ma, int mb, const c64* A, int64_t lda, const c64* B,
                            int64_t ldb, c64* G, int ldg, int upper) {
  constexpr int TM = 64, TN = 64, WM = 2, WN = 2, BK = 8, STAGES = 4, OCC = 2;
  const int ntm = (ma + TM - 1) / TM, ntn = (mb + TN - 1) / TN;
  const int ntiles = upper ? ntm * (ntm + 1) / 2 : ntm * ntn;
  int nsplit = (ctx->sm_count * OCC) / ntiles;
  if (nsplit < 1) nsplit = 1;
  const int64_t min_rows = 16 * BK;
  if ((int64_t)nsplit * min_rows > n) nsplit = (int)((n + min_rows - 1) / min_rows);
  if (nsplit < 1) nsplit = 1;
  int64_t rps = (n + nsplit - 1) / nsplit;
  rps = (rps + BK - 1) / BK * BK;
  nsplit = (int)((n + rps - 1) / rps);
  const int64_t split_stride = (int64_t)ma * mb;
  c64* part = (c64*)ctx_scratch(ctx, sizeof(c64) * split_stride * nsplit);
  if (!part) return -1;
  constexpr size_t smem = sizeof(c64) * (size_t)STAGES * (TM + TN) * (BK + 4);
  auto k = gram_zmma_kernel<TM, TN, WM, WN, BK, STAGES, OCC>;
  LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<dim3(ntiles, nsplit), WM * WN * 32, smem, ctx->stream>>>(A, lda, B, ldb, ma, mb, n, rps, upper, ntm, part,
                                                              split_stride, ma);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  const int64_t tot = (int64_t)ma * mb;
  gram_reduce_kernel<c64><<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(part, split_stride, nsplit, ma, mb,
                                                                                 upper, G, ldg);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

static int launch_nn_zmma(lb2_ctx* ctx, int64_t n, int kd, int nb, c64 alpha, const c64* S, int64_t lds, const c64* C,
                          int ldc, c64 beta, c64* Out, int64_t ldo) {
  constexpr int TM = 64, TN = 64, WM = 2, WN = 2, BK = 8, STAGES = 4, OCC = 2;
  const int nct = (nb + TN - 1) / TN;
  const int64_t nrt = (n + TM - 1) / TM;
  constexpr size_t smem = sizeof(c64) * (size_t)STAGES * (BK * (TM + 2) + TN * (BK + 4));
  auto k = tall_nn_zmma_kernel<TM, TN, WM, WN, BK, STAGES, OCC>;
  LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const bool bz = (beta.re == 0.0 && beta.im == 0.0);
  k<<<(unsigned)(nrt * nct), WM * WN * 32, smem, ctx->stream>>>(S, lds, C, ldc, Out, ldo, n, kd, nb, nct, alpha, beta,
                                                               bz ? 1 : 0);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// deterministic split reduction (+ Hermitian mirror) for float partials produced outside this file (gram_tc5.cu)
int gram_reduce_f32(lb2_ctx* ctx, const float* part, int64_t split_stride, int nsplit, int ma, int mb, int mirror,
                    float* G, int ldg) {
  const int64_t tot = (int64_t)ma * mb;
  gram_reduce_kernel<float><<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(part, split_stride, nsplit, ma, mb,
                                                                                   mirror, G, ldg);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// f64 Gram on the tile x equal-split grid (gram_dmma_kernel): every CTA of an n-range runs in the same wave, so the
// operands come from HBM once and from L2 for every further tile.
int gram_tiles_f64(lb2_ctx* ctx, int64_t n, int ma, int mb, const double* A, int64_t lda, const double* B,
                   int64_t ldb, double* G, int ldg, int upper) {
  int tile = ctx->gram_tile;
  if (tile == 0) {
    // cost model: padded tile area (upper: tiles on/above the diagonal) x wave quantisation / relative speed
    const int cand[3] = {128, 96, 64};
    const int occ[3] = {1, 2, 3};
    const double speed[3] = {1.0, 0.97, 0.85};
    double best = 1e300;
    for (int i = 0; i < 3; i++) {
      const int t = cand[i];
      const int ntm = (ma + t - 1) / t, ntn = (mb + t - 1) / t;
      const int ntiles = upper ? ntm * (ntm + 1) / 2 : ntm * ntn;
      const int slots = ctx->sm_count * occ[i];
      int nsplit = slots / ntiles;
      if (nsplit < 1) nsplit = 1;
      const double waves = (double)(ntiles * nsplit + slots - 1) / slots;   // >= 1 when ntiles > slots
      const double fill = (double)ntiles * nsplit / (std::ceil(waves) * slots);
      const double cost = (double)ntiles * t * t / (speed[i] * fill);
      if (cost < best) { best = cost; tile = t; }
    }
  }
  if (tile == 128) return launch_gram_dmma<128, 128, 2, 4, 16, 4, 1>(ctx, n, ma, mb, A, lda, B, ldb, G, ldg, upper);
  if (tile == 96) return launch_gram_dmma<96, 96, 2, 4, 16, 3, 2>(ctx, n, ma, mb, A, lda, B, ldb, G, ldg, upper);
  return launch_gram_dmma<64, 64, 2, 2, 16, 3, 3>(ctx, n, ma, mb, A, lda, B, ldb, G, ldg, upper);
}

// G = A^H B.  upper != 0: A and B span the same columns of a Hermitian product (G = G^H): only tiles on
// or above the diagonal are computed and the result is mirrored, so all of G is valid on return.
template <typename T>
int gram(lb2_ctx* ctx, int64_t n, int ma, int mb, const T* A, int64_t lda, const T* B, int64_t ldb, T* G,
         int ldg, int upper) {
  if (ma <= 0 || mb <= 0) return 0;
  if (upper && ma != mb) return -2;
  if (n <= 0) {
    LB2_CUDA_OK(cudaMemset2DAsync(G, sizeof(T) * ldg, 0, sizeof(T) * ma, mb, ctx->stream));
    return 0;
  }
  if constexpr (std::is_same<T, double>::value) {
    if (!ctx->force_simt) {
      if (lb2_gram_i8_on(ctx, n, ma, mb)) {   // tcgen05 kind::i8 on an Ozaki split (gram_i8.cu)
        const int rc = gram_i8_f64(ctx, n, ma, mb, A, lda, B, ldb, G, ldg, upper);
        if (rc != -100) return rc;
      }
      ctx->oz_tag_ptr = nullptr;   // this product leaves no slices behind: a projection that follows must not find older ones
      // work-list kernel (gram_wl.cu) for Hermitian products; forced for every shape with gram_wl = 1
      if (ctx->gram_wl == 1 || (ctx->gram_wl < 0 && ctx->gram_tile == 0 && upper && n >= 4096))
        return gram_wl_f64(ctx, n, ma, mb, A, lda, B, ldb, G, ldg, upper);
      return gram_tiles_f64(ctx, n, ma, mb, A, lda, B, ldb, G, ldg, upper);
    }
  }
  if constexpr (std::is_same<T, c64>::value) {
    if (!ctx->force_simt) return launch_gram_zmma(ctx, n, ma, mb, A, lda, B, ldb, G, ldg, upper);
  }
  if constexpr (std::is_same<T, float>::value) {
    if (!ctx->force_simt) {
      if (ctx->gram_tc5 != 0 && n >= 1024) {   // tcgen05 / TMEM path (gram_tc5.cu); -100 = operands not 16-byte aligned
        const int rc = gram_tc5_f32(ctx, n, ma, mb, A, lda, B, ldb, G, ldg, upper);
        if (rc != -100) return rc;
      }
      return launch_gram_tf32(ctx, n, ma, mb, A, lda, B, ldb, G, ldg, upper);
    }
  }
  return launch_gram_simt<T>(ctx, n, ma, mb, A, lda, B, ldb, G, ldg, upper);
}

// Column-block products of the cached-Gram pass: G0[0:m,0:nw] = S^H W0 and, when W1 != null,
// G1[0:m,0:nw] = S^H W1.  tri_c0 >= 0: rows tri_c0.. of the results are a Hermitian nw x nw block; its strictly-lower
// part MAY be left unwritten (callers read the upper part only).  f64: one work-list launch for both products
// (gram_wl.cu); other types: one rectangular product each.
template <typename T>
int gram_cols(lb2_ctx* ctx, int64_t n, int m, int nw, const T* S, int64_t lds, const T* W0, int64_t ldw0, T* G0,
              int ldg0, const T* W1, int64_t ldw1, T* G1, int ldg1, int tri_c0) {
  if (m <= 0 || nw <= 0) return 0;
  if constexpr (std::is_same<T, double>::value) {
    if (!ctx->force_simt && lb2_gram_i8_on(ctx, n, m, nw)) {   // tcgen05 kind::i8 on an Ozaki split (gram_i8.cu)
      const int rc = gram_cols_i8_f64(ctx, n, m, nw, S, lds, W0, ldw0, G0, ldg0, W1, ldw1, G1, ldg1, tri_c0);
      if (rc != -100) return rc;
    }
    ctx->oz_tag_ptr = nullptr;   // (see gram: no stale slices for the projections of this pass)
    if (!ctx->force_simt && ctx->gram_wl != 0 && ctx->gram_tile == 0 && n >= 4096)
      return gram_wl_cols_f64(ctx, n, m, nw, S, lds, W0, ldw0, G0, ldg0, W1, ldw1, G1, ldg1, tri_c0);
  }
  if (int rc = gram<T>(ctx, n, m, nw, S, lds, W0, ldw0, G0, ldg0, 0)) return rc;
  if (W1 && G1) return gram<T>(ctx, n, m, nw, S, lds, W1, ldw1, G1, ldg1, 0);
  return 0;
}

template <int TM, int TN, int WM, int WN, int BK, int STAGES>
static int launch_nn_dmma(lb2_ctx* ctx, int64_t n, int kd, int nb, double alpha, const double* S,
                          int64_t lds, const double* C, int ldc, double beta, double* Out, int64_t ldo) {
  const int nct = (nb + TN - 1) / TN;
  const int64_t nrt = (n + TM - 1) / TM;
  const bool veca = (lds % 2 == 0) && ((uintptr_t)S % 16 == 0);
  const bool vecb = (ldc % 2 == 0) && ((uintptr_t)C % 16 == 0);
  constexpr size_t smem = sizeof(double) * (size_t)STAGES * (BK * (TM + 4) + TN * (BK + 4));
  const unsigned grid = (unsigned)(nrt * nct);
  dim3 block(WM * WN * 32);
#define LB2_NN_LAUNCH(VA, VB)                                                                         \
  {                                                                                                   \
    auto k = tall_nn_dmma_kernel<TM, TN, WM, WN, BK, STAGES, VA, VB>;                                 \
    LB2_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    k<<<grid, block, smem, ctx->stream>>>(S, lds, C, ldc, Out, ldo, n, kd, nb, nct, alpha, beta);     \
  }
  if (veca && vecb) LB2_NN_LAUNCH(true, true)
  else if (veca) LB2_NN_LAUNCH(true, false)
  else if (vecb) LB2_NN_LAUNCH(false, true)
  else LB2_NN_LAUNCH(false, false)
#undef LB2_NN_LAUNCH
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// Out = alpha * S C + beta * Out.  Out must not alias S (the drivers ping-pong slabs instead).
template <typename T>
int tall_nn(lb2_ctx* ctx, int64_t n, int kd, int nb, T alpha, const T* S, int64_t lds, const T* C, int ldc,
            T beta, T* Out, int64_t ldo) {
  if (n <= 0 || nb <= 0) return 0;
  if constexpr (std::is_same<T, double>::value) {
    // int8 tensor path (gram_i8.cu): plain products always; updates (alpha, beta general) when the slices of S are already there
    if (!ctx->force_simt && kd > 0 && ctx->nn_i8 != 0 &&
        ((alpha == 1.0 && beta == 0.0 && lb2_gram_i8_on(ctx, n, kd, nb)) || (lb2_gram_i8_on(ctx, n) && oz_slices_cached(ctx, S, n, kd, lds)))) {
      const int rc = tall_nn_i8_f64(ctx, n, kd, nb, alpha, S, lds, C, ldc, beta, Out, ldo);
      if (rc != -100) return rc;
    }
    if (!ctx->force_simt && kd > 0) {
      // 128-wide column tiles plus one narrower remainder tile (32/64/96) so that padding stays below 32 columns
      const int forced = ctx->nn_tile;
      const int nfull = (forced == 0 || forced == 128) ? nb / 128 : 0;
      int rc = 0;
      if (nfull > 0 && ctx->nn_persist != 0 && ctx->nn_bk != 16 && ctx->nn_warps != 16) {
        // persistent 128 x 128 tiles: one CTA per SM, cp.async ring running across the tiles
        const int nct = nfull;
        const int64_t nitems = ((n + 127) / 128) * nct;
        const bool veca = (lds % 2 == 0) && ((uintptr_t)S % 16 == 0);
        const bool vecb = (ldc % 2 == 0) && ((uintptr_t)C % 16 == 0);
        constexpr size_t smem = sizeof(double) * (size_t)3 * (32 * (128 + 4) + 128 * (32 + 4));
        const unsigned grid = (unsigned)std::min<int64_t>(nitems, ctx->sm_count);
#define LB2_NNP_LAUNCH(VA, VB)                                                                          \
  {                                                                                                     \
    auto kp = tall_nn_persist_kernel<32, 3, VA, VB>;                                                    \
    LB2_CUDA_OK(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    kp<<<grid, 256, smem, ctx->stream>>>(S, lds, C, ldc, Out, ldo, n, kd, nfull * 128, nct, nitems, alpha, beta, ctx->nn_stagger ? 4 : 8); \
  }
        if (veca && vecb) LB2_NNP_LAUNCH(true, true)
        else if (veca) LB2_NNP_LAUNCH(true, false)
        else if (vecb) LB2_NNP_LAUNCH(false, true)
        else LB2_NNP_LAUNCH(false, false)
#undef LB2_NNP_LAUNCH
        ctx->launches++;
        LB2_CUDA_OK(cudaGetLastError());
      } else if (nfull > 0 && ctx->nn_warps == 16) {
        // experiment: 16 warps (4 per scheduler) with 32 x 32 warp tiles instead of 8 warps with 64 x 32
        rc = launch_nn_dmma<128, 128, 4, 4, 32, 3>(ctx, n, kd, nfull * 128, alpha, S, lds, C, ldc, beta, Out, ldo);
      } else if (nfull > 0) {
        if (ctx->nn_bk != 16)
          rc = launch_nn_dmma<128, 128, 2, 4, 32, 3>(ctx, n, kd, nfull * 128, alpha, S, lds, C, ldc, beta, Out, ldo);
        else
          rc = launch_nn_dmma<128, 128, 2, 4, 16, 4>(ctx, n, kd, nfull * 128, alpha, S, lds, C, ldc, beta, Out, ldo);
      }
      const int rem = nb - nfull * 128;
      if (rc == 0 && rem > 0) {
        const double* Cr = C + (int64_t)nfull * 128 * ldc;
        double* Or = Out + (int64_t)nfull * 128 * ldo;
        // remainder tile widths in steps of 16 columns (16/48/80 run 4-warp CTAs, two per SM)
        const int w = forced ? forced : (rem + 15) / 16 * 16;
        if (w <= 16) rc = launch_nn_dmma<128, 16, 4, 1, 16, 4>(ctx, n, kd, rem, alpha, S, lds, Cr, ldc, beta, Or, ldo);
        else if (w <= 32) rc = launch_nn_dmma<128, 32, 8, 1, 16, 4>(ctx, n, kd, rem, alpha, S, lds, Cr, ldc, beta, Or, ldo);
        else if (w <= 48) rc = launch_nn_dmma<128, 48, 2, 2, 16, 4>(ctx, n, kd, rem, alpha, S, lds, Cr, ldc, beta, Or, ldo);
        else if (w <= 64) rc = launch_nn_dmma<128, 64, 4, 2, 16, 4>(ctx, n, kd, rem, alpha, S, lds, Cr, ldc, beta, Or, ldo);
        else if (w <= 80) rc = launch_nn_dmma<128, 80, 2, 2, 16, 4>(ctx, n, kd, rem, alpha, S, lds, Cr, ldc, beta, Or, ldo);
        else if (w <= 96) rc = launch_nn_dmma<128, 96, 4, 2, 16, 4>(ctx, n, kd, rem, alpha, S, lds, Cr, ldc, beta, Or, ldo);
        else rc = launch_nn_dmma<128, 128, 2, 4, 16, 4>(ctx, n, kd, rem, alpha, S, lds, Cr, ldc, beta, Or, ldo);
      }
      return rc;
    }
  }
  if constexpr (std::is_same<T, c64>::value) {
    if (!ctx->force_simt && kd > 0) return launch_nn_zmma(ctx, n, kd, nb, alpha, S, lds, C, ldc, beta, Out, ldo);
  }
  if constexpr (std::is_same<T, float>::value) {
    if (!ctx->force_simt && kd > 0) {
      if (ctx->gram_tc5 != 0 && n >= 1024) {   // tcgen05 / TMEM path (nn_tc5.cu); -100 = S not 16-byte aligned
        const int rc = nn_tc5_f32(ctx, n, kd, nb, alpha, S, lds, C, ldc, beta, Out, ldo);
        if (rc != -100) return rc;
      }
      return launch_nn_tf32(ctx, n, kd, nb, alpha, S, lds, C, ldc, beta, Out, ldo);
    }
  }
  const int nct = (nb + 63) / 64;
  const int64_t nrt = (n + 63) / 64;
  const bool bz = (real_(beta) == real_t<T>(0)) && (abs2_(beta) == real_t<T>(0));
  tall_nn_simt_kernel<T><<<(unsigned)(nrt * nct), 256, 0, ctx->stream>>>(S, lds, C, ldc, Out, ldo, n, kd, nb,
                                                                       nct, alpha, beta, bz ? 1 : 0);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

#define LB2_INST(T)                                                                                   \
  template int gram<T>(lb2_ctx*, int64_t, int, int, const T*, int64_t, const T*, int64_t, T*, int, int); \
  template int tall_nn<T>(lb2_ctx*, int64_t, int, int, T, const T*, int64_t, const T*, int, T, T*, int64_t); \
  template int gram_cols<T>(lb2_ctx*, int64_t, int, int, const T*, int64_t, const T*, int64_t, T*, int, const T*, int64_t, T*, int, int);
LB2_INST(float)
LB2_INST(double)
LB2_INST(c32)
LB2_INST(c64)
#undef LB2_INST

}  // namespace lb2

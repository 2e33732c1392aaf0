// lobpcg_b200/csrc/common.cuh — scalar traits, error handling and small device helpers shared by every
// kernel file.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

namespace lb2 {

// ---------------------------------------------------------------------------------------------------
// Scalars.  The reference instantiates everything for f32/f64/c32/c64 (include/lobpcg/types.h:11-20);
// Cx<R> is layout-compatible with C99 `R _Complex` (re, im interleaved).
// ---------------------------------------------------------------------------------------------------
template <typename R>
struct alignas(2 * sizeof(R)) Cx {
  R re, im;
};
using c32 = Cx<float>;
using c64 = Cx<double>;

template <typename T> struct Sc;
template <> struct Sc<float>  { using real = float;  static constexpr bool cplx = false; static constexpr char prefix = 's'; };
template <> struct Sc<double> { using real = double; static constexpr bool cplx = false; static constexpr char prefix = 'd'; };
template <> struct Sc<c32>    { using real = float;  static constexpr bool cplx = true;  static constexpr char prefix = 'c'; };
template <> struct Sc<c64>    { using real = double; static constexpr bool cplx = true;  static constexpr char prefix = 'z'; };
template <typename T> using real_t = typename Sc<T>::real;

#define LB2_HD __host__ __device__ __forceinline__

LB2_HD float  zero_of(float)  { return 0.f; }
LB2_HD double zero_of(double) { return 0.0; }
template <typename R> LB2_HD Cx<R> zero_of(Cx<R>) { return Cx<R>{R(0), R(0)}; }
template <typename T> LB2_HD T zero() { return zero_of(T{}); }

LB2_HD float  from_real(float r, float)   { return r; }
LB2_HD double from_real(double r, double) { return r; }
template <typename R> LB2_HD Cx<R> from_real(R r, Cx<R>) { return Cx<R>{r, R(0)}; }
template <typename T> LB2_HD T make(real_t<T> r) { return from_real(r, T{}); }

LB2_HD float  conj_(float a)  { return a; }
LB2_HD double conj_(double a) { return a; }
template <typename R> LB2_HD Cx<R> conj_(Cx<R> a) { return Cx<R>{a.re, -a.im}; }

LB2_HD float  real_(float a)  { return a; }
LB2_HD double real_(double a) { return a; }
template <typename R> LB2_HD R real_(Cx<R> a) { return a.re; }

LB2_HD float  abs2_(float a)  { return a * a; }
LB2_HD double abs2_(double a) { return a * a; }
template <typename R> LB2_HD R abs2_(Cx<R> a) { return a.re * a.re + a.im * a.im; }

LB2_HD float  add_(float a, float b)   { return a + b; }
LB2_HD double add_(double a, double b) { return a + b; }
template <typename R> LB2_HD Cx<R> add_(Cx<R> a, Cx<R> b) { return Cx<R>{a.re + b.re, a.im + b.im}; }

LB2_HD float  sub_(float a, float b)   { return a - b; }
LB2_HD double sub_(double a, double b) { return a - b; }
template <typename R> LB2_HD Cx<R> sub_(Cx<R> a, Cx<R> b) { return Cx<R>{a.re - b.re, a.im - b.im}; }

LB2_HD float  mul_(float a, float b)   { return a * b; }
LB2_HD double mul_(double a, double b) { return a * b; }
template <typename R> LB2_HD Cx<R> mul_(Cx<R> a, Cx<R> b) {
  return Cx<R>{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
// scale by a real
LB2_HD float  rscale_(float a, float s)   { return a * s; }
LB2_HD double rscale_(double a, double s) { return a * s; }
template <typename R> LB2_HD Cx<R> rscale_(Cx<R> a, R s) { return Cx<R>{a.re * s, a.im * s}; }

// acc += a*b
LB2_HD void fma_(float& acc, float a, float b)    { acc = fmaf(a, b, acc); }
LB2_HD void fma_(double& acc, double a, double b) { acc = fma(a, b, acc); }
template <typename R> LB2_HD void fma_(Cx<R>& acc, Cx<R> a, Cx<R> b) {
  acc.re += a.re * b.re - a.im * b.im;
  acc.im += a.re * b.im + a.im * b.re;
}
// acc += conj(a)*b
LB2_HD void fmac_(float& acc, float a, float b)    { acc = fmaf(a, b, acc); }
LB2_HD void fmac_(double& acc, double a, double b) { acc = fma(a, b, acc); }
template <typename R> LB2_HD void fmac_(Cx<R>& acc, Cx<R> a, Cx<R> b) {
  acc.re += a.re * b.re + a.im * b.im;
  acc.im += a.re * b.im - a.im * b.re;
}

// ---------------------------------------------------------------------------------------------------
// Errors: the C ABI never throws (SURVEY §8b "Errors"); failures print to stderr and return non-zero.
// ---------------------------------------------------------------------------------------------------
#define LB2_CUDA_OK(expr)                                                                           \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      fprintf(stderr, "lobpcg_b200: CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(_e),         \
              __FILE__, __LINE__, #expr);                                                           \
      return -1;                                                                                    \
    }                                                                                               \
  } while (0)

#define LB2_CUDA_OKV(expr)                                                                          \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      fprintf(stderr, "lobpcg_b200: CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(_e),         \
              __FILE__, __LINE__, #expr);                                                           \
    }                                                                                               \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// cp.async (LDGSTS) with zero-fill: copies `src_bytes` (<= BYTES) from global and zero-fills the rest.
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(void* smem_dst, const void* gmem_src, int src_bytes) {
  const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  if constexpr (BYTES == 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(src_bytes));
  } else if constexpr (BYTES == 8) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(src_bytes));
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(src_bytes));
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// FP64 tensor-core tile: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4 (the native shape on
// sm_100a; the m16n8k{4,8,16} PTX shapes are decomposed into this one by ptxas).
// Fragment ownership (lane = 4*g + t): a = A[g][t], b = B[t][g], c[0..1] = C[g][2t..2t+1].
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// TF32 tensor-core tile (legacy mma.sync path; SASS HMMA.1688.F32.TF32): D(16x8) += A(16x8,row) * B(8x8,col).
// lane = 4*g + t: a0=A[g][t] a1=A[g+8][t] a2=A[g][t+4] a3=A[g+8][t+4]; b0=B[t][g] b1=B[t+4][g];
// c0=C[g][2t] c1=C[g][2t+1] c2=C[g+8][2t] c3=C[g+8][2t+1].
__device__ __forceinline__ void mma_tf32_1688(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// split an fp32 value into tf32 hi + tf32 lo (x ~= hi + lo to ~2^-22 relative): 3xTF32 keeps fp32-level accuracy
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(lo) : "f"(r));
}

}  // namespace lb2

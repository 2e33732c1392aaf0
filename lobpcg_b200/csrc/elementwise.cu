// lobpcg_b200/csrc/elementwise.cu — HBM-bound streaming kernels of the hot path.
//
//   K7/K8  fused residual + column norms   W = AX - BX diag(lambda), ||W[:,j]||^2 in the same pass
//          replaces memcpy + k axpy (src/residual/residual_impl.inc:38-57) and k nrm2 (:88-98).
//   K8     column / Frobenius sums of squares (nrm2 call sites in src/ortho/ortho_drop_impl.inc:62-117,
//          src/core/lobpcg_impl.inc:93, src/residual/estimate_norm_impl.inc:43-49).
//   K10    counter-based uniform fill (replaces libc rand(), estimate_norm_impl.inc:19-35) and scalings.
//
// Reductions are two-stage and deterministic: grid (row chunks, columns) writes one partial per CTA,
// a finishing kernel sums the chunks of each column in order.
#include <algorithm>
#include "common.cuh"
#include "context.h"
#include "kernels.h"

namespace lb2 {

static constexpr int EW_THREADS = 256;
static constexpr int EW_UNROLL = 4;

__host__ __device__ inline int ew_chunks(int64_t n, int nc, int sm_count) {
  // enough CTAs to fill the machine ~4x, at least 2048 rows per CTA
  int64_t per_col = (4LL * sm_count * 8 + nc - 1) / nc;
  int64_t maxc = (n + 2047) / 2048;
  if (per_col > maxc) per_col = maxc;
  if (per_col < 1) per_col = 1;
  return (int)per_col;
}

template <typename R>
__device__ __forceinline__ R block_sum(R v) {
  __shared__ R red[EW_THREADS / 32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  R s = 0;
  if (w == 0) {
    s = (lane < EW_THREADS / 32) ? red[lane] : R(0);
    s = warp_sum(s);
  }
  return s;  // valid in warp 0
}

template <typename T, bool WRITE, bool HAVE_B>
__global__ void __launch_bounds__(EW_THREADS)
    residual_kernel(int64_t n, int64_t rows_per_chunk, const T* __restrict__ AX, int64_t ldax,
                    const T* __restrict__ BX, int64_t ldbx, const real_t<T>* __restrict__ lambda,
                    T* __restrict__ W, int64_t ldw, real_t<T>* __restrict__ partial, int nchunks) {
  using R = real_t<T>;
  const int j = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk;
  const int64_t r1 = min(n, r0 + rows_per_chunk);
  const R lam = HAVE_B ? lambda[j] : R(0);
  const T* ax = AX + (int64_t)j * ldax;
  const T* bx = HAVE_B ? BX + (int64_t)j * ldbx : nullptr;
  T* w = WRITE ? W + (int64_t)j * ldw : nullptr;
  R s = 0;
  for (int64_t i = r0 + threadIdx.x; i < r1; i += (int64_t)EW_THREADS * EW_UNROLL) {
    T a[EW_UNROLL], b[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; u++) {
      const int64_t ii = i + (int64_t)u * EW_THREADS;
      a[u] = zero<T>();
      b[u] = zero<T>();
      if (ii < r1) {
        a[u] = ax[ii];
        if (HAVE_B) b[u] = bx[ii];
      }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; u++) {
      const int64_t ii = i + (int64_t)u * EW_THREADS;
      if (ii < r1) {
        T v = HAVE_B ? sub_(a[u], rscale_(b[u], lam)) : a[u];
        s += abs2_(v);
        if (WRITE) w[ii] = v;
      }
    }
  }
  s = block_sum<R>(s);
  if (partial && threadIdx.x == 0) partial[(int64_t)j * nchunks + blockIdx.x] = s;
}

template <typename R>
__global__ void finish_sums_kernel(const R* __restrict__ partial, int nchunks, int nc, R* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nc) return;
  R s = 0;
  for (int c = 0; c < nchunks; c++) s += partial[(int64_t)j * nchunks + c];
  out[j] = s;
}

template <typename T>
int residual(lb2_ctx* ctx, int64_t n, int nc, const T* AX, int64_t ldax, const T* BX, int64_t ldbx,
             const real_t<T>* lambda, T* W, int64_t ldw, real_t<T>* sumsq) {
  using R = real_t<T>;
  if (n <= 0 || nc <= 0) return 0;
  const int nchunks = ew_chunks(n, nc, ctx->sm_count);
  int64_t rpc = (n + nchunks - 1) / nchunks;
  R* partial = nullptr;
  if (sumsq) {
    partial = (R*)ctx_scratch(ctx, sizeof(R) * (size_t)nchunks * nc);
    if (!partial) return -1;
  }
  dim3 grid(nchunks, nc);
  const bool haveb = (BX != nullptr);
  if (W) {
    if (haveb) residual_kernel<T, true, true><<<grid, EW_THREADS, 0, ctx->stream>>>(n, rpc, AX, ldax, BX, ldbx, lambda, W, ldw, partial, nchunks);
    else residual_kernel<T, true, false><<<grid, EW_THREADS, 0, ctx->stream>>>(n, rpc, AX, ldax, BX, ldbx, lambda, W, ldw, partial, nchunks);
  } else {
    if (haveb) residual_kernel<T, false, true><<<grid, EW_THREADS, 0, ctx->stream>>>(n, rpc, AX, ldax, BX, ldbx, lambda, W, ldw, partial, nchunks);
    else residual_kernel<T, false, false><<<grid, EW_THREADS, 0, ctx->stream>>>(n, rpc, AX, ldax, BX, ldbx, lambda, W, ldw, partial, nchunks);
  }
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  if (sumsq) {
    finish_sums_kernel<R><<<(nc + 127) / 128, 128, 0, ctx->stream>>>(partial, nchunks, nc, sumsq);
    ctx->launches++;
    LB2_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

// Residual norms plus the drift monitor of the cached-Gram pass (solver.cu): for column j
//   out[j]        = ||AX_j - lambda_j BX_j||^2
//   out[nc + j]   = Re(x_j^H AX_j)      (Rayleigh quotient numerator: must equal the Ritz value theta_j)
//   out[2 nc + j] = Re(x_j^H BX_j)      (must equal 1)
// X, AX (and BX) are streamed once; BX == nullptr means B = I (BX = X: two streams, same as the plain norm pass).
template <typename T, bool HAVE_B>
__global__ void __launch_bounds__(EW_THREADS)
    residual_monitor_kernel(int64_t n, int64_t rows_per_chunk, const T* __restrict__ X, int64_t ldx,
                            const T* __restrict__ AX, int64_t ldax, const T* __restrict__ BX, int64_t ldbx,
                            const real_t<T>* __restrict__ lambda, real_t<T>* __restrict__ partial, int nchunks, int nc) {
  using R = real_t<T>;
  const int j = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk;
  const int64_t r1 = min(n, r0 + rows_per_chunk);
  const R lam = lambda[j];
  const T* x = X + (int64_t)j * ldx;
  const T* ax = AX + (int64_t)j * ldax;
  const T* bx = HAVE_B ? BX + (int64_t)j * ldbx : x;
  R s0 = 0, s1 = 0, s2 = 0;
  for (int64_t i = r0 + threadIdx.x; i < r1; i += (int64_t)EW_THREADS * EW_UNROLL) {
    T a[EW_UNROLL], b[EW_UNROLL], xv[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; u++) {
      const int64_t ii = i + (int64_t)u * EW_THREADS;
      a[u] = b[u] = xv[u] = zero<T>();
      if (ii < r1) {
        a[u] = ax[ii];
        xv[u] = x[ii];
        b[u] = HAVE_B ? bx[ii] : xv[u];
      }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; u++) {
      const T v = sub_(a[u], rscale_(b[u], lam));
      s0 += abs2_(v);
      T d1 = zero<T>(), d2 = zero<T>();
      fmac_(d1, xv[u], a[u]);
      fmac_(d2, xv[u], b[u]);
      s1 += real_(d1);
      s2 += real_(d2);
    }
  }
  s0 = block_sum<R>(s0);
  __syncthreads();
  s1 = block_sum<R>(s1);
  __syncthreads();
  s2 = block_sum<R>(s2);
  if (threadIdx.x == 0) {
    partial[((int64_t)j) * nchunks + blockIdx.x] = s0;
    partial[((int64_t)(nc + j)) * nchunks + blockIdx.x] = s1;
    partial[((int64_t)(2 * nc + j)) * nchunks + blockIdx.x] = s2;
  }
}

template <typename T>
int residual_monitor(lb2_ctx* ctx, int64_t n, int nc, const T* X, int64_t ldx, const T* AX, int64_t ldax, const T* BX,
                     int64_t ldbx, const real_t<T>* lambda, real_t<T>* out3) {
  using R = real_t<T>;
  if (n <= 0 || nc <= 0) return 0;
  const int nchunks = ew_chunks(n, nc, ctx->sm_count);
  const int64_t rpc = (n + nchunks - 1) / nchunks;
  R* partial = (R*)ctx_scratch(ctx, sizeof(R) * (size_t)nchunks * nc * 3);
  if (!partial) return -1;
  dim3 grid(nchunks, nc);
  if (BX) residual_monitor_kernel<T, true><<<grid, EW_THREADS, 0, ctx->stream>>>(n, rpc, X, ldx, AX, ldax, BX, ldbx, lambda, partial, nchunks, nc);
  else residual_monitor_kernel<T, false><<<grid, EW_THREADS, 0, ctx->stream>>>(n, rpc, X, ldx, AX, ldax, BX, ldbx, lambda, partial, nchunks, nc);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  finish_sums_kernel<R><<<(3 * nc + 127) / 128, 128, 0, ctx->stream>>>(partial, nchunks, 3 * nc, out3);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T>
int col_sumsq(lb2_ctx* ctx, int64_t n, int nc, const T* X, int64_t ldx, real_t<T>* sumsq) {
  return residual<T>(ctx, n, nc, X, ldx, nullptr, 0, nullptr, nullptr, 0, sumsq);
}

template <typename R>
__global__ void sum_reals_kernel(int nc, const R* __restrict__ v, R* __restrict__ out) {
  R s = 0;
  for (int i = threadIdx.x; i < nc; i += EW_THREADS) s += v[i];
  s = block_sum<R>(s);
  if (threadIdx.x == 0) out[0] = s;
}
template <typename R>
int sum_reals(lb2_ctx* ctx, int nc, const R* v, R* out) {
  sum_reals_kernel<R><<<1, EW_THREADS, 0, ctx->stream>>>(nc, v, out);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- splitmix64 uniform fill (bit-identical to lobpcg_b200/problems.py::splitmix_uniform) -------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + (idx + 1ULL) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ void uni(uint64_t seed, uint64_t c, double& o) {
  o = (double)(splitmix64(seed, c) >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}
__device__ __forceinline__ void uni(uint64_t seed, uint64_t c, float& o) {
  o = (float)(splitmix64(seed, c) >> 40) * (1.0f / 16777216.0f) - 0.5f;
}
template <typename R>
__device__ __forceinline__ void uni(uint64_t seed, uint64_t c, Cx<R>& o) {
  uni(seed, 2 * c, o.re);
  uni(seed, 2 * c + 1, o.im);
}

template <typename T>
__global__ void fill_uniform_kernel(int64_t n, int nc, T* __restrict__ X, int64_t ldx, uint64_t seed,
                                    int64_t n_global, int64_t row0) {
  const int64_t tot = n * nc;
  for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < tot;
       id += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = id % n, j = id / n;
    T v;
    uni(seed, (uint64_t)j * (uint64_t)n_global + (uint64_t)(row0 + i), v);
    X[i + j * ldx] = v;
  }
}
template <typename T>
int fill_uniform(lb2_ctx* ctx, int64_t n, int nc, T* X, int64_t ldx, uint64_t seed, int64_t n_global,
                 int64_t row0) {
  if (n <= 0 || nc <= 0) return 0;
  fill_uniform_kernel<T><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(n, nc, X, ldx, seed, n_global, row0);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T>
__global__ void scale_cols_kernel(int64_t n, int nc, T* __restrict__ X, int64_t ldx,
                                  const real_t<T>* __restrict__ s, real_t<T> s0) {
  const int j = blockIdx.y;
  const real_t<T> f = s ? s[j] : s0;
  T* x = X + (int64_t)j * ldx;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = rscale_(x[i], f);
}
template <typename T>
int scale_cols(lb2_ctx* ctx, int64_t n, int nc, T* X, int64_t ldx, const real_t<T>* s, real_t<T> s0) {
  if (n <= 0 || nc <= 0) return 0;
  int gx = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (gx < 1) gx = 1;
  scale_cols_kernel<T><<<dim3(gx, nc), 256, 0, ctx->stream>>>(n, nc, X, ldx, s, s0);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T>
__global__ void normalize_by_kernel(int64_t n, T* __restrict__ x, const real_t<T>* __restrict__ sumsq) {
  using R = real_t<T>;
  const R ss = sumsq[0];
  if (!(ss > R(0))) return;
  const R f = R(1) / sqrt(ss);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = rscale_(x[i], f);
}
template <typename T>
int normalize_by(lb2_ctx* ctx, int64_t n, T* x, const real_t<T>* sumsq) {
  int gx = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (gx < 1) gx = 1;
  normalize_by_kernel<T><<<gx, 256, 0, ctx->stream>>>(n, x, sumsq);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T>
int copy_block(lb2_ctx* ctx, int64_t n, int nc, const T* X, int64_t ldx, T* Y, int64_t ldy) {
  if (n <= 0 || nc <= 0) return 0;
  LB2_CUDA_OK(cudaMemcpy2DAsync(Y, sizeof(T) * ldy, X, sizeof(T) * ldx, sizeof(T) * n, nc,
                                cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Chebyshev polynomial preconditioner (built-in T operator): the two vector updates of the
// three-term recurrence, fused so that one step touches each block once.
//   init:    d = x / theta, y = d
//   update:  r_out = r_in - Ad;  d_out = c1 d_in + c2 r_out;  y += d_out   (r_in is x on the first step; d ping-pongs so that
//            a neighbour rank may still be reading d_in's halo plane while this rank already runs the update)
// HBM-bound: 3 (init) / 7 (update) block streams of n x nc scalars.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    cheb_init_kernel(int64_t n, const T* __restrict__ X, int64_t ldx, T* __restrict__ D, T* __restrict__ Y, int64_t ldy,
                     int64_t ldw, real_t<T> inv_theta) {
  const int j = blockIdx.y;
  const T* x = X + (int64_t)j * ldx;
  T* d = D + (int64_t)j * ldw;
  T* y = Y + (int64_t)j * ldy;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const T v = rscale_(x[i], inv_theta);
    d[i] = v;
    y[i] = v;
  }
}
template <typename T>
__global__ void __launch_bounds__(256)
    cheb_update_kernel(int64_t n, const T* __restrict__ AD, const T* __restrict__ Rin, int64_t ldrin,
                       T* __restrict__ Rout, const T* __restrict__ Din, T* __restrict__ Dout, int64_t ldw,
                       T* __restrict__ Y, int64_t ldy, real_t<T> c1, real_t<T> c2, int write_r) {
  const int j = blockIdx.y;
  const T* ad = AD + (int64_t)j * ldw;
  const T* rin = Rin + (int64_t)j * ldrin;
  T* rout = Rout + (int64_t)j * ldw;
  const T* din = Din + (int64_t)j * ldw;
  T* dout = Dout + (int64_t)j * ldw;
  T* y = Y + (int64_t)j * ldy;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const T r = sub_(rin[i], ad[i]);
    const T dn = add_(rscale_(din[i], c1), rscale_(r, c2));
    if (write_r) rout[i] = r;
    dout[i] = dn;
    y[i] = add_(y[i], dn);
  }
}
template <typename T>
int cheb_init(lb2_ctx* ctx, int64_t n, int nc, const T* X, int64_t ldx, T* D, T* Y, int64_t ldy, int64_t ldw,
              real_t<T> inv_theta) {
  if (n <= 0 || nc <= 0) return 0;
  int gx = (int)std::min<int64_t>((n + 256 * 4 - 1) / (256 * 4), 1 << 20);
  cheb_init_kernel<T><<<dim3(std::max(gx, 1), nc), 256, 0, ctx->stream>>>(n, X, ldx, D, Y, ldy, ldw, inv_theta);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}
template <typename T>
int cheb_update(lb2_ctx* ctx, int64_t n, int nc, const T* AD, const T* Rin, int64_t ldrin, T* Rout, const T* Din, T* Dout,
                int64_t ldw, T* Y, int64_t ldy, real_t<T> c1, real_t<T> c2, bool write_r) {
  if (n <= 0 || nc <= 0) return 0;
  int gx = (int)std::min<int64_t>((n + 256 * 4 - 1) / (256 * 4), 1 << 20);
  cheb_update_kernel<T><<<dim3(std::max(gx, 1), nc), 256, 0, ctx->stream>>>(n, AD, Rin, ldrin, Rout, Din, Dout, ldw, Y, ldy,
                                                                           c1, c2, write_r ? 1 : 0);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// precision conversion of a block (mixed-precision preconditioner: double <-> float, c64 <-> c32)
template <typename TO, typename TI>
__device__ __forceinline__ TO cvt_scalar(TI v) {
  if constexpr (Sc<TI>::cplx) return TO{(typename Sc<TO>::real)v.re, (typename Sc<TO>::real)v.im};
  else return (TO)v;
}
template <typename TO, typename TI>
__global__ void __launch_bounds__(256)
    convert_block_kernel(int64_t n, const TI* __restrict__ X, int64_t ldx, TO* __restrict__ Y, int64_t ldy) {
  const TI* x = X + (int64_t)blockIdx.y * ldx;
  TO* y = Y + (int64_t)blockIdx.y * ldy;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = cvt_scalar<TO, TI>(x[i]);
}
template <typename TO, typename TI>
int convert_block(lb2_ctx* ctx, int64_t n, int nc, const TI* X, int64_t ldx, TO* Y, int64_t ldy) {
  if (n <= 0 || nc <= 0) return 0;
  int gx = (int)std::min<int64_t>((n + 256 * 4 - 1) / (256 * 4), 1 << 20);
  convert_block_kernel<TO, TI><<<dim3(std::max(gx, 1), nc), 256, 0, ctx->stream>>>(n, X, ldx, Y, ldy);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}
template int convert_block<float, double>(lb2_ctx*, int64_t, int, const double*, int64_t, float*, int64_t);
template int convert_block<double, float>(lb2_ctx*, int64_t, int, const float*, int64_t, double*, int64_t);
template int convert_block<c32, c64>(lb2_ctx*, int64_t, int, const c64*, int64_t, c32*, int64_t);
template int convert_block<c64, c32>(lb2_ctx*, int64_t, int, const c32*, int64_t, c64*, int64_t);

#define LB2_INST(T)                                                                                     \
  template int residual<T>(lb2_ctx*, int64_t, int, const T*, int64_t, const T*, int64_t, const real_t<T>*, \
                           T*, int64_t, real_t<T>*);                                                    \
  template int col_sumsq<T>(lb2_ctx*, int64_t, int, const T*, int64_t, real_t<T>*);                     \
  template int residual_monitor<T>(lb2_ctx*, int64_t, int, const T*, int64_t, const T*, int64_t, const T*, int64_t, const real_t<T>*, real_t<T>*); \
  template int fill_uniform<T>(lb2_ctx*, int64_t, int, T*, int64_t, uint64_t, int64_t, int64_t);                \
  template int scale_cols<T>(lb2_ctx*, int64_t, int, T*, int64_t, const real_t<T>*, real_t<T>);         \
  template int normalize_by<T>(lb2_ctx*, int64_t, T*, const real_t<T>*);                                \
  template int copy_block<T>(lb2_ctx*, int64_t, int, const T*, int64_t, T*, int64_t);                   \
  template int cheb_init<T>(lb2_ctx*, int64_t, int, const T*, int64_t, T*, T*, int64_t, int64_t, real_t<T>); \
  template int cheb_update<T>(lb2_ctx*, int64_t, int, const T*, const T*, int64_t, T*, const T*, T*, int64_t, T*, int64_t, \
                              real_t<T>, real_t<T>, bool);
LB2_INST(float)
LB2_INST(double)
LB2_INST(c32)
LB2_INST(c64)
#undef LB2_INST
template int sum_reals<float>(lb2_ctx*, int, const float*, float*);
template int sum_reals<double>(lb2_ctx*, int, const double*, double*);

}  // namespace lb2

// lobpcg_b200/csrc/smalldense.cu — the small (<= 3k x 3k) projected problem, kept on the device.
//
// D1-D8 of SURVEY.md §2b: diagonal scaling, Cholesky + condition estimate + D R^-1, congruence transform,
// symmetric/Hermitian eigensolve, QR of Z_{1,perp}^T, SVQB transform, orthogonality error.  These are
// O(k^3) on <= 900 x 900 matrices (< 1 % of the reference's time); the factorizations call
// cuSOLVER/cuBLAS as plain library routines, everything around them is small hand-written kernels so that
// no matrix leaves the device (the host only reads back `info`, a condition number or a column count).
//
// Reference call sites: src/rayleigh/rayleigh_ritz_impl.inc:49-96, rayleigh_ritz_modified_impl.inc:70-269,
// src/ortho/svqb_impl.inc:64-98, src/ortho/ortho_err_upper_impl.inc:2-13.
#include "common.cuh"
#include "context.h"
#include "smalldense.h"

namespace lb2 {

template <typename T> struct CudaType;
template <> struct CudaType<float>  { static constexpr cudaDataType v = CUDA_R_32F; static constexpr cudaDataType r = CUDA_R_32F; };
template <> struct CudaType<double> { static constexpr cudaDataType v = CUDA_R_64F; static constexpr cudaDataType r = CUDA_R_64F; };
template <> struct CudaType<c32>    { static constexpr cudaDataType v = CUDA_C_32F; static constexpr cudaDataType r = CUDA_R_32F; };
template <> struct CudaType<c64>    { static constexpr cudaDataType v = CUDA_C_64F; static constexpr cudaDataType r = CUDA_R_64F; };

#define LB2_SOLVER_OK(expr)                                                                          \
  do {                                                                                               \
    cusolverStatus_t _s = (expr);                                                                    \
    if (_s != CUSOLVER_STATUS_SUCCESS) {                                                             \
      fprintf(stderr, "lobpcg_b200: cuSOLVER status %d at %s:%d (%s)\n", (int)_s, __FILE__, __LINE__, #expr); \
      return -1;                                                                                     \
    }                                                                                                \
  } while (0)
#define LB2_BLAS_OK(expr)                                                                            \
  do {                                                                                               \
    cublasStatus_t _s = (expr);                                                                      \
    if (_s != CUBLAS_STATUS_SUCCESS) {                                                               \
      fprintf(stderr, "lobpcg_b200: cuBLAS status %d at %s:%d (%s)\n", (int)_s, __FILE__, __LINE__, #expr); \
      return -1;                                                                                     \
    }                                                                                                \
  } while (0)

static cusolverDnParams_t g_params = nullptr;

int sd_init(lb2_ctx* ctx) {
  if (!ctx->cublas) {
    LB2_BLAS_OK(cublasCreate(&ctx->cublas));
    LB2_BLAS_OK(cublasSetStream(ctx->cublas, ctx->stream));
    LB2_BLAS_OK(cublasSetPointerMode(ctx->cublas, CUBLAS_POINTER_MODE_HOST));
  }
  if (!ctx->cusolver) {
    LB2_SOLVER_OK(cusolverDnCreate(&ctx->cusolver));
    LB2_SOLVER_OK(cusolverDnSetStream(ctx->cusolver, ctx->stream));
  }
  if (!g_params) LB2_SOLVER_OK(cusolverDnCreateParams(&g_params));
  if (!ctx->dev_info) LB2_CUDA_OK(cudaMalloc(&ctx->dev_info, sizeof(int) * 4));
  return 0;
}

static int ensure_ws(lb2_ctx* ctx, size_t dev_bytes, size_t host_bytes) {
  if (dev_bytes > ctx->solver_ws_bytes) {
    LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    if (ctx->solver_ws) cudaFree(ctx->solver_ws);
    ctx->solver_ws = nullptr;
    const size_t want = dev_bytes + (dev_bytes >> 2) + 1024;
    LB2_CUDA_OK(cudaMalloc(&ctx->solver_ws, want));
    ctx->solver_ws_bytes = want;
  }
  if (host_bytes > ctx->solver_hws_bytes) {
    LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    free(ctx->solver_hws);
    ctx->solver_hws = malloc(host_bytes + 1024);
    if (!ctx->solver_hws) return -1;
    ctx->solver_hws_bytes = host_bytes + 1024;
  }
  return 0;
}

static int read_info(lb2_ctx* ctx, int* h_info) {
  LB2_CUDA_OK(cudaMemcpyAsync(h_info, ctx->dev_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  LB2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

template <typename T>
int sd_potrf_upper(lb2_ctx* ctx, int m, T* A, int lda, int* h_info) {
  if (sd_init(ctx)) return -1;
  size_t wd = 0, wh = 0;
  LB2_SOLVER_OK(cusolverDnXpotrf_bufferSize(ctx->cusolver, g_params, CUBLAS_FILL_MODE_UPPER, m, CudaType<T>::v, A,
                                            lda, CudaType<T>::v, &wd, &wh));
  if (ensure_ws(ctx, wd, wh)) return -1;
  LB2_SOLVER_OK(cusolverDnXpotrf(ctx->cusolver, g_params, CUBLAS_FILL_MODE_UPPER, m, CudaType<T>::v, A, lda,
                                 CudaType<T>::v, ctx->solver_ws, wd, ctx->solver_hws, wh, ctx->dev_info));
  ctx->launches++;
  return read_info(ctx, h_info);
}

template <typename T>
int sd_syevd_upper(lb2_ctx* ctx, int m, T* A, int lda, real_t<T>* w, int* h_info) {
  if (sd_init(ctx)) return -1;
  size_t wd = 0, wh = 0;
  LB2_SOLVER_OK(cusolverDnXsyevd_bufferSize(ctx->cusolver, g_params, CUSOLVER_EIG_MODE_VECTOR,
                                            CUBLAS_FILL_MODE_UPPER, m, CudaType<T>::v, A, lda, CudaType<T>::r, w,
                                            CudaType<T>::v, &wd, &wh));
  if (ensure_ws(ctx, wd, wh)) return -1;
  LB2_SOLVER_OK(cusolverDnXsyevd(ctx->cusolver, g_params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, m,
                                 CudaType<T>::v, A, lda, CudaType<T>::r, w, CudaType<T>::v, ctx->solver_ws, wd,
                                 ctx->solver_hws, wh, ctx->dev_info));
  ctx->launches++;
  return read_info(ctx, h_info);
}

// orgqr / ungqr have no 64-bit generic API: typed overloads
static cusolverStatus_t orgqr_bs(cusolverDnHandle_t h, int m, int n, int k, const float* A, int lda, const float* tau, int* lw) { return cusolverDnSorgqr_bufferSize(h, m, n, k, A, lda, tau, lw); }
static cusolverStatus_t orgqr_bs(cusolverDnHandle_t h, int m, int n, int k, const double* A, int lda, const double* tau, int* lw) { return cusolverDnDorgqr_bufferSize(h, m, n, k, A, lda, tau, lw); }
static cusolverStatus_t orgqr_bs(cusolverDnHandle_t h, int m, int n, int k, const c32* A, int lda, const c32* tau, int* lw) { return cusolverDnCungqr_bufferSize(h, m, n, k, (const cuComplex*)A, lda, (const cuComplex*)tau, lw); }
static cusolverStatus_t orgqr_bs(cusolverDnHandle_t h, int m, int n, int k, const c64* A, int lda, const c64* tau, int* lw) { return cusolverDnZungqr_bufferSize(h, m, n, k, (const cuDoubleComplex*)A, lda, (const cuDoubleComplex*)tau, lw); }
static cusolverStatus_t orgqr_run(cusolverDnHandle_t h, int m, int n, int k, float* A, int lda, const float* tau, float* w, int lw, int* info) { return cusolverDnSorgqr(h, m, n, k, A, lda, tau, w, lw, info); }
static cusolverStatus_t orgqr_run(cusolverDnHandle_t h, int m, int n, int k, double* A, int lda, const double* tau, double* w, int lw, int* info) { return cusolverDnDorgqr(h, m, n, k, A, lda, tau, w, lw, info); }
static cusolverStatus_t orgqr_run(cusolverDnHandle_t h, int m, int n, int k, c32* A, int lda, const c32* tau, c32* w, int lw, int* info) { return cusolverDnCungqr(h, m, n, k, (cuComplex*)A, lda, (const cuComplex*)tau, (cuComplex*)w, lw, info); }
static cusolverStatus_t orgqr_run(cusolverDnHandle_t h, int m, int n, int k, c64* A, int lda, const c64* tau, c64* w, int lw, int* info) { return cusolverDnZungqr(h, m, n, k, (cuDoubleComplex*)A, lda, (const cuDoubleComplex*)tau, (cuDoubleComplex*)w, lw, info); }

// A (rows x cols, rows >= cols) <- Q of its thin QR.  tau: cols scalars of scratch.
template <typename T>
int sd_qr_q(lb2_ctx* ctx, int rows, int cols, T* A, int lda, T* tau) {
  if (sd_init(ctx)) return -1;
  size_t wd = 0, wh = 0;
  LB2_SOLVER_OK(cusolverDnXgeqrf_bufferSize(ctx->cusolver, g_params, rows, cols, CudaType<T>::v, A, lda,
                                            CudaType<T>::v, tau, CudaType<T>::v, &wd, &wh));
  int lw = 0;
  LB2_SOLVER_OK(orgqr_bs(ctx->cusolver, rows, cols, cols, A, lda, tau, &lw));
  const size_t need = wd > sizeof(T) * (size_t)lw ? wd : sizeof(T) * (size_t)lw;
  if (ensure_ws(ctx, need, wh)) return -1;
  // both steps report through their own info word (the reference checks geqrf / orgqr, rayleigh_ritz_modified_impl.inc:233-246);
  // they are read together with the next host synchronisation of the caller (sd_qr_info)
  LB2_SOLVER_OK(cusolverDnXgeqrf(ctx->cusolver, g_params, rows, cols, CudaType<T>::v, A, lda, CudaType<T>::v, tau,
                                 CudaType<T>::v, ctx->solver_ws, wd, ctx->solver_hws, wh, ctx->dev_info + 2));
  LB2_SOLVER_OK(orgqr_run(ctx->cusolver, rows, cols, cols, A, lda, tau, (T*)ctx->solver_ws, lw, ctx->dev_info + 3));
  ctx->launches += 2;
  return 0;
}

// ---- cuBLAS typed shims ---------------------------------------------------------------------------
static cublasStatus_t gemm_(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const float* al, const float* A, int lda, const float* B, int ldb, const float* be, float* C, int ldc) { return cublasSgemm(h, ta, tb, m, n, k, al, A, lda, B, ldb, be, C, ldc); }
static cublasStatus_t gemm_(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const double* al, const double* A, int lda, const double* B, int ldb, const double* be, double* C, int ldc) { return cublasDgemm(h, ta, tb, m, n, k, al, A, lda, B, ldb, be, C, ldc); }
static cublasStatus_t gemm_(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const c32* al, const c32* A, int lda, const c32* B, int ldb, const c32* be, c32* C, int ldc) { return cublasCgemm(h, ta, tb, m, n, k, (const cuComplex*)al, (const cuComplex*)A, lda, (const cuComplex*)B, ldb, (const cuComplex*)be, (cuComplex*)C, ldc); }
static cublasStatus_t gemm_(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const c64* al, const c64* A, int lda, const c64* B, int ldb, const c64* be, c64* C, int ldc) { return cublasZgemm(h, ta, tb, m, n, k, (const cuDoubleComplex*)al, (const cuDoubleComplex*)A, lda, (const cuDoubleComplex*)B, ldb, (const cuDoubleComplex*)be, (cuDoubleComplex*)C, ldc); }
static cublasStatus_t trsm_(cublasHandle_t h, cublasSideMode_t s, cublasFillMode_t u, cublasOperation_t t, cublasDiagType_t d, int m, int n, const float* al, const float* A, int lda, float* B, int ldb) { return cublasStrsm(h, s, u, t, d, m, n, al, A, lda, B, ldb); }
static cublasStatus_t trsm_(cublasHandle_t h, cublasSideMode_t s, cublasFillMode_t u, cublasOperation_t t, cublasDiagType_t d, int m, int n, const double* al, const double* A, int lda, double* B, int ldb) { return cublasDtrsm(h, s, u, t, d, m, n, al, A, lda, B, ldb); }
static cublasStatus_t trsm_(cublasHandle_t h, cublasSideMode_t s, cublasFillMode_t u, cublasOperation_t t, cublasDiagType_t d, int m, int n, const c32* al, const c32* A, int lda, c32* B, int ldb) { return cublasCtrsm(h, s, u, t, d, m, n, (const cuComplex*)al, (const cuComplex*)A, lda, (cuComplex*)B, ldb); }
static cublasStatus_t trsm_(cublasHandle_t h, cublasSideMode_t s, cublasFillMode_t u, cublasOperation_t t, cublasDiagType_t d, int m, int n, const c64* al, const c64* A, int lda, c64* B, int ldb) { return cublasZtrsm(h, s, u, t, d, m, n, (const cuDoubleComplex*)al, (const cuDoubleComplex*)A, lda, (cuDoubleComplex*)B, ldb); }

// C = op(A) op(B), opa: 'N' or 'H' (conjugate transpose; plain transpose for real types); small matrices.
template <typename T>
int sd_gemm(lb2_ctx* ctx, char opa, int m, int n, int k, const T* A, int lda, const T* B, int ldb, T* C, int ldc) {
  if (sd_init(ctx)) return -1;
  if (m <= 0 || n <= 0) return 0;
  const T one = make<T>(1), zer = zero<T>();
  const cublasOperation_t ta = (opa == 'N') ? CUBLAS_OP_N : (Sc<T>::cplx ? CUBLAS_OP_C : CUBLAS_OP_T);
  LB2_BLAS_OK(gemm_(ctx->cublas, ta, CUBLAS_OP_N, m, n, k, &one, A, lda, B, ldb, &zer, C, ldc));
  ctx->launches++;
  return 0;
}

// X <- X R^-1 (R upper triangular, non-unit): the reference's trsm_run (blas_wrapper.h:364-386)
template <typename T>
int sd_trsm_run(lb2_ctx* ctx, int rows, int m, const T* R, int ldr, T* X, int ldx) {
  if (sd_init(ctx)) return -1;
  const T one = make<T>(1);
  LB2_BLAS_OK(trsm_(ctx->cublas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, rows,
                    m, &one, R, ldr, X, ldx));
  ctx->launches++;
  return 0;
}

// ---- hand-written small kernels ---------------------------------------------------------------------
// D[i] = 1/sqrt(|G_ii|) (1 if zero); G <- D G D.  rayleigh_ritz_impl.inc:49-60, svqb_impl.inc:65-73.
template <typename T>
__global__ void dscale_kernel(int m, T* __restrict__ G, int ldg, real_t<T>* __restrict__ D, int phase) {
  using R = real_t<T>;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (phase == 0) {
    if (idx < m) {
      const R gii = sqrt(abs2_(G[idx + (int64_t)idx * ldg]));
      D[idx] = (gii > R(0)) ? R(1) / sqrt(gii) : R(1);
    }
  } else {
    if (idx < m * m) {
      const int i = idx % m, j = idx / m;
      T* p = G + i + (int64_t)j * ldg;
      *p = rscale_(*p, D[i] * D[j]);
    }
  }
}
template <typename T>
int sd_dscale(lb2_ctx* ctx, int m, T* G, int ldg, real_t<T>* D) {
  dscale_kernel<T><<<(m + 127) / 128, 128, 0, ctx->stream>>>(m, G, ldg, D, 0);
  dscale_kernel<T><<<(m * m + 255) / 256, 256, 0, ctx->stream>>>(m, G, ldg, D, 1);
  ctx->launches += 2;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// M (m x m) <- diag(D)
template <typename T>
__global__ void set_diag_kernel(int m, T* __restrict__ M, int ldm, const real_t<T>* __restrict__ D) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < m * m) {
    const int i = idx % m, j = idx / m;
    M[i + (int64_t)j * ldm] = (i == j) ? make<T>(D[i]) : zero<T>();
  }
}
template <typename T>
int sd_set_diag(lb2_ctx* ctx, int m, T* M, int ldm, const real_t<T>* D) {
  set_diag_kernel<T><<<(m * m + 255) / 256, 256, 0, ctx->stream>>>(m, M, ldm, D);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// rcond_1(R) = 1 / (||R||_1 ||R^-1||_1) with R^-1 = diag(1/D) * DinvR (exact, where LAPACK trcon estimates;
// rayleigh_ritz_modified_impl.inc:169-178).  One CTA per column (coalesced column sums), then one CTA takes the maxima;
// out[0] = rcond.  (The first version walked the columns with one thread each: 1.1 ms at m = 900, ncu launch list r02.)
template <typename T>
__global__ void __launch_bounds__(128)
    rcond_colsum_kernel(int m, const T* __restrict__ Rm, int ldr, const T* __restrict__ DinvR, int ldd,
                        const real_t<T>* __restrict__ D, real_t<T>* __restrict__ sums) {
  using R = real_t<T>;
  const int j = blockIdx.x;
  R c1 = 0, c2 = 0;
  for (int i = threadIdx.x; i <= j; i += blockDim.x) {
    c1 += sqrt(abs2_(Rm[i + (int64_t)j * ldr]));
    c2 += sqrt(abs2_(DinvR[i + (int64_t)j * ldd])) / D[i];
  }
  __shared__ R s1[128], s2[128];
  s1[threadIdx.x] = c1;
  s2[threadIdx.x] = c2;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { sums[j] = s1[0]; sums[m + j] = s2[0]; }
}
template <typename R>
__global__ void rcond_finish_kernel(int m, const R* __restrict__ sums, R* __restrict__ out) {
  __shared__ R s1[256], s2[256];
  R n1 = 0, n2 = 0;
  for (int j = threadIdx.x; j < m; j += blockDim.x) { n1 = fmax(n1, sums[j]); n2 = fmax(n2, sums[m + j]); }
  s1[threadIdx.x] = n1;
  s2[threadIdx.x] = n2;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s1[threadIdx.x] = fmax(s1[threadIdx.x], s1[threadIdx.x + o]);
      s2[threadIdx.x] = fmax(s2[threadIdx.x], s2[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const R p = s1[0] * s2[0];
    out[0] = (p > R(0) && isfinite((double)p)) ? R(1) / p : R(0);
  }
}
template <typename T>
int sd_rcond(lb2_ctx* ctx, int m, const T* Rm, int ldr, const T* DinvR, int ldd, const real_t<T>* D,
             real_t<T>* out_dev) {
  using R = real_t<T>;
  R* sums = (R*)ctx_scratch(ctx, sizeof(R) * 2 * (size_t)m);
  if (!sums) return -1;
  rcond_colsum_kernel<T><<<m, 128, 0, ctx->stream>>>(m, Rm, ldr, DinvR, ldd, D, sums);
  rcond_finish_kernel<R><<<1, 256, 0, ctx->stream>>>(m, sums, out_dev);
  ctx->launches += 2;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// Out (rows x cols, ld ldo) <- plain transpose of In[r0:r0+cols, c0:c0+rows] :  Out[j,i] = In[r0+i, c0+j]
// (Z_{1,perp}^T extraction, rayleigh_ritz_modified_impl.inc:100-104 — no conjugation, as the reference)
template <typename T>
__global__ void transpose_kernel(int rows, int cols, const T* __restrict__ In, int ldi, T* __restrict__ Out, int ldo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < rows * cols) {
    const int j = idx % rows, i = idx / rows;
    Out[j + (int64_t)i * ldo] = In[i + (int64_t)j * ldi];
  }
}
template <typename T>
int sd_transpose(lb2_ctx* ctx, int rows, int cols, const T* In, int ldi, T* Out, int ldo) {
  if (rows <= 0 || cols <= 0) return 0;
  transpose_kernel<T><<<(rows * cols + 255) / 256, 256, 0, ctx->stream>>>(rows, cols, In, ldi, Out, ldo);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// SVQB transform (svqb_impl.inc:82-98): thresh = tau*max|lam|; for retained j (all j unless drop and
// |lam_j| < thresh): Tm[:,c] = D .* V[:,j] / sqrt(max(|lam_j|, thresh)).  count[0] = retained columns.
template <typename T>
__global__ void svqb_transform_kernel(int m, const T* __restrict__ V, int ldv, const real_t<T>* __restrict__ lam,
                                      const real_t<T>* __restrict__ D, real_t<T> tau, int drop,
                                      T* __restrict__ Tm, int ldt, int* __restrict__ count) {
  using R = real_t<T>;
  __shared__ R smax[256];
  extern __shared__ int spos[];   // m + 1 entries (dynamic)
  R mx = 0;
  for (int j = threadIdx.x; j < m; j += blockDim.x) mx = fmax(mx, fabs(lam[j]));
  smax[threadIdx.x] = mx;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + o]);
    __syncthreads();
  }
  const R thresh = tau * smax[0];
  if (threadIdx.x == 0) {
    int c = 0;
    for (int j = 0; j < m; j++) {
      const bool keep = !(drop && fabs(lam[j]) < thresh);
      spos[j] = keep ? c : -1;
      c += keep ? 1 : 0;
    }
    count[0] = c;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
    const int i = idx % m, j = idx / m;
    const int c = spos[j];
    if (c >= 0) {
      const R df = R(1) / sqrt(fmax(fabs(lam[j]), thresh));
      Tm[i + (int64_t)c * ldt] = rscale_(V[i + (int64_t)j * ldv], D[i] * df);
    }
  }
}
template <typename T>
int sd_svqb_transform(lb2_ctx* ctx, int m, const T* V, int ldv, const real_t<T>* lam, const real_t<T>* D,
                      real_t<T> tau, int drop, T* Tm, int ldt, int* count_dev) {
  if (m > 12000) {   // 48 KB of dynamic shared memory; Solver::prepare rejects such block sizes up front
    fprintf(stderr, "lobpcg_b200: svqb block wider than 12000 columns is not supported\n");
    return -1;
  }
  svqb_transform_kernel<T><<<1, 256, sizeof(int) * (size_t)(m + 1), ctx->stream>>>(m, V, ldv, lam, D, tau, drop, Tm, ldt, count_dev);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// ||G - I||_F from the upper triangle (ortho_err_upper_impl.inc:2-13) and ||G||_F of a full matrix.
template <typename T>
__global__ void ortho_err_kernel(int rows, int cols, const T* __restrict__ G, int ldg, int mode,
                                 real_t<T>* __restrict__ out) {
  using R = real_t<T>;
  __shared__ R red[256];
  R s = 0;
  for (int idx = threadIdx.x; idx < rows * cols; idx += blockDim.x) {
    const int i = idx % rows, j = idx / rows;
    const R a = sqrt(abs2_(G[i + (int64_t)j * ldg]));
    if (mode == 0) {  // upper-triangle identity error
      if (i == j) s += (a - R(1)) * (a - R(1));
      else if (i < j) s += R(2) * a * a;
    } else {
      s += a * a;
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sqrt(red[0]);
}
template <typename T>
int sd_ortho_err_upper(lb2_ctx* ctx, int m, const T* G, int ldg, real_t<T>* out_dev) {
  ortho_err_kernel<T><<<1, 256, 0, ctx->stream>>>(m, m, G, ldg, 0, out_dev);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}
template <typename T>
int sd_frob(lb2_ctx* ctx, int rows, int cols, const T* G, int ldg, real_t<T>* out_dev) {
  ortho_err_kernel<T><<<1, 256, 0, ctx->stream>>>(rows, cols, G, ldg, 1, out_dev);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}


// general small triangular solve with the upper factor R: side 'L': op(R) X = B, side 'R': X op(R) = B; op 'N' or 'H'
template <typename T>
int sd_trsm_upper(lb2_ctx* ctx, char side, char op, int rows, int cols, const T* R, int ldr, T* X, int ldx) {
  if (sd_init(ctx)) return -1;
  const T one = make<T>(1);
  const cublasOperation_t t = (op == 'N') ? CUBLAS_OP_N : (Sc<T>::cplx ? CUBLAS_OP_C : CUBLAS_OP_T);
  LB2_BLAS_OK(trsm_(ctx->cublas, side == 'L' ? CUBLAS_SIDE_LEFT : CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, t,
                    CUBLAS_DIAG_NON_UNIT, rows, cols, &one, R, ldr, X, ldx));
  ctx->launches++;
  return 0;
}

// C = alpha * op(A) B + beta * C on small matrices
template <typename T>
int sd_gemm_ab(lb2_ctx* ctx, char opa, int m, int n, int k, T alpha, const T* A, int lda, const T* B, int ldb, T beta,
               T* C, int ldc) {
  if (sd_init(ctx)) return -1;
  if (m <= 0 || n <= 0) return 0;
  const cublasOperation_t ta = (opa == 'N') ? CUBLAS_OP_N : (Sc<T>::cplx ? CUBLAS_OP_C : CUBLAS_OP_T);
  LB2_BLAS_OK(gemm_(ctx->cublas, ta, CUBLAS_OP_N, m, n, k, &alpha, A, lda, B, ldb, &beta, C, ldc));
  ctx->launches++;
  return 0;
}

// Indefinite RR back end (replaces GGEV + B-normalisation + signature sort of the reference,
// src/rayleigh/indefinite_rr_impl.inc:73-145): input mu (ascending) and V with V^H G_B V = diag(mu) (columns of
// R^-1 W); output columns scaled to |v^H G_B v| = 1, theta = 1/mu, signature = sign(mu), ordered positive
// signature first with theta ascending, then negative signature with theta descending (bubble_sort_sig_impl.inc).
template <typename T>
__global__ void indef_finalize_kernel(int m, const real_t<T>* __restrict__ mu, const T* __restrict__ V, int ldv,
                                      T* __restrict__ VR, int ldo, real_t<T>* __restrict__ theta,
                                      int8_t* __restrict__ sig) {
  using R = real_t<T>;
  __shared__ int npos_s;
  if (threadIdx.x == 0) {
    int c = 0;
    for (int j = 0; j < m; j++) c += (mu[j] >= R(0)) ? 1 : 0;
    npos_s = c;
  }
  __syncthreads();
  const int npos = npos_s, nneg = m - npos;
  // mu ascending: indices [0,nneg) negative (theta descending as listed), [nneg,m) positive (theta ascending when reversed)
  for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
    const int i = idx % m, jo = idx / m;                 // output column jo
    const int js = (jo < npos) ? (m - 1 - jo) : (jo - npos);
    const R a = fabs(mu[js]);
    const R sc = (a > R(1e-30)) ? R(1) / sqrt(a) : R(1);
    VR[i + (int64_t)jo * ldo] = rscale_(V[i + (int64_t)js * ldv], sc);
    if (i == 0) {
      theta[jo] = (a > R(1e-30)) ? R(1) / mu[js] : ((mu[js] >= R(0)) ? R(1e30) : R(-1e30));
      sig[jo] = (mu[js] >= R(0)) ? 1 : -1;
    }
  }
  (void)nneg;
}
template <typename T>
int sd_indef_finalize(lb2_ctx* ctx, int m, const real_t<T>* mu, const T* V, int ldv, T* VR, int ldo,
                      real_t<T>* theta, int8_t* sig) {
  indef_finalize_kernel<T><<<1, 256, 0, ctx->stream>>>(m, mu, V, ldv, VR, ldo, theta, sig);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// Cp = [0; Cx[nx:m, :]]  (indefinite_rr_modified_impl.inc:229-235)
template <typename T>
__global__ void cp_lower_kernel(int m, int nx, const T* __restrict__ Cx, T* __restrict__ Cp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < m * nx) {
    const int i = idx % m;
    Cp[idx] = (i < nx) ? zero<T>() : Cx[idx];
  }
}
template <typename T>
int sd_cp_lower(lb2_ctx* ctx, int m, int nx, const T* Cx, T* Cp) {
  cp_lower_kernel<T><<<(m * nx + 255) / 256, 256, 0, ctx->stream>>>(m, nx, Cx, Cp);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// Gram matrix of S = [X P | W] from the cached [X P] block and the freshly contracted W columns (cached-Gram pass,
// solver.cu: rr_modified): G (m x m) = [[Gc, Gw_top], [Gw_top^H, herm(Gw_bottom)]].  Only the upper triangles of Gc and of
// the W x W block of Gw are read (the column-block kernel leaves the strictly-lower tiles of that block unwritten); the
// result is exactly Hermitian.  Gc == nullptr: identity (ortho branch, B-orthonormal basis).
template <typename T>
__global__ void assemble_gram_kernel(int m, int mxp, const T* __restrict__ Gc, int ldc, const T* __restrict__ Gw, int ldw,
                                     T* __restrict__ G, int ldg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * m) return;
  const int i = idx % m, j = idx / m;
  const int lo = i < j ? i : j, hi = i < j ? j : i;   // (lo, hi) is the upper-triangle twin of (i, j)
  T v;
  if (hi < mxp) v = Gc ? Gc[lo + (int64_t)hi * ldc] : ((lo == hi) ? make<T>(real_t<T>(1)) : zero<T>());
  else v = Gw[lo + (int64_t)(hi - mxp) * ldw];
  if (i > j) v = conj_(v);
  if (i == j) v = make<T>(real_(v));
  G[i + (int64_t)j * ldg] = v;
}
template <typename T>
int sd_assemble_gram(lb2_ctx* ctx, int m, int mxp, const T* Gc, int ldc, const T* Gw, int ldw, T* G, int ldg) {
  if (m <= 0) return 0;
  assemble_gram_kernel<T><<<(m * m + 255) / 256, 256, 0, ctx->stream>>>(m, mxp, Gc, ldc, Gw, ldw, G, ldg);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- general (non-definite) projected pencil: G_A v = theta G_B v through a non-symmetric standard problem -----------
// The reference hands (G_A, G_B) to LAPACK GGEV (src/rayleigh/indefinite_rr_modified_impl.inc:94-118); cuSOLVER has no
// GGEV, so the pencil is reduced with an LU solve, M = G_B^-1 G_A (G_B = S^H B S is Hermitian indefinite but non-singular
// for a B-orthonormalised basis), and M goes to cusolverDnXgeev in complex arithmetic.  Used when S^H A S is not positive
// definite (otherwise the Hermitian form in Solver::rr_indef applies).

// B (m x nrhs) <- A^-1 B by LU with partial pivoting; A is overwritten by its factors.  ipiv: m int64 on the device.
template <typename T>
int sd_lu_solve(lb2_ctx* ctx, int m, T* A, int lda, T* B, int ldb, int nrhs, int64_t* ipiv, int* h_info) {
  if (sd_init(ctx)) return -1;
  size_t wd = 0, wh = 0;
  LB2_SOLVER_OK(cusolverDnXgetrf_bufferSize(ctx->cusolver, g_params, m, m, CudaType<T>::v, A, lda, CudaType<T>::v, &wd, &wh));
  if (ensure_ws(ctx, wd, wh)) return -1;
  LB2_SOLVER_OK(cusolverDnXgetrf(ctx->cusolver, g_params, m, m, CudaType<T>::v, A, lda, ipiv, CudaType<T>::v, ctx->solver_ws,
                                 wd, ctx->solver_hws, wh, ctx->dev_info));
  ctx->launches++;
  if (read_info(ctx, h_info)) return -1;
  if (*h_info != 0) return 0;
  LB2_SOLVER_OK(cusolverDnXgetrs(ctx->cusolver, g_params, CUBLAS_OP_N, m, nrhs, CudaType<T>::v, A, lda, ipiv, CudaType<T>::v,
                                 B, ldb, ctx->dev_info));
  ctx->launches++;
  return read_info(ctx, h_info);
}

template <typename T, typename CT>
__global__ void to_complex_kernel(int m, const T* __restrict__ A, int lda, CT* __restrict__ C, int ldc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * m) return;
  const int i = idx % m, j = idx / m;
  const T a = A[i + (int64_t)j * lda];
  if constexpr (Sc<T>::cplx) C[i + (int64_t)j * ldc] = a;
  else C[i + (int64_t)j * ldc] = CT{a, real_t<T>(0)};
}

// eigenvalues W (complex, m) and right eigenvectors VR (complex, m x m) of the m x m matrix M (type T, destroyed is Mc)
// Mc, W, VR: complex scratch of m*m, m, m*m elements of ComplexOf<T>.
template <typename T>
int sd_geev(lb2_ctx* ctx, int m, const T* M, int ldm, void* Mc_, void* W_, void* VR_, int* h_info) {
  using CT = typename ComplexOf<T>::type;
  if (sd_init(ctx)) return -1;
  CT* Mc = (CT*)Mc_;
  CT* W = (CT*)W_;
  CT* VR = (CT*)VR_;
  to_complex_kernel<T, CT><<<(m * m + 255) / 256, 256, 0, ctx->stream>>>(m, M, ldm, Mc, m);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  size_t wd = 0, wh = 0;
  constexpr cudaDataType ct = CudaType<CT>::v;
  LB2_SOLVER_OK(cusolverDnXgeev_bufferSize(ctx->cusolver, g_params, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_VECTOR, m, ct,
                                           Mc, m, ct, W, ct, nullptr, m, ct, VR, m, ct, &wd, &wh));
  if (ensure_ws(ctx, wd, wh)) return -1;
  LB2_SOLVER_OK(cusolverDnXgeev(ctx->cusolver, g_params, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_VECTOR, m, ct, Mc, m, ct,
                                W, ct, nullptr, m, ct, VR, m, ct, ctx->solver_ws, wd, ctx->solver_hws, wh, ctx->dev_info));
  ctx->launches++;
  return read_info(ctx, h_info);
}

// theta[j] = Re(W[j]);  V[:, j] = VR[:, j] rotated so that its largest component is real positive (a real pencil with a
// real eigenvalue has a real eigenvector up to that phase; the reference's real GGEV returns exactly that), cast to T.
template <typename T, typename CT>
__global__ void geev_extract_kernel(int m, const CT* __restrict__ W, const CT* __restrict__ VR, real_t<T>* __restrict__ theta,
                                    T* __restrict__ V, int ldv) {
  using R = real_t<T>;
  const int j = blockIdx.x;
  __shared__ R smax[128];
  __shared__ int sidx[128];
  R best = -1;
  int bi = 0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const R a = abs2_(VR[i + (int64_t)j * m]);
    if (a > best) { best = a; bi = i; }
  }
  smax[threadIdx.x] = best;
  sidx[threadIdx.x] = bi;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const bool take = smax[threadIdx.x + o] > smax[threadIdx.x] ||
                        (smax[threadIdx.x + o] == smax[threadIdx.x] && sidx[threadIdx.x + o] < sidx[threadIdx.x]);
      if (take) { smax[threadIdx.x] = smax[threadIdx.x + o]; sidx[threadIdx.x] = sidx[threadIdx.x + o]; }
    }
    __syncthreads();
  }
  const CT piv = VR[sidx[0] + (int64_t)j * m];
  const R pa = sqrt(abs2_(piv));
  const CT ph = (pa > R(0)) ? CT{piv.re / pa, -piv.im / pa} : CT{R(1), R(0)};   // conj(piv) / |piv|
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const CT v = mul_(VR[i + (int64_t)j * m], ph);
    if constexpr (Sc<T>::cplx) V[i + (int64_t)j * ldv] = v;
    else V[i + (int64_t)j * ldv] = v.re;
  }
  if (threadIdx.x == 0) theta[j] = W[j].re;
}
template <typename T>
int sd_geev_extract(lb2_ctx* ctx, int m, const void* W, const void* VR, real_t<T>* theta, T* V, int ldv) {
  using CT = typename ComplexOf<T>::type;
  geev_extract_kernel<T, CT><<<m, 128, 0, ctx->stream>>>(m, (const CT*)W, (const CT*)VR, theta, V, ldv);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// B-normalisation of eigenvector columns (indefinite_rr_modified_impl.inc:136-176): E = V^H G_B V given, column j is scaled
// by 1 / sqrt(|E_jj|) (when |E_jj| > 1e-30); sig (may be null) receives sign(Re E_jj).
template <typename T>
__global__ void bnormalize_kernel(int m, T* __restrict__ V, int ldv, const T* __restrict__ E, int lde, int8_t* __restrict__ sig) {
  using R = real_t<T>;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * m) return;
  const int i = idx % m, j = idx / m;
  const T d = E[j + (int64_t)j * lde];
  const R a = sqrt(abs2_(d));
  if (a > R(1e-30)) V[i + (int64_t)j * ldv] = rscale_(V[i + (int64_t)j * ldv], R(1) / sqrt(a));
  if (sig && i == 0) sig[j] = (real_(d) >= R(0)) ? 1 : -1;
}
template <typename T>
int sd_bnormalize(lb2_ctx* ctx, int m, T* V, int ldv, const T* E, int lde, int8_t* sig) {
  bnormalize_kernel<T><<<(m * m + 255) / 256, 256, 0, ctx->stream>>>(m, V, ldv, E, lde, sig);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// quality numbers of the reference (indefinite_rr_modified_impl.inc:178-192): out[0] = || E with |E_jj| - 1 on the diagonal ||_F,
// out[1] = ||V||_F, out[2] = ||G_B V||_F
template <typename T>
__global__ void indef_quality_kernel(int m, const T* __restrict__ E, const T* __restrict__ V, const T* __restrict__ GV,
                                     real_t<T>* __restrict__ out) {
  using R = real_t<T>;
  __shared__ R r0[256], r1[256], r2[256];
  R a = 0, b = 0, c = 0;
  for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
    const int i = idx % m, j = idx / m;
    const R e = (i == j) ? sqrt(abs2_(E[idx])) - R(1) : sqrt(abs2_(E[idx]));
    a += e * e;
    b += abs2_(V[idx]);
    c += abs2_(GV[idx]);
  }
  r0[threadIdx.x] = a; r1[threadIdx.x] = b; r2[threadIdx.x] = c;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { r0[threadIdx.x] += r0[threadIdx.x + o]; r1[threadIdx.x] += r1[threadIdx.x + o]; r2[threadIdx.x] += r2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = sqrt(r0[0]); out[1] = sqrt(r1[0]); out[2] = sqrt(r2[0]); }
}
template <typename T>
int sd_indef_quality(lb2_ctx* ctx, int m, const T* E, const T* V, const T* GV, real_t<T>* out_dev) {
  indef_quality_kernel<T><<<1, 256, 0, ctx->stream>>>(m, E, V, GV, out_dev);   // E, V, GV packed (ld = m)
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

// signature sort (bubble_sort_sig_impl.inc: positive signature first with theta ascending, then negative with theta
// descending; stable) + column reordering: Vout[:, rank(j)] = V[:, j]
template <typename T>
__global__ void indef_sort_kernel(int m, const real_t<T>* __restrict__ theta, const int8_t* __restrict__ sig,
                                  const T* __restrict__ V, int ldv, T* __restrict__ Vout, int ldo,
                                  real_t<T>* __restrict__ theta_out, int8_t* __restrict__ sig_out) {
  const int j = blockIdx.x;
  __shared__ int rank_s;
  __shared__ int cnt[128];
  const int sj = sig[j];
  const real_t<T> tj = theta[j];
  int c = 0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    if (i == j) continue;
    const int si = sig[i];
    const real_t<T> ti = theta[i];
    bool before;   // does i come before j ?
    if (si > 0 && sj < 0) before = true;
    else if (si < 0 && sj > 0) before = false;
    else if (si > 0) before = (ti < tj) || (ti == tj && i < j);
    else before = (ti > tj) || (ti == tj && i < j);
    c += before ? 1 : 0;
  }
  cnt[threadIdx.x] = c;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) cnt[threadIdx.x] += cnt[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) { rank_s = cnt[0]; theta_out[cnt[0]] = tj; sig_out[cnt[0]] = (int8_t)sj; }
  __syncthreads();
  const int r = rank_s;
  for (int i = threadIdx.x; i < m; i += blockDim.x) Vout[i + (int64_t)r * ldo] = V[i + (int64_t)j * ldv];
}
template <typename T>
int sd_indef_sort(lb2_ctx* ctx, int m, const real_t<T>* theta, const int8_t* sig, const T* V, int ldv, T* Vout, int ldo,
                  real_t<T>* theta_out, int8_t* sig_out) {
  indef_sort_kernel<T><<<m, 128, 0, ctx->stream>>>(m, theta, sig, V, ldv, Vout, ldo, theta_out, sig_out);
  ctx->launches++;
  LB2_CUDA_OK(cudaGetLastError());
  return 0;
}

#define LB2_INST(T)                                                                                         \
  template int sd_potrf_upper<T>(lb2_ctx*, int, T*, int, int*);                                             \
  template int sd_syevd_upper<T>(lb2_ctx*, int, T*, int, real_t<T>*, int*);                                 \
  template int sd_qr_q<T>(lb2_ctx*, int, int, T*, int, T*);                                                 \
  template int sd_gemm<T>(lb2_ctx*, char, int, int, int, const T*, int, const T*, int, T*, int);            \
  template int sd_trsm_run<T>(lb2_ctx*, int, int, const T*, int, T*, int);                                  \
  template int sd_dscale<T>(lb2_ctx*, int, T*, int, real_t<T>*);                                            \
  template int sd_set_diag<T>(lb2_ctx*, int, T*, int, const real_t<T>*);                                    \
  template int sd_rcond<T>(lb2_ctx*, int, const T*, int, const T*, int, const real_t<T>*, real_t<T>*);      \
  template int sd_transpose<T>(lb2_ctx*, int, int, const T*, int, T*, int);                                 \
  template int sd_svqb_transform<T>(lb2_ctx*, int, const T*, int, const real_t<T>*, const real_t<T>*, real_t<T>, int, T*, int, int*); \
  template int sd_ortho_err_upper<T>(lb2_ctx*, int, const T*, int, real_t<T>*);                             \
  template int sd_frob<T>(lb2_ctx*, int, int, const T*, int, real_t<T>*);                                  \
  template int sd_trsm_upper<T>(lb2_ctx*, char, char, int, int, const T*, int, T*, int);                  \
  template int sd_gemm_ab<T>(lb2_ctx*, char, int, int, int, T, const T*, int, const T*, int, T, T*, int); \
  template int sd_indef_finalize<T>(lb2_ctx*, int, const real_t<T>*, const T*, int, T*, int, real_t<T>*, int8_t*); \
  template int sd_cp_lower<T>(lb2_ctx*, int, int, const T*, T*);                                           \
  template int sd_assemble_gram<T>(lb2_ctx*, int, int, const T*, int, const T*, int, T*, int);          \
  template int sd_lu_solve<T>(lb2_ctx*, int, T*, int, T*, int, int, int64_t*, int*);                       \
  template int sd_geev<T>(lb2_ctx*, int, const T*, int, void*, void*, void*, int*);                        \
  template int sd_geev_extract<T>(lb2_ctx*, int, const void*, const void*, real_t<T>*, T*, int);           \
  template int sd_bnormalize<T>(lb2_ctx*, int, T*, int, const T*, int, int8_t*);                           \
  template int sd_indef_quality<T>(lb2_ctx*, int, const T*, const T*, const T*, real_t<T>*);               \
  template int sd_indef_sort<T>(lb2_ctx*, int, const real_t<T>*, const int8_t*, const T*, int, T*, int, real_t<T>*, int8_t*);
LB2_INST(float)
LB2_INST(double)
LB2_INST(c32)
LB2_INST(c64)
#undef LB2_INST

}  // namespace lb2

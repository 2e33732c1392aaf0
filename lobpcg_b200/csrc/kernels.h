// lobpcg_b200/csrc/kernels.h — internal C++ (templated) interface of the device kernels.  The C ABI in
// include/lobpcg_b200.h instantiates these for s/d/c/z.  All pointers are DEVICE pointers, all matrices
// column-major, all work is enqueued on ctx->stream.  Return 0 on success.
#pragma once
#include <stdint.h>
#include <type_traits>
#include "common.cuh"
#include "context.h"

namespace lb2 {

// ---- dense.cu ---------------------------------------------------------------------------------------
template <typename T>
int gram(lb2_ctx* ctx, int64_t n, int ma, int mb, const T* A, int64_t lda, const T* B, int64_t ldb, T* G,
         int ldg, int upper);
template <typename T>
int tall_nn(lb2_ctx* ctx, int64_t n, int kd, int nb, T alpha, const T* S, int64_t lds, const T* C, int ldc,
            T beta, T* Out, int64_t ldo);

// column-block products G0 = S^H W0, G1 = S^H W1 (W1 may be null) of the cached-Gram pass; see dense.cu
template <typename T>
int gram_cols(lb2_ctx* ctx, int64_t n, int m, int nw, const T* S, int64_t lds, const T* W0, int64_t ldw0, T* G0,
              int ldg0, const T* W1, int64_t ldw1, T* G1, int ldg1, int tri_c0);

// gram_wl.cu: f64 Gram, work-list kernel (masked diagonal / ragged tiles, tiles laid end to end over the CTAs)
int gram_wl_f64(lb2_ctx* ctx, int64_t n, int ma, int mb, const double* A, int64_t lda, const double* B, int64_t ldb,
                double* G, int ldg, int upper);
// dense.cu: f64 Gram on the tile x equal-split grid (first kernel; still used for rectangular products and strips)
int gram_tiles_f64(lb2_ctx* ctx, int64_t n, int ma, int mb, const double* A, int64_t lda, const double* B,
                   int64_t ldb, double* G, int ldg, int upper);
int gram_reduce_f32(lb2_ctx* ctx, const float* part, int64_t split_stride, int nsplit, int ma, int mb, int mirror,
                    float* G, int ldg);
// gram_i8.cu: f64 Gram on tcgen05 kind::i8 through an Ozaki split into 7 int8 slices (option gram_i8); -100 = not available
int gram_i8_f64(lb2_ctx* ctx, int64_t n, int ma, int mb, const double* A, int64_t lda, const double* B, int64_t ldb, double* G,
                int ldg, int upper);
int gram_cols_i8_f64(lb2_ctx* ctx, int64_t n, int m, int nw, const double* S, int64_t lds, const double* W0, int64_t ldw0, double* G0,
                     int ldg0, const double* W1, int64_t ldw1, double* G1, int ldg1, int tri_c0);
// Out = alpha S C + beta Out on the same slices, S read as an MN-major int8 operand; -100 = not available
int tall_nn_i8_f64(lb2_ctx* ctx, int64_t n, int kd, int nb, double alpha, const double* S, int64_t lds, const double* C, int ldc,
                   double beta, double* Out, int64_t ldo);
bool oz_slices_cached(const lb2_ctx* ctx, const void* S, int64_t n, int kd, int64_t lds);
int oz_plan_check(int m, int nw, int nprod, int tri_c0, int64_t n, int nworkers, int mode, double* stats);   // host-only schedule check
int oz_stats_query(lb2_ctx* ctx, double* out);   // {split ms, MMA kernel ms, reduce ms, calls} of the int8 Gram since the last query
// gram_tc5.cu: float Gram on tcgen05 (kind::tf32, 3xTF32 split, accumulator in TMEM); -100 = alignment not met
int gram_tc5_f32(lb2_ctx* ctx, int64_t n, int ma, int mb, const float* A, int64_t lda, const float* B, int64_t ldb,
                 float* G, int ldg, int upper);
// nn_tc5.cu: float projection Out = alpha S C + beta Out on tcgen05 (3xTF32, TMEM accumulator); -100 = alignment not met
int nn_tc5_f32(lb2_ctx* ctx, int64_t n, int kd, int nb, float alpha, const float* S, int64_t lds, const float* C, int ldc,
               float beta, float* Out, int64_t ldo);
// gram_wl.cu: column-block products of the cached-Gram pass: G0[0:m,0:nw] = S^H W0, G1[0:m,0:nw] = S^H W1 (W1 may be null) in
// one launch; tri_c0 >= 0: rows tri_c0.. of the results are a Hermitian block whose strictly-lower tiles are skipped
int gram_wl_cols_f64(lb2_ctx* ctx, int64_t n, int m, int nw, const double* S, int64_t lds, const double* W0, int64_t ldw0,
                     double* G0, int ldg0, const double* W1, int64_t ldw1, double* G1, int ldg1, int tri_c0);
int gram_wl_cols_plan_check(int m, int nw, int nprod, int tri_c0, int64_t n, int ncta, int BK, double* stats);
int gram_wl_plan_check(int ma, int mb, int upper, int64_t n, int ncta, int BK, double* stats);
int gram_wl_plan_sharing(int ma, int mb, int upper, int64_t n, int ncta, int BK, int phase, int window, int samples,
                         double* share);

// ---- elementwise.cu ---------------------------------------------------------------------------------
// W[:,j] = AX[:,j] - lambda[j] * BX[:,j]  (W may be null: norms only); sumsq[j] = ||W[:,j]||^2 (may be null)
template <typename T>
int residual(lb2_ctx* ctx, int64_t n, int nc, const T* AX, int64_t ldax, const T* BX, int64_t ldbx,
             const real_t<T>* lambda, T* W, int64_t ldw, real_t<T>* sumsq);
// residual norms + drift monitor: out3[j] = ||AX_j - lambda_j BX_j||^2, out3[nc+j] = Re(x_j^H AX_j), out3[2nc+j] = Re(x_j^H BX_j);
// BX == null: B = I
template <typename T>
int residual_monitor(lb2_ctx* ctx, int64_t n, int nc, const T* X, int64_t ldx, const T* AX, int64_t ldax, const T* BX,
                     int64_t ldbx, const real_t<T>* lambda, real_t<T>* out3);
// sumsq[j] = ||X[:,j]||_2^2, j < nc
template <typename T>
int col_sumsq(lb2_ctx* ctx, int64_t n, int nc, const T* X, int64_t ldx, real_t<T>* sumsq);
// out[0] = sum_j sumsq[j]  (Frobenius^2) on device
template <typename R>
int sum_reals(lb2_ctx* ctx, int nc, const R* v, R* out);
template <typename T>
int fill_uniform(lb2_ctx* ctx, int64_t n, int nc, T* X, int64_t ldx, uint64_t seed, int64_t n_global,
                 int64_t row0);  // X[i,j] <- uniform(seed, counter = j*n_global + row0 + i)
// X[:,j] *= s[j] (s on device) or X *= s0 (s == null)
template <typename T>
int scale_cols(lb2_ctx* ctx, int64_t n, int nc, T* X, int64_t ldx, const real_t<T>* s, real_t<T> s0);
// x *= 1/sqrt(sumsq[0]) when sumsq[0] > 0 (power iteration step, estimate_norm)
template <typename T>
int normalize_by(lb2_ctx* ctx, int64_t n, T* x, const real_t<T>* sumsq);
template <typename T>
int copy_block(lb2_ctx* ctx, int64_t n, int nc, const T* X, int64_t ldx, T* Y, int64_t ldy);

// Chebyshev preconditioner steps: d = x/theta, y = d;  r_out = r_in - AD, d = c1 d + c2 r_out, y += d.
// Din, Dout, AD, Rout share the leading dimension ldw.
template <typename T>
int cheb_init(lb2_ctx* ctx, int64_t n, int nc, const T* X, int64_t ldx, T* D, T* Y, int64_t ldy, int64_t ldw,
              real_t<T> inv_theta);
template <typename T>
int cheb_update(lb2_ctx* ctx, int64_t n, int nc, const T* AD, const T* Rin, int64_t ldrin, T* Rout, const T* Din, T* Dout,
                int64_t ldw, T* Y, int64_t ldy, real_t<T> c1, real_t<T> c2, bool write_r);

// Y = (TO) X, block-wise (mixed-precision preconditioner)
template <typename TO, typename TI>
int convert_block(lb2_ctx* ctx, int64_t n, int nc, const TI* X, int64_t ldx, TO* Y, int64_t ldy);

// ---- spmm.cu ----------------------------------------------------------------------------------------
struct StencilDesc {
  int gx, gy, gz;          // local grid (gz = local planes of the z-slab)
  double cdiag, coff;      // y = (cdiag + v + shift) x + coff * sum(neighbours)
  double shift;            // extra diagonal shift (BdG: K + shift)
  const void* potential;   // real_t<T>[n_local] or null
  const void* halo_lo;     // T plane (gx*gy) below z=0 per column, or null (Dirichlet)
  const void* halo_hi;     // T plane above z=gz-1 per column, or null
  int64_t halo_ld;         // column stride of the halo planes
  int bdg;                 // 1: operator acts on [u;v] (n = 2*gx*gy*gz), coupling d / conj(d)
  double dre, dim;
};
template <typename T>
int spmm_stencil(lb2_ctx* ctx, const StencilDesc& d, int nc, const T* X, int64_t ldx, T* Y, int64_t ldy);
// Chebyshev-step epilogue of the stencil kernel (preconditioner T = p(A)): see spmm.cu
template <typename T>
struct ChebEpilogue {
  const T* rin = nullptr;   // residual block read (x on the first step), leading dimension ldrin
  int64_t ldrin = 0;
  T* rout = nullptr;        // residual block written when write_r (leading dimension ldw)
  T* dout = nullptr;        // new search direction block (leading dimension ldw)
  int64_t ldw = 0;
  real_t<T> c1 = 0, c2 = 0;
  int write_r = 0;
};
template <typename T>
int spmm_stencil_cheb(lb2_ctx* ctx, const StencilDesc& d, int nc, const T* Din, int64_t ldd, T* Yacc, int64_t ldy,
                      const ChebEpilogue<T>& ep);
// neighbour blocks of a row-partitioned CSR apply: the same block vector in the lower / upper neighbour's arena (device
// pointers mapped through CUDA IPC, column stride = that rank's local row count); null where there is no neighbour
struct CsrHalo {
  const void* lo = nullptr;
  const void* hi = nullptr;
  int64_t ld_lo = 0, ld_hi = 0;
};
template <typename T>
int spmm_csr(lb2_ctx* ctx, int64_t n, const int64_t* rowptr, const int32_t* col, const T* val, int nc,
             const T* X, int64_t ldx, T* Y, int64_t ldy, const CsrHalo* halo = nullptr, int64_t nnz = -1);
// CSR with a shared-memory window of X around the diagonal (spmm.cu: csr_win_kernel); H = half-width in rows (<= 256)
template <typename T>
int spmm_csr_window(lb2_ctx* ctx, int64_t n, const int64_t* rowptr, const int32_t* col, const T* val, int nc,
                    const T* X, int64_t ldx, T* Y, int64_t ldy, int H);
template <typename T>
int spmm_diag(lb2_ctx* ctx, int64_t n, const real_t<T>* d, int nc, const T* X, int64_t ldx, T* Y, int64_t ldy);

}  // namespace lb2

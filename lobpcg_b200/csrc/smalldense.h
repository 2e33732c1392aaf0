// lobpcg_b200/csrc/smalldense.h — device-resident small dense helpers (see smalldense.cu).
#pragma once
#include "common.cuh"
#include "context.h"

namespace lb2 {
template <typename T> struct ComplexOf { using type = T; };   // complex type of the same precision
template <> struct ComplexOf<float> { using type = c32; };
template <> struct ComplexOf<double> { using type = c64; };
int sd_init(lb2_ctx* ctx);
template <typename T> int sd_potrf_upper(lb2_ctx* ctx, int m, T* A, int lda, int* h_info);
template <typename T> int sd_syevd_upper(lb2_ctx* ctx, int m, T* A, int lda, real_t<T>* w, int* h_info);
template <typename T> int sd_qr_q(lb2_ctx* ctx, int rows, int cols, T* A, int lda, T* tau);
template <typename T> int sd_gemm(lb2_ctx* ctx, char opa, int m, int n, int k, const T* A, int lda, const T* B, int ldb, T* C, int ldc);
template <typename T> int sd_trsm_run(lb2_ctx* ctx, int rows, int m, const T* R, int ldr, T* X, int ldx);
template <typename T> int sd_dscale(lb2_ctx* ctx, int m, T* G, int ldg, real_t<T>* D);
template <typename T> int sd_set_diag(lb2_ctx* ctx, int m, T* M, int ldm, const real_t<T>* D);
template <typename T> int sd_rcond(lb2_ctx* ctx, int m, const T* Rm, int ldr, const T* DinvR, int ldd, const real_t<T>* D, real_t<T>* out_dev);
template <typename T> int sd_transpose(lb2_ctx* ctx, int rows, int cols, const T* In, int ldi, T* Out, int ldo);
template <typename T> int sd_svqb_transform(lb2_ctx* ctx, int m, const T* V, int ldv, const real_t<T>* lam, const real_t<T>* D, real_t<T> tau, int drop, T* Tm, int ldt, int* count_dev);
template <typename T> int sd_ortho_err_upper(lb2_ctx* ctx, int m, const T* G, int ldg, real_t<T>* out_dev);
template <typename T> int sd_frob(lb2_ctx* ctx, int rows, int cols, const T* G, int ldg, real_t<T>* out_dev);
template <typename T> int sd_trsm_upper(lb2_ctx* ctx, char side, char op, int rows, int cols, const T* R, int ldr, T* X, int ldx);
template <typename T> int sd_gemm_ab(lb2_ctx* ctx, char opa, int m, int n, int k, T alpha, const T* A, int lda, const T* B, int ldb, T beta, T* C, int ldc);
template <typename T> int sd_indef_finalize(lb2_ctx* ctx, int m, const real_t<T>* mu, const T* V, int ldv, T* VR, int ldo, real_t<T>* theta, int8_t* sig);
template <typename T> int sd_cp_lower(lb2_ctx* ctx, int m, int nx, const T* Cx, T* Cp);
// general projected pencil (smalldense.cu: "general (non-definite) projected pencil")
template <typename T> int sd_lu_solve(lb2_ctx* ctx, int m, T* A, int lda, T* B, int ldb, int nrhs, int64_t* ipiv, int* h_info);
template <typename T> int sd_geev(lb2_ctx* ctx, int m, const T* M, int ldm, void* Mc, void* W, void* VR, int* h_info);
template <typename T> int sd_geev_extract(lb2_ctx* ctx, int m, const void* W, const void* VR, real_t<T>* theta, T* V, int ldv);
template <typename T> int sd_bnormalize(lb2_ctx* ctx, int m, T* V, int ldv, const T* E, int lde, int8_t* sig);
template <typename T> int sd_indef_quality(lb2_ctx* ctx, int m, const T* E, const T* V, const T* GV, real_t<T>* out_dev);
template <typename T> int sd_indef_sort(lb2_ctx* ctx, int m, const real_t<T>* theta, const int8_t* sig, const T* V, int ldv, T* Vout, int ldo, real_t<T>* theta_out, int8_t* sig_out);
template <typename T> int sd_assemble_gram(lb2_ctx* ctx, int m, int mxp, const T* Gc, int ldc, const T* Gw, int ldw, T* G, int ldg);
}  // namespace lb2

/* oracle/shim/cblas.h — minimal CBLAS prototypes (public netlib CBLAS API) for building the
 * reference's sources UNMODIFIED against the OpenBLAS shipped inside the scipy wheel
 * (symbols carry a scipy_ prefix there; see scipy_prefix.h).  TEST INFRASTRUCTURE ONLY.
 * Only the entry points include/lobpcg/blas_wrapper.h (reference) uses are declared. */
#ifndef ORACLE_SHIM_CBLAS_H
#define ORACLE_SHIM_CBLAS_H
#include "scipy_prefix.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef enum { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO;
typedef enum { CblasNonUnit = 131, CblasUnit = 132 } CBLAS_DIAG;
typedef enum { CblasLeft = 141, CblasRight = 142 } CBLAS_SIDE;
typedef CBLAS_ORDER CBLAS_LAYOUT;

float  cblas_snrm2(int n, const float *x, int incx);
double cblas_dnrm2(int n, const double *x, int incx);
float  cblas_scnrm2(int n, const void *x, int incx);
double cblas_dznrm2(int n, const void *x, int incx);
float  cblas_sdot(int n, const float *x, int incx, const float *y, int incy);
double cblas_ddot(int n, const double *x, int incx, const double *y, int incy);
void cblas_cdotc_sub(int n, const void *x, int incx, const void *y, int incy, void *ret);
void cblas_zdotc_sub(int n, const void *x, int incx, const void *y, int incy, void *ret);
void cblas_saxpy(int n, float alpha, const float *x, int incx, float *y, int incy);
void cblas_sscal(int n, float alpha, float *x, int incx);
void cblas_scopy(int n, const float *x, int incx, float *y, int incy);
void cblas_sgemm(CBLAS_ORDER o, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int m, int n, int k, float alpha, const float *A, int lda, const float *B, int ldb, float beta, float *C, int ldc);
void cblas_strsm(CBLAS_ORDER o, CBLAS_SIDE s, CBLAS_UPLO u, CBLAS_TRANSPOSE ta, CBLAS_DIAG d, int m, int n, float alpha, const float *A, int lda, float *B, int ldb);
void cblas_daxpy(int n, double alpha, const double *x, int incx, double *y, int incy);
void cblas_dscal(int n, double alpha, double *x, int incx);
void cblas_dcopy(int n, const double *x, int incx, double *y, int incy);
void cblas_dgemm(CBLAS_ORDER o, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb, double beta, double *C, int ldc);
void cblas_dtrsm(CBLAS_ORDER o, CBLAS_SIDE s, CBLAS_UPLO u, CBLAS_TRANSPOSE ta, CBLAS_DIAG d, int m, int n, double alpha, const double *A, int lda, double *B, int ldb);
void cblas_caxpy(int n, const void *alpha, const void *x, int incx, void *y, int incy);
void cblas_cscal(int n, const void *alpha, void *x, int incx);
void cblas_ccopy(int n, const void *x, int incx, void *y, int incy);
void cblas_cgemm(CBLAS_ORDER o, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int m, int n, int k, const void *alpha, const void *A, int lda, const void *B, int ldb, const void *beta, void *C, int ldc);
void cblas_ctrsm(CBLAS_ORDER o, CBLAS_SIDE s, CBLAS_UPLO u, CBLAS_TRANSPOSE ta, CBLAS_DIAG d, int m, int n, const void *alpha, const void *A, int lda, void *B, int ldb);
void cblas_zaxpy(int n, const void *alpha, const void *x, int incx, void *y, int incy);
void cblas_zscal(int n, const void *alpha, void *x, int incx);
void cblas_zcopy(int n, const void *x, int incx, void *y, int incy);
void cblas_zgemm(CBLAS_ORDER o, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int m, int n, int k, const void *alpha, const void *A, int lda, const void *B, int ldb, const void *beta, void *C, int ldc);
void cblas_ztrsm(CBLAS_ORDER o, CBLAS_SIDE s, CBLAS_UPLO u, CBLAS_TRANSPOSE ta, CBLAS_DIAG d, int m, int n, const void *alpha, const void *A, int lda, void *B, int ldb);
void cblas_ssyrk(CBLAS_ORDER o, CBLAS_UPLO u, CBLAS_TRANSPOSE t, int n, int k, float alpha, const float *A, int lda, float beta, float *C, int ldc);
void cblas_dsyrk(CBLAS_ORDER o, CBLAS_UPLO u, CBLAS_TRANSPOSE t, int n, int k, double alpha, const double *A, int lda, double beta, double *C, int ldc);
void cblas_cherk(CBLAS_ORDER o, CBLAS_UPLO u, CBLAS_TRANSPOSE t, int n, int k, float alpha, const void *A, int lda, float beta, void *C, int ldc);
void cblas_zherk(CBLAS_ORDER o, CBLAS_UPLO u, CBLAS_TRANSPOSE t, int n, int k, double alpha, const void *A, int lda, double beta, void *C, int ldc);
void openblas_set_num_threads(int n);
int openblas_get_num_threads(void);
char *openblas_get_config(void);
#ifdef __cplusplus
}
#endif
#endif

/* oracle/shim/lapacke.h — minimal LAPACKE prototypes (public LAPACKE C API) — see cblas.h.
 * TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_LAPACKE_H
#define ORACLE_SHIM_LAPACKE_H
#include <complex.h>
#include "scipy_prefix.h"
#ifdef __cplusplus
extern "C" {
#endif
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
typedef int lapack_int;
typedef float _Complex lapack_complex_float;
typedef double _Complex lapack_complex_double;

lapack_int LAPACKE_spotrf(int layout, char uplo, lapack_int n, float *a, lapack_int lda);
lapack_int LAPACKE_strcon(int layout, char norm, char uplo, char diag, lapack_int n, const float *a, lapack_int lda, float *rcond);
lapack_int LAPACKE_sgeqrf(int layout, lapack_int m, lapack_int n, float *a, lapack_int lda, float *tau);
lapack_int LAPACKE_ssyev(int layout, char jobz, char uplo, lapack_int n, float *a, lapack_int lda, float *w);
lapack_int LAPACKE_sorgqr(int layout, lapack_int m, lapack_int n, lapack_int k, float *a, lapack_int lda, const float *tau);
lapack_int LAPACKE_sgeev(int layout, char jobvl, char jobvr, lapack_int n, float *a, lapack_int lda, float *wr, float *wi, float *vl, lapack_int ldvl, float *vr, lapack_int ldvr);
lapack_int LAPACKE_sggev(int layout, char jobvl, char jobvr, lapack_int n, float *a, lapack_int lda, float *b, lapack_int ldb, float *alphar, float *alphai, float *beta, float *vl, lapack_int ldvl, float *vr, lapack_int ldvr);
lapack_int LAPACKE_dpotrf(int layout, char uplo, lapack_int n, double *a, lapack_int lda);
lapack_int LAPACKE_dtrcon(int layout, char norm, char uplo, char diag, lapack_int n, const double *a, lapack_int lda, double *rcond);
lapack_int LAPACKE_dgeqrf(int layout, lapack_int m, lapack_int n, double *a, lapack_int lda, double *tau);
lapack_int LAPACKE_dsyev(int layout, char jobz, char uplo, lapack_int n, double *a, lapack_int lda, double *w);
lapack_int LAPACKE_dorgqr(int layout, lapack_int m, lapack_int n, lapack_int k, double *a, lapack_int lda, const double *tau);
lapack_int LAPACKE_dgeev(int layout, char jobvl, char jobvr, lapack_int n, double *a, lapack_int lda, double *wr, double *wi, double *vl, lapack_int ldvl, double *vr, lapack_int ldvr);
lapack_int LAPACKE_dggev(int layout, char jobvl, char jobvr, lapack_int n, double *a, lapack_int lda, double *b, lapack_int ldb, double *alphar, double *alphai, double *beta, double *vl, lapack_int ldvl, double *vr, lapack_int ldvr);
lapack_int LAPACKE_cpotrf(int layout, char uplo, lapack_int n, lapack_complex_float *a, lapack_int lda);
lapack_int LAPACKE_ctrcon(int layout, char norm, char uplo, char diag, lapack_int n, const lapack_complex_float *a, lapack_int lda, float *rcond);
lapack_int LAPACKE_cgeqrf(int layout, lapack_int m, lapack_int n, lapack_complex_float *a, lapack_int lda, lapack_complex_float *tau);
lapack_int LAPACKE_cheev(int layout, char jobz, char uplo, lapack_int n, lapack_complex_float *a, lapack_int lda, float *w);
lapack_int LAPACKE_cungqr(int layout, lapack_int m, lapack_int n, lapack_int k, lapack_complex_float *a, lapack_int lda, const lapack_complex_float *tau);
lapack_int LAPACKE_cgeev(int layout, char jobvl, char jobvr, lapack_int n, lapack_complex_float *a, lapack_int lda, lapack_complex_float *w, lapack_complex_float *vl, lapack_int ldvl, lapack_complex_float *vr, lapack_int ldvr);
lapack_int LAPACKE_cggev(int layout, char jobvl, char jobvr, lapack_int n, lapack_complex_float *a, lapack_int lda, lapack_complex_float *b, lapack_int ldb, lapack_complex_float *alpha, lapack_complex_float *beta, lapack_complex_float *vl, lapack_int ldvl, lapack_complex_float *vr, lapack_int ldvr);
lapack_int LAPACKE_zpotrf(int layout, char uplo, lapack_int n, lapack_complex_double *a, lapack_int lda);
lapack_int LAPACKE_ztrcon(int layout, char norm, char uplo, char diag, lapack_int n, const lapack_complex_double *a, lapack_int lda, double *rcond);
lapack_int LAPACKE_zgeqrf(int layout, lapack_int m, lapack_int n, lapack_complex_double *a, lapack_int lda, lapack_complex_double *tau);
lapack_int LAPACKE_zheev(int layout, char jobz, char uplo, lapack_int n, lapack_complex_double *a, lapack_int lda, double *w);
lapack_int LAPACKE_zungqr(int layout, lapack_int m, lapack_int n, lapack_int k, lapack_complex_double *a, lapack_int lda, const lapack_complex_double *tau);
lapack_int LAPACKE_zgeev(int layout, char jobvl, char jobvr, lapack_int n, lapack_complex_double *a, lapack_int lda, lapack_complex_double *w, lapack_complex_double *vl, lapack_int ldvl, lapack_complex_double *vr, lapack_int ldvr);
lapack_int LAPACKE_zggev(int layout, char jobvl, char jobvr, lapack_int n, lapack_complex_double *a, lapack_int lda, lapack_complex_double *b, lapack_int ldb, lapack_complex_double *alpha, lapack_complex_double *beta, lapack_complex_double *vl, lapack_int ldvl, lapack_complex_double *vr, lapack_int ldvr);
#ifdef __cplusplus
}
#endif
#endif

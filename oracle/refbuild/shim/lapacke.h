/* Prototype-only LAPACKE shim (ours); see cblas.h in this directory.
 * TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_LAPACKE_H
#define ORACLE_SHIM_LAPACKE_H
#include <complex.h>
#ifdef __cplusplus
extern "C" {
#endif
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
typedef int lapack_int;
typedef double _Complex lapack_complex_double;

double LAPACKE_dlange(int layout, char norm, lapack_int m, lapack_int n, const double *a, lapack_int lda);
lapack_int LAPACKE_dgesv(int layout, lapack_int n, lapack_int nrhs, double *a, lapack_int lda, lapack_int *ipiv, double *b, lapack_int ldb);
lapack_int LAPACKE_dgeev(int layout, char jobvl, char jobvr, lapack_int n, double *a, lapack_int lda, double *wr, double *wi, double *vl, lapack_int ldvl, double *vr, lapack_int ldvr);
lapack_int LAPACKE_dgesvd(int layout, char jobu, char jobvt, lapack_int m, lapack_int n, double *a, lapack_int lda, double *s, double *u, lapack_int ldu, double *vt, lapack_int ldvt, double *superb);
lapack_int LAPACKE_dgeqrf(int layout, lapack_int m, lapack_int n, double *a, lapack_int lda, double *tau);
lapack_int LAPACKE_dormqr(int layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k, const double *a, lapack_int lda, const double *tau, double *c, lapack_int ldc);
lapack_int LAPACKE_dgetrf(int layout, lapack_int m, lapack_int n, double *a, lapack_int lda, lapack_int *ipiv);
lapack_int LAPACKE_dgetri(int layout, lapack_int n, double *a, lapack_int lda, const lapack_int *ipiv);

double LAPACKE_zlange(int layout, char norm, lapack_int m, lapack_int n, const lapack_complex_double *a, lapack_int lda);
lapack_int LAPACKE_zgesv(int layout, lapack_int n, lapack_int nrhs, lapack_complex_double *a, lapack_int lda, lapack_int *ipiv, lapack_complex_double *b, lapack_int ldb);
lapack_int LAPACKE_zgeev(int layout, char jobvl, char jobvr, lapack_int n, lapack_complex_double *a, lapack_int lda, lapack_complex_double *w, lapack_complex_double *vl, lapack_int ldvl, lapack_complex_double *vr, lapack_int ldvr);
lapack_int LAPACKE_zgesvd(int layout, char jobu, char jobvt, lapack_int m, lapack_int n, lapack_complex_double *a, lapack_int lda, double *s, lapack_complex_double *u, lapack_int ldu, lapack_complex_double *vt, lapack_int ldvt, double *superb);
lapack_int LAPACKE_zgeqrf(int layout, lapack_int m, lapack_int n, lapack_complex_double *a, lapack_int lda, lapack_complex_double *tau);
lapack_int LAPACKE_zunmqr(int layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k, const lapack_complex_double *a, lapack_int lda, const lapack_complex_double *tau, lapack_complex_double *c, lapack_int ldc);
lapack_int LAPACKE_zgetrf(int layout, lapack_int m, lapack_int n, lapack_complex_double *a, lapack_int lda, lapack_int *ipiv);
lapack_int LAPACKE_zgetri(int layout, lapack_int n, lapack_complex_double *a, lapack_int lda, const lapack_int *ipiv);
#ifdef __cplusplus
}
#endif
#endif

/* Prototype-only CBLAS shim (ours) so that the unmodified reference sources
 * under /root/reference/src compile without a system cblas.h.  The symbols are
 * resolved at link time by the OpenBLAS that ships in the image (see
 * oracle/refbuild/build_ref.sh).  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_CBLAS_H
#define ORACLE_SHIM_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
typedef enum CBLAS_ORDER CBLAS_LAYOUT;

double cblas_ddot(const int n, const double *x, const int incx, const double *y, const int incy);
double cblas_dnrm2(const int n, const double *x, const int incx);
void cblas_daxpy(const int n, const double alpha, const double *x, const int incx, double *y, const int incy);
void cblas_dcopy(const int n, const double *x, const int incx, double *y, const int incy);
void cblas_dscal(const int n, const double alpha, double *x, const int incx);
void cblas_dgemm(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const enum CBLAS_TRANSPOSE tb,
                 const int m, const int n, const int k, const double alpha, const double *a, const int lda,
                 const double *b, const int ldb, const double beta, double *c, const int ldc);
void cblas_dger(const enum CBLAS_ORDER order, const int m, const int n, const double alpha,
                const double *x, const int incx, const double *y, const int incy, double *a, const int lda);

void cblas_zscal(const int n, const void *alpha, void *x, const int incx);
void cblas_zgeru(const enum CBLAS_ORDER order, const int m, const int n, const void *alpha,
                 const void *x, const int incx, const void *y, const int incy, void *a, const int lda);
void cblas_zgemm(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const enum CBLAS_TRANSPOSE tb,
                 const int m, const int n, const int k, const void *alpha, const void *a, const int lda,
                 const void *b, const int ldb, const void *beta, void *c, const int ldc);
void cblas_zdotu_sub(const int n, const void *x, const int incx, const void *y, const int incy, void *ret);
void cblas_zdotc_sub(const int n, const void *x, const int incx, const void *y, const int incy, void *ret);
#ifdef __cplusplus
}
#endif
#endif

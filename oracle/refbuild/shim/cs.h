/* Minimal CSparse-compatible interface (ours).  The reference links against
 * SuiteSparse CXSparse, which is not installed in this image and is not part
 * of /root/reference; cs_stub.c in this directory provides from-scratch
 * implementations of the entry points the mesh/functional path needs
 * (transpose, gaxpy, add, multiply, alloc helpers).  The direct solvers are
 * stubs that report failure: they are off the functional hot path.
 * TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_CS_H
#define ORACLE_SHIM_CS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
#define CS_INT int
typedef struct cs_sparse {
    int nzmax;   /* maximum number of entries */
    int m;       /* rows */
    int n;       /* columns */
    int *p;      /* column pointers (n+1) */
    int *i;      /* row indices */
    double *x;   /* values or NULL */
    int nz;      /* -1 for compressed-column */
} cs;

void *cs_free(void *p);
void *cs_realloc(void *p, int n, size_t size, int *ok);
cs *cs_transpose(const cs *A, int values);
cs *cs_add(const cs *A, const cs *B, double alpha, double beta);
cs *cs_multiply(const cs *A, const cs *B);
int cs_gaxpy(const cs *A, const double *x, double *y);
int cs_lusol(int order, const cs *A, double *b, double tol);
int cs_qrsol(int order, const cs *A, double *b);
#ifdef __cplusplus
}
#endif
#endif

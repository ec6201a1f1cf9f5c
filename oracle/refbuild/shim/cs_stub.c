/* From-scratch implementations of the handful of CSparse entry points the
 * reference's mesh/functional path calls (see cs.h here).  Compressed-column
 * only.  TEST INFRASTRUCTURE ONLY. */
#include <stdlib.h>
#include <string.h>
#include "cs.h"

void *cs_free(void *p) { if (p) free(p); return NULL; }

void *cs_realloc(void *p, int n, size_t size, int *ok) {
    size_t bytes = (size_t)(n > 1 ? n : 1) * (size > 1 ? size : 1);
    void *q = realloc(p, bytes);
    *ok = (q != NULL);
    return q ? q : p;
}

static cs *cs_new(int m, int n, int nzmax, int values) {
    cs *A = calloc(1, sizeof(cs));
    if (!A) return NULL;
    A->m = m; A->n = n; A->nzmax = (nzmax > 1 ? nzmax : 1); A->nz = -1;
    A->p = malloc(sizeof(int) * (size_t)(n + 1));
    A->i = malloc(sizeof(int) * (size_t)A->nzmax);
    A->x = values ? malloc(sizeof(double) * (size_t)A->nzmax) : NULL;
    if (!A->p || !A->i || (values && !A->x)) { free(A->p); free(A->i); free(A->x); free(A); return NULL; }
    return A;
}

/* Stable counting-sort transpose: entries of each output column come out in
 * ascending source-column order. */
cs *cs_transpose(const cs *A, int values) {
    int m = A->m, n = A->n, nnz = A->p[n];
    int withx = values && A->x;
    cs *C = cs_new(n, m, nnz, withx);
    if (!C) return NULL;
    int *w = calloc((size_t)(m > 1 ? m : 1), sizeof(int));
    if (!w) { free(C->p); free(C->i); free(C->x); free(C); return NULL; }
    for (int k = 0; k < nnz; k++) w[A->i[k]]++;
    int run = 0;
    for (int r = 0; r < m; r++) { C->p[r] = run; int c = w[r]; w[r] = run; run += c; }
    C->p[m] = run;
    for (int j = 0; j < n; j++) {
        for (int k = A->p[j]; k < A->p[j + 1]; k++) {
            int q = w[A->i[k]]++;
            C->i[q] = j;
            if (withx) C->x[q] = A->x[k];
        }
    }
    free(w);
    return C;
}

int cs_gaxpy(const cs *A, const double *x, double *y) {
    if (!A || !x || !y || !A->x) return 0;
    for (int j = 0; j < A->n; j++)
        for (int k = A->p[j]; k < A->p[j + 1]; k++) y[A->i[k]] += A->x[k] * x[j];
    return 1;
}

/* x += beta*A(:,j) scattered into a dense work vector, recording new rows in C */
static int scatter_col(const cs *A, int j, double beta, int *w, double *x, int mark, cs *C, int nz) {
    for (int k = A->p[j]; k < A->p[j + 1]; k++) {
        int r = A->i[k];
        double v = A->x ? A->x[k] : 1.0;
        if (w[r] < mark) { w[r] = mark; C->i[nz++] = r; if (x) x[r] = beta * v; }
        else if (x) x[r] += beta * v;
    }
    return nz;
}

static int grow(cs *C, int need) {
    if (need <= C->nzmax) return 1;
    int nzmax = 2 * C->nzmax + need;
    int *ni = realloc(C->i, sizeof(int) * (size_t)nzmax);
    if (!ni) return 0;
    C->i = ni;
    if (C->x) { double *nx = realloc(C->x, sizeof(double) * (size_t)nzmax); if (!nx) return 0; C->x = nx; }
    C->nzmax = nzmax;
    return 1;
}

cs *cs_add(const cs *A, const cs *B, double alpha, double beta) {
    if (!A || !B || A->m != B->m || A->n != B->n) return NULL;
    int m = A->m, n = A->n;
    cs *C = cs_new(m, n, A->p[n] + B->p[n], 1);
    int *w = calloc((size_t)(m > 1 ? m : 1), sizeof(int));
    double *x = malloc(sizeof(double) * (size_t)(m > 1 ? m : 1));
    if (!C || !w || !x) { free(w); free(x); return NULL; }
    int nz = 0;
    for (int j = 0; j < n; j++) {
        C->p[j] = nz;
        nz = scatter_col(A, j, alpha, w, x, j + 1, C, nz);
        nz = scatter_col(B, j, beta, w, x, j + 1, C, nz);
        for (int k = C->p[j]; k < nz; k++) C->x[k] = x[C->i[k]];
    }
    C->p[n] = nz;
    free(w); free(x);
    return C;
}

cs *cs_multiply(const cs *A, const cs *B) {
    if (!A || !B || A->n != B->m) return NULL;
    int m = A->m, n = B->n;
    cs *C = cs_new(m, n, A->p[A->n] + B->p[B->n], 1);
    int *w = calloc((size_t)(m > 1 ? m : 1), sizeof(int));
    double *x = malloc(sizeof(double) * (size_t)(m > 1 ? m : 1));
    if (!C || !w || !x) { free(w); free(x); return NULL; }
    int nz = 0;
    for (int j = 0; j < n; j++) {
        if (!grow(C, nz + m)) { free(w); free(x); return NULL; }
        C->p[j] = nz;
        for (int k = B->p[j]; k < B->p[j + 1]; k++)
            nz = scatter_col(A, B->i[k], B->x ? B->x[k] : 1.0, w, x, j + 1, C, nz);
        for (int k = C->p[j]; k < nz; k++) C->x[k] = x[C->i[k]];
    }
    C->p[n] = nz;
    free(w); free(x);
    return C;
}

int cs_lusol(int order, const cs *A, double *b, double tol) { (void)order; (void)A; (void)b; (void)tol; return 0; }
int cs_qrsol(int order, const cs *A, double *b) { (void)order; (void)A; (void)b; return 0; }

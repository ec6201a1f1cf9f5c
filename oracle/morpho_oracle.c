/* morpho_oracle.c — CPU ORACLE (plain C restatement).  TEST INFRASTRUCTURE ONLY.
 * See morpho_oracle.h for scope, pinning and the rules on who may load this.
 *
 * Every routine names the reference function (Morpho-lang/morpho 0.6.4,
 * paths relative to /root/reference/) whose serial algorithm it restates.
 * Operation ORDER is kept identical to the reference wherever floating-point
 * results depend on it (element loop order, Kahan summation, FD stencils,
 * accumulation into the gradient).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "morpho_oracle.h"

#define MO_EPS DBL_EPSILON /* MORPHO_EPS, src/build.h:50 */
#define MO_MAXNV 4
#define MO_MAXPSIZE 16

/* ------------------------------------------------------------------ *
 * Level-1 BLAS / small LAPACK as the reference calls them
 * ------------------------------------------------------------------ */

/* functional_vecdot → cblas_ddot (functional.c:1885-1887): netlib order, left to right. */
static double vecdot(int n, const double *a, const double *b) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* functional_vecnorm → cblas_dnrm2 (functional.c:1880-1882).  The x86-64 OpenBLAS
 * kernel accumulates the sum of squares in 80-bit extended precision and takes
 * one square root (verified bit-for-bit against OpenBLAS 0.3.15 on 60k random
 * vectors); restated with long double. */
static double vecnorm(int n, const double *a) {
    long double s = 0.0L;
    for (int i = 0; i < n; i++) s += (long double) a[i] * (long double) a[i];
    return (double) sqrtl(s);
}

static void vecsub(int n, const double *a, const double *b, double *out) { for (int i = 0; i < n; i++) out[i] = a[i] - b[i]; }
static void vecadd(int n, const double *a, const double *b, double *out) { for (int i = 0; i < n; i++) out[i] = a[i] + b[i]; }
static void vecscale(int n, double l, const double *a, double *out) { for (int i = 0; i < n; i++) out[i] = l * a[i]; }
static void vecaddscale(int n, const double *a, double l, const double *b, double *out) { for (int i = 0; i < n; i++) out[i] = a[i] + l * b[i]; }

/* functional_veccross (functional.c:1890-1894) */
static void veccross(const double *a, const double *b, double *out) {
    out[0] = a[1] * b[2] - a[2] * b[1];
    out[1] = a[2] * b[0] - a[0] * b[2];
    out[2] = a[0] * b[1] - a[1] * b[0];
}

/* matrix_addtocolumnptr → cblas_daxpy (matrix.c:478-483): y += alpha*x, no FMA */
static void addtocolumn(double *frc, int dim, int col, double alpha, const double *x) {
    volatile double t;
    for (int k = 0; k < dim; k++) { t = alpha * x[k]; frc[(size_t) col * dim + k] += t; }
}

/* matrix_solvesmall → LAPACKE_dgesv (matrix.c:129-139, 638-644): unblocked LU with
 * partial pivoting (dgetf2) followed by forward/back substitution (dgetrs), column major. */
static int solvesmall(int n, double *a, int nrhs, double *b) {
    int piv[MO_MAXNV];
    for (int j = 0; j < n; j++) {
        int p = j; double big = fabs(a[j + j * n]);
        for (int i = j + 1; i < n; i++) if (fabs(a[i + j * n]) > big) { big = fabs(a[i + j * n]); p = i; }
        piv[j] = p;
        if (a[p + j * n] == 0.0) return 0;
        if (p != j) for (int c = 0; c < n; c++) { double t = a[j + c * n]; a[j + c * n] = a[p + c * n]; a[p + c * n] = t; }
        double r = 1.0 / a[j + j * n];
        for (int i = j + 1; i < n; i++) a[i + j * n] *= r;
        for (int c = j + 1; c < n; c++)
            for (int i = j + 1; i < n; i++) a[i + c * n] -= a[i + j * n] * a[j + c * n];
    }
    for (int r = 0; r < nrhs; r++) {
        double *x = b + (size_t) r * n;
        for (int j = 0; j < n; j++) if (piv[j] != j) { double t = x[j]; x[j] = x[piv[j]]; x[piv[j]] = t; }
        for (int j = 0; j < n; j++) for (int i = j + 1; i < n; i++) x[i] -= x[j] * a[i + j * n];
        for (int j = n - 1; j >= 0; j--) { x[j] /= a[j + j * n]; for (int i = 0; i < j; i++) x[i] -= x[j] * a[i + j * n]; }
    }
    return 1;
}

/* ------------------------------------------------------------------ *
 * functional_fdstepsize (functional.c:42-58); fddelta1 = cbrt(eps) set at
 * functional.c:6230, fddelta2 = eps^(1/4) at 6231.
 * ------------------------------------------------------------------ */
double mo_fdstepsize(double x, int order) {
    double h = pow(MO_EPS, 1.0 / 3.0);
    if (order == 2) h = pow(MO_EPS, 1.0 / 4.0);
    double absx = fabs(x);
    if (absx > 1) h *= absx;
    volatile double temp = x + h;
    return temp - x;
}

/* ------------------------------------------------------------------ *
 * Evaluation context
 * ------------------------------------------------------------------ */
typedef struct {
    const mo_mesh *m;
    const mo_functional *f;
    double *vert;        /* working copy of the vertex matrix (FD perturbs and restores it) */
    int g;               /* grade looped over */
    /* LineCurvatureSq: conn(1,0) transpose of the line elements */
    int *ltptr, *ltidx;
    /* EquiElement: grade of the elements compared, conn(grade,0) transpose, mean weight */
    int eg, *etptr, *etidx;
    double wmean;
} ctx;

static const double *X(const ctx *c, int v) { return c->vert + (size_t) v * c->m->dim; }

/* ------------------------------------------------------------------ *
 * Per-element integrands
 * ------------------------------------------------------------------ */

/* length_integrand (functional.c:1948-1958) */
static int length_integrand(const ctx *c, int nv, const int *vid, double *out) {
    if (nv != 2) return 0;
    int dim = c->m->dim; double s0[3];
    vecsub(dim, X(c, vid[1]), X(c, vid[0]), s0);
    *out = vecnorm(dim, s0);
    return 1;
}

/* areaenclosed_integrand (functional.c:2000-2016) */
static int areaenclosed_integrand(const ctx *c, int nv, const int *vid, double *out) {
    int dim = c->m->dim; double cx[3], normcx;
    const double *x0 = X(c, vid[0]), *x1 = X(c, vid[1]);
    (void) nv;
    if (dim == 2) { cx[0] = x0[0] * x1[1] - x0[1] * x1[0]; normcx = fabs(cx[0]); }
    else { veccross(x0, x1, cx); normcx = vecnorm(dim, cx); }
    *out = 0.5 * normcx;
    return 1;
}

/* area_integrand (functional.c:2063-2076): 3-vectors zero padded when dim==2 */
static int area_integrand(const ctx *c, int nv, const int *vid, double *out) {
    if (nv != 3) return 0;
    int dim = c->m->dim; double s0[3] = {0, 0, 0}, s1[3] = {0, 0, 0}, cx[3];
    vecsub(dim, X(c, vid[1]), X(c, vid[0]), s0);
    vecsub(dim, X(c, vid[2]), X(c, vid[1]), s1);
    veccross(s0, s1, cx);
    *out = 0.5 * vecnorm(3, cx);
    return 1;
}

/* volumeenclosed_integrand (functional.c:2131-2139) */
static int volumeenclosed_integrand(const ctx *c, int nv, const int *vid, double *out) {
    int dim = c->m->dim; double cx[3];
    (void) nv;
    veccross(X(c, vid[0]), X(c, vid[1]), cx);
    *out = fabs(vecdot(dim, cx, X(c, vid[2]))) / 6.0;
    return 1;
}

/* volume_integrand (functional.c:2185-2198) */
static int volume_integrand(const ctx *c, int nv, const int *vid, double *out) {
    int dim = c->m->dim; double s10[3], s20[3], s30[3], cx[3];
    (void) nv;
    vecsub(dim, X(c, vid[1]), X(c, vid[0]), s10);
    vecsub(dim, X(c, vid[2]), X(c, vid[0]), s20);
    vecsub(dim, X(c, vid[3]), X(c, vid[0]), s30);
    veccross(s20, s30, cx);
    *out = fabs(vecdot(dim, s10, cx)) / 6.0;
    return 1;
}

/* functional_elementsize (functional.c:1901-1908) */
static int elementsize(const ctx *c, int g, int nv, const int *vid, double *out) {
    switch (g) {
        case 1: return length_integrand(c, nv, vid, out);
        case 2: return area_integrand(c, nv, vid, out);
        case 3: return volume_integrand(c, nv, vid, out);
    }
    return 0;
}

/* linecurvsq_integrand (functional.c:2992-3042); neighbours via mesh_findneighbors
 * (mesh.c:838-883) = edges incident to the vertex in ascending edge id (transpose
 * of conn(0,1)); no symmetries (out of scope). */
static int linecurvsq_integrand(const ctx *c, int id, double *out) {
    int dim = c->m->dim; double result = 0.0, s0[3], s1[3], *s[2] = {s0, s1}, sgn = -1.0;
    int nn = c->ltptr[id + 1] - c->ltptr[id];
    if (nn > 0) {
        if (nn != 2) { *out = 0.0; return 1; }
        for (int i = 0; i < 2; i++) {
            const int *entries = c->m->conn[1] + 2 * (size_t) c->ltidx[c->ltptr[id] + i];
            vecsub(dim, X(c, entries[0]), X(c, entries[1]), s[i]);
            if (!(entries[0] == id)) sgn *= -1;
        }
        double s0s0 = vecdot(dim, s0, s0), s0s1 = vecdot(dim, s0, s1), s1s1 = vecdot(dim, s1, s1);
        s0s0 = sqrt(s0s0); s1s1 = sqrt(s1s1);
        if (s0s0 < MO_EPS || s1s1 < MO_EPS) return 0;
        double u = sgn * s0s1 / s0s0 / s1s1, len = 0.5 * (s0s0 + s1s1);
        if (u < 1) u = acos(u); else u = 0;
        result = u * u / len;
        if (c->f->integrandonly) result /= len;
    }
    *out = result;
    return 1;
}

/* gradsq_computeperpendicular (functional.c:3485-3503) */
static int computeperpendicular(int n, const double *s1, const double *s2, double *out) {
    double s1s2 = vecdot(n, s1, s2), s2s2 = vecdot(n, s2, s2), temp[3];
    if (fabs(s2s2) < MO_EPS) return 0;
    vecscale(n, s1s2 / s2s2, s2, temp);
    vecsub(n, s1, temp, out);
    double sout = vecnorm(n, out);
    if (fabs(sout) < MO_EPS) return 0;
    vecscale(n, 1 / (sout * sout), out, out);
    return 1;
}

/* gradsq_evaluategradient (functional.c:3511-3542): P1 gradient on a triangle */
static int evaluategradient(const ctx *c, const double *fld, int psize, int nv, const int *vid, double *out) {
    int dim = c->m->dim; double s[3][3], t[3][3];
    const double *x[3] = {X(c, vid[0]), X(c, vid[1]), X(c, vid[2])};
    (void) nv;
    vecsub(dim, x[1], x[0], s[0]);
    vecsub(dim, x[2], x[1], s[1]);
    vecsub(dim, x[0], x[2], s[2]);
    computeperpendicular(dim, s[2], s[1], t[0]);
    computeperpendicular(dim, s[0], s[2], t[1]);
    computeperpendicular(dim, s[1], s[0], t[2]);
    for (int i = 0; i < dim * psize; i++) out[i] = 0;
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < psize; i++)
            vecaddscale(dim, &out[i * dim], fld[(size_t) vid[j] * psize + i], t[j], &out[i * dim]);
    return 1;
}

/* gradsq_evaluategradient3d (functional.c:3584-3620): 3x3 solve per tetrahedron */
static int evaluategradient3d(const ctx *c, const double *fld, int psize, int nv, const int *vid, double *out) {
    int dim = c->m->dim; double xarray[9], xtarray[9];
    for (int i = 1; i < nv; i++) vecsub(dim, X(c, vid[i]), X(c, vid[0]), &xarray[(i - 1) * dim]);
    for (int r = 0; r < dim; r++) for (int cc = 0; cc < dim; cc++) xtarray[cc + r * dim] = xarray[r + cc * dim];
    for (int i = 0; i < psize; i++)
        for (int j = 0; j < dim; j++)
            out[i * dim + j] = fld[(size_t) vid[j + 1] * psize + i] - fld[(size_t) vid[0] * psize + i];
    solvesmall(dim, xtarray, psize, out);
    return 1;
}

/* gradsq_integrand (functional.c:3657-3676) */
static int gradsq_integrand(const ctx *c, int nv, const int *vid, double *out) {
    int dim = c->m->dim, psize = c->f->psize[0];
    double size = 0, grad[MO_MAXPSIZE * 3];
    if (!elementsize(c, c->g, nv, vid, &size)) return 0;
    if (c->g == 2) { if (!evaluategradient(c, c->f->field[0], psize, nv, vid, grad)) return 0; }
    else if (c->g == 3) { if (!evaluategradient3d(c, c->f->field[0], psize, nv, vid, grad)) return 0; }
    else return 0;
    double gradnrm = vecnorm(psize * dim, grad);
    *out = gradnrm * gradnrm * size;
    return 1;
}

/* normsq_integrand (functional.c:4149-4160): grade-0 loop, q.q at the vertex */
static int normsq_integrand(const ctx *c, int id, double *out) {
    int psize = c->f->psize[0];
    const double *q = c->f->field[0] + (size_t) id * psize;
    *out = vecdot(psize, q, q);
    return 1;
}

/* nematic_bcintf / nematic_bcintfg (functional.c:3818-3834) */
static double bcintf(int n, const double *f) { double sum = 0; for (int i = 0; i < n; i++) sum += f[i]; return sum / n; }
static double bcintfg(int n, const double *f, const double *g) {
    double sum = 0;
    for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) sum += f[i] * g[j]; sum += f[i] * g[i]; }
    return sum / (n * (n + 1));
}

/* nematic_integrand (functional.c:3837-3923) */
static int nematic_integrand(const ctx *c, int nv, const int *vid, double *out) {
    const mo_functional *f = c->f; int dim = c->m->dim, psize = f->psize[0];
    double size = 0, gradnnraw[MO_MAXPSIZE * 3], gradnn[MO_MAXPSIZE * 3], divnn, curlnn[3] = {0, 0, 0};
    for (int i = 0; i < psize * 3; i++) { gradnn[i] = 0.0; gradnnraw[i] = 0.0; }
    if (!elementsize(c, c->g, nv, vid, &size)) return 0;
    const double *nn[MO_MAXNV];
    for (int i = 0; i < nv; i++) nn[i] = f->field[0] + (size_t) vid[i] * psize;
    if (c->g == 2) { if (!evaluategradient(c, f->field[0], psize, nv, vid, gradnnraw)) return 0; }
    else if (c->g == 3) { if (!evaluategradient3d(c, f->field[0], psize, nv, vid, gradnnraw)) return 0; }
    for (int j = 0; j < 3; j++) for (int i = 0; i < dim; i++) gradnn[3 * j + i] = gradnnraw[dim * j + i];
    divnn = 0.0; for (int i = 0; i < 3; i++) divnn += gradnn[i + 3 * i]; /* matrix_trace */
    curlnn[0] = gradnn[7] - gradnn[5];
    curlnn[1] = gradnn[2] - gradnn[6];
    curlnn[2] = gradnn[3] - gradnn[1];
    double ctwst[6] = {curlnn[0] * curlnn[0], curlnn[1] * curlnn[1], curlnn[2] * curlnn[2],
                       2 * curlnn[0] * curlnn[1], 2 * curlnn[1] * curlnn[2], 2 * curlnn[2] * curlnn[0]};
    double cbnd[6] = {ctwst[1] + ctwst[2], ctwst[0] + ctwst[2], ctwst[0] + ctwst[1], -ctwst[3], -ctwst[4], -ctwst[5]};
    double nnt[3][MO_MAXNV];
    for (int i = 0; i < nv; i++) for (int j = 0; j < 3; j++) nnt[j][i] = nn[i][j];
    double integrals[6] = {bcintfg(nv, nnt[0], nnt[0]), bcintfg(nv, nnt[1], nnt[1]), bcintfg(nv, nnt[2], nnt[2]),
                           bcintfg(nv, nnt[0], nnt[1]), bcintfg(nv, nnt[1], nnt[2]), bcintfg(nv, nnt[2], nnt[0])};
    double splay = 0.0, twist = 0.0, bend = 0.0, chol = 0.0;
    splay = 0.5 * f->ksplay * size * divnn * divnn;
    for (int i = 0; i < 6; i++) { twist += ctwst[i] * integrals[i]; bend += cbnd[i] * integrals[i]; }
    twist *= 0.5 * f->ktwist * size;
    bend *= 0.5 * f->kbend * size;
    if (f->haspitch) {
        for (int i = 0; i < 3; i++) chol += -2 * curlnn[i] * bcintf(nv, nnt[i]) * f->pitch;
        chol += (f->pitch * f->pitch);
        chol *= 0.5 * f->ktwist * size;
    }
    *out = splay + twist + bend + chol;
    return 1;
}

/* nematicelectric_integrand (functional.c:4053-4092); electric potential = scalar field[1] */
static int nematicelectric_integrand(const ctx *c, int nv, const int *vid, double *out) {
    const mo_functional *f = c->f; int dim = c->m->dim, psize = f->psize[0];
    double size = 0, ee[3] = {0, 0, 0};
    if (!elementsize(c, c->g, nv, vid, &size)) return 0;
    const double *nn[MO_MAXNV];
    for (int i = 0; i < nv; i++) nn[i] = f->field[0] + (size_t) vid[i] * psize;
    if (c->g == 2) { if (!evaluategradient(c, f->field[1], f->psize[1], nv, vid, ee)) return 0; }
    else if (c->g == 3) { if (!evaluategradient3d(c, f->field[1], f->psize[1], nv, vid, ee)) return 0; }
    /* in 2D the reference reads ee[2] and nnt[2] past the end of its dim-sized arrays (functional.c:4067, 4077-4087);
     * the golden outputs record the value those terms had there: zero */
    double nnt[3][MO_MAXNV] = {{0}};
    for (int i = 0; i < nv; i++) for (int j = 0; j < dim; j++) nnt[j][i] = nn[i][j];
    double total = ee[0] * ee[0] * bcintfg(nv, nnt[0], nnt[0]) +
                   ee[1] * ee[1] * bcintfg(nv, nnt[1], nnt[1]) +
                   ee[2] * ee[2] * bcintfg(nv, nnt[2], nnt[2]) +
                   2 * ee[0] * ee[1] * bcintfg(nv, nnt[0], nnt[1]) +
                   2 * ee[1] * ee[2] * bcintfg(nv, nnt[1], nnt[2]) +
                   2 * ee[2] * ee[0] * bcintfg(nv, nnt[2], nnt[0]);
    *out = size * total;
    return 1;
}

/* equielement_integrand (functional.c:2847-2893): the element is a vertex; the sizes of the elements of grade eg
 * incident to it (conn(eg,0), ascending element id) are compared with their mean */
static int equielement_integrand(const ctx *c, int id, double *out) {
    int nconn = c->etptr[id + 1] - c->etptr[id];
    const int *conn = c->etidx + c->etptr[id];
    if (nconn == 1) { *out = 0; return 1; }
    double *size = malloc(sizeof(double) * (size_t) (nconn > 0 ? nconn : 1)), mean = 0.0, total = 0.0;
    for (int i = 0; i < nconn; i++) {
        elementsize(c, c->eg, c->eg + 1, c->m->conn[c->eg] + (size_t) conn[i] * (c->eg + 1), &size[i]);
        mean += size[i];
    }
    mean /= ((double) nconn);
    if (fabs(mean) < MO_EPS) { free(size); return 0; }
    if (!c->f->weight || fabs(c->wmean) < MO_EPS) {
        for (int i = 0; i < nconn; i++) total += (1.0 - size[i] / mean) * (1.0 - size[i] / mean);
    } else {
        double wmean = 0.0;
        for (int i = 0; i < nconn; i++) wmean += (conn[i] < c->f->nweight ? c->f->weight[conn[i]] : 1.0);   /* matrix_getelement out of bounds leaves 1.0 */
        wmean /= ((double) nconn);
        if (fabs(wmean) < MO_EPS) wmean = 1.0;
        for (int i = 0; i < nconn; i++) {
            double w = (conn[i] < c->f->nweight ? c->f->weight[conn[i]] : 1.0);
            double term = (1.0 - w * size[i] / mean / wmean);
            total += term * term;
        }
    }
    free(size);
    *out = total;
    return 1;
}

static int integrand(const ctx *c, int id, int nv, const int *vid, double *out) {
    switch (c->f->functional) {
        case MO_EQUIELEMENT: return equielement_integrand(c, id, out);
        case MO_LENGTH: return length_integrand(c, nv, vid, out);
        case MO_AREAENCLOSED: return areaenclosed_integrand(c, nv, vid, out);
        case MO_AREA: return area_integrand(c, nv, vid, out);
        case MO_VOLUMEENCLOSED: return volumeenclosed_integrand(c, nv, vid, out);
        case MO_VOLUME: return volume_integrand(c, nv, vid, out);
        case MO_LINECURVATURESQ: return linecurvsq_integrand(c, id, out);
        case MO_GRADSQ: return gradsq_integrand(c, nv, vid, out);
        case MO_NORMSQ: return normsq_integrand(c, id, out);
        case MO_NEMATIC: return nematic_integrand(c, nv, vid, out);
        case MO_NEMATICELECTRIC: return nematicelectric_integrand(c, nv, vid, out);
    }
    return 0;
}

/* ------------------------------------------------------------------ *
 * Analytic per-element gradients
 * ------------------------------------------------------------------ */

/* length_gradient_scale (functional.c:1960-1972), scale = 1 */
static int length_gradient(const ctx *c, int nv, const int *vid, double *frc) {
    int dim = c->m->dim; double s0[3], norm;
    (void) nv;
    vecsub(dim, X(c, vid[1]), X(c, vid[0]), s0);
    norm = vecnorm(dim, s0);
    if (norm < MO_EPS) return 0;
    addtocolumn(frc, dim, vid[0], -1.0 / norm * 1.0, s0);
    addtocolumn(frc, dim, vid[1], 1. / norm * 1.0, s0);
    return 1;
}

/* area_gradient_scale (functional.c:2079-2103), scale = 1 */
static int area_gradient(const ctx *c, int nv, const int *vid, double *frc) {
    int dim = c->m->dim; double s0[3] = {0, 0, 0}, s1[3] = {0, 0, 0}, s01[3], s010[3], s011[3], norm;
    (void) nv;
    vecsub(dim, X(c, vid[1]), X(c, vid[0]), s0);
    vecsub(dim, X(c, vid[2]), X(c, vid[1]), s1);
    veccross(s0, s1, s01);
    norm = vecnorm(3, s01);
    if (norm < MO_EPS) return 0;
    veccross(s01, s0, s010);
    veccross(s01, s1, s011);
    addtocolumn(frc, dim, vid[0], 0.5 / norm * 1.0, s011);
    addtocolumn(frc, dim, vid[2], 0.5 / norm * 1.0, s010);
    vecadd(dim, s010, s011, s0);
    addtocolumn(frc, dim, vid[1], -0.5 / norm * 1.0, s0);
    return 1;
}

/* volumeenclosed_gradient (functional.c:2142-2164) */
static int volumeenclosed_gradient(const ctx *c, int nv, const int *vid, double *frc, int *err) {
    int dim = c->m->dim; double cx[3], dot;
    const double *x0 = X(c, vid[0]), *x1 = X(c, vid[1]), *x2 = X(c, vid[2]);
    (void) nv;
    veccross(x0, x1, cx);
    dot = vecdot(dim, cx, x2);
    if (fabs(dot) <= DBL_MIN) { *err = MO_ERR_VOLENCLZERO; return 0; }
    dot /= fabs(dot);
    addtocolumn(frc, dim, vid[2], dot / 6.0, cx);
    veccross(x1, x2, cx);
    addtocolumn(frc, dim, vid[0], dot / 6.0, cx);
    veccross(x2, x0, cx);
    addtocolumn(frc, dim, vid[1], dot / 6.0, cx);
    return 1;
}

/* volume_gradient_scale (functional.c:2201-2227), scale = 1 */
static int volume_gradient(const ctx *c, int nv, const int *vid, double *frc) {
    int dim = c->m->dim; double s10[3], s20[3], s30[3], s31[3], s21[3], cx[3], uu;
    const double *x0 = X(c, vid[0]), *x1 = X(c, vid[1]), *x2 = X(c, vid[2]), *x3 = X(c, vid[3]);
    (void) nv;
    vecsub(dim, x1, x0, s10); vecsub(dim, x2, x0, s20); vecsub(dim, x3, x0, s30);
    vecsub(dim, x3, x1, s31); vecsub(dim, x2, x1, s21);
    veccross(s20, s30, cx);
    uu = vecdot(dim, s10, cx);
    uu = (uu > 0 ? 1.0 : -1.0);
    addtocolumn(frc, dim, vid[1], uu / 6.0 * 1.0, cx);
    veccross(s31, s21, cx);
    addtocolumn(frc, dim, vid[0], uu / 6.0 * 1.0, cx);
    veccross(s30, s10, cx);
    addtocolumn(frc, dim, vid[2], uu / 6.0 * 1.0, cx);
    veccross(s10, s20, cx);
    addtocolumn(frc, dim, vid[3], uu / 6.0 * 1.0, cx);
    return 1;
}

/* ------------------------------------------------------------------ *
 * Dispatch plumbing
 * ------------------------------------------------------------------ */

int mo_default_grade(const mo_mesh *m, const mo_functional *f) {
    int maxg = 0;
    for (int g = m->dim; g > 0; g--) if (m->conn[g] && m->nel[g] > 0) { maxg = g; break; } /* mesh_maxgrade, mesh.c:334-340 */
    switch (f->functional) {
        case MO_LENGTH: case MO_AREAENCLOSED: return 1;
        case MO_AREA: case MO_VOLUMEENCLOSED: return 2;
        case MO_VOLUME: return 3;
        case MO_LINECURVATURESQ: case MO_NORMSQ: case MO_EQUIELEMENT: return 0;
        case MO_GRADSQ: case MO_NEMATIC: case MO_NEMATICELECTRIC:
            return (f->grade > 0 ? f->grade : maxg); /* gradsq_prepareref, functional.c:3622-3642 */
    }
    return -1;
}

int mo_count(const mo_mesh *m, const mo_functional *f, int *n) {
    int g = mo_default_grade(m, f);
    if (g < 0) return MO_ERR_ARGS;
    if (g == 0) { *n = m->nv; return MO_OK; }
    if (!m->conn[g]) return MO_ERR_ELNOTFOUND;
    *n = m->nel[g];
    return MO_OK;
}

static int ctx_init(ctx *c, const mo_mesh *m, const mo_functional *f) {
    memset(c, 0, sizeof(*c));
    c->m = m; c->f = f;
    c->g = mo_default_grade(m, f);
    if (c->g < 0) return MO_ERR_ARGS;
    if (c->g > 0 && !m->conn[c->g]) return MO_ERR_ELNOTFOUND;
    if ((f->functional == MO_GRADSQ || f->functional == MO_NEMATIC || f->functional == MO_NEMATICELECTRIC || f->functional == MO_NORMSQ)
        && (!f->field[0] || f->psize[0] < 1 || f->psize[0] > MO_MAXPSIZE)) return MO_ERR_ARGS;
    if (f->functional == MO_NEMATICELECTRIC && (!f->field[1] || f->psize[1] != 1)) return MO_ERR_ARGS;
    c->vert = malloc(sizeof(double) * (size_t) m->nv * m->dim);
    if (!c->vert) return MO_ERR_ARGS;
    memcpy(c->vert, m->vert, sizeof(double) * (size_t) m->nv * m->dim);
    if (f->functional == MO_LINECURVATURESQ) {
        if (!m->conn[1]) { free(c->vert); return MO_ERR_ELNOTFOUND; }
        c->ltptr = malloc(sizeof(int) * (size_t) (m->nv + 1));
        c->ltidx = malloc(sizeof(int) * (size_t) m->nel[1] * 2 + 1);
        mo_transpose(m->nv, m->nel[1], 2, m->conn[1], c->ltptr, c->ltidx);
    }
    if (f->functional == MO_EQUIELEMENT) {
        /* equielement_prepareref (functional.c:2775-2806): grade outside 0..maxgrade means the highest grade present */
        int maxg = 0;
        for (int g = m->dim; g > 0; g--) if (m->conn[g] && m->nel[g] > 0) { maxg = g; break; }
        c->eg = (f->grade < 0 || f->grade > maxg) ? maxg : f->grade;
        if (c->eg < 1 || !m->conn[c->eg]) { free(c->vert); return MO_ERR_ARGS; }
        c->etptr = malloc(sizeof(int) * (size_t) (m->nv + 1));
        c->etidx = malloc(sizeof(int) * ((size_t) m->nel[c->eg] * (c->eg + 1) + 1));
        mo_transpose(m->nv, m->nel[c->eg], c->eg + 1, m->conn[c->eg], c->etptr, c->etidx);
        if (f->weight && f->nweight > 0) {       /* matrix_sum / ncols */
            double sum = 0.0, cc = 0.0;
            for (int i = 0; i < f->nweight; i++) { double y = f->weight[i] - cc, t = sum + y; cc = (t - sum) - y; sum = t; }
            c->wmean = sum / f->nweight;
        }
    }
    return MO_OK;
}

static void ctx_free(ctx *c) { free(c->vert); free(c->ltptr); free(c->ltidx); free(c->etptr); free(c->etidx); }

static void element(const ctx *c, int id, int *nv, const int **vid, int *self) {
    if (c->g == 0) { *self = id; *nv = 1; *vid = self; }
    else { *nv = c->g + 1; *vid = c->m->conn[c->g] + (size_t) id * (c->g + 1); }
}

/* functional_sumintegrandX (functional.c:202-267): ascending id (or selection order), Kahan */
int mo_total(const mo_mesh *m, const mo_functional *f, const int *sel, int nsel, double *out) {
    ctx c; int rc = ctx_init(&c, m, f); if (rc) return rc;
    int n; mo_count(m, f, &n);
    double sum = 0.0, cc = 0.0, y, t, result;
    int cnt = sel ? nsel : n;
    for (int k = 0; k < cnt; k++) {
        int i = sel ? sel[k] : k, nv, self; const int *vid;
        element(&c, i, &nv, &vid, &self);
        if (integrand(&c, i, nv, vid, &result)) { y = result - cc; t = sum + y; cc = (t - sum) - y; sum = t; }
        else { ctx_free(&c); return MO_ERR_DEGENERATE; }
    }
    *out = sum;
    ctx_free(&c);
    return MO_OK;
}

/* functional_mapintegrandX (functional.c:274-349): 1 x n, unselected entries stay 0 */
int mo_integrand(const mo_mesh *m, const mo_functional *f, const int *sel, int nsel, double *out) {
    ctx c; int rc = ctx_init(&c, m, f); if (rc) return rc;
    int n; mo_count(m, f, &n);
    memset(out, 0, sizeof(double) * (size_t) n);
    int cnt = sel ? nsel : n;
    for (int k = 0; k < cnt; k++) {
        int i = sel ? sel[k] : k, nv, self; const int *vid; double result;
        element(&c, i, &nv, &vid, &self);
        if (integrand(&c, i, nv, vid, &result)) out[i] = result;
        else { ctx_free(&c); return MO_ERR_DEGENERATE; }
    }
    ctx_free(&c);
    return MO_OK;
}

/* functional_numericalgradient / functional_numericalremotegradient (functional.c:487-533) */
static int numericalgrad_vertex(ctx *c, int eid, int vtx, int nv, const int *vid, double *frc) {
    int dim = c->m->dim; double f0, fp, fm, x0, eps;
    for (int k = 0; k < dim; k++) {
        f0 = frc[(size_t) vtx * dim + k];
        x0 = c->vert[(size_t) vtx * dim + k];
        eps = mo_fdstepsize(x0, 1);
        c->vert[(size_t) vtx * dim + k] = x0 + eps;
        if (!integrand(c, eid, nv, vid, &fp)) return 0;
        c->vert[(size_t) vtx * dim + k] = x0 - eps;
        if (!integrand(c, eid, nv, vid, &fm)) return 0;
        c->vert[(size_t) vtx * dim + k] = x0;
        frc[(size_t) vtx * dim + k] = f0 + (fp - fm) / (2 * eps);
    }
    return 1;
}

/* linecurvsq_dependencies (functional.c:2964-2989): self first, then the other
 * vertex of each incident edge in ascending edge id */
static int dependencies(const ctx *c, int id, int *deps) {
    int n = 0;
    if (c->f->functional == MO_EQUIELEMENT) {
        /* equielement_dependencies (functional.c:2817-2844): the other vertices of the incident elements, in order of
         * discovery, each once; not the vertex itself */
        for (int q = c->etptr[id]; q < c->etptr[id + 1]; q++) {
            const int *entries = c->m->conn[c->eg] + (size_t) c->etidx[q] * (c->eg + 1);
            for (int j = 0; j <= c->eg; j++) {
                if (entries[j] == id) continue;
                int seen = 0;
                for (int k = 0; k < n; k++) if (deps[k] == entries[j]) seen = 1;
                if (!seen) deps[n++] = entries[j];
            }
        }
        return n;
    }
    deps[n++] = id;
    for (int q = c->ltptr[id]; q < c->ltptr[id + 1]; q++) {
        const int *entries = c->m->conn[1] + 2 * (size_t) c->ltidx[q];
        for (int j = 0; j < 2; j++) { if (entries[j] == id) continue; deps[n++] = entries[j]; }
    }
    return n;
}

/* functional_mapgradientX (functional.c:356-416) and functional_mapnumericalgradientX
 * (functional.c:557-651) */
int mo_gradient(const mo_mesh *m, const mo_functional *f, const int *sel, int nsel, double *out) {
    ctx c; int rc = ctx_init(&c, m, f); if (rc) return rc;
    int n, dim = m->dim; mo_count(m, f, &n);
    memset(out, 0, sizeof(double) * (size_t) m->nv * dim);
    int cnt = sel ? nsel : n, err = MO_ERR_DEGENERATE;
    int analytic = (f->functional == MO_LENGTH || f->functional == MO_AREA ||
                    f->functional == MO_VOLUMEENCLOSED || f->functional == MO_VOLUME);
    int *deps = NULL;
    if (f->functional == MO_LINECURVATURESQ) deps = malloc(sizeof(int) * (size_t) (2 * m->nel[1] + 2));
    if (f->functional == MO_EQUIELEMENT) deps = malloc(sizeof(int) * ((size_t) m->nel[c.eg] * (c.eg + 1) + 2));
    for (int k = 0; k < cnt; k++) {
        int i = sel ? sel[k] : k, nv, self, ok = 1; const int *vid;
        element(&c, i, &nv, &vid, &self);
        if (analytic) {
            switch (f->functional) {
                case MO_LENGTH: ok = length_gradient(&c, nv, vid, out); break;
                case MO_AREA: ok = area_gradient(&c, nv, vid, out); break;
                case MO_VOLUMEENCLOSED: ok = volumeenclosed_gradient(&c, nv, vid, out, &err); break;
                case MO_VOLUME: ok = volume_gradient(&c, nv, vid, out); break;
            }
        } else {
            for (int j = 0; j < nv && ok; j++) ok = numericalgrad_vertex(&c, i, vid[j], nv, vid, out);
            if (ok && deps) {
                int nd = dependencies(&c, i, deps);
                for (int j = 0; j < nd && ok; j++) {
                    /* the element loop skips dependencies that are element vertices
                     * (functional.c:626); the selection loop does not (functional.c:602-606) */
                    if (!sel) { int found = 0; for (int q = 0; q < nv; q++) if (vid[q] == deps[j]) found = 1; if (found) continue; }
                    ok = numericalgrad_vertex(&c, i, deps[j], nv, vid, out);
                }
            }
        }
        if (!ok) { free(deps); ctx_free(&c); return err; }
    }
    free(deps);
    ctx_free(&c);
    return MO_OK;
}

/* functional_mapnumericalfieldgradient (functional.c:1428-1503) with one task;
 * functional_numericalfieldgradientmapfn (1377-1425, non-fespace branch) and
 * functional_numericalfieldgrad (1305-1328). */
int mo_fieldgradient(const mo_mesh *m, const mo_functional *f, const int *sel, int nsel, double *out) {
    ctx c; int rc = ctx_init(&c, m, f); if (rc) return rc;
    int w = f->wrt; if (w < 0 || w > 1 || !f->field[w]) { ctx_free(&c); return MO_ERR_ARGS; }
    int n, psize = f->psize[w]; mo_count(m, f, &n);
    double *fld = f->field[w];
    memset(out, 0, sizeof(double) * (size_t) m->nv * psize);
    int cnt = sel ? nsel : n;
    for (int k = 0; k < cnt; k++) {
        int i = sel ? sel[k] : k, nv, self; const int *vid;
        element(&c, i, &nv, &vid, &self);
        for (int q = 0; q < nv; q++) {
            for (int j = 0; j < psize; j++) {
                size_t idx = (size_t) vid[q] * psize + j;
                double fr, fl, f0 = fld[idx], eps = mo_fdstepsize(f0, 1);
                fld[idx] += eps;
                if (!integrand(&c, i, nv, vid, &fr)) { fld[idx] = f0; ctx_free(&c); return MO_ERR_DEGENERATE; }
                fld[idx] = f0 - eps;
                if (!integrand(&c, i, nv, vid, &fl)) { fld[idx] = f0; ctx_free(&c); return MO_ERR_DEGENERATE; }
                fld[idx] = f0;
                out[idx] += (fr - fl) / (2 * eps);
            }
        }
    }
    ctx_free(&c);
    return MO_OK;
}

/* ------------------------------------------------------------------ *
 * Connectivity
 * ------------------------------------------------------------------ */

/* sparse_transpose → cs_transpose (sparse.c:1227-1244): stable counting sort */
int mo_transpose(int nv, int nel, int nvper, const int *conn, int *tptr, int *tidx) {
    int *w = calloc((size_t) nv + 1, sizeof(int));
    if (!w) return MO_ERR_ARGS;
    size_t nnz = (size_t) nel * nvper;
    for (size_t k = 0; k < nnz; k++) w[conn[k]]++;
    int run = 0;
    for (int v = 0; v < nv; v++) { tptr[v] = run; int cnt = w[v]; w[v] = run; run += cnt; }
    tptr[nv] = run;
    for (int e = 0; e < nel; e++) for (int k = 0; k < nvper; k++) tidx[w[conn[(size_t) e * nvper + k]]++] = e;
    free(w);
    return MO_OK;
}

static int cmp_edge(const void *a, const void *b) {
    const long long *x = a, *y = b;
    if (x[0] != y[0]) return (x[0] > y[0]) - (x[0] < y[0]);
    return (x[1] > y[1]) - (x[1] < y[1]);
}

/* mesh_addgrade (mesh.c:419-489) for g = 1: candidate pairs in element order and
 * lexicographic counter order; first occurrence defines the new id. */
int mo_edges_from(int nel, int nvper, const int *conn, int *edges_out) {
    int npairs = nvper * (nvper - 1) / 2;
    size_t ncand = (size_t) nel * npairs;
    long long *key = malloc(sizeof(long long) * 2 * ncand);
    if (!key) return -1;
    size_t q = 0;
    for (int e = 0; e < nel; e++)
        for (int a = 0; a < nvper; a++)
            for (int b = a + 1; b < nvper; b++) {
                long long lo = conn[(size_t) e * nvper + a], hi = conn[(size_t) e * nvper + b];
                key[2 * q] = lo * 2147483648LL + hi; key[2 * q + 1] = (long long) q; q++;
            }
    qsort(key, ncand, 2 * sizeof(long long), cmp_edge);
    /* mark first occurrences */
    char *first = calloc(ncand, 1);
    for (size_t i = 0; i < ncand; i++) if (i == 0 || key[2 * i] != key[2 * (i - 1)]) first[key[2 * i + 1]] = 1;
    int n = 0; q = 0;
    for (int e = 0; e < nel; e++)
        for (int a = 0; a < nvper; a++)
            for (int b = a + 1; b < nvper; b++) {
                if (first[q]) { edges_out[2 * n] = conn[(size_t) e * nvper + a]; edges_out[2 * n + 1] = conn[(size_t) e * nvper + b]; n++; }
                q++;
            }
    free(first); free(key);
    return n;
}

/* functional_binbounds (functional.c:863-879) */
void mo_binbounds(int nel, int nbins, int *bounds) {
    int defsize = nel / nbins, rem = nel % nbins, b = 0;
    for (int i = 0; i <= nbins; i++) { bounds[i] = b; if (i < nbins) b += defsize + (i < rem ? 1 : 0); }
}

/* ------------------------------------------------------------------ *
 * PairwisePotential, Coulomb closures (modules/functionals.morpho:105-139)
 * Matrix arithmetic of the script: dx = x_i - x_j (matrix sub), r = dx.norm()
 * (dnrm2), out[i] += 0.5*f ; gradient: df = grad(r)/r ; column_i = df*dx + column_i.
 * ------------------------------------------------------------------ */
/* The closures func(r) / grad(r) a script hands to PairwisePotential, for the potential families the library knows
 * (include/morpho_cuda.h MCU_PW_*), written as the closures of tests/golden/pairwise_cases.py write them
 * (Morpho's ^ is pow(), veneer.c / vm.c OP_POW). */
static double pw_func(int kind, const double *p, double r) {
    switch (kind) {
        case 0: return p[0] / r;                                         /* fn (r) a/r */
        case 1: return p[0] / pow(r, p[1]);                              /* fn (r) a/r^p */
        case 2: return p[0] * exp(-r / p[1]) / r;                        /* fn (r) a*exp(-r/l)/r */
        case 3: return 4 * p[0] * (pow(p[1] / r, 12) - pow(p[1] / r, 6)); /* fn (r) 4 e ((s/r)^12-(s/r)^6) */
    }
    return NAN;
}
static double pw_grad(int kind, const double *p, double r) {
    switch (kind) {
        case 0: return -p[0] / pow(r, 2);
        case 1: return -p[0] * p[1] / pow(r, p[1] + 1);
        case 2: return -p[0] * exp(-r / p[1]) * (1 / pow(r, 2) + 1 / (p[1] * r));
        case 3: return 4 * p[0] * (-12 * pow(p[1] / r, 12) + 6 * pow(p[1] / r, 6)) / r;
    }
    return NAN;
}

/* PairwisePotential.integrand (functionals.morpho:105-121) */
int mo_pairwise_integrand(int n, int dim, const double *x, int kind, const double *params, double cutoff, double *out) {
    memset(out, 0, sizeof(double) * (size_t) n);
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            double dx[3];
            vecsub(dim, x + (size_t) i * dim, x + (size_t) j * dim, dx);
            double r = vecnorm(dim, dx);
            if (cutoff > 0 && r > cutoff) continue;
            double fv = pw_func(kind, params, r);
            out[i] += 0.5 * fv;
            out[j] += 0.5 * fv;
        }
    return MO_OK;
}

/* PairwisePotential.gradient (functionals.morpho:123-139) */
int mo_pairwise_gradient(int n, int dim, const double *x, int kind, const double *params, double cutoff, double *out) {
    memset(out, 0, sizeof(double) * (size_t) n * dim);
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            double dx[3];
            vecsub(dim, x + (size_t) i * dim, x + (size_t) j * dim, dx);
            double r = vecnorm(dim, dx);
            if (cutoff > 0 && r > cutoff) continue;
            double df = pw_grad(kind, params, r) / r;
            for (int k = 0; k < dim; k++) {
                out[(size_t) i * dim + k] = df * dx[k] + out[(size_t) i * dim + k];
                out[(size_t) j * dim + k] = -df * dx[k] + out[(size_t) j * dim + k];
            }
        }
    return MO_OK;
}

/* Functional.total (functionals.morpho:50-53): integrand(mesh).sum() — Matrix.sum is a Kahan sum (matrix.c matrix_sum) */
double mo_pairwise_total(int n, int dim, const double *x, int kind, const double *params, double cutoff) {
    double *e = malloc(sizeof(double) * (size_t) (n > 0 ? n : 1));
    mo_pairwise_integrand(n, dim, x, kind, params, cutoff, e);
    double sum = 0.0, c = 0.0;
    for (int i = 0; i < n; i++) { double y = e[i] - c, t = sum + y; c = (t - sum) - y; sum = t; }
    free(e);
    return sum;
}

int mo_pairwise_coulomb_integrand(int n, int dim, const double *x, double cutoff, double *out) {
    const double one = 1.0;
    return mo_pairwise_integrand(n, dim, x, 0, &one, cutoff, out);
}
int mo_pairwise_coulomb_gradient(int n, int dim, const double *x, double cutoff, double *out) {
    const double one = 1.0;
    return mo_pairwise_gradient(n, dim, x, 0, &one, cutoff, out);
}

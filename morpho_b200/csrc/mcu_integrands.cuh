// mcu_integrands.cuh — per-element integrands of all supported functionals on data that
// has already been gathered into registers/local memory (ElemData).  Finite-difference
// gradients re-evaluate these with one perturbed entry, exactly like the reference's
// functional_numericalgrad / functional_numericalfieldgrad (functional.c:1142-1162,
// 1305-1328) re-evaluates its integrand callbacks.
//
// When MCU_STRICT is defined (translation unit compiled with -fmad=false) vector norms
// use the extended-precision emulation norm_dd() so that the arithmetic follows the
// reference's (un-fused C arithmetic + an extended-precision dnrm2) as closely as a
// GPU can; this matters because central differences amplify integrand round-off by
// 1/(2h) ~ 8e4.
#pragma once

#include "mcu_device.cuh"

struct ElemData {
    double x[MCU_MAXNV][3];             // vertex coordinates, zero padded when dim == 2
    double q[MCU_MAXNV][MCU_MAXPSIZE];  // primary field values at the element's vertices
    double phi[MCU_MAXNV];              // secondary (scalar) field values
};

__device__ __forceinline__ double vnorm(int n, const double *a) {
#ifdef MCU_STRICT
    return norm_dd(n, a);
#else
    return sqrt(dotn(n, a, a));
#endif
}

// functional_elementsize (functional.c:1901-1908)
__device__ __forceinline__ double elem_size(int g, const double (*x)[3]) {
#ifdef MCU_STRICT
    if (g == 1) {
        double s0[3] = {x[1][0] - x[0][0], x[1][1] - x[0][1], x[1][2] - x[0][2]};
        return norm_dd(3, s0);
    } else if (g == 2) {
        double s0[3] = {x[1][0] - x[0][0], x[1][1] - x[0][1], x[1][2] - x[0][2]};
        double s1[3] = {x[2][0] - x[1][0], x[2][1] - x[1][1], x[2][2] - x[1][2]};
        double cx[3];
        cross3(s0, s1, cx);
        return 0.5 * norm_dd(3, cx);
    }
    return volume_of(x);
#else
    if (g == 1) return length_of(x);
    if (g == 2) return area_of(x);
    return volume_of(x);
#endif
}

// gradsq_computeperpendicular (functional.c:3485-3503): out = (s1 - (s1.s2/s2.s2) s2) / |.|^2
__device__ __forceinline__ void perpendicular(int n, const double *s1, const double *s2, double *out) {
    double s1s2 = dotn(n, s1, s2), s2s2 = dotn(n, s2, s2);
    if (fabs(s2s2) < MCU_EPS) return;  // reference returns false and the caller ignores it
    double r = s1s2 / s2s2;
    for (int k = 0; k < n; k++) out[k] = s1[k] - r * s2[k];
    double sout = vnorm(n, out);
    if (fabs(sout) < MCU_EPS) return;
    double inv = 1 / (sout * sout);
    for (int k = 0; k < n; k++) out[k] = inv * out[k];
}

// gradsq_evaluategradient (functional.c:3511-3542): P1 gradient on a triangle.
// vals[j*stride + i] = value i at vertex j ; out[i*dim + k]
__device__ __forceinline__ void p1grad_tri(int dim, const double (*x)[3], int nvals, const double *vals, int stride, double *out) {
    double s[3][3], t[3][3];
    for (int k = 0; k < 3; k++) {
        s[0][k] = x[1][k] - x[0][k];
        s[1][k] = x[2][k] - x[1][k];
        s[2][k] = x[0][k] - x[2][k];
        t[0][k] = t[1][k] = t[2][k] = 0.0;
    }
    perpendicular(dim, s[2], s[1], t[0]);
    perpendicular(dim, s[0], s[2], t[1]);
    perpendicular(dim, s[1], s[0], t[2]);
#pragma unroll
    for (int i = 0; i < nvals * dim; i++) out[i] = 0;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < nvals; i++)
#pragma unroll
            for (int k = 0; k < dim; k++) out[i * dim + k] = out[i * dim + k] + vals[j * stride + i] * t[j][k];
}

// gradsq_evaluategradient3d (functional.c:3584-3620): rows (x_i - x_0)^T, rhs f_i - f_0, solved by
// LU with partial pivoting in the order LAPACK's dgetf2 + dgetrs use (matrix.c:129-139).
// Split in two so that callers that change only the VALUES (finite differences with respect to a field) factor once.
struct TetLU { double a[9]; int piv[3]; };

__device__ __forceinline__ void p1grad_tet_factor(const double (*x)[3], TetLU &f) {
    // Same operations in the same order as dgetf2 on the 3x3 system; the row exchanges are written as
    // conditional swaps with constant indices (a run-time pivot index would put a[] in local memory).
    double *a = f.a;  // column major Mt: Mt[r + 3 c] = (x_{r+1} - x_0)[c]
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) a[r + 3 * c] = x[r + 1][c] - x[0][c];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        int p = j;
        double big = fabs(a[j + 3 * j]);
#pragma unroll
        for (int i = j + 1; i < 3; i++)
            if (fabs(a[i + 3 * j]) > big) { big = fabs(a[i + 3 * j]); p = i; }
        f.piv[j] = p;
#pragma unroll
        for (int i = j + 1; i < 3; i++)
            if (p == i) {
#pragma unroll
                for (int c = 0; c < 3; c++) { double tt = a[j + 3 * c]; a[j + 3 * c] = a[i + 3 * c]; a[i + 3 * c] = tt; }
            }
        double r = 1.0 / a[j + 3 * j];
#pragma unroll
        for (int i = j + 1; i < 3; i++) a[i + 3 * j] *= r;
#pragma unroll
        for (int c = j + 1; c < 3; c++)
#pragma unroll
            for (int i = j + 1; i < 3; i++) a[i + 3 * c] -= a[i + 3 * j] * a[j + 3 * c];
    }
}

// dgetrs for one right-hand side: values v0..v3 at the four vertices, out = gradient (3 numbers)
__device__ __forceinline__ void p1grad_tet_solve(const TetLU &f, double v0, double v1, double v2, double v3, double *out) {
    const double *a = f.a;
    double b[3] = {v1 - v0, v2 - v0, v3 - v0};
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = j + 1; i < 3; i++)
            if (f.piv[j] == i) { double tt = b[j]; b[j] = b[i]; b[i] = tt; }
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = j + 1; i < 3; i++) b[i] -= b[j] * a[i + 3 * j];
#pragma unroll
    for (int j = 2; j >= 0; j--) {
        b[j] /= a[j + 3 * j];
#pragma unroll
        for (int i = 0; i < j; i++) b[i] -= b[j] * a[i + 3 * j];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) out[k] = b[k];
}

__device__ __forceinline__ void p1grad_tet(const double (*x)[3], int nvals, const double *vals, int stride, double *out) {
    TetLU f;
    p1grad_tet_factor(x, f);
#pragma unroll
    for (int v = 0; v < nvals; v++) p1grad_tet_solve(f, vals[v], vals[stride + v], vals[2 * stride + v], vals[3 * stride + v], out + v * 3);
}

// the geometric part of p1grad_tri: the three "perpendiculars" t[j]; grad(value) = sum_j value_j t[j]
__device__ __forceinline__ void p1grad_tri_factor(int dim, const double (*x)[3], double (*t)[3]) {
    double s[3][3];
    for (int k = 0; k < 3; k++) {
        s[0][k] = x[1][k] - x[0][k];
        s[1][k] = x[2][k] - x[1][k];
        s[2][k] = x[0][k] - x[2][k];
        t[0][k] = t[1][k] = t[2][k] = 0.0;
    }
    perpendicular(dim, s[2], s[1], t[0]);
    perpendicular(dim, s[0], s[2], t[1]);
    perpendicular(dim, s[1], s[0], t[2]);
}
__device__ __forceinline__ void p1grad_tri_apply(int dim, const double (*t)[3], double v0, double v1, double v2, double *out) {
#pragma unroll
    for (int k = 0; k < 3; k++) if (k < dim) out[k] = ((0.0 + v0 * t[0][k]) + v1 * t[1][k]) + v2 * t[2][k];
}

// What the field functionals need from the vertex positions alone: element size and the factors of the P1 gradient.
// The finite-difference field gradients perturb field values only, so they compute it once per element.
struct ElemGeom { double size; TetLU lu; double t[3][3]; };
__device__ __forceinline__ void elem_geom(int g, int dim, const double (*x)[3], ElemGeom &gm) {
    gm.size = elem_size(g, x);
    if (g == 2) p1grad_tri_factor(dim, x, gm.t);
    else if (g == 3) p1grad_tet_factor(x, gm.lu);
}
// p1grad_tri / p1grad_tet with the factors taken from gm: the same operations on the values, in the same order
__device__ __forceinline__ void p1grad_cached(int g, int dim, const ElemGeom &gm, int nvals, const double *vals, int stride, double *out) {
#pragma unroll
    for (int v = 0; v < nvals; v++) {
        if (g == 2) p1grad_tri_apply(dim, gm.t, vals[v], vals[stride + v], vals[2 * stride + v], out + v * dim);
        else p1grad_tet_solve(gm.lu, vals[v], vals[stride + v], vals[2 * stride + v], vals[3 * stride + v], out + v * 3);
    }
}

// nematic_bcintf / nematic_bcintfg (functional.c:3818-3834): exact barycentric integrals
__device__ __forceinline__ double bcintf(int n, const double *f) {
    double sum = 0;
#pragma unroll
    for (int i = 0; i < n; i++) sum += f[i];
    return sum / n;
}
__device__ __forceinline__ double bcintfg(int n, const double *f, const double *g) {
    double sum = 0;
#pragma unroll
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < n; j++) sum += f[i] * g[j];
        sum += f[i] * g[i];
    }
    return sum / (n * (n + 1));
}

// gradsq_integrand (functional.c:3657-3676): size * |grad q|_F^2
__device__ __forceinline__ double gradsq_integrand(const McuParams &p, const ElemData &d, int nv, int dim, int ps, const ElemGeom *gm = nullptr) {
    const int g = nv - 1;
    double grad[MCU_MAXPSIZE * 3];
    double size = gm ? gm->size : elem_size(g, d.x);
    if (gm) p1grad_cached(g, dim, *gm, ps, &d.q[0][0], MCU_MAXPSIZE, grad);
    else if (g == 2) p1grad_tri(dim, d.x, ps, &d.q[0][0], MCU_MAXPSIZE, grad);
    else p1grad_tet(d.x, ps, &d.q[0][0], MCU_MAXPSIZE, grad);
    double gradnrm = vnorm(ps * dim, grad);
    return gradnrm * gradnrm * size;
}

// nematic_integrand (functional.c:3837-3923): Frank energy with a P1 director
__device__ __forceinline__ double nematic_integrand(const McuParams &p, const ElemData &d, int nv, int dim, int ps, const ElemGeom *gm = nullptr) {
    const int g = nv - 1;
    double gradnnraw[MCU_MAXPSIZE * 3], gradnn[9], curlnn[3];
#pragma unroll
    for (int i = 0; i < 9; i++) gradnn[i] = 0.0;
#pragma unroll
    for (int i = 0; i < ps * 3; i++) gradnnraw[i] = 0.0;
    double size = gm ? gm->size : elem_size(g, d.x);
    if (gm) p1grad_cached(g, dim, *gm, ps, &d.q[0][0], MCU_MAXPSIZE, gradnnraw);
    else if (g == 2) p1grad_tri(dim, d.x, ps, &d.q[0][0], MCU_MAXPSIZE, gradnnraw);
    else if (g == 3) p1grad_tet(d.x, ps, &d.q[0][0], MCU_MAXPSIZE, gradnnraw);
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < dim; i++) gradnn[3 * j + i] = gradnnraw[dim * j + i];
    double divnn = 0.0;
    for (int i = 0; i < 3; i++) divnn += gradnn[4 * i];
    curlnn[0] = gradnn[7] - gradnn[5];
    curlnn[1] = gradnn[2] - gradnn[6];
    curlnn[2] = gradnn[3] - gradnn[1];
    double ctwst[6] = {curlnn[0] * curlnn[0], curlnn[1] * curlnn[1], curlnn[2] * curlnn[2],
                       2 * curlnn[0] * curlnn[1], 2 * curlnn[1] * curlnn[2], 2 * curlnn[2] * curlnn[0]};
    double cbnd[6] = {ctwst[1] + ctwst[2], ctwst[0] + ctwst[2], ctwst[0] + ctwst[1], -ctwst[3], -ctwst[4], -ctwst[5]};
    double nnt[3][MCU_MAXNV];
#pragma unroll
    for (int i = 0; i < nv; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) nnt[j][i] = d.q[i][j];
    double integrals[6] = {bcintfg(nv, nnt[0], nnt[0]), bcintfg(nv, nnt[1], nnt[1]), bcintfg(nv, nnt[2], nnt[2]),
                           bcintfg(nv, nnt[0], nnt[1]), bcintfg(nv, nnt[1], nnt[2]), bcintfg(nv, nnt[2], nnt[0])};
    double splay, twist = 0.0, bend = 0.0, chol = 0.0;
    splay = 0.5 * p.ksplay * size * divnn * divnn;
#pragma unroll
    for (int i = 0; i < 6; i++) { twist += ctwst[i] * integrals[i]; bend += cbnd[i] * integrals[i]; }
    twist *= 0.5 * p.ktwist * size;
    bend *= 0.5 * p.kbend * size;
    if (p.haspitch) {
#pragma unroll
        for (int i = 0; i < 3; i++) chol += -2 * curlnn[i] * bcintf(nv, nnt[i]) * p.pitch;
        chol += (p.pitch * p.pitch);
        chol *= 0.5 * p.ktwist * size;
    }
    return splay + twist + bend + chol;
}

// nematicelectric_integrand (functional.c:4053-4092): size * E_i E_j <n_i n_j>, E = grad(phi)
__device__ __forceinline__ double nematicelectric_integrand(const McuParams &p, const ElemData &d, int nv, int dim, int ps, const ElemGeom *gm = nullptr) {
    const int g = nv - 1;
    double ee[3] = {0.0, 0.0, 0.0};
    double size = gm ? gm->size : elem_size(g, d.x);
    if (gm) p1grad_cached(g, dim, *gm, 1, d.phi, 1, ee);
    else if (g == 2) p1grad_tri(dim, d.x, 1, d.phi, 1, ee);
    else if (g == 3) p1grad_tet(d.x, 1, d.phi, 1, ee);
    double nnt[3][MCU_MAXNV];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < nv; i++) nnt[j][i] = (j < dim ? d.q[i][j] : 0.0);
    double total = ee[0] * ee[0] * bcintfg(nv, nnt[0], nnt[0]) +
                   ee[1] * ee[1] * bcintfg(nv, nnt[1], nnt[1]) +
                   ee[2] * ee[2] * bcintfg(nv, nnt[2], nnt[2]) +
                   2 * ee[0] * ee[1] * bcintfg(nv, nnt[0], nnt[1]) +
                   2 * ee[1] * ee[2] * bcintfg(nv, nnt[1], nnt[2]) +
                   2 * ee[2] * ee[0] * bcintfg(nv, nnt[2], nnt[0]);
    return size * total;
}

// One integrand evaluation on gathered data.  F is a compile-time functional id.
// (nv, dim, ps) = vertices per element, coordinates per vertex, doubles per field entry: literal constants in the
// shape-specialised kernels (everything below unrolls and the element data stays in registers), p's values otherwise.
template <int F>
__device__ __forceinline__ double eval_integrand(const McuParams &p, const ElemData &d, int nv, int dim, int ps, const ElemGeom *gm = nullptr) {
    if (F == MCU_F_LENGTH) return elem_size(1, d.x);
    if (F == MCU_F_AREAENCLOSED) {
#ifdef MCU_STRICT
        if (p.dim == 2) return 0.5 * fabs(d.x[0][0] * d.x[1][1] - d.x[0][1] * d.x[1][0]);
        double cx[3];
        cross3(d.x[0], d.x[1], cx);
        return 0.5 * norm_dd(3, cx);
#else
        return areaenclosed_of(d.x, p.dim);
#endif
    }
    if (F == MCU_F_AREA) return elem_size(2, d.x);
    if (F == MCU_F_VOLUMEENCLOSED) return volumeenclosed_of(d.x);
    if (F == MCU_F_VOLUME) return volume_of(d.x);
    if (F == MCU_F_GRADSQ) return gradsq_integrand(p, d, nv, dim, ps, gm);
    if (F == MCU_F_NEMATIC) return nematic_integrand(p, d, nv, dim, ps, gm);
    if (F == MCU_F_NEMATICELECTRIC) return nematicelectric_integrand(p, d, nv, dim, ps, gm);
    if (F == MCU_F_NORMSQ) return dotn(ps, d.q[0], d.q[0]);  // normsq_integrand, functional.c:4149-4160
    return 0.0;
}

template <int F>
struct FTraits {
    static constexpr bool needs_q = (F == MCU_F_GRADSQ || F == MCU_F_NEMATIC || F == MCU_F_NEMATICELECTRIC || F == MCU_F_NORMSQ);
    static constexpr bool needs_phi = (F == MCU_F_NEMATICELECTRIC);
    static constexpr bool analytic = (F == MCU_F_LENGTH || F == MCU_F_AREA || F == MCU_F_VOLUMEENCLOSED || F == MCU_F_VOLUME);
};

// Gather one element's data.  ids: nvper vertex ids.
template <int F>
__device__ __forceinline__ void gather_elem(const McuParams &p, const McuView &v, const int *ids, ElemData &d, int nv, int dim, int ps) {
    // constant trip count with a guard: d.x stays in registers (a run-time trip count puts it in local memory)
#pragma unroll
    for (int j = 0; j < MCU_MAXNV; j++) {
        if (j >= nv) continue;
        const double *xv = v.vert + (size_t) ids[j] * dim;
        d.x[j][0] = xv[0];
        d.x[j][1] = xv[1];
        d.x[j][2] = (dim == 3 ? xv[2] : 0.0);
        if (FTraits<F>::needs_q) {
            const double *qv = v.f0 + (size_t) ids[j] * ps;
#pragma unroll
            for (int i = 0; i < ps; i++) d.q[j][i] = qv[i];
        }
        if (FTraits<F>::needs_phi) d.phi[j] = v.f1[ids[j]];
    }
}

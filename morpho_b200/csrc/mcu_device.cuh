// mcu_device.cuh — device-side building blocks shared by the kernel translation units.
//
// Per-element arithmetic of the functionals, written from the mathematical
// definitions the reference implements (citations per function, paths relative to
// /root/reference/).  All FP64.  No atomics anywhere: gradients are assembled by
// gathers with a fixed summation order.
#pragma once

#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>

#include "../../include/morpho_cuda.h"

#define MCU_MAXNV 4
#define MCU_MAXPSIZE 9
#define MCU_EPS DBL_EPSILON  // MORPHO_EPS (src/build.h:50)

// error bits raised by kernels into a device flag word
#define MCU_FLAG_DEGENERATE 1u
#define MCU_FLAG_VOLENCLZERO 2u

// Plain-old-data view of a mesh/functional handed to kernels by value.
struct McuParams {
    int functional;
    int g;        // grade looped over
    int nvper;    // g + 1 (1 for grade 0)
    int dim;      // 2 or 3
    int nv;       // number of vertices
    int nel;      // number of elements of grade g in the mesh (integrand length)
    int psize0, psize1;
    double ksplay, ktwist, kbend, pitch;
    int haspitch, integrandonly;
    int wrt;
    int egrade;       // EquiElement: grade of the elements whose sizes are compared
    int nweight;      // EquiElement: number of weights (0: unweighted)
    double wmean;     // EquiElement: mean weight (equielement_prepareref, functional.c:2797-2803)
};

struct McuView {
    const double *vert;   // nv*dim
    const int *rix;       // element → vertex ids (nel*nvper); nullptr for grade 0
    const int *eids;      // selection: compact index → element id; nullptr = all
    int n;                // number of elements to visit (nel or selection count)
    const double *f0;     // field 0 (nv*psize0) or nullptr
    const double *f1;     // field 1 (nv*psize1) or nullptr
    // vertex → compact element incidence (CSR transpose), element ids ascending per vertex
    const int *tptr;
    const int *tidx;
    // LineCurvatureSq: line elements and their transpose; EquiElement: the elements of grade egrade and their transpose
    const int *lrix;
    const int *ltptr;
    const int *ltidx;
    const unsigned char *selmask;  // grade-0 selections: 1 where the vertex element is selected
    const double *weight;          // EquiElement: one weight per element, or nullptr
    unsigned *flags;               // device status word (MCU_FLAG_*)
};

// ---------------------------------------------------------------------------------
// small vector helpers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void cross3(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ double dotn(int n, const double *a, const double *b) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

// Error-free transformations (used to emulate the extended-precision dnrm2 the reference's
// BLAS performs, functional.c:1880-1882, and for compensated reductions).
__device__ __forceinline__ void two_sum(double a, double b, double &s, double &e) {
    s = __dadd_rn(a, b);
    double bb = __dsub_rn(s, a);
    e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
}
__device__ __forceinline__ void two_prod(double a, double b, double &p, double &e) {
    p = __dmul_rn(a, b);
    e = __fma_rn(a, b, -p);
}

// sqrt(sum x_i^2) with the sum of squares carried in double-double: agrees with an
// 80-bit accumulation + one rounding in all but astronomically rare cases.
__device__ __forceinline__ double norm_dd(int n, const double *a) {
    double hi = 0.0, lo = 0.0;
    for (int i = 0; i < n; i++) {
        double p, pe, s, se;
        two_prod(a[i], a[i], p, pe);
        two_sum(hi, p, s, se);
        lo = __dadd_rn(lo, __dadd_rn(se, pe));
        hi = s;
    }
    double s = __dadd_rn(hi, lo);
    if (s <= 0.0) return 0.0;
    double r = sqrt(s);
    // one correction step with the residual (hi - r*r) + lo
    double rr, rre;
    two_prod(r, r, rr, rre);
    double res = __dadd_rn(__dadd_rn(__dsub_rn(hi, rr), -rre), lo);
    return __dadd_rn(r, __ddiv_rn(res, __dmul_rn(2.0, r)));
}

// functional_fdstepsize (functional.c:46-58): h = cbrt(eps)*max(1,|x|), made representable.
__device__ __forceinline__ double fdstepsize(double x) {
    const double fddelta1 = 6.0554544523933395e-06;  // pow(DBL_EPSILON, 1.0/3.0), functional.c:6230
    double h = fddelta1;
    double absx = fabs(x);
    if (absx > 1) h = __dmul_rn(h, absx);
    double temp = __dadd_rn(x, h);
    return __dsub_rn(temp, x);
}

// ---------------------------------------------------------------------------------
// Element sizes and analytic gradients.  X is double[nv][3] (zero padded for dim 2).
// ---------------------------------------------------------------------------------

// length_integrand (functional.c:1948-1958)
__device__ __forceinline__ double length_of(const double (*x)[3]) {
    double s0[3] = {x[1][0] - x[0][0], x[1][1] - x[0][1], x[1][2] - x[0][2]};
    return sqrt(dot3(s0, s0));
}

// area_integrand (functional.c:2063-2076): A = |s0 x s1| / 2, s0 = x1-x0, s1 = x2-x1
__device__ __forceinline__ double area_of(const double (*x)[3]) {
    double s0[3] = {x[1][0] - x[0][0], x[1][1] - x[0][1], x[1][2] - x[0][2]};
    double s1[3] = {x[2][0] - x[1][0], x[2][1] - x[1][1], x[2][2] - x[1][2]};
    double cx[3];
    cross3(s0, s1, cx);
    return 0.5 * sqrt(dot3(cx, cx));
}

// volumeenclosed_integrand (functional.c:2131-2139): |(x0 x x1).x2| / 6
__device__ __forceinline__ double volumeenclosed_of(const double (*x)[3]) {
    double cx[3];
    cross3(x[0], x[1], cx);
    return fabs(dot3(cx, x[2])) / 6.0;
}

// volume_integrand (functional.c:2185-2198): |s10.(s20 x s30)| / 6
__device__ __forceinline__ double volume_of(const double (*x)[3]) {
    double s10[3], s20[3], s30[3], cx[3];
    for (int k = 0; k < 3; k++) { s10[k] = x[1][k] - x[0][k]; s20[k] = x[2][k] - x[0][k]; s30[k] = x[3][k] - x[0][k]; }
    cross3(s20, s30, cx);
    return fabs(dot3(s10, cx)) / 6.0;
}

// areaenclosed_integrand (functional.c:2000-2016)
__device__ __forceinline__ double areaenclosed_of(const double (*x)[3], int dim) {
    if (dim == 2) return 0.5 * fabs(x[0][0] * x[1][1] - x[0][1] * x[1][0]);
    double cx[3];
    cross3(x[0], x[1], cx);
    return 0.5 * sqrt(dot3(cx, cx));
}

// length_gradient_scale (functional.c:1960-1972): -+ s0/|s0|
__device__ __forceinline__ bool length_grad(const double (*x)[3], double (*g)[3]) {
    double s0[3] = {x[1][0] - x[0][0], x[1][1] - x[0][1], x[1][2] - x[0][2]};
    double norm = sqrt(dot3(s0, s0));
    if (norm < MCU_EPS) return false;
    double a = 1.0 / norm;
    for (int k = 0; k < 3; k++) { g[0][k] = -a * s0[k]; g[1][k] = a * s0[k]; }
    return true;
}

// area_gradient_scale (functional.c:2079-2103):
//   n = s0 x s1 ; v0 += (n x s1)/(2|n|) ; v2 += (n x s0)/(2|n|) ; v1 -= ((n x s0)+(n x s1))/(2|n|)
__device__ __forceinline__ bool area_grad(const double (*x)[3], double (*g)[3]) {
    double s0[3] = {x[1][0] - x[0][0], x[1][1] - x[0][1], x[1][2] - x[0][2]};
    double s1[3] = {x[2][0] - x[1][0], x[2][1] - x[1][1], x[2][2] - x[1][2]};
    double s01[3], s010[3], s011[3];
    cross3(s0, s1, s01);
    double n2 = dot3(s01, s01);
    double norm = sqrt(n2);
    if (norm < MCU_EPS) return false;
    cross3(s01, s0, s010);
    cross3(s01, s1, s011);
    double a = 0.5 / norm;
    for (int k = 0; k < 3; k++) {
        g[0][k] = a * s011[k];
        g[2][k] = a * s010[k];
        g[1][k] = -a * (s010[k] + s011[k]);
    }
    return true;
}

// volumeenclosed_gradient (functional.c:2142-2164): sigma/6 * (x1 x x2, x2 x x0, x0 x x1)
// returns 0 ok, 1 → VolEnclZero
__device__ __forceinline__ int volumeenclosed_grad(const double (*x)[3], double (*g)[3]) {
    double c01[3], c12[3], c20[3];
    cross3(x[0], x[1], c01);
    double dot = dot3(c01, x[2]);
    if (fabs(dot) <= DBL_MIN) return 1;
    double s = (dot > 0 ? 1.0 : -1.0) / 6.0;
    cross3(x[1], x[2], c12);
    cross3(x[2], x[0], c20);
    for (int k = 0; k < 3; k++) { g[2][k] = s * c01[k]; g[0][k] = s * c12[k]; g[1][k] = s * c20[k]; }
    return 0;
}

// volume_gradient_scale (functional.c:2201-2227)
__device__ __forceinline__ bool volume_grad(const double (*x)[3], double (*g)[3]) {
    double s10[3], s20[3], s30[3], s31[3], s21[3], cx[3];
    for (int k = 0; k < 3; k++) {
        s10[k] = x[1][k] - x[0][k]; s20[k] = x[2][k] - x[0][k]; s30[k] = x[3][k] - x[0][k];
        s31[k] = x[3][k] - x[1][k]; s21[k] = x[2][k] - x[1][k];
    }
    cross3(s20, s30, cx);
    double uu = dot3(s10, cx);
    uu = (uu > 0 ? 1.0 : -1.0) / 6.0;
    for (int k = 0; k < 3; k++) g[1][k] = uu * cx[k];
    cross3(s31, s21, cx);
    for (int k = 0; k < 3; k++) g[0][k] = uu * cx[k];
    cross3(s30, s10, cx);
    for (int k = 0; k < 3; k++) g[2][k] = uu * cx[k];
    cross3(s10, s20, cx);
    for (int k = 0; k < 3; k++) g[3][k] = uu * cx[k];
    return true;
}

// ---------------------------------------------------------------------------------
// Compensated (sum, comp) pairs for deterministic block reductions of totals.
// The reference Kahan-sums in ascending element order (functional.c:253); we carry an
// error term through a fixed reduction tree instead, which is at least as accurate.
// ---------------------------------------------------------------------------------
struct CompSum {
    double s, c;
};
__device__ __forceinline__ void comp_add(CompSum &a, double v) {
    double s, e;
    two_sum(a.s, v, s, e);
    a.s = s;
    a.c = __dadd_rn(a.c, e);
}
__device__ __forceinline__ void comp_merge(CompSum &a, const CompSum &b) {
    double s, e;
    two_sum(a.s, b.s, s, e);
    a.s = s;
    a.c = __dadd_rn(__dadd_rn(a.c, b.c), e);
}
__device__ __forceinline__ CompSum warp_reduce(CompSum v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        CompSum o;
        o.s = __shfl_down_sync(0xffffffffu, v.s, off);
        o.c = __shfl_down_sync(0xffffffffu, v.c, off);
        comp_merge(v, o);
    }
    return v;
}
// result valid in thread 0
template <int BLOCK>
__device__ __forceinline__ CompSum block_reduce(CompSum v) {
    __shared__ CompSum sh[BLOCK / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_reduce(v);
    if (lane == 0) sh[w] = v;
    __syncthreads();
    CompSum r = {0.0, 0.0};
    if (w == 0) {
        if (lane < BLOCK / 32) r = sh[lane];
        r = warp_reduce(r);
    }
    return r;
}

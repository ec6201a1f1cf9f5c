// mcu_integral.cu — quadrature of an integrand over line, area and volume elements on the device.
//
// What it replaces: integrate_integrate (src/geometry/integrate.c:769-800), the reference's default ("old")
// element integrator behind LineIntegral / AreaIntegral / VolumeIntegral (lineintegral_integrand functional.c:5176-5245,
// areaintegral_integrand 5354-5428, volumeintegral_integrand 5471-5546): integrate the integrand over the element in
// barycentric coordinates with an adaptive nested rule, then multiply by the element size.
//   lines      Gauss 7 / Kronrod 15 pair (integrate.c:38-196): accept when |K15 - G7| * 2^-depth / |global estimate|
//              < 1e-6, else bisect and average the halves;
//   triangles  Grundmann-Moeller rules of degree 3, 5, 7 on 4 / 10 / 20 nested points (integrate.c:205-415, "Walkington"):
//              accept the degree-5 value when it agrees with degree 3, else the degree-7 value when it agrees with
//              degree 5, else split into four and average.
//   tetrahedra the 35- and 56-point symmetric rules of Shunn & Ham, J. Comput. Appl. Math. 236, 4348 (2012), which the
//              reference tabulates (integrate.c:480-577): accept the 56-point value when |R56 - R35| * 8^-depth /
//              |global estimate| < 1e-6, else split into the reference's eight sub-tetrahedra (integrate.c:662-669) and
//              average (integrate_volint, integrate.c:677-752).
// In the reference the integrand is a Morpho closure called through the VM at every node.  A closure cannot
// run on the device; here the integrand is a small postfix PROGRAM (mcu_expr_op, include/morpho_cuda.h) over the
// interpolated position and the interpolated components of P1 fields — what a closure→device translator
// would emit, and what morpho_b200/functionals.py produces by tracing a Python callable.
//
// The rule constants are generated, not tabulated: Gauss-Kronrod nodes/weights are the classical 15-point values,
// the simplex rules come from the Grundmann-Moeller formula  w_i = (-1)^i 2^-2s (d+n-2i)^d / (i! (d+n-i)!),
// nodes (2 beta + 1)/(d+n-2i), |beta| = s-i  (d = 2s+1, n = 2).
// The recursion is unrolled into a depth-first walk with an explicit stack; sub-elements are addressed in the
// barycentric coordinates of the root element, so a leaf contributes  value * (1/2 or 1/4)^depth  — the same number
// the reference's nested averages produce, up to the association of the sums (1e-16).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "mcu_host.h"
#include "mcu_integrands.cuh"      // perpendicular / p1grad_tri: the P1 gradient of GradSq, reused by grad()
#include "mcu_quadrules.h"

extern long g_mcu_launches;

#define MI_MAXQ 40        // interpolated field components + element constants (tangent / normal / gradients)
#define MI_MAXG 12        // field components whose gradient the integrand uses
#define MI_MAXOPS 256
#define MI_STACK 24       // expression stack
#define MI_MAXDEPTH_LINE 30
#define MI_MAXDEPTH_TRI 10
#define MI_MAXDEPTH_TET 5   // the reference recurses without a limit here; 8^5 leaves of 91 nodes bound a thread's work
#define MI_TET_A 35
#define MI_TET_B 56
#define MI_GOAL 1e-6      // INTEGRATE_ACCURACYGOAL (integrate.h:22)

struct MiRules {
    double gk_t[15], gk_wg[15], gk_wk[15];     // line nodes in [0,1] and the two weight sets (Gauss weights 0 beyond 7)
    double tri[20][3];                         // triangle nodes (barycentric)
    double w1[2], w2[3], w3[4];                // weights per group: centroid, degree-3 ring, degree-5 ring, degree-7 ring
    double tet[MI_TET_A + MI_TET_B][4];        // tetrahedron nodes (barycentric): the 35-point rule, then the 56-point rule
    double tetw[MI_TET_A + MI_TET_B];
};
__constant__ MiRules c_rules;
__constant__ mcu_expr_op c_prog[MI_MAXOPS];

// ---- the rule-table integrator (integrate.c:2168-2865), selected by a `method` dictionary ---------------------------
// A configured integrator is a CHAIN of rules rule -> ext -> ext ... over one node array (an extension re-uses the
// function values of the rule it extends) and, when the rule has no extension, a second chain for the error rule.
#define MQ_MAXLEV 6
#define MQ_MAXNODES 126
#define MQ_MAXW 256
#define MQ_ZTOL 1e-15           // INTEGRATE_ZEROCHECK
#define MQ_MAXIT 1000           // INTEGRATE_MAXITERATIONS
struct MqChain {
    int nlev, nn[MQ_MAXLEV], woff[MQ_MAXLEV];
    double nodes[MQ_MAXNODES * 4];
    double w[MQ_MAXW];
};
struct MqMethod {
    int adapt, has_err;
    MqChain main, err;
};
__constant__ MqMethod c_method;
// subdivision rules (integrate.c:1988-2157): new points = midpoints, new elements by vertex index (original vertices first)
__constant__ unsigned char c_sub_line[2][2] = {{2, 1}, {0, 2}};
__constant__ unsigned char c_sub_tri[4][3] = {{0, 3, 5}, {3, 1, 4}, {3, 4, 5}, {5, 4, 2}};

struct MiView {
    int dim, g, nops, nq, nel;
    const double *vert;
    const int *rix;
    const int *eids;                           // selection or nullptr
    int n;                                     // elements to visit
    const double *fld[MI_MAXQ];                // per flattened component: field base pointer
    int fstride[MI_MAXQ], fcomp[MI_MAXQ];      // psize of that field, component inside an entry
    // Element constants are carried as extra "quantities" with the same value at every vertex (their interpolant is
    // the constant): slots nq_in .. nq-1, filled by mi_derive from the (possibly perturbed) vertex data.
    int nq_in;                                 // components read from fields; the rest are derived
    int tn_slot;                               // first of 3 slots holding the tangent (g = 1) / normal (g = 2), or -1
    int ngrad, gsrc[MI_MAXG], gslot[MI_MAXG];  // gradient of component gsrc[i] in slots gslot[i] .. +2
    int method;                                // 0: integrate_integrate (the default integrator); 1: c_method
    unsigned *fail;                            // device word: set when an element needs more work items than a thread holds
};

// morpho_extendedcomparevalue on two floats (value.c:48-68): 0 when equal within DBL_EPSILON relative, +1 when b > a, else -1
__device__ __forceinline__ int mi_cmp(double a, double b) {
    if (a == b) return 0;
    const double diff = fabs(a - b), absa = fabs(a), absb = fabs(b), absmax = absa > absb ? absa : absb;
    if (diff == 0.0 || (absmax > DBL_MIN && diff / absmax <= DBL_EPSILON)) return 0;
    return b > a ? 1 : -1;
}

// the postfix machine: ops c_prog[off .. off + len)
__device__ __forceinline__ double mi_run(int off, int len, const double *x, const double *q) {
    double st[MI_STACK], rg[32];
    int sp = 0;
    for (int i = off; i < off + len; i++) {
        const mcu_expr_op op = c_prog[i];
        switch (op.op) {
            case MCU_X_CONST: st[sp++] = op.c; break;
            case MCU_X_POS: st[sp++] = x[op.a]; break;
            case MCU_X_FIELD: st[sp++] = q[op.a]; break;
            case MCU_X_ADD: sp--; st[sp - 1] = st[sp - 1] + st[sp]; break;
            case MCU_X_SUB: sp--; st[sp - 1] = st[sp - 1] - st[sp]; break;
            case MCU_X_MUL: sp--; st[sp - 1] = st[sp - 1] * st[sp]; break;
            case MCU_X_DIV: sp--; st[sp - 1] = st[sp - 1] / st[sp]; break;
            case MCU_X_NEG: st[sp - 1] = -st[sp - 1]; break;
            case MCU_X_POWI: st[sp - 1] = pow(st[sp - 1], (double) op.a); break;
            case MCU_X_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
            case MCU_X_SIN: st[sp - 1] = sin(st[sp - 1]); break;
            case MCU_X_COS: st[sp - 1] = cos(st[sp - 1]); break;
            case MCU_X_EXP: st[sp - 1] = exp(st[sp - 1]); break;
            case MCU_X_LOG: st[sp - 1] = log(st[sp - 1]); break;
            case MCU_X_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
            case MCU_X_POW: sp--; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
            case MCU_X_ATAN2: sp--; st[sp - 1] = atan2(st[sp - 1], st[sp]); break;
            case MCU_X_TAN: st[sp - 1] = tan(st[sp - 1]); break;
            case MCU_X_ASIN: st[sp - 1] = asin(st[sp - 1]); break;
            case MCU_X_ACOS: st[sp - 1] = acos(st[sp - 1]); break;
            case MCU_X_ATAN: st[sp - 1] = atan(st[sp - 1]); break;
            case MCU_X_SINH: st[sp - 1] = sinh(st[sp - 1]); break;
            case MCU_X_COSH: st[sp - 1] = cosh(st[sp - 1]); break;
            case MCU_X_TANH: st[sp - 1] = tanh(st[sp - 1]); break;
            case MCU_X_FLOOR: st[sp - 1] = floor(st[sp - 1]); break;
            case MCU_X_CEIL: st[sp - 1] = ceil(st[sp - 1]); break;
            case MCU_X_LOG10: st[sp - 1] = log10(st[sp - 1]); break;
            case MCU_X_LT: sp--; st[sp - 1] = mi_cmp(st[sp - 1], st[sp]) > 0 ? 1.0 : 0.0; break;
            case MCU_X_LE: sp--; st[sp - 1] = mi_cmp(st[sp - 1], st[sp]) >= 0 ? 1.0 : 0.0; break;
            case MCU_X_EQ: sp--; st[sp - 1] = mi_cmp(st[sp - 1], st[sp]) == 0 ? 1.0 : 0.0; break;
            case MCU_X_NE: sp--; st[sp - 1] = mi_cmp(st[sp - 1], st[sp]) != 0 ? 1.0 : 0.0; break;
            case MCU_X_NOT: st[sp - 1] = st[sp - 1] != 0.0 ? 0.0 : 1.0; break;
            case MCU_X_SELECT: sp -= 2; st[sp - 1] = st[sp - 1] != 0.0 ? st[sp] : st[sp + 1]; break;
            case MCU_X_STORE: rg[op.a & 31] = st[sp - 1]; break;
            case MCU_X_LOAD: st[sp++] = rg[op.a & 31]; break;
            default: break;
        }
    }
    return sp > 0 ? st[sp - 1] : 0.0;
}
__device__ __forceinline__ double mi_eval(const MiView &v, const double *x, const double *q) { return mi_run(0, v.nops, x, q); }

// integrand at the point with barycentric coordinates lam[0..nv) of the ROOT element
template <int NV>
__device__ __forceinline__ double mi_at(const MiView &v, const double (*xv)[3], const double (*qv)[MI_MAXQ], const double *lam) {
    double x[3] = {0.0, 0.0, 0.0}, q[MI_MAXQ];
#pragma unroll
    for (int k = 0; k < NV; k++)
#pragma unroll
        for (int j = 0; j < 3; j++) x[j] += lam[k] * xv[k][j];          // integrate_interpolateposition*
    for (int i = 0; i < v.nq; i++) {
        double s = lam[0] * qv[0][i];
#pragma unroll
        for (int k = 1; k < NV; k++) s += lam[k] * qv[k][i];            // integrate_interpolatequantities*
        q[i] = s;
    }
    return mi_eval(v, x, q);
}

__device__ double mi_line(const MiView &v, const double (*xv)[3], const double (*qv)[MI_MAXQ]) {
    unsigned stack_idx[MI_MAXDEPTH_LINE + 2];
    int stack_d[MI_MAXDEPTH_LINE + 2];
    int top = 0;
    stack_idx[0] = 0; stack_d[0] = 0; top = 1;
    double sum = 0.0, gest = 0.0;
    while (top > 0) {
        top--;
        const int d = stack_d[top];
        const unsigned idx = stack_idx[top];
        const double w = ldexp(1.0, -d), tlo = (double) idx * w;
        double r1 = 0.0, r2 = 0.0;
        for (int i = 0; i < 15; i++) {
            const double t = tlo + c_rules.gk_t[i] * w;
            const double lam[2] = {1.0 - t, t};
            const double f = mi_at<2>(v, xv, qv, lam);
            if (i < 7) r1 += f * c_rules.gk_wg[i];
            r2 += f * c_rules.gk_wk[i];
        }
        r1 *= 0.5; r2 *= 0.5;
        if (d == 0) gest = fabs(r2);
        double eps = (r2 - r1) * w;
        if (gest > MCU_EPS) eps /= gest;
        if (fabs(eps) < MI_GOAL || d >= MI_MAXDEPTH_LINE) sum += r2 * w;
        else {
            stack_d[top] = d + 1; stack_idx[top] = 2 * idx + 1; top++;     // right half (visited second)
            stack_d[top] = d + 1; stack_idx[top] = 2 * idx; top++;         // left half
        }
    }
    return sum;
}

__device__ double mi_tri(const MiView &v, const double (*xv)[3], const double (*qv)[MI_MAXQ]) {
    // a sub-triangle = three corners in the barycentric coordinates of the root
    double sb[3 * MI_MAXDEPTH_TRI + 4][9];
    int sd[3 * MI_MAXDEPTH_TRI + 4];
    int top = 0;
    for (int i = 0; i < 9; i++) sb[0][i] = (i % 4 == 0) ? 1.0 : 0.0;
    sd[0] = 0; top = 1;
    double sum = 0.0, gest = 0.0;
    while (top > 0) {
        top--;
        const int d = sd[top];
        double B[9];
        for (int i = 0; i < 9; i++) B[i] = sb[top][i];
        const double af = ldexp(1.0, -2 * d);
        double r[20];
        auto node = [&](int i) {
            double lam[3];
            for (int k = 0; k < 3; k++) lam[k] = c_rules.tri[i][0] * B[k] + c_rules.tri[i][1] * B[3 + k] + c_rules.tri[i][2] * B[6 + k];
            return mi_at<3>(v, xv, qv, lam);
        };
        for (int i = 0; i < 10; i++) r[i] = node(i);
        const double rr = r[1] + r[2] + r[3];
        const double rr2 = r[4] + r[5] + r[6] + r[7] + r[8] + r[9];
        const double r1 = c_rules.w1[0] * r[0] + c_rules.w1[1] * rr;
        const double r2 = c_rules.w2[0] * r[0] + c_rules.w2[1] * rr + c_rules.w2[2] * rr2;
        if (d == 0) gest = fabs(r2);
        double eps = (r2 - r1) * af;
        if (gest > MCU_EPS) eps /= gest;
        if (fabs(eps) < MI_GOAL) { sum += 2.0 * r2 * af; continue; }
        for (int i = 10; i < 20; i++) r[i] = node(i);
        double rr3 = 0.0;
        for (int i = 10; i < 20; i++) rr3 += r[i];
        const double r3 = c_rules.w3[0] * r[0] + c_rules.w3[1] * rr + c_rules.w3[2] * rr2 + c_rules.w3[3] * rr3;
        if (d == 0) gest = fabs(2.0 * r3);
        eps = (r2 - r3) * af;
        if (gest > MCU_EPS) eps /= gest;
        if (fabs(eps) < MI_GOAL || d >= MI_MAXDEPTH_TRI) { sum += 2.0 * r3 * af; continue; }
        // quadrisect: (0, 01, 20) (01, 1, 12) (20, 12, 2) (01, 12, 20); pushed in reverse so they are visited in this order
        double m01[3], m12[3], m20[3];
        for (int k = 0; k < 3; k++) { m01[k] = 0.5 * (B[k] + B[3 + k]); m12[k] = 0.5 * (B[3 + k] + B[6 + k]); m20[k] = 0.5 * (B[6 + k] + B[k]); }
        const double *kids[4][3] = {{B, m01, m20}, {m01, B + 3, m12}, {m20, m12, B + 6}, {m01, m12, m20}};
        for (int c = 3; c >= 0; c--) {
            for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) sb[top][3 * j + k] = kids[c][j][k];
            sd[top] = d + 1; top++;
        }
    }
    return sum;
}

// integrate_volint (integrate.c:677-752).  A sub-tetrahedron = four corners in the barycentric coordinates of the root.
// Children (integrate.c:662-669) of corners 0..3 and edge midpoints 4 = 01, 5 = 02, 6 = 03, 7 = 12, 8 = 13, 9 = 23.
__constant__ unsigned char c_tetsub[8][4] = {{1, 4, 7, 8}, {0, 4, 7, 9}, {0, 4, 8, 9}, {4, 7, 8, 9}, {0, 5, 7, 9}, {0, 6, 8, 9}, {2, 5, 7, 9}, {3, 6, 8, 9}};

__device__ double mi_tet(const MiView &v, const double (*xv)[3], const double (*qv)[MI_MAXQ]) {
    double sb[7 * MI_MAXDEPTH_TET + 2][16];
    int sd[7 * MI_MAXDEPTH_TET + 2];
    int top = 0;
    for (int i = 0; i < 16; i++) sb[0][i] = (i % 5 == 0) ? 1.0 : 0.0;
    sd[0] = 0; top = 1;
    double sum = 0.0, gest = 0.0;
    while (top > 0) {
        top--;
        const int d = sd[top];
        double B[10][4];
        for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) B[j][k] = sb[top][4 * j + k];
        const double af = ldexp(1.0, -3 * d);
        double r1 = 0.0, r2 = 0.0;
        for (int i = 0; i < MI_TET_A + MI_TET_B; i++) {
            double lam[4];
            for (int k = 0; k < 4; k++)
                lam[k] = c_rules.tet[i][0] * B[0][k] + c_rules.tet[i][1] * B[1][k] + c_rules.tet[i][2] * B[2][k] + c_rules.tet[i][3] * B[3][k];
            const double f = mi_at<4>(v, xv, qv, lam);
            if (i < MI_TET_A) r1 += c_rules.tetw[i] * f;
            else r2 += c_rules.tetw[i] * f;
        }
        if (d == 0) gest = fabs(r2);
        double eps = (r2 - r1) * af;
        if (gest > MCU_EPS) eps /= gest;
        if (fabs(eps) < MI_GOAL || d >= MI_MAXDEPTH_TET) { sum += r2 * af; continue; }
        const int pair[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
        for (int m = 0; m < 6; m++) for (int k = 0; k < 4; k++) B[4 + m][k] = 0.5 * (B[pair[m][0]][k] + B[pair[m][1]][k]);
        for (int c = 7; c >= 0; c--) {                         // pushed in reverse: visited in the reference's order
            for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) sb[top][4 * j + k] = B[c_tetsub[c][j]][k];
            sd[top] = d + 1; top++;
        }
    }
    return sum;
}

// One work item: weight, value, lower-order value, error estimate and the element's vertices in the barycentric
// coordinates of the ROOT element (the reference keeps them on a vertex stack and refers to them by id).
template <int NB> struct MqItem { double weight, val, lval, err; double v[NB * NB]; };

// integrator_quadrature (integrate.c:2445-2504) with `chain` as rule -> ext -> ...
template <int NB>
__device__ void mq_chain(const MiView &v, const double (*xv)[3], const double (*qv)[MI_MAXQ], const MqChain &ch, MqItem<NB> &it, double *f) {
    auto evalfn = [&](int i0, int i1) {
        for (int i = i0; i < i1; i++) {
            double lam[NB];
            for (int j = 0; j < NB; j++) lam[j] = 0.0;
            for (int k = 0; k < NB; k++) for (int j = 0; j < NB; j++) lam[j] += it.v[k * NB + j] * ch.nodes[i * NB + k];   // integrator_transformtorefelement
            f[i] = mi_at<NB>(v, xv, qv, lam);
        }
    };
    auto wsum = [&](int lev) { double r = 0.0; for (int i = 0; i < ch.nn[lev]; i++) r += f[i] * ch.w[ch.woff[lev] + i]; return r; };
    evalfn(0, ch.nn[0]);
    double rprev = wsum(0), rcur = rprev, eps = 0.0;
    it.lval = it.val = it.weight * rprev;
    if (ch.nlev > 1) {
        for (int l = 1; l < ch.nlev; l++) {
            evalfn(ch.nn[l - 1], ch.nn[l]);
            if (l > 1) rprev = rcur;
            rcur = wsum(l);
            eps = fabs(rcur - rprev);
            if (fabs(rcur) < MQ_ZTOL || fabs(eps / rcur) < MI_GOAL) break;
        }
        it.lval = it.weight * rprev;
        it.val = it.weight * rcur;
        it.err = it.weight * eps;
    }
}
template <int NB>
__device__ void mq_quadrature(const MiView &v, const double (*xv)[3], const double (*qv)[MI_MAXQ], MqItem<NB> &it, double *f) {
    mq_chain<NB>(v, xv, qv, c_method.main, it, f);
    if (c_method.main.nlev == 1 && c_method.has_err) {      // no extension: the error rule gives the better value and the error
        const double temp = it.val;
        mq_chain<NB>(v, xv, qv, c_method.err, it, f);
        it.lval = temp;
        it.err = fabs(it.val - temp);
    }
}

// integrator_integrate (integrate.c:2759-2835): a binary heap of work items ordered by error estimate; the worst item is
// subdivided until the summed error is below the goal.  `heap` holds `cap` items; *overflow is set when that is not enough.
template <int NB>
__device__ double mq_integrate(const MiView &v, const double (*xv)[3], const double (*qv)[MI_MAXQ], MqItem<NB> *heap, int cap, bool &overflow) {
    constexpr int NSUB = NB == 2 ? 2 : (NB == 3 ? 4 : 8), NPTS = NB == 2 ? 1 : (NB == 3 ? 3 : 6);
    double f[MQ_MAXNODES];
    int count = 0;
    auto push = [&](const MqItem<NB> &w) {                   // integrator_pushworkitem
        heap[count++] = w;
        for (int i = count - 1, p; i > 0; i = p) {
            p = (i - 1) / 2;
            if (heap[i].err > heap[p].err) { MqItem<NB> t = heap[i]; heap[i] = heap[p]; heap[p] = t; } else break;
        }
    };
    auto pop = [&](MqItem<NB> &w) {                          // integrator_popworkitem
        w = heap[0];
        const int n = count - 1;
        if (n > 0) heap[0] = heap[n];
        count--;
        for (int i = 0, p, q; i < n; i = p) {
            p = 2 * i + 1; q = p + 1;
            if (q < n && heap[q].err > heap[p].err) p = q;
            if (p < n && heap[p].err > heap[i].err) { MqItem<NB> t = heap[i]; heap[i] = heap[p]; heap[p] = t; } else break;
        }
    };
    double val = 0.0, errest = 0.0;
    auto estimate = [&]() {                                  // integrator_estimate: Kahan sums over the heap, last entry first
        double sv = 0.0, cv = 0.0, se = 0.0, ce = 0.0;
        for (int i = count - 1; i >= 0; i--) {
            const double yv = heap[i].val - cv, ye = heap[i].err - ce;
            const double tv = sv + yv, te = se + ye;
            cv = (tv - sv) - yv; ce = (te - se) - ye;
            sv = tv; se = te;
        }
        val = sv; errest = se;
    };
    MqItem<NB> work;
    work.weight = 1.0; work.val = work.lval = work.err = 0.0;
    for (int i = 0; i < NB * NB; i++) work.v[i] = (i % (NB + 1) == 0) ? 1.0 : 0.0;
    mq_quadrature<NB>(v, xv, qv, work, f);
    push(work);
    estimate();
    if (c_method.adapt)
        for (int it = 0; it <= MQ_MAXIT; it++) {
            if (fabs(val) < MQ_ZTOL || fabs(errest / val) < MI_GOAL) break;
            if (count - 1 + NSUB > cap) { overflow = true; break; }
            pop(work);
            // integrator_subdivide: midpoints in the coordinates of the root element
            double P[NB + NPTS][NB];
            for (int k = 0; k < NB; k++) for (int j = 0; j < NB; j++) P[k][j] = work.v[k * NB + j];
            if (NB == 2) { for (int j = 0; j < NB; j++) P[2][j] = 0.5 * P[0][j] + 0.5 * P[1][j]; }
            else if (NB == 3) {
                for (int j = 0; j < NB; j++) {
                    P[3][j] = 0.5 * P[0][j] + 0.5 * P[1][j] + 0.0 * P[2][j];
                    P[4][j] = 0.0 * P[0][j] + 0.5 * P[1][j] + 0.5 * P[2][j];
                    P[5][j] = 0.5 * P[0][j] + 0.0 * P[1][j] + 0.5 * P[2][j];
                }
            } else {
                const int pair[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
                for (int m = 0; m < 6; m++) for (int j = 0; j < NB; j++) {
                    double s = 0.0;
                    for (int k = 0; k < 4; k++) s += ((k == pair[m][0] || k == pair[m][1]) ? 0.5 : 0.0) * P[k][j];
                    P[4 + m][j] = s;
                }
            }
            MqItem<NB> kid[NSUB];
            bool bad = false;
            for (int c = 0; c < NSUB; c++) {
                kid[c].val = kid[c].lval = kid[c].err = 0.0;
                kid[c].weight = work.weight * (1.0 / NSUB);
                if (kid[c].weight < val * MCU_EPS) bad = true;         // IntgrtrSbdvsns: the reference gives up with an error
                for (int k = 0; k < NB; k++) {
                    const int src = NB == 2 ? c_sub_line[c][k] : (NB == 3 ? c_sub_tri[c][k] : c_tetsub[c][k]);
                    for (int j = 0; j < NB; j++) kid[c].v[k * NB + j] = P[src][j];
                }
            }
            if (bad) { overflow = true; break; }
            for (int c = 0; c < NSUB; c++) mq_quadrature<NB>(v, xv, qv, kid[c], f);
            // integrator_sharpenerrorestimate (Laurie, BIT 23 (1983) 258)
            double a2 = 0.0, b2 = 0.0;
            for (int c = 0; c < NSUB; c++) { a2 += kid[c].val; b2 += kid[c].lval; }
            const double a1 = work.val, b1 = work.lval;
            if (fabs(a2 - a1) < fabs(b2 - b1) && fabs(a2 - b2) < fabs(a1 - b1)) {
                const double sigma = fabs((a2 - a1) / (b2 - b1 - a2 + a1));
                for (int c = 0; c < NSUB; c++) kid[c].err *= sigma;
            }
            // integrator_update
            double dval = 0.0, derr = 0.0;
            val -= work.val; errest -= work.err;
            for (int c = 0; c < NSUB; c++) { dval += kid[c].val; derr += kid[c].err; push(kid[c]); }
            val += dval; errest += derr;
        }
    estimate();
    return val;
}

// items a thread keeps in local memory (16 KB): enough for 255 / 52 / 14 subdivisions
#define MQ_CAP_LINE 256
#define MQ_CAP_TRI 157
#define MQ_CAP_TET 102
__device__ double mi_method(const MiView &v, const double (*xv)[3], const double (*qv)[MI_MAXQ]) {
    bool overflow = false;
    double r;
    double raw[2048];                  // one buffer for the three item types (all doubles): 256 x 8, 157 x 13, 102 x 20 <= 2048
    static_assert(sizeof(MqItem<2>) * MQ_CAP_LINE <= sizeof(raw) && sizeof(MqItem<3>) * MQ_CAP_TRI <= sizeof(raw) && sizeof(MqItem<4>) * MQ_CAP_TET <= sizeof(raw), "heap buffer");
    if (v.g == 1) r = mq_integrate<2>(v, xv, qv, reinterpret_cast<MqItem<2> *>(raw), MQ_CAP_LINE, overflow);
    else if (v.g == 2) r = mq_integrate<3>(v, xv, qv, reinterpret_cast<MqItem<3> *>(raw), MQ_CAP_TRI, overflow);
    else r = mq_integrate<4>(v, xv, qv, reinterpret_cast<MqItem<4> *>(raw), MQ_CAP_TET, overflow);
    if (overflow && v.fail) atomicOr(v.fail, 1u);
    return r;
}

// tangent() / normal() / grad(q) of the element (integral_evaluatetangent, integral_evaluatenormal,
// integral_evaluategradient → gradsq_evaluategradient; functional.c:4402-4425, 4444-4470, 4716-4726)
__device__ __forceinline__ void mi_derive(const MiView &v, const double (*xv)[3], double (*qv)[MI_MAXQ]) {
    const int nv = v.g + 1;
    if (v.tn_slot >= 0) {
        double t[3];
        if (v.g == 1) { t[0] = xv[1][0] - xv[0][0]; t[1] = xv[1][1] - xv[0][1]; t[2] = xv[1][2] - xv[0][2]; }
        else {
            const double s0[3] = {xv[1][0] - xv[0][0], xv[1][1] - xv[0][1], xv[1][2] - xv[0][2]};
            const double s1[3] = {xv[2][0] - xv[1][0], xv[2][1] - xv[1][1], xv[2][2] - xv[1][2]};
            cross3(s0, s1, t);
        }
        const double nrm = sqrt(dot3(t, t));
        if (fabs(nrm) > MCU_EPS) { const double r = 1.0 / nrm; t[0] *= r; t[1] *= r; t[2] *= r; }
        for (int k = 0; k < nv; k++) for (int j = 0; j < 3; j++) qv[k][v.tn_slot + j] = t[j];
    }
    for (int i = 0; i < v.ngrad; i++) {
        double vals[4] = {qv[0][v.gsrc[i]], qv[1][v.gsrc[i]], qv[2][v.gsrc[i]], v.g == 3 ? qv[3][v.gsrc[i]] : 0.0}, gr[3] = {0.0, 0.0, 0.0};
        if (v.g == 3) p1grad_tet(xv, 1, vals, 1, gr);          // gradsq_evaluategradient3d (functional.c:3584-3620)
        else p1grad_tri(v.dim, xv, 1, vals, 1, gr);
        for (int k = 0; k < nv; k++) for (int j = 0; j < 3; j++) qv[k][v.gslot[i] + j] = gr[j];
    }
}

// integral over one element times its size: "*out *= elref.elementsize" (functional.c:5227, 5410)
__device__ __forceinline__ double mi_element(const MiView &v, const double (*xv)[3], double (*qv)[MI_MAXQ]) {
    if (v.nq > v.nq_in) mi_derive(v, xv, qv);
    double val, size;
    if (v.g == 1) {
        val = v.method ? mi_method(v, xv, qv) : mi_line(v, xv, qv);
        double s[3] = {xv[1][0] - xv[0][0], xv[1][1] - xv[0][1], xv[1][2] - xv[0][2]};
        size = sqrt(dot3(s, s));
    } else if (v.g == 2) {
        val = v.method ? mi_method(v, xv, qv) : mi_tri(v, xv, qv);
        size = area_of(xv);
    } else {
        val = v.method ? mi_method(v, xv, qv) : mi_tet(v, xv, qv);
        size = volume_of(xv);
    }
    return val * size;
}

__device__ __forceinline__ void mi_gather(const MiView &v, int e, double (*xv)[3], double (*qv)[MI_MAXQ]) {
    const int nv = v.g + 1;
    for (int k = 0; k < nv; k++) {
        const int id = v.rix[(size_t) e * nv + k];
        for (int j = 0; j < 3; j++) xv[k][j] = (j < v.dim) ? v.vert[(size_t) id * v.dim + j] : 0.0;
        for (int i = 0; i < v.nq_in; i++) qv[k][i] = v.fld[i][(size_t) id * v.fstride[i] + v.fcomp[i]];
    }
}

__global__ void __launch_bounds__(128) k_integral(MiView v, double *out) {
    int ce = blockIdx.x * 128 + threadIdx.x;
    if (ce >= v.n) return;
    const int e = v.eids ? v.eids[ce] : ce;
    double xv[4][3], qv[4][MI_MAXQ];
    mi_gather(v, e, xv, qv);
    out[e] = mi_element(v, xv, qv);
}

// Central differences of one element with respect to one of its vertices: the coordinates (qbase < 0) or the ncomp
// components of a field stored there (functional_numericalgrad, functional.c:1142-1162; functional_numericalfieldgrad,
// 1305-1328): f(x0 + eps) and f(x0 - eps) are full adaptive quadratures.  One thread per (element, corner); the
// results go to scratch[(ce * nv + corner) * ncomp + c] and are summed per vertex in ascending element id by
// k_integral_gather — the order in which the reference accumulates them.
__global__ void __launch_bounds__(128) k_integral_fd(MiView v, int qbase, int ncomp, double *scratch) {
    const int nv = v.g + 1;
    long t = (long) blockIdx.x * 128 + threadIdx.x;
    if (t >= (long) v.n * nv) return;
    const int ce = (int) (t / nv), corner = (int) (t - (long) ce * nv);
    const int e = v.eids ? v.eids[ce] : ce;
    double xv[4][3], qv[4][MI_MAXQ];
    mi_gather(v, e, xv, qv);
    for (int c = 0; c < ncomp; c++) {
        double *slot = qbase < 0 ? &xv[corner][c] : &qv[corner][qbase + c];
        const double f0 = *slot, eps = fdstepsize(f0);
        *slot = f0 + eps;
        const double fp = mi_element(v, xv, qv);
        *slot = f0 - eps;
        const double fm = mi_element(v, xv, qv);
        *slot = f0;
        scratch[((size_t) ce * nv + corner) * ncomp + c] = (fp - fm) / (2 * eps);
    }
}

__global__ void __launch_bounds__(128) k_integral_gather(int nvtx, int nv, int ncomp, const int *__restrict__ tptr, const int *__restrict__ tidx,
                                                         const int *__restrict__ eids, const int *__restrict__ rix,
                                                         const double *__restrict__ scratch, double *__restrict__ out) {
    int vtx = blockIdx.x * 128 + threadIdx.x;
    if (vtx >= nvtx) return;
    double acc[MI_MAXQ];
    for (int c = 0; c < ncomp; c++) acc[c] = 0.0;
    for (int q = tptr[vtx]; q < tptr[vtx + 1]; q++) {
        const int ce = tidx[q], e = eids ? eids[ce] : ce;
        int corner = 0;
        for (int j = 0; j < nv; j++) if (rix[(size_t) e * nv + j] == vtx) corner = j;
        const double *src = scratch + ((size_t) ce * nv + corner) * ncomp;
        for (int c = 0; c < ncomp; c++) acc[c] = acc[c] + src[c];
    }
    for (int c = 0; c < ncomp; c++) out[(size_t) vtx * ncomp + c] = acc[c];
}

static __global__ void __launch_bounds__(256) k_isum(const double *__restrict__ v, int n, CompSum *partials) {
    CompSum acc = {0.0, 0.0};
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) comp_add(acc, v[i]);
    CompSum r = block_reduce<256>(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
}
static __global__ void __launch_bounds__(256) k_isum_finish(const CompSum *partials, int n, double *out) {
    CompSum acc = {0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += 256) comp_merge(acc, partials[i]);
    CompSum r = block_reduce<256>(acc);
    if (threadIdx.x == 0) out[0] = r.s + r.c;
}

static void make_rules(MiRules &R) {
    // Gauss-Kronrod 7/15 on [-1,1]: (node, Gauss weight, Kronrod weight); the Gauss nodes first (their order fixes
    // only the order of the sums)
    const double xg[4] = {0.949107912342758524526189684047851, 0.741531185599394439863864773280788,
                          0.405845151377397166906606412076961, 0.0};
    const double wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                          0.381830050505118944950369775488975, 0.417959183673469387755102040816327};
    const double wkg[4] = {0.063092092629978553290700663189204, 0.140653259715525918745189590510238,
                           0.190350578064785409913256402421014, 0.209482141084727828012999174891714};
    const double xk[4] = {0.991455371120812639206854697526329, 0.864864423359769072789712788640926,
                          0.586087235467691130294144838258730, 0.207784955007898467600689403773245};
    const double wk[4] = {0.022935322010529224963732008058970, 0.104790010322250183839876322541518,
                          0.169004726639267902826583426598550, 0.204432940075298892414161999234649};
    int n = 0;
    auto put = [&](double x, double g, double k) { R.gk_t[n] = 0.5 * (1.0 + x); R.gk_wg[n] = g; R.gk_wk[n] = k; n++; };
    for (int i = 0; i < 3; i++) { put(-xg[i], wg[i], wkg[i]); put(xg[i], wg[i], wkg[i]); }
    put(0.0, wg[3], wkg[3]);
    for (int i = 0; i < 4; i++) { put(-xk[i], 0.0, wk[i]); put(xk[i], 0.0, wk[i]); }
    // Grundmann-Moeller on the triangle (n = 2), s = 1, 2, 3
    auto fact = [](int k) { double f = 1.0; for (int i = 2; i <= k; i++) f *= i; return f; };
    auto gmw = [&](int s, int i) {
        int d = 2 * s + 1;
        return ((i & 1) ? -1.0 : 1.0) * std::ldexp(1.0, -2 * s) * std::pow((double) (d + 2 - 2 * i), d) / (fact(i) * fact(d + 2 - i));
    };
    // the formula's weights sum to 1/2 (the area of the reference triangle); the integrators double the weighted sum
    for (int s = 1; s <= 3; s++) {
        double *w = (s == 1) ? R.w1 : (s == 2 ? R.w2 : R.w3);
        for (int i = 0; i <= s; i++) w[s - i] = gmw(s, i);      // group index: 0 = centroid (i = s), 1 = denominators 5, ...
    }
    // nodes: centroid; (3,1,1)/5; (5,1,1)/7 and (3,3,1)/7; (7,1,1)/9, (5,3,1)/9 and (3,3,3)/9
    int p = 0;
    auto node = [&](double a, double b, double c, double den) { R.tri[p][0] = a / den; R.tri[p][1] = b / den; R.tri[p][2] = c / den; p++; };
    node(1, 1, 1, 3);
    node(3, 1, 1, 5); node(1, 3, 1, 5); node(1, 1, 3, 5);
    node(5, 1, 1, 7); node(1, 5, 1, 7); node(1, 1, 5, 7); node(3, 3, 1, 7); node(3, 1, 3, 7); node(1, 3, 3, 7);
    node(7, 1, 1, 9); node(1, 7, 1, 9); node(1, 1, 7, 9);
    node(3, 5, 1, 9); node(3, 1, 5, 9); node(5, 3, 1, 9); node(5, 1, 3, 9); node(1, 3, 5, 9); node(1, 5, 3, 9);
    node(3, 3, 3, 9);
    // Tetrahedron: the 35- and 56-point rules of Shunn & Ham (2012) as the reference tabulates them (integrate.c:480-577),
    // stored here by symmetry orbit: generator values a, b, c, the weight, and the reference's order of the
    // permutations inside the orbit (it fixes only the order of the weighted sums).
    struct Orbit { const char *perms; double a, b, c, w; };
    static const char *P4 = "abbb babb bbab bbba", *P1 = "aaaa", *P6 = "aacc acac acca caac caca ccaa";
    static const char *P12A = "abcc bacc acbc bcac accb bcca cabc cbac cacb cbca ccab ccba";
    static const char *P12B = "abbc babc bbac abcb bacb bbca acbb bcab bcba cabb cbab cbba";
    static const Orbit rule35[] = {
        {P4, 0.9197896733368800, 0.0267367755543735, 0.0, 0.0021900463965388},
        {P12A, 0.1740356302468940, 0.7477598884818090, 0.0391022406356488, 0.0143395670177665},
        {P6, 0.4547545999844830, 0.0, 0.0452454000155172, 0.0250305395686746},
        {P12B, 0.5031186450145980, 0.2232010379623150, 0.0504792790607720, 0.0479839333057554},
        {P1, 0.2500000000000000, 0.0, 0.0, 0.0931745731195340}};
    static const Orbit rule56[] = {
        {P4, 0.9551438045408220, 0.0149520651530592, 0.0, 0.0010373112336140},
        {P12A, 0.7799760084415400, 0.1518319491659370, 0.0340960211962615, 0.0096016645399480},
        {P12A, 0.3549340560639790, 0.5526556431060170, 0.0462051504150017, 0.0164493976798232},
        {P12B, 0.5381043228880020, 0.2281904610687610, 0.0055147549744775, 0.0153747766513310},
        {P12B, 0.1961837595745600, 0.3523052600879940, 0.0992057202494530, 0.0293520118375230},
        {P4, 0.5965649956210170, 0.1344783347929940, 0.0, 0.0366291366405108}};
    int t = 0;
    auto expand = [&](const Orbit *o, int norb) {
        for (int k = 0; k < norb; k++)
            for (const char *q = o[k].perms; *q; q += (q[4] ? 5 : 4)) {
                for (int j = 0; j < 4; j++) R.tet[t][j] = q[j] == 'a' ? o[k].a : (q[j] == 'b' ? o[k].b : o[k].c);
                R.tetw[t++] = o[k].w;
            }
    };
    expand(rule35, 5);
    expand(rule56, 6);
}

// arguments shared by mcu_integral and mcu_integral_gradient: validates, uploads the rules and the program, fills the view
// integrator_configure + the upload of the configured chains (c_method)
static bool fill_chain(const McuQuadRules &Q, int rule, MqChain &ch) {
    memset(&ch, 0, sizeof(ch));
    const int nb = Q.rules[rule].grade + 1;
    int woff = 0, last = rule;
    for (int r = rule; r >= 0; r = Q.rules[r].ext) {
        if (ch.nlev == MQ_MAXLEV) return false;
        const McuQuadRule &R = Q.rules[r];
        const int nw = std::max(R.nnodes, (int) R.weights.size());
        if (R.nnodes > MQ_MAXNODES || woff + nw > MQ_MAXW) return false;
        ch.nn[ch.nlev] = R.nnodes; ch.woff[ch.nlev] = woff;
        for (int i = 0; i < R.nnodes; i++) ch.w[woff + i] = R.weights[i];
        woff += R.nnodes;
        ch.nlev++;
        last = r;
    }
    const std::vector<double> &nodes = Q.families[Q.rules[last].family];
    for (int i = 0; i < Q.rules[last].nnodes * nb; i++) ch.nodes[i] = nodes[i];
    return true;
}

static int method_setup(int grade, const mcu_quad_method *method, cudaStream_t st) {
    const McuQuadRules &Q = McuQuadRules::get();
    int rule = -1, errrule = -1;
    char name[sizeof(method->rule) + 1];
    memcpy(name, method->rule, sizeof(method->rule));
    name[sizeof(method->rule)] = 0;
    if (!Q.configure(grade, method->adapt != 0, method->degree, name, rule, errrule)) {
        mcu_set_error("no quadrature rule '%s' / degree %d for grade %d (the reference raises IntgrtrRlNtFnd / IntgrtrRlUnavlb)", name, method->degree, grade);
        return MCU_ERR_UNSUPPORTED;
    }
    static MqMethod M;                          // pinned for the asynchronous copy below
    memset(&M, 0, sizeof(M));
    M.adapt = method->adapt != 0;
    M.has_err = errrule >= 0;
    if (!fill_chain(Q, rule, M.main) || (errrule >= 0 && !fill_chain(Q, errrule, M.err))) { mcu_set_error("quadrature rule chain too long"); return MCU_ERR_UNSUPPORTED; }
    if (cudaMemcpyToSymbolAsync(c_method, &M, sizeof(M), 0, cudaMemcpyHostToDevice, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return MCU_ERR_CUDA;
    return MCU_OK;
}

extern "C" int mcu_debug_quadrule(int index, char name[32], int info[5], double *nodes, double *weights, int cap) {
    const McuQuadRules &Q = McuQuadRules::get();
    if (index < 0 || index >= (int) Q.rules.size()) return MCU_ERR_ARGS;
    const McuQuadRule &R = Q.rules[index];
    if (name) { memset(name, 0, 32); strncpy(name, R.name.c_str(), 31); }
    if (info) { info[0] = R.grade; info[1] = R.order; info[2] = R.nnodes; info[3] = R.ext; info[4] = (int) R.weights.size(); }
    const int nb = R.grade + 1;
    if (nodes) for (int i = 0; i < R.nnodes * nb && i < cap; i++) nodes[i] = Q.families[R.family][i];
    if (weights) for (int i = 0; i < (int) R.weights.size() && i < cap; i++) weights[i] = R.weights[i];
    return MCU_OK;
}

extern "C" int mcu_debug_quadconfigure(int grade, int adapt, int degree, const char *name, int out[2]) {
    int rule, err;
    if (!McuQuadRules::get().configure(grade, adapt != 0, degree, name, rule, err)) return MCU_ERR_UNSUPPORTED;
    out[0] = rule; out[1] = err;
    return MCU_OK;
}

static unsigned *g_integral_fail = nullptr;     // device word: an element ran out of work items (rule-table integrator)

static int integral_setup(mcu_mesh *m, int grade, int nops, const mcu_expr_op *prog, int nfields, mcu_field *const *fields,
                          mcu_selection *sel, MiView &v, cudaStream_t st, const mcu_quad_method *method = nullptr) {
    if (!m || (grade < 1 || grade > 3) || nops < 1 || nops > MI_MAXOPS || !prog || nfields < 0) { mcu_set_error("bad integral arguments"); return MCU_ERR_ARGS; }
    if (!m->conn[grade].present) { mcu_set_error("mesh has no elements of grade %d", grade); return MCU_ERR_ELNOTFOUND; }
    if (sel && (sel->m != m || sel->g != grade)) { mcu_set_error("selection does not belong to this mesh / grade"); return MCU_ERR_ARGS; }
    static bool rules_up = false;
    if (!rules_up) {
        MiRules R;
        make_rules(R);
        if (cudaMemcpyToSymbol(c_rules, &R, sizeof(R)) != cudaSuccess) { mcu_set_error("rule upload failed"); return MCU_ERR_CUDA; }
        rules_up = true;
    }
    memset(&v, 0, sizeof(v));
    if (method) {
        int rm = method_setup(grade, method, st);
        if (rm) return rm;
        if (!g_integral_fail && cudaMalloc((void **) &g_integral_fail, sizeof(unsigned)) != cudaSuccess) return MCU_ERR_ALLOC;
        if (cudaMemsetAsync(g_integral_fail, 0, sizeof(unsigned), st) != cudaSuccess) return MCU_ERR_CUDA;
        v.method = 1; v.fail = g_integral_fail;
    }
    v.dim = m->dim; v.g = grade; v.nops = nops; v.nel = m->conn[grade].nel;
    v.vert = m->d_vert; v.rix = m->conn[grade].d_rix;
    if (sel) { int rs = mcu_ensure_selection(sel); if (rs) return rs; }
    v.eids = sel ? sel->d_eids : nullptr;
    v.n = sel ? (int) sel->ids.size() : v.nel;
    for (int f = 0; f < nfields; f++) {
        if (!fields || !fields[f] || fields[f]->m != m) { mcu_set_error("field does not belong to this mesh"); return MCU_ERR_ARGS; }
        for (int c = 0; c < fields[f]->psize; c++) {
            if (v.nq == MI_MAXQ) { mcu_set_error("too many field components (max %d)", MI_MAXQ); return MCU_ERR_UNSUPPORTED; }
            v.fld[v.nq] = fields[f]->d_data; v.fstride[v.nq] = fields[f]->psize; v.fcomp[v.nq] = c; v.nq++;
        }
    }
    // validate the program (stack discipline, operand ranges) and give the element constants it uses their slots
    v.nq_in = v.nq;
    v.tn_slot = -1;
    std::vector<mcu_expr_op> rw(prog, prog + nops);
    int depth = 0;
    unsigned stored = 0;
    for (int i = 0; i < nops; i++) {
        mcu_expr_op &op = rw[i];
        if (op.op == MCU_X_TANGENT || op.op == MCU_X_NORMAL) {
            // the reference raises TNGNT_GRD / NRML_GRD on the wrong grade; veccross needs three coordinates
            if (op.a < 0 || op.a > 2 || (op.op == MCU_X_TANGENT ? grade != 1 : (grade != 2 || m->dim != 3))) { mcu_set_error("tangent()/normal() not defined for this grade"); return MCU_ERR_UNSUPPORTED; }
            if (v.tn_slot < 0) {
                if (v.nq + 3 > MI_MAXQ) { mcu_set_error("too many integrand quantities"); return MCU_ERR_UNSUPPORTED; }
                v.tn_slot = v.nq; v.nq += 3;
            }
            op.op = MCU_X_FIELD; op.a = v.tn_slot + op.a;
        } else if (op.op == MCU_X_GRADQ) {
            const int comp = op.a / 3, k = op.a % 3;
            if ((grade != 2 && !(grade == 3 && m->dim == 3)) || op.a < 0 || comp >= v.nq_in || k >= m->dim) { mcu_set_error("grad() needs a triangle or a tetrahedron and a field component"); return MCU_ERR_UNSUPPORTED; }
            int gi = -1;
            for (int j = 0; j < v.ngrad; j++) if (v.gsrc[j] == comp) gi = j;
            if (gi < 0) {
                if (v.ngrad == MI_MAXG || v.nq + 3 > MI_MAXQ) { mcu_set_error("too many gradients in the integrand"); return MCU_ERR_UNSUPPORTED; }
                gi = v.ngrad++;
                v.gsrc[gi] = comp; v.gslot[gi] = v.nq; v.nq += 3;
            }
            op.op = MCU_X_FIELD; op.a = v.gslot[gi] + k;
        }
        if (op.op == MCU_X_CONST || op.op == MCU_X_POS || op.op == MCU_X_FIELD) {
            if ((op.op == MCU_X_POS && (op.a < 0 || op.a > 2)) || (op.op == MCU_X_FIELD && (op.a < 0 || op.a >= v.nq))) { mcu_set_error("integrand operand out of range at op %d", i); return MCU_ERR_ARGS; }
            depth++;
        } else if ((op.op >= MCU_X_ADD && op.op <= MCU_X_DIV) || op.op == MCU_X_POW || op.op == MCU_X_ATAN2 || (op.op >= MCU_X_LT && op.op <= MCU_X_NE)) { if (depth < 2) { mcu_set_error("integrand stack underflow at op %d", i); return MCU_ERR_ARGS; } depth--; }
        else if ((op.op >= MCU_X_NEG && op.op <= MCU_X_LOG10) || op.op == MCU_X_NOT) { if (depth < 1) { mcu_set_error("integrand stack underflow at op %d", i); return MCU_ERR_ARGS; } }
        else if (op.op == MCU_X_SELECT) { if (depth < 3) { mcu_set_error("integrand stack underflow at op %d", i); return MCU_ERR_ARGS; } depth -= 2; }
        else if (op.op == MCU_X_STORE) { if (depth < 1 || op.a < 0 || op.a > 31) { mcu_set_error("bad register store at op %d", i); return MCU_ERR_ARGS; } stored |= 1u << op.a; }
        else if (op.op == MCU_X_LOAD) { if (op.a < 0 || op.a > 31 || !(stored & (1u << op.a))) { mcu_set_error("register loaded before it is stored at op %d", i); return MCU_ERR_ARGS; } depth++; }
        else { mcu_set_error("unknown integrand op %d", op.op); return MCU_ERR_ARGS; }
        if (depth > MI_STACK) { mcu_set_error("integrand stack too deep"); return MCU_ERR_UNSUPPORTED; }
    }
    if (depth != 1) { mcu_set_error("integrand program leaves %d values", depth); return MCU_ERR_ARGS; }
    prog = rw.data();                           // lives on this frame: the copy below is completed before returning
    if (cudaMemcpyToSymbolAsync(c_prog, prog, sizeof(mcu_expr_op) * nops, 0, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) return MCU_ERR_CUDA;
    return MCU_OK;
}

// ---- ScalarPotential (functional.c:2255-2372): a function of the position evaluated at every vertex ----
struct PwView {
    int dim, n, nprog;
    const double *vert;
    const int *ids;                            // selected vertices or nullptr
    int off[3], len[3];                        // the programs inside c_prog
};

__global__ void __launch_bounds__(256) k_pw_values(PwView v, double *out) {
    int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= v.n) return;
    const int id = v.ids ? v.ids[c] : c;
    double x[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < v.dim; k++) x[k] = v.vert[(size_t) id * v.dim + k];
    out[id] = mi_run(v.off[0], v.len[0], x, x);
}

// gradient: the gradient function's components if given (scalarpotential_gradient adds the returned Matrix to the
// zeroed column), else central differences of the potential (functional_numericalgrad with the vertex as element)
__global__ void __launch_bounds__(256) k_pw_grad(PwView v, double *out) {
    int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= v.n) return;
    const int id = v.ids ? v.ids[c] : c;
    double x[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < v.dim; k++) x[k] = v.vert[(size_t) id * v.dim + k];
    for (int k = 0; k < v.dim; k++) {
        double g;
        if (v.nprog > 1) g = 0.0 + 1.0 * mi_run(v.off[k], v.len[k], x, x);
        else {
            const double x0 = x[k], eps = fdstepsize(x0);
            x[k] = x0 + eps;
            const double fp = mi_run(v.off[0], v.len[0], x, x);
            x[k] = x0 - eps;
            const double fm = mi_run(v.off[0], v.len[0], x, x);
            x[k] = x0;
            g = 0.0 + (fp - fm) / (2 * eps);
        }
        out[(size_t) id * v.dim + k] = g;
    }
}

static int validate_program(const mcu_expr_op *prog, int nops, int dim, int nq, bool element_ops) {
    int depth = 0;
    unsigned stored = 0;
    for (int i = 0; i < nops; i++) {
        const mcu_expr_op &op = prog[i];
        if (op.op == MCU_X_CONST || op.op == MCU_X_POS || op.op == MCU_X_FIELD) {
            if ((op.op == MCU_X_POS && (op.a < 0 || op.a >= dim)) || (op.op == MCU_X_FIELD && (op.a < 0 || op.a >= nq))) { mcu_set_error("operand out of range at op %d", i); return MCU_ERR_ARGS; }
            depth++;
        } else if ((op.op >= MCU_X_ADD && op.op <= MCU_X_DIV) || op.op == MCU_X_POW || op.op == MCU_X_ATAN2 || (op.op >= MCU_X_LT && op.op <= MCU_X_NE)) { if (depth < 2) { mcu_set_error("stack underflow at op %d", i); return MCU_ERR_ARGS; } depth--; }
        else if ((op.op >= MCU_X_NEG && op.op <= MCU_X_LOG10) || op.op == MCU_X_NOT) { if (depth < 1) { mcu_set_error("stack underflow at op %d", i); return MCU_ERR_ARGS; } }
        else if (op.op == MCU_X_SELECT) { if (depth < 3) { mcu_set_error("stack underflow at op %d", i); return MCU_ERR_ARGS; } depth -= 2; }
        else if (op.op == MCU_X_STORE) { if (depth < 1 || op.a < 0 || op.a > 31) { mcu_set_error("bad register store at op %d", i); return MCU_ERR_ARGS; } stored |= 1u << op.a; }
        else if (op.op == MCU_X_LOAD) { if (op.a < 0 || op.a > 31 || !(stored & (1u << op.a))) { mcu_set_error("register loaded before it is stored at op %d", i); return MCU_ERR_ARGS; } depth++; }
        else if (element_ops && op.op >= MCU_X_TANGENT && op.op <= MCU_X_GRADQ) depth++;
        else { mcu_set_error("op %d not allowed here", op.op); return MCU_ERR_ARGS; }
        if (depth > MI_STACK) { mcu_set_error("expression stack too deep"); return MCU_ERR_UNSUPPORTED; }
    }
    if (depth != 1) { mcu_set_error("program leaves %d values", depth); return MCU_ERR_ARGS; }
    return MCU_OK;
}

extern "C" int mcu_pointwise(mcu_mesh *m, int nprog, const int *nops, const mcu_expr_op *progs, mcu_selection *sel, int mode, double *host_out) {
    if (!m || !nops || !progs || !host_out || nprog < 1 || nprog > 3) { mcu_set_error("bad pointwise arguments"); return MCU_ERR_ARGS; }
    if (mode != MCU_TOTAL && mode != MCU_INTEGRAND && mode != MCU_GRADIENT) { mcu_set_error("bad pointwise mode"); return MCU_ERR_ARGS; }
    if ((mode != MCU_GRADIENT && nprog != 1) || (mode == MCU_GRADIENT && nprog != 1 && nprog != m->dim)) { mcu_set_error("pointwise: wrong number of programs"); return MCU_ERR_ARGS; }
    if (sel && (sel->m != m || sel->g != 0)) { mcu_set_error("selection does not belong to this mesh / grade 0"); return MCU_ERR_ARGS; }
    PwView v;
    memset(&v, 0, sizeof(v));
    int total_ops = 0;
    for (int k = 0; k < nprog; k++) {
        if (nops[k] < 1 || total_ops + nops[k] > MI_MAXOPS) { mcu_set_error("pointwise programs too long"); return MCU_ERR_UNSUPPORTED; }
        int rc = validate_program(progs + total_ops, nops[k], m->dim, 0, false);
        if (rc) return rc;
        v.off[k] = total_ops; v.len[k] = nops[k];
        total_ops += nops[k];
    }
    cudaStream_t st = mcu_stream();
    if (sel) { int rs = mcu_ensure_selection(sel); if (rs) return rs; }
    v.dim = m->dim; v.nprog = nprog; v.vert = m->d_vert;
    v.ids = sel ? sel->d_eids : nullptr;
    v.n = sel ? (int) sel->ids.size() : m->nv;
    if (cudaMemcpyToSymbolAsync(c_prog, progs, sizeof(mcu_expr_op) * total_ops, 0, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) return MCU_ERR_CUDA;
    const size_t nout = mode == MCU_GRADIENT ? (size_t) m->nv * m->dim : (size_t) m->nv;
    double *d_out = nullptr, *d_tot = nullptr;
    CompSum *d_part = nullptr;
    int rc = MCU_OK;
    if (cudaMalloc((void **) &d_out, sizeof(double) * std::max<size_t>(nout, 1)) != cudaSuccess) return MCU_ERR_ALLOC;
    cudaMemsetAsync(d_out, 0, sizeof(double) * std::max<size_t>(nout, 1), st);
    if (v.n > 0) {
        if (mode == MCU_GRADIENT) k_pw_grad<<<(v.n + 255) / 256, 256, 0, st>>>(v, d_out);
        else k_pw_values<<<(v.n + 255) / 256, 256, 0, st>>>(v, d_out);
        g_mcu_launches += 1;
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && mode == MCU_TOTAL) {
        int rb = std::max(1, std::min(1024, (m->nv + 255) / 256));
        if (cudaMalloc((void **) &d_part, sizeof(CompSum) * rb) != cudaSuccess || cudaMalloc((void **) &d_tot, 8) != cudaSuccess) rc = MCU_ERR_ALLOC;
        else {
            k_isum<<<rb, 256, 0, st>>>(d_out, m->nv, d_part);
            k_isum_finish<<<1, 256, 0, st>>>(d_part, rb, d_tot);
            g_mcu_launches += 2;
            e = cudaMemcpyAsync(host_out, d_tot, 8, cudaMemcpyDeviceToHost, st);
        }
    } else if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, d_out, sizeof(double) * nout, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { mcu_set_error("pointwise evaluation failed: %s", cudaGetErrorString(e)); rc = MCU_ERR_CUDA; }
    cudaFree(d_out);
    if (d_part) cudaFree(d_part);
    if (d_tot) cudaFree(d_tot);
    return rc;
}

/* Gradient of the integral by central differences, like the reference (LineIntegral/AreaIntegral gradient and
 * fieldgradient are functional_mapnumericalgradient / functional_mapnumericalfieldgradient over the quadrature,
 * functional.c:5240, 5319-5340, 5434-5458).  wrt < 0: vertex positions, host_out nv * dim doubles; wrt = k: the k-th
 * field, host_out nv * psize_k doubles (the layout of that Field's store). */
// after the kernels of a rule-table evaluation: did any element run out of work items (or hit the reference's
// IntgrtrSbdvsns condition)?  Such a call is handed back to the reference.
static int method_check(const MiView &v, cudaStream_t st) {
    if (!v.method) return MCU_OK;
    unsigned f = 0;
    if (cudaMemcpyAsync(&f, v.fail, sizeof(unsigned), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return MCU_ERR_CUDA;
    if (f) { mcu_set_error("rule-table integrator: an element needs more subdivisions than the device keeps (or cannot be subdivided further)"); return MCU_ERR_UNSUPPORTED; }
    return MCU_OK;
}

extern "C" int mcu_integral_gradient(mcu_mesh *m, int grade, int nops, const mcu_expr_op *prog, int nfields, mcu_field *const *fields,
                                     mcu_selection *sel, int wrt, double *host_out) {
    return mcu_integral_gradient_method(m, grade, nops, prog, nfields, fields, sel, wrt, nullptr, host_out);
}

extern "C" int mcu_integral_gradient_method(mcu_mesh *m, int grade, int nops, const mcu_expr_op *prog, int nfields, mcu_field *const *fields,
                                            mcu_selection *sel, int wrt, const mcu_quad_method *method, double *host_out) {
    if (!host_out || wrt >= nfields) { mcu_set_error("bad integral gradient arguments"); return MCU_ERR_ARGS; }
    cudaStream_t st = mcu_stream();
    MiView v;
    int rc = integral_setup(m, grade, nops, prog, nfields, fields, sel, v, st, method);
    if (rc) return rc;
    int qbase = -1, ncomp = m->dim;
    if (wrt >= 0) {
        qbase = 0;
        for (int f = 0; f < wrt; f++) qbase += fields[f]->psize;
        ncomp = fields[wrt]->psize;
    }
    const int nv = grade + 1;
    const int *tptr, *tidx;
    if (sel) { tptr = sel->d_tptr; tidx = sel->d_tidx; }
    else {
        rc = mcu_ensure_transpose(m, grade);
        if (rc) return rc;
        tptr = m->conn[grade].d_tptr; tidx = m->conn[grade].d_tidx;
    }
    double *d_scratch = nullptr, *d_out = nullptr;
    const size_t ns = (size_t) std::max(v.n, 1) * nv * ncomp, no = (size_t) m->nv * ncomp;
    if (cudaMalloc((void **) &d_scratch, sizeof(double) * ns) != cudaSuccess || cudaMalloc((void **) &d_out, sizeof(double) * std::max<size_t>(no, 1)) != cudaSuccess) {
        if (d_scratch) cudaFree(d_scratch);
        mcu_set_error("integral gradient: allocation failed");
        return MCU_ERR_ALLOC;
    }
    const long nt = (long) v.n * nv;
    if (nt > 0) { k_integral_fd<<<(unsigned) ((nt + 127) / 128), 128, 0, st>>>(v, qbase, ncomp, d_scratch); g_mcu_launches += 1; }
    k_integral_gather<<<(m->nv + 127) / 128, 128, 0, st>>>(m->nv, nv, ncomp, tptr, tidx, v.eids, v.rix, d_scratch, d_out);
    g_mcu_launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, d_out, sizeof(double) * no, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { mcu_set_error("integral gradient failed: %s", cudaGetErrorString(e)); rc = MCU_ERR_CUDA; }
    if (rc == MCU_OK) rc = method_check(v, st);
    cudaFree(d_scratch);
    cudaFree(d_out);
    return rc;
}

extern "C" int mcu_integral(mcu_mesh *m, int grade, int nops, const mcu_expr_op *prog, int nfields, mcu_field *const *fields,
                            mcu_selection *sel, int mode, double *host_out) {
    return mcu_integral_method(m, grade, nops, prog, nfields, fields, sel, mode, nullptr, host_out);
}

extern "C" int mcu_integral_method(mcu_mesh *m, int grade, int nops, const mcu_expr_op *prog, int nfields, mcu_field *const *fields,
                                   mcu_selection *sel, int mode, const mcu_quad_method *method, double *host_out) {
    if (!host_out || (mode != MCU_TOTAL && mode != MCU_INTEGRAND)) { mcu_set_error("bad integral arguments"); return MCU_ERR_ARGS; }
    cudaStream_t st = mcu_stream();
    MiView v;
    int rc0 = integral_setup(m, grade, nops, prog, nfields, fields, sel, v, st, method);
    if (rc0) return rc0;
    double *d_int = nullptr, *d_tot = nullptr;
    CompSum *d_part = nullptr;
    int rc = MCU_OK;
    const int nel = v.nel;
    if (cudaMalloc((void **) &d_int, sizeof(double) * std::max(nel, 1)) != cudaSuccess) return MCU_ERR_ALLOC;
    cudaMemsetAsync(d_int, 0, sizeof(double) * std::max(nel, 1), st);
    if (v.n > 0) { k_integral<<<(v.n + 127) / 128, 128, 0, st>>>(v, d_int); g_mcu_launches += 1; }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && mode == MCU_TOTAL) {
        int rb = std::max(1, std::min(1024, (nel + 255) / 256));
        if (cudaMalloc((void **) &d_part, sizeof(CompSum) * rb) != cudaSuccess || cudaMalloc((void **) &d_tot, 8) != cudaSuccess) rc = MCU_ERR_ALLOC;
        else {
            k_isum<<<rb, 256, 0, st>>>(d_int, nel, d_part);
            k_isum_finish<<<1, 256, 0, st>>>(d_part, rb, d_tot);
            g_mcu_launches += 2;
            e = cudaMemcpyAsync(host_out, d_tot, 8, cudaMemcpyDeviceToHost, st);
        }
    } else if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, d_int, sizeof(double) * nel, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { mcu_set_error("integral failed: %s", cudaGetErrorString(e)); rc = MCU_ERR_CUDA; }
    if (rc == MCU_OK) rc = method_check(v, st);
    cudaFree(d_int);
    if (d_part) cudaFree(d_part);
    if (d_tot) cudaFree(d_tot);
    return rc;
}

// mcu_kernels_impl.cuh — generic map kernels (first, always-available path).
//
//   k_total      : Σ_e integrand(e)            functional_sumintegrand   (functional.c:962-996)
//   k_integrand  : integrand(e) per element    functional_mapintegrand   (functional.c:1049-1085)
//   k_grad_gather: analytic gradients, CSR-transpose vertex gather
//                                              functional_mapgradient    (functional.c:1092-1135)
//   k_fd_grad    : central-difference gradient functional_mapnumericalgradient (1193-1247)
//   k_fd_fieldgrad: central-difference field gradient
//                                              functional_mapnumericalfieldgradient (1428-1503)
//
// The gradient kernels run one thread per VERTEX and walk that vertex's incident elements
// in ascending element id (the transpose is a stable counting sort), so every output
// entry is produced by exactly one thread with the same summation order as the
// reference's serial loop: deterministic, no atomics.
//
// Included by two translation units: mcu_kernels_fast.cu (analytic functionals, FMA on)
// and mcu_kernels_strict.cu (-fmad=false, MCU_STRICT: finite-difference family).
#pragma once

#include "mcu_integrands.cuh"
#include "mcu_kernels.h"

#define MCU_BLOCK 256

// Shape code of a kernel instantiation: 0 = generic (vertices per element, dimension and field width read from the
// parameters at run time); otherwise nv | ps << 3 with dim = 3 — literal constants, so that every loop over them
// unrolls and the gathered element data lives in registers instead of a local-memory stack frame.
#define MCU_SHAPE(nv, ps) ((nv) | ((ps) << 3))
template <int SH> struct Shape {
    static __device__ __forceinline__ int nv(const McuParams &p) { return SH ? (SH & 7) : p.nvper; }
    static __device__ __forceinline__ int dim(const McuParams &p) { return SH ? 3 : p.dim; }
    static __device__ __forceinline__ int ps(const McuParams &p) { return SH ? (SH >> 3) : p.psize0; }
};
// element vertex `corner` (run-time) / component k, without run-time indexing of the register arrays
__device__ __forceinline__ double elem_getx(const ElemData &d, int corner, int k) {
    double r = 0.0;
#pragma unroll
    for (int j = 0; j < MCU_MAXNV; j++) if (j == corner) r = d.x[j][k];
    return r;
}
__device__ __forceinline__ void elem_setx(ElemData &d, int corner, int k, double val) {
#pragma unroll
    for (int j = 0; j < MCU_MAXNV; j++) if (j == corner) d.x[j][k] = val;
}
__device__ __forceinline__ double elem_getq(const ElemData &d, int corner, int i) {
    double r = 0.0;
#pragma unroll
    for (int j = 0; j < MCU_MAXNV; j++) if (j == corner) r = d.q[j][i];
    return r;
}
__device__ __forceinline__ void elem_setq(ElemData &d, int corner, int i, double val) {
#pragma unroll
    for (int j = 0; j < MCU_MAXNV; j++) if (j == corner) d.q[j][i] = val;
}
__device__ __forceinline__ double elem_getphi(const ElemData &d, int corner) {
    double r = 0.0;
#pragma unroll
    for (int j = 0; j < MCU_MAXNV; j++) if (j == corner) r = d.phi[j];
    return r;
}
__device__ __forceinline__ void elem_setphi(ElemData &d, int corner, double val) {
#pragma unroll
    for (int j = 0; j < MCU_MAXNV; j++) if (j == corner) d.phi[j] = val;
}

template <int F, int SH>
__global__ void __launch_bounds__(MCU_BLOCK) k_total(McuParams p, McuView v, CompSum *partials) {
    CompSum acc = {0.0, 0.0};
    const int nv = Shape<SH>::nv(p), dim = Shape<SH>::dim(p), ps = Shape<SH>::ps(p);
    const int stride = gridDim.x * MCU_BLOCK;
    // unrolled: the gathers of several elements are in flight per thread (a single dependent id -> coordinate chain
    // per iteration leaves the memory system idle)
#pragma unroll 4
    for (int ce = blockIdx.x * MCU_BLOCK + threadIdx.x; ce < v.n; ce += stride) {
        int e = v.eids ? v.eids[ce] : ce;
        int ids[MCU_MAXNV];
        if (p.g == 0) ids[0] = e;
        else {
#pragma unroll
            for (int j = 0; j < MCU_MAXNV; j++) if (j < nv) ids[j] = v.rix[(size_t) e * nv + j];
        }
        ElemData d;
        gather_elem<F>(p, v, ids, d, nv, dim, ps);
        comp_add(acc, eval_integrand<F>(p, d, nv, dim, ps));
    }
    CompSum r = block_reduce<MCU_BLOCK>(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

// Single-block, fixed-order reduction of the per-block partials.
static __global__ void __launch_bounds__(MCU_BLOCK) k_total_finish(const CompSum *partials, int n, double *out) {
    CompSum acc = {0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += MCU_BLOCK) comp_merge(acc, partials[i]);
    CompSum r = block_reduce<MCU_BLOCK>(acc);
    if (threadIdx.x == 0) out[0] = r.s + r.c;
}

template <int F, int SH>
__global__ void __launch_bounds__(MCU_BLOCK) k_integrand(McuParams p, McuView v, double *out) {
    int ce = blockIdx.x * MCU_BLOCK + threadIdx.x;
    if (ce >= v.n) return;
    const int nv = Shape<SH>::nv(p), dim = Shape<SH>::dim(p), ps = Shape<SH>::ps(p);
    int e = v.eids ? v.eids[ce] : ce;
    int ids[MCU_MAXNV];
    if (p.g == 0) ids[0] = e;
    else {
#pragma unroll
        for (int j = 0; j < MCU_MAXNV; j++) if (j < nv) ids[j] = v.rix[(size_t) e * nv + j];
    }
    ElemData d;
    gather_elem<F>(p, v, ids, d, nv, dim, ps);
    out[e] = eval_integrand<F>(p, d, nv, dim, ps);
}

template <int F>
__device__ __forceinline__ bool analytic_grad(const double (*x)[3], double (*g)[3], unsigned *flags) {
    if (F == MCU_F_LENGTH) { if (!length_grad(x, g)) { *flags |= MCU_FLAG_DEGENERATE; return false; } }
    if (F == MCU_F_AREA) { if (!area_grad(x, g)) { *flags |= MCU_FLAG_DEGENERATE; return false; } }
    if (F == MCU_F_VOLUMEENCLOSED) { if (volumeenclosed_grad(x, g)) { *flags |= MCU_FLAG_VOLENCLZERO; return false; } }
    if (F == MCU_F_VOLUME) volume_grad(x, g);
    return true;
}

template <int F>
__global__ void __launch_bounds__(MCU_BLOCK) k_grad_gather(McuParams p, McuView v, double *out, unsigned *flags) {
    int vtx = blockIdx.x * MCU_BLOCK + threadIdx.x;
    if (vtx >= p.nv) return;
    double acc[3] = {0.0, 0.0, 0.0};
    unsigned myflags = 0;
    for (int q = v.tptr[vtx]; q < v.tptr[vtx + 1]; q++) {
        int ce = v.tidx[q];
        int e = v.eids ? v.eids[ce] : ce;
        // fully unrolled with constant indices (guards instead of run-time trip counts and a select instead of
        // g[corner]): the small arrays stay in registers — indexed dynamically they live in local memory
        int ids[MCU_MAXNV];
        double x[MCU_MAXNV][3], g[MCU_MAXNV][3];
#pragma unroll
        for (int j = 0; j < MCU_MAXNV; j++) {
            ids[j] = (j < p.nvper) ? v.rix[(size_t) e * p.nvper + j] : -1;
            const double *xv = v.vert + (size_t) (ids[j] < 0 ? 0 : ids[j]) * p.dim;
            x[j][0] = (j < p.nvper) ? xv[0] : 0.0;
            x[j][1] = (j < p.nvper) ? xv[1] : 0.0;
            x[j][2] = (j < p.nvper && p.dim == 3) ? xv[2] : 0.0;
        }
        if (analytic_grad<F>(x, g, &myflags)) {
#pragma unroll
            for (int j = 0; j < MCU_MAXNV; j++) {
                const bool mine = ids[j] == vtx;
#pragma unroll
                for (int k = 0; k < 3; k++) acc[k] += mine ? g[j][k] : 0.0;
            }
        }
    }
    for (int k = 0; k < p.dim; k++) out[(size_t) vtx * p.dim + k] = acc[k];
    if (myflags) atomicOr(flags, myflags);  // status word only; never data
}

// The finite-difference kernels evaluate heavy integrands (Nematic: ~700 instructions per evaluation) with long FP64
// dependency chains; unconstrained they take ~150 registers = one 256-thread block per SM and stall on latency.
// Two resident blocks (<= 128 registers) hide it better.
#ifndef MCU_FD_MINBLOCKS
#define MCU_FD_MINBLOCKS 2
#endif

// Central differences with respect to the coordinates of one vertex, summed over the
// incident elements in ascending id: frc = f0 + (fp - fm)/(2 eps) per coordinate
// (functional_numericalgrad, functional.c:1142-1162).
template <int F, int SH>
__global__ void __launch_bounds__(MCU_BLOCK, MCU_FD_MINBLOCKS) k_fd_grad(McuParams p, McuView v, double *out) {
    int vtx = blockIdx.x * MCU_BLOCK + threadIdx.x;
    if (vtx >= p.nv) return;
    const int nv = Shape<SH>::nv(p), dim = Shape<SH>::dim(p), ps = Shape<SH>::ps(p);
    double acc[3] = {0.0, 0.0, 0.0};
    if (p.g == 0) {
        // vertex elements whose integrand does not depend on positions (NormSq): FD gives exact zeros
    } else {
        for (int q = v.tptr[vtx]; q < v.tptr[vtx + 1]; q++) {
            int ce = v.tidx[q];
            int e = v.eids ? v.eids[ce] : ce;
            int ids[MCU_MAXNV];
            int corner = 0;
#pragma unroll
            for (int j = 0; j < MCU_MAXNV; j++) {
                if (j >= nv) continue;
                ids[j] = v.rix[(size_t) e * nv + j];
                if (ids[j] == vtx) corner = j;
            }
            ElemData d;
            gather_elem<F>(p, v, ids, d, nv, dim, ps);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (k >= dim) continue;
                double x0 = elem_getx(d, corner, k);
                double eps = fdstepsize(x0);
                elem_setx(d, corner, k, x0 + eps);
                double fp = eval_integrand<F>(p, d, nv, dim, ps);
                elem_setx(d, corner, k, x0 - eps);
                double fm = eval_integrand<F>(p, d, nv, dim, ps);
                elem_setx(d, corner, k, x0);
                acc[k] = acc[k] + (fp - fm) / (2 * eps);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) if (k < dim) out[(size_t) vtx * dim + k] = acc[k];
}

// Central differences with respect to the field degrees of freedom stored at one vertex
// (functional_numericalfieldgrad, functional.c:1305-1328).
template <int F, int SH>
__global__ void __launch_bounds__(MCU_BLOCK, MCU_FD_MINBLOCKS) k_fd_fieldgrad(McuParams p, McuView v, double *out) {
    int vtx = blockIdx.x * MCU_BLOCK + threadIdx.x;
    if (vtx >= p.nv) return;
    const int nv = Shape<SH>::nv(p), dim = Shape<SH>::dim(p), ps0 = Shape<SH>::ps(p);
    const int ps = (p.wrt == 0 ? ps0 : p.psize1);
    double acc[MCU_MAXPSIZE];
#pragma unroll
    for (int i = 0; i < MCU_MAXPSIZE; i++) acc[i] = 0.0;
    if (p.g == 0) {
        // grade-0 functional (NormSq): the element is the vertex itself
        bool selected = true;
        if (v.eids) selected = (v.tptr[vtx + 1] > v.tptr[vtx]);
        if (selected) {
            int ids[1] = {vtx};
            ElemData d;
            gather_elem<F>(p, v, ids, d, 1, dim, ps0);
            for (int i = 0; i < ps; i++) {
                double f0 = d.q[0][i];
                double eps = fdstepsize(f0);
                d.q[0][i] = f0 + eps;
                double fr = eval_integrand<F>(p, d, 1, dim, ps0);
                d.q[0][i] = f0 - eps;
                double fl = eval_integrand<F>(p, d, 1, dim, ps0);
                d.q[0][i] = f0;
                acc[i] += (fr - fl) / (2 * eps);
            }
        }
    } else {
        for (int q = v.tptr[vtx]; q < v.tptr[vtx + 1]; q++) {
            int ce = v.tidx[q];
            int e = v.eids ? v.eids[ce] : ce;
            int ids[MCU_MAXNV];
            int corner = 0;
#pragma unroll
            for (int j = 0; j < MCU_MAXNV; j++) {
                if (j >= nv) continue;
                ids[j] = v.rix[(size_t) e * nv + j];
                if (ids[j] == vtx) corner = j;
            }
            ElemData d;
            gather_elem<F>(p, v, ids, d, nv, dim, ps0);
            if (SH) {
                // shape known at compile time: constant indices everywhere
#pragma unroll
                for (int i = 0; i < MCU_MAXPSIZE; i++) {
                    if (i >= ps) continue;
                    double f0 = (p.wrt == 0) ? elem_getq(d, corner, i) : elem_getphi(d, corner);
                    double eps = fdstepsize(f0);
                    if (p.wrt == 0) elem_setq(d, corner, i, f0 + eps); else elem_setphi(d, corner, f0 + eps);
                    double fr = eval_integrand<F>(p, d, nv, dim, ps0);
                    if (p.wrt == 0) elem_setq(d, corner, i, f0 - eps); else elem_setphi(d, corner, f0 - eps);
                    double fl = eval_integrand<F>(p, d, nv, dim, ps0);
                    if (p.wrt == 0) elem_setq(d, corner, i, f0); else elem_setphi(d, corner, f0);
                    acc[i] += (fr - fl) / (2 * eps);
                }
            } else {
                for (int i = 0; i < ps; i++) {
                    double *slot = (p.wrt == 0 ? &d.q[corner][i] : &d.phi[corner]);
                    double f0 = *slot;
                    double eps = fdstepsize(f0);
                    *slot = f0 + eps;
                    double fr = eval_integrand<F>(p, d, nv, dim, ps0);
                    *slot = f0 - eps;
                    double fl = eval_integrand<F>(p, d, nv, dim, ps0);
                    *slot = f0;
                    acc[i] += (fr - fl) / (2 * eps);
                }
            }
        }
    }
    if (SH) {
#pragma unroll
        for (int i = 0; i < MCU_MAXPSIZE; i++) if (i < ps) out[(size_t) vtx * ps + i] = acc[i];
    } else {
        for (int i = 0; i < ps; i++) out[(size_t) vtx * ps + i] = acc[i];
    }
}

// Element-centric form of the two finite-difference maps (p.g > 0).  The vertex-centric kernels above gather and
// re-evaluate every element once per (vertex, coordinate) from scattered threads; here one thread owns one
// (element, corner): the nv threads of an element are neighbours in a warp and read the same rows, the central
// differences go to scratch[(ce * nv + corner) * nc + c], and k_fd_sum adds them per vertex in ascending element id
// — the order in which functional_numericalgrad / functional_numericalfieldgrad accumulate (functional.c:1142-1162,
// 1305-1328), so the result is bit-identical to the vertex-centric kernels'.
template <int F, int SH, bool FIELD>
__global__ void __launch_bounds__(MCU_BLOCK, MCU_FD_MINBLOCKS) k_fd_elem(McuParams p, McuView v, double *scratch) {
    const int nv = Shape<SH>::nv(p), dim = Shape<SH>::dim(p), ps0 = Shape<SH>::ps(p);
    const int nc = FIELD ? (p.wrt == 0 ? ps0 : p.psize1) : dim;
    const long t = (long) blockIdx.x * MCU_BLOCK + threadIdx.x;
    if (t >= (long) v.n * nv) return;
    const int ce = (int) (t / nv), corner = (int) (t - (long) ce * nv);
    const int e = v.eids ? v.eids[ce] : ce;
    int ids[MCU_MAXNV];
#pragma unroll
    for (int j = 0; j < MCU_MAXNV; j++) if (j < nv) ids[j] = v.rix[(size_t) e * nv + j];
    ElemData d;
    gather_elem<F>(p, v, ids, d, nv, dim, ps0);
    double *o = scratch + (size_t) t * nc;
    // field perturbations leave the geometry alone: element size and P1-gradient factors are computed once
    ElemGeom gm;
    const ElemGeom *gmp = nullptr;
    if (FIELD && (F == MCU_F_GRADSQ || F == MCU_F_NEMATIC || F == MCU_F_NEMATICELECTRIC) && nv >= 3) { elem_geom(nv - 1, dim, d.x, gm); gmp = &gm; }
    if (!FIELD) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (k >= dim) continue;
            double x0 = elem_getx(d, corner, k);
            double eps = fdstepsize(x0);
            elem_setx(d, corner, k, x0 + eps);
            double fp = eval_integrand<F>(p, d, nv, dim, ps0);
            elem_setx(d, corner, k, x0 - eps);
            double fm = eval_integrand<F>(p, d, nv, dim, ps0);
            elem_setx(d, corner, k, x0);
            o[k] = (fp - fm) / (2 * eps);
        }
    } else if (SH) {
#pragma unroll
        for (int i = 0; i < MCU_MAXPSIZE; i++) {
            if (i >= nc) continue;
            double f0 = (p.wrt == 0) ? elem_getq(d, corner, i) : elem_getphi(d, corner);
            double eps = fdstepsize(f0);
            if (p.wrt == 0) elem_setq(d, corner, i, f0 + eps); else elem_setphi(d, corner, f0 + eps);
            double fr = eval_integrand<F>(p, d, nv, dim, ps0, gmp);
            if (p.wrt == 0) elem_setq(d, corner, i, f0 - eps); else elem_setphi(d, corner, f0 - eps);
            double fl = eval_integrand<F>(p, d, nv, dim, ps0, gmp);
            if (p.wrt == 0) elem_setq(d, corner, i, f0); else elem_setphi(d, corner, f0);
            o[i] = (fr - fl) / (2 * eps);
        }
    } else {
        for (int i = 0; i < nc; i++) {
            double *slot = (p.wrt == 0 ? &d.q[corner][i] : &d.phi[corner]);
            double f0 = *slot;
            double eps = fdstepsize(f0);
            *slot = f0 + eps;
            double fr = eval_integrand<F>(p, d, nv, dim, ps0, gmp);
            *slot = f0 - eps;
            double fl = eval_integrand<F>(p, d, nv, dim, ps0, gmp);
            *slot = f0;
            o[i] = (fr - fl) / (2 * eps);
        }
    }
}

// GradSq.fieldgradient, specialised: a perturbation of one field value changes only that component's gradient, and
// the geometry (LU factors of the tetrahedron / perpendiculars of the triangle, element size) not at all.  The
// arithmetic of every number that IS recomputed is the one gradsq_integrand performs (same helper functions, same
// order), so the differences are bit-identical to re-evaluating the whole integrand twice per degree of freedom.
template <int NV, int PS>
__global__ void __launch_bounds__(MCU_BLOCK, MCU_FD_MINBLOCKS) k_gradsq_fd_elem(McuParams p, McuView v, double *scratch) {
    const long t = (long) blockIdx.x * MCU_BLOCK + threadIdx.x;
    if (t >= (long) v.n * NV) return;
    const int ce = (int) (t / NV), corner = (int) (t - (long) ce * NV);
    const int e = v.eids ? v.eids[ce] : ce;
    double x[MCU_MAXNV][3], q[NV][PS];
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int id = v.rix[(size_t) e * NV + j];
#pragma unroll
        for (int k = 0; k < 3; k++) x[j][k] = v.vert[(size_t) id * 3 + k];
#pragma unroll
        for (int i = 0; i < PS; i++) q[j][i] = v.f0[(size_t) id * PS + i];
    }
    const double size = elem_size(NV - 1, x);
    TetLU lu;
    double tt[3][3];
    if (NV == 4) p1grad_tet_factor(x, lu); else p1grad_tri_factor(3, x, tt);
    double grad[PS * 3];
#pragma unroll
    for (int i = 0; i < PS; i++) {
        if (NV == 4) p1grad_tet_solve(lu, q[0][i], q[1][i], q[2][i], q[3 % NV][i], grad + 3 * i);
        else p1grad_tri_apply(3, tt, q[0][i], q[1][i], q[2][i], grad + 3 * i);
    }
    double *o = scratch + (size_t) t * PS;
#pragma unroll
    for (int i = 0; i < PS; i++) {
        double f0 = 0.0;
#pragma unroll
        for (int j = 0; j < NV; j++) if (j == corner) f0 = q[j][i];
        const double eps = fdstepsize(f0);
        double f[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const double val = s == 0 ? f0 + eps : f0 - eps;
            double w[4];
#pragma unroll
            for (int j = 0; j < 4; j++) w[j] = (j < NV) ? (j == corner ? val : q[j % NV][i]) : 0.0;
            double g3[3], tmp[PS * 3];
            if (NV == 4) p1grad_tet_solve(lu, w[0], w[1], w[2], w[3], g3); else p1grad_tri_apply(3, tt, w[0], w[1], w[2], g3);
#pragma unroll
            for (int c = 0; c < PS * 3; c++) tmp[c] = (c / 3 == i) ? g3[c % 3] : grad[c];
            const double nrm = vnorm(PS * 3, tmp);
            f[s] = nrm * nrm * size;
        }
        o[i] = (f[0] - f[1]) / (2 * eps);
    }
}

static __global__ void __launch_bounds__(MCU_BLOCK) k_fd_sum(int nvtx, int nv, int nc, McuView v, const double *__restrict__ scratch, double *__restrict__ out) {
    int vtx = blockIdx.x * MCU_BLOCK + threadIdx.x;
    if (vtx >= nvtx) return;
    double acc[MCU_MAXPSIZE];
#pragma unroll
    for (int c = 0; c < MCU_MAXPSIZE; c++) acc[c] = 0.0;
    for (int q = v.tptr[vtx]; q < v.tptr[vtx + 1]; q++) {
        const int ce = v.tidx[q], e = v.eids ? v.eids[ce] : ce;
        int corner = 0;
        for (int j = 0; j < nv; j++) if (v.rix[(size_t) e * nv + j] == vtx) corner = j;
        const double *src = scratch + ((size_t) ce * nv + corner) * nc;
#pragma unroll
        for (int c = 0; c < MCU_MAXPSIZE; c++) if (c < nc) acc[c] = acc[c] + src[c];
    }
#pragma unroll
    for (int c = 0; c < MCU_MAXPSIZE; c++) if (c < nc) out[(size_t) vtx * nc + c] = acc[c];
}

// ---------------------------------------------------------------------------------
// host-side launch helpers shared by both translation units
// ---------------------------------------------------------------------------------
static inline int mcu_total_grid(int n) {
    int blocks = (n + MCU_BLOCK - 1) / MCU_BLOCK;
    int cap = 148 * 8;  // 8 resident 256-thread CTAs per SM on 148 SMs
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return blocks;
}

extern long g_mcu_launches;

// Shape of the call, if one of the specialised instantiations fits (field functionals on 3-D triangles /
// tetrahedra with the usual field widths); 0 = generic kernel.
template <int F>
static inline int mcu_shape_of(const McuParams &p) {
    if (p.dim != 3 || (p.nvper != 3 && p.nvper != 4)) return 0;
    if (F == MCU_F_NEMATIC || F == MCU_F_NEMATICELECTRIC) return p.psize0 == 3 ? MCU_SHAPE(p.nvper, 3) : 0;
    if (F == MCU_F_GRADSQ) return (p.psize0 == 1 || p.psize0 == 3 || p.psize0 == 5) ? MCU_SHAPE(p.nvper, p.psize0) : 0;
    return 0;
}
#define MCU_SHAPE_SWITCH(F, sh, LAUNCH) \
    do { \
        if constexpr (F == MCU_F_NEMATIC || F == MCU_F_NEMATICELECTRIC || F == MCU_F_GRADSQ) { \
            if (sh == MCU_SHAPE(3, 3)) { LAUNCH(MCU_SHAPE(3, 3)); break; } \
            if (sh == MCU_SHAPE(4, 3)) { LAUNCH(MCU_SHAPE(4, 3)); break; } \
        } \
        if constexpr (F == MCU_F_GRADSQ) { \
            if (sh == MCU_SHAPE(3, 1)) { LAUNCH(MCU_SHAPE(3, 1)); break; } \
            if (sh == MCU_SHAPE(4, 1)) { LAUNCH(MCU_SHAPE(4, 1)); break; } \
            if (sh == MCU_SHAPE(3, 5)) { LAUNCH(MCU_SHAPE(3, 5)); break; } \
            if (sh == MCU_SHAPE(4, 5)) { LAUNCH(MCU_SHAPE(4, 5)); break; } \
        } \
        LAUNCH(0); \
    } while (0)

template <int F>
static cudaError_t run_total(const McuParams &p, const McuView &v, CompSum *partials, double *d_out, cudaStream_t st) {
    int blocks = mcu_total_grid(v.n);
    const int sh = mcu_shape_of<F>(p);
#define L_(S) k_total<F, S><<<blocks, MCU_BLOCK, 0, st>>>(p, v, partials)
    MCU_SHAPE_SWITCH(F, sh, L_);
#undef L_
    k_total_finish<<<1, MCU_BLOCK, 0, st>>>(partials, blocks, d_out);
    g_mcu_launches += 2;
    return cudaGetLastError();
}

template <int F>
static cudaError_t run_integrand(const McuParams &p, const McuView &v, double *d_out, cudaStream_t st) {
    if (v.n > 0) {
        const int sh = mcu_shape_of<F>(p);
#define L_(S) k_integrand<F, S><<<(v.n + MCU_BLOCK - 1) / MCU_BLOCK, MCU_BLOCK, 0, st>>>(p, v, d_out)
        MCU_SHAPE_SWITCH(F, sh, L_);
#undef L_
        g_mcu_launches += 1;
    }
    return cudaGetLastError();
}

template <int F>
static cudaError_t run_grad_gather(const McuParams &p, const McuView &v, double *d_out, unsigned *flags, cudaStream_t st) {
    k_grad_gather<F><<<(p.nv + MCU_BLOCK - 1) / MCU_BLOCK, MCU_BLOCK, 0, st>>>(p, v, d_out, flags);
    g_mcu_launches += 1;
    return cudaGetLastError();
}

double *mcu_scratch(size_t bytes);      // grow-only device buffer (mcu_api.cu); nullptr if it cannot be had
long mcu_opt(const char *key, long dflt);

// element-centric when the scratch fits (option fd_elementwise, default 1), else the vertex-centric kernels
template <int F, bool FIELD>
static bool run_fd_elementwise(const McuParams &p, const McuView &v, double *d_out, cudaStream_t st) {
    // measured (scripts/field_bench.py, scripts/tet_bench.py): the element-centric form wins for field gradients
    // and loses a few per cent for position gradients
    const long choice = mcu_opt("fd_elementwise", -1);          // -1: the default per kind, 0 / 1: forced
    if (p.g <= 0 || !(choice < 0 ? FIELD : choice != 0)) return false;
    const int nc = FIELD ? (p.wrt == 0 ? p.psize0 : p.psize1) : p.dim;
    if (nc > MCU_MAXPSIZE) return false;
    const long nt = (long) v.n * p.nvper;
    double *scratch = mcu_scratch(sizeof(double) * (size_t) std::max(nt, 1L) * nc);
    if (!scratch) return false;
    const int sh = mcu_shape_of<F>(p);
    bool done = false;
    if constexpr (F == MCU_F_GRADSQ && FIELD) {
        if (nt > 0 && p.wrt == 0 && p.dim == 3 && mcu_opt("gradsq_fast_fieldgradient", 1)) {
            const unsigned gb = (unsigned) ((nt + MCU_BLOCK - 1) / MCU_BLOCK);
#define G_(NV, PS) if (!done && p.nvper == NV && p.psize0 == PS) { k_gradsq_fd_elem<NV, PS><<<gb, MCU_BLOCK, 0, st>>>(p, v, scratch); done = true; }
            G_(3, 1) G_(3, 3) G_(3, 5) G_(4, 1) G_(4, 3) G_(4, 5)
#undef G_
            if (done) g_mcu_launches += 1;
        }
    }
    if (nt > 0 && !done) {
#define L_(S) k_fd_elem<F, S, FIELD><<<(unsigned) ((nt + MCU_BLOCK - 1) / MCU_BLOCK), MCU_BLOCK, 0, st>>>(p, v, scratch)
        MCU_SHAPE_SWITCH(F, sh, L_);
#undef L_
        g_mcu_launches += 1;
    }
    k_fd_sum<<<(p.nv + MCU_BLOCK - 1) / MCU_BLOCK, MCU_BLOCK, 0, st>>>(p.nv, p.nvper, nc, v, scratch, d_out);
    g_mcu_launches += 1;
    return true;
}

template <int F>
static cudaError_t run_fd_grad(const McuParams &p, const McuView &v, double *d_out, cudaStream_t st) {
    if (run_fd_elementwise<F, false>(p, v, d_out, st)) return cudaGetLastError();
    const int sh = mcu_shape_of<F>(p);
#define L_(S) k_fd_grad<F, S><<<(p.nv + MCU_BLOCK - 1) / MCU_BLOCK, MCU_BLOCK, 0, st>>>(p, v, d_out)
    MCU_SHAPE_SWITCH(F, sh, L_);
#undef L_
    g_mcu_launches += 1;
    return cudaGetLastError();
}

template <int F>
static cudaError_t run_fd_fieldgrad(const McuParams &p, const McuView &v, double *d_out, cudaStream_t st) {
    if (run_fd_elementwise<F, true>(p, v, d_out, st)) return cudaGetLastError();
    const int sh = mcu_shape_of<F>(p);
#define L_(S) k_fd_fieldgrad<F, S><<<(p.nv + MCU_BLOCK - 1) / MCU_BLOCK, MCU_BLOCK, 0, st>>>(p, v, d_out)
    MCU_SHAPE_SWITCH(F, sh, L_);
#undef L_
    g_mcu_launches += 1;
    return cudaGetLastError();
}

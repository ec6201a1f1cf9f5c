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

template <int F>
__global__ void __launch_bounds__(MCU_BLOCK) k_total(McuParams p, McuView v, CompSum *partials) {
    CompSum acc = {0.0, 0.0};
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
            for (int j = 0; j < MCU_MAXNV; j++) if (j < p.nvper) ids[j] = v.rix[(size_t) e * p.nvper + j];
        }
        ElemData d;
        gather_elem<F>(p, v, ids, d);
        comp_add(acc, eval_integrand<F>(p, d));
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

template <int F>
__global__ void __launch_bounds__(MCU_BLOCK) k_integrand(McuParams p, McuView v, double *out) {
    int ce = blockIdx.x * MCU_BLOCK + threadIdx.x;
    if (ce >= v.n) return;
    int e = v.eids ? v.eids[ce] : ce;
    int ids[MCU_MAXNV];
    if (p.g == 0) ids[0] = e;
    else for (int j = 0; j < p.nvper; j++) ids[j] = v.rix[(size_t) e * p.nvper + j];
    ElemData d;
    gather_elem<F>(p, v, ids, d);
    out[e] = eval_integrand<F>(p, d);
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

// Central differences with respect to the coordinates of one vertex, summed over the
// incident elements in ascending id: frc = f0 + (fp - fm)/(2 eps) per coordinate
// (functional_numericalgrad, functional.c:1142-1162).
template <int F>
__global__ void __launch_bounds__(MCU_BLOCK) k_fd_grad(McuParams p, McuView v, double *out) {
    int vtx = blockIdx.x * MCU_BLOCK + threadIdx.x;
    if (vtx >= p.nv) return;
    double acc[3] = {0.0, 0.0, 0.0};
    if (p.g == 0) {
        // vertex elements whose integrand does not depend on positions (NormSq): FD gives exact zeros
    } else {
        for (int q = v.tptr[vtx]; q < v.tptr[vtx + 1]; q++) {
            int ce = v.tidx[q];
            int e = v.eids ? v.eids[ce] : ce;
            int ids[MCU_MAXNV];
            int corner = 0;
            for (int j = 0; j < p.nvper; j++) {
                ids[j] = v.rix[(size_t) e * p.nvper + j];
                if (ids[j] == vtx) corner = j;
            }
            ElemData d;
            gather_elem<F>(p, v, ids, d);
            for (int k = 0; k < p.dim; k++) {
                double x0 = d.x[corner][k];
                double eps = fdstepsize(x0);
                d.x[corner][k] = x0 + eps;
                double fp = eval_integrand<F>(p, d);
                d.x[corner][k] = x0 - eps;
                double fm = eval_integrand<F>(p, d);
                d.x[corner][k] = x0;
                acc[k] = acc[k] + (fp - fm) / (2 * eps);
            }
        }
    }
    for (int k = 0; k < p.dim; k++) out[(size_t) vtx * p.dim + k] = acc[k];
}

// Central differences with respect to the field degrees of freedom stored at one vertex
// (functional_numericalfieldgrad, functional.c:1305-1328).
template <int F>
__global__ void __launch_bounds__(MCU_BLOCK) k_fd_fieldgrad(McuParams p, McuView v, double *out) {
    int vtx = blockIdx.x * MCU_BLOCK + threadIdx.x;
    if (vtx >= p.nv) return;
    const int ps = (p.wrt == 0 ? p.psize0 : p.psize1);
    double acc[MCU_MAXPSIZE];
    for (int i = 0; i < ps; i++) acc[i] = 0.0;
    if (p.g == 0) {
        // grade-0 functional (NormSq): the element is the vertex itself
        bool selected = true;
        if (v.eids) selected = (v.tptr[vtx + 1] > v.tptr[vtx]);
        if (selected) {
            int ids[1] = {vtx};
            ElemData d;
            gather_elem<F>(p, v, ids, d);
            for (int i = 0; i < ps; i++) {
                double f0 = d.q[0][i];
                double eps = fdstepsize(f0);
                d.q[0][i] = f0 + eps;
                double fr = eval_integrand<F>(p, d);
                d.q[0][i] = f0 - eps;
                double fl = eval_integrand<F>(p, d);
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
            for (int j = 0; j < p.nvper; j++) {
                ids[j] = v.rix[(size_t) e * p.nvper + j];
                if (ids[j] == vtx) corner = j;
            }
            ElemData d;
            gather_elem<F>(p, v, ids, d);
            for (int i = 0; i < ps; i++) {
                double *slot = (p.wrt == 0 ? &d.q[corner][i] : &d.phi[corner]);
                double f0 = *slot;
                double eps = fdstepsize(f0);
                *slot = f0 + eps;
                double fr = eval_integrand<F>(p, d);
                *slot = f0 - eps;
                double fl = eval_integrand<F>(p, d);
                *slot = f0;
                acc[i] += (fr - fl) / (2 * eps);
            }
        }
    }
    for (int i = 0; i < ps; i++) out[(size_t) vtx * ps + i] = acc[i];
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

template <int F>
static cudaError_t run_total(const McuParams &p, const McuView &v, CompSum *partials, double *d_out, cudaStream_t st) {
    int blocks = mcu_total_grid(v.n);
    k_total<F><<<blocks, MCU_BLOCK, 0, st>>>(p, v, partials);
    k_total_finish<<<1, MCU_BLOCK, 0, st>>>(partials, blocks, d_out);
    g_mcu_launches += 2;
    return cudaGetLastError();
}

template <int F>
static cudaError_t run_integrand(const McuParams &p, const McuView &v, double *d_out, cudaStream_t st) {
    if (v.n > 0) {
        k_integrand<F><<<(v.n + MCU_BLOCK - 1) / MCU_BLOCK, MCU_BLOCK, 0, st>>>(p, v, d_out);
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

template <int F>
static cudaError_t run_fd_grad(const McuParams &p, const McuView &v, double *d_out, cudaStream_t st) {
    k_fd_grad<F><<<(p.nv + MCU_BLOCK - 1) / MCU_BLOCK, MCU_BLOCK, 0, st>>>(p, v, d_out);
    g_mcu_launches += 1;
    return cudaGetLastError();
}

template <int F>
static cudaError_t run_fd_fieldgrad(const McuParams &p, const McuView &v, double *d_out, cudaStream_t st) {
    k_fd_fieldgrad<F><<<(p.nv + MCU_BLOCK - 1) / MCU_BLOCK, MCU_BLOCK, 0, st>>>(p, v, d_out);
    g_mcu_launches += 1;
    return cudaGetLastError();
}

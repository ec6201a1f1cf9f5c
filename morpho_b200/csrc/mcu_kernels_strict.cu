// mcu_kernels_strict.cu — finite-difference family and field functionals.
// Compiled with -fmad=false: the reference's C arithmetic is un-fused (gcc, x86-64 baseline)
// and central differences amplify integrand round-off by 1/(2h) ~ 8e4.
#define MCU_STRICT 1
#include "mcu_kernels_impl.cuh"

bool mcu_strict_supports(int f, int mode) {
    switch (f) {
        case MCU_F_AREAENCLOSED: return mode == MCU_GRADIENT;
        case MCU_F_LINECURVATURESQ: case MCU_F_EQUIELEMENT: return mode == MCU_TOTAL || mode == MCU_INTEGRAND || mode == MCU_GRADIENT;
        case MCU_F_GRADSQ: case MCU_F_NEMATIC: case MCU_F_NEMATICELECTRIC: case MCU_F_NORMSQ:
            return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------
// LineCurvatureSq (functional.c:2940-3058): the element is a vertex; its integrand reads the
// two line elements incident to it (conn(1,0), mesh_findneighbors mesh.c:838-883).
// (pv, pk, delta): optional perturbation of coordinate pk of vertex pv used by the FD kernel.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double lc_coord(const McuParams &p, const McuView &v, int id, int k, int pv, int pk, double pval) {
    if (k >= p.dim) return 0.0;
    if (id == pv && k == pk) return pval;
    return v.vert[(size_t) id * p.dim + k];
}

__device__ double linecurvsq_integrand(const McuParams &p, const McuView &v, int id, int pv, int pk, double pval, bool *ok) {
    int nn = v.ltptr[id + 1] - v.ltptr[id];
    if (nn != 2) return 0.0;
    double s[2][3], sgn = -1.0;
    for (int i = 0; i < 2; i++) {
        int edge = v.ltidx[v.ltptr[id] + i];
        int e0 = v.lrix[2 * (size_t) edge], e1 = v.lrix[2 * (size_t) edge + 1];
        for (int k = 0; k < 3; k++) s[i][k] = lc_coord(p, v, e0, k, pv, pk, pval) - lc_coord(p, v, e1, k, pv, pk, pval);
        if (e0 != id) sgn *= -1;
    }
    double s0s0 = dotn(p.dim, s[0], s[0]), s0s1 = dotn(p.dim, s[0], s[1]), s1s1 = dotn(p.dim, s[1], s[1]);
    s0s0 = sqrt(s0s0);
    s1s1 = sqrt(s1s1);
    if (s0s0 < MCU_EPS || s1s1 < MCU_EPS) { *ok = false; return 0.0; }
    double u = sgn * s0s1 / s0s0 / s1s1, len = 0.5 * (s0s0 + s1s1);
    if (u < 1) u = acos(u); else u = 0;
    double result = u * u / len;
    if (p.integrandonly) result /= len;
    return result;
}

__global__ void __launch_bounds__(MCU_BLOCK) k_lc_total(McuParams p, McuView v, CompSum *partials, unsigned *flags) {
    CompSum acc = {0.0, 0.0};
    bool ok = true;
    const int stride = gridDim.x * MCU_BLOCK;
    for (int ce = blockIdx.x * MCU_BLOCK + threadIdx.x; ce < v.n; ce += stride) {
        int e = v.eids ? v.eids[ce] : ce;
        comp_add(acc, linecurvsq_integrand(p, v, e, -1, 0, 0.0, &ok));
    }
    if (!ok) atomicOr(flags, MCU_FLAG_DEGENERATE);
    CompSum r = block_reduce<MCU_BLOCK>(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

__global__ void __launch_bounds__(MCU_BLOCK) k_lc_integrand(McuParams p, McuView v, double *out, unsigned *flags) {
    int ce = blockIdx.x * MCU_BLOCK + threadIdx.x;
    if (ce >= v.n) return;
    int e = v.eids ? v.eids[ce] : ce;
    bool ok = true;
    out[e] = linecurvsq_integrand(p, v, e, -1, 0, 0.0, &ok);
    if (!ok) atomicOr(flags, MCU_FLAG_DEGENERATE);
}

// Gradient of Σ_i integrand(i) with respect to vertex vtx: elements that depend on vtx are vtx itself
// and the other end of each incident line (linecurvsq_dependencies, functional.c:2964-2989); the
// reference visits elements in ascending id and adds (fp-fm)/(2 eps) per coordinate
// (functional.c:557-651).  `sel` (0/1 per vertex element) restricts the element loop; with a
// selection the reference differentiates the element's own vertex twice (functional.c:596-606
// has no functional_containsvertex test) — reproduced via `selfcount`.
__global__ void __launch_bounds__(MCU_BLOCK) k_lc_grad(McuParams p, McuView v, const unsigned char *selmask, double *out, unsigned *flags) {
    int vtx = blockIdx.x * MCU_BLOCK + threadIdx.x;
    if (vtx >= p.nv) return;
    double acc[3] = {0.0, 0.0, 0.0};
    bool ok = true;
    int nn = v.ltptr[vtx + 1] - v.ltptr[vtx];
    // walk dependent elements in ascending id: merge {vtx} with the (unsorted) neighbour list
    int last = -1;
    for (int it = 0; it <= nn; it++) {
        // next smallest candidate id > last
        int cand = 0x7fffffff;
        if (vtx > last) cand = vtx;
        for (int i = 0; i < nn; i++) {
            int edge = v.ltidx[v.ltptr[vtx] + i];
            int e0 = v.lrix[2 * (size_t) edge], e1 = v.lrix[2 * (size_t) edge + 1];
            int other = (e0 == vtx ? e1 : e0);
            if (other > last && other < cand) cand = other;
        }
        if (cand == 0x7fffffff) break;
        last = cand;
        if (selmask && !selmask[cand]) continue;
        int reps = (selmask && cand == vtx) ? 2 : 1;
        for (int r = 0; r < reps; r++)
            for (int k = 0; k < p.dim; k++) {
                double x0 = v.vert[(size_t) vtx * p.dim + k];
                double eps = fdstepsize(x0);
                double fp = linecurvsq_integrand(p, v, cand, vtx, k, x0 + eps, &ok);
                double fm = linecurvsq_integrand(p, v, cand, vtx, k, x0 - eps, &ok);
                acc[k] = acc[k] + (fp - fm) / (2 * eps);
            }
    }
    for (int k = 0; k < p.dim; k++) out[(size_t) vtx * p.dim + k] = acc[k];
    if (!ok) atomicOr(flags, MCU_FLAG_DEGENERATE);
}

// ---------------------------------------------------------------------------------
// EquiElement (functional.c:2775-2922): the element is a vertex; its integrand compares the sizes of the
// elements of grade p.egrade incident to it (conn(egrade,0): v.ltptr / v.ltidx, ascending element id; their
// vertices in v.lrix) with their mean, optionally weighted.  Used as a mesh regulariser
// (examples/tactoid/tactoid.morpho:16).  (pv, pk, pval): the perturbed coordinate of the FD kernel.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double eq_size(const McuParams &p, const McuView &v, int el, int pv, int pk, double pval) {
    double x[MCU_MAXNV][3];
    const int nvp = p.egrade + 1;
    for (int j = 0; j < MCU_MAXNV; j++) {
        if (j >= nvp) break;
        const int id = v.lrix[(size_t) el * nvp + j];
        for (int k = 0; k < 3; k++) x[j][k] = lc_coord(p, v, id, k, pv, pk, pval);
    }
    return elem_size(p.egrade, x);
}

__device__ double equielement_integrand(const McuParams &p, const McuView &v, int id, int pv, int pk, double pval, bool *ok) {
    const int q0 = v.ltptr[id], nconn = v.ltptr[id + 1] - q0;
    if (nconn == 1) return 0.0;
    // two passes over the incident elements (mean, then the deviations) instead of an array of sizes in local memory:
    // the sizes are recomputed with identical arithmetic, so the result is the reference's
    double mean = 0.0, total = 0.0;
    for (int i = 0; i < nconn; i++) mean = mean + eq_size(p, v, v.ltidx[q0 + i], pv, pk, pval);
    mean = mean / ((double) nconn);
    if (fabs(mean) < MCU_EPS) { *ok = false; return 0.0; }
    if (!v.weight || fabs(p.wmean) < MCU_EPS) {
        for (int i = 0; i < nconn; i++) {
            const double t = 1.0 - eq_size(p, v, v.ltidx[q0 + i], pv, pk, pval) / mean;
            total = total + t * t;
        }
    } else {
        double wmean = 0.0;
        for (int i = 0; i < nconn; i++) { const int el = v.ltidx[q0 + i]; wmean = wmean + (el < p.nweight ? v.weight[el] : 1.0); }
        wmean = wmean / ((double) nconn);
        if (fabs(wmean) < MCU_EPS) wmean = 1.0;
        for (int i = 0; i < nconn; i++) {
            const int el = v.ltidx[q0 + i];
            const double w = (el < p.nweight ? v.weight[el] : 1.0);
            const double t = 1.0 - w * eq_size(p, v, el, pv, pk, pval) / mean / wmean;
            total = total + t * t;
        }
    }
    return total;
}

__global__ void __launch_bounds__(MCU_BLOCK) k_eq_total(McuParams p, McuView v, CompSum *partials, unsigned *flags) {
    CompSum acc = {0.0, 0.0};
    bool ok = true;
    const int stride = gridDim.x * MCU_BLOCK;
    for (int ce = blockIdx.x * MCU_BLOCK + threadIdx.x; ce < v.n; ce += stride) {
        int e = v.eids ? v.eids[ce] : ce;
        comp_add(acc, equielement_integrand(p, v, e, -1, 0, 0.0, &ok));
    }
    if (!ok) atomicOr(flags, MCU_FLAG_DEGENERATE);
    CompSum r = block_reduce<MCU_BLOCK>(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

__global__ void __launch_bounds__(MCU_BLOCK) k_eq_integrand(McuParams p, McuView v, double *out, unsigned *flags) {
    int ce = blockIdx.x * MCU_BLOCK + threadIdx.x;
    if (ce >= v.n) return;
    int e = v.eids ? v.eids[ce] : ce;
    bool ok = true;
    out[e] = equielement_integrand(p, v, e, -1, 0, 0.0, &ok);
    if (!ok) atomicOr(flags, MCU_FLAG_DEGENERATE);
}

// Gradient with respect to vertex vtx: the vertex elements that depend on it are vtx itself and every other vertex
// of its incident elements (equielement_dependencies, functional.c:2817-2844, is symmetric); the reference visits the
// vertex elements in ascending id and adds (fp - fm) / (2 eps) per coordinate (functional.c:557-651, 1142-1162).
__global__ void __launch_bounds__(MCU_BLOCK) k_eq_grad(McuParams p, McuView v, const unsigned char *selmask, double *out, unsigned *flags) {
    int vtx = blockIdx.x * MCU_BLOCK + threadIdx.x;
    if (vtx >= p.nv) return;
    double acc[3] = {0.0, 0.0, 0.0};
    bool ok = true;
    const int q0 = v.ltptr[vtx], nconn = v.ltptr[vtx + 1] - q0, nvp = p.egrade + 1;
    int last = -1;
    for (;;) {
        // next vertex element in ascending id: the smallest id > last among vtx and the vertices of its elements
        int cand = 0x7fffffff;
        if (vtx > last) cand = vtx;
        for (int i = 0; i < nconn; i++) {
            const int *ev = v.lrix + (size_t) v.ltidx[q0 + i] * nvp;
            for (int j = 0; j < nvp; j++) { const int o = ev[j]; if (o > last && o < cand) cand = o; }
        }
        if (cand == 0x7fffffff) break;
        last = cand;
        if (selmask && !selmask[cand]) continue;
        for (int k = 0; k < p.dim; k++) {
            double x0 = v.vert[(size_t) vtx * p.dim + k];
            double eps = fdstepsize(x0);
            double fp = equielement_integrand(p, v, cand, vtx, k, x0 + eps, &ok);
            double fm = equielement_integrand(p, v, cand, vtx, k, x0 - eps, &ok);
            acc[k] = acc[k] + (fp - fm) / (2 * eps);
        }
    }
    for (int k = 0; k < p.dim; k++) out[(size_t) vtx * p.dim + k] = acc[k];
    if (!ok) atomicOr(flags, MCU_FLAG_DEGENERATE);
}

#define DISPATCH(call)                                                    \
    switch (p.functional) {                                               \
        case MCU_F_AREAENCLOSED: return call(MCU_F_AREAENCLOSED);         \
        case MCU_F_GRADSQ: return call(MCU_F_GRADSQ);                     \
        case MCU_F_NORMSQ: return call(MCU_F_NORMSQ);                     \
        case MCU_F_NEMATIC: return call(MCU_F_NEMATIC);                   \
        case MCU_F_NEMATICELECTRIC: return call(MCU_F_NEMATICELECTRIC);   \
    }                                                                     \
    return cudaErrorInvalidValue;

cudaError_t mcu_strict_total(const McuParams &p, const McuView &v, CompSum *partials, double *d_out, cudaStream_t st) {
    if (p.functional == MCU_F_LINECURVATURESQ) {
        int blocks = mcu_total_grid(v.n);
        k_lc_total<<<blocks, MCU_BLOCK, 0, st>>>(p, v, partials, v.flags);
        k_total_finish<<<1, MCU_BLOCK, 0, st>>>(partials, blocks, d_out);
        g_mcu_launches += 2;
        return cudaGetLastError();
    }
    if (p.functional == MCU_F_EQUIELEMENT) {
        int blocks = mcu_total_grid(v.n);
        k_eq_total<<<blocks, MCU_BLOCK, 0, st>>>(p, v, partials, v.flags);
        k_total_finish<<<1, MCU_BLOCK, 0, st>>>(partials, blocks, d_out);
        g_mcu_launches += 2;
        return cudaGetLastError();
    }
#define CALL(F) run_total<F>(p, v, partials, d_out, st)
    DISPATCH(CALL)
#undef CALL
}

cudaError_t mcu_strict_integrand(const McuParams &p, const McuView &v, double *d_out, cudaStream_t st) {
    if (p.functional == MCU_F_LINECURVATURESQ) {
        if (v.n > 0) {
            k_lc_integrand<<<(v.n + MCU_BLOCK - 1) / MCU_BLOCK, MCU_BLOCK, 0, st>>>(p, v, d_out, v.flags);
            g_mcu_launches += 1;
        }
        return cudaGetLastError();
    }
    if (p.functional == MCU_F_EQUIELEMENT) {
        if (v.n > 0) {
            k_eq_integrand<<<(v.n + MCU_BLOCK - 1) / MCU_BLOCK, MCU_BLOCK, 0, st>>>(p, v, d_out, v.flags);
            g_mcu_launches += 1;
        }
        return cudaGetLastError();
    }
#define CALL(F) run_integrand<F>(p, v, d_out, st)
    DISPATCH(CALL)
#undef CALL
}

cudaError_t mcu_strict_gradient(const McuParams &p, const McuView &v, double *d_out, unsigned *flags, cudaStream_t st) {
    if (p.functional == MCU_F_LINECURVATURESQ) {
        k_lc_grad<<<(p.nv + MCU_BLOCK - 1) / MCU_BLOCK, MCU_BLOCK, 0, st>>>(p, v, v.selmask, d_out, flags);
        g_mcu_launches += 1;
        return cudaGetLastError();
    }
    if (p.functional == MCU_F_EQUIELEMENT) {
        k_eq_grad<<<(p.nv + MCU_BLOCK - 1) / MCU_BLOCK, MCU_BLOCK, 0, st>>>(p, v, v.selmask, d_out, flags);
        g_mcu_launches += 1;
        return cudaGetLastError();
    }
#define CALL(F) run_fd_grad<F>(p, v, d_out, st)
    DISPATCH(CALL)
#undef CALL
}

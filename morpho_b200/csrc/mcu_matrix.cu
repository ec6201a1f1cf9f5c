// mcu_matrix.cu — device-resident Matrix objects and the vector operations optimize.morpho applies to them
// (SURVEY.md 8f N1).
//
// A line search of the reference's optimisers (modules/optimize.morpho:461-486, 229-303) is, per trial step,
//     vsave = target.clone(); target.acc(-s, force); energy = sum of totals; mesh.setvertexmatrix(vsave)
// and per iteration a handful of gradients combined with inner / norm / acc / scalar multiples / column writes
// (fixgrad, optimize.morpho:686-693).  In the reference these are cblas_daxpy / ddot / dnrm2 / dscal calls on host
// arrays (src/linalg/matrix.c:495-500, 503-509, 526-529, 557-565, 586-604, 621-626).  Here the same operations run
// as HBM-bound kernels on matrices that never leave the device: a 10M-triangle step moves a few scalars over PCIe
// instead of 360 MB.
//
// Reductions are deterministic: every CTA folds a fixed contiguous chunk with compensated (two-sum) accumulation,
// partials are combined in CTA order by one finishing CTA — the same bits run to run, independent of scheduling.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "mcu_host.h"

extern long g_mcu_launches;

struct mcu_matrix {
    int nrows = 0, ncols = 0;
    size_t n = 0;
    double *d = nullptr;
    CompSum *d_part = nullptr;      // per-CTA partials of the reductions
    double *d_scalar = nullptr;     // result of a reduction
};

#define MX_THREADS 256
#define MX_RED_CTAS 592             // 148 SMs x 4

static double *g_hres = nullptr;    // pinned landing place of reduction results

static inline int mx_grid(size_t n) {
    size_t want = (n / 2 + MX_THREADS - 1) / MX_THREADS;       // two elements per thread and trip
    return (int) std::max<size_t>(1, std::min<size_t>(want, 148 * 8));
}

// out = a * sa + b * sb (b may be null: out = a * sa).  FUSED: one rounding per element (what cblas_daxpy / dscal do
// on FMA hardware); otherwise every product and sum is rounded (matrix_addscalar's `*=` then `+=`).
template <bool HASB>
static __global__ void __launch_bounds__(MX_THREADS) k_mx_axpby(double *__restrict__ out, const double *__restrict__ a, double sa,
                                                                 const double *__restrict__ b, double sb, size_t n) {
    const size_t stride = (size_t) gridDim.x * MX_THREADS;
    const size_t n2 = n / 2;
    const double2 *a2 = (const double2 *) a, *b2 = (const double2 *) b;
    double2 *o2 = (double2 *) out;
    for (size_t i = (size_t) blockIdx.x * MX_THREADS + threadIdx.x; i < n2; i += stride) {
        double2 x = a2[i], r;
        if (HASB) { double2 y = b2[i]; r.x = fma(sb, y.x, sa == 1.0 ? x.x : sa * x.x); r.y = fma(sb, y.y, sa == 1.0 ? x.y : sa * x.y); }
        else { r.x = sa * x.x; r.y = sa * x.y; }
        o2[i] = r;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const size_t i = n - 1;
        out[i] = HASB ? fma(sb, b[i], sa == 1.0 ? a[i] : sa * a[i]) : sa * a[i];
    }
}

// out = a * sgn + beta with two roundings (matrix_addscalar, matrix.c:557-565)
static __global__ void __launch_bounds__(MX_THREADS) k_mx_xpa(double *__restrict__ out, const double *__restrict__ a, double sgn, double beta, size_t n) {
    const size_t stride = (size_t) gridDim.x * MX_THREADS;
    for (size_t i = (size_t) blockIdx.x * MX_THREADS + threadIdx.x; i < n; i += stride) out[i] = __dadd_rn(__dmul_rn(a[i], sgn), beta);
}

// MODE 0: sum a_i b_i ; 1: sum a_i^2 ; 2: sum a_i
template <int MODE>
static __global__ void __launch_bounds__(MX_THREADS) k_mx_reduce(const double *__restrict__ a, const double *__restrict__ b, size_t n,
                                                                  CompSum *__restrict__ part) {
    const size_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const size_t lo = (size_t) blockIdx.x * chunk, hi = (lo + chunk < n) ? lo + chunk : n;
    CompSum acc = {0.0, 0.0};
    for (size_t i = lo + threadIdx.x; i < hi; i += MX_THREADS) {
        const double x = a[i];
        comp_add(acc, MODE == 0 ? x * b[i] : (MODE == 1 ? x * x : x));
    }
    CompSum r = block_reduce<MX_THREADS>(acc);
    if (threadIdx.x == 0) part[blockIdx.x] = r;
}
template <bool SQRT>
static __global__ void __launch_bounds__(MX_THREADS) k_mx_reduce_finish(const CompSum *__restrict__ part, int n, double *out) {
    CompSum acc = {0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += MX_THREADS) comp_merge(acc, part[i]);
    CompSum r = block_reduce<MX_THREADS>(acc);
    if (threadIdx.x == 0) out[0] = SQRT ? sqrt(r.s + r.c) : (r.s + r.c);
}

static __global__ void __launch_bounds__(MX_THREADS) k_mx_zero_columns(double *__restrict__ a, int nrows, const int *__restrict__ cols, int ncols_sel) {
    const int t = blockIdx.x * MX_THREADS + threadIdx.x;
    if (t >= ncols_sel * nrows) return;
    const int c = cols[t / nrows], r = t % nrows;
    a[(size_t) c * nrows + r] = 0.0;
}

#define CKM(call)                                                                                                   \
    do {                                                                                                            \
        cudaError_t _e = (call);                                                                                    \
        if (_e != cudaSuccess) { mcu_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); return MCU_ERR_CUDA; } \
    } while (0)

static int mx_same(const mcu_matrix *a, const mcu_matrix *b) { return a && b && a->nrows == b->nrows && a->ncols == b->ncols; }

extern "C" {

mcu_matrix *mcu_matrix_create(int nrows, int ncols) {
    if (nrows < 0 || ncols < 0) { mcu_set_error("bad matrix dimensions"); return nullptr; }
    if (mcu_init_if_needed() != MCU_OK) return nullptr;
    mcu_matrix *m = new mcu_matrix();
    m->nrows = nrows; m->ncols = ncols; m->n = (size_t) nrows * ncols;
    // 64 spare bytes: a matrix bound as a mesh's vertex store is read by bulk copies padded to an even row count
    if (cudaMalloc((void **) &m->d, sizeof(double) * std::max<size_t>(m->n, 1) + 64) != cudaSuccess ||
        cudaMalloc((void **) &m->d_part, sizeof(CompSum) * MX_RED_CTAS + 64) != cudaSuccess) {
        cudaGetLastError();
        mcu_set_error("cudaMalloc failed for a %d x %d matrix", nrows, ncols);
        if (m->d) cudaFree(m->d);
        delete m;
        return nullptr;
    }
    m->d_scalar = (double *) (m->d_part + MX_RED_CTAS);
    return m;
}

int mcu_matrix_destroy(mcu_matrix *m) {
    if (!m) return MCU_OK;
    cudaStreamSynchronize(mcu_stream());
    cudaFree(m->d); cudaFree(m->d_part);
    delete m;
    return MCU_OK;
}

int mcu_matrix_dims(const mcu_matrix *m, int *nrows, int *ncols) {
    if (!m) return MCU_ERR_ARGS;
    if (nrows) *nrows = m->nrows;
    if (ncols) *ncols = m->ncols;
    return MCU_OK;
}

double *mcu_matrix_device(mcu_matrix *m) { return m ? m->d : nullptr; }

int mcu_matrix_upload(mcu_matrix *m, const double *host) {
    if (!m || !host) return MCU_ERR_ARGS;
    CKM(cudaMemcpyAsync(m->d, host, sizeof(double) * m->n, cudaMemcpyHostToDevice, mcu_stream()));
    return MCU_OK;
}

int mcu_matrix_download(mcu_matrix *m, double *host) {
    if (!m || !host) return MCU_ERR_ARGS;
    CKM(cudaMemcpyAsync(host, m->d, sizeof(double) * m->n, cudaMemcpyDeviceToHost, mcu_stream()));
    CKM(cudaStreamSynchronize(mcu_stream()));
    return MCU_OK;
}

/* matrix_copy (matrix.c:503-509): dst <- src */
int mcu_matrix_copy(mcu_matrix *dst, const mcu_matrix *src) {
    if (!mx_same(dst, src)) { mcu_set_error("Matrices have incompatible shape."); return MCU_ERR_ARGS; }
    if (dst->d != src->d) CKM(cudaMemcpyAsync(dst->d, src->d, sizeof(double) * src->n, cudaMemcpyDeviceToDevice, mcu_stream()));
    return MCU_OK;
}

int mcu_matrix_zero(mcu_matrix *m) {
    if (!m) return MCU_ERR_ARGS;
    CKM(cudaMemsetAsync(m->d, 0, sizeof(double) * m->n, mcu_stream()));
    return MCU_OK;
}

/* out <- sa * a + sb * b ; b == NULL: out <- sa * a.  Covers matrix_axpy (y <- alpha x + y: out = y, a = y, sa = 1,
 * b = x, sb = alpha; matrix.c:495-500), Matrix + Matrix / Matrix - Matrix (clone then axpy, matrix.c:1098-1112) and
 * matrix_scale (matrix.c:526-529).  out may alias a or b. */
int mcu_matrix_axpby(mcu_matrix *out, const mcu_matrix *a, double sa, const mcu_matrix *b, double sb) {
    if (!mx_same(out, a) || (b && !mx_same(out, b))) { mcu_set_error("Matrices have incompatible shape."); return MCU_ERR_ARGS; }
    if (out->n == 0) return MCU_OK;
    if (b) k_mx_axpby<true><<<mx_grid(out->n), MX_THREADS, 0, mcu_stream()>>>(out->d, a->d, sa, b->d, sb, out->n);
    else k_mx_axpby<false><<<mx_grid(out->n), MX_THREADS, 0, mcu_stream()>>>(out->d, a->d, sa, nullptr, 0.0, out->n);
    g_mcu_launches += 1;
    CKM(cudaGetLastError());
    return MCU_OK;
}

/* matrix_addscalar (matrix.c:557-565) on a clone: out <- a * sgn + beta, each operation rounded */
int mcu_matrix_xpa(mcu_matrix *out, const mcu_matrix *a, double sgn, double beta) {
    if (!mx_same(out, a)) { mcu_set_error("Matrices have incompatible shape."); return MCU_ERR_ARGS; }
    if (out->n == 0) return MCU_OK;
    k_mx_xpa<<<mx_grid(out->n * 2), MX_THREADS, 0, mcu_stream()>>>(out->d, a->d, sgn, beta, out->n);
    g_mcu_launches += 1;
    CKM(cudaGetLastError());
    return MCU_OK;
}

static int mx_reduce(int mode, mcu_matrix *a, const mcu_matrix *b, double *host_out) {
    if (!a || !host_out || (mode == 0 && !mx_same(a, b))) { mcu_set_error("Matrices have incompatible shape."); return MCU_ERR_ARGS; }
    if (!g_hres) CKM(cudaMallocHost((void **) &g_hres, 64));
    cudaStream_t st = mcu_stream();
    const int grid = (int) std::max<size_t>(1, std::min<size_t>(MX_RED_CTAS, (a->n + 4 * MX_THREADS - 1) / (4 * MX_THREADS)));
    if (mode == 0) k_mx_reduce<0><<<grid, MX_THREADS, 0, st>>>(a->d, b->d, a->n, a->d_part);
    else if (mode == 1) k_mx_reduce<1><<<grid, MX_THREADS, 0, st>>>(a->d, nullptr, a->n, a->d_part);
    else k_mx_reduce<2><<<grid, MX_THREADS, 0, st>>>(a->d, nullptr, a->n, a->d_part);
    if (mode == 1) k_mx_reduce_finish<true><<<1, MX_THREADS, 0, st>>>(a->d_part, grid, a->d_scalar);
    else k_mx_reduce_finish<false><<<1, MX_THREADS, 0, st>>>(a->d_part, grid, a->d_scalar);
    g_mcu_launches += 2;
    CKM(cudaGetLastError());
    CKM(cudaMemcpyAsync(g_hres, a->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, st));
    CKM(cudaStreamSynchronize(st));
    *host_out = *g_hres;
    return MCU_OK;
}

/* matrix_inner (matrix.c:621-626, cblas_ddot): Frobenius inner product */
int mcu_matrix_inner(mcu_matrix *a, const mcu_matrix *b, double *host_out) { return mx_reduce(0, a, b, host_out); }
/* matrix_norm(MATRIX_NORM_FROBENIUS) (matrix.c:586-589, cblas_dnrm2) */
int mcu_matrix_norm(mcu_matrix *a, double *host_out) { return mx_reduce(1, a, nullptr, host_out); }
/* matrix_sum (matrix.c:591-604, Kahan) */
int mcu_matrix_sum(mcu_matrix *a, double *host_out) { return mx_reduce(2, a, nullptr, host_out); }

/* matrix_setcolumn / matrix_getcolumn (matrix.c:447-463): one column from / to host memory */
int mcu_matrix_set_column(mcu_matrix *m, int col, const double *host_col) {
    if (!m || !host_col || col < 0 || col >= m->ncols) { mcu_set_error("Matrix index out of bounds."); return MCU_ERR_ARGS; }
    CKM(cudaMemcpyAsync(m->d + (size_t) col * m->nrows, host_col, sizeof(double) * m->nrows, cudaMemcpyHostToDevice, mcu_stream()));
    return MCU_OK;
}
int mcu_matrix_get_column(mcu_matrix *m, int col, double *host_col) {
    if (!m || !host_col || col < 0 || col >= m->ncols) { mcu_set_error("Matrix index out of bounds."); return MCU_ERR_ARGS; }
    CKM(cudaMemcpyAsync(host_col, m->d + (size_t) col * m->nrows, sizeof(double) * m->nrows, cudaMemcpyDeviceToHost, mcu_stream()));
    CKM(cudaStreamSynchronize(mcu_stream()));
    return MCU_OK;
}

/* Optimizer.fixgrad (optimize.morpho:686-693): zero the columns of the fixed vertices, in one launch */
int mcu_matrix_zero_columns(mcu_matrix *m, int n, const int *host_cols) {
    if (!m || n < 0 || (n > 0 && !host_cols)) return MCU_ERR_ARGS;
    if (n == 0 || m->nrows == 0) return MCU_OK;
    for (int i = 0; i < n; i++) if (host_cols[i] < 0 || host_cols[i] >= m->ncols) { mcu_set_error("Matrix index out of bounds."); return MCU_ERR_ARGS; }
    int *d_cols = nullptr;
    cudaStream_t st = mcu_stream();
    CKM(cudaMallocAsync((void **) &d_cols, sizeof(int) * n, st));
    CKM(cudaMemcpyAsync(d_cols, host_cols, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    CKM(cudaStreamSynchronize(st));                      // host_cols belongs to the caller
    k_mx_zero_columns<<<(n * m->nrows + MX_THREADS - 1) / MX_THREADS, MX_THREADS, 0, st>>>(m->d, m->nrows, d_cols, n);
    g_mcu_launches += 1;
    CKM(cudaGetLastError());
    CKM(cudaFreeAsync(d_cols, st));
    return MCU_OK;
}

}  // extern "C"

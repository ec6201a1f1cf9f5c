// mcu_api.cu — C ABI of libmorpho_cuda.so (include/morpho_cuda.h): device mirrors of
// Mesh / Field / Selection, the evaluation dispatcher and measurement helpers.
//
// Host-side logic only; the kernels live in mcu_kernels_fast.cu, mcu_kernels_strict.cu,
// mcu_patch.cu and mcu_pairwise.cu.  Nothing here computes a functional on the CPU: when
// no CUDA device is usable every compute entry point fails with MCU_ERR_CUDA.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "mcu_host.h"
#include "mcu_patch.h"
#include "mcu_block.h"

long g_mcu_launches = 0;

static char g_err[512] = "";
static bool g_inited = false;
static int g_device = -1;
static cudaStream_t g_own_stream = nullptr;
static cudaStream_t g_stream = nullptr;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static unsigned *g_hflags = nullptr;   // pinned
static double *g_hscalar = nullptr;    // pinned
static std::map<std::string, long> g_opts;
static void *g_flush_buf = nullptr;
static double *g_scratch = nullptr;         // grow-only scratch of the element-centric finite-difference maps
static size_t g_scratch_bytes = 0;
static size_t g_flush_bytes = 0;

void mcu_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

cudaStream_t mcu_stream() { return g_stream; }
int mcu_init_if_needed() { return g_inited ? MCU_OK : mcu_init(g_device < 0 ? 0 : g_device); }
long mcu_opt(const char *key, long dflt) {
    auto it = g_opts.find(key);
    return it == g_opts.end() ? dflt : it->second;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            mcu_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return MCU_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

#define NEED_INIT()                                                     \
    do {                                                                \
        if (!g_inited) {                                                \
            int _rc = mcu_init(g_device < 0 ? 0 : g_device);            \
            if (_rc) return _rc;                                        \
        }                                                               \
    } while (0)

double *mcu_scratch(size_t bytes) {
    if (bytes <= g_scratch_bytes) return g_scratch;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return nullptr;
    if (bytes > free_b + g_scratch_bytes || bytes > total_b / 2) return nullptr;      // leave room for everything else
    cudaStreamSynchronize(g_stream);                     // the old buffer may still be in use by queued kernels
    if (g_scratch) cudaFree(g_scratch);
    g_scratch = nullptr; g_scratch_bytes = 0;
    bytes += bytes / 8;
    if (cudaMalloc((void **) &g_scratch, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    g_scratch_bytes = bytes;
    return g_scratch;
}

extern "C" {

int mcu_abi_version(void) { return MCU_ABI_VERSION; }

const char *mcu_last_error(void) { return g_err; }

const char *mcu_error_id(int status) {
    switch (status) {
        case MCU_ERR_ARGS: return "FnctlArgs";
        case MCU_ERR_ELNOTFOUND: return "FnctlELNtFnd";
        case MCU_ERR_VOLENCLZERO: return "VolEnclZero";
        case MCU_ERR_ALLOC: return "Alloc";
        default: return "";
    }
}

int mcu_init(int device) {
    if (g_inited) return MCU_OK;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        mcu_set_error("no CUDA device available (%s); libmorpho_cuda has no CPU fallback",
                      e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return MCU_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { mcu_set_error("device %d out of range (0..%d)", device, ndev - 1); return MCU_ERR_ARGS; }
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        mcu_set_error("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
        return MCU_ERR_CUDA;
    }
    g_device = device;
    CK(cudaStreamCreateWithFlags(&g_own_stream, cudaStreamNonBlocking));
    g_stream = g_own_stream;
    CK(cudaEventCreate(&g_ev0));
    CK(cudaEventCreate(&g_ev1));
    CK(cudaMallocHost((void **) &g_hflags, 64));
    CK(cudaMallocHost((void **) &g_hscalar, 64));
    *g_hflags = 0;
    g_inited = true;
    return MCU_OK;
}

static void graveyard_purge(const struct mcu_mesh *m);
int mcu_finalize(void) {
    if (!g_inited) return MCU_OK;
    cudaStreamSynchronize(g_stream);
    graveyard_purge(nullptr);
    if (g_scratch) cudaFree(g_scratch);
    g_scratch = nullptr; g_scratch_bytes = 0;
    if (g_flush_buf) cudaFree(g_flush_buf);
    g_flush_buf = nullptr; g_flush_bytes = 0;
    cudaFreeHost(g_hflags); cudaFreeHost(g_hscalar);
    cudaEventDestroy(g_ev0); cudaEventDestroy(g_ev1);
    cudaStreamDestroy(g_own_stream);
    g_own_stream = g_stream = nullptr;
    g_inited = false;
    return MCU_OK;
}

int mcu_set_option(const char *key, long value) { g_opts[key] = value; return MCU_OK; }
long mcu_get_option(const char *key) { return mcu_opt(key, 0); }

int mcu_set_stream(void *s) {
    NEED_INIT();
    g_stream = s ? (cudaStream_t) s : g_own_stream;
    return MCU_OK;
}

// Wait for the stream.  Status raised by asynchronously launched maps is per mesh: mcu_mesh_status.
int mcu_sync(void) {
    NEED_INIT();
    CK(cudaStreamSynchronize(g_stream));
    return MCU_OK;
}

/* ------------------------------------------------------------------ Mesh */

mcu_mesh *mcu_mesh_create(int dim, int nv) {
    if (!g_inited && mcu_init(g_device < 0 ? 0 : g_device)) return nullptr;
    if ((dim != 2 && dim != 3) || nv < 0) { mcu_set_error("mesh dim must be 2 or 3"); return nullptr; }
    mcu_mesh *m = new mcu_mesh();
    m->dim = dim; m->nv = nv;
    // + one spare row pair: the patch kernel copies runs of vertices padded to an even count (mcu_patch.cu)
    size_t bytes = sizeof(double) * ((size_t) std::max(nv, 1) + 2) * dim;
    if (cudaMalloc((void **) &m->d_vert, bytes) != cudaSuccess ||
        cudaMalloc((void **) &m->d_partials, sizeof(CompSum) * 148 * 8) != cudaSuccess ||
        cudaMalloc((void **) &m->d_scalar, 64) != cudaSuccess ||
        cudaMalloc((void **) &m->d_flags, 64) != cudaSuccess) {
        mcu_set_error("cudaMalloc failed for mesh mirror (%zu bytes)", bytes);
        delete m;
        return nullptr;
    }
    cudaMemsetAsync(m->d_flags, 0, 64, g_stream);
    m->d_vert_own = m->d_vert;
    return m;
}

static void conn_free(mcu_mesh *m, McuConn &c) {
    if (c.patches && m->halo) mcu_halo_detach_mesh(m);       // a fused exchange pushed through this patch set
    if (c.d_rix) cudaFree(c.d_rix);
    if (c.d_tptr) cudaFree(c.d_tptr);
    if (c.d_tidx) cudaFree(c.d_tidx);
    if (c.patches) { mcu_patchset_free(c.patches); c.patches = nullptr; }
    if (c.blocks) { mcu_blockset_free(c.blocks); c.blocks = nullptr; }
    c = McuConn();
}

int mcu_mesh_destroy(mcu_mesh *m) {
    if (!m) return MCU_OK;
    cudaStreamSynchronize(g_stream);
    graveyard_purge(m);
    for (int g = 0; g < 4; g++) conn_free(m, m->conn[g]);
    cudaFree(m->d_vert_own); cudaFree(m->d_partials); cudaFree(m->d_scalar); cudaFree(m->d_flags);
    if (m->d_result) cudaFree(m->d_result);
    delete m;
    return MCU_OK;
}

int mcu_mesh_set_vertices(mcu_mesh *m, const double *host) {
    if (!m || !host) { mcu_set_error("null argument"); return MCU_ERR_ARGS; }
    m->d_vert = m->d_vert_own; m->bound = nullptr;       // an upload always lands in the mesh's own buffer
    CK(cudaMemcpyAsync(m->d_vert, host, sizeof(double) * (size_t) m->nv * m->dim, cudaMemcpyHostToDevice, g_stream));
    m->vert_version++;
    return MCU_OK;
}

/* Mesh.setvertexmatrix with a device-resident Matrix (mesh.c:1145-1160 swaps the pointer mesh->vert; so does this):
 * the kernels read the vertices straight from `vert`, no copy.  NULL returns to the mesh's own buffer, whose contents
 * are then stale until the next mcu_mesh_set_vertices.  The matrix must outlive the binding. */
int mcu_mesh_bind_vertices(mcu_mesh *m, mcu_matrix *vert) {
    if (!m) { mcu_set_error("null argument"); return MCU_ERR_ARGS; }
    if (!vert) { m->d_vert = m->d_vert_own; m->bound = nullptr; return MCU_OK; }
    int nr = 0, nc = 0;
    mcu_matrix_dims(vert, &nr, &nc);
    if (nr != m->dim || nc != m->nv) { mcu_set_error("Vertex matrix dimensions inconsistent with mesh."); return MCU_ERR_ARGS; }
    m->d_vert = mcu_matrix_device(vert); m->bound = vert;
    m->vert_version++;
    return MCU_OK;
}

int mcu_mesh_set_vertices_device(mcu_mesh *m, const double *dev) {
    if (!m || !dev) { mcu_set_error("null argument"); return MCU_ERR_ARGS; }
    if (dev != m->d_vert)
        CK(cudaMemcpyAsync(m->d_vert, dev, sizeof(double) * (size_t) m->nv * m->dim, cudaMemcpyDeviceToDevice, g_stream));
    m->vert_version++;
    return MCU_OK;
}

int mcu_mesh_get_vertices(mcu_mesh *m, double *host) {
    if (!m || !host) { mcu_set_error("null argument"); return MCU_ERR_ARGS; }
    CK(cudaMemcpyAsync(host, m->d_vert, sizeof(double) * (size_t) m->nv * m->dim, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return MCU_OK;
}

double *mcu_mesh_vertices_device(mcu_mesh *m) { return m ? m->d_vert : nullptr; }

int mcu_mesh_set_connectivity(mcu_mesh *m, int g, int nel, const int *rix) {
    if (!m || g < 1 || g > 3 || g > m->dim || nel < 0 || (nel > 0 && !rix)) { mcu_set_error("bad connectivity arguments"); return MCU_ERR_ARGS; }
    CK(cudaStreamSynchronize(g_stream));
    McuConn &c = m->conn[g];
    conn_free(m, c);
    c.g = g; c.nvper = g + 1; c.nel = nel;
    size_t n = (size_t) nel * c.nvper;
    for (size_t i = 0; i < n; i++)
        if (rix[i] < 0 || rix[i] >= m->nv) { mcu_set_error("vertex id %d out of range in grade %d", rix[i], g); c = McuConn(); return MCU_ERR_ARGS; }
    c.h_rix.assign(rix, rix + n);
    CK(cudaMalloc((void **) &c.d_rix, sizeof(int) * std::max<size_t>(n, 1)));
    CK(cudaMemcpyAsync(c.d_rix, c.h_rix.data(), sizeof(int) * n, cudaMemcpyHostToDevice, g_stream));
    c.present = true;
    m->conn_version++;
    return MCU_OK;
}

int mcu_mesh_drop_connectivity(mcu_mesh *m, int g) {
    if (!m || g < 1 || g > 3) { mcu_set_error("bad connectivity arguments"); return MCU_ERR_ARGS; }
    if (!m->conn[g].present) return MCU_OK;
    CK(cudaStreamSynchronize(g_stream));
    conn_free(m, m->conn[g]);
    m->conn_version++;
    return MCU_OK;
}

int mcu_mesh_counts(const mcu_mesh *m, int *dim, int *nv, int nel[4]) {
    if (!m) return MCU_ERR_ARGS;
    if (dim) *dim = m->dim;
    if (nv) *nv = m->nv;
    if (nel) for (int g = 0; g < 4; g++) nel[g] = (g == 0 ? m->nv : (m->conn[g].present ? m->conn[g].nel : -1));
    return MCU_OK;
}

}  // extern "C"

// Stable counting-sort transpose: element ids ascending per vertex — the same result as the
// reference's sparse_transpose → cs_transpose (sparse.c:1227-1244).
void mcu_host_transpose(int nv, int n, int nvper, const int *rix, const int *eids, std::vector<int> &tptr, std::vector<int> &tidx) {
    tptr.assign((size_t) nv + 1, 0);
    tidx.resize((size_t) n * nvper);
    for (int ce = 0; ce < n; ce++) {
        const int *ids = rix + (size_t) (eids ? eids[ce] : ce) * nvper;
        for (int j = 0; j < nvper; j++) tptr[ids[j] + 1]++;
    }
    for (int v = 0; v < nv; v++) tptr[v + 1] += tptr[v];
    std::vector<int> cur(tptr.begin(), tptr.end() - 1);
    for (int ce = 0; ce < n; ce++) {
        const int *ids = rix + (size_t) (eids ? eids[ce] : ce) * nvper;
        for (int j = 0; j < nvper; j++) tidx[cur[ids[j]]++] = ce;
    }
}

extern "C" {

int mcu_ensure_transpose(mcu_mesh *m, int g) {
    McuConn &c = m->conn[g];
    if (c.d_tptr) return MCU_OK;
    std::vector<int> tptr, tidx;
    mcu_host_transpose(m->nv, c.nel, c.nvper, c.h_rix.data(), nullptr, tptr, tidx);
    CK(cudaMalloc((void **) &c.d_tptr, sizeof(int) * tptr.size()));
    CK(cudaMalloc((void **) &c.d_tidx, sizeof(int) * std::max<size_t>(tidx.size(), 1)));
    CK(cudaMemcpyAsync(c.d_tptr, tptr.data(), sizeof(int) * tptr.size(), cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(c.d_tidx, tidx.data(), sizeof(int) * tidx.size(), cudaMemcpyHostToDevice, g_stream));
    CK(cudaStreamSynchronize(g_stream));  // host vectors go out of scope
    return MCU_OK;
}

int mcu_mesh_get_transpose(mcu_mesh *m, int g, int *tptr, int *tidx) {
    if (!m || g < 1 || g > 3 || !m->conn[g].present) { mcu_set_error("no elements of grade %d", g); return MCU_ERR_ELNOTFOUND; }
    int rc = mcu_ensure_transpose(m, g);
    if (rc) return rc;
    McuConn &c = m->conn[g];
    CK(cudaMemcpyAsync(tptr, c.d_tptr, sizeof(int) * ((size_t) m->nv + 1), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaMemcpyAsync(tidx, c.d_tidx, sizeof(int) * (size_t) c.nel * c.nvper, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return MCU_OK;
}

/* ------------------------------------------------------------------ Field */

mcu_field *mcu_field_create(mcu_mesh *m, int psize) {
    if (!m || psize < 1 || psize > MCU_MAXPSIZE) { mcu_set_error("field psize must be 1..%d", MCU_MAXPSIZE); return nullptr; }
    mcu_field *f = new mcu_field();
    f->m = m; f->psize = psize;
    if (cudaMalloc((void **) &f->d_data, sizeof(double) * (size_t) std::max(m->nv, 1) * psize) != cudaSuccess) {
        mcu_set_error("cudaMalloc failed for field mirror");
        delete f;
        return nullptr;
    }
    return f;
}
int mcu_field_destroy(mcu_field *f) {
    if (!f) return MCU_OK;
    cudaStreamSynchronize(g_stream);
    cudaFree(f->d_data);
    delete f;
    return MCU_OK;
}
int mcu_field_set(mcu_field *f, const double *host) {
    if (!f || !host) return MCU_ERR_ARGS;
    CK(cudaMemcpyAsync(f->d_data, host, sizeof(double) * (size_t) f->m->nv * f->psize, cudaMemcpyHostToDevice, g_stream));
    return MCU_OK;
}
int mcu_field_set_device(mcu_field *f, const double *dev) {
    if (!f || !dev) return MCU_ERR_ARGS;
    if (dev != f->d_data)
        CK(cudaMemcpyAsync(f->d_data, dev, sizeof(double) * (size_t) f->m->nv * f->psize, cudaMemcpyDeviceToDevice, g_stream));
    return MCU_OK;
}
int mcu_field_get(mcu_field *f, double *host) {
    if (!f || !host) return MCU_ERR_ARGS;
    CK(cudaMemcpyAsync(host, f->d_data, sizeof(double) * (size_t) f->m->nv * f->psize, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return MCU_OK;
}
double *mcu_field_device(mcu_field *f) { return f ? f->d_data : nullptr; }

/* ------------------------------------------------------------------ Selection */

/* Callers that follow the reference's calling convention (the extension: a Selection object arrives with every call,
 * functional.c:82-102) create and destroy a selection mirror per evaluation.  Destroyed mirrors rest in a small
 * graveyard; a new selection with the same mesh, grade, connectivity version and ids takes its predecessor's device
 * arrays and block decomposition over instead of building them again. */
#define MCU_SEL_GRAVEYARD 8
static mcu_selection *g_sel_graveyard[MCU_SEL_GRAVEYARD];
static int g_sel_buried = 0;
static void selection_release(mcu_selection *s);

static void graveyard_purge(const mcu_mesh *m) {       // m == nullptr: everything
    int k = 0;
    for (int i = 0; i < g_sel_buried; i++) {
        mcu_selection *s = g_sel_graveyard[i];
        if (!m || s->m == m) { selection_release(s); delete s; }
        else g_sel_graveyard[k++] = s;
    }
    g_sel_buried = k;
}

mcu_selection *mcu_selection_create(mcu_mesh *m, int g, int n, const int *ids) {
    if (!m || g < 0 || g > 3 || n < 0 || (n > 0 && !ids)) { mcu_set_error("bad selection arguments"); return nullptr; }
    int limit = (g == 0 ? m->nv : (m->conn[g].present ? m->conn[g].nel : 0));
    mcu_selection *s = new mcu_selection();
    s->m = m; s->g = g;
    s->ids.assign(ids, ids + n);
    std::sort(s->ids.begin(), s->ids.end());
    s->ids.erase(std::unique(s->ids.begin(), s->ids.end()), s->ids.end());
    for (int i = 0; i < g_sel_buried; i++) {
        mcu_selection *old = g_sel_graveyard[i];
        if (old->m == m && old->g == g && old->conn_version == m->conn_version && old->ids == s->ids) {
            g_sel_graveyard[i] = g_sel_graveyard[--g_sel_buried];
            delete s;
            return old;
        }
    }
    for (int id : s->ids)
        if (id < 0 || id >= limit) { mcu_set_error("selected id %d out of range for grade %d", id, g); delete s; return nullptr; }
    s->conn_version = m->conn_version;
    return s;
}

static void selection_release(mcu_selection *s) {
    if (s->d_eids) cudaFree(s->d_eids);
    if (s->d_tptr) cudaFree(s->d_tptr);
    if (s->d_tidx) cudaFree(s->d_tidx);
    if (s->d_mask) cudaFree(s->d_mask);
    s->d_eids = s->d_tptr = s->d_tidx = nullptr; s->d_mask = nullptr;
    if (s->blocks) { mcu_blockset_free(s->blocks); s->blocks = nullptr; }
    s->block_failed = false;
}

int mcu_selection_destroy(mcu_selection *s) {
    if (!s) return MCU_OK;
    cudaStreamSynchronize(g_stream);
    if (g_sel_buried == MCU_SEL_GRAVEYARD) {                 // the oldest goes for good
        selection_release(g_sel_graveyard[0]);
        delete g_sel_graveyard[0];
        for (int i = 1; i < g_sel_buried; i++) g_sel_graveyard[i - 1] = g_sel_graveyard[i];
        g_sel_buried--;
    }
    g_sel_graveyard[g_sel_buried++] = s;
    return MCU_OK;
}

int mcu_selection_get_ids(mcu_selection *s, int *out, int cap) {
    if (!s) return -1;
    int n = (int) s->ids.size();
    if (out) for (int i = 0; i < n && i < cap; i++) out[i] = s->ids[i];
    return n;
}

int mcu_ensure_selection(mcu_selection *s) {
    mcu_mesh *m = s->m;
    if (s->d_eids && s->conn_version == m->conn_version) return MCU_OK;
    selection_release(s);
    int n = (int) s->ids.size();
    CK(cudaMalloc((void **) &s->d_eids, sizeof(int) * std::max(n, 1)));
    CK(cudaMemcpyAsync(s->d_eids, s->ids.data(), sizeof(int) * n, cudaMemcpyHostToDevice, g_stream));
    std::vector<int> tptr, tidx;
    std::vector<unsigned char> mask;
    if (s->g == 0) {
        // grade-0 elements are the vertices themselves: "connectivity" is the identity
        tptr.assign((size_t) m->nv + 1, 0);
        mask.assign((size_t) std::max(m->nv, 1), 0);
        for (int id : s->ids) { tptr[id + 1] = 1; mask[id] = 1; }
        for (int v = 0; v < m->nv; v++) tptr[v + 1] += tptr[v];
        tidx.resize(std::max(n, 1));
        for (int i = 0; i < n; i++) tidx[i] = i;
        CK(cudaMalloc((void **) &s->d_mask, mask.size()));
        CK(cudaMemcpyAsync(s->d_mask, mask.data(), mask.size(), cudaMemcpyHostToDevice, g_stream));
    } else {
        McuConn &c = m->conn[s->g];
        if (!c.present) { mcu_set_error("no elements of grade %d", s->g); return MCU_ERR_ELNOTFOUND; }
        mcu_host_transpose(m->nv, n, c.nvper, c.h_rix.data(), s->ids.data(), tptr, tidx);
    }
    CK(cudaMalloc((void **) &s->d_tptr, sizeof(int) * tptr.size()));
    CK(cudaMalloc((void **) &s->d_tidx, sizeof(int) * std::max<size_t>(tidx.size(), 1)));
    CK(cudaMemcpyAsync(s->d_tptr, tptr.data(), sizeof(int) * tptr.size(), cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(s->d_tidx, tidx.data(), sizeof(int) * tidx.size(), cudaMemcpyHostToDevice, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    s->conn_version = m->conn_version;
    return MCU_OK;
}

/* ------------------------------------------------------------------ Evaluation */

static int mesh_maxgrade(const mcu_mesh *m) {  // mesh_maxgrade, mesh.c:334-340
    for (int g = m->dim; g > 0; g--) if (m->conn[g].present) return g;
    return 0;
}

static int functional_grade(const mcu_mesh *m, const mcu_functional *f) {
    switch (f->functional) {
        case MCU_F_LENGTH: case MCU_F_AREAENCLOSED: return 1;
        case MCU_F_AREA: case MCU_F_VOLUMEENCLOSED: return 2;
        case MCU_F_VOLUME: return 3;
        case MCU_F_LINECURVATURESQ: case MCU_F_NORMSQ: case MCU_F_EQUIELEMENT: return 0;
        case MCU_F_GRADSQ: case MCU_F_NEMATIC: case MCU_F_NEMATICELECTRIC:
            return f->grade > 0 ? f->grade : mesh_maxgrade(m);  // gradsq_prepareref, functional.c:3622-3642
    }
    return -1;
}

long mcu_result_size(mcu_mesh *m, const mcu_functional *f, int mode) {
    if (!m || !f) return -1;
    int g = functional_grade(m, f);
    if (g < 0) return -1;
    switch (mode) {
        case MCU_TOTAL: return 1;
        case MCU_INTEGRAND: return g == 0 ? m->nv : (m->conn[g].present ? m->conn[g].nel : -1);
        case MCU_GRADIENT: return (long) m->nv * m->dim;
        case MCU_FIELDGRADIENT: {
            mcu_field *fl = (f->wrt >= 0 && f->wrt < 2) ? f->field[f->wrt] : nullptr;
            return fl ? (long) m->nv * fl->psize : -1;
        }
    }
    return -1;
}

}  // extern "C"
int mcu_ensure_patches(mcu_mesh *m, int g) {
    if (!m || g != 2 || m->dim != 3 || !m->conn[g].present) return MCU_ERR_UNSUPPORTED;
    McuConn &c = m->conn[g];
    if (c.patches) return MCU_OK;
    // a refused plan is remembered until the connectivity changes (conn_free resets the flag); planning needs real
    // coordinates (the bisection works on positions), so not before the first upload
    if (c.patch_failed || m->vert_version == 0) return MCU_ERR_UNSUPPORTED;
    cudaStream_t st = g_stream;
    CK(cudaStreamSynchronize(st));
    std::vector<double> hv((size_t) m->nv * 3);
    CK(cudaMemcpy(hv.data(), m->d_vert, sizeof(double) * hv.size(), cudaMemcpyDeviceToHost));
    c.patches = mcu_patchset_build(m->nv, c.nel, c.h_rix.data(), hv.data(), st);
    if (!c.patches) c.patch_failed = true;
    return c.patches ? MCU_OK : MCU_ERR_UNSUPPORTED;
}
// Block decomposition of grade g (of the selected elements if sel): built on first use from the current positions,
// kept until the connectivity (or the selection) changes.  nullptr: the planner refused (remembered).
static McuBlockSet *mcu_ensure_blocks(mcu_mesh *m, int g, mcu_selection *sel) {
    McuConn &c = m->conn[g];
    if (!c.present || m->vert_version == 0) return nullptr;
    McuBlockSet **slot = sel ? &sel->blocks : &c.blocks;
    bool *failed = sel ? &sel->block_failed : &c.block_failed;
    if (sel && sel->blocks && sel->blocks_version != m->conn_version) { mcu_blockset_free(sel->blocks); sel->blocks = nullptr; sel->block_failed = false; }
    if (*slot) return *slot;
    if (*failed) return nullptr;
    if (cudaStreamSynchronize(g_stream) != cudaSuccess) return nullptr;
    std::vector<double> hv((size_t) m->nv * m->dim);
    if (cudaMemcpy(hv.data(), m->d_vert, sizeof(double) * hv.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return nullptr;
    *slot = mcu_blockset_build(m->nv, m->dim, sel ? (int) sel->ids.size() : c.nel, c.nvper, c.h_rix.data(), sel ? sel->ids.data() : nullptr, hv.data(), g_stream);
    if (!*slot) *failed = true;
    if (sel) sel->blocks_version = m->conn_version;
    return *slot;
}

extern "C" {

static int eval_impl(mcu_mesh *m, const mcu_functional *f, mcu_selection *sel, int mode, double *d_out) {
    const int F = f->functional;
    const int g = functional_grade(m, f);
    if (g < 0) { mcu_set_error("unknown functional %d", F); return MCU_ERR_UNSUPPORTED; }
    const bool fast = mcu_fast_supports(F, mode), strict = mcu_strict_supports(F, mode);
    if (!fast && !strict) { mcu_set_error("functional %d has no accelerated method %d", F, mode); return MCU_ERR_UNSUPPORTED; }
    if (g > 0 && !m->conn[g].present) { mcu_set_error("mesh has no elements of grade %d", g); return MCU_ERR_ELNOTFOUND; }
    if (g > m->dim) return MCU_ERR_ELNOTFOUND;
    if (sel && (sel->m != m)) { mcu_set_error("selection belongs to another mesh"); return MCU_ERR_ARGS; }
    if ((F == MCU_F_VOLUMEENCLOSED || F == MCU_F_VOLUME) && m->dim != 3) { mcu_set_error("functional needs a 3D mesh"); return MCU_ERR_ARGS; }

    McuParams p;
    memset(&p, 0, sizeof(p));
    p.functional = F; p.g = g; p.nvper = (g == 0 ? 1 : g + 1); p.dim = m->dim; p.nv = m->nv;
    p.nel = (g == 0 ? m->nv : m->conn[g].nel);
    p.ksplay = f->ksplay; p.ktwist = f->ktwist; p.kbend = f->kbend; p.pitch = f->pitch;
    p.haspitch = f->haspitch; p.integrandonly = f->integrandonly; p.wrt = f->wrt;

    McuView v;
    memset(&v, 0, sizeof(v));
    v.vert = m->d_vert;
    v.rix = (g > 0 ? m->conn[g].d_rix : nullptr);
    v.n = p.nel;
    v.flags = m->d_flags;

    const bool needs_field = (F == MCU_F_GRADSQ || F == MCU_F_NORMSQ || F == MCU_F_NEMATIC || F == MCU_F_NEMATICELECTRIC);
    if (needs_field) {
        if (!f->field[0] || f->field[0]->m != m) { mcu_set_error("functional needs a field on this mesh"); return MCU_ERR_ARGS; }
        p.psize0 = f->field[0]->psize; v.f0 = f->field[0]->d_data;
        if ((F == MCU_F_NEMATIC || F == MCU_F_NEMATICELECTRIC) && p.psize0 < 3 && m->dim == 3) { mcu_set_error("director field needs 3 components"); return MCU_ERR_ARGS; }
        if ((F == MCU_F_GRADSQ || F == MCU_F_NEMATIC || F == MCU_F_NEMATICELECTRIC) && g != 2 && g != 3) { mcu_set_error("functional acts on grade 2 or 3"); return MCU_ERR_ARGS; }
        if (g == 3 && m->dim != 3) return MCU_ERR_ARGS;
        if (F == MCU_F_NEMATICELECTRIC) {
            if (!f->field[1] || f->field[1]->m != m || f->field[1]->psize != 1) { mcu_set_error("NematicElectric needs a scalar potential field"); return MCU_ERR_ARGS; }
            p.psize1 = 1; v.f1 = f->field[1]->d_data;
        }
        if (mode == MCU_FIELDGRADIENT) {
            if (f->wrt < 0 || f->wrt > 1 || !f->field[f->wrt]) { mcu_set_error("fieldgradient: bad target field"); return MCU_ERR_ARGS; }
            if (f->wrt == 1 && F != MCU_F_NEMATICELECTRIC) return MCU_ERR_ARGS;
        }
    } else if (mode == MCU_FIELDGRADIENT) {
        mcu_set_error("functional has no fieldgradient");
        return MCU_ERR_UNSUPPORTED;
    }

    if (sel) {
        if (sel->g != g) {
            // the reference reads sel->selected[g] (functional.c:226-229): a selection holding ids of another grade
            // selects nothing of this one — the total is 0, integrand and gradients are zero matrices
            long nout = (mode == MCU_TOTAL) ? 1 : (mode == MCU_INTEGRAND ? p.nel : (mode == MCU_GRADIENT ? (long) m->nv * m->dim
                        : (long) m->nv * f->field[f->wrt]->psize));
            CK(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t) nout, g_stream));
            return MCU_OK;
        }
        int rc = mcu_ensure_selection(sel);
        if (rc) return rc;
        v.eids = sel->d_eids; v.n = (int) sel->ids.size();
        v.tptr = sel->d_tptr; v.tidx = sel->d_tidx; v.selmask = sel->d_mask;
    }
    if (F == MCU_F_LINECURVATURESQ) {
        if (!m->conn[1].present) { mcu_set_error("LineCurvatureSq needs line elements"); return MCU_ERR_ELNOTFOUND; }
        int rc = mcu_ensure_transpose(m, 1);
        if (rc) return rc;
        v.lrix = m->conn[1].d_rix; v.ltptr = m->conn[1].d_tptr; v.ltidx = m->conn[1].d_tidx;
    }
    double *d_weight = nullptr;
    if (F == MCU_F_EQUIELEMENT) {
        // equielement_prepareref (functional.c:2775-2806): a grade outside 0..maxgrade means the highest grade present
        const int maxg = mesh_maxgrade(m);
        const int eg = (f->grade < 0 || f->grade > maxg) ? maxg : f->grade;
        if (eg < 1 || !m->conn[eg].present) { mcu_set_error("EquiElement needs elements of grade >= 1"); return MCU_ERR_UNSUPPORTED; }
        int rc = mcu_ensure_transpose(m, eg);
        if (rc) return rc;
        p.egrade = eg;
        v.lrix = m->conn[eg].d_rix; v.ltptr = m->conn[eg].d_tptr; v.ltidx = m->conn[eg].d_tidx;
        if (f->host_weight && f->nweight > 0) {
            double sum = 0.0, cc = 0.0;                       // matrix_sum (matrix.c:591-603) / ncols
            for (int i = 0; i < f->nweight; i++) { double y = f->host_weight[i] - cc, t = sum + y; cc = (t - sum) - y; sum = t; }
            p.wmean = sum / f->nweight; p.nweight = f->nweight;
            CK(cudaMallocAsync((void **) &d_weight, sizeof(double) * (size_t) f->nweight, g_stream));
            CK(cudaMemcpyAsync(d_weight, f->host_weight, sizeof(double) * (size_t) f->nweight, cudaMemcpyHostToDevice, g_stream));
            v.weight = d_weight;
        }
    }
    const bool gather = (mode == MCU_GRADIENT || mode == MCU_FIELDGRADIENT) && g > 0;
    if (gather && !sel) {
        int rc = mcu_ensure_transpose(m, g);
        if (rc) return rc;
        v.tptr = m->conn[g].d_tptr; v.tidx = m->conn[g].d_tidx;
    }

    cudaStream_t st = g_stream;
    cudaError_t e = cudaSuccess;
    switch (mode) {
        case MCU_TOTAL:
            // large triangle meshes: the patch kernel (coordinates staged once per patch, every triangle counted once)
            if (fast && !sel && g == 2 && m->dim == 3 && (F == MCU_F_AREA || F == MCU_F_VOLUMEENCLOSED) &&
                mcu_opt("gradient_path", MCU_PATH_AUTO) != MCU_PATH_GATHER && m->conn[g].nel >= mcu_opt("patch_min_elements", 4096) &&
                mcu_ensure_patches(m, g) == MCU_OK) {
                e = mcu_patch_total(F, *m->conn[g].patches, m->d_vert, d_out, m->d_flags, st);
                break;
            }
            e = fast ? mcu_fast_total(p, v, m->d_partials, d_out, st) : mcu_strict_total(p, v, m->d_partials, d_out, st);
            break;
        case MCU_INTEGRAND:
            if (sel) CK(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t) p.nel, st));
            e = fast ? mcu_fast_integrand(p, v, d_out, st) : mcu_strict_integrand(p, v, d_out, st);
            break;
        case MCU_GRADIENT: {
            long path = mcu_opt("gradient_path", MCU_PATH_AUTO);
            bool patchable = fast && !sel && g == 2 && m->dim == 3 && (F == MCU_F_AREA || F == MCU_F_VOLUMEENCLOSED);
            if (path == MCU_PATH_PATCH && !patchable) { mcu_set_error("patch path not applicable"); return MCU_ERR_UNSUPPORTED; }
            bool use_patch = patchable && (path == MCU_PATH_PATCH || (path == MCU_PATH_AUTO && m->conn[g].nel >= mcu_opt("patch_min_elements", 4096)));
            if (use_patch) {
                McuConn &c = m->conn[g];
                if (!c.patches && mcu_ensure_patches(m, g) != MCU_OK) {
                    if (path == MCU_PATH_PATCH) return MCU_ERR_UNSUPPORTED;
                    use_patch = false;
                }
                if (use_patch) {
                    McuHaloPush push;
                    bool pushing = m->halo && mcu_halo_push_params(m->halo, m->halo_slot, m->dim, &push);
                    e = mcu_patch_gradient(F, *c.patches, m->d_vert, d_out, m->d_flags, st, pushing ? &push : nullptr);
                    break;
                }
            }
            // any grade / selections: the shared-memory block kernel (mcu_block.cu); small meshes stay on the gather kernel
            const bool blockable = fast && g >= 1 && (F == MCU_F_LENGTH || F == MCU_F_AREA || F == MCU_F_VOLUMEENCLOSED || F == MCU_F_VOLUME);
            if (path == MCU_PATH_BLOCK && !blockable) { mcu_set_error("block path not applicable"); return MCU_ERR_UNSUPPORTED; }
            if (blockable && (path == MCU_PATH_BLOCK || (path == MCU_PATH_AUTO && v.n >= mcu_opt("block_min_elements", 32768)))) {
                McuBlockSet *bs = mcu_ensure_blocks(m, g, sel);
                if (bs) { e = mcu_block_gradient(F, *bs, m->d_vert, m->nv, m->dim, d_out, m->d_flags, st); break; }
                if (path == MCU_PATH_BLOCK) return MCU_ERR_UNSUPPORTED;
            }
            e = fast ? mcu_fast_gradient(p, v, d_out, m->d_flags, st) : mcu_strict_gradient(p, v, d_out, m->d_flags, st);
            break;
        }
        case MCU_FIELDGRADIENT:
            e = mcu_strict_fieldgradient(p, v, d_out, st);
            break;
        default:
            return MCU_ERR_ARGS;
    }
    if (d_weight) cudaFreeAsync(d_weight, g_stream);
    if (e != cudaSuccess) { mcu_set_error("kernel launch failed: %s", cudaGetErrorString(e)); return MCU_ERR_CUDA; }
    return MCU_OK;
}

static int check_flags(mcu_mesh *m) {
    // copy the status word back together with the result; reset it for the next call
    CK(cudaMemcpyAsync(g_hflags, m->d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaMemsetAsync(m->d_flags, 0, sizeof(unsigned), g_stream));
    return MCU_OK;
}

static int flags_to_status(unsigned fl) {
    if (fl & MCU_FLAG_VOLENCLZERO) { mcu_set_error("VolumeEnclosed detected an element of zero size. Check that a mesh point is not coincident with the origin."); return MCU_ERR_VOLENCLZERO; }
    if (fl & MCU_FLAG_DEGENERATE) { mcu_set_error("degenerate element: the reference returns nil here"); return MCU_ERR_DEGENERATE; }
    return MCU_OK;
}

// Status of the asynchronous (mcu_eval_device) evaluations on this mesh since the last mcu_mesh_status / mcu_eval:
// waits for the stream, reads and clears the status word.
int mcu_mesh_status(mcu_mesh *m) {
    NEED_INIT();
    if (!m) return MCU_ERR_ARGS;
    int rc = check_flags(m);
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_stream));
    return flags_to_status(*g_hflags);
}

int mcu_eval_device(mcu_mesh *m, const mcu_functional *f, mcu_selection *sel, int mode, double *dev_out) {
    NEED_INIT();
    if (!m || !f || !dev_out) { mcu_set_error("null argument"); return MCU_ERR_ARGS; }
    return eval_impl(m, f, sel, mode, dev_out);
}

/* the gradient / integrand straight into a device-resident Matrix (the object the builtin returns stays on the device) */
int mcu_eval_matrix(mcu_mesh *m, const mcu_functional *f, mcu_selection *sel, int mode, mcu_matrix *out) {
    NEED_INIT();
    if (!m || !f || !out) { mcu_set_error("null argument"); return MCU_ERR_ARGS; }
    long n = mcu_result_size(m, f, mode);
    int nr = 0, nc = 0;
    mcu_matrix_dims(out, &nr, &nc);
    if (n < 0 || (long) nr * nc != n) { mcu_set_error("result matrix has the wrong size"); return MCU_ERR_ARGS; }
    return eval_impl(m, f, sel, mode, mcu_matrix_device(out));
}

int mcu_eval(mcu_mesh *m, const mcu_functional *f, mcu_selection *sel, int mode, double *host_out) {
    NEED_INIT();
    if (!m || !f || !host_out) { mcu_set_error("null argument"); return MCU_ERR_ARGS; }
    long n = mcu_result_size(m, f, mode);
    if (n < 0) {
        int g = functional_grade(m, f);
        if (g > 0 && g <= 3 && !m->conn[g].present) { mcu_set_error("mesh has no elements of grade %d", g); return MCU_ERR_ELNOTFOUND; }
        mcu_set_error("cannot size the result");
        return MCU_ERR_ARGS;
    }
    double *d_out = m->d_scalar;
    if (mode != MCU_TOTAL) {
        if ((size_t) n > m->result_cap) {
            CK(cudaStreamSynchronize(g_stream));
            if (m->d_result) cudaFree(m->d_result);
            m->d_result = nullptr; m->result_cap = 0;
            CK(cudaMalloc((void **) &m->d_result, sizeof(double) * (size_t) std::max<long>(n, 1)));
            m->result_cap = (size_t) n;
        }
        d_out = m->d_result;
    }
    // status raised by earlier mcu_eval_device calls belongs to those (mcu_mesh_status), not to this evaluation
    CK(cudaMemsetAsync(m->d_flags, 0, sizeof(unsigned), g_stream));
    int rc = eval_impl(m, f, sel, mode, d_out);
    if (rc) return rc;
    if (n > 0) CK(cudaMemcpyAsync(host_out, d_out, sizeof(double) * (size_t) n, cudaMemcpyDeviceToHost, g_stream));
    rc = check_flags(m);
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_stream));
    return flags_to_status(*g_hflags);
}

/* ------------------------------------------------------------------ Pairwise */

int mcu_pairwise_device(int n, int dim, const double *dev_x, int kind, const double *params, double cutoff, int mode, double *dev_out) {
    NEED_INIT();
    if (n < 0 || (dim != 2 && dim != 3) || !dev_x || !dev_out) { mcu_set_error("bad pairwise arguments"); return MCU_ERR_ARGS; }
    static CompSum *partials = nullptr;
    if (!partials) CK(cudaMalloc((void **) &partials, sizeof(CompSum) * 4096));
    int status = MCU_OK;
    cudaError_t e = mcu_pairwise_launch(n, dim, dev_x, kind, params, cutoff, mode, dev_out, partials, g_stream, 0, -1, &status);
    if (e != cudaSuccess) { mcu_set_error("pairwise launch failed: %s", cudaGetErrorString(e)); return MCU_ERR_CUDA; }
    if (status) mcu_set_error("bad pairwise potential (kind %d)", kind);
    return status;
}

/* rows [row0, row0 + nrows) of the all-pairs interaction: the block one GPU computes when the N x N matrix is
 * partitioned by rows across GPUs (every GPU holds all n positions).  dev_out: nrows (MCU_INTEGRAND), nrows * dim
 * (MCU_GRADIENT) doubles, or the partial energy of those rows (MCU_TOTAL, 1 double). */
int mcu_pairwise_rows_device(int n, int dim, const double *dev_x, int kind, const double *params, double cutoff, int mode,
                             int row0, int nrows, double *dev_out) {
    NEED_INIT();
    if (n < 0 || (dim != 2 && dim != 3) || !dev_x || !dev_out || row0 < 0 || nrows < 0 || row0 + nrows > n) { mcu_set_error("bad pairwise arguments"); return MCU_ERR_ARGS; }
    static CompSum *partials = nullptr;
    if (!partials) CK(cudaMalloc((void **) &partials, sizeof(CompSum) * 4096));
    int status = MCU_OK;
    cudaError_t e = mcu_pairwise_launch(n, dim, dev_x, kind, params, cutoff, mode, dev_out, partials, g_stream, row0, nrows, &status);
    if (e != cudaSuccess) { mcu_set_error("pairwise launch failed: %s", cudaGetErrorString(e)); return MCU_ERR_CUDA; }
    if (status) mcu_set_error("bad pairwise potential (kind %d)", kind);
    return status;
}

int mcu_pairwise(int n, int dim, const double *host_x, int kind, const double *params, double cutoff, int mode, double *host_out) {
    NEED_INIT();
    if (n < 0 || !host_x || !host_out) return MCU_ERR_ARGS;
    double *d_x = nullptr, *d_out = nullptr;
    size_t nout = (mode == MCU_TOTAL ? 1 : (mode == MCU_INTEGRAND ? (size_t) n : (size_t) n * dim));
    CK(cudaMalloc((void **) &d_x, sizeof(double) * std::max<size_t>((size_t) n * dim, 1)));
    CK(cudaMalloc((void **) &d_out, sizeof(double) * std::max<size_t>(nout, 1)));
    CK(cudaMemcpyAsync(d_x, host_x, sizeof(double) * (size_t) n * dim, cudaMemcpyHostToDevice, g_stream));
    int rc = mcu_pairwise_device(n, dim, d_x, kind, params, cutoff, mode, d_out);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(host_out, d_out, sizeof(double) * nout, cudaMemcpyDeviceToHost, g_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
        if (e != cudaSuccess) { mcu_set_error("pairwise copy failed: %s", cudaGetErrorString(e)); rc = MCU_ERR_CUDA; }
    }
    cudaFree(d_x); cudaFree(d_out);
    return rc;
}

int mcu_pairwise_program(int nops_f, const mcu_expr_op *prog_f, int nops_g, const mcu_expr_op *prog_g) {
    int rc = mcu_pairwise_store_program(nops_f, prog_f, nops_g, prog_g);
    if (rc) mcu_set_error("pairwise program rejected (unsupported op, unbalanced stack, or longer than 64 ops)");
    return rc;
}

/* ------------------------------------------------------------------ Partition / measurement */

// functional_binbounds (functional.c:863-879): nel/nbins per bin, the remainder spread over the first bins
int mcu_binbounds(int nel, int nbins, int *bounds) {
    if (nbins < 1 || nel < 0 || !bounds) return MCU_ERR_ARGS;
    int base = nel / nbins, rem = nel % nbins, b = 0;
    for (int i = 0; i <= nbins; i++) { bounds[i] = b; if (i < nbins) b += base + (i < rem ? 1 : 0); }
    return MCU_OK;
}

int mcu_timer_start(void) { NEED_INIT(); CK(cudaEventRecord(g_ev0, g_stream)); return MCU_OK; }
int mcu_timer_stop(float *ms) {
    NEED_INIT();
    CK(cudaEventRecord(g_ev1, g_stream));
    CK(cudaEventSynchronize(g_ev1));
    CK(cudaEventElapsedTime(ms, g_ev0, g_ev1));
    return MCU_OK;
}
long mcu_launch_count(void) { return g_mcu_launches; }

int mcu_device_malloc(void **ptr, size_t bytes) { NEED_INIT(); CK(cudaMalloc(ptr, bytes)); return MCU_OK; }
int mcu_device_free(void *ptr) { CK(cudaFree(ptr)); return MCU_OK; }

int mcu_flush_l2(void) {
    NEED_INIT();
    if (!g_flush_buf) {
        g_flush_bytes = (size_t) 256 << 20;  // 2x the 126 MB L2
        CK(cudaMalloc(&g_flush_buf, g_flush_bytes));
    }
    static int toggle = 0;
    CK(cudaMemsetAsync(g_flush_buf, ++toggle & 0xff, g_flush_bytes, g_stream));
    return MCU_OK;
}

}  // extern "C"

// mcu_halo.cu — completion of the gradient rows of vertices shared between GPUs, over NVLink peer memory.
//
// Multi-GPU evaluation partitions the ELEMENTS in contiguous id ranges, the reference's own scheme
// (functional_binbounds, functional.c:863-879; there every thread fills a private full-size gradient matrix which is
// merged with matrix_axpy, functional.c:1092-1135).  Here every rank (one process per GPU) computes partial
// gradient rows for the vertices its elements touch; rows of vertices touched by several ranks have to be summed.
//
// No collective library call sits on this path.  Every rank owns one device block
//     [ flags: int[world] | recv[2][world][max_m][ncomp] doubles ]
// exported with cudaIpcGetMemHandle and mapped by all peers (NVLink / NVSwitch P2P):
//   k_halo_exchange (ONE launch per exchange)
//     push     each (peer q, k) pair stores this rank's partial row of the k-th vertex shared with q straight into
//              q's recv[parity][rank][k] (remote stores); the last CTA to finish fences and raises flags[rank] = epoch
//              on every peer;
//     combine  waits until all peers' flags reached the epoch, then sums, per shared vertex, the contributions of all
//              sharing ranks in ASCENDING RANK ORDER (own partial included at its place) and writes the completed row
//              back — the same order on every rank, so shared rows end up bitwise identical everywhere and the
//              result is deterministic.
// Two receive buffers alternate (parity = epoch & 1): a rank can be at most one exchange ahead of a peer, because
// finishing exchange e needs every peer's flag e, which a peer raises only after it finished combining e-1.
#include <cstdio>
#include <cstring>
#include <vector>

#include "mcu_host.h"
#include "mcu_patch.h"

extern long g_mcu_launches;

struct mcu_halo {
    int rank = 0, world = 1, ncomp = 0, n_shared = 0, max_m = 0, n_pairs = 0;
    int epoch = 0;
    // plan (device)
    int *d_shared_local = nullptr;   // n_shared local vertex ids (ascending global id)
    int *d_peer_ptr = nullptr;       // world+1 offsets into d_peer_pos
    int *d_peer_pos = nullptr;       // positions in the shared list, per peer, ascending global id
    int *d_contrib_ptr = nullptr;    // n_shared+1
    int *d_contrib_rank = nullptr;   // sharing ranks, ascending (own rank included)
    int *d_contrib_k = nullptr;      // index in the (rank, me) pair list; -1 for my own partial
    // the exported block and the peers' mappings
    unsigned char *d_block = nullptr;
    size_t block_bytes = 0, recv_off = 0;
    std::vector<void *> peer_block;  // world pointers (own block at [rank])
    void **d_peer_block = nullptr;   // device copy
    unsigned *d_counter = nullptr;
    unsigned *d_status = nullptr;    // bit 0: a peer did not show up in time
    // host copies of the pair lists (for mcu_halo_attach) and the fused mode switch
    std::vector<int> h_shared_local, h_peer_ptr, h_peer_pos;
    bool fused = false;              // pushes ride in the patch gradient kernel; the exchange kernel only combines
    unsigned pushed_mask = 0;        // fused mode: column blocks (slots) a gradient kernel pushed since the last exchange
    long long timeout_cycles = 240000000000LL;   // ~2 minutes at 1.9 GHz (option "halo_timeout_ms")
    bool failed = false;             // a peer timed out in an earlier exchange: reported by the next call
    mcu_mesh *mesh = nullptr;
};

#define HALO_MAXGRAD 4
struct HaloGrads {
    double *g[HALO_MAXGRAD];
};

// One launch per exchange.  Phase 1 (push): every (peer q, k) pair stores this rank's partial row of the k-th vertex
// shared with q straight into q's recv[parity][rank][k] (remote stores); one system-scope fence per CTA, and the last
// CTA to get there raises flags[rank] = epoch on every peer.  Phase 2 (combine): every CTA waits until all peers'
// flags reached the epoch — peers raise them independently of this grid, and the grid is small enough (<= 64 CTAs)
// to be resident as a whole — then sums its share of the shared vertices in ascending rank order.
__global__ void __launch_bounds__(256)
k_halo_exchange(HaloGrads grads, int ngrad, int dim, const int *__restrict__ shared_local,
                const int *__restrict__ peer_ptr, const int *__restrict__ peer_pos, void *const *__restrict__ peer_block,
                const int *__restrict__ contrib_ptr, const int *__restrict__ contrib_rank, const int *__restrict__ contrib_k,
                unsigned char *block, size_t recv_off, int rank, int world, int max_m, int ncomp, int parity, int n_pairs,
                int n_shared, unsigned *counter, unsigned *status, int epoch, long long timeout_cycles) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_pairs; t += gridDim.x * blockDim.x) {
        int q = 0;
        while (q + 1 < world && peer_ptr[q + 1] <= t) q++;
        int k = t - peer_ptr[q];
        int v = shared_local[peer_pos[t]];
        double *dst = (double *) ((unsigned char *) peer_block[q] + recv_off) +
                      (((size_t) parity * world + rank) * max_m + k) * ncomp;
        for (int gi = 0; gi < ngrad; gi++)
            for (int c = 0; c < dim; c++) dst[gi * dim + c] = grads.g[gi][(size_t) v * dim + c];   // remote stores
    }
    __syncthreads();
    __shared__ int ok;
    if (threadIdx.x == 0) {
        // n_pairs == 0 (fused mode): the rows were stored by the gradient kernels, which have completed; nothing of
        // this grid has to be ordered before the flags, so one CTA raises them right away.
        bool raise = (n_pairs == 0) ? (blockIdx.x == 0) : false;
        if (n_pairs > 0) {
            __threadfence_system();                    // cumulative: covers the block's stores ordered by the barrier
            unsigned done = atomicAdd(counter, 1u);
            if (done == gridDim.x - 1) { *counter = 0; raise = true; }
        }
        if (raise) {
            __threadfence_system();
            for (int q = 0; q < world; q++)
                if (q != rank) ((volatile int *) peer_block[q])[rank] = epoch;    // "my rows of exchange `epoch` have landed"
        }
        volatile int *flags = (volatile int *) block;
        int good = 1;
        long long t0 = clock64();
        for (int q = 0; q < world && good; q++) {
            if (q == rank) continue;
            while (flags[q] < epoch) {
                if (clock64() - t0 > timeout_cycles) { good = 0; break; }   // a peer is gone; do not hang the GPU
                __nanosleep(20);
            }
        }
        // acquire side: order the flag observation before the (cache-bypassing, ld.cv) reads of the received rows
        // that the other threads issue after the barrier below
        __threadfence_system();
        ok = good;
        if (!good) atomicOr(status, 1u);
    }
    __syncthreads();
    if (!ok) return;
    const double *recv = (const double *) (block + recv_off);
    const int ncols = ngrad * dim;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_shared * ncols; t += gridDim.x * blockDim.x) {
        int i = t / ncols, col = t - i * ncols, gi = col / dim, c = col - gi * dim;
        int v = shared_local[i];
        double own = grads.g[gi][(size_t) v * dim + c], sum = 0.0;
        for (int j = contrib_ptr[i]; j < contrib_ptr[i + 1]; j++) {
            int q = contrib_rank[j], k = contrib_k[j];
            double x = (k < 0) ? own : __ldcv(recv + (((size_t) parity * world + q) * max_m + k) * ncomp + col);
            sum = __dadd_rn(sum, x);                  // ascending rank order, identical on every sharing rank
        }
        grads.g[gi][(size_t) v * dim + c] = sum;
    }
}

static int *upload_ints(const int *h, size_t n) {
    int *d = nullptr;
    if (cudaMalloc((void **) &d, sizeof(int) * (n ? n : 1)) != cudaSuccess) return nullptr;
    if (n && cudaMemcpy(d, h, sizeof(int) * n, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
    return d;
}

extern "C" mcu_halo *mcu_halo_create(int rank, int world, int ncomp, int n_shared, const int *shared_local,
                                     const int *peer_ptr, const int *peer_pos, int max_m, const int *contrib_ptr,
                                     const int *contrib_rank, const int *contrib_k) {
    if (rank < 0 || world < 1 || rank >= world || ncomp < 1 || n_shared < 0 || max_m < 0) { mcu_set_error("bad halo arguments"); return nullptr; }
    mcu_halo *h = new mcu_halo();
    h->rank = rank; h->world = world; h->ncomp = ncomp; h->n_shared = n_shared; h->max_m = max_m;
    h->n_pairs = peer_ptr[world];
    h->h_shared_local.assign(shared_local, shared_local + n_shared);
    h->h_peer_ptr.assign(peer_ptr, peer_ptr + world + 1);
    h->h_peer_pos.assign(peer_pos, peer_pos + h->n_pairs);
    h->d_shared_local = upload_ints(shared_local, (size_t) n_shared);
    h->d_peer_ptr = upload_ints(peer_ptr, (size_t) world + 1);
    h->d_peer_pos = upload_ints(peer_pos, (size_t) h->n_pairs);
    h->d_contrib_ptr = upload_ints(contrib_ptr, (size_t) n_shared + 1);
    h->d_contrib_rank = upload_ints(contrib_rank, (size_t) contrib_ptr[n_shared]);
    h->d_contrib_k = upload_ints(contrib_k, (size_t) contrib_ptr[n_shared]);
    h->recv_off = 256;
    h->block_bytes = h->recv_off + sizeof(double) * 2 * (size_t) world * (size_t) (max_m ? max_m : 1) * ncomp;
    bool ok = h->d_shared_local && h->d_peer_ptr && h->d_peer_pos && h->d_contrib_ptr && h->d_contrib_rank && h->d_contrib_k &&
              cudaMalloc((void **) &h->d_block, h->block_bytes) == cudaSuccess &&
              cudaMemset(h->d_block, 0, h->block_bytes) == cudaSuccess &&
              cudaMalloc((void **) &h->d_counter, 64) == cudaSuccess && cudaMemset(h->d_counter, 0, 64) == cudaSuccess &&
              cudaMalloc((void **) &h->d_peer_block, sizeof(void *) * world) == cudaSuccess;
    if (!ok) { mcu_set_error("halo allocation failed: %s", cudaGetErrorString(cudaGetLastError())); delete h; return nullptr; }
    h->d_status = h->d_counter + 4;
    h->peer_block.assign((size_t) world, nullptr);
    h->peer_block[rank] = h->d_block;
    return h;
}

extern "C" int mcu_halo_export(mcu_halo *h, unsigned char handle[64]) {
    if (!h) return MCU_ERR_ARGS;
    cudaIpcMemHandle_t ipc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t e = cudaIpcGetMemHandle(&ipc, h->d_block);
    if (e != cudaSuccess) { mcu_set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return MCU_ERR_CUDA; }
    memcpy(handle, &ipc, 64);
    return MCU_OK;
}

extern "C" int mcu_halo_connect(mcu_halo *h, const unsigned char *handles) {
    if (!h || !handles) return MCU_ERR_ARGS;
    for (int q = 0; q < h->world; q++) {
        if (q == h->rank) continue;
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, handles + 64 * (size_t) q, 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { mcu_set_error("cudaIpcOpenMemHandle(rank %d): %s", q, cudaGetErrorString(e)); return MCU_ERR_CUDA; }
        h->peer_block[q] = p;
    }
    if (cudaMemcpy(h->d_peer_block, h->peer_block.data(), sizeof(void *) * h->world, cudaMemcpyHostToDevice) != cudaSuccess) return MCU_ERR_CUDA;
    return MCU_OK;
}

extern "C" int mcu_halo_exchange(mcu_halo *h, int ngrad, double *const *dev_grads, int dim) {
    if (!h || ngrad < 1 || ngrad > HALO_MAXGRAD || ngrad * dim > h->ncomp) { mcu_set_error("bad halo exchange arguments"); return MCU_ERR_ARGS; }
    if (h->world == 1) return MCU_OK;
    if (h->failed) { mcu_set_error("halo exchange: a peer did not arrive in time in an earlier exchange"); return MCU_ERR_CUDA; }
    // fused mode: the exchange kernel may skip its push phase only when a gradient kernel really pushed every column
    // block of this exchange (mcu_halo_push_params was taken for slots 0..ngrad-1 since the last exchange).  Any other
    // route to the gradients (gather path on small meshes, selections, other functionals, a slot out of range, a
    // patch set freed by mcu_mesh_set_connectivity) leaves rows un-pushed: push them here.
    const unsigned want = (ngrad >= 32) ? 0xffffffffu : ((1u << ngrad) - 1u);
    const bool all_pushed = h->fused && (h->pushed_mask & want) == want;
    h->pushed_mask = 0;
    {
        long ms = mcu_opt("halo_timeout_ms", 0);
        if (ms > 0) h->timeout_cycles = (long long) ms * 2000000LL;
    }
    HaloGrads g;
    for (int i = 0; i < HALO_MAXGRAD; i++) g.g[i] = i < ngrad ? dev_grads[i] : nullptr;
    cudaStream_t st = mcu_stream();
    h->epoch++;
    int parity = h->epoch & 1;
    int work = h->n_pairs > h->n_shared * ngrad * dim ? h->n_pairs : h->n_shared * ngrad * dim;
    int nb = (work + 255) / 256;
    nb = nb < 1 ? 1 : (nb > 64 ? 64 : nb);             // small enough to be resident as a whole (the CTAs wait for peers)
    k_halo_exchange<<<nb, 256, 0, st>>>(g, ngrad, dim, h->d_shared_local, h->d_peer_ptr, h->d_peer_pos, h->d_peer_block,
                                        h->d_contrib_ptr, h->d_contrib_rank, h->d_contrib_k, h->d_block, h->recv_off, h->rank,
                                        h->world, h->max_m, h->ncomp, parity, all_pushed ? 0 : h->n_pairs, h->n_shared, h->d_counter, h->d_status,
                                        h->epoch, h->timeout_cycles);
    g_mcu_launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { mcu_set_error("halo launch failed: %s", cudaGetErrorString(e)); return MCU_ERR_CUDA; }
    return MCU_OK;
}

/* 0 = all exchanges so far completed; 1 = a peer did not arrive in time (synchronises the stream) */
extern "C" int mcu_halo_status(mcu_halo *h) {
    if (!h) return MCU_ERR_ARGS;
    unsigned s = 0;
    if (cudaStreamSynchronize(mcu_stream()) != cudaSuccess) return MCU_ERR_CUDA;
    if (cudaMemcpy(&s, h->d_status, sizeof(unsigned), cudaMemcpyDeviceToHost) != cudaSuccess) return MCU_ERR_CUDA;
    if (s) h->failed = true;             // later exchanges fail loudly instead of summing incomplete rows
    return (int) s;
}

/* Fused mode: from now on the patch gradient kernel of `m` (triangles) stores the rows of shared vertices into the
 * peers' receive buffers itself, column block `slot` (mcu_halo_bind selects the slot before each evaluation), and
 * mcu_halo_exchange only raises the flags, waits and combines. */
extern "C" int mcu_halo_attach(mcu_halo *h, mcu_mesh *m) {
    if (!h || !m) return MCU_ERR_ARGS;
    int rc = mcu_ensure_patches(m, 2);
    if (rc != MCU_OK) { mcu_set_error("halo attach: the mesh has no patch decomposition"); return rc; }
    std::vector<int> vtx, peer, kk;
    for (int q = 0; q < h->world; q++)
        for (int t = h->h_peer_ptr[q]; t < h->h_peer_ptr[q + 1]; t++) {
            vtx.push_back(h->h_shared_local[h->h_peer_pos[t]]);
            peer.push_back(q);
            kk.push_back(t - h->h_peer_ptr[q]);
        }
    rc = mcu_patchset_attach_push(m->conn[2].patches, m->nv, (int) vtx.size(), vtx.data(), peer.data(), kk.data(), mcu_stream());
    if (rc != MCU_OK) { mcu_set_error("halo attach failed"); return rc; }
    h->fused = true; h->mesh = m;
    m->halo = h; m->halo_slot = 0;
    return MCU_OK;
}

extern "C" int mcu_halo_bind(mcu_halo *h, int slot) {
    if (!h || !h->mesh || slot < 0) return MCU_ERR_ARGS;
    h->mesh->halo_slot = slot;
    return MCU_OK;
}

bool mcu_halo_push_params(mcu_halo *h, int slot, int dim, McuHaloPush *out) {
    if (!h || !h->fused || (slot + 1) * dim > h->ncomp) return false;
    out->entries = nullptr;               // filled in by the launcher
    out->peer_block = (void *const *) h->d_peer_block;
    out->recv_off = h->recv_off;
    out->rank = h->rank; out->world = h->world; out->max_m = h->max_m; out->ncomp = h->ncomp;
    out->parity = (h->epoch + 1) & 1;     // the exchange that follows this evaluation
    out->col0 = slot * dim;
    if (slot < 32) h->pushed_mask |= 1u << slot;
    return true;
}

// the patch set a fused exchange pushed through is gone (mcu_mesh_set_connectivity / mcu_mesh_destroy)
void mcu_halo_detach_mesh(mcu_mesh *m) {
    if (!m || !m->halo) return;
    m->halo->fused = false; m->halo->mesh = nullptr; m->halo->pushed_mask = 0;
    m->halo = nullptr;
}

extern "C" int mcu_halo_destroy(mcu_halo *h) {
    if (!h) return MCU_OK;
    if (h->mesh && h->mesh->halo == h) h->mesh->halo = nullptr;
    cudaStreamSynchronize(mcu_stream());
    for (int q = 0; q < h->world; q++)
        if (q != h->rank && h->peer_block[q]) cudaIpcCloseMemHandle(h->peer_block[q]);
    cudaFree(h->d_shared_local); cudaFree(h->d_peer_ptr); cudaFree(h->d_peer_pos);
    cudaFree(h->d_contrib_ptr); cudaFree(h->d_contrib_rank); cudaFree(h->d_contrib_k);
    cudaFree(h->d_block); cudaFree(h->d_counter); cudaFree(h->d_peer_block);
    delete h;
    return MCU_OK;
}

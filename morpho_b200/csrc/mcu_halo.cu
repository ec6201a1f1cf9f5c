// mcu_halo.cu — completion of the gradient rows of vertices shared between GPUs, over NVLink peer memory.
//
// Multi-GPU evaluation partitions the ELEMENTS in contiguous id ranges, the reference's own scheme
// (functional_binbounds, functional.c:863-879; there every thread fills a private full-size gradient matrix which is
// merged with matrix_axpy, functional.c:1092-1135).  Here every rank (one process per GPU) computes partial
// gradient rows for the vertices its elements touch; rows of vertices touched by several ranks have to be summed.
//
// No collective library call sits on this path.  Every rank owns one device block
//     [ control words (MCU_HALO_CTRL_BYTES) | recv[2][world][max_m][ncomp] doubles ]
// exported with cudaIpcGetMemHandle and mapped by all peers (NVLink / NVSwitch P2P).  A row carries one COLUMN BLOCK
// ("slot") per gradient exchanged in a step (Area: columns 0..2, VolumeEnclosed: 3..5); every slot has its own flags and
// its own device-side epoch counter, so the exchange of one gradient is independent of the other's:
//   push     this rank's partial row of the k-th vertex shared with peer q is stored straight into q's
//            recv[parity][rank][k] (remote stores) — by the patch gradient kernel itself (fused mode: the patches that own
//            shared vertices are handed out FIRST, and the last of them to finish raises flags[slot][rank] = epoch on
//            every peer and the local "own rows written" word, long before the kernel ends), or by k_halo_slot when the
//            gradient came from another kernel;
//   combine  k_halo_slot waits until the peers' flags (and, fused, the local word) reached the epoch, then sums, per
//            shared vertex, the contributions of all sharing ranks in ASCENDING RANK ORDER (own partial included at its
//            place) and writes the completed row back — the same order on every rank, so shared rows end up bitwise
//            identical everywhere and the result is deterministic; the last CTA to finish bumps the slot's epoch.
// k_halo_slot is a grid of single-warp CTAs with <= 32 registers: such a CTA fits beside two resident CTAs of the
// persistent patch kernel on every SM, so on the library's SECOND STREAM (mcu_halo_exchange_async) the whole exchange of
// one gradient runs underneath the gradient kernels and costs the step only the final join.
// Two receive buffers alternate (parity = epoch & 1): a rank can be at most one exchange of a slot ahead of a peer,
// because finishing exchange e needs every peer's flag e, which a peer raises from its gradient kernel of step e, which is
// stream-ordered behind the peer's combine of e-1.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "mcu_host.h"
#include "mcu_patch.h"

extern long g_mcu_launches;

struct mcu_halo {
    int rank = 0, world = 1, ncomp = 0, n_shared = 0, max_m = 0, n_pairs = 0;
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
    unsigned *d_counter = nullptr;   // per slot: CTAs past the push phase [slot], CTAs finished [4 + slot]
    unsigned *d_status = nullptr;    // bit 0: a peer did not show up in time
    // host copies of the pair lists (for mcu_halo_attach) and the fused mode switch
    std::vector<int> h_shared_local, h_peer_ptr, h_peer_pos;
    bool fused = false;              // pushes ride in the patch gradient kernel; k_halo_slot only combines
    unsigned pushed_mask = 0;        // fused mode: column blocks (slots) a gradient kernel pushed since their last exchange
    long long timeout_cycles = 240000000000LL;   // ~2 minutes at 1.9 GHz (option "halo_timeout_ms")
    bool failed = false;             // a peer timed out in an earlier exchange: reported by the next call
    mcu_mesh *mesh = nullptr;
    // second stream: exchanges that run underneath the gradient kernels (mcu_halo_exchange_async / mcu_halo_join)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_main = nullptr, ev_side[MCU_HALO_MAXSLOT] = {nullptr, nullptr, nullptr, nullptr};
    bool side_pending[MCU_HALO_MAXSLOT] = {false, false, false, false};
};

extern "C" int mcu_halo_join(mcu_halo *h);
static int halo_join_slot(mcu_halo *h, int slot);

#define HALO_WARPS_MAX 296           // single-warp CTAs: two per SM
#define HALO_WARPS_ALONE 2048        // ... alone on the library stream
#define HALO_WARPS_OVERTAKING 48     // ... when the launch can get in front of the kernel whose rows it waits for (see halo_slot_launch)

// One launch per (exchange, slot).  Single-warp CTAs.
__global__ void __maxnreg__(32)
k_halo_slot(double *__restrict__ grad, double *__restrict__ grad2, int nslots, int dim, int slot, const int *__restrict__ shared_local,
            const int *__restrict__ peer_ptr, const int *__restrict__ peer_pos, void *const *__restrict__ peer_block,
            const int *__restrict__ contrib_ptr, const int *__restrict__ contrib_rank, const int *__restrict__ contrib_k,
            unsigned char *block, size_t recv_off, int rank, int world, int max_m, int ncomp, int n_pairs,
            int n_shared, unsigned *counter, unsigned *status, int wait_own, long long timeout_cycles) {
    volatile int *ctrl = (volatile int *) block;
    // One launch completes `nslots` (1 or 2) consecutive column blocks: grad -> block `slot`, grad2 -> block `slot + 1`.
    // Flags and the wait use the FIRST block's epoch; every block keeps its own epoch counter (and with it its own buffer
    // parity), bumped together at the end, so merged and single exchanges of a block can follow each other in any order.
    // The epochs cannot move while any CTA of this launch is alive: they are bumped by the LAST CTA to finish.
    const int epoch = ctrl[MCU_HALO_EPOCH(slot)] + 1;
    const int parity0 = epoch & 1;
    const int parity1 = nslots > 1 ? ((ctrl[MCU_HALO_EPOCH(slot + 1)] + 1) & 1) : 0;
    const int lane = threadIdx.x;
    if (n_pairs > 0) {
        // push phase (the gradient did not come from the fused patch kernel): this launch is stream-ordered behind it
        for (int t = blockIdx.x * 32 + lane; t < n_pairs * nslots; t += gridDim.x * 32) {
            const int sl = t >= n_pairs ? 1 : 0, tp = t - sl * n_pairs;
            int q = 0;
            while (q + 1 < world && peer_ptr[q + 1] <= tp) q++;
            const int k = tp - peer_ptr[q];
            const int v = shared_local[peer_pos[tp]];
            const double *src = (sl ? grad2 : grad) + (size_t) v * dim;
            double *dst = (double *) ((unsigned char *) peer_block[q] + recv_off) +
                          (((size_t) (sl ? parity1 : parity0) * world + rank) * max_m + k) * ncomp + (slot + sl) * dim;
            for (int c = 0; c < dim; c++) dst[c] = src[c];                       // remote stores
        }
        __threadfence_system();
        __syncwarp();
        if (lane == 0 && atomicAdd(counter + slot, 1u) == gridDim.x - 1) {
            counter[slot] = 0;
            __threadfence_system();
            for (int q = 0; q < world; q++)
                if (q != rank) ((volatile int *) peer_block[q])[MCU_HALO_FLAG(slot, rank)] = epoch;    // "my rows of exchange `epoch` have landed"
        }
    }
    int good = 1;
    if (lane == 0) {
        const long long t0 = clock64();
        if (wait_own) while (ctrl[MCU_HALO_OWNDONE(slot)] < epoch) { if (clock64() - t0 > timeout_cycles) { good = 0; break; } __nanosleep(40); }
        for (int q = 0; q < world && good; q++) {
            if (q == rank || peer_ptr[q + 1] == peer_ptr[q]) continue;        // nothing shared with q
            while (ctrl[MCU_HALO_FLAG(slot, q)] < epoch) {
                if (clock64() - t0 > timeout_cycles) { good = 0; break; }      // a peer is gone; do not hang the GPU
                __nanosleep(40);
            }
        }
        // acquire side: order the flag observations before the (cache-bypassing) reads of the rows below
        __threadfence_system();
        if (!good) atomicOr(status, 1u);
    }
    good = __shfl_sync(0xffffffffu, good, 0);
    if (good) {
        const double *recv = (const double *) (block + recv_off);
        const int per = n_shared * dim;
        for (int t = blockIdx.x * 32 + lane; t < per * nslots; t += gridDim.x * 32) {
            const int sl = t >= per ? 1 : 0, tt = t - sl * per;
            const int i = tt / dim, c = tt - i * dim;
            const int v = shared_local[i];
            double *g = (sl ? grad2 : grad) + (size_t) v * dim + c;
            const int parity = sl ? parity1 : parity0, col0 = (slot + sl) * dim;
            const double own = __ldcv(g);
            double sum = 0.0;
            for (int j = contrib_ptr[i]; j < contrib_ptr[i + 1]; j++) {
                const int q = contrib_rank[j], k = contrib_k[j];
                const double x = (k < 0) ? own : __ldcv(recv + (((size_t) parity * world + q) * max_m + k) * ncomp + col0 + c);
                sum = __dadd_rn(sum, x);                  // ascending rank order, identical on every sharing rank
            }
            *g = sum;
        }
    }
    __threadfence();
    __syncwarp();
    if (lane == 0 && atomicAdd(counter + 4 + slot, 1u) == gridDim.x - 1) {
        counter[4 + slot] = 0;
        __threadfence();
        if (nslots > 1) ctrl[MCU_HALO_EPOCH(slot + 1)] = ctrl[MCU_HALO_EPOCH(slot + 1)] + 1;
        ctrl[MCU_HALO_EPOCH(slot)] = epoch;               // this exchange of the column block(s) is complete
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
    if (world > MCU_HALO_MAXWORLD) { mcu_set_error("halo: at most %d ranks", MCU_HALO_MAXWORLD); delete h; return nullptr; }
    h->recv_off = MCU_HALO_CTRL_BYTES;
    h->block_bytes = h->recv_off + sizeof(double) * 2 * (size_t) world * (size_t) (max_m ? max_m : 1) * ncomp;
    bool ok = h->d_shared_local && h->d_peer_ptr && h->d_peer_pos && h->d_contrib_ptr && h->d_contrib_rank && h->d_contrib_k &&
              cudaMalloc((void **) &h->d_block, h->block_bytes) == cudaSuccess &&
              cudaMemset(h->d_block, 0, h->block_bytes) == cudaSuccess &&
              cudaMalloc((void **) &h->d_counter, 64) == cudaSuccess && cudaMemset(h->d_counter, 0, 64) == cudaSuccess &&
              cudaMalloc((void **) &h->d_peer_block, sizeof(void *) * world) == cudaSuccess;
    if (!ok) { mcu_set_error("halo allocation failed: %s", cudaGetErrorString(cudaGetLastError())); delete h; return nullptr; }
    h->d_status = h->d_counter + 12;
    if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_side[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_side[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_side[2], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_side[3], cudaEventDisableTiming) != cudaSuccess) {
        mcu_set_error("halo: stream / event creation failed"); delete h; return nullptr;
    }
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

static int halo_slot_launch(mcu_halo *h, int slot, double *dev_grad, int dim, cudaStream_t st, double *dev_grad2 = nullptr) {
    const int nslots = dev_grad2 ? 2 : 1;
    const bool pushed = h->fused && slot < 32 && (h->pushed_mask & (1u << slot));
    if (slot < 32) h->pushed_mask &= ~(1u << slot);
    if (nslots > 1 && slot + 1 < 32) h->pushed_mask &= ~(1u << (slot + 1));
    {
        long ms = mcu_opt("halo_timeout_ms", 0);
        if (ms > 0) h->timeout_cycles = (long long) ms * 2000000LL;
    }
    const int work = nslots * std::max(pushed ? 0 : h->n_pairs, h->n_shared * dim);
    int nb = (work + 31) / 32;
    // A launch that may OVERTAKE the gradient kernel it waits for (fused pushes, second stream) must leave most SMs
    // alone: its CTAs spin until the gradient kernel has written the own rows, and an SM that hosts a spinning CTA
    // cannot take a CTA of the patch kernel when the two kernels ask for different shared-memory carve-outs (the
    // patch kernel needs the largest one) — with a spinning CTA on every SM the gradient kernel never starts and both
    // sit there until the time-out (seen at N = 2 with >= 1500 shared vertices: 199-296 CTAs on 148 SMs).
    // On the second stream the launch shares the SMs with the next gradient kernel (VolumeEnclosed 77 -> 88 us beside 296
    // of these warps at N = 2).  Fewer warps are NOT the answer: the loops below are chains of dependent loads, and with
    // 32 warps the exchange of 11 414 shared rows at N = 8 took ~100 us per gradient instead of ~30
    // (profiles/r2_bench_n8_sidecap32.json against profiles/r2_bench_n2.json).
    // On the library stream nothing else runs beside the exchange: as many warps as there are rows (one item per lane
    // and phase — with 296 warps the merged exchange of 2 x 10 000 rows took 47 us, all of it load latency).  Every CTA
    // of a launch has to be resident at once (a CTA waits for the peers' flags, and those are raised only after ALL CTAs
    // of the peer's launch have pushed): 2048 single-warp CTAs are 14 per SM, under the 32 an SM holds.
    const int cap = pushed ? HALO_WARPS_OVERTAKING : (st == h->side ? HALO_WARPS_MAX : HALO_WARPS_ALONE);
    nb = nb < 1 ? 1 : (nb > cap ? cap : nb);
    static bool carveout_set = false;
    if (!carveout_set) {               // the same carve-out as the patch kernel: CTAs of both can share an SM
        cudaFuncSetAttribute(k_halo_slot, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carveout_set = true;
    }
    // rows the fused patch kernel pushed: only wait and combine (and wait for the local rows, the launch may overtake
    // the gradient kernel); any other route to the gradient (gather kernel on small meshes, selections, other
    // functionals, a patch set freed by mcu_mesh_set_connectivity) leaves rows un-pushed: push them here
    k_halo_slot<<<nb, 32, 0, st>>>(dev_grad, dev_grad2, nslots, dim, slot, h->d_shared_local, h->d_peer_ptr, h->d_peer_pos, h->d_peer_block,
                                   h->d_contrib_ptr, h->d_contrib_rank, h->d_contrib_k, h->d_block, h->recv_off, h->rank, h->world,
                                   h->max_m, h->ncomp, pushed ? 0 : h->n_pairs, h->n_shared, h->d_counter, h->d_status, pushed ? 1 : 0,
                                   h->timeout_cycles);
    g_mcu_launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { mcu_set_error("halo launch failed: %s", cudaGetErrorString(e)); return MCU_ERR_CUDA; }
    return MCU_OK;
}

static int halo_check(mcu_halo *h, int ngrad, int dim) {
    if (!h || ngrad < 1 || ngrad > MCU_HALO_MAXSLOT || ngrad * dim > h->ncomp) { mcu_set_error("bad halo exchange arguments"); return MCU_ERR_ARGS; }
    if (h->failed) { mcu_set_error("halo exchange: a peer did not arrive in time in an earlier exchange"); return MCU_ERR_CUDA; }
    return MCU_OK;
}

/* complete the shared rows of ngrad gradients (column blocks 0 .. ngrad-1) on the library stream */
extern "C" int mcu_halo_exchange(mcu_halo *h, int ngrad, double *const *dev_grads, int dim) {
    int rc = halo_check(h, ngrad, dim);
    if (rc) return rc;
    if (h->world == 1 || h->n_shared == 0) return MCU_OK;
    rc = mcu_halo_join(h);
    // Two gradients whose rows all have to be pushed here go in ONE launch: one flag round trip instead of two
    // ("halo_merge" = 0 keeps one launch per gradient).  Rows the fused patch kernel already pushed keep the
    // per-gradient launches (their flags were raised per column block).
    const bool any_pushed = h->fused && (h->pushed_mask & ((1u << ngrad) - 1u));
    const bool merge = mcu_opt("halo_merge", 1) != 0 && !any_pushed;
    for (int s = 0; s < ngrad && !rc;) {
        if (merge && s + 1 < ngrad) { rc = halo_slot_launch(h, s, dev_grads[s], dim, mcu_stream(), dev_grads[s + 1]); s += 2; }
        else { rc = halo_slot_launch(h, s, dev_grads[s], dim, mcu_stream()); s += 1; }
    }
    return rc;
}

/* The same for ONE gradient (column block `slot`), on the library's second stream: called right after the evaluation
 * that produced dev_grad was queued, the exchange runs underneath that kernel and whatever is queued next on the
 * library stream.  mcu_halo_join makes the library stream wait for the exchanges in flight (it is called implicitly by
 * the next evaluation on the attached mesh and by mcu_halo_exchange). */
extern "C" int mcu_halo_exchange_async(mcu_halo *h, int slot, double *dev_grad, int dim) {
    int rc = halo_check(h, slot + 1, dim);
    if (rc) return rc;
    if (slot < 0 || !dev_grad) return MCU_ERR_ARGS;
    if (h->world == 1 || h->n_shared == 0) return MCU_OK;
    rc = halo_join_slot(h, slot);
    if (rc) return rc;
    const bool pushed = h->fused && slot < 32 && (h->pushed_mask & (1u << slot));
    if (!pushed) {
        // the rows have to be read from the finished gradient: order the second stream behind the library stream
        if (cudaEventRecord(h->ev_main, mcu_stream()) != cudaSuccess || cudaStreamWaitEvent(h->side, h->ev_main, 0) != cudaSuccess) return MCU_ERR_CUDA;
    }
    rc = halo_slot_launch(h, slot, dev_grad, dim, h->side);
    if (rc == MCU_OK && cudaEventRecord(h->ev_side[slot], h->side) != cudaSuccess) rc = MCU_ERR_CUDA;
    h->side_pending[slot] = true;
    return rc;
}

/* the library stream waits for the exchange of `slot` that is in flight on the second stream */
static int halo_join_slot(mcu_halo *h, int slot) {
    if (!h->side_pending[slot]) return MCU_OK;
    if (cudaStreamWaitEvent(mcu_stream(), h->ev_side[slot], 0) != cudaSuccess) {
        mcu_set_error("halo join failed: %s", cudaGetErrorString(cudaGetLastError()));
        return MCU_ERR_CUDA;
    }
    h->side_pending[slot] = false;
    return MCU_OK;
}

extern "C" int mcu_halo_join(mcu_halo *h) {
    if (!h) return MCU_ERR_ARGS;
    int rc = MCU_OK;
    for (int s = 0; s < MCU_HALO_MAXSLOT && !rc; s++) rc = halo_join_slot(h, s);
    return rc;
}

/* 0 = all exchanges so far completed; 1 = a peer did not arrive in time (synchronises the stream) */
extern "C" int mcu_halo_status(mcu_halo *h) {
    if (!h) return MCU_ERR_ARGS;
    unsigned s = 0;
    if (mcu_halo_join(h) != MCU_OK || cudaStreamSynchronize(mcu_stream()) != cudaSuccess) return MCU_ERR_CUDA;
    if (cudaMemcpy(&s, h->d_status, sizeof(unsigned), cudaMemcpyDeviceToHost) != cudaSuccess) return MCU_ERR_CUDA;
    if (s) {
        h->failed = true;                // later exchanges fail loudly instead of summing incomplete rows
        // what every rank was waiting for: per slot the completed-exchange count, the own-rows word and the peers' flags
        int ctrl[72];
        if (cudaMemcpy(ctrl, h->d_block, sizeof(ctrl), cudaMemcpyDeviceToHost) == cudaSuccess) {
            for (int slot = 0; slot < MCU_HALO_MAXSLOT; slot++) {
                fprintf(stderr, "[mcu_halo] rank %d slot %d: exchanges completed %d, own rows at %d, peer flags", h->rank, slot,
                        ctrl[MCU_HALO_EPOCH(slot)], ctrl[MCU_HALO_OWNDONE(slot)]);
                for (int q = 0; q < h->world; q++) fprintf(stderr, " %d", ctrl[MCU_HALO_FLAG(slot, q)]);
                fprintf(stderr, "\n");
            }
        }
    }
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
    out->slot = slot;
    out->col0 = slot * dim;
    if (slot < MCU_HALO_MAXSLOT) halo_join_slot(h, slot);   // the previous exchange of this column block has to be through
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
    if (h->side) cudaStreamSynchronize(h->side);
    cudaStreamSynchronize(mcu_stream());
    if (h->side) cudaStreamDestroy(h->side);
    if (h->ev_main) cudaEventDestroy(h->ev_main);
    for (int i = 0; i < MCU_HALO_MAXSLOT; i++) if (h->ev_side[i]) cudaEventDestroy(h->ev_side[i]);
    for (int q = 0; q < h->world; q++)
        if (q != h->rank && h->peer_block[q]) cudaIpcCloseMemHandle(h->peer_block[q]);
    cudaFree(h->d_shared_local); cudaFree(h->d_peer_ptr); cudaFree(h->d_peer_pos);
    cudaFree(h->d_contrib_ptr); cudaFree(h->d_contrib_rank); cudaFree(h->d_contrib_k);
    cudaFree(h->d_block); cudaFree(h->d_counter); cudaFree(h->d_peer_block);
    delete h;
    return MCU_OK;
}

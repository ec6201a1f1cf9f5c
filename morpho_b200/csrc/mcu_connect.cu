// mcu_connect.cu — connectivity construction on the device: Mesh.addgrade.
//
// mesh_addgrade (mesh.c:419-489) walks the elements of the next higher grade in id order, generates for each the
// (g+1)-subsets of its vertices in lexicographic counter order and keeps the subsets not seen before (an n-tuple hash,
// ~4 us per entry on the host): the new element ids are the order of FIRST DISCOVERY.  Here:
//   1. one thread per candidate q = element * ncomb + combination writes its tuple as sort keys;
//   2. a stable radix sort (CUB) orders the candidates by tuple, equal tuples staying in ascending q;
//   3. the head of every run of equal tuples is the first discovery: flag first[q];
//   4. an exclusive scan of the flags in q order gives the new element id; a last kernel emits the tuples.
// The result is bit-identical to the reference's (and the oracle's mo_edges_from).  Tuples of 2 ids use one 64-bit
// key; tuples of 3 ids (faces of tetrahedra) two stable passes (last id first).
#include <cub/cub.cuh>
#include <algorithm>
#include <vector>

#include "mcu_host.h"

extern long g_mcu_launches;

struct Combos {
    int n, ncomb, nvper;
    int idx[6][3];
};

__global__ void k_candidates(Combos cb, int nel, const int *__restrict__ rix, unsigned long long *__restrict__ key_ab,
                             unsigned *__restrict__ key_c, unsigned *__restrict__ val) {
    long q = (long) blockIdx.x * blockDim.x + threadIdx.x;
    long ncand = (long) nel * cb.ncomb;
    if (q >= ncand) return;
    int e = (int) (q / cb.ncomb), c = (int) (q - (long) e * cb.ncomb);
    const int *ids = rix + (size_t) e * cb.nvper;
    unsigned long long a = (unsigned) ids[cb.idx[c][0]], b = (unsigned) ids[cb.idx[c][1]];
    key_ab[q] = (a << 32) | b;
    if (cb.n == 3) key_c[q] = (unsigned) ids[cb.idx[c][2]];
    val[q] = (unsigned) q;
}

__global__ void k_gather_keys(long n, const unsigned *__restrict__ val, const unsigned long long *__restrict__ src, unsigned long long *__restrict__ dst) {
    long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[val[i]];
}

// sorted position i is a run head when its tuple differs from the one before it
__global__ void k_heads(long n, int tuple, const unsigned long long *__restrict__ key_sorted, const unsigned *__restrict__ val_sorted,
                        const unsigned *__restrict__ key_c, unsigned *__restrict__ first) {
    long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool head = (i == 0) || key_sorted[i] != key_sorted[i - 1];
    if (!head && tuple == 3) head = key_c[val_sorted[i]] != key_c[val_sorted[i - 1]];
    if (head) first[val_sorted[i]] = 1u;
}

__global__ void k_emit(Combos cb, int nel, const int *__restrict__ rix, const unsigned *__restrict__ first, const unsigned *__restrict__ newid, int *__restrict__ out) {
    long q = (long) blockIdx.x * blockDim.x + threadIdx.x;
    long ncand = (long) nel * cb.ncomb;
    if (q >= ncand || !first[q]) return;
    int e = (int) (q / cb.ncomb), c = (int) (q - (long) e * cb.ncomb);
    const int *ids = rix + (size_t) e * cb.nvper;
    int *o = out + (size_t) newid[q] * cb.n;
    for (int k = 0; k < cb.n; k++) o[k] = ids[cb.idx[c][k]];
}

#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { mcu_set_error("%s: %s", #call, cudaGetErrorString(e_)); rc = MCU_ERR_CUDA; goto done; } } while (0)

extern "C" int mcu_mesh_addgrade(mcu_mesh *m, int g, int *nel_out) {
    if (!m || g < 1 || g > 2) { mcu_set_error("addgrade: grade must be 1 or 2"); return MCU_ERR_ARGS; }
    if (m->conn[g].present) { if (nel_out) *nel_out = m->conn[g].nel; return MCU_OK; }
    int h = 0;
    for (int k = g + 1; k <= 3 && !h; k++) if (m->conn[k].present) h = k;
    if (!h) { mcu_set_error("addgrade: no higher grade present"); return MCU_ERR_ELNOTFOUND; }
    const McuConn &src = m->conn[h];
    Combos cb;
    cb.n = g + 1; cb.nvper = h + 1; cb.ncomb = 0;
    // (g+1)-subsets of {0..h} in lexicographic counter order (mesh.c:450-470)
    if (cb.n == 2) { for (int a = 0; a <= h; a++) for (int b = a + 1; b <= h; b++) { cb.idx[cb.ncomb][0] = a; cb.idx[cb.ncomb][1] = b; cb.idx[cb.ncomb][2] = 0; cb.ncomb++; } }
    else { for (int a = 0; a <= h; a++) for (int b = a + 1; b <= h; b++) for (int c = b + 1; c <= h; c++) { cb.idx[cb.ncomb][0] = a; cb.idx[cb.ncomb][1] = b; cb.idx[cb.ncomb][2] = c; cb.ncomb++; } }
    const long ncand = (long) src.nel * cb.ncomb;
    if (ncand >= 4294967295L) { mcu_set_error("addgrade: too many candidates"); return MCU_ERR_UNSUPPORTED; }
    cudaStream_t st = mcu_stream();
    int rc = MCU_OK;
    unsigned long long *key[2] = {nullptr, nullptr};
    unsigned *val[2] = {nullptr, nullptr}, *key_c = nullptr, *key_c2 = nullptr, *first = nullptr, *newid = nullptr;
    void *tmp = nullptr;
    int *d_out = nullptr;
    size_t tmp_bytes = 0, t1 = 0, t2 = 0, t3 = 0;
    int blocks = (int) ((ncand + 255) / 256);
    unsigned last_first = 0, last_id = 0;
    int nnew = 0;
    std::vector<int> h_out;
    if (ncand == 0) { nnew = 0; goto publish; }
    CKC(cudaMalloc((void **) &key[0], 8 * ncand)); CKC(cudaMalloc((void **) &key[1], 8 * ncand));
    CKC(cudaMalloc((void **) &val[0], 4 * ncand)); CKC(cudaMalloc((void **) &val[1], 4 * ncand));
    CKC(cudaMalloc((void **) &key_c, 4 * ncand)); CKC(cudaMalloc((void **) &key_c2, 4 * ncand));
    CKC(cudaMalloc((void **) &first, 4 * ncand)); CKC(cudaMalloc((void **) &newid, 4 * ncand));
    CKC(cub::DeviceRadixSort::SortPairs(nullptr, t1, key[0], key[1], val[0], val[1], (int) ncand, 0, 64, st));
    CKC(cub::DeviceRadixSort::SortPairs(nullptr, t2, key_c, key_c2, val[0], val[1], (int) ncand, 0, 32, st));
    CKC(cub::DeviceScan::ExclusiveSum(nullptr, t3, first, newid, (int) ncand, st));
    tmp_bytes = std::max(t1, std::max(t2, t3));
    CKC(cudaMalloc(&tmp, tmp_bytes));
    k_candidates<<<blocks, 256, 0, st>>>(cb, src.nel, src.d_rix, key[0], key_c, val[0]);
    if (cb.n == 3) {
        // least significant id first, then a stable pass on the leading pair
        CKC(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_c, key_c2, val[0], val[1], (int) ncand, 0, 32, st));
        k_gather_keys<<<blocks, 256, 0, st>>>(ncand, val[1], key[0], key[1]);
        CKC(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key[1], key[0], val[1], val[0], (int) ncand, 0, 64, st));
        // sorted keys in key[0], sorted candidates in val[0]; key_c still indexed by candidate
        CKC(cudaMemsetAsync(first, 0, 4 * ncand, st));
        k_heads<<<blocks, 256, 0, st>>>(ncand, 3, key[0], val[0], key_c, first);
    } else {
        CKC(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key[0], key[1], val[0], val[1], (int) ncand, 0, 64, st));
        CKC(cudaMemsetAsync(first, 0, 4 * ncand, st));
        k_heads<<<blocks, 256, 0, st>>>(ncand, 2, key[1], val[1], key_c, first);
    }
    CKC(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, first, newid, (int) ncand, st));
    CKC(cudaMemcpyAsync(&last_first, first + (ncand - 1), 4, cudaMemcpyDeviceToHost, st));
    CKC(cudaMemcpyAsync(&last_id, newid + (ncand - 1), 4, cudaMemcpyDeviceToHost, st));
    CKC(cudaStreamSynchronize(st));
    nnew = (int) (last_id + last_first);
    CKC(cudaMalloc((void **) &d_out, sizeof(int) * (size_t) std::max(nnew, 1) * cb.n));
    k_emit<<<blocks, 256, 0, st>>>(cb, src.nel, src.d_rix, first, newid, d_out);
    g_mcu_launches += 4;
    h_out.resize((size_t) nnew * cb.n);
    CKC(cudaMemcpyAsync(h_out.data(), d_out, sizeof(int) * h_out.size(), cudaMemcpyDeviceToHost, st));
    CKC(cudaStreamSynchronize(st));
publish:
    {
        McuConn &c = m->conn[g];
        c.g = g; c.nvper = g + 1; c.nel = nnew;
        c.h_rix.swap(h_out);
        c.d_rix = d_out; d_out = nullptr;
        if (!c.d_rix) cudaMalloc((void **) &c.d_rix, 16);
        c.present = true;
        m->conn_version++;
        if (nel_out) *nel_out = nnew;
    }
done:
    for (int i = 0; i < 2; i++) { if (key[i]) cudaFree(key[i]); if (val[i]) cudaFree(val[i]); }
    if (key_c) cudaFree(key_c);
    if (key_c2) cudaFree(key_c2);
    if (first) cudaFree(first);
    if (newid) cudaFree(newid);
    if (tmp) cudaFree(tmp);
    if (d_out) cudaFree(d_out);
    return rc;
}

/* host copy of the incidence of grade g (nel * (g+1) ints, element ids in the reference's order) */
extern "C" int mcu_mesh_get_connectivity(mcu_mesh *m, int g, int *host_rix) {
    if (!m || g < 1 || g > 3 || !m->conn[g].present || !host_rix) { mcu_set_error("get_connectivity: grade not present"); return MCU_ERR_ARGS; }
    const McuConn &c = m->conn[g];
    std::copy(c.h_rix.begin(), c.h_rix.end(), host_rix);
    return MCU_OK;
}

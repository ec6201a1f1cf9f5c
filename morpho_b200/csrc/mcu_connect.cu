// mcu_connect.cu — connectivity construction on the device: Mesh.addgrade, the boundary selection, grade-lowering matrices.
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

// ---------------------------------------------------------------------------------
// Boundary selection (selection_selectboundary, selection.c:211-232): the elements of grade max-1 contained in exactly
// one element of the top grade.  The reference builds the raise matrix conn(max, max-1) with hash lookups; here the
// facets of the top elements are generated as sort keys (the same candidates as addgrade), sorted by their leading
// pair of ids, and every existing element of grade max-1 counts its occurrences with a binary search (+ a scan of
// the short run of equal leading pairs for the third id of a triangular face).
// ---------------------------------------------------------------------------------
__global__ void k_count_containing(int nsub, int tuple, const int *__restrict__ sub_rix, long ncand,
                                   const unsigned long long *__restrict__ key_sorted, const unsigned *__restrict__ val_sorted,
                                   const unsigned *__restrict__ key_c, unsigned *__restrict__ flag) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nsub) return;
    const int *ids = sub_rix + (size_t) e * tuple;
    const unsigned long long key = ((unsigned long long) (unsigned) ids[0] << 32) | (unsigned) ids[1];
    long lo = 0, hi = ncand;
    while (lo < hi) { long mid = (lo + hi) >> 1; if (key_sorted[mid] < key) lo = mid + 1; else hi = mid; }
    int count = 0;
    for (long i = lo; i < ncand && key_sorted[i] == key; i++)
        if (tuple == 2 || key_c[val_sorted[i]] == (unsigned) ids[2]) count++;
    flag[e] = (count == 1) ? 1u : 0u;
}

__global__ void k_vertex_degree_one(int nv, const int *__restrict__ tptr, unsigned *__restrict__ flag) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nv) flag[v] = (tptr[v + 1] - tptr[v] == 1) ? 1u : 0u;
}

__global__ void k_compact(int n, const unsigned *__restrict__ flag, const unsigned *__restrict__ pos, int *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[pos[i]] = i;
}

/* ids (ascending) of the boundary elements of grade maxgrade-1; *grade_out = that grade.  cap = room in host_ids;
 * returns MCU_OK and the count in *n_out (ids beyond cap are dropped, the count is still the full one). */
extern "C" int mcu_mesh_boundary(mcu_mesh *m, int *grade_out, int *n_out, int *host_ids, int cap) {
    if (!m || !n_out) { mcu_set_error("boundary: null argument"); return MCU_ERR_ARGS; }
    int top = 0;
    for (int g = m->dim; g > 0 && !top; g--) if (m->conn[g].present) top = g;
    if (top < 1) { mcu_set_error("boundary: the mesh has no elements"); return MCU_ERR_ELNOTFOUND; }       // SELECTION_BND
    const int bnd = top - 1;
    if (grade_out) *grade_out = bnd;
    cudaStream_t st = mcu_stream();
    int rc = MCU_OK, n_items = 0, nsel = 0;
    unsigned *flag = nullptr, *pos = nullptr, *val[2] = {nullptr, nullptr}, *key_c = nullptr;
    unsigned long long *key[2] = {nullptr, nullptr};
    void *tmp = nullptr;
    int *d_ids = nullptr;
    size_t t1 = 0, t2 = 0;
    unsigned last_flag = 0, last_pos = 0;
    if (bnd == 0) {
        rc = mcu_ensure_transpose(m, 1);
        if (rc) return rc;
        n_items = m->nv;
        CKC(cudaMalloc((void **) &flag, 4 * (size_t) std::max(n_items, 1))); CKC(cudaMalloc((void **) &pos, 4 * (size_t) std::max(n_items, 1)));
        if (n_items > 0) k_vertex_degree_one<<<(n_items + 255) / 256, 256, 0, st>>>(n_items, m->conn[1].d_tptr, flag);
    } else {
        if (!m->conn[bnd].present) { mcu_set_error("boundary: the mesh has no elements of grade %d (addgrade first)", bnd); return MCU_ERR_ELNOTFOUND; }
        const McuConn &src = m->conn[top], &sub = m->conn[bnd];
        Combos cb;
        cb.n = bnd + 1; cb.nvper = top + 1; cb.ncomb = 0;
        if (cb.n == 2) { for (int a = 0; a <= top; a++) for (int b = a + 1; b <= top; b++) { cb.idx[cb.ncomb][0] = a; cb.idx[cb.ncomb][1] = b; cb.idx[cb.ncomb][2] = 0; cb.ncomb++; } }
        else { for (int a = 0; a <= top; a++) for (int b = a + 1; b <= top; b++) for (int c = b + 1; c <= top; c++) { cb.idx[cb.ncomb][0] = a; cb.idx[cb.ncomb][1] = b; cb.idx[cb.ncomb][2] = c; cb.ncomb++; } }
        const long ncand = (long) src.nel * cb.ncomb;
        if (ncand >= 4294967295L) { mcu_set_error("boundary: too many candidates"); return MCU_ERR_UNSUPPORTED; }
        n_items = sub.nel;
        const size_t nc = (size_t) std::max<long>(ncand, 1);
        CKC(cudaMalloc((void **) &key[0], 8 * nc)); CKC(cudaMalloc((void **) &key[1], 8 * nc));
        CKC(cudaMalloc((void **) &val[0], 4 * nc)); CKC(cudaMalloc((void **) &val[1], 4 * nc));
        CKC(cudaMalloc((void **) &key_c, 4 * nc));
        CKC(cudaMalloc((void **) &flag, 4 * (size_t) std::max(n_items, 1))); CKC(cudaMalloc((void **) &pos, 4 * (size_t) std::max(n_items, 1)));
        if (ncand > 0) {
            CKC(cub::DeviceRadixSort::SortPairs(nullptr, t1, key[0], key[1], val[0], val[1], (int) ncand, 0, 64, st));
            CKC(cudaMalloc(&tmp, std::max<size_t>(t1, 16)));
            k_candidates<<<(int) ((ncand + 255) / 256), 256, 0, st>>>(cb, src.nel, src.d_rix, key[0], key_c, val[0]);
            CKC(cub::DeviceRadixSort::SortPairs(tmp, t1, key[0], key[1], val[0], val[1], (int) ncand, 0, 64, st));
            cudaFree(tmp); tmp = nullptr;
        }
        if (n_items > 0) k_count_containing<<<(n_items + 255) / 256, 256, 0, st>>>(n_items, cb.n, sub.d_rix, ncand, key[1], val[1], key_c, flag);
        g_mcu_launches += 2;
    }
    if (n_items > 0) {
        CKC(cub::DeviceScan::ExclusiveSum(nullptr, t2, flag, pos, n_items, st));
        CKC(cudaMalloc(&tmp, std::max<size_t>(t2, 16)));
        CKC(cub::DeviceScan::ExclusiveSum(tmp, t2, flag, pos, n_items, st));
        CKC(cudaMemcpyAsync(&last_flag, flag + (n_items - 1), 4, cudaMemcpyDeviceToHost, st));
        CKC(cudaMemcpyAsync(&last_pos, pos + (n_items - 1), 4, cudaMemcpyDeviceToHost, st));
        CKC(cudaStreamSynchronize(st));
        nsel = (int) (last_flag + last_pos);
        if (nsel > 0 && host_ids && cap > 0) {
            CKC(cudaMalloc((void **) &d_ids, sizeof(int) * (size_t) nsel));
            k_compact<<<(n_items + 255) / 256, 256, 0, st>>>(n_items, flag, pos, d_ids);
            CKC(cudaMemcpyAsync(host_ids, d_ids, sizeof(int) * (size_t) std::min(nsel, cap), cudaMemcpyDeviceToHost, st));
            CKC(cudaStreamSynchronize(st));
            g_mcu_launches += 1;
        }
    }
    *n_out = nsel;
done:
    for (int i = 0; i < 2; i++) { if (key[i]) cudaFree(key[i]); if (val[i]) cudaFree(val[i]); }
    if (key_c) cudaFree(key_c);
    if (flag) cudaFree(flag);
    if (pos) cudaFree(pos);
    if (tmp) cudaFree(tmp);
    if (d_ids) cudaFree(d_ids);
    return rc;
}

// ---------------------------------------------------------------------------------
// Grade-lowering connectivity conn(row, col), 1 <= row < col (mesh_addlowermatrix, mesh.c:621-660): for every element of
// grade col the elements of grade row all of whose vertices it contains.  The reference merges the incidence lists of
// the element's vertices, sorts them and counts repeats (mesh_matchelements, mesh.c:584-617), then inserts the
// matches into a dictionary-of-keys matrix; here the elements of grade row are sorted by their vertex tuple once and
// every (element, subset) looks its tuple up by binary search.  Columns come out with ascending ids, as the
// reference's conversion to compressed columns leaves them.
// ---------------------------------------------------------------------------------
__global__ void k_elem_keys(int nel, int n, const int *__restrict__ rix, unsigned long long *__restrict__ key_ab, unsigned *__restrict__ val) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nel) return;
    const int *ids = rix + (size_t) e * n;
    key_ab[e] = ((unsigned long long) (unsigned) ids[0] << 32) | (unsigned) ids[1];
    val[e] = (unsigned) e;
}

// two elements of grade row with the same vertices: the reference would list both; left to it
__global__ void k_dup_check(int nel, int n, const int *__restrict__ rix, const unsigned long long *__restrict__ key_sorted,
                            const unsigned *__restrict__ val_sorted, unsigned *__restrict__ dup) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 1 || i >= nel || key_sorted[i] != key_sorted[i - 1]) return;
    if (n == 2) { *dup = 1u; return; }
    // triangles: the same leading pair is common (faces around an edge); compare the third id within the run
    const int c = rix[(size_t) val_sorted[i] * 3 + 2];
    for (int j = i - 1; j >= 0 && key_sorted[j] == key_sorted[i]; j--)
        if (rix[(size_t) val_sorted[j] * 3 + 2] == c) { *dup = 1u; return; }
}

__global__ void k_lower(Combos cb, int ncol_el, const int *__restrict__ col_rix, int nrow_el, const int *__restrict__ row_rix,
                        const unsigned long long *__restrict__ key_sorted, const unsigned *__restrict__ val_sorted,
                        int *__restrict__ found, unsigned *__restrict__ count) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ncol_el) return;
    const int *ids = col_rix + (size_t) e * cb.nvper;
    int got[6], n = 0;
    for (int c = 0; c < cb.ncomb; c++) {
        const unsigned a = (unsigned) ids[cb.idx[c][0]], b = (unsigned) ids[cb.idx[c][1]];
        const unsigned long long key = ((unsigned long long) a << 32) | b;
        long lo = 0, hi = nrow_el;
        while (lo < hi) { long mid = (lo + hi) >> 1; if (key_sorted[mid] < key) lo = mid + 1; else hi = mid; }
        int id = -1;
        for (long i = lo; i < nrow_el && key_sorted[i] == key; i++) {
            const int cand = (int) val_sorted[i];
            if (cb.n == 2 || row_rix[(size_t) cand * 3 + 2] == ids[cb.idx[c][2]]) { id = cand; break; }
        }
        if (id < 0) continue;
        int k = n++;                                   // insertion into the ascending list
        while (k > 0 && got[k - 1] > id) { got[k] = got[k - 1]; k--; }
        got[k] = id;
    }
    for (int k = 0; k < n; k++) found[(size_t) e * cb.ncomb + k] = got[k];
    count[e] = (unsigned) n;
}

__global__ void k_lower_emit(int ncol_el, int ncomb, const int *__restrict__ found, const unsigned *__restrict__ count,
                             const unsigned *__restrict__ cptr, int *__restrict__ out) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ncol_el) return;
    for (unsigned k = 0; k < count[e]; k++) out[cptr[e] + k] = found[(size_t) e * ncomb + k];
}

extern "C" int mcu_mesh_lower(mcu_mesh *m, int row, int col, int *host_cptr, int *host_rix, int *nnz_out, int *nrows_out, int *ncols_out) {
    if (!m || row < 1 || row > 2 || col <= row || col > 3 || !host_cptr || !host_rix || !nnz_out) { mcu_set_error("lower: bad arguments"); return MCU_ERR_ARGS; }
    if (!m->conn[row].present || !m->conn[col].present) { mcu_set_error("lower: grades %d and %d must be present", row, col); return MCU_ERR_ELNOTFOUND; }
    const McuConn &R = m->conn[row], &Cc = m->conn[col];
    Combos cb;
    cb.n = row + 1; cb.nvper = col + 1; cb.ncomb = 0;
    if (cb.n == 2) { for (int a = 0; a <= col; a++) for (int b = a + 1; b <= col; b++) { cb.idx[cb.ncomb][0] = a; cb.idx[cb.ncomb][1] = b; cb.idx[cb.ncomb][2] = 0; cb.ncomb++; } }
    else { for (int a = 0; a <= col; a++) for (int b = a + 1; b <= col; b++) for (int c = b + 1; c <= col; c++) { cb.idx[cb.ncomb][0] = a; cb.idx[cb.ncomb][1] = b; cb.idx[cb.ncomb][2] = c; cb.ncomb++; } }
    cudaStream_t st = mcu_stream();
    int rc = MCU_OK;
    unsigned long long *key[2] = {nullptr, nullptr};
    unsigned *val[2] = {nullptr, nullptr}, *count = nullptr, *cptr = nullptr, *dup = nullptr;
    int *found = nullptr, *d_out = nullptr;
    void *tmp = nullptr;
    size_t t1 = 0, t2 = 0;
    unsigned h_dup = 0, last_c = 0, last_p = 0;
    const int nr = R.nel, nc = Cc.nel;
    host_cptr[0] = 0;
    *nnz_out = 0;
    if (nrows_out) *nrows_out = 0;
    if (ncols_out) *ncols_out = 0;
    if (nc == 0) return MCU_OK;
    CKC(cudaMalloc((void **) &key[0], 8 * (size_t) std::max(nr, 1))); CKC(cudaMalloc((void **) &key[1], 8 * (size_t) std::max(nr, 1)));
    CKC(cudaMalloc((void **) &val[0], 4 * (size_t) std::max(nr, 1))); CKC(cudaMalloc((void **) &val[1], 4 * (size_t) std::max(nr, 1)));
    CKC(cudaMalloc((void **) &count, 4 * (size_t) nc)); CKC(cudaMalloc((void **) &cptr, 4 * (size_t) nc));
    CKC(cudaMalloc((void **) &found, sizeof(int) * (size_t) nc * cb.ncomb));
    CKC(cudaMalloc((void **) &dup, 4));
    CKC(cudaMemsetAsync(dup, 0, 4, st));
    if (nr > 0) {
        CKC(cub::DeviceRadixSort::SortPairs(nullptr, t1, key[0], key[1], val[0], val[1], nr, 0, 64, st));
        CKC(cudaMalloc(&tmp, std::max<size_t>(t1, 16)));
        k_elem_keys<<<(nr + 255) / 256, 256, 0, st>>>(nr, cb.n, R.d_rix, key[0], val[0]);
        CKC(cub::DeviceRadixSort::SortPairs(tmp, t1, key[0], key[1], val[0], val[1], nr, 0, 64, st));
        cudaFree(tmp); tmp = nullptr;
        k_dup_check<<<(nr + 255) / 256, 256, 0, st>>>(nr, cb.n, R.d_rix, key[1], val[1], dup);
    }
    k_lower<<<(nc + 255) / 256, 256, 0, st>>>(cb, nc, Cc.d_rix, nr, R.d_rix, key[1], val[1], found, count);
    CKC(cub::DeviceScan::ExclusiveSum(nullptr, t2, count, cptr, nc, st));
    CKC(cudaMalloc(&tmp, std::max<size_t>(t2, 16)));
    CKC(cub::DeviceScan::ExclusiveSum(tmp, t2, count, cptr, nc, st));
    CKC(cudaMemcpyAsync(&h_dup, dup, 4, cudaMemcpyDeviceToHost, st));
    CKC(cudaMemcpyAsync(&last_c, count + (nc - 1), 4, cudaMemcpyDeviceToHost, st));
    CKC(cudaMemcpyAsync(&last_p, cptr + (nc - 1), 4, cudaMemcpyDeviceToHost, st));
    CKC(cudaStreamSynchronize(st));
    g_mcu_launches += 4;
    if (h_dup) { mcu_set_error("lower: two elements of grade %d have the same vertices", row); rc = MCU_ERR_UNSUPPORTED; goto done; }
    {
        const int nnz = (int) (last_c + last_p);
        *nnz_out = nnz;
        CKC(cudaMalloc((void **) &d_out, sizeof(int) * (size_t) std::max(nnz, 1)));
        k_lower_emit<<<(nc + 255) / 256, 256, 0, st>>>(nc, cb.ncomb, found, count, cptr, d_out);
        CKC(cudaMemcpyAsync(host_cptr, cptr, sizeof(int) * (size_t) nc, cudaMemcpyDeviceToHost, st));
        CKC(cudaMemcpyAsync(host_rix, d_out, sizeof(int) * (size_t) nnz, cudaMemcpyDeviceToHost, st));
        CKC(cudaStreamSynchronize(st));
        host_cptr[nc] = nnz;
        int mx = -1;
        for (int i = 0; i < nnz; i++) mx = std::max(mx, host_rix[i]);
        if (nrows_out) *nrows_out = mx + 1;            // the size a dictionary-of-keys matrix grows to (sparse.c:137-160):
        int lastcol = nc;                              // largest row / column index with an entry, plus one
        while (lastcol > 0 && host_cptr[lastcol] == host_cptr[lastcol - 1]) lastcol--;
        if (ncols_out) *ncols_out = lastcol;
        g_mcu_launches += 1;
    }
done:
    for (int i = 0; i < 2; i++) { if (key[i]) cudaFree(key[i]); if (val[i]) cudaFree(val[i]); }
    if (count) cudaFree(count);
    if (cptr) cudaFree(cptr);
    if (found) cudaFree(found);
    if (dup) cudaFree(dup);
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

// mcu_block.cu — owner-computes gradient kernel for ANY simplicial grade and for Selections.
//
// Replaces the serial scatter loop of functional_mapgradientX (functional.c:356-416) for the analytic gradients
// (Length 1960-1972, Area 2079-2103, VolumeEnclosed 2142-2164, Volume 2201-2227) where the persistent triangle
// patch kernel (mcu_patch.cu) does not apply: tetrahedra (BASELINE config 5: Volume.gradient on 50.9M tets), line
// elements, and any grade restricted to a Selection (functional.c:383-394).  The generic CSR-transpose gather
// (k_grad_gather) re-reads the incidence and gathers 24-byte rows from L2 once per (vertex, element, corner): on the
// tetrahedral cube that is ~18 GB of L2 traffic per gradient (2.3 ms, 0.08 of the HBM roofline).
//
// Here the vertices that have elements are cut into BLOCKS of up to 256 "owned" vertices that are compact in space
// (Morton order of the positions at planning time).  One CTA per block:
//   1. stages the coordinates of the block's local vertices (owned + every other vertex of their elements, at most
//      1023: slot indices are 10 bit) in shared memory, 24 KB, SoA;
//   2. one thread per owned vertex walks that vertex's elements in ASCENDING ELEMENT ID — the reference's summation
//      order (F8), the same as k_grad_gather's — reading one packed 32-bit entry per element (the shared-memory slots
//      of the element's other vertices, sliced ELL per warp: coalesced) and evaluates its own corner's gradient;
//   3. writes its gradient row once.  No atomics; the result does not depend on the block decomposition.
// Eight CTAs fit an SM, so the staging of one block overlaps the arithmetic of the others.
//
// HBM traffic per tetrahedron (V = T/6): 16 B entries (4 corners x 4 B) + ~10 B coordinates incl. halo + 4 B rows
// written: above the 24 algorithmic bytes of SURVEY 8(d) by the halo re-reads only.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "mcu_host.h"
#include "mcu_block.h"

extern long g_mcu_launches;

#define BK_THREADS 256
#define BK_MAX_LOCAL 1023
#define BK_NULL 0xffffffffu

// ---------------------------------------------------------------------------------
// Planner (host, multi-threaded over blocks)
// ---------------------------------------------------------------------------------

static inline uint64_t spread3(uint64_t x) {          // 21 bits -> every third bit
    x &= 0x1fffff;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}

namespace {
struct BlockOut {                 // what one candidate block turns into
    std::vector<int> loc;         // global ids of the local vertices, owned first
    std::vector<uint32_t> ent;    // sliced ELL entries, warp after warp (lines, triangles)
    std::vector<uint16_t> ent16;  // tetrahedra: link-strip entries, warp after warp
    McuBlockHdr h;
};

// open-addressing map vertex id -> slot for the at most 1023 local vertices of a block
struct SlotMap {
    int key[4096], val[4096];
    void clear() { memset(key, 0xff, sizeof(key)); }
    int *find(int k) {
        unsigned h = ((unsigned) k * 2654435761u) >> 20;
        while (key[h] != -1 && key[h] != k) h = (h + 1) & 4095;
        return key[h] == k ? &val[h] : (key[h] = k, val[h] = -1, &val[h]);
    }
};
}  // namespace

// Tetrahedra: the elements of an owned vertex u are the triangles (a, b, c) of its LINK (the other three vertices of
// every incident tetrahedron).  Neighbouring link triangles share an edge, so walking the link as triangle STRIPS
// fetches ONE new vertex from shared memory per tetrahedron instead of three (the shared-memory pipe, not HBM or the
// FP64 pipe, is what bounds the one-entry-per-element form: profiles/r2_block_v1_summary.txt).  The thread keeps the
// current triangle (a, b, c), c the newest vertex; a 16-bit entry = 10-bit slot of the next vertex x + 2-bit op:
//   BK_OP_STRIP  cross the edge (b, c): (a, b, c) <- (b, c, x), evaluate
//   BK_OP_FAN    cross the edge (a, c): (a, b, c) <- (a, c, x), evaluate
//   BK_OP_LOAD   (a, b, c) <- (b, c, x) without evaluating: the first two vertices of a new strip
//   (lanes with shorter walks are padded with LOAD entries of slot 0)
#define BK_OP_STRIP 0u
#define BK_OP_FAN 1u
#define BK_OP_LOAD 2u

static void link_strips(int d, const int (*tri)[3], std::vector<uint16_t> &out) {
    // adjacency across the three edges of every link triangle (edge k = the one opposite vertex k)
    std::vector<int> nb((size_t) d * 3, -1);
    for (int t = 0; t < d; t++)
        for (int k = 0; k < 3; k++) {
            if (nb[(size_t) t * 3 + k] >= 0) continue;
            const int e0 = tri[t][(k + 1) % 3], e1 = tri[t][(k + 2) % 3];
            for (int t2 = t + 1; t2 < d; t2++) {
                int k2 = -1, hit = 0;
                for (int j = 0; j < 3; j++) { if (tri[t2][j] == e0 || tri[t2][j] == e1) hit++; else k2 = j; }
                if (hit == 2 && k2 >= 0 && nb[(size_t) t2 * 3 + k2] < 0) { nb[(size_t) t * 3 + k] = t2; nb[(size_t) t2 * 3 + k2] = t; break; }
            }
        }
    std::vector<char> done((size_t) d, 0);
    auto free_nbrs = [&](int t) { int c = 0; for (int k = 0; k < 3; k++) { int o = nb[(size_t) t * 3 + k]; if (o >= 0 && !done[o]) c++; } return c; };
    int left = d;
    while (left > 0) {
        // start at the unvisited triangle with the fewest unvisited neighbours (ends of strips first), lowest index on ties
        int t0 = -1, best = 4;
        for (int t = 0; t < d; t++) if (!done[t]) { int f = free_nbrs(t); if (f < best) { best = f; t0 = t; } }
        done[t0] = 1; left--;
        // orient the start so that the edge towards the next triangle is (b, c)
        int knext = -1;
        for (int k = 0; k < 3 && knext < 0; k++) { int o = nb[(size_t) t0 * 3 + k]; if (o >= 0 && !done[o]) knext = k; }
        int a = tri[t0][knext < 0 ? 0 : knext], b = tri[t0][((knext < 0 ? 0 : knext) + 1) % 3], c = tri[t0][((knext < 0 ? 0 : knext) + 2) % 3];
        out.push_back((uint16_t) (a | BK_OP_LOAD << 10)); out.push_back((uint16_t) (b | BK_OP_LOAD << 10)); out.push_back((uint16_t) (c | BK_OP_STRIP << 10));
        int cur = t0;
        for (;;) {
            // unvisited neighbours across (b, c) [strip] and (a, c) [fan]
            int cand[2] = {-1, -1};
            for (int k = 0; k < 3; k++) {
                int o = nb[(size_t) cur * 3 + k];
                if (o < 0 || done[o]) continue;
                const int opp = tri[cur][k];
                if (opp == a) cand[0] = o; else if (opp == b) cand[1] = o;
            }
            int pick = -1;
            if (cand[0] >= 0 && cand[1] >= 0) pick = free_nbrs(cand[0]) <= free_nbrs(cand[1]) ? 0 : 1;
            else if (cand[0] >= 0) pick = 0;
            else if (cand[1] >= 0) pick = 1;
            if (pick < 0) break;
            const int nxt = cand[pick];
            int x = -1;
            for (int j = 0; j < 3; j++) { const int w = tri[nxt][j]; if (w != c && w != (pick == 0 ? b : a)) x = w; }
            out.push_back((uint16_t) (x | (pick == 0 ? BK_OP_STRIP : BK_OP_FAN) << 10));
            if (pick == 0) { a = b; b = c; c = x; } else { b = c; c = x; }
            done[nxt] = 1; left--;
            cur = nxt;
        }
    }
}

// Build the block owning `owned[0..n)` (ascending ids).  false: more than BK_MAX_LOCAL local vertices.
static bool build_block(const int *owned, int n, int nvper, const int *rix, const int *eids, const std::vector<int> &tptr,
                        const std::vector<int> &tidx, SlotMap &map, std::vector<int> &halo, BlockOut &out) {
    map.clear();
    halo.clear();
    for (int i = 0; i < n; i++) *map.find(owned[i]) = i;
    for (int i = 0; i < n; i++) {
        const int u = owned[i];
        for (int q = tptr[u]; q < tptr[u + 1]; q++) {
            const int *ids = rix + (size_t) (eids ? eids[tidx[q]] : tidx[q]) * nvper;
            for (int j = 0; j < nvper; j++) {
                int *s = map.find(ids[j]);
                if (*s == -1) { *s = -2; halo.push_back(ids[j]); if (n + (int) halo.size() > BK_MAX_LOCAL) return false; }
            }
        }
    }
    std::sort(halo.begin(), halo.end());
    for (size_t i = 0; i < halo.size(); i++) *map.find(halo[i]) = n + (int) i;
    out.loc.assign(owned, owned + n);
    out.loc.insert(out.loc.end(), halo.begin(), halo.end());
    memset(&out.h, 0, sizeof(out.h));
    out.h.n_owned = n; out.h.n_local = (int) out.loc.size();
    out.ent.clear();
    out.ent16.clear();
    if (nvper == 4) {
        std::vector<uint16_t> walk[32];
        std::vector<int> tri;
        for (int w = 0; w * 32 < n; w++) {
            size_t deg = 0;
            for (int l = 0; l < 32; l++) {
                walk[l].clear();
                if (w * 32 + l >= n) continue;
                const int u = owned[w * 32 + l];
                tri.clear();
                for (int q = tptr[u]; q < tptr[u + 1]; q++) {
                    const int *ids = rix + (size_t) (eids ? eids[tidx[q]] : tidx[q]) * 4;
                    bool seen = false;
                    for (int j = 0; j < 4; j++) {
                        if (ids[j] == u && !seen) { seen = true; continue; }
                        tri.push_back(*map.find(ids[j]));
                    }
                }
                link_strips((int) (tri.size() / 3), (const int (*)[3]) tri.data(), walk[l]);
                deg = std::max(deg, walk[l].size());
            }
            // groups of four entries per lane (one 8-byte load, issued a group ahead of its use); lanes with shorter
            // walks are padded with LOAD entries of slot 0, which evaluate nothing
            const size_t groups = (deg + 3) / 4;
            if (groups > 255) return false;
            out.h.ent_off[w] = (unsigned) out.ent16.size();
            out.h.deg[w] = (unsigned char) groups;
            const size_t base = out.ent16.size();
            out.ent16.resize(base + groups * 128, (uint16_t) (BK_OP_LOAD << 10));
            for (int l = 0; l < 32; l++) for (size_t k = 0; k < walk[l].size(); k++) out.ent16[base + ((k / 4) * 32 + l) * 4 + (k & 3)] = walk[l][k];
        }
        return true;
    }
    for (int w = 0; w * 32 < n; w++) {
        int deg = 0;
        for (int l = 0; l < 32 && w * 32 + l < n; l++) { const int u = owned[w * 32 + l]; deg = std::max(deg, tptr[u + 1] - tptr[u]); }
        if (deg > 255) return false;                 // a vertex of extreme valence: left to the gather kernel
        out.h.ent_off[w] = (unsigned) out.ent.size();
        out.h.deg[w] = (unsigned char) deg;
        const size_t base = out.ent.size();
        out.ent.resize(base + (size_t) deg * 32, BK_NULL);
        for (int l = 0; l < 32 && w * 32 + l < n; l++) {
            const int u = owned[w * 32 + l];
            for (int q = tptr[u], k = 0; q < tptr[u + 1]; q++, k++) {
                const int *ids = rix + (size_t) (eids ? eids[tidx[q]] : tidx[q]) * nvper;
                uint32_t e = 0;
                int sh = 0;
                bool seen = false;
                for (int j = 0; j < nvper; j++) {
                    if (ids[j] == u && !seen) { seen = true; continue; }      // the owner's own corner (once)
                    e |= (uint32_t) *map.find(ids[j]) << sh;
                    sh += 10;
                }
                out.ent[base + (size_t) k * 32 + l] = e;
            }
        }
    }
    return true;
}

bool mcu_block_plan(int nv, int dim, int n, int nvper, const int *rix, const int *eids, const double *vert, McuBlockPlan &plan) {
    const bool timing = getenv("MCU_PLAN_TIMING") != nullptr;
    auto now_s = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now_s();
    std::vector<int> tptr, tidx;
    mcu_host_transpose(nv, n, nvper, rix, eids, tptr, tidx);
    const double t1 = now_s();

    // active vertices in Morton order of their positions
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    std::vector<std::pair<uint64_t, int>> keyed;
    for (int v = 0; v < nv; v++) {
        if (tptr[v + 1] == tptr[v]) continue;
        for (int k = 0; k < dim; k++) { double x = vert[(size_t) v * dim + k]; lo[k] = std::min(lo[k], x); hi[k] = std::max(hi[k], x); }
        keyed.push_back({0, v});
    }
    const size_t na = keyed.size();
    double scale[3] = {0, 0, 0};
    for (int k = 0; k < dim; k++) scale[k] = hi[k] > lo[k] ? 2097151.0 / (hi[k] - lo[k]) : 0.0;
    int nthreads = (int) std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    if (na < 100000) nthreads = 1;
    auto parallel = [&](size_t count, auto &&fn) {
        if (nthreads == 1) { fn(0, count, 0); return; }
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) {
            size_t a = count * t / nthreads, b = count * (t + 1) / nthreads;
            th.emplace_back([&fn, a, b, t] { fn(a, b, t); });
        }
        for (auto &x : th) x.join();
    };
    parallel(na, [&](size_t a, size_t b, int) {
        for (size_t i = a; i < b; i++) {
            const int v = keyed[i].second;
            uint64_t key = 0;
            for (int k = 0; k < dim; k++) key |= spread3((uint64_t) ((vert[(size_t) v * dim + k] - lo[k]) * scale[k])) << k;
            keyed[i].first = key;
        }
    });
    // sort: per-thread chunks, then pairwise merges
    {
        std::vector<size_t> cut(nthreads + 1);
        for (int t = 0; t <= nthreads; t++) cut[t] = na * t / nthreads;
        parallel((size_t) nthreads, [&](size_t a, size_t b, int) { for (size_t t = a; t < b; t++) std::sort(keyed.begin() + cut[t], keyed.begin() + cut[t + 1]); });
        for (int width = 1; width < nthreads; width *= 2) {
            std::vector<std::thread> th;
            for (int t = 0; t + width < nthreads; t += 2 * width) {
                size_t a = cut[t], m = cut[t + width], b = cut[std::min(t + 2 * width, nthreads)];
                th.emplace_back([&keyed, a, m, b] { std::inplace_merge(keyed.begin() + a, keyed.begin() + m, keyed.begin() + b); });
            }
            for (auto &x : th) x.join();
        }
    }
    const double t2 = now_s();

    // candidate blocks: chunks of BK_THREADS consecutive vertices of the order, halved while they do not fit
    const size_t ncand = (na + BK_THREADS - 1) / BK_THREADS;
    std::vector<std::vector<BlockOut>> per_cand(ncand);
    std::atomic<bool> failed(false);
    parallel(ncand, [&](size_t a, size_t b, int) {
        SlotMap *map = new SlotMap();
        std::vector<int> halo, owned;
        std::vector<std::pair<size_t, size_t>> todo;
        for (size_t c = a; c < b && !failed; c++) {
            todo.clear();
            todo.push_back({c * BK_THREADS, std::min(na, (c + 1) * BK_THREADS)});
            while (!todo.empty()) {
                auto rg = todo.back();
                todo.pop_back();
                owned.clear();
                for (size_t i = rg.first; i < rg.second; i++) owned.push_back(keyed[i].second);
                std::sort(owned.begin(), owned.end());
                BlockOut bo;
                if (build_block(owned.data(), (int) owned.size(), nvper, rix, eids, tptr, tidx, *map, halo, bo)) { per_cand[c].push_back(std::move(bo)); continue; }
                if (rg.second - rg.first <= 1) { failed = true; break; }
                size_t mid = rg.first + ((rg.second - rg.first) / 2 + 31) / 32 * 32;
                if (mid >= rg.second) mid = rg.first + (rg.second - rg.first) / 2;
                todo.push_back({mid, rg.second});
                todo.push_back({rg.first, mid});
            }
        }
        delete map;
    });
    if (failed) return false;
    const double t3 = now_s();

    std::vector<McuBlockHdr> &hdr = plan.hdr;
    hdr.clear();
    plan.n_halo = 0;
    size_t nloc = 0, nent = 0;
    const bool strips = (nvper == 4);
    plan.strips = strips;
    for (auto &pc : per_cand) for (auto &bo : pc) { nloc += bo.loc.size(); nent += strips ? bo.ent16.size() : bo.ent.size(); }
    if (nent >= 0xfffffff0u) return false;
    std::vector<int> &loc = plan.loc;
    std::vector<uint32_t> &ent = plan.ent;
    loc.assign(nloc, 0);
    ent.assign(strips ? (nent + 1) / 2 : nent, 0u);
    nloc = nent = 0;
    for (auto &pc : per_cand) for (auto &bo : pc) {
        McuBlockHdr h = bo.h;
        h.loc_off = (int) nloc;
        for (int w = 0; w < 8; w++) h.ent_off[w] += (unsigned) nent;
        memcpy(loc.data() + nloc, bo.loc.data(), bo.loc.size() * sizeof(int));
        if (strips) memcpy((uint16_t *) ent.data() + nent, bo.ent16.data(), bo.ent16.size() * sizeof(uint16_t));
        else memcpy(ent.data() + nent, bo.ent.data(), bo.ent.size() * sizeof(uint32_t));
        nloc += bo.loc.size(); nent += strips ? bo.ent16.size() : bo.ent.size();
        plan.n_halo += h.n_local - h.n_owned;
        hdr.push_back(h);
    }
    plan.n_active = (long) na; plan.all_active = (na == (size_t) nv);
    if (timing)
        fprintf(stderr, "[mcu_block_plan] nv=%d n=%d nvper=%d: %zu blocks, %.2f local vertices per owned, %zu entries; transpose %.2fs, order %.2fs, "
                        "blocks %.2fs (%d threads)\n", nv, n, nvper, hdr.size(), na ? (double) nloc / na : 0.0, nent, t1 - t0, t2 - t1, t3 - t2, nthreads);
    return true;
}

McuBlockSet *mcu_blockset_build(int nv, int dim, int n, int nvper, const int *rix, const int *eids, const double *vert, cudaStream_t st) {
    McuBlockPlan plan;
    if (!mcu_block_plan(nv, dim, n, nvper, rix, eids, vert, plan)) return nullptr;
    const std::vector<McuBlockHdr> &hdr = plan.hdr;
    const std::vector<int> &loc = plan.loc;
    const std::vector<uint32_t> &ent = plan.ent;
    const size_t nloc = loc.size(), nent = ent.size();        // 32-bit words (two strip entries per word)
    McuBlockSet *bs = new McuBlockSet();
    bs->n_halo = plan.n_halo;
    bs->strips = plan.strips;
    bs->n_blocks = (int) hdr.size(); bs->n_active = plan.n_active; bs->n_entries = (long) nent; bs->all_active = plan.all_active;
    bool ok = cudaMalloc((void **) &bs->d_hdr, sizeof(McuBlockHdr) * std::max<size_t>(hdr.size(), 1)) == cudaSuccess &&
              cudaMalloc((void **) &bs->d_loc, sizeof(int) * std::max<size_t>(nloc, 1)) == cudaSuccess &&
              cudaMalloc((void **) &bs->d_ent, sizeof(uint32_t) * std::max<size_t>(nent, 1)) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(bs->d_hdr, hdr.data(), sizeof(McuBlockHdr) * hdr.size(), cudaMemcpyHostToDevice, st) == cudaSuccess &&
         cudaMemcpyAsync(bs->d_loc, loc.data(), sizeof(int) * nloc, cudaMemcpyHostToDevice, st) == cudaSuccess &&
         cudaMemcpyAsync(bs->d_ent, ent.data(), sizeof(uint32_t) * nent, cudaMemcpyHostToDevice, st) == cudaSuccess &&
         cudaStreamSynchronize(st) == cudaSuccess;
    if (!ok) { mcu_blockset_free(bs); return nullptr; }
    return bs;
}

// introspection for the CPU tests of the host logic (no CUDA call): counts = {blocks, local vertices, entries, active
// vertices, all_active}; arrays may be NULL (sizes from a first call)
extern "C" int mcu_debug_block_plan(int nv, int dim, int n, int nvper, const int *rix, const int *eids, const double *vert, long counts[5],
                                    int *hdr_out, int *loc_out, unsigned *ent_out) {
    McuBlockPlan plan;
    if (!mcu_block_plan(nv, dim, n, nvper, rix, eids, vert, plan)) return MCU_ERR_UNSUPPORTED;
    counts[0] = (long) plan.hdr.size(); counts[1] = (long) plan.loc.size(); counts[2] = (long) plan.ent.size();
    counts[3] = plan.n_active; counts[4] = plan.all_active ? 1 : 0;
    if (hdr_out) memcpy(hdr_out, plan.hdr.data(), plan.hdr.size() * sizeof(McuBlockHdr));
    if (loc_out) memcpy(loc_out, plan.loc.data(), plan.loc.size() * sizeof(int));
    if (ent_out) memcpy(ent_out, plan.ent.data(), plan.ent.size() * sizeof(uint32_t));
    return MCU_OK;
}

void mcu_blockset_free(McuBlockSet *bs) {
    if (!bs) return;
    if (bs->d_hdr) cudaFree(bs->d_hdr);
    if (bs->d_loc) cudaFree(bs->d_loc);
    if (bs->d_ent) cudaFree(bs->d_ent);
    delete bs;
}

// ---------------------------------------------------------------------------------
// Kernel
// ---------------------------------------------------------------------------------

__device__ __forceinline__ double bk_rsqrt(double x) {      // as rsqrt_normal of mcu_patch.cu
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x * y, y, 1.0);
    double t = fma(e, 0.375, 0.5);
    return fma(y * e, t, y);
}

// Gradient of the element's size with respect to the owner p; the other vertices come in the element's own order, but
// every expression below is invariant under their permutations:
//   Length          (p - a) / |p - a|                                                           functional.c:1960-1972
//   Area            e1 = a-p, e2 = b-p, N = e1 x e2 ;  N x (e2-e1) / (2|N|)                        functional.c:2079-2103
//   VolumeEnclosed  sign(p.(a x b)) (a x b) / 6                                                    functional.c:2142-2164
//   Volume          N = (b-a) x (c-a) ;  sign((p-a).N) N / 6   ( = the reference's four cross products, each
//                   written from its own vertex: d/dp |det[a-p, b-p, c-p]| / 6 )                   functional.c:2201-2227
template <int F>
__device__ __forceinline__ void bk_corner(const double *sx, const double *sy, const double *sz, double px, double py, double pz,
                                          unsigned e, double *acc, unsigned &fl) {
    const int ia = e & 1023, ib = (e >> 10) & 1023, ic = (e >> 20) & 1023;
    if (F == MCU_F_LENGTH) {
        const double dx = px - sx[ia], dy = py - sy[ia], dz = pz - sz[ia];
        const double n2 = dx * dx + dy * dy + dz * dz;
        if (n2 < MCU_EPS * MCU_EPS) { fl |= MCU_FLAG_DEGENERATE; return; }
        const double r = bk_rsqrt(n2);
        acc[0] += dx * r; acc[1] += dy * r; acc[2] += dz * r;
    } else if (F == MCU_F_AREA) {
        const double e1x = sx[ia] - px, e1y = sy[ia] - py, e1z = sz[ia] - pz;
        const double e2x = sx[ib] - px, e2y = sy[ib] - py, e2z = sz[ib] - pz;
        const double nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
        const double n2 = nx * nx + ny * ny + nz * nz;
        if (n2 < MCU_EPS * MCU_EPS) { fl |= MCU_FLAG_DEGENERATE; return; }
        const double r = 0.5 * bk_rsqrt(n2);
        const double dx = e2x - e1x, dy = e2y - e1y, dz = e2z - e1z;
        acc[0] += (ny * dz - nz * dy) * r; acc[1] += (nz * dx - nx * dz) * r; acc[2] += (nx * dy - ny * dx) * r;
    } else if (F == MCU_F_VOLUMEENCLOSED) {
        // un-fused, as the reference computes it: on a fine closed surface the vertex sums cancel to ~1e-3 of their terms
        const double ax = sx[ia], ay = sy[ia], az = sz[ia], bx = sx[ib], by = sy[ib], bz = sz[ib];
        const double cx = __dsub_rn(__dmul_rn(ay, bz), __dmul_rn(az, by)), cy = __dsub_rn(__dmul_rn(az, bx), __dmul_rn(ax, bz)),
                     cz = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
        const double d = __dadd_rn(__dadd_rn(__dmul_rn(px, cx), __dmul_rn(py, cy)), __dmul_rn(pz, cz));
        if (fabs(d) <= DBL_MIN) { fl |= MCU_FLAG_VOLENCLZERO; return; }
        const double s = d > 0 ? 1.0 / 6.0 : -1.0 / 6.0;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(s, cx)); acc[1] = __dadd_rn(acc[1], __dmul_rn(s, cy)); acc[2] = __dadd_rn(acc[2], __dmul_rn(s, cz));
    } else {
        const double ax = sx[ia], ay = sy[ia], az = sz[ia];
        const double ux = sx[ib] - ax, uy = sy[ib] - ay, uz = sz[ib] - az;
        const double vx = sx[ic] - ax, vy = sy[ic] - ay, vz = sz[ic] - az;
        const double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
        const double d = (px - ax) * nx + (py - ay) * ny + (pz - az) * nz;
        const double s = d > 0 ? 1.0 / 6.0 : -1.0 / 6.0;
        acc[0] += s * nx; acc[1] += s * ny; acc[2] += s * nz;
    }
}

template <int F>
__global__ void __launch_bounds__(BK_THREADS) k_block_grad(const McuBlockHdr *__restrict__ hdr, const int *__restrict__ loc,
                                                           const unsigned *__restrict__ ent, const double *__restrict__ vert, int dim,
                                                           double *__restrict__ out, unsigned *flags) {
    __shared__ double sx[BK_MAX_LOCAL + 1], sy[BK_MAX_LOCAL + 1], sz[BK_MAX_LOCAL + 1];
    __shared__ McuBlockHdr h;
    if (threadIdx.x < sizeof(McuBlockHdr) / 4) ((int *) &h)[threadIdx.x] = ((const int *) (hdr + blockIdx.x))[threadIdx.x];
    __syncthreads();
    const int *myloc = loc + h.loc_off;
    int gid = -1;
    for (int i = threadIdx.x; i < h.n_local; i += BK_THREADS) {
        const int id = myloc[i];
        if (i == (int) threadIdx.x) gid = id;
        const double *x = vert + (size_t) id * dim;
        sx[i] = x[0]; sy[i] = x[1]; sz[i] = (dim == 3) ? x[2] : 0.0;
    }
    __syncthreads();
    if ((int) threadIdx.x >= h.n_owned) return;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double px = sx[threadIdx.x], py = sy[threadIdx.x], pz = sz[threadIdx.x];
    double acc[3] = {0.0, 0.0, 0.0};
    unsigned fl = 0;
    const unsigned *e = ent + h.ent_off[w] + lane;
    const int deg = h.deg[w];
#pragma unroll 4
    for (int j = 0; j < deg; j++) {
        const unsigned u = e[(size_t) j * 32];
        if (u != BK_NULL) bk_corner<F>(sx, sy, sz, px, py, pz, u, acc, fl);
    }
    double *row = out + (size_t) gid * dim;
    row[0] = acc[0]; row[1] = acc[1];
    if (dim == 3) row[2] = acc[2];
    if (fl) atomicOr(flags, fl);            // status word only; never data
}

// Volume on tetrahedra: the link of the owned vertex walked as triangle strips (see link_strips).  Every entry fetches
// one vertex; (a, b, c) is the current link triangle.  Branch free: LOAD and PAD entries evaluate with weight zero.
__global__ void __launch_bounds__(BK_THREADS) k_block_grad_tet(const McuBlockHdr *__restrict__ hdr, const int *__restrict__ loc,
                                                               const unsigned short *__restrict__ ent, const double *__restrict__ vert,
                                                               double *__restrict__ out) {
    __shared__ double sx[BK_MAX_LOCAL + 1], sy[BK_MAX_LOCAL + 1], sz[BK_MAX_LOCAL + 1];
    __shared__ McuBlockHdr h;
    if (threadIdx.x < sizeof(McuBlockHdr) / 4) ((int *) &h)[threadIdx.x] = ((const int *) (hdr + blockIdx.x))[threadIdx.x];
    __syncthreads();
    const int *myloc = loc + h.loc_off;
    int gid = -1;
    for (int i = threadIdx.x; i < h.n_local; i += BK_THREADS) {
        const int id = myloc[i];
        if (i == (int) threadIdx.x) gid = id;
        const double *x = vert + (size_t) id * 3;
        sx[i] = x[0]; sy[i] = x[1]; sz[i] = x[2];
    }
    __syncthreads();
    if ((int) threadIdx.x >= h.n_owned) return;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double px = sx[threadIdx.x], py = sy[threadIdx.x], pz = sz[threadIdx.x];
    double ax = 0, ay = 0, az = 0, bx = 0, by = 0, bz = 0, cx = 0, cy = 0, cz = 0;
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
    // four entries per lane and group in one 8-byte load, fetched one group ahead (ent_off is a multiple of 128 entries)
    const uint2 *e = (const uint2 *) (ent + h.ent_off[w]) + lane;
    const int groups = h.deg[w];
    // (volatile asm keeps the loads at the top of the iteration, two groups ahead: the compiler otherwise sinks them
    //  to the bottom, right in front of their first use)
    auto fetch = [&](int j) -> uint2 {
        uint2 r = make_uint2(0, 0);
        if (j < groups) asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(e + (size_t) j * 32));
        return r;
    };
    uint2 cur = fetch(0), nxt = fetch(1);
    for (int j = 0; j < groups; j++) {
        const uint2 nxt2 = fetch(j + 2);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned u = ((k < 2 ? cur.x : cur.y) >> (16 * (k & 1))) & 0xffffu;
            const unsigned op = u >> 10;
            const int s = u & 1023;
            const bool fan = (op == BK_OP_FAN);
            ax = fan ? ax : bx; ay = fan ? ay : by; az = fan ? az : bz;
            bx = cx; by = cy; bz = cz;
            cx = sx[s]; cy = sy[s]; cz = sz[s];
            const double ux = bx - ax, uy = by - ay, uz = bz - az;
            const double vx = cx - ax, vy = cy - ay, vz = cz - az;
            const double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
            const double d = (px - ax) * nx + (py - ay) * ny + (pz - az) * nz;
            double wgt = d > 0 ? 1.0 / 6.0 : -1.0 / 6.0;
            wgt = (op >= BK_OP_LOAD) ? 0.0 : wgt;
            g0 += wgt * nx; g1 += wgt * ny; g2 += wgt * nz;
        }
        cur = nxt; nxt = nxt2;
    }
    double *row = out + (size_t) gid * 3;
    row[0] = g0; row[1] = g1; row[2] = g2;
}

cudaError_t mcu_block_gradient(int functional, const McuBlockSet &bs, const double *d_vert, int nv, int dim, double *d_out, unsigned *flags, cudaStream_t st) {
    if (!bs.all_active) {                   // vertices without (selected) elements keep zero rows
        cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t) nv * dim, st);
        if (e != cudaSuccess) return e;
    }
    if (bs.n_blocks == 0) return cudaSuccess;
    static bool carveout_set = false;
    if (!carveout_set) {                    // 25 KB of shared memory per CTA: let eight of them share an SM
        cudaFuncSetAttribute(k_block_grad_tet, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_block_grad<MCU_F_LENGTH>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_block_grad<MCU_F_AREA>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_block_grad<MCU_F_VOLUMEENCLOSED>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carveout_set = true;
    }
    if (bs.strips) {
        if (functional != MCU_F_VOLUME || dim != 3) return cudaErrorInvalidValue;
        k_block_grad_tet<<<bs.n_blocks, BK_THREADS, 0, st>>>(bs.d_hdr, bs.d_loc, (const unsigned short *) bs.d_ent, d_vert, d_out);
        g_mcu_launches += 1;
        return cudaGetLastError();
    }
    switch (functional) {
        case MCU_F_LENGTH: k_block_grad<MCU_F_LENGTH><<<bs.n_blocks, BK_THREADS, 0, st>>>(bs.d_hdr, bs.d_loc, bs.d_ent, d_vert, dim, d_out, flags); break;
        case MCU_F_AREA: k_block_grad<MCU_F_AREA><<<bs.n_blocks, BK_THREADS, 0, st>>>(bs.d_hdr, bs.d_loc, bs.d_ent, d_vert, dim, d_out, flags); break;
        case MCU_F_VOLUMEENCLOSED: k_block_grad<MCU_F_VOLUMEENCLOSED><<<bs.n_blocks, BK_THREADS, 0, st>>>(bs.d_hdr, bs.d_loc, bs.d_ent, d_vert, dim, d_out, flags); break;
        default: return cudaErrorInvalidValue;
    }
    g_mcu_launches += 1;
    return cudaGetLastError();
}

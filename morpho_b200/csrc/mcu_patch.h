// mcu_patch.h — data structures of the patch decomposition behind the triangle-mesh gradient kernel
// (mcu_patch.cu).  See DESIGN.md §3 for the layout and §4 for the kernel.
#pragma once
#include <cstdint>
#include <vector>

#define MCU_PATCH_MAX_SLOTS 1024      // slot indices are 10 bit
#ifndef PK_CONSUMER_WARPS
#define PK_CONSUMER_WARPS 16
#endif
#define MCU_PATCH_MAX_OWNED (32 * PK_CONSUMER_WARPS)   // one consumer thread per owned vertex
#define MCU_PATCH_DEFAULT_OWNED MCU_PATCH_MAX_OWNED
#define MCU_PATCH_MAX_TAB_BYTES 24576 // staged table of one patch (ring entries + owned ids)

// One patch = up to 512 "owned" vertices (their gradient rows are produced by this patch's pass) plus the
// halo: every other vertex that shares a triangle with an owned one.  The local vertices occupy SLOTS of a
// shared-memory coordinate array.  Slots are laid out as RUNS of consecutive global vertex ids (each run starts
// at an even id and has even length, so that its 24-byte rows form a 16-byte aligned, 16-byte granular block of
// the vertex matrix): one bulk-async copy per run moves the coordinates HBM → shared memory.
struct McuPatchHeader {   // 32 bytes
    int run_off;    // first run descriptor of this patch in McuPatchPlan::runs
    int n_runs;
    int tab_off;    // offset of the staged table in McuPatchPlan::tab, in 16-byte units
    int tab_bytes;  // multiple of 16
    int n_owned;
    int n_slots;    // even, <= MCU_PATCH_MAX_SLOTS
    int n_slices;   // ceil(n_owned / 32); slice = the 32 owned vertices handled by one warp
    int tx_bytes;   // bytes the copies of this patch deliver: 24 * n_slots + tab_bytes
};

struct McuPatchRun {      // 8 bytes
    int gid;              // first global vertex id of the run (even)
    unsigned short slot;  // first slot (even)
    unsigned short n;     // vertices in the run (even)
};

// What the producer warps read per patch, addressed by the patch index alone (one latency, prefetchable):
#ifndef MCU_PATCH_DESC_RUNS
#define MCU_PATCH_DESC_RUNS 64        // = 32 lanes x producer warps
#endif
struct McuPatchTri {      // 16 bytes: the triangle list of a patch (totals), see McuPatchPlan::tri
    int off;              // offset of the staged list in McuPatchPlan::tri, in 16-byte units
    int bytes;            // multiple of 16
    int n_tri;
    int pad;
};
struct McuPatchDesc {     // 560 bytes
    McuPatchHeader h;
    McuPatchRun runs[MCU_PATCH_DESC_RUNS];   // the first runs of the patch; the rest (rare) at runs[h.run_off + 64 ..]
    McuPatchTri t;
};

// Staged table of a patch (32-bit words):
//   [0] index of the patch [1] n_slices [2] index of the patch's 17 per-warp push offsets [3] number of push entries (both 0 unless a
//       halo exchange is attached: rows of vertices shared with other GPUs are also stored into the peers' receive
//       buffers, see mcu_halo.cu)
//   [4 .. 4+32*n_slices)            global id of the owned vertex of every (slice, lane); -1 = idle lane
//   [.. + round4(n_slices+1))       entry offset of every slice (relative to the entry array), last = n_ent
//   n_ent 16-bit entries, sliced ELL: entry k of lane l of slice s at ents[off[s] + 32*k + l]
// Entries of one owned vertex: [0] its own slot; [1..] the ring of neighbouring vertices as STRIPS: consecutive
// entries (r_k, r_k+1) form one triangle with the owner unless r_k+1 carries the START bit (it then opens a new
// strip).  Columns are padded to the slice's length by repeating the last entry with the START bit.
#define MCU_PATCH_ENT_START 0x8000u
#define MCU_PATCH_ENT_SLOT 0x03ffu

struct McuPatchPlan {
    int nv = 0, nel = 0;
    std::vector<McuPatchHeader> hdr;
    std::vector<McuPatchRun> runs;
    std::vector<uint32_t> tab;       // all staged tables, each 16-byte aligned
    // Triangle lists for the TOTALS: every triangle belongs to the patch that owns its smallest vertex id (all three
    // vertices are local to that patch), so a total visits each triangle exactly once, one thread per triangle.
    // Staged list (32-bit words): [0] index of the patch [1] n_tri [2],[3] 0, then n_tri entries of 4 x 16 bit
    // (slot of vertex 0, 1, 2 in ascending global id, 0), padded to 16 bytes.
    std::vector<McuPatchTri> tri_hdr;
    std::vector<uint32_t> tri;
    int max_slots = 0, max_tab_bytes = 0, max_owned = 0, max_runs = 0;
    long n_halo = 0;                 // halo vertices over all patches (excluding alignment padding)
    long n_pad = 0;                  // slots that only exist for alignment
    // shared-memory model of the ring reads (64-bit loads of one half warp hit 16 distinct bank pairs or conflict)
    long lds_wavefronts = 0, lds_wavefronts_ideal = 0;
};

// Push of shared rows from inside the gradient kernel (multi-GPU): where thread `thread` of a patch stores a copy
// of its row — receive buffer of rank `peer`, row `k` of the (this rank, peer) pair list.
struct McuPushEntry { int thread, peer, k, pad; };
// Layout of the control words (ints) at the start of every rank's receive block (mcu_halo.cu):
#define MCU_HALO_MAXSLOT 4
#define MCU_HALO_MAXWORLD 16
#define MCU_HALO_FLAG(slot, q) ((slot) * MCU_HALO_MAXWORLD + (q))   // "rank q's rows of column block `slot` have landed": epoch
#define MCU_HALO_OWNDONE(slot) (64 + (slot))                        // this rank's own partial rows of `slot` are written: epoch
#define MCU_HALO_EPOCH(slot) (68 + (slot))                          // completed exchanges of `slot` (the device-side epoch)
#define MCU_HALO_CTRL_BYTES 1024
struct McuHaloPush {      // kernel parameter; entries == nullptr: no exchange attached
    const McuPushEntry *entries;
    const int *wptr;                 // per boundary patch 17 offsets into entries: entries of warp w = [wptr[w], wptr[w+1])
    const int *order;                // hand-out order of the patches: the ones that push first
    int *done;                       // counter of finished pushing patches (self re-arming)
    void *const *peer_block;         // device array: base of every rank's receive block (own block at [rank])
    unsigned long long recv_off;     // byte offset of recv[][][][] in a block
    int rank, world, max_m, ncomp, slot, col0;   // col0 = first column of this gradient in a row (slot * dim)
    int n_push_patches;              // patches with push entries: the last of them to finish raises the flags
};

struct McuPatchSet {
    McuPatchPlan plan;               // headers + statistics (bulk arrays are dropped after upload)
    std::vector<int> owned_gid;      // owner thread order: vertex of every (patch, slice, lane), -1 = idle lane
    std::vector<long> own_ptr;       // per patch offset into owned_gid (n_patches + 1)
    std::vector<int> tab_word_off;   // per patch: word offset of its table header in d_tab
    McuPushEntry *d_push = nullptr;
    int *d_push_wptr = nullptr;
    int *d_order = nullptr;          // hand-out order with the pushing patches first (attached exchange), + the done counter behind it
    int n_push_patches = 0;
    struct CompSum *d_partials = nullptr;   // totals: one compensated partial sum per (patch, consumer warp)
    int n_patches = 0;
    McuPatchDesc *d_desc = nullptr;
    McuPatchRun *d_runs = nullptr;
    uint32_t *d_tab = nullptr;
    uint32_t *d_tri = nullptr;
    int *d_sched = nullptr;          // [0] next patch to hand out, [1] CTAs that ran out of work (self re-arming)
    int coords_bytes = 0, stage_bytes = 0, stages = 0, ctas_per_sm = 0;
};

bool mcu_patch_plan(int nv, int nel, const int *rix, const double *vert, int max_owned, McuPatchPlan &plan);

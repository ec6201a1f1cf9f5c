// mcu_block.h — block decomposition behind the generic owner-computes gradient kernel (mcu_block.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>

// One block = up to 256 owned vertices (8 warps, one thread per vertex) + the other vertices of their elements.
struct McuBlockHdr {          // 56 bytes
    int loc_off;              // first entry of the block's local vertex list in McuBlockSet::d_loc (owned vertices first)
    int n_local;              // <= 1023: entries address local vertices with 10-bit slots
    int n_owned;
    int pad;
    unsigned ent_off[8];      // per warp: offset of its sliced-ELL entries in d_ent; entry k of lane l at ent_off[w] + 32 k + l
    unsigned char deg[8];     // per warp: entries per lane (the largest number of elements of one of its vertices)
};
// Entry = the shared-memory slots of the element's OTHER vertices in the element's order, 10 bits each
// (1 for a line, 2 for a triangle); 0xffffffff pads lanes with fewer elements.  Tetrahedra use 16-bit entries that walk
// the link of the owned vertex as triangle strips (ent_off then counts 16-bit entries): see mcu_block.cu.

struct McuBlockSet {
    int n_blocks = 0;
    long n_active = 0;        // vertices with at least one (selected) element
    long n_halo = 0;          // local vertices that are not owned, over all blocks
    long n_entries = 0;
    bool all_active = false;  // every vertex of the mesh is owned by some block: no rows to zero
    bool strips = false;      // tetrahedra: d_ent holds 16-bit link-strip entries
    McuBlockHdr *d_hdr = nullptr;
    int *d_loc = nullptr;
    uint32_t *d_ent = nullptr;
};

struct McuBlockPlan {          // host side of a decomposition
    std::vector<McuBlockHdr> hdr;
    std::vector<int> loc;
    std::vector<uint32_t> ent;
    long n_active = 0, n_halo = 0;
    bool all_active = false;
    bool strips = false;      // tetrahedra: ent holds 16-bit link-strip entries, two per word (see mcu_block.cu)
};
bool mcu_block_plan(int nv, int dim, int n, int nvper, const int *rix, const int *eids, const double *vert, McuBlockPlan &plan);

// n elements of nvper vertices each (eids != nullptr: the selected element ids, rix indexed by them); vert: host
// positions (nv x dim) used only to order the vertices in space.  nullptr when the mesh cannot be cut (a vertex with
// more than 255 elements or more than 1022 neighbours).
McuBlockSet *mcu_blockset_build(int nv, int dim, int n, int nvper, const int *rix, const int *eids, const double *vert, cudaStream_t st);
void mcu_blockset_free(McuBlockSet *bs);
cudaError_t mcu_block_gradient(int functional, const McuBlockSet &bs, const double *d_vert, int nv, int dim, double *d_out, unsigned *flags, cudaStream_t st);

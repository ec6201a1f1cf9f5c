// mcu_patch.h — data structures of the shared-memory patch decomposition (mcu_patch.cu).
#pragma once
#include <cstdint>
#include <vector>

#define MCU_PATCH_MAX_LOCAL 1024      // local vertex ids are 10 bit
#define MCU_PATCH_MAX_COLORS 64
#define MCU_PATCH_DEFAULT_OWNED 768

struct McuPatchHeader {
    int vtx_off;    // offset of this patch's local vertex list (owned first, then halo)
    int n_owned;
    int n_local;
    int elem_off;   // offset of the packed triangles, grouped by colour
    int n_elem;
    int color_off;  // offset of n_colors+1 relative triangle offsets
    int n_colors;
    int slice_off;  // offset of n_slices+1 relative entry offsets (slice = 32 consecutive owned vertices)
    int n_slices;
    int ent_off;    // offset of this patch's ring entries (sliced ELL of uint16: [slice][k][lane])
    int n_ent;      // entries of this patch (multiple of 32)
    int pad1;
};

#define MCU_PATCH_ENT_NONE 0xffffu    // padding entry of the sliced-ELL ring table
#define MCU_PATCH_ENT_START 0x8000u  // this ring vertex opens a new strip (no triangle with the previous entry)

struct McuPatchPlan {
    int nv = 0, nel = 0, max_colors = 0;
    std::vector<McuPatchHeader> hdr;
    std::vector<int> vtx_ids;
    std::vector<uint32_t> elems;     // a | b << 10 | c << 20 (local indices of the ascending global ids)
    std::vector<int> elem_ids;       // global triangle id of each packed slot (tests only)
    std::vector<int> color_ptr;
    std::vector<int> slice_ptr;      // per patch n_slices+1 offsets (in entries) relative to ent_off
    std::vector<uint16_t> ents;      // ring strips: 10-bit local vertex index | START flag (see mcu_patch.cu)
    int max_local = 0, max_owned = 0, max_ents = 0;
};

struct McuPatchSet {
    McuPatchPlan plan;               // headers + statistics (bulk arrays are dropped after upload)
    int n_patches = 0;
    McuPatchHeader *d_hdr = nullptr;
    int *d_vtx = nullptr;
    uint32_t *d_elems = nullptr;
    int *d_color = nullptr;
    int *d_slice = nullptr;
    uint16_t *d_ents = nullptr;
};

bool mcu_patch_plan(int nv, int nel, const int *rix, const double *vert, int max_owned, McuPatchPlan &plan);

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
    int pad;
};

struct McuPatchPlan {
    int nv = 0, nel = 0, max_colors = 0;
    std::vector<McuPatchHeader> hdr;
    std::vector<int> vtx_ids;
    std::vector<uint32_t> elems;     // a | b << 10 | c << 20 (local indices of the ascending global ids)
    std::vector<int> elem_ids;       // global triangle id of each packed slot (tests only)
    std::vector<int> color_ptr;
};

struct McuPatchSet {
    McuPatchPlan plan;               // headers + statistics (bulk arrays are dropped after upload)
    int n_patches = 0;
    McuPatchHeader *d_hdr = nullptr;
    int *d_vtx = nullptr;
    uint32_t *d_elems = nullptr;
    int *d_color = nullptr;
};

bool mcu_patch_plan(int nv, int nel, const int *rix, const double *vert, int max_owned, McuPatchPlan &plan);

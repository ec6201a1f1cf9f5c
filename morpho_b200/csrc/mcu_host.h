// mcu_host.h — host-side structures behind the opaque handles of include/morpho_cuda.h.
#pragma once
#include <cuda_runtime.h>
#include <vector>

#include "mcu_kernels.h"

struct McuPatchSet;
struct McuBlockSet;

// element → vertex incidence of one grade: the CCS arrays of mesh_getconnectivityelement(m,0,g)
// (mesh.c:296-307; cptr is implicit, e*(g+1)) plus derived device structures.
struct McuConn {
    bool present = false;
    int g = 0, nvper = 0, nel = 0;
    std::vector<int> h_rix;          // host copy, used to derive transposes / patches / selections
    int *d_rix = nullptr;
    int *d_tptr = nullptr, *d_tidx = nullptr;   // vertex → element transpose (lazily built)
    McuPatchSet *patches = nullptr;  // shared-memory patch decomposition (lazily built, triangles)
    bool patch_failed = false;       // the planner refused this connectivity: do not plan again until it changes
    McuBlockSet *blocks = nullptr;   // block decomposition of the generic owner-computes gradient kernel (mcu_block.cu), lazily built
    bool block_failed = false;
};

struct mcu_mesh {
    int dim = 0, nv = 0;
    double *d_vert = nullptr;        // the vertex store the kernels read: the mesh's own buffer or a bound mcu_matrix
    double *d_vert_own = nullptr;    // the mesh's own buffer
    struct mcu_matrix *bound = nullptr;
    McuConn conn[4];
    long vert_version = 0, conn_version = 0;
    CompSum *d_partials = nullptr;
    double *d_scalar = nullptr;
    unsigned *d_flags = nullptr;
    double *d_result = nullptr;      // staging for mcu_eval (host-out) results
    size_t result_cap = 0;
    struct mcu_halo *halo = nullptr; // multi-GPU: exchange whose pushes ride in the patch gradient kernel (mcu_halo_bind)
    int halo_slot = 0;
};

struct mcu_field {
    mcu_mesh *m = nullptr;
    int psize = 0;
    double *d_data = nullptr;
};

struct mcu_selection {
    mcu_mesh *m = nullptr;
    int g = 0;
    std::vector<int> ids;            // sorted, unique
    int *d_eids = nullptr;
    int *d_tptr = nullptr, *d_tidx = nullptr;   // transpose restricted to the selected elements
    unsigned char *d_mask = nullptr; // grade 0 only
    long conn_version = -1;
    McuBlockSet *blocks = nullptr;   // block decomposition restricted to the selected elements (lazily built)
    bool block_failed = false;
    long blocks_version = -1;
};

void mcu_set_error(const char *fmt, ...);
cudaStream_t mcu_stream();
int mcu_init_if_needed();
long mcu_opt(const char *key, long dflt);
void mcu_host_transpose(int nv, int n, int nvper, const int *rix, const int *eids, std::vector<int> &tptr, std::vector<int> &tidx);

// patch decomposition (mcu_patch.cu)
McuPatchSet *mcu_patchset_build(int nv, int nel, const int *h_rix, const double *h_vert, cudaStream_t st);
void mcu_patchset_free(McuPatchSet *ps);
extern "C" int mcu_ensure_transpose(mcu_mesh *m, int g);       // builds conn[g].d_tptr / d_tidx if absent
extern "C" int mcu_ensure_selection(mcu_selection *s);   // uploads ids / transpose / mask of a selection if stale
int mcu_ensure_patches(mcu_mesh *m, int g);        // builds conn[g].patches if absent (MCU_ERR_UNSUPPORTED if it cannot)
struct McuHaloPush;
bool mcu_halo_push_params(struct mcu_halo *h, int slot, int dim, McuHaloPush *out);
void mcu_halo_detach_mesh(mcu_mesh *m);

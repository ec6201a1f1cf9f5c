// mcu_kernels.h — host-callable launchers implemented by the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include "mcu_device.cuh"

// Analytic functionals (Length, AreaEnclosed[total/integrand], Area, VolumeEnclosed, Volume); FMA on.
cudaError_t mcu_fast_total(const McuParams &p, const McuView &v, CompSum *partials, double *d_out, cudaStream_t st);
cudaError_t mcu_fast_integrand(const McuParams &p, const McuView &v, double *d_out, cudaStream_t st);
cudaError_t mcu_fast_gradient(const McuParams &p, const McuView &v, double *d_out, unsigned *flags, cudaStream_t st);
bool mcu_fast_supports(int functional, int mode);

// Finite-difference family and field functionals; compiled with -fmad=false (MCU_STRICT).
cudaError_t mcu_strict_total(const McuParams &p, const McuView &v, CompSum *partials, double *d_out, cudaStream_t st);
cudaError_t mcu_strict_integrand(const McuParams &p, const McuView &v, double *d_out, cudaStream_t st);
cudaError_t mcu_strict_gradient(const McuParams &p, const McuView &v, double *d_out, unsigned *flags, cudaStream_t st);
cudaError_t mcu_strict_fieldgradient(const McuParams &p, const McuView &v, double *d_out, cudaStream_t st);
bool mcu_strict_supports(int functional, int mode);

// Shared-memory patch kernel for triangle meshes (Area / VolumeEnclosed gradients).
struct McuPatchSet;
struct McuHaloPush;
cudaError_t mcu_patch_gradient(int functional, const McuPatchSet &ps, const double *d_vert, double *d_out,
                               unsigned *flags, cudaStream_t st, const McuHaloPush *push);
cudaError_t mcu_patch_total(int functional, const McuPatchSet &ps, const double *d_vert, double *d_out, unsigned *flags, cudaStream_t st);
int mcu_patchset_attach_push(McuPatchSet *ps, int nv, int n, const int *vertex, const int *peer, const int *k, cudaStream_t st);

// All-pairs kernels.
cudaError_t mcu_pairwise_launch(int n, int dim, const double *d_x, int kind, const double *params, double cutoff, int mode,
                                double *d_out, CompSum *partials, cudaStream_t st, int row0 = 0, int nrows = -1, int *status = nullptr);
int mcu_pairwise_store_program(int nf, const mcu_expr_op *pf, int ng, const mcu_expr_op *pg);

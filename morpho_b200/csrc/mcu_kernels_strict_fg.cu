// mcu_kernels_strict_fg.cu — the field-gradient maps of the finite-difference family (functional.c:1305-1503), split
// from mcu_kernels_strict.cu only to halve the compile time of that translation unit.  Same flags (-fmad=false).
#define MCU_STRICT 1
#include "mcu_kernels_impl.cuh"

cudaError_t mcu_strict_fieldgradient(const McuParams &p, const McuView &v, double *d_out, cudaStream_t st) {
    switch (p.functional) {
        case MCU_F_GRADSQ: return run_fd_fieldgrad<MCU_F_GRADSQ>(p, v, d_out, st);
        case MCU_F_NORMSQ: return run_fd_fieldgrad<MCU_F_NORMSQ>(p, v, d_out, st);
        case MCU_F_NEMATIC: return run_fd_fieldgrad<MCU_F_NEMATIC>(p, v, d_out, st);
        case MCU_F_NEMATICELECTRIC: return run_fd_fieldgrad<MCU_F_NEMATICELECTRIC>(p, v, d_out, st);
    }
    return cudaErrorInvalidValue;
}

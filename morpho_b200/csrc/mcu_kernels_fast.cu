// mcu_kernels_fast.cu — analytic functionals on the generic map kernels (FMA contraction on).
#include "mcu_kernels_impl.cuh"

bool mcu_fast_supports(int f, int mode) {
    switch (f) {
        case MCU_F_LENGTH: case MCU_F_AREA: case MCU_F_VOLUMEENCLOSED: case MCU_F_VOLUME:
            return mode == MCU_TOTAL || mode == MCU_INTEGRAND || mode == MCU_GRADIENT;
        case MCU_F_AREAENCLOSED:
            return mode == MCU_TOTAL || mode == MCU_INTEGRAND;
    }
    return false;
}

#define DISPATCH(call)                                                   \
    switch (p.functional) {                                              \
        case MCU_F_LENGTH: return call(MCU_F_LENGTH);                    \
        case MCU_F_AREAENCLOSED: return call(MCU_F_AREAENCLOSED);        \
        case MCU_F_AREA: return call(MCU_F_AREA);                        \
        case MCU_F_VOLUMEENCLOSED: return call(MCU_F_VOLUMEENCLOSED);    \
        case MCU_F_VOLUME: return call(MCU_F_VOLUME);                    \
    }                                                                    \
    return cudaErrorInvalidValue;

cudaError_t mcu_fast_total(const McuParams &p, const McuView &v, CompSum *partials, double *d_out, cudaStream_t st) {
#define CALL(F) run_total<F>(p, v, partials, d_out, st)
    DISPATCH(CALL)
#undef CALL
}

cudaError_t mcu_fast_integrand(const McuParams &p, const McuView &v, double *d_out, cudaStream_t st) {
#define CALL(F) run_integrand<F>(p, v, d_out, st)
    DISPATCH(CALL)
#undef CALL
}

cudaError_t mcu_fast_gradient(const McuParams &p, const McuView &v, double *d_out, unsigned *flags, cudaStream_t st) {
    switch (p.functional) {
        case MCU_F_LENGTH: return run_grad_gather<MCU_F_LENGTH>(p, v, d_out, flags, st);
        case MCU_F_AREA: return run_grad_gather<MCU_F_AREA>(p, v, d_out, flags, st);
        case MCU_F_VOLUMEENCLOSED: return run_grad_gather<MCU_F_VOLUMEENCLOSED>(p, v, d_out, flags, st);
        case MCU_F_VOLUME: return run_grad_gather<MCU_F_VOLUME>(p, v, d_out, flags, st);
    }
    return cudaErrorInvalidValue;
}

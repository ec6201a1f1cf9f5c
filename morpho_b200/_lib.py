"""ctypes binding of libmorpho_cuda.so (include/morpho_cuda.h).

There is no fallback: if the shared library is missing, or no sm_100 device is
usable, every compute call raises.  Nothing here imports the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MORPHO_CUDA_LIB", os.path.join(HERE, "lib", "libmorpho_cuda.so"))   # env: kernel-variant experiments

MCU_OK, MCU_ERR_ARGS, MCU_ERR_ELNOTFOUND, MCU_ERR_DEGENERATE, MCU_ERR_VOLENCLZERO, \
    MCU_ERR_UNSUPPORTED, MCU_ERR_CUDA, MCU_ERR_ALLOC = range(8)
MCU_TOTAL, MCU_INTEGRAND, MCU_GRADIENT, MCU_FIELDGRADIENT = range(4)
MCU_PATH_AUTO, MCU_PATH_GATHER, MCU_PATH_PATCH, MCU_PATH_BLOCK = range(4)
MCU_PW_COULOMB, MCU_PW_LAURENT, MCU_PW_YUKAWA, MCU_PW_PROGRAM = range(4)

FUNCTIONAL_IDS = {
    "Length": 1, "AreaEnclosed": 2, "Area": 3, "VolumeEnclosed": 4, "Volume": 5,
    "LineCurvatureSq": 6, "GradSq": 7, "NormSq": 8, "Nematic": 9, "NematicElectric": 10, "EquiElement": 11,
}


class McuFunctional(C.Structure):
    _fields_ = [("functional", C.c_int), ("grade", C.c_int),
                ("ksplay", C.c_double), ("ktwist", C.c_double), ("kbend", C.c_double), ("pitch", C.c_double),
                ("haspitch", C.c_int), ("integrandonly", C.c_int),
                ("field", C.c_void_p * 2), ("wrt", C.c_int),
                ("host_weight", C.c_void_p), ("nweight", C.c_int)]


class MorphoError(RuntimeError):
    """A Morpho runtime error: ``id`` is the reference's error id (functional.h:97-188)."""

    def __init__(self, status, ident, msg):
        super().__init__(f"Error '{ident}': {msg}" if ident else f"libmorpho_cuda status {status}: {msg}")
        self.status, self.id, self.msg = status, ident, msg


_lib = None


def lib():
    """Load (once) and return the C-ABI library; fail loudly when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C morpho_b200/csrc). morpho_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    sig = {
        "mcu_abi_version": (C.c_int, []),
        "mcu_init": (C.c_int, [C.c_int]),
        "mcu_finalize": (C.c_int, []),
        "mcu_last_error": (C.c_char_p, []),
        "mcu_error_id": (C.c_char_p, [C.c_int]),
        "mcu_set_option": (C.c_int, [C.c_char_p, C.c_long]),
        "mcu_get_option": (C.c_long, [C.c_char_p]),
        "mcu_sync": (C.c_int, []),
        "mcu_set_stream": (C.c_int, [vp]),
        "mcu_mesh_create": (vp, [C.c_int, C.c_int]),
        "mcu_mesh_destroy": (C.c_int, [vp]),
        "mcu_mesh_set_vertices": (C.c_int, [vp, vp]),
        "mcu_mesh_set_vertices_device": (C.c_int, [vp, vp]),
        "mcu_mesh_get_vertices": (C.c_int, [vp, vp]),
        "mcu_mesh_vertices_device": (vp, [vp]),
        "mcu_mesh_set_connectivity": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "mcu_mesh_counts": (C.c_int, [vp, ip, ip, ip]),
        "mcu_mesh_drop_connectivity": (C.c_int, [vp, C.c_int]),
        "mcu_mesh_addgrade": (C.c_int, [vp, C.c_int, ip]),
        "mcu_mesh_boundary": (C.c_int, [vp, ip, ip, vp, C.c_int]),
        "mcu_mesh_lower": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, ip, ip, ip]),
        "mcu_integral": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, C.c_int, vp]),
        "mcu_debug_patch_triangles": (C.c_int, [C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, vp]),
        "mcu_pointwise": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_int, vp]),
        "mcu_integral_gradient": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, C.c_int, vp]),
        "mcu_mesh_get_connectivity": (C.c_int, [vp, C.c_int, vp]),
        "mcu_mesh_get_transpose": (C.c_int, [vp, C.c_int, vp, vp]),
        "mcu_field_create": (vp, [vp, C.c_int]),
        "mcu_field_destroy": (C.c_int, [vp]),
        "mcu_field_set": (C.c_int, [vp, vp]),
        "mcu_field_set_device": (C.c_int, [vp, vp]),
        "mcu_field_get": (C.c_int, [vp, vp]),
        "mcu_field_device": (vp, [vp]),
        "mcu_selection_create": (vp, [vp, C.c_int, C.c_int, vp]),
        "mcu_selection_destroy": (C.c_int, [vp]),
        "mcu_selection_get_ids": (C.c_int, [vp, vp, C.c_int]),
        "mcu_eval": (C.c_int, [vp, C.POINTER(McuFunctional), vp, C.c_int, vp]),
        "mcu_eval_device": (C.c_int, [vp, C.POINTER(McuFunctional), vp, C.c_int, vp]),
        "mcu_result_size": (C.c_long, [vp, C.POINTER(McuFunctional), C.c_int]),
        "mcu_mesh_status": (C.c_int, [vp]),
        "mcu_matrix_create": (vp, [C.c_int, C.c_int]),
        "mcu_matrix_destroy": (C.c_int, [vp]),
        "mcu_matrix_dims": (C.c_int, [vp, ip, ip]),
        "mcu_matrix_device": (vp, [vp]),
        "mcu_matrix_upload": (C.c_int, [vp, vp]),
        "mcu_matrix_download": (C.c_int, [vp, vp]),
        "mcu_matrix_copy": (C.c_int, [vp, vp]),
        "mcu_matrix_zero": (C.c_int, [vp]),
        "mcu_matrix_axpby": (C.c_int, [vp, vp, C.c_double, vp, C.c_double]),
        "mcu_matrix_xpa": (C.c_int, [vp, vp, C.c_double, C.c_double]),
        "mcu_matrix_inner": (C.c_int, [vp, vp, dp]),
        "mcu_matrix_norm": (C.c_int, [vp, dp]),
        "mcu_matrix_sum": (C.c_int, [vp, dp]),
        "mcu_matrix_set_column": (C.c_int, [vp, C.c_int, vp]),
        "mcu_matrix_get_column": (C.c_int, [vp, C.c_int, vp]),
        "mcu_matrix_zero_columns": (C.c_int, [vp, C.c_int, vp]),
        "mcu_mesh_bind_vertices": (C.c_int, [vp, vp]),
        "mcu_eval_matrix": (C.c_int, [vp, C.POINTER(McuFunctional), vp, C.c_int, vp]),
        "mcu_pairwise_program": (C.c_int, [C.c_int, vp, C.c_int, vp]),
        "mcu_pairwise": (C.c_int, [C.c_int, C.c_int, vp, C.c_int, vp, C.c_double, C.c_int, vp]),
        "mcu_pairwise_device": (C.c_int, [C.c_int, C.c_int, vp, C.c_int, vp, C.c_double, C.c_int, vp]),
        "mcu_pairwise_rows_device": (C.c_int, [C.c_int, C.c_int, vp, C.c_int, vp, C.c_double, C.c_int, C.c_int, C.c_int, vp]),
        "mcu_binbounds": (C.c_int, [C.c_int, C.c_int, vp]),
        "mcu_timer_start": (C.c_int, []),
        "mcu_timer_stop": (C.c_int, [C.POINTER(C.c_float)]),
        "mcu_launch_count": (C.c_long, []),
        "mcu_device_malloc": (C.c_int, [C.POINTER(vp), C.c_size_t]),
        "mcu_device_free": (C.c_int, [vp]),
        "mcu_flush_l2": (C.c_int, []),
        "mcu_debug_dfma_peak": (C.c_int, [dp]),
        "mcu_halo_create": (vp, [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp]),
        "mcu_halo_export": (C.c_int, [vp, vp]),
        "mcu_halo_connect": (C.c_int, [vp, vp]),
        "mcu_halo_exchange": (C.c_int, [vp, C.c_int, vp, C.c_int]),
        "mcu_halo_exchange_async": (C.c_int, [vp, C.c_int, vp, C.c_int]),
        "mcu_halo_join": (C.c_int, [vp]),
        "mcu_halo_status": (C.c_int, [vp]),
        "mcu_halo_attach": (C.c_int, [vp, vp]),
        "mcu_halo_bind": (C.c_int, [vp, C.c_int]),
        "mcu_halo_destroy": (C.c_int, [vp]),
        "mcu_debug_patch_profile": (C.c_int, [vp]),
        "mcu_debug_quadrule": (C.c_int, [C.c_int, vp, ip, vp, vp, C.c_int]),
        "mcu_debug_quadconfigure": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_char_p, ip]),
        "mcu_integral_method": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, C.c_int, vp, vp]),
        "mcu_integral_gradient_method": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, C.c_int, vp, vp]),
        "mcu_debug_patch_plan": (C.c_int, [C.c_int, C.c_int, vp, vp, C.c_int, C.POINTER(C.c_long), vp, vp, vp]),
        "mcu_debug_block_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.POINTER(C.c_long), vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    if L.mcu_abi_version() != 1:
        raise ImportError("libmorpho_cuda.so ABI version mismatch")
    _lib = L
    return L


def check(status):
    """Translate a non-zero status into MorphoError (ids as in functional.h:97-188)."""
    if status == MCU_OK:
        return
    L = lib()
    raise MorphoError(status, L.mcu_error_id(status).decode(), L.mcu_last_error().decode())


def init(device=None):
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", os.environ.get("MORPHO_CUDA_DEVICE", "0")))
    check(lib().mcu_init(int(device)))


def set_option(key, value):
    check(lib().mcu_set_option(key.encode(), int(value)))

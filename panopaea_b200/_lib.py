"""ctypes binding of libpanopaea_b200.so (the C ABI in include/panopaea_b200.h).

There is no CPU fallback: if the shared library is missing this module raises, and every
compute entry point of the library itself fails with PANO_ERR_CUDA when no device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "lib", "libpanopaea_b200.so")

OK, ERR_INVALID, ERR_SHAPE, ERR_CUDA, ERR_UNIMPLEMENTED, ERR_TIMEOUT, ERR_COMM = range(7)
F64, F32 = 0, 1
SIMPLEX0, SIMPLEX1, SIMPLEX2 = 0, 1, 2
COMP_ALL, COMP_VY, COMP_VX = 0, 1, 2
PRECOND_IDENTITY, PRECOND_JACOBI, PRECOND_MULTIGRID = 0, 1, 2


class PanoError(RuntimeError):
    """Raised for any non-zero return code; the Rust shim would panic! here, as the reference does."""

    def __init__(self, code, message):
        super().__init__(f"[pano error {code}] {message}")
        self.code = code


class Rect(C.Structure):
    _fields_ = [("y0", C.c_int64), ("y1", C.c_int64), ("x0", C.c_int64), ("x1", C.c_int64)]


class PcgInfo(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("applies", C.c_int32), ("final_residual", C.c_double), ("rhs_max", C.c_double)]

    def as_dict(self):
        return dict(iterations=self.iterations, applies=self.applies, final_residual=self.final_residual, rhs_max=self.rhs_max)


class StepParams(C.Structure):
    _fields_ = [("timestep", C.c_double), ("threshold", C.c_double), ("max_iterations", C.c_int32), ("precond", C.c_int32),
                ("inflow", Rect), ("inflow_density", C.c_double), ("inflow_vy", C.c_double), ("obstacle", Rect)]


class Box(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("z0", "z1", "y0", "y1", "x0", "x1")]


class Step3Params(C.Structure):
    _fields_ = [("timestep", C.c_double), ("threshold", C.c_double), ("max_iterations", C.c_int32), ("precond", C.c_int32),
                ("inflow", Box), ("inflow_density", C.c_double), ("inflow_vy", C.c_double), ("obstacle", Box)]


_P = C.c_void_p
_SIGS = {
    # name: (restype, argtypes)
    "pano_version": (C.c_char_p, []),
    "pano_last_error": (C.c_char_p, []),
    "pano_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pano_ctx_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "pano_ctx_destroy": (C.c_int, [_P]),
    "pano_ctx_sync": (C.c_int, [_P]),
    "pano_ctx_stream": (C.c_int, [_P, C.POINTER(_P)]),
    "pano_ctx_num_sms": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "pano_ctx_launch_count": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "pano_timer_start": (C.c_int, [_P]),
    "pano_timer_stop_ms": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "pano_timer_mark": (C.c_int, [_P]),
    "pano_timer_marks_ms": (C.c_int, [_P, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]),
    "pano_ctx_step_times": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "pano_ctx_cg_profile": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "pano_ctx_cg_profile_ctas": (C.c_int, [_P, C.POINTER(C.c_int64), C.c_int]),
    "pano_ctx_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "pano_ctx_get_option": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_int64)]),
    "pano_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "pano_host_free": (C.c_int, [_P]),
    "pano_field_new": (C.c_int, [_P, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.POINTER(_P)]),
    "pano_field_free": (C.c_int, [_P]),
    "pano_field_num_elem": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t)]),
    "pano_field_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "pano_field_device_ptr": (C.c_int, [_P, C.POINTER(_P)]),
    "pano_field_upload": (C.c_int, [_P, _P, C.c_size_t]),
    "pano_field_download": (C.c_int, [_P, _P, C.c_size_t]),
    "pano_field_fill": (C.c_int, [_P, C.c_double]),
    "pano_field_fill_rect": (C.c_int, [_P, C.c_int, Rect, C.c_double]),
    "pano_field_assign": (C.c_int, [_P, _P]),
    "pano_field_swap": (C.c_int, [_P, _P]),
    "pano_field_scaled_add": (C.c_int, [_P, C.c_double, _P]),
    "pano_field_scale": (C.c_int, [_P, C.c_double]),
    "pano_field_xpby": (C.c_int, [_P, _P, C.c_double]),
    "pano_field_dot": (C.c_int, [_P, _P, C.POINTER(C.c_double)]),
    "pano_field_norm_max": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "pano_hodge_0_primal": (C.c_int, [_P, _P]),
    "pano_hodge_2_dual": (C.c_int, [_P, _P]),
    "pano_hodge_1_primal": (C.c_int, [_P, _P]),
    "pano_hodge_1_dual": (C.c_int, [_P, _P]),
    "pano_hodge_2_primal": (C.c_int, [_P, _P]),
    "pano_hodge_0_dual": (C.c_int, [_P, _P]),
    "pano_derivative_0_primal": (C.c_int, [_P, _P]),
    "pano_derivative_1_primal": (C.c_int, [_P, _P]),
    "pano_derivative_0_dual": (C.c_int, [_P, _P]),
    "pano_derivative_1_dual": (C.c_int, [_P, _P]),
    "pano_advect": (C.c_int, [_P, _P, C.c_double, _P]),
    "pano_advect_mac": (C.c_int, [_P, _P, C.c_double, _P]),
    "pano_advect_all": (C.c_int, [_P, _P, _P, _P, C.c_double]),
    "pano_neg_divergence": (C.c_int, [_P, _P, Rect, C.POINTER(C.c_double)]),
    "pano_laplacian_apply": (C.c_int, [_P, _P, C.c_double, Rect]),
    "pano_project": (C.c_int, [_P, _P, C.c_double]),
    "pano_pcg_solve": (C.c_int, [C.c_int, _P, _P, C.c_int32, C.c_double, _P, _P, _P, C.c_double, Rect, C.POINTER(PcgInfo)]),
    "pano_jacobi_apply": (C.c_int, [_P, _P, C.c_double, Rect]),
    "pano_mg_create": (C.c_int, [_P, C.c_size_t, C.c_size_t, C.c_double, Rect, C.POINTER(_P)]),
    "pano_mg_destroy": (C.c_int, [_P]),
    "pano_mg_apply": (C.c_int, [_P, _P, _P]),
    "pano_mg_levels": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pano_fluid_step": (C.c_int, [C.POINTER(StepParams), _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(PcgInfo)]),
    "pano_fluid_step_host": (C.c_int, [_P, C.POINTER(StepParams), C.c_size_t, C.c_size_t, _P, _P, _P, C.POINTER(PcgInfo)]),
    "pano_density_to_u8": (C.c_int, [_P, C.c_double, C.c_double, _P]),
    "pano_slab_range": (C.c_int, [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "pano_dist_create": (C.c_int, [_P, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.POINTER(StepParams), C.POINTER(_P)]),
    "pano_dist_destroy": (C.c_int, [_P]),
    "pano_dist_window": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "pano_dist_ipc_handle": (C.c_int, [_P, _P]),
    "pano_dist_connect": (C.c_int, [_P, C.c_int, _P]),
    "pano_dist_set_max_ctas": (C.c_int, [_P, C.c_int]),
    "pano_dist_upload": (C.c_int, [_P, C.c_int, _P]),
    "pano_dist_download": (C.c_int, [_P, C.c_int, _P, C.POINTER(C.c_size_t)]),
    "pano_dist_step": (C.c_int, [_P]),
    "pano_dist_solve": (C.c_int, [_P]),
    "pano_dist_sync": (C.c_int, [_P, C.POINTER(PcgInfo)]),
    "pano_dist_step_host": (C.c_int, [_P, _P, _P, _P, C.POINTER(PcgInfo)]),
    "pano_field3_new": (C.c_int, [_P, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(_P)]),
    "pano_field3_num_elem": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t)]),
    "pano_field3_dim": (C.c_int, [_P, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "pano_field3_fill_box": (C.c_int, [_P, C.c_int, Box, C.c_double]),
    "pano_trilinear": (C.c_double, [C.c_double] * 11),
    "pano_advect3": (C.c_int, [_P, _P, C.c_double, _P]),
    "pano_advect3_mac": (C.c_int, [_P, _P, C.c_double, _P]),
    "pano_advect3_all": (C.c_int, [_P, _P, _P, _P, C.c_double]),
    "pano_neg_divergence3": (C.c_int, [_P, _P, Box, C.POINTER(C.c_double)]),
    "pano_laplacian3_apply": (C.c_int, [_P, _P, C.c_double, Box]),
    "pano_project3": (C.c_int, [_P, _P, C.c_double]),
    "pano_pcg3_solve": (C.c_int, [C.c_int, _P, _P, C.c_int32, C.c_double, _P, _P, _P, C.c_double, Box, C.POINTER(PcgInfo)]),
    "pano_fluid3_step": (C.c_int, [C.POINTER(Step3Params), _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(PcgInfo)]),
    "pano_fluid3_step_host": (C.c_int, [_P, C.POINTER(Step3Params), C.c_size_t, C.c_size_t, C.c_size_t, _P, _P, _P, C.POINTER(PcgInfo)]),
}
IPC_HANDLE_BYTES = 64

_lib = None


def exported_names():
    return sorted(_SIGS)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -m panopaea_b200.build` (nvcc, sm_100a). "
            "panopaea_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)          # AttributeError here = the .so is stale against the header
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise PanoError(rc, load().pano_last_error().decode("utf-8", "replace"))

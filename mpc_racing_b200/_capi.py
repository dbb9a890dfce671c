"""ctypes binding of include/mpcr.h (the C ABI of libmpcr.so).

This is the whole host<->device boundary: plain pointers and sizes, no torch
types.  The library is built in-tree by ``__graft_entry__.build()`` /
``mpc_racing_b200.build``; importing the solver without it fails loudly - there
is no CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpcr.so")


class MpcrConfig(C.Structure):
    """Mirror of ``struct mpcr_config`` (include/mpcr.h)."""
    _fields_ = [
        ("N", C.c_int32), ("model", C.c_int32), ("wrap_constraint", C.c_int32),
        ("lane_constraint", C.c_int32), ("device", C.c_int32),
        ("pointers_on_device", C.c_int32), ("max_iter", C.c_int32),
        ("acceptable_iter", C.c_int32),
        ("Ts", C.c_double),
        ("lambda_s", C.c_double), ("alpha_L", C.c_double),
        ("min_steer", C.c_double), ("max_steer", C.c_double), ("min_throttle", C.c_double),
        ("max_steer_delta", C.c_double), ("min_steer_delta", C.c_double),
        ("max_throttle_delta", C.c_double), ("min_throttle_delta", C.c_double),
        ("q_v_max", C.c_double), ("v_max", C.c_double), ("min_s_delta", C.c_double),
        ("alpha_c", C.c_double), ("d_max", C.c_double), ("q_v_y", C.c_double),
        ("n_exp", C.c_double), ("beta_delta", C.c_double),
        ("tol", C.c_double), ("acceptable_tol", C.c_double),
        ("dual_inf_tol", C.c_double), ("constr_viol_tol", C.c_double),
        ("compl_inf_tol", C.c_double), ("mu_init", C.c_double),
        ("solve_kernel", C.c_int32), ("integrator", C.c_int32),
        ("restoration", C.c_int32), ("schedule", C.c_int32),
        ("polish_tol", C.c_double),
    ]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
SOLVE_ARGTYPES = [C.c_void_p] * 8 + [C.c_void_p] * 9   # 8 inputs, 9 outputs (raw addresses)

# every symbol include/mpcr.h declares
EXPORTED_SYMBOLS = [
    "mpcr_version", "mpcr_config_default", "mpcr_create", "mpcr_destroy", "mpcr_last_error",
    "mpcr_lam_g_stride", "mpcr_solve_batch", "mpcr_solve_batch_async", "mpcr_solve_batch_trace", "mpcr_last_kernel_ms",
    "mpcr_launch_info", "mpcr_last_launch_count", "mpcr_fp64_peak_tflops",
    "mpcr_track_create", "mpcr_track_destroy", "mpcr_track_eval", "mpcr_track_coeffs",
    "mpcr_track_project", "mpcr_track_lookup_error", "mpcr_plant_step", "mpcr_model_eval",
    "mpcr_rollout_prepare", "mpcr_rollout_apply", "mpcr_set_dispatch_keys",
]


class MpcrRolloutBufs(C.Structure):
    """Mirror of ``struct mpcr_rollout_bufs`` (include/mpcr.h): device pointers as raw addresses."""
    _fields_ = [
        ("R", C.c_int32), ("N", C.c_int32), ("n_seg", C.c_int32), ("plant_model", C.c_int32),
        ("estimate_state", C.c_int32), ("start_control_at", C.c_int32),
        ("Ts", C.c_double), ("window", C.c_double), ("lookback", C.c_double), ("lookahead", C.c_double),
        ("target", C.c_double),
        ("x", C.c_void_p), ("x_mpc", C.c_void_p), ("progress", C.c_void_p), ("yaw_prev", C.c_void_p),
        ("u_prev", C.c_void_p), ("cx", C.c_void_p), ("cy", C.c_void_p), ("dist", C.c_void_p),
        ("sign", C.c_void_p), ("travelled", C.c_void_p), ("lap_time", C.c_void_p), ("max_err", C.c_void_p),
        ("seg_bounds", C.c_void_p), ("seg_t", C.c_void_p), ("steps", C.c_void_p), ("fails", C.c_void_p),
        ("done", C.c_void_p), ("counters", C.c_void_p),
    ]

_lib = None


class MpcrError(RuntimeError):
    pass


def load_library(path=None):
    """Load libmpcr.so; raises MpcrError when it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("MPCR_LIB") or LIB_PATH       # MPCR_LIB: development builds side by side
    if not os.path.exists(p):
        raise MpcrError(
            f"{p} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
            "g.build()').  mpc_racing_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    lib.mpcr_version.restype = C.c_int
    lib.mpcr_config_default.argtypes = [C.POINTER(MpcrConfig)]
    lib.mpcr_config_default.restype = None
    lib.mpcr_create.argtypes = [C.POINTER(MpcrConfig), C.POINTER(C.c_void_p)]
    lib.mpcr_create.restype = C.c_int
    lib.mpcr_destroy.argtypes = [C.c_void_p]
    lib.mpcr_destroy.restype = C.c_int
    lib.mpcr_last_error.argtypes = [C.c_void_p]
    lib.mpcr_last_error.restype = C.c_char_p
    lib.mpcr_lam_g_stride.argtypes = [C.c_void_p]
    lib.mpcr_lam_g_stride.restype = C.c_int
    lib.mpcr_solve_batch.argtypes = [C.c_void_p, C.c_int] + SOLVE_ARGTYPES + [C.c_void_p]
    lib.mpcr_solve_batch.restype = C.c_int
    lib.mpcr_solve_batch_async.argtypes = [C.c_void_p, C.c_int] + SOLVE_ARGTYPES + [C.c_void_p]
    lib.mpcr_solve_batch_async.restype = C.c_int
    lib.mpcr_solve_batch_trace.argtypes = [C.c_void_p, C.c_int] + SOLVE_ARGTYPES + [C.c_void_p, C.c_void_p]
    lib.mpcr_solve_batch_trace.restype = C.c_int
    lib.mpcr_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.mpcr_last_kernel_ms.restype = C.c_int
    lib.mpcr_launch_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.mpcr_launch_info.restype = C.c_int
    lib.mpcr_last_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    lib.mpcr_last_launch_count.restype = C.c_int
    lib.mpcr_fp64_peak_tflops.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.mpcr_fp64_peak_tflops.restype = C.c_int
    V = C.c_void_p
    lib.mpcr_track_create.argtypes = [V, C.c_int, V, V, V, C.c_int, V, V, C.POINTER(V)]
    lib.mpcr_track_create.restype = C.c_int
    lib.mpcr_track_destroy.argtypes = [V]
    lib.mpcr_track_destroy.restype = C.c_int
    lib.mpcr_track_eval.argtypes = [V, C.c_int, V, V, V]
    lib.mpcr_track_eval.restype = C.c_int
    lib.mpcr_track_coeffs.argtypes = [V, C.c_int, V, C.c_double, V, V, V]
    lib.mpcr_track_coeffs.restype = C.c_int
    lib.mpcr_track_project.argtypes = [V, C.c_int, V, V, V, V, V, V, V, V]
    lib.mpcr_track_project.restype = C.c_int
    lib.mpcr_track_lookup_error.argtypes = [V, C.c_int, V, C.c_double, V, V]
    lib.mpcr_track_lookup_error.restype = C.c_int
    lib.mpcr_plant_step.argtypes = [V, C.c_int, C.c_int, V, V, C.c_double, V, V]
    lib.mpcr_plant_step.restype = C.c_int
    lib.mpcr_model_eval.argtypes = [V, C.c_int, C.c_int, C.c_int, V, V, C.c_double, V, V, V, V, V]
    lib.mpcr_model_eval.restype = C.c_int
    lib.mpcr_set_dispatch_keys.argtypes = [V, V]
    lib.mpcr_set_dispatch_keys.restype = C.c_int
    lib.mpcr_rollout_prepare.argtypes = [V, C.POINTER(MpcrRolloutBufs), V]
    lib.mpcr_rollout_prepare.restype = C.c_int
    lib.mpcr_rollout_apply.argtypes = [V, C.POINTER(MpcrRolloutBufs), V, V, V]
    lib.mpcr_rollout_apply.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def default_config(lib=None):
    cfg = MpcrConfig()
    (lib or load_library()).mpcr_config_default(C.byref(cfg))
    return cfg

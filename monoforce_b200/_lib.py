"""ctypes binding of the C ABI in include/monoforce_b200.h.

The shared library is REQUIRED: there is no CPU or eager-PyTorch fallback anywhere in
this package.  If it is missing, `load()` raises with the build command.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# MFB_LIB_PATH lets kernel experiments load an alternative build of the SAME ABI (development only)
LIB_PATH = os.environ.get("MFB_LIB_PATH") or os.path.join(_PKG, "libmonoforce_b200.so")

MFB_F32, MFB_F64 = 0, 1
MFB_STEP_LOOP, MFB_ODEINT_EULER = 0, 1
MFB_PHYSICS_LOSS_MAX_BLOCKS = 4096

EXPORTED_SYMBOLS = (
    "mfb_rollout_workspace_bytes", "mfb_rollout_forward", "mfb_rollout_backward", "mfb_rollout_forward_host",
    "mfb_lift_splat_forward", "mfb_lift_splat_backward", "mfb_lift_splat_forward_bf16", "mfb_conv_bn_act_bf16",
    "mfb_conv2d_bf16", "mfb_upsample_concat_nhwc_bf16", "mfb_stem_conv_bf16", "mfb_dwconv_bn_silu_bf16", "mfb_se_fold_bf16",
    "mfb_cast_f32_to_bf16", "mfb_terrain_postproc", "mfb_path_postproc", "mfb_physics_loss",
    "mfb_last_error", "mfb_abi_version", "mfb_kernel_launches", "mfb_release_scratch",
)


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("N", "H", "W", "Cin", "Ho", "Wo", "Cout", "KH", "KW", "stride", "pad_h", "pad_w",
                                         "act", "per_image_weights", "n_heads")] + \
               [("head_act", C.c_int32 * 4), ("head_bias", C.c_float * 4), ("head_lo", C.c_float * 4), ("head_hi", C.c_float * 4)]


class RolloutDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("T", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("n_tracks", C.c_int32), ("variant", C.c_int32), ("traj_per_map", C.c_int32),
        ("map_stride", C.c_int64),
        ("mass", C.c_double), ("gravity", C.c_double), ("stiffness", C.c_double), ("damping", C.c_double),
        ("grid_res", C.c_double), ("d_max", C.c_double), ("dt", C.c_double), ("omega_max", C.c_double),
        ("robot_Ly", C.c_double),
        ("I_inv", C.c_double * 9),
        ("joint_positions", C.c_double * 12),
    ]


_IN = ("z_grid", "friction", "controls", "x0", "xd0", "R0", "omega0", "points", "part_id", "ts", "joint_angles")
_OUT = ("Xs", "Xds", "Rs", "Omegas", "F_springs", "F_frictions", "x0z", "cost", "contact_sum")


class RolloutBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _IN + _OUT] + [("workspace", C.c_void_p), ("workspace_bytes", C.c_int64)]


_GIN = ("g_Xs", "g_Xds", "g_Rs", "g_Omegas", "g_F_springs", "g_F_frictions", "g_x0z")
_GOUT = ("g_z_grid", "g_friction", "g_controls", "g_x0", "g_xd0", "g_R0", "g_omega0", "g_joint_angles")


class RolloutGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _GIN + _GOUT]


_lib = None


def load() -> C.CDLL:
    """Loads libmonoforce_b200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found. monoforce_b200 has no CPU fallback: build the CUDA library first with "
            f"`python -m monoforce_b200.build` (needs nvcc, targets sm_100a).")
    lib = C.CDLL(LIB_PATH)
    lib.mfb_rollout_forward.argtypes = [C.POINTER(RolloutDesc), C.POINTER(RolloutBuffers), C.c_int, C.c_void_p]
    lib.mfb_rollout_forward.restype = C.c_int
    lib.mfb_rollout_backward.argtypes = [C.POINTER(RolloutDesc), C.POINTER(RolloutBuffers), C.POINTER(RolloutGrads),
                                         C.c_int, C.c_void_p]
    lib.mfb_rollout_backward.restype = C.c_int
    lib.mfb_rollout_forward_host.argtypes = [C.POINTER(RolloutDesc), C.POINTER(RolloutBuffers), C.c_int, C.c_int]
    lib.mfb_rollout_forward_host.restype = C.c_int
    lib.mfb_rollout_workspace_bytes.argtypes = [C.POINTER(RolloutDesc), C.c_int]
    lib.mfb_rollout_workspace_bytes.restype = C.c_int64
    lib.mfb_lift_splat_forward.argtypes = [C.c_void_p] * 3 + [C.c_int] * 8 + [C.c_void_p]
    lib.mfb_lift_splat_forward.restype = C.c_int
    lib.mfb_lift_splat_backward.argtypes = [C.c_void_p] * 4 + [C.c_int] * 8 + [C.c_void_p]
    lib.mfb_lift_splat_backward.restype = C.c_int
    lib.mfb_conv_bn_act_bf16.argtypes = [C.c_void_p] * 5 + [C.c_int] * 7 + [C.c_void_p]
    lib.mfb_conv_bn_act_bf16.restype = C.c_int
    lib.mfb_conv2d_bf16.argtypes = [C.POINTER(ConvDesc)] + [C.c_void_p] * 9
    lib.mfb_conv2d_bf16.restype = C.c_int
    lib.mfb_lift_splat_forward_bf16.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p]
    lib.mfb_lift_splat_forward_bf16.restype = C.c_int
    lib.mfb_upsample_concat_nhwc_bf16.argtypes = [C.c_void_p] * 3 + [C.c_int] * 8 + [C.c_void_p]
    lib.mfb_upsample_concat_nhwc_bf16.restype = C.c_int
    lib.mfb_stem_conv_bf16.argtypes = [C.c_void_p] * 4 + [C.c_int] * 7 + [C.c_void_p]
    lib.mfb_stem_conv_bf16.restype = C.c_int
    lib.mfb_dwconv_bn_silu_bf16.argtypes = [C.c_void_p] * 5 + [C.c_int] * 10 + [C.c_void_p]
    lib.mfb_dwconv_bn_silu_bf16.restype = C.c_int
    lib.mfb_se_fold_bf16.argtypes = [C.c_void_p, C.c_float] + [C.c_void_p] * 6 + [C.c_int] * 5 + [C.c_void_p]
    lib.mfb_se_fold_bf16.restype = C.c_int
    lib.mfb_cast_f32_to_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
    lib.mfb_cast_f32_to_bf16.restype = C.c_int
    lib.mfb_terrain_postproc.argtypes = [C.c_void_p] * 3 + [C.c_longlong] + [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p]
    lib.mfb_terrain_postproc.restype = C.c_int
    lib.mfb_path_postproc.argtypes = [C.c_void_p] * 4 + [C.c_int] * 2 + [C.c_void_p]
    lib.mfb_path_postproc.restype = C.c_int
    lib.mfb_physics_loss.argtypes = ([C.c_void_p] * 4 + [C.c_int64] * 2 + [C.c_int] * 3 + [C.c_double, C.c_int] +
                                     [C.c_void_p] * 3 + [C.c_int, C.c_void_p])
    lib.mfb_physics_loss.restype = C.c_int
    lib.mfb_last_error.restype = C.c_char_p
    lib.mfb_abi_version.restype = C.c_int
    lib.mfb_kernel_launches.restype = C.c_longlong
    lib.mfb_release_scratch.restype = None
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mfb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def kernel_launches() -> int:
    return int(load().mfb_kernel_launches())

"""ctypes binding of libb200gs.so (include/b200gs.h).

The product path has no CPU fallback: if the CUDA library is missing or cannot be loaded this
module raises, loudly, at first use.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200GS_LIB_PATH") or os.path.join(_HERE, "lib", "libb200gs.so")  # env: A/B builds

EXPORTED_SYMBOLS = (
    "b200gs_forward", "b200gs_backward", "b200gs_mark_visible", "b200gs_buffer_sizes",
    "b200gs_last_error", "b200gs_version", "b200gs_launch_count",
    "b200gs_profile_enable", "b200gs_profile_read", "b200gs_stage_name", "b200gs_export_rgb8",
    "b200gs_set_option", "b200gs_ply_activate", "b200gs_transform_gaussians",
    "b200gs_extract_alpha", "b200gs_photometric_loss", "b200gs_photometric_loss_backward",
    "b200gs_ssim_forward", "b200gs_ssim_backward", "b200gs_adam_step", "b200gs_geom_layout",
    "b200gs_context_create", "b200gs_context_destroy", "b200gs_context_forward", "b200gs_context_ticket_wait",
    "b200gs_context_backward", "b200gs_context_query", "b200gs_policy_pair_capacity", "b200gs_policy_bin_shift",
    "b200gs_graph_instantiate", "b200gs_graph_launch", "b200gs_graph_exec_destroy",
)
ADAM_MAX_GROUPS = 8
DEFER_PAIR_CHECK = 1
FORWARD_ONLY = 2
OUT_RGB8 = 4
NUM_STAGES = 8


class B200GSParams(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("sh_degree", C.c_int32), ("M", C.c_int32),
        ("image_height", C.c_int32), ("image_width", C.c_int32),
        ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
        ("prefiltered", C.c_int32), ("debug", C.c_int32), ("near_plane", C.c_float), ("flags", C.c_int32),
        ("pair_capacity_hint", C.c_int64),
    ]


class B200GSPlyLayout(C.Structure):
    _fields_ = [("stride", C.c_int32), ("off_xyz", C.c_int32), ("off_fdc", C.c_int32), ("off_frest", C.c_int32),
                ("n_rest", C.c_int32), ("off_opacity", C.c_int32), ("off_scale", C.c_int32), ("off_rot", C.c_int32)]


class B200GSAdamGroup(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int64), ("lr", C.c_float), ("reserved", C.c_float)]


RESIZE_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)


class B200GSAlloc(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("resize", RESIZE_FN)]


_lib = None


class B200GSError(RuntimeError):
    pass


def lib():
    """Load libb200gs.so; raise if the CUDA extension has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200GSError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, fp = C.c_void_p, C.c_void_p
    L.b200gs_forward.restype = C.c_int
    L.b200gs_forward.argtypes = [C.POINTER(B200GSParams), fp, fp, fp, fp, fp, fp, fp, fp, fp, fp, fp,
                                 fp, vp, B200GSAlloc, B200GSAlloc, B200GSAlloc,
                                 vp, vp]      # num_rendered: int32* (pinned host memory in deferred mode)
    L.b200gs_backward.restype = C.c_int
    L.b200gs_backward.argtypes = [C.POINTER(B200GSParams), fp, fp, fp, fp, fp, fp, fp, fp, fp, fp, fp,
                                  vp, vp, vp, vp, C.c_int32, fp,
                                  fp, fp, fp, fp, fp, fp, fp, fp, B200GSAlloc, vp]
    L.b200gs_mark_visible.restype = C.c_int
    L.b200gs_mark_visible.argtypes = [C.c_int32, fp, fp, fp, vp, vp]
    L.b200gs_buffer_sizes.restype = C.c_int
    L.b200gs_buffer_sizes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                      C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.b200gs_last_error.restype = C.c_char_p
    L.b200gs_version.restype = C.c_int
    L.b200gs_launch_count.restype = C.c_int64
    L.b200gs_launch_count.argtypes = [C.c_int]
    L.b200gs_export_rgb8.restype = C.c_int
    L.b200gs_export_rgb8.argtypes = [fp, C.c_int32, C.c_int32, vp, vp]
    L.b200gs_ply_activate.restype = C.c_int
    L.b200gs_ply_activate.argtypes = [C.c_int32, fp, C.POINTER(B200GSPlyLayout), fp, fp, fp, fp, fp, vp]
    L.b200gs_transform_gaussians.restype = C.c_int
    L.b200gs_transform_gaussians.argtypes = [C.c_int32, fp, fp, vp, fp, fp, C.c_int32, fp, fp, vp]
    L.b200gs_photometric_loss.restype = C.c_int
    L.b200gs_photometric_loss.argtypes = [fp, fp, C.c_int64, C.c_float, C.c_float, fp, vp]
    L.b200gs_photometric_loss_backward.restype = C.c_int
    L.b200gs_photometric_loss_backward.argtypes = [fp, fp, C.c_int64, C.c_float, C.c_float, C.c_float, fp, fp, vp]
    L.b200gs_ssim_forward.restype = C.c_int
    L.b200gs_ssim_forward.argtypes = [fp, fp, C.c_int32, C.c_int32, C.c_int32, fp, fp, vp]
    L.b200gs_ssim_backward.restype = C.c_int
    L.b200gs_ssim_backward.argtypes = [fp, fp, fp, C.c_int32, C.c_int32, C.c_int32, C.c_float, fp, fp, vp]
    L.b200gs_adam_step.restype = C.c_int
    L.b200gs_adam_step.argtypes = [C.POINTER(B200GSAdamGroup), C.c_int32, C.c_float, C.c_float, C.c_float,
                                   C.c_int32, vp]
    L.b200gs_extract_alpha.restype = C.c_int
    L.b200gs_extract_alpha.argtypes = [vp, C.c_int32, C.c_int32, fp, vp]
    L.b200gs_geom_layout.restype = C.c_int
    L.b200gs_geom_layout.argtypes = [C.c_int32, C.POINTER(C.c_size_t)]
    L.b200gs_policy_pair_capacity.restype = C.c_int64
    L.b200gs_policy_pair_capacity.argtypes = [C.c_int64]
    L.b200gs_policy_bin_shift.restype = C.c_int32
    L.b200gs_policy_bin_shift.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_float]
    L.b200gs_graph_instantiate.restype = C.c_int
    L.b200gs_graph_instantiate.argtypes = [vp, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.b200gs_graph_launch.restype = C.c_int
    L.b200gs_graph_launch.argtypes = [vp, vp]
    L.b200gs_graph_exec_destroy.restype = C.c_int
    L.b200gs_graph_exec_destroy.argtypes = [vp]
    L.b200gs_context_create.restype = C.c_int
    L.b200gs_context_create.argtypes = [C.POINTER(C.c_void_p)]
    L.b200gs_context_destroy.restype = C.c_int
    L.b200gs_context_destroy.argtypes = [C.c_void_p]
    L.b200gs_context_forward.restype = C.c_int
    L.b200gs_context_forward.argtypes = [C.c_void_p, C.POINTER(B200GSParams), fp, fp, fp, fp, fp, fp, fp, fp, fp, fp, fp,
                                         fp, vp, B200GSAlloc, B200GSAlloc, B200GSAlloc, C.c_int32,
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int32), vp]
    L.b200gs_context_ticket_wait.restype = C.c_int
    L.b200gs_context_ticket_wait.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.b200gs_context_backward.restype = C.c_int
    L.b200gs_context_backward.argtypes = [C.c_void_p, C.POINTER(B200GSParams), C.c_int32, fp, fp, fp, fp, fp, fp, fp, fp,
                                          fp, fp, fp, vp, vp, vp, vp, C.c_int32, fp, fp, fp, fp, fp, fp, fp, fp, fp, vp]
    L.b200gs_context_query.restype = C.c_int
    L.b200gs_context_query.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int32)]
    L.b200gs_set_option.restype = C.c_int
    L.b200gs_set_option.argtypes = [C.c_char_p, C.c_int]
    L.b200gs_profile_enable.argtypes = [C.c_int]
    L.b200gs_profile_read.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int]
    L.b200gs_stage_name.restype = C.c_char_p
    L.b200gs_stage_name.argtypes = [C.c_int]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise B200GSError(f"libb200gs error {rc}: {lib().b200gs_last_error().decode()}")


def launch_count(reset: bool = False) -> int:
    return int(lib().b200gs_launch_count(1 if reset else 0))


def profile_enable(on: bool) -> None:
    lib().b200gs_profile_enable(1 if on else 0)


def profile_read(reset: bool = True) -> dict:
    """{stage name: (total ms, launches)} accumulated since the last reset."""
    L = lib()
    ms = (C.c_float * NUM_STAGES)()
    calls = (C.c_int32 * NUM_STAGES)()
    check(L.b200gs_profile_read(ms, calls, 1 if reset else 0))
    return {L.b200gs_stage_name(i).decode(): (float(ms[i]), int(calls[i])) for i in range(NUM_STAGES)}


def set_option(name: str, value: int) -> None:
    """Tuning knobs of the library ("bin_shift": -1 auto / 0..5, "gather": 0 TMA / 1 LDGSTS)."""
    check(lib().b200gs_set_option(name.encode(), int(value)))


def geom_layout(P: int) -> dict:
    """Byte offsets of the per-Gaussian arrays inside a forward call's geom buffer (b200gs_geom_layout)."""
    off = (C.c_size_t * 5)()
    check(lib().b200gs_geom_layout(C.c_int32(int(P)), off))
    return dict(zip(("rec", "depth_key", "tiles", "offsets", "clamped"), (int(o) for o in off)))

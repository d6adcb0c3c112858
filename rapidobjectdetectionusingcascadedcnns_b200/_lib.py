"""ctypes binding of librodc.so (the C ABI declared in include/rodc.h).

There is deliberately NO fallback: if the library is missing it is built in-tree with nvcc
(``build.py``); if that fails, or a compute entry point is called without a CUDA device, the
error propagates.  The oracle under ``oracle/`` is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
vp = C.c_void_p


class WindowSource(C.Structure):
    """``rodc_window_source`` (include/rodc.h)."""
    _fields_ = [("kind", C.c_int32), ("dense", vp), ("plan", vp), ("pyramid", vp),
                ("frames_base", vp), ("frame_off", vp), ("frame_hw", vp), ("meta", vp),
                ("gather_base", vp), ("win_off", vp), ("win_pitch", vp)]


class ScaledSlot(C.Structure):
    """``rodc_scaled_slot`` (include/rodc.h)."""
    _fields_ = [("scale", C.c_double), ("base_off", C.c_int64), ("frame_stride", C.c_int64),
                ("width", C.c_int32), ("height", C.c_int32), ("frame_base", C.c_int32), ("individually", C.c_int32)]


SRC_DENSE, SRC_PYRAMID, SRC_RECROP, SRC_GATHER = 0, 1, 2, 3
IMPL_TENSOR, IMPL_SIMT = 0, 1

# name -> (restype, argtypes); every symbol include/rodc.h declares
SIGNATURES = {
    "rodc_version": (C.c_int, []),
    "rodc_last_error": (C.c_char_p, []),
    "rodc_device_info": (C.c_int, [c_i32p, c_i32p, c_i32p]),
    "rodc_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(vp)]),
    "rodc_plan_destroy": (None, [vp]),
    "rodc_plan_n_levels": (C.c_int, [vp]),
    "rodc_plan_n_windows": (C.c_int64, [vp]),
    "rodc_plan_fits": (C.c_int, [vp]),
    "rodc_plan_levels": (C.c_int, [vp, c_f64p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i64p, c_i64p]),
    "rodc_plan_windows": (C.c_int, [vp, c_i32p, c_i32p, c_i32p, c_i32p]),
    "rodc_plan_pyramid_bytes": (C.c_int64, [vp]),
    "rodc_plan_scratch_bytes": (C.c_int64, [vp]),
    "rodc_pyramid": (C.c_int, [vp, vp, C.c_int, vp, vp, vp]),
    "rodc_extract_windows": (C.c_int, [vp, vp, C.c_int, vp, vp]),
    "rodc_window_meta": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
    "rodc_window_gather_rows": (C.c_int, [vp, C.c_int, C.c_int64, vp, vp, vp]),
    "rodc_window_levels": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
    "rodc_select_rows": (C.c_int, [vp, vp, C.c_int64, C.c_int, vp, vp]),
    "rodc_plan_create_scales": (C.c_int, [C.c_int, C.c_int, C.c_int, c_f64p, C.POINTER(vp)]),
    "rodc_window_scaled_rows": (C.c_int, [vp, vp, C.c_int64, vp, C.c_int, C.c_int, C.c_int64, vp, vp, vp, vp]),
    "rodc_source_windows": (C.c_int, [vp, C.c_int64, C.c_int64, C.c_int, vp, vp]),
    "rodc_recrop": (C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_int, vp, vp]),
    "rodc_stage_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
    "rodc_stage_create_general": (C.c_int, [C.c_int, C.c_int, c_i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp),
                                            C.c_int, C.c_int, vp, vp, vp, vp, C.POINTER(vp)]),
    "rodc_stage_destroy": (None, [vp]),
    "rodc_stage_workspace_bytes": (C.c_int64, [vp, C.c_int64, C.c_int]),
    "rodc_stage_forward": (C.c_int, [vp, C.POINTER(WindowSource), C.c_int64, C.c_int64, vp, C.c_int, vp, vp,
                                     vp, C.c_int64, C.c_int, vp]),
    "rodc_upload_frames": (C.c_int, [C.POINTER(vp), c_i64p, c_i64p, C.c_int, vp, vp]),
    "rodc_stage_forward_dyn": (C.c_int, [vp, C.POINTER(WindowSource), C.c_int64, C.c_int64, vp, vp, C.c_int, vp, vp,
                                         vp, C.c_int64, C.c_int, vp]),
    "rodc_compact_dyn": (C.c_int, [vp, C.c_int64, vp, C.c_double, vp, vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int,
                                   vp, vp, C.c_int, vp, vp, vp, vp]),
    "rodc_group_rectangles_dyn": (C.c_int, [vp, C.c_int64, vp, C.c_int, C.c_int, C.c_double, vp, vp, vp, vp]),
    "rodc_stage_features_f32": (C.c_int, [vp, C.POINTER(WindowSource), C.c_int64, C.c_int64, vp, vp]),
    "rodc_compact_scratch_bytes": (C.c_int64, [C.c_int64]),
    "rodc_compact": (C.c_int, [vp, C.c_int64, C.c_double, vp, vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int,
                               vp, vp, C.c_int, vp, vp, vp, vp]),
    "rodc_group_scratch_bytes": (C.c_int64, [C.c_int64, C.c_int]),
    "rodc_profile_enable": (C.c_int, [C.c_int]),
    "rodc_profile_collect": (C.c_int, [c_f64p, c_f64p, c_i64p, c_f64p, c_f64p, c_i64p]),
    "rodc_profile_table": (C.c_int, [C.c_int, c_i32p, c_i32p, c_f64p, c_f64p, c_f64p, c_i64p, c_i64p]),
    "rodc_launch_count": (C.c_int64, []),
    "rodc_kernel_class_name": (C.c_char_p, [C.c_int]),
    "rodc_group_rectangles": (C.c_int, [vp, C.c_int64, C.c_int, C.c_int, C.c_double, vp, vp, vp, vp]),
}


class RodcError(RuntimeError):
    pass


_PATH_OVERRIDE = None


def use_library(path: str) -> None:
    """Point the binding at another build of the same ABI BEFORE the first call (``profiles/`` scripts load
    ``librodc_probes.so`` this way).  A code-level switch on purpose: no environment variable can swap the
    library under a user of the package."""
    global _PATH_OVERRIDE
    if _LIB is not None:
        raise RodcError("use_library() must be called before the library is loaded")
    _PATH_OVERRIDE = path


def lib() -> C.CDLL:
    """Load librodc.so and declare every signature.  The library is (re)built first whenever its content stamp
    does not match the sources in ``csrc/`` -- a stale ``.so`` is never used silently."""
    global _LIB
    if _LIB is None:
        path = _PATH_OVERRIDE or _build.LIB
        if _PATH_OVERRIDE is None and (not _build.is_current() or os.environ.get("RODC_REBUILD")):
            _build.build_library(force=bool(os.environ.get("RODC_REBUILD")))
        handle = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().rodc_last_error()
        raise RodcError("%s failed (%d): %s" % (what or "librodc call", rc, (msg or b"").decode()))

"""ctypes binding of ``libpano360_b200.so`` (the C ABI in include/pano360_b200.h).

There is no CPU fallback: if the library is missing the import of any compute
entry point raises, and every call checks the returned status and raises
``RuntimeError`` with ``p360_last_error`` text.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

import torch  # noqa: F401  (loads libcudart first so the library shares torch's runtime)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpano360_b200.so")

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes; every function returns int.  Kept in one table so the
# symbol-export test can walk it next to the header.
SIGNATURES = {
    "p360_version": [],
    "p360_last_error": [C.c_char_p, _i],
    "p360_device_info": [_i, C.POINTER(C.c_int32)],
    "p360_pack_rgbx": [_vp, _i, _i, _vp, _vp],
    "p360_pack_rgbx_rect": [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "p360_pack_rgbx_batch": [_vp, _i, _i, _i, _vp],
    "p360_copy_rect": [_vp, _i64, _vp, _i64, _i64, _i64, _vp],
    "p360_source_rects": [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "p360_warp_batch": [_vp, _vp, _i, _vp, _vp, _i, _vp],
    "p360_seam_plan_build": [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp],
    "p360_warp_tiles": [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "p360_owner_update": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp],
    "p360_owner_decode": [_vp, _vp, _i64, _vp],
    "p360_gauss_blur": [_vp, _vp, _vp, _i, _i, C.POINTER(C.c_float), _i, _vp],
    "p360_blur_set_taps": [_i, C.POINTER(C.c_float), _i, _vp],
    "p360_gauss_blur_batch": [_vp, _i, _i, _i, _vp, _vp],
    "p360_owned_boxes": [_vp, _vp, _i, _i, _i, _vp],
    "p360_tile_maps_build": [_vp, _vp, _vp, _i, _i, _i, _vp, _vp],
    "p360_pyramid_dims": [_i, _i, _i, C.POINTER(C.c_int32)],
    "p360_pyramid_reduce_batch": [_vp, _i, _i, _i, _vp, _i, _vp, _vp],
    "p360_multiband_collapse": [_vp, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "p360_linear_collapse": [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp],
    "p360_paste_collapse": [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp],
    "p360_pair_stats_blocks": [_i, _i],
    "p360_pair_overlap_stats": [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "p360_cover_update": [_vp, _i, _i, _i, _i, _vp, _i, _vp],
    "p360_owner_to_alpha": [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp],
    "p360_band_accumulate": [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp],
    "p360_exact_collapse": [_vp, _i, _vp, _vp, _i, _i, _vp],
    "p360_crop_scratch_bytes": [_i, _i],
    "p360_crop_rect": [_vp, _i, _i, _vp, _vp, _vp],
    "p360_resize_u8": [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp],
}
MAX_LEVELS = 8

# NumPy mirrors of the job records in include/pano360_b200.h (filled on the
# host, shipped to the device as raw bytes)
WARP_JOB = np.dtype([("src", "u8"), ("lut", "u8"), ("hat_y", "u8"), ("hat_x", "u8"), ("ray_x", "u8"),
                     ("ray_z", "u8"), ("ray_y", "u8"), ("out", "u8"), ("invalid", "u8"), ("kr", "f8", (9,)),
                     ("h", "i4"), ("w", "i4"), ("c", "i4"), ("pw", "i4"), ("ph", "i4"), ("x0", "i4"),
                     ("y0", "i4"), ("col0", "i4"), ("row0", "i4"), ("patch", "i4"), ("half_w", "f4"),
                     ("half_h", "f4"), ("max_x", "f4"), ("max_y", "f4"), ("inv_2w", "f4"), ("inv_2h", "f4"),
                     ("ty0", "i4"), ("ty1", "i4"), ("tx0", "i4"), ("tx1", "i4")])
BLUR_JOB = np.dtype([("in", "u8"), ("out", "u8"), ("tmp", "u8"), ("w", "i4"), ("h", "i4"), ("slot", "i4"),
                     ("shift", "i4"), ("patch", "u8"), ("pad", "i4"), ("grow", "i4")])
BAND_PATCH = np.dtype([("rgba", "u8"), ("invalid", "u8"), ("d2", "u8"), ("d4", "u8"),
                       ("low", "u8", (MAX_LEVELS - 1,)), ("x0", "i4"), ("y0", "i4"), ("pw", "i4"),
                       ("ph", "i4"), ("w4", "i4"), ("h4", "i4"), ("pad", "i4"), ("index", "i4"),
                       ("own", "i4", (4,))])
TILE_MAPS = np.dtype([("present", "u8"), ("cand", "u8"), ("need", "u8"), ("multi", "u8"), ("work", "u8"),
                      ("work_count", "u8"), ("wneed", "u8"), ("tiles_x", "i4"), ("tiles_y", "i4"), ("words", "i4"), ("row0", "i4"),
                      ("reach_x", "i4"), ("reach_y", "i4"), ("work_cap", "i4"), ("h_rows", "i4")])
PAIR_JOB = np.dtype([("src_i", "u8"), ("src_j", "u8"), ("inv", "f8", (9,))])
PACK_JOB = np.dtype([("src", "u8"), ("dst", "u8"), ("h", "i4"), ("w", "i4"), ("r0", "i4"), ("r1", "i4"), ("c0", "i4"), ("c1", "i4")])
assert PACK_JOB.itemsize == 40
assert PAIR_JOB.itemsize == 88
assert WARP_JOB.itemsize == 224 and BLUR_JOB.itemsize == 56 and BAND_PATCH.itemsize == 136
assert TILE_MAPS.itemsize == 88
OWN_OFFSET = BAND_PATCH.fields["own"][1]

# entry points whose int return is a value, not a status
_VALUE_RETURN = {"p360_version", "p360_pair_stats_blocks", "p360_crop_scratch_bytes"}

_lib = None
launch_count = 0      # kernels launched through this binding (bench.py: gpu_launches)
_LAUNCHES = {"p360_pack_rgbx": 1, "p360_pack_rgbx_rect": 1, "p360_pack_rgbx_batch": 1, "p360_source_rects": 1, "p360_warp_batch": 1, "p360_owner_update": 1, "p360_owner_decode": 1,
             "p360_gauss_blur": 2, "p360_gauss_blur_batch": 2, "p360_pyramid_reduce_batch": 1,
             "p360_owned_boxes": 1,            # (+1 scan kernel per pass with maps, counted by the caller)
             "p360_tile_maps_build": 3, "p360_seam_plan_build": 3, "p360_warp_tiles": 1,
             "p360_multiband_collapse": 1, "p360_linear_collapse": 1, "p360_paste_collapse": 1,
             "p360_pair_overlap_stats": 2, "p360_cover_update": 1, "p360_crop_rect": 3,
             "p360_owner_to_alpha": 1, "p360_band_accumulate": 1, "p360_exact_collapse": 1,
             "p360_resize_u8": 1}


def load():
    """dlopen the library (once) and declare prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if _build.stale():                   # missing, or older than csrc/ or the header: never run stale kernels
        try:
            _build.build()
        except Exception as exc:
            raise RuntimeError(
                f"{LIB_PATH} is missing or out of date and could not be rebuilt ({exc}); build it with "
                "`python -m pano360_b200.build` (pano360_b200 has no CPU fallback)") from exc
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _i64 if name == "p360_crop_scratch_bytes" else _i
    _lib = lib
    return lib


def last_error():
    buf = C.create_string_buffer(512)
    load().p360_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def call(name, *args):
    """Invoke an entry point; raise on a non-zero status."""
    global launch_count
    rc = getattr(load(), name)(*args)
    if name in _VALUE_RETURN:
        return rc
    if rc != 0:
        raise RuntimeError(f"{name} failed with status {rc}: {last_error()}")
    launch_count += _LAUNCHES.get(name, 0)
    return 0


def ptr(t):
    """Device (or host) address of a torch tensor / None."""
    return None if t is None else t.data_ptr()

"""ctypes binding of libdsc.so -- exactly the symbols declared in include/dsc.h.

This is the stub a maintainer of the reference adds to call the engine.
There is no CPU fallback: if the shared library cannot be loaded the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DSC_LIB") or os.path.join(_HERE, "libdsc.so")     # DSC_LIB: A/B builds only

DSC_OK, DSC_ERR_INVALID, DSC_ERR_CUDA, DSC_ERR_NOMEM, DSC_ERR_CAPACITY = 0, -1, -2, -3, -4
MEM_HOST, MEM_DEVICE = 0, 1
VARIANT_A, VARIANT_B = 0, 1
FLAG_GAP_F32, FLAG_EXACT_ONLY, FLAG_GAP_RING_F32, FLAG_GAP_SECTOR_F32 = 1, 2, 4, 8

EXPORTS = (
    "dsc_version", "dsc_create", "dsc_destroy", "dsc_last_error", "dsc_set_timing", "dsc_last_kernel_ms",
    "dsc_kernel_launches", "dsc_set_option", "dsc_host_alloc", "dsc_host_free", "dsc_sc_build_batch",
    "dsc_sc_build_batch_f64", "dsc_sc_normalize01",
    "dsc_set_polar_tables", "dsc_raster_boxes", "dsc_reconstruct_batch", "dsc_covered_mask",
    "dsc_gather_cells", "dsc_label_maps", "dsc_label_heatmap", "dsc_voxel_downsample",
)


class DscParams(C.Structure):
    _fields_ = [
        ("num_ring", C.c_int32),
        ("num_sector", C.c_int32),
        ("max_length", C.c_double),
        ("variant", C.c_int32),
        ("full_height", C.c_int32),
        ("height_offset", C.c_float),
        ("flags", C.c_uint32),
    ]


class DscScanStats(C.Structure):
    _fields_ = [
        ("n_points", C.c_int64),
        ("n_nonfinite", C.c_int64),
        ("n_sector_overflow", C.c_int64),
        ("n_clamped_ring", C.c_int64),
        ("n_out_of_range", C.c_int64),
        ("n_slow_path", C.c_int64),
        ("n_boundary", C.c_int64),
    ]


STATS_FIELDS = tuple(n for n, _ in DscScanStats._fields_)


class DscError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dsc error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load libdsc.so (building it with nvcc first if the in-tree .so is missing or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()          # raises if nvcc is unavailable -- no fallback
    L = C.CDLL(LIB_PATH)
    if os.environ.get("DSC_LIB"):
        # A/B builds (profiles/microbench/build_variant.sh, older revisions) may lack newer entry points: bind what exists
        class _Missing:
            argtypes = restype = None

            def __call__(self, *a):
                raise DscError(DSC_ERR_INVALID, "this DSC_LIB build does not export the entry point")
        for name in EXPORTS:
            if not hasattr(L, name):
                setattr(L, name, _Missing())
    vp, ip = C.c_void_p, C.c_int
    L.dsc_version.restype = C.c_char_p
    L.dsc_create.argtypes = [ip, C.POINTER(vp)]
    L.dsc_destroy.argtypes = [vp]
    L.dsc_last_error.argtypes = [vp]
    L.dsc_last_error.restype = C.c_char_p
    L.dsc_set_timing.argtypes = [vp, ip]
    L.dsc_last_kernel_ms.argtypes = [vp]
    L.dsc_last_kernel_ms.restype = C.c_float
    L.dsc_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.dsc_kernel_launches.argtypes = [vp]
    L.dsc_kernel_launches.restype = C.c_int64
    L.dsc_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.dsc_host_free.argtypes = [vp, vp]
    L.dsc_sc_build_batch.argtypes = [vp, C.POINTER(DscParams), vp, vp, ip, ip, ip, vp, vp, vp, vp]
    L.dsc_sc_build_batch_f64.argtypes = [vp, C.POINTER(DscParams), C.c_double, vp, vp, ip, ip, ip, vp, vp, vp]
    L.dsc_sc_normalize01.argtypes = [vp, vp, ip, C.c_int64, ip, vp, vp]
    L.dsc_set_polar_tables.argtypes = [vp, ip, ip, vp, vp, vp]
    L.dsc_raster_boxes.argtypes = [vp, vp, ip, ip, vp, vp, C.c_int64, vp, vp]
    L.dsc_reconstruct_batch.argtypes = [vp, vp, ip, ip, C.c_double, ip, vp, vp, vp, C.c_int64, vp]
    L.dsc_covered_mask.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, ip, vp, vp]
    L.dsc_gather_cells.argtypes = [vp, vp, vp, C.c_int64, ip, vp, vp]
    L.dsc_label_maps.argtypes = [vp, vp, vp, vp, vp, ip, ip, ip, ip, ip, ip, vp, vp, vp]
    L.dsc_label_heatmap.argtypes = [vp, vp, vp, ip, ip, ip, ip, vp, ip, ip, vp, vp]
    L.dsc_voxel_downsample.argtypes = [vp, vp, C.c_int64, ip, ip, C.c_double, ip, vp, C.c_int64, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int:      # default restype
            fn.restype = C.c_int
    _lib = L
    return L

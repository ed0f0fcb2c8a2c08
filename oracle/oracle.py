"""ctypes front-end of the CPU oracle (``oracle/dsc_oracle.c``).

TEST INFRASTRUCTURE ONLY -- importable from ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product package
``distropointclouds_b200`` must never import this module.

The oracle restates, in scalar C with individually rounded operations, the arithmetic the
reference performs through NumPy:

* variant A  ``ScanContext.ptcloud2sc_vectorized``   scaner_version2.py:64-115
* variant B  ``ScanContextComputer._scan_context``   scan_context_node.py:102-157
* ``cells_covered_by_box``                            kitti_dataset.py:204-253
* ``pedestrian_to_scan_context_coords``               kitti_dataset.py:129-159
* ``scan_context_to_pointcloud``                      reconstruct_pointcloud.py:47-77
* ``to_tensor01``                                     tiny_sc_ped.py:54-60
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdsc_oracle.so")

VARIANT_A, VARIANT_B = 0, 1


class OrcParams(C.Structure):
    _fields_ = [
        ("num_ring", C.c_int32),
        ("num_sector", C.c_int32),
        ("max_length", C.c_double),
        ("variant", C.c_int32),
        ("full_height", C.c_int32),
        ("height_offset", C.c_float),
        ("gap_f32", C.c_int32),
    ]


class OrcStats(C.Structure):
    _fields_ = [
        ("n_points", C.c_int64),
        ("n_nonfinite", C.c_int64),
        ("n_sector_overflow", C.c_int64),
        ("n_clamped_ring", C.c_int64),
        ("n_out_of_range", C.c_int64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (``make -C oracle``)."""
    srcs = [os.path.join(_HERE, f) for f in ("dsc_oracle.c", "dsc_oracle_f64.c", "dsc_oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp, dp, ip, lp = (C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32),
                          C.POINTER(C.c_int64))
        L.orc_atan2f.argtypes = [C.c_float, C.c_float]
        L.orc_atan2f.restype = C.c_float
        L.orc_atanf.argtypes = [C.c_float]
        L.orc_atanf.restype = C.c_float
        L.orc_atan2f_array.argtypes = [fp, fp, fp, C.c_int64]
        L.orc_atan2f_array.restype = None
        L.orc_sc_build.argtypes = [C.POINTER(OrcParams), fp, C.c_int64, C.c_int, fp, fp,
                                   C.POINTER(OrcStats)]
        L.orc_sc_build.restype = C.c_int
        L.orc_sc_build_batch.argtypes = [C.POINTER(OrcParams), fp, lp, C.c_int, C.c_int, fp,
                                         C.POINTER(OrcStats), C.c_int]
        L.orc_sc_build_batch.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        L.orc_atan2.argtypes = [C.c_double, C.c_double]
        L.orc_atan2.restype = C.c_double
        L.orc_atan2_array.argtypes = [dp, dp, dp, C.c_int64]
        L.orc_atan2_array.restype = None
        L.orc_sc_build_f64.argtypes = [C.POINTER(OrcParams), C.c_double, dp, C.c_int64, C.c_int, dp, dp, fp,
                                       C.POINTER(OrcStats), lp]
        L.orc_sc_build_f64.restype = C.c_int
        L.orc_voxel_downsample.argtypes = [dp, C.c_int64, C.c_int, C.c_double, dp, C.c_int64, lp]
        L.orc_voxel_downsample.restype = C.c_int64
        L.orc_normalize01.argtypes = [fp, C.c_int64, fp, fp, fp]
        L.orc_normalize01.restype = None
        L.orc_raster_box.argtypes = [C.c_int, C.c_int, dp, dp, dp] + [C.c_double] * 6 + [ip, C.c_int64]
        L.orc_raster_box.restype = C.c_int64
        L.orc_reconstruct.argtypes = [C.c_int, C.c_int, dp, dp, dp, fp, C.c_int, C.c_double, dp, ip,
                                      C.c_int64]
        L.orc_reconstruct.restype = C.c_int64
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


@dataclass
class Params:
    num_ring: int = 160
    num_sector: int = 720
    max_length: float = 20.0
    variant: int = VARIANT_A
    full_height: bool = True
    height_offset: float = 2.0
    gap_f32: bool = False

    def c(self) -> OrcParams:
        return OrcParams(int(self.num_ring), int(self.num_sector), float(self.max_length),
                         int(self.variant), int(bool(self.full_height)), float(self.height_offset),
                         3 if self.gap_f32 is True else int(self.gap_f32))       # True: both gaps; 1: ring only; 2: sector only


def atan2f(y, x) -> np.ndarray:
    y = np.ascontiguousarray(y, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(y)
    lib().orc_atan2f_array(_p(y, C.c_float), _p(x, C.c_float), _p(out, C.c_float), y.size)
    return out


def atan2(y, x) -> np.ndarray:
    """float64 arctan2, fdlibm flavour (dsc_oracle_f64.c)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(y)
    lib().orc_atan2_array(_p(y, C.c_double), _p(x, C.c_double), _p(out, C.c_double), y.size)
    return out


def sc_build_f64(points: np.ndarray, params: Params = Params(), theta_override=None):
    """One float64 scan (what Open3D's voxel_down_sample hands the reference) -> (grid, stats dict).
    Variant A returns a float64 grid (scaner_version2.py:97-108), variant B a float32 grid
    (scan_context_node.py:150-151).  stats has the extra key ``n_boundary``."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    if pts.ndim != 2 or pts.shape[1] < 3:
        raise ValueError("points must be (N,>=3)")
    n = pts.shape[0]
    is_a = params.variant == VARIANT_A
    out = np.empty((params.num_ring, params.num_sector), dtype=np.float64 if is_a else np.float32)
    st = OrcStats()
    nb = C.c_int64(0)
    th = None
    if theta_override is not None:
        th = np.ascontiguousarray(theta_override, dtype=np.float64)
        assert th.size == n
    cp = params.c()
    rc = lib().orc_sc_build_f64(C.byref(cp), float(params.height_offset), _p(pts, C.c_double), n, pts.shape[1],
                                _p(th, C.c_double) if th is not None else None,
                                _p(out, C.c_double) if is_a else None, _p(out, C.c_float) if not is_a else None,
                                C.byref(st), C.byref(nb))
    if rc:
        raise ValueError(f"orc_sc_build_f64 rc={rc}")
    d = st.as_dict()
    d["n_boundary"] = int(nb.value)
    return out, d


def voxel_downsample(points: np.ndarray, voxel_size: float):
    """Open3D's published voxel_down_sample algorithm (parity unpinned: Open3D is not installed), output sorted by
    voxel index.  float32 input is widened exactly, as Vector3dVector does.  Returns ((M,3) float64, n_dropped)."""
    pts = np.ascontiguousarray(np.asarray(points)[:, :3], dtype=np.float64)
    out = np.empty((max(len(pts), 1), 3), np.float64)
    nd = C.c_int64(0)
    k = lib().orc_voxel_downsample(_p(pts, C.c_double), len(pts), 3, float(voxel_size), _p(out, C.c_double), len(out),
                                   C.byref(nd))
    if k < 0:
        raise ValueError(f"orc_voxel_downsample rc={k}")
    return out[:k].copy(), int(nd.value)


def as_packed(points: np.ndarray):
    """Return (contiguous float32 buffer, n, stride_in_floats) for an (N,3|4) array or view."""
    pts = np.asarray(points)
    if pts.dtype != np.float32:
        raise TypeError("oracle takes float32 points")
    n = pts.shape[0]
    if pts.ndim != 2 or pts.shape[1] not in (3, 4):
        raise ValueError("points must be (N,3) or (N,4)")
    if n and pts.strides == (16, 4) and pts.shape[1] == 3 and pts.base is not None:
        # the reference's scan[:, :-1] view -- re-materialise the packed xyzi rows
        pts = np.ascontiguousarray(pts)
        return pts, n, 3
    pts = np.ascontiguousarray(pts)
    return pts, n, pts.shape[1]


def sc_build(points: np.ndarray, params: Params = Params(), theta_override=None):
    """One scan -> (grid float32 (R,S), stats dict)."""
    pts, n, stride = as_packed(points)
    out = np.empty((params.num_ring, params.num_sector), dtype=np.float32)
    st = OrcStats()
    th = None
    if theta_override is not None:
        th = np.ascontiguousarray(theta_override, dtype=np.float32)
        assert th.size == n
    cp = params.c()
    rc = lib().orc_sc_build(C.byref(cp), _p(pts, C.c_float), n, stride,
                            _p(th, C.c_float) if th is not None else None, _p(out, C.c_float),
                            C.byref(st))
    if rc:
        raise ValueError(f"orc_sc_build rc={rc}")
    return out, st.as_dict()


def sc_build_batch(points: np.ndarray, offsets: np.ndarray, params: Params = Params(),
                   n_threads: int = 0):
    """Packed (total,4|3) float32 points + int64 offsets[n+1] -> (n,R,S) float32 grids."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = offsets.size - 1
    out = np.empty((n, params.num_ring, params.num_sector), dtype=np.float32)
    st = (OrcStats * n)()
    cp = params.c()
    rc = lib().orc_sc_build_batch(C.byref(cp), _p(pts, C.c_float), _p(offsets, C.c_int64), n,
                                  pts.shape[1], _p(out, C.c_float), st, n_threads)
    if rc:
        raise ValueError(f"orc_sc_build_batch rc={rc}")
    return out, [s.as_dict() for s in st]


def max_threads() -> int:
    return int(lib().orc_max_threads())


def normalize01(grid: np.ndarray) -> np.ndarray:
    g = np.ascontiguousarray(grid, dtype=np.float32)
    out = np.empty_like(g)
    lib().orc_normalize01(_p(g, C.c_float), g.size, _p(out, C.c_float), None, None)
    return out


def polar_tables(num_ring=160, num_sector=720, max_length=20.0):
    """Cell-centre tables, built with NumPy exactly as the reference builds them
    (kitti_dataset.py:219-229, reconstruct_pointcloud.py:52-59)."""
    gap_ring = max_length / num_ring
    gap_sector = 360.0 / num_sector
    r_centers = (np.arange(num_ring) + 0.5) * gap_ring
    theta = np.deg2rad((np.arange(num_sector) + 0.5) * gap_sector)
    return (np.ascontiguousarray(r_centers, dtype=np.float64),
            np.ascontiguousarray(np.cos(theta), dtype=np.float64),
            np.ascontiguousarray(np.sin(theta), dtype=np.float64))


def cells_covered_by_box(x, y, l, w, yaw_lidar, num_ring=160, num_sector=720, max_length=20.0):
    rc, ct, st = polar_tables(num_ring, num_sector, max_length)
    cap = num_ring * num_sector
    cells = np.empty((cap, 2), dtype=np.int32)
    k = lib().orc_raster_box(num_ring, num_sector, _p(rc, C.c_double), _p(ct, C.c_double),
                             _p(st, C.c_double), float(x), float(y), float(l), float(w),
                             float(np.cos(yaw_lidar)), float(np.sin(yaw_lidar)),
                             _p(cells, C.c_int32), cap)
    return cells[:k].astype(np.int64)


def pedestrian_to_scan_context_coords(px, py, num_ring=160, num_sector=720, max_length=20,
                                      atan2=None):
    """kitti_dataset.py:129-159 -- scalar, in the dtype of the inputs (np.float32 or float).
    ``atan2`` lets the caller choose the arctan2 flavour for float32 inputs."""
    gap_ring = max_length / num_ring
    gap_sector = 360 / num_sector
    if isinstance(px, np.float32):
        a = np.float32(atan2f([py], [px])[0]) if atan2 is None else np.float32(atan2(py, px))
        theta = np.rad2deg(a) % 360
    else:
        theta = np.rad2deg(np.arctan2(py, px)) % 360
    r = np.sqrt(px ** 2 + py ** 2)
    return int(np.minimum(r / gap_ring, num_ring - 1)), int(theta / gap_sector)


def reconstruct(grid: np.ndarray, num_ring=160, num_sector=720, max_length=20.0, full_height=True,
                height_offset=2.0):
    g = np.ascontiguousarray(grid, dtype=np.float32)
    rc, ct, st = polar_tables(num_ring, num_sector, max_length)
    cap = num_ring * num_sector
    pts = np.empty((cap, 3), dtype=np.float64)
    ids = np.empty((cap,), dtype=np.int32)
    k = lib().orc_reconstruct(num_ring, num_sector, _p(rc, C.c_double), _p(ct, C.c_double),
                              _p(st, C.c_double), _p(g, C.c_float), int(bool(full_height)),
                              float(height_offset), _p(pts, C.c_double), _p(ids, C.c_int32), cap)
    return pts[:k].copy(), ids[:k].copy()


# ---- label-side training tensors: plain NumPy/Python restatement of the reference loops --------

def label_weight_mask(ped_list, H=160, W=720):
    """ScanContextPedestrianDataset.__getitem__ (distro_model/tiny_sc_pred_v2.py:57-85): (mask, weight_map) float32."""
    mask = np.zeros((H, W), dtype=np.float32)
    weight_map = np.ones((H, W), dtype=np.float32)
    for ped in ped_list:
        covered = ped.get('covered_cells', [])
        distance = ped.get('distance', 10.0)
        num_cells = len(covered)
        if num_cells > 0:
            size_weight = min(100.0 / num_cells, 5.0)                        # :67
            distance_weight = 1.0 + (distance / 20.0) ** 2                   # :70
            ped_weight = min(size_weight * distance_weight, 10.0)            # :73
        else:
            ped_weight = 1.0
        for r, s in covered:
            if 0 <= r < H and 0 <= s < W:                                    # :81
                mask[r, s] = 1.0
                weight_map[r, s] = max(weight_map[r, s], ped_weight)         # float32 store of the float64 maximum
    return mask, weight_map


def label_clip_wrap_mask(ped_list, H=160, W=720):
    """ScanContextPedDataset.__getitem__ (distro_model/train_sc_ped.py:175-183): ring clipped, sector np.mod-wrapped."""
    mask = np.zeros((H, W), dtype=np.float32)
    for ped in ped_list:
        cells = np.array(ped.get("covered_cells", []), dtype=np.int32)
        if cells.size:
            rr = np.clip(cells[:, 0], 0, H - 1)
            ss = np.mod(cells[:, 1], W)
            mask[rr, ss] = 1.0
    return mask


def label_center_heatmap(ped_list, H=160, W=720, sigma=1.8):
    """ScanContextPedDataset.__getitem__ (train_sc_ped.py:165-173) with gaussian2d (:74-79)."""
    heat = np.zeros((H, W), dtype=np.float32)
    rr = np.arange(H)[:, None]
    ss = np.arange(W)[None, :]
    for ped in ped_list:
        r = int(ped.get("center_ring_idx", -1))
        s = int(ped.get("center_sector_idx", -1))
        if 0 <= r < H and 0 <= s < W:
            d2 = (rr - r) ** 2 + (ss - s) ** 2
            with np.errstate(under="ignore"):
                heat = np.maximum(heat, np.exp(-0.5 * d2 / (sigma ** 2)))
    return heat

"""Patch the CUDA engine into the *imported reference modules* (armando-genis/DistroPointclouds).

    import scaner_version2, reconstruct_pointcloud          # the reference, unmodified
    import distropointclouds_b200.compat as dsc_compat
    dsc_compat.install()                                    # hot-path methods now run on the GPU

``install()`` finds the reference classes by name in ``sys.modules`` (or in the modules passed
explicitly) and replaces exactly the hot-path methods of SURVEY 8(a); everything else of the
reference (file walking, JSON, plotting, ROS) keeps running as it is.  Class statics are read from
the patched class at call time.  ``uninstall()`` restores the originals.
"""
from __future__ import annotations

import sys
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

from . import _ffi
from .engine import ScanContextParams, default_engine, variant_a_gap_dtypes
from .scan_context import _raise_like_reference, _is_np_int

_STATS = _ffi.STATS_FIELDS
_saved: List[Tuple[object, str, object]] = []
device = 0          # GPU the patched methods run on
strict = True       # raise IndexError where variant A of the reference would


# ---- replacements -----------------------------------------------------------------------------

def _ptcloud2sc_vectorized(self, ptcloud, num_sector, num_ring, max_length):
    """scaner_version2.py:64-115 / kitti_dataset.py:47-98 on K1 (variant A)."""
    cls = type(self)
    params = ScanContextParams(
        num_ring=int(num_ring), num_sector=int(num_sector), max_length=float(max_length), variant=_ffi.VARIANT_A,
        full_height=bool(cls.use_full_height_range), height_offset=float(cls.custom_height_offset),
        gap_f32=variant_a_gap_dtypes(num_sector, num_ring, max_length))
    grid, stats = default_engine(device).build_any(ptcloud, params)       # float32 (.bin) or float64 (down-sampled)
    cls.last_stats = dict(zip(_STATS, (int(v) for v in stats)))
    if strict:
        _raise_like_reference(stats)
    return grid


def _sc_to_pointcloud_loop(self, sc_data):
    """scaner_version2.py:216-262 on K3 (returns points, colors)."""
    cls = type(self)
    pts, _ = default_engine(device).reconstruct_host(
        np.asarray(sc_data).astype(np.float32), int(cls.ring_res[0]), int(cls.sector_res[0]), float(cls.max_length),
        bool(cls.use_full_height_range), float(cls.custom_height_offset))[0]
    if pts.shape[0] == 0:
        return np.array([]), np.array([])
    return pts, np.full((pts.shape[0], 3), 0.7, dtype=float)


def _make_cells_covered_by_box(ref_cls):
    def cells_covered_by_box(x, y, l, w, yaw_lidar, num_ring=None, num_sector=None, max_length=None):
        """kitti_dataset.py:204-253 on K2."""
        if num_ring is None: num_ring = int(ref_cls.ring_res[0])
        if num_sector is None: num_sector = int(ref_cls.sector_res[0])
        if max_length is None: max_length = float(ref_cls.max_length)
        cells, _ = default_engine(device).raster_boxes_host(
            np.array([[x, y, l, w, yaw_lidar]], dtype=np.float64), int(num_ring), int(num_sector), float(max_length))
        return cells[0]
    return staticmethod(cells_covered_by_box)


def _scan_context_b(self, pts_xyz):
    """scan_context_node.py:102-157 on K1 (variant B)."""
    pts = np.asarray(pts_xyz)
    if pts.size == 0:
        return np.zeros((self.ring_res, self.sector_res), dtype=np.float32)
    params = ScanContextParams(int(self.ring_res), int(self.sector_res), float(self.max_length), _ffi.VARIANT_B,
                               bool(self.use_full_height_range), float(self.custom_height_offset))
    # the live node hands over float64 (Open3D-downsampled, scan_context_node.py:100,161); ROS float32 also works
    grid, stats = default_engine(device).build_any(pts, params)
    self.last_stats = dict(zip(_STATS, (int(v) for v in stats)))
    return grid


def _downsample_voxel(pts_xyz, voxel):
    """scan_context_node.py:93-100 on K7 (Open3D's published algorithm, parity unpinned; output sorted by voxel)."""
    pts = np.asarray(pts_xyz)
    if pts.size == 0:
        return pts_xyz
    return default_engine(device).voxel_downsample_host(pts, float(voxel))[0]


def _rec_sc_to_pointcloud(self, sc_data):
    """reconstruct_pointcloud.py:47-77 on K3 (returns points, colors, bin_ids)."""
    sc = np.asarray(sc_data)
    if sc.dtype != np.float32:
        sc32 = sc.astype(np.float32)
        if not np.array_equal(sc32.astype(sc.dtype), sc):
            raise TypeError("grid values must be float32-representable (the sc_*.npy dtype)")
        sc = sc32
    pts, ids = default_engine(device).reconstruct_host(
        sc, int(self.ring_res), int(self.sector_res), float(self.max_length),
        bool(self.use_full_height_range), float(self.custom_height_offset))[0]
    return pts, np.full((pts.shape[0], 3), 0.7, dtype=float), ids


def _rec_add_covered_cells(self, sc_data, covered_cells):
    """reconstruct_pointcloud.py:104-130 on K5."""
    if not covered_cells:
        return np.empty((0, 3)), np.empty((0, 3))
    pts = default_engine(device).gather_cells_host(
        np.ascontiguousarray(sc_data, dtype=np.float32), np.array(covered_cells, dtype=int), int(self.ring_res),
        int(self.sector_res), float(self.max_length))
    if not self.use_full_height_range:
        pts = pts[pts[:, 2] > 0]
        pts[:, 2] = (pts[:, 2].astype(np.float32) - np.float32(self.custom_height_offset)).astype(np.float64)
    return pts, np.tile(np.array([[1.0, 0.0, 0.0]]), (pts.shape[0], 1))


def _to_tensor01(x):
    """tiny_sc_ped.py:54-60 on K4."""
    import torch
    g = np.ascontiguousarray(np.asarray(x).astype(np.float32))
    return torch.from_numpy(default_engine(device).normalize01_host(g)).unsqueeze(0)


# ---- patching ---------------------------------------------------------------------------------

def _patch(owner, name, new):
    _saved.append((owner, name, owner.__dict__.get(name) if isinstance(owner, type) else getattr(owner, name)))
    setattr(owner, name, new)


def _candidates(modules: Optional[Iterable[object]]):
    if modules is not None:
        return list(modules)
    return [m for m in list(sys.modules.values())
            if m is not None and not getattr(m, "__name__", "").startswith("distropointclouds_b200")]


def install(modules: Optional[Iterable[object]] = None, device_index: int = 0,
            patch_voxel: bool = True) -> Dict[str, List[str]]:
    """Patch every reference class found.  Returns {module name: [patched attributes]}.
    patch_voxel=False leaves ScanContextComputer._downsample_voxel to Open3D (where it is installed)."""
    global device
    device = int(device_index)
    done: Dict[str, List[str]] = {}
    seen = set()
    for m in _candidates(modules):
        d = getattr(m, "__dict__", None)
        if not isinstance(d, dict):
            continue
        names: List[str] = []
        sc = d.get("ScanContext")
        if isinstance(sc, type) and sc.__module__ == getattr(m, "__name__", None) and sc not in seen \
                and "ptcloud2sc_vectorized" in sc.__dict__:
            seen.add(sc)
            _patch(sc, "ptcloud2sc_vectorized", _ptcloud2sc_vectorized)
            names.append("ScanContext.ptcloud2sc_vectorized")
            if "scan_context_to_pointcloud" in sc.__dict__:
                _patch(sc, "scan_context_to_pointcloud", _sc_to_pointcloud_loop)
                names.append("ScanContext.scan_context_to_pointcloud")
            if "cells_covered_by_box" in sc.__dict__:
                _patch(sc, "cells_covered_by_box", _make_cells_covered_by_box(sc))
                names.append("ScanContext.cells_covered_by_box")
        comp = d.get("ScanContextComputer")
        if isinstance(comp, type) and comp.__module__ == getattr(m, "__name__", None) and comp not in seen \
                and "_scan_context" in comp.__dict__:
            seen.add(comp)
            _patch(comp, "_scan_context", _scan_context_b)
            names.append("ScanContextComputer._scan_context")
            if patch_voxel and "_downsample_voxel" in comp.__dict__:
                _patch(comp, "_downsample_voxel", staticmethod(_downsample_voxel))
                names.append("ScanContextComputer._downsample_voxel")
        rec = d.get("ScanContextReconstructor")
        if isinstance(rec, type) and rec.__module__ == getattr(m, "__name__", None) and rec not in seen \
                and "scan_context_to_pointcloud" in rec.__dict__:
            seen.add(rec)
            _patch(rec, "scan_context_to_pointcloud", _rec_sc_to_pointcloud)
            _patch(rec, "add_covered_cells_to_pointcloud", _rec_add_covered_cells)
            names += ["ScanContextReconstructor.scan_context_to_pointcloud",
                      "ScanContextReconstructor.add_covered_cells_to_pointcloud"]
        if callable(d.get("to_tensor01")) and isinstance(d.get("TinyCNN"), type) \
                and getattr(d["to_tensor01"], "__module__", None) == getattr(m, "__name__", None):
            _patch(m, "to_tensor01", _to_tensor01)
            names.append("to_tensor01")
        if names:
            done[getattr(m, "__name__", repr(m))] = names
    return done


def uninstall() -> None:
    while _saved:
        owner, name, old = _saved.pop()
        setattr(owner, name, old)

"""Drop-in mirror of reconstruct_pointcloud.py's ScanContextReconstructor on the CUDA engine.

Numeric members only (K3 inverse expansion, K5 covered-cell gather / np.isin recolour mask);
the Open3D window of the reference (reconstruct_pointcloud.py:166-202) is out of scope.
"""
from __future__ import annotations

import json
import os

import numpy as np

from .engine import default_engine


class ScanContextReconstructor:
    """reconstruct_pointcloud.py:25-229."""

    # statics, same names/defaults as reconstruct_pointcloud.py:32-36
    use_full_height_range = 1
    custom_height_offset = 2.0
    load_json_covered_cells = 1
    device = 0

    def __init__(self, scan_context_dir='test_data'):
        self.scan_context_dir = scan_context_dir
        self.pedestrian_json_dir = './test_data/'
        self.sector_res = 720
        self.ring_res = 160
        self.max_length = 20

    def _eng(self):
        return default_engine(type(self).device)

    # -- reconstruct_pointcloud.py:47-77 -----------------------------------------------------------
    def scan_context_to_pointcloud(self, sc_data):
        """(R,S) grid -> (points (K,3) float64, colors (K,3) float64 all 0.7, bin_ids (K,) int32),
        row-major cell order."""
        sc = np.asarray(sc_data)
        if sc.dtype != np.float32:
            sc32 = sc.astype(np.float32)
            if not np.array_equal(sc32.astype(sc.dtype), sc):
                raise TypeError("grid values must be float32-representable (the sc_*.npy dtype)")
            sc = sc32
        pts, ids = self._eng().reconstruct_host(sc, self.ring_res, self.sector_res, float(self.max_length),
                                                bool(type(self).use_full_height_range),
                                                float(type(self).custom_height_offset))[0]
        colors = np.full((pts.shape[0], 3), 0.7, dtype=float)
        return pts, colors, ids

    # -- reconstruct_pointcloud.py:81-102: the union of every pedestrian's covered_cells list ------
    def load_covered_cells_from_json(self, json_path):
        """All [ring, sector] pairs of a pedestrians_*.json, in file order (duplicates kept: every pedestrian is
        listed twice upstream).  A missing or malformed file yields an empty list, as in the reference."""
        if not os.path.isfile(json_path):
            return []
        try:
            with open(json_path) as f:
                records = json.load(f)
        except (OSError, ValueError):
            return []
        if not isinstance(records, list):
            return []
        return [cell for rec in records if isinstance(rec, dict) for cell in rec.get('covered_cells', [])]

    # -- reconstruct_pointcloud.py:104-130 ----------------------------------------------------------
    def add_covered_cells_to_pointcloud(self, sc_data, covered_cells):
        if not covered_cells:
            return np.empty((0, 3)), np.empty((0, 3))
        sc = np.ascontiguousarray(sc_data, dtype=np.float32)
        pts = self._eng().gather_cells_host(sc, np.array(covered_cells, dtype=int), self.ring_res,
                                            self.sector_res, float(self.max_length))
        if not type(self).use_full_height_range:
            valid = pts[:, 2] > 0
            pts = pts[valid]
            # :126 z - offset in the grid's dtype (float32)
            pts[:, 2] = (pts[:, 2].astype(np.float32) - np.float32(type(self).custom_height_offset)).astype(np.float64)
        colors = np.tile(np.array([[1.0, 0.0, 0.0]]), (pts.shape[0], 1))
        return pts, colors

    # -- reconstruct_pointcloud.py:153-164 (the numeric core of the visualiser) ----------------------
    def covered_mask(self, base_bin_ids, covered_cells):
        """np.isin(base_bin_ids, clip(r)*S + clip(s)) as a bool array."""
        if len(covered_cells) == 0:
            return np.zeros(len(base_bin_ids), dtype=bool)
        return self._eng().covered_mask_host(np.asarray(base_bin_ids, dtype=np.int32),
                                             np.array(covered_cells, dtype=int).reshape(-1, 2), self.ring_res,
                                             self.sector_res, float(self.max_length))

    def reconstruct_with_covered_cells(self, sc_file_path, json_file_path=None):
        """Everything visualize_pointcloud_with_covered_cells computes before opening the window
        (reconstruct_pointcloud.py:132-164): points, colours (covered bins painted red), bin ids."""
        sc_data = np.load(sc_file_path)
        base_points, base_colors, base_bin_ids = self.scan_context_to_pointcloud(sc_data)
        red = 0
        if bool(type(self).load_json_covered_cells) and json_file_path and os.path.exists(json_file_path):
            covered = self.load_covered_cells_from_json(json_file_path)
            if covered:
                m = self.covered_mask(base_bin_ids, covered)
                red = int(m.sum())
                if red:
                    base_colors[m] = np.array([1.0, 0.0, 0.0])
        return base_points, base_colors, base_bin_ids, red

    def process_all_files(self):
        """reconstruct_pointcloud.py:205-229 without the GUI: returns {file: (points, colors, ids, n_red)}."""
        out = {}
        if not os.path.exists(self.scan_context_dir):
            print(f"Scan context directory not found: {self.scan_context_dir}")
            return out
        for npy_file in sorted(f for f in os.listdir(self.scan_context_dir) if f.endswith('.npy')):
            base_name = npy_file.replace('sc_', '').replace('.npy', '')
            out[npy_file] = self.reconstruct_with_covered_cells(
                os.path.join(self.scan_context_dir, npy_file),
                os.path.join(self.pedestrian_json_dir, f"pedestrians_{base_name}.json"))
        return out

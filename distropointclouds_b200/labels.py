"""Pedestrian label projection onto the scan-context grid (distro_model/data_process/kitti_dataset.py:129-253).

``cells_covered_by_box`` -- the 115 200-cell float64 test per box -- runs on the GPU (K2);
the per-box scalar arithmetic (centre cell, distance, angle) stays on the host in NumPy with
the very expressions the reference uses, as SURVEY 8(a) a4/a6 prescribes.
"""
from __future__ import annotations

import json
import os
from typing import List, Optional, Sequence

import numpy as np

from .engine import default_engine

PEDESTRIAN_ID = 1   # config/kitti_config.py:5-11  CLASS_NAME_TO_ID['Pedestrian']


def _statics():
    from .scan_context import ScanContext
    return ScanContext


def pedestrian_to_scan_context_coords(pedestrian_x, pedestrian_y):
    """kitti_dataset.py:129-159 -- true division then int() truncation, in the dtype of the inputs."""
    SC = _statics()
    num_ring = SC.ring_res[0]
    num_sector = SC.sector_res[0]
    max_length = SC.max_length
    gap_ring = max_length / num_ring
    gap_sector = 360 / num_sector
    theta = np.rad2deg(np.arctan2(pedestrian_y, pedestrian_x)) % 360
    r = np.sqrt(pedestrian_x ** 2 + pedestrian_y ** 2)
    ring_idx = int(np.minimum(r / gap_ring, num_ring - 1))
    sector_idx = int(theta / gap_sector)
    return ring_idx, sector_idx


def cells_covered_by_boxes(boxes_xylwyaw, num_ring=None, num_sector=None, max_length=None, device=None,
                           want_mask=False):
    """Batched K2: (B,5) x,y,l,w,yaw -> list of (K,2) int64 [ring, sector] arrays (row-major)."""
    SC = _statics()
    if num_ring is None: num_ring = int(SC.ring_res[0])
    if num_sector is None: num_sector = int(SC.sector_res[0])
    if max_length is None: max_length = float(SC.max_length)
    eng = default_engine(SC.device if device is None else device)
    return eng.raster_boxes_host(np.asarray(boxes_xylwyaw, dtype=np.float64).reshape(-1, 5), int(num_ring),
                                 int(num_sector), float(max_length), want_mask=want_mask)


def cells_covered_by_box(x, y, l, w, yaw_lidar, num_ring=None, num_sector=None, max_length=None):
    """kitti_dataset.py:204-253 -- all cells whose centre lies inside the rotated l x w footprint.
    Returns np.ndarray (K,2) int64 rows [ring_idx, sector_idx] in row-major order."""
    cells, _ = cells_covered_by_boxes([[x, y, l, w, yaw_lidar]], num_ring, num_sector, max_length)
    return cells[0]


def find_pedestrians_in_scan_context(labels):
    """kitti_dataset.py:161-201 -- labels (N,8) [class,x,y,z,h,w,l,yaw] -> list of dicts.
    All pedestrian boxes of the frame are rasterised in one K2 launch."""
    rows = [lab for lab in labels if int(lab[0]) == PEDESTRIAN_ID]
    if not rows:
        return []
    boxes = [[float(l_[1]), float(l_[2]), float(l_[6]), float(l_[5]), float(l_[7])] for l_ in rows]
    covered, _ = cells_covered_by_boxes(boxes)
    out = []
    for lab, cov in zip(rows, covered):
        class_id, x, y, z, h, w, l, yaw = lab
        ring_idx, sector_idx = pedestrian_to_scan_context_coords(x, y)
        out.append({
            'class': int(class_id),
            'x': float(x), 'y': float(y), 'z': float(z),
            'h': float(h), 'w': float(w), 'l': float(l), 'yaw': float(yaw),
            'center_ring_idx': int(ring_idx),
            'center_sector_idx': int(sector_idx),
            'covered_cells': cov.tolist(),
            'distance': float(np.hypot(x, y)),
            'angle': float((np.degrees(np.arctan2(y, x)) % 360.0)),
        })
    return out


def pedestrians_from_boxes(boxes_xyzhwlyaw: Sequence[Sequence[float]], class_id: int = PEDESTRIAN_ID):
    """The calib-derived half of the JSON (kitti_dataset.py:530-567) once x,y,z,h,w,l,yaw_lidar are
    known as Python floats (the camera->LiDAR calibration algebra stays with the caller)."""
    if len(boxes_xyzhwlyaw) == 0:
        return []
    covered, _ = cells_covered_by_boxes([[b[0], b[1], b[5], b[4], b[6]] for b in boxes_xyzhwlyaw])
    out = []
    for b, cov in zip(boxes_xyzhwlyaw, covered):
        x, y, z, h, w, l, yaw = (float(v) for v in b)
        ring_idx, sector_idx = pedestrian_to_scan_context_coords(x, y)
        out.append({
            'class': class_id, 'x': x, 'y': y, 'z': z, 'h': h, 'w': w, 'l': l, 'yaw': yaw,
            'center_ring_idx': int(ring_idx), 'center_sector_idx': int(sector_idx),
            'covered_cells': cov.tolist(),
            'distance': float(np.hypot(x, y)),
            'angle': float((np.degrees(np.arctan2(y, x)) % 360.0)),
        })
    return out


def build_mask_from_json(H: int, W: int, json_path: str) -> np.ndarray:
    """tiny_sc_ped.py:62-73 -- binary mask from the 'covered_cells' lists (out-of-range cells ignored)."""
    m = np.zeros((H, W), dtype=np.uint8)
    if not os.path.exists(json_path):
        return m
    with open(json_path, "r") as f:
        data = json.load(f)
    cells = [c for ped in data for c in ped.get("covered_cells", [])]
    if not cells:
        return m
    c = np.asarray(cells, dtype=np.int64).reshape(-1, 2)
    ok = (c[:, 0] >= 0) & (c[:, 0] < H) & (c[:, 1] >= 0) & (c[:, 1] < W)
    c = c[ok]
    if c.shape[0] == 0:
        return m
    SC = _statics()
    eng = default_engine(SC.device)
    ids = np.arange(H * W, dtype=np.int32)
    # mask[i] = i in {r*W+s}: the K5 membership kernel over the whole grid
    return eng.covered_mask_host(ids, c, H, W, float(SC.max_length)).astype(np.uint8).reshape(H, W)


# ---- label-side training tensors --------------------------------------------------------------
# The dataset classes of the reference rebuild these from the pedestrians_*.json lists in Python on every __getitem__;
# here the per-cell work runs on the device (K6) and only the per-pedestrian scalars stay on the host.

def pedestrian_weight(num_cells: int, distance: float) -> float:
    """distro_model/tiny_sc_pred_v2.py:64-76 -- adaptive weight of one pedestrian (Python-float arithmetic)."""
    if num_cells > 0:
        size_weight = min(100.0 / num_cells, 5.0)
        distance_weight = 1.0 + (distance / 20.0) ** 2
        return min(size_weight * distance_weight, 10.0)
    return 1.0


def _csr_cells(ped_list):
    cells = [np.asarray(p.get('covered_cells', []), dtype=np.int64).reshape(-1, 2) for p in ped_list]
    offs = np.zeros(len(cells) + 1, np.int64)
    np.cumsum([c.shape[0] for c in cells], out=offs[1:])
    flat = np.concatenate(cells, axis=0) if cells else np.zeros((0, 2), np.int64)
    return flat, offs


def weight_and_mask_maps(ped_list, H: int = 160, W: int = 720, device=None):
    """ScanContextPedestrianDataset.__getitem__'s mask and adaptive weight map (tiny_sc_pred_v2.py:57-85):
    mask[r,s] = 1 on every in-range covered cell, weight[r,s] = max(1, weights of the pedestrians covering it).
    Returns (mask float32 (H,W), weight_map float32 (H,W))."""
    SC = _statics()
    eng = default_engine(SC.device if device is None else device)
    cells, offs = _csr_cells(ped_list)
    w = np.array([pedestrian_weight(int(offs[i + 1] - offs[i]), p.get('distance', 10.0)) for i, p in enumerate(ped_list)],
                 dtype=np.float64).astype(np.float32)
    mask, weight = eng.label_maps_host(cells, offs, w, None, 1, H, W, mode=0, want_weight=True)
    return mask[0], weight[0]


def clip_wrap_mask(ped_list, H: int = 160, W: int = 720, device=None) -> np.ndarray:
    """ScanContextPedDataset.__getitem__'s optional segmentation mask (train_sc_ped.py:175-183): ring indices
    clipped to [0, H-1], sector indices wrapped with np.mod."""
    SC = _statics()
    eng = default_engine(SC.device if device is None else device)
    cells, offs = _csr_cells(ped_list)
    mask, _ = eng.label_maps_host(cells, offs, None, None, 1, H, W, mode=1, want_weight=False)
    return mask[0]


def gaussian_table(sigma: float, H: int = 160, W: int = 720) -> np.ndarray:
    """exp(-0.5 * d2 / sigma**2) for the integer squared distances d2 = 0 .. (H-1)^2 + (W-1)^2, evaluated with the
    expression of gaussian2d (train_sc_ped.py:78-79) and cut after the last non-zero value."""
    d2 = np.arange((H - 1) ** 2 + (W - 1) ** 2 + 1)
    with np.errstate(under="ignore"):
        t = np.exp(-0.5 * d2 / (sigma ** 2))
    nz = np.flatnonzero(t)
    return np.ascontiguousarray(t[: int(nz[-1]) + 1] if nz.size else t[:1])


def center_heatmap(ped_list, H: int = 160, W: int = 720, sigma: float = 1.8, device=None) -> np.ndarray:
    """ScanContextPedDataset.__getitem__'s centre heat-map (train_sc_ped.py:165-173): np.maximum over the pedestrians
    whose centre cell lies in the grid of gaussian2d((H,W), centre, sigma).  float64 as soon as one centre is valid
    (np.maximum promotes), float32 zeros otherwise -- the reference's dtypes."""
    centers = np.array([[int(p.get("center_ring_idx", -1)), int(p.get("center_sector_idx", -1))] for p in ped_list],
                       dtype=np.int64).reshape(-1, 2)
    ok = (centers[:, 0] >= 0) & (centers[:, 0] < H) & (centers[:, 1] >= 0) & (centers[:, 1] < W)
    if not ok.any():
        return np.zeros((H, W), dtype=np.float32)
    SC = _statics()
    eng = default_engine(SC.device if device is None else device)
    return eng.label_heatmap_host(centers[ok], gaussian_table(sigma, H, W), None, 1, H, W)[0]

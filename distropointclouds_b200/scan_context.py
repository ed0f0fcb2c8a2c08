"""Drop-in mirrors of the reference's scan-context generators, running on the CUDA engine.

Same class names, class statics, method names, argument meaning, return dtypes and error
behaviour as the reference, so the reference's scripts keep working when these classes are
patched in (``distropointclouds_b200.compat.install()``):

* :class:`ScanContext`          -- scaner_version2.py:25-262 (offline generator, variant A);
  also accepts the KITTI-dataset constructor form ``ScanContext(ptcloud_xyz)`` of
  distro_model/data_process/kitti_dataset.py:24-127 and carries its label static methods.
* :class:`ScanContextComputer`  -- scan_context_node.py:24-188 (live generator, variant B).

GUI / plotting members of the reference (Open3D windows, matplotlib PNGs) are out of scope and
not mirrored.  float32 clouds (the .bin dtype) are binned in float32, float64 clouds (what Open3D's
voxel_down_sample returns) in float64, exactly as NumPy promotes the reference's expressions.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from . import _ffi
from .engine import Engine, ScanContextParams, default_engine, pack_scans, variant_a_gap_dtypes
from . import labels as _labels

_STATS = _ffi.STATS_FIELDS


def _is_np_int(v) -> bool:
    return isinstance(v, np.integer)


def _raise_like_reference(stats_row):
    """The reference indexes its grid with garbage for such points (scaner_version2.py:105-108)
    and dies with IndexError; mirror that."""
    bad = int(stats_row[_STATS.index("n_nonfinite")])
    over = int(stats_row[_STATS.index("n_sector_overflow")])
    if bad:
        raise IndexError(f"index -2147483648 is out of bounds ({bad} point(s) with non-finite polar coordinates)")
    if over:
        raise IndexError(f"sector index out of bounds ({over} point(s) whose angle rounds to 360.0)")


class ScanContext:
    """Offline generator (variant A).  scaner_version2.py:25-262 / kitti_dataset.py:24-253."""

    # static variables -- same names and defaults as scaner_version2.py:28-47
    viz = 0
    store_pointcloud = 0
    use_downsampling = 0
    downcell_size = 0.2
    use_full_height_range = 1
    custom_height_offset = 2.0
    load_json_covered_cells = 0
    sector_res = np.array([720])
    ring_res = np.array([160])
    max_length = 20

    # engine-side switches (not in the reference)
    strict = True            # raise IndexError where the reference would
    device = 0
    last_stats = None        # per-scan counters of the last ptcloud2sc_vectorized call

    def __init__(self, bin_dir, bin_file_name: Optional[str] = None):
        if bin_file_name is None:
            # kitti_dataset.py:43-45 form: ScanContext(ptcloud_xyz)
            self.ptcloud_xyz = bin_dir
            self.bin_dir = self.bin_file_name = self.bin_path = None
        else:
            # scaner_version2.py:49-54 form (note: plain string concatenation, as upstream)
            self.bin_dir = bin_dir
            self.bin_file_name = bin_file_name
            self.bin_path = bin_dir + bin_file_name
            self.ptcloud_xyz = None
        self.scancontexts = self.genSCs()

    # -- scaner_version2.py:57-62 ----------------------------------------------------------------
    def load_velo_scan(self):
        scan = np.fromfile(self.bin_path, dtype=np.float32)
        scan = scan.reshape((-1, 4))
        return scan[:, :-1]

    @classmethod
    def _engine(cls) -> Engine:
        return default_engine(cls.device)

    # -- scaner_version2.py:64-115 ---------------------------------------------------------------
    def ptcloud2sc_vectorized(self, ptcloud, num_sector, num_ring, max_length):
        """(N,3) points -> (num_ring, num_sector) float64 grid, empty cells 0.  float32 points (the .bin dtype)
        give float32-representable values; float64 points (use_downsampling=1, :126-133) are binned in float64
        and keep their float64 heights.  Callers cast to float32 on save."""
        cls = type(self) if isinstance(self, ScanContext) else ScanContext
        params = ScanContextParams(
            num_ring=int(num_ring), num_sector=int(num_sector), max_length=float(max_length),
            variant=_ffi.VARIANT_A, full_height=bool(cls.use_full_height_range),
            height_offset=float(cls.custom_height_offset),
            # np.int64 resolutions (the class statics) make the gaps np.float64 -> float64 `//`;
            # Python ints make them weak scalars -> float32 `//`   (scaner_version2.py:70-71,93-94)
            gap_f32=variant_a_gap_dtypes(num_sector, num_ring, max_length),
        )
        grid, stats = cls._engine().build_any(ptcloud, params)
        ScanContext.last_stats = dict(zip(_STATS, (int(v) for v in stats)))
        if cls.strict:
            _raise_like_reference(stats)
        return grid

    # -- scaner_version2.py:117-178 / kitti_dataset.py:100-127 ------------------------------------
    def genSCs(self):
        ptcloud_xyz = self.load_velo_scan() if self.ptcloud_xyz is None else self.ptcloud_xyz
        if bool(type(self).use_downsampling):                # scaner_version2.py:126-133: float64 voxel means (K7)
            ptcloud_xyz, _ = type(self)._engine().voxel_downsample_host(np.asarray(ptcloud_xyz), float(type(self).downcell_size))
        SCs = []
        for res in range(len(type(self).sector_res)):
            num_sector = type(self).sector_res[res]
            num_ring = type(self).ring_res[res]
            SCs.append(self.ptcloud2sc_vectorized(ptcloud_xyz, num_sector, num_ring, type(self).max_length))
        return SCs

    # -- scaner_version2.py:216-262 (loop twin of the reconstructor; returns points, colors) ------
    def scan_context_to_pointcloud(self, sc_data):
        cls = type(self)
        R, S = int(cls.ring_res[0]), int(cls.sector_res[0])
        sc = np.asarray(sc_data)
        pts, _ = cls._engine().reconstruct_host(sc.astype(np.float32), R, S, float(cls.max_length),
                                                bool(cls.use_full_height_range),
                                                float(cls.custom_height_offset))[0]
        colors = np.full((pts.shape[0], 3), 0.7, dtype=float)
        if pts.shape[0] == 0:
            return np.array([]), np.array([])
        return pts, colors

    # -- label side (kitti_dataset.py:129-253) ----------------------------------------------------
    pedestrian_to_scan_context_coords = staticmethod(_labels.pedestrian_to_scan_context_coords)
    find_pedestrians_in_scan_context = staticmethod(_labels.find_pedestrians_in_scan_context)
    cells_covered_by_box = staticmethod(_labels.cells_covered_by_box)


def build_scan_contexts(scans, num_ring=160, num_sector=720, max_length=20.0, full_height=True,
                        height_offset=2.0, variant=_ffi.VARIANT_A, device=0, want_norm=False):
    """Batch API: list of (N,3|4) float32 arrays -> (n, R, S) float32 grids (+ normalised, stats)."""
    params = ScanContextParams(num_ring, num_sector, max_length, variant, full_height, height_offset)
    return default_engine(device).build_scans(list(scans), params, want_norm=want_norm)


class ScanContextComputer:
    """Live generator (variant B).  scan_context_node.py:24-188."""

    def __init__(self, voxel_size: float = 0.20, max_length: float = 80.0, sector_res: int = 720,
                 ring_res: int = 160, use_full_height_range: bool = False, custom_height_offset: float = 2.0,
                 visualize: bool = False, store_pointcloud_png: bool = False, output_dir: str = "scan_results",
                 dl_images_dir: str = "scan_dl_images_live", cmap: str = "viridis", device: int = 0):
        self.voxel_size = voxel_size
        self.max_length = max_length
        self.sector_res = sector_res
        self.ring_res = ring_res
        self.use_full_height_range = bool(use_full_height_range)
        self.custom_height_offset = float(custom_height_offset)
        self.visualize = visualize
        self.store_pointcloud_png = store_pointcloud_png
        self.output_dir = output_dir
        self.dl_images_dir = dl_images_dir
        self.cmap = cmap
        self.device = device
        self.last_stats = None
        os.makedirs(self.output_dir, exist_ok=True)       # scan_context_node.py:57-58
        os.makedirs(self.dl_images_dir, exist_ok=True)

    # -- scan_context_node.py:93-100 --------------------------------------------------------------
    @staticmethod
    def _downsample_voxel(pts_xyz: np.ndarray, voxel: float) -> np.ndarray:
        """Open3D's voxel_down_sample on the GPU (K7): float64 voxel means, like np.asarray(d.points).  Open3D is not
        installed here, so this restates its published algorithm (parity unpinned); the points come out sorted by
        voxel index instead of in Open3D's hash-map order, which the scan context does not depend on."""
        pts = np.asarray(pts_xyz)
        if pts.size == 0:
            return pts_xyz
        out, _ = default_engine(ScanContextComputer.device_default).voxel_downsample_host(pts, float(voxel))
        return out

    device_default = 0

    # -- scan_context_node.py:102-157 -------------------------------------------------------------
    def _scan_context(self, pts_xyz: np.ndarray) -> np.ndarray:
        pts = np.asarray(pts_xyz)
        if pts.size == 0:                                   # :107-108
            return np.zeros((self.ring_res, self.sector_res), dtype=np.float32)
        params = ScanContextParams(int(self.ring_res), int(self.sector_res), float(self.max_length),
                                   _ffi.VARIANT_B, self.use_full_height_range, self.custom_height_offset)
        grid, stats = default_engine(self.device).build_any(pts, params)      # float32 or float64 cloud
        self.last_stats = dict(zip(_STATS, (int(v) for v in stats)))
        return grid

    # -- scan_context_node.py:159-188 (numeric part: down-sample, grid, .npy; the PNG render is out of scope) --
    def compute_and_save(self, pts_xyz: np.ndarray, stamp_stem: str):
        pts_ds = self._downsample_voxel(pts_xyz, self.voxel_size)     # :161  float64 voxel means
        sc = self._scan_context(pts_ds)                               # :168  binned in float64 -> float32 grid
        out_npy = os.path.join(self.output_dir, f"sc_{stamp_stem}.npy")
        np.save(out_npy, sc.astype(np.float32))                       # :185-186
        return sc, None, out_npy

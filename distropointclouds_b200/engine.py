"""Python front-end of the C-ABI engine: one ``Engine`` per GPU.

Host arrays are NumPy; device arrays are torch CUDA tensors (torch is only the allocator and
stream provider here -- every kernel that runs is in libdsc.so).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import weakref
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi


@dataclass(frozen=True)
class ScanContextParams:
    """The statics the reference keeps on its classes (scaner_version2.py:28-47,
    kitti_dataset.py:28-41, reconstruct_pointcloud.py:32-45, scan_context_node.py:31-45)."""
    num_ring: int = 160
    num_sector: int = 720
    max_length: float = 20.0
    variant: int = _ffi.VARIANT_A
    full_height: bool = True
    height_offset: float = 2.0
    gap_f32: object = False     # variant A: which `//` NumPy runs in float32 -- False: none, True / 3: both gaps,
    #                             1: the ring gap only, 2: the sector gap only (see variant_a_gap_dtypes)
    exact_only: bool = False    # disable the screened fast path

    def c(self) -> _ffi.DscParams:
        g = 3 if self.gap_f32 is True else int(self.gap_f32)
        flags = ((_ffi.FLAG_GAP_RING_F32 if g & 1 else 0) | (_ffi.FLAG_GAP_SECTOR_F32 if g & 2 else 0) |
                 (_ffi.FLAG_EXACT_ONLY if self.exact_only else 0))
        return _ffi.DscParams(int(self.num_ring), int(self.num_sector), float(self.max_length),
                              int(self.variant), int(bool(self.full_height)), float(self.height_offset), flags)

    @property
    def cells(self) -> int:
        return int(self.num_ring) * int(self.num_sector)


def variant_a_gap_dtypes(num_sector, num_ring, max_length) -> int:
    """Which of variant A's two floor-divides NumPy evaluates in float32 for a float32 cloud (scaner_version2.py:70-71,
    93-94): ``gap_ring = max_length / num_ring`` and ``gap_sector = 360 / num_sector`` are np.float64 when the resolution
    (or max_length) is a NumPy scalar -- the class statics are np.int64 -- and the `//` then runs in float64; with plain
    Python numbers the gap is a weak Python float and the `//` stays float32.  Under NumPy 1.x (value-based casting) a
    float64 *scalar* does not upcast a float32 array either, so both run in float32.  Returns the gap_f32 mask
    (bit 0: ring gap, bit 1: sector gap); it only matters for gaps that are not powers of two."""
    if int(np.__version__.split(".")[0]) < 2:
        return 3
    ring64 = isinstance(num_ring, np.generic) or isinstance(max_length, np.generic)
    sector64 = isinstance(num_sector, np.generic)
    return (0 if ring64 else 1) | (0 if sector64 else 2)


def polar_tables(num_ring: int, num_sector: int, max_length: float):
    """Cell-centre tables built with NumPy exactly as the reference builds them
    (kitti_dataset.py:219-229; reconstruct_pointcloud.py:52-59)."""
    gap_ring = max_length / num_ring
    gap_sector = 360.0 / num_sector
    r_centers = (np.arange(num_ring) + 0.5) * gap_ring
    theta = np.deg2rad((np.arange(num_sector) + 0.5) * gap_sector)
    return (np.ascontiguousarray(r_centers, dtype=np.float64),
            np.ascontiguousarray(np.cos(theta), dtype=np.float64),
            np.ascontiguousarray(np.sin(theta), dtype=np.float64))


def _ptr(a) -> Optional[int]:
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a.data_ptr())      # torch tensor


def pack_scans(scans: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray, int]:
    """Concatenate per-scan (N,3|4) float32 arrays into one packed buffer.
    Returns (buffer, offsets int64[n+1], stride_bytes).  The reference's ``scan[:, :-1]`` view of
    an (N,4) file buffer is passed through without a copy when there is a single scan."""
    if len(scans) == 0:
        return np.zeros((0, 4), np.float32), np.zeros(1, np.int64), 16
    cols = {int(s.shape[1]) for s in scans}
    for s in scans:
        if s.dtype != np.float32:
            raise TypeError("scan-context engine takes float32 points (the .bin dtype); got %s" % s.dtype)
        if s.ndim != 2 or s.shape[1] not in (3, 4):
            raise ValueError("points must be (N,3) or (N,4)")
    offsets = np.zeros(len(scans) + 1, np.int64)
    np.cumsum([s.shape[0] for s in scans], out=offsets[1:])
    if len(scans) == 1:
        s = scans[0]
        if s.shape[0] == 0:
            return np.zeros((0, 4), np.float32), offsets, 16
        if s.flags.c_contiguous:
            return s, offsets, 4 * s.shape[1]
        if s.shape[1] == 3 and s.strides == (16, 4):
            # strided xyz view of packed xyzi rows: hand the rows over as they lie in memory.
            # The last row's padding float may lie outside the view, so only do it when the
            # base buffer really owns 4 floats per row.
            base = s.base
            if isinstance(base, np.ndarray) and base.dtype == np.float32 and base.flags.c_contiguous \
                    and base.size >= 4 * s.shape[0] and base.ctypes.data == s.ctypes.data:
                return base.reshape(-1)[: 4 * s.shape[0]].reshape(-1, 4), offsets, 16
    stride_cols = 4 if cols == {4} else 3
    buf = np.empty((int(offsets[-1]), stride_cols), np.float32)
    for s, a, b in zip(scans, offsets[:-1], offsets[1:]):
        buf[a:b] = s[:, :stride_cols]
    return buf, offsets, 4 * stride_cols


def pack_scans_f64(scans: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray, int]:
    """float64 clouds (Open3D-downsampled: np.asarray(pcd.points)) -> one packed (N_total, 3) float64 buffer,
    offsets, stride in bytes.  A single C-contiguous (N,3|4) array is passed through without a copy."""
    offsets = np.zeros(len(scans) + 1, np.int64)
    if len(scans) == 0:
        return np.zeros((0, 3), np.float64), offsets, 24
    for s in scans:
        if s.ndim != 2 or s.shape[1] not in (3, 4):
            raise ValueError("points must be (N,3) or (N,4)")
    np.cumsum([s.shape[0] for s in scans], out=offsets[1:])
    if len(scans) == 1 and scans[0].dtype == np.float64 and scans[0].flags.c_contiguous and scans[0].shape[0]:
        return scans[0], offsets, 8 * scans[0].shape[1]
    buf = np.empty((int(offsets[-1]), 3), np.float64)
    for s, a, b in zip(scans, offsets[:-1], offsets[1:]):
        buf[a:b] = s[:, :3]
    return buf, offsets, 24


class Engine:
    """Owns one dsc_ctx (one GPU).  Not thread-safe; must not be used from forked children."""

    def __init__(self, device: int = 0):
        self._lib = _ffi.load()
        h = C.c_void_p()
        rc = self._lib.dsc_create(int(device), C.byref(h))
        if rc != _ffi.DSC_OK:
            raise _ffi.DscError(rc, "dsc_create failed (no CUDA device? the engine has no CPU fallback)")
        self._handle = h
        self._pid = os.getpid()
        self.device = int(device)
        self._tables_key = None
        self._pinned = {}

    @property
    def _h(self):
        # a CUDA context does not survive fork(): DataLoader(num_workers>0) children must not use the
        # parent's engine (kitti_dataloader.py builds its loader with fork workers)
        if self._pid != os.getpid():
            raise RuntimeError("dsc engine used from a forked child process: CUDA cannot be used after fork(); "
                               "run the DataLoader with num_workers=0 or multiprocessing_context='spawn'")
        if self._handle is None:
            raise RuntimeError("dsc engine is closed")
        return self._handle

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_handle", None) is not None and self._pid == os.getpid():
            for addr, fin in list(getattr(self, "_pinned", {}).items()):      # pinned buffers die with the engine
                fin.detach()
                self._free_pinned(addr)
            self._lib.dsc_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int, allow=()):
        if rc != _ffi.DSC_OK and rc not in allow:
            raise _ffi.DscError(rc, self._lib.dsc_last_error(self._h).decode())
        return rc

    @property
    def version(self) -> str:
        return self._lib.dsc_version().decode()

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.dsc_kernel_launches(self._h))

    def set_option(self, name: str, value: int):
        self._check(self._lib.dsc_set_option(self._h, name.encode(), int(value)))

    def set_timing(self, enabled: bool):
        self._check(self._lib.dsc_set_timing(self._h, int(enabled)))

    @property
    def last_kernel_ms(self) -> float:
        return float(self._lib.dsc_last_kernel_ms(self._h))

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """NumPy array over page-locked host memory.  The allocation is released when the array (and every view of
        it) has been garbage-collected, or when the engine is closed, whichever comes first -- do not touch such an
        array after ``close()``."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._check(self._lib.dsc_host_alloc(self._h, n, C.byref(p)))
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        addr = int(p.value)
        self._pinned[addr] = weakref.finalize(buf, self._free_pinned, addr)     # buf lives as long as arr.base does
        return arr

    def _free_pinned(self, addr: int):
        if self._pinned.pop(addr, None) is not None and self._handle is not None and self._pid == os.getpid():
            self._lib.dsc_host_free(self._handle, C.c_void_p(addr))

    # -- K1 ------------------------------------------------------------------------------------
    def build_host(self, pts: np.ndarray, offsets: np.ndarray, params: ScanContextParams,
                   stride_bytes: Optional[int] = None, want_norm: bool = False, want_stats: bool = True,
                   out: Optional[np.ndarray] = None, out_norm: Optional[np.ndarray] = None):
        """Host buffers in, host buffers out (H2D/D2H inside).  Returns (grids, norm|None, stats|None)."""
        pts = np.asarray(pts)
        if pts.dtype != np.float32 or not pts.flags.c_contiguous:
            raise TypeError("pts must be a C-contiguous float32 array")
        if stride_bytes is None:
            stride_bytes = 4 * pts.shape[1]
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.size - 1
        if out is None:
            out = np.empty((n, params.num_ring, params.num_sector), np.float32)
        if want_norm and out_norm is None:
            out_norm = np.empty_like(out)
        stats = np.zeros((n, len(_ffi.STATS_FIELDS)), np.int64) if want_stats else None
        cp = params.c()
        self._check(self._lib.dsc_sc_build_batch(self._h, C.byref(cp), _ptr(pts), _ptr(offsets), n,
                                                 int(stride_bytes), _ffi.MEM_HOST, _ptr(out),
                                                 _ptr(out_norm) if want_norm else None, _ptr(stats), None))
        return out, (out_norm if want_norm else None), stats

    def build_device(self, pts, offsets: np.ndarray, params: ScanContextParams, out, out_norm=None,
                     stats=None, stride_bytes: int = 16, stream: Optional[int] = None):
        """Device-resident batch.  ``pts``/``out``/``out_norm``/``stats`` are torch CUDA tensors
        (stats: int64 [n,6]); ``offsets`` is a host int64 array.  Stream-ordered, asynchronous."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.size - 1
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
        cp = params.c()
        self._check(self._lib.dsc_sc_build_batch(self._h, C.byref(cp), _ptr(pts), _ptr(offsets), n,
                                                 int(stride_bytes), _ffi.MEM_DEVICE, _ptr(out), _ptr(out_norm),
                                                 _ptr(stats), C.c_void_p(stream)))

    def build_scans(self, scans: Sequence[np.ndarray], params: ScanContextParams, want_norm=False):
        pts, offsets, stride = pack_scans(scans)
        return self.build_host(pts, offsets, params, stride_bytes=stride, want_norm=want_norm)

    # -- K1, float64 clouds --------------------------------------------------------------------
    def build_host_f64(self, pts: np.ndarray, offsets: np.ndarray, params: ScanContextParams,
                       stride_bytes: Optional[int] = None, want_stats: bool = True):
        """float64 points (what Open3D's voxel_down_sample returns; scan_context_node.py:93-100,161-168,
        scaner_version2.py:126-133): the polar chain runs in float64.  Returns (grids, stats|None); the grids
        are float64 for variant A (float64 heights) and float32 for variant B (heights cast before the max)."""
        pts = np.asarray(pts)
        if pts.dtype != np.float64 or not pts.flags.c_contiguous:
            raise TypeError("pts must be a C-contiguous float64 array")
        if stride_bytes is None:
            stride_bytes = 8 * pts.shape[1]
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.size - 1
        out = np.empty((n, params.num_ring, params.num_sector),
                       np.float64 if params.variant == _ffi.VARIANT_A else np.float32)
        stats = np.zeros((n, len(_ffi.STATS_FIELDS)), np.int64) if want_stats else None
        cp = params.c()
        self._check(self._lib.dsc_sc_build_batch_f64(self._h, C.byref(cp), float(params.height_offset), _ptr(pts),
                                                     _ptr(offsets), n, int(stride_bytes), _ffi.MEM_HOST, _ptr(out),
                                                     _ptr(stats), None))
        return out, stats

    def build_device_f64(self, pts, offsets: np.ndarray, params: ScanContextParams, out, stats=None,
                         stride_bytes: int = 24, stream: Optional[int] = None):
        """Device-resident float64 batch (torch CUDA tensors; ``out`` float64 for variant A, float32 for B)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
        cp = params.c()
        self._check(self._lib.dsc_sc_build_batch_f64(self._h, C.byref(cp), float(params.height_offset), _ptr(pts),
                                                     _ptr(offsets), offsets.size - 1, int(stride_bytes),
                                                     _ffi.MEM_DEVICE, _ptr(out), _ptr(stats), C.c_void_p(stream)))

    def build_scans_f64(self, scans: Sequence[np.ndarray], params: ScanContextParams):
        pts, offsets, stride = pack_scans_f64([np.asarray(s) for s in scans])
        return self.build_host_f64(pts, offsets, params, stride_bytes=stride)

    def build_any(self, points: np.ndarray, params: ScanContextParams):
        """One cloud of either dtype -> (grid, stats row), binned in the dtype NumPy would use: float32 clouds in
        float32 (K1), everything else (float64, integer, float16 ...) promoted to float64 as NumPy's
        x*x, sqrt and arctan2 would.  Variant A grids come back as float64 either way, like the reference's."""
        pts = np.asarray(points)
        if pts.dtype == np.float32:
            grids, _, stats = self.build_scans([pts], params)
            g = grids[0].astype(np.float64) if params.variant == _ffi.VARIANT_A else grids[0]
            return g, stats[0]
        if pts.dtype != np.float64:
            if pts.dtype == np.float16 or not np.can_cast(pts.dtype, np.float64, "safe"):
                raise TypeError(f"cannot bin points of dtype {pts.dtype}")
            pts = pts.astype(np.float64)
        grids, stats = self.build_scans_f64([pts], params)
        return grids[0], stats[0]

    # -- K4 ------------------------------------------------------------------------------------
    def normalize01_host(self, grids: np.ndarray) -> np.ndarray:
        g = np.ascontiguousarray(grids, dtype=np.float32)
        g3 = g.reshape((-1,) + g.shape[-2:]) if g.ndim >= 2 else g.reshape(1, -1)
        out = np.empty_like(g3)
        cells = int(np.prod(g3.shape[1:]))
        self._check(self._lib.dsc_sc_normalize01(self._h, _ptr(g3), g3.shape[0], cells, _ffi.MEM_HOST,
                                                 _ptr(out), None))
        return out.reshape(g.shape)

    def normalize01_device(self, grids, out, stream: Optional[int] = None):
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
        n = int(grids.shape[0])
        cells = int(grids.numel() // max(n, 1))
        self._check(self._lib.dsc_sc_normalize01(self._h, _ptr(grids), n, cells, _ffi.MEM_DEVICE, _ptr(out),
                                                 C.c_void_p(stream)))

    # -- tables --------------------------------------------------------------------------------
    def set_tables(self, num_ring: int, num_sector: int, max_length: float):
        key = (int(num_ring), int(num_sector), float(max_length))
        if key == self._tables_key:
            return
        rc_, ct, st = polar_tables(*key)
        self._check(self._lib.dsc_set_polar_tables(self._h, key[0], key[1], _ptr(rc_), _ptr(ct), _ptr(st)))
        self._tables_key = key

    # -- K2 ------------------------------------------------------------------------------------
    def raster_boxes_host(self, boxes_xylwyaw: np.ndarray, num_ring=160, num_sector=720, max_length=20.0,
                          want_mask: bool = False):
        """boxes: (B,5) float64 x,y,l,w,yaw.  Returns (list of (K,2) int64 arrays, mask|None)."""
        self.set_tables(num_ring, num_sector, max_length)
        b = np.asarray(boxes_xylwyaw, dtype=np.float64).reshape(-1, 5)
        nb = b.shape[0]
        packed = np.empty((nb, 6), np.float64)
        packed[:, :4] = b[:, :4]
        packed[:, 4] = np.cos(b[:, 4])      # kitti_dataset.py:236
        packed[:, 5] = np.sin(b[:, 4])      # kitti_dataset.py:237
        offs = np.zeros(nb + 1, np.int64)
        mask = np.zeros((num_ring, num_sector), np.uint8) if want_mask else None
        cap = max(1024, 256 * nb)
        while True:
            cells = np.empty((cap, 2), np.int32)
            rc = self._lib.dsc_raster_boxes(self._h, _ptr(packed), nb, _ffi.MEM_HOST, _ptr(offs), _ptr(cells),
                                            cap, _ptr(mask), None)
            if rc == _ffi.DSC_ERR_CAPACITY:
                cap = int(offs[nb])
                continue
            self._check(rc)
            break
        c64 = cells[: int(offs[nb])].astype(np.int64)            # one conversion, then views per box
        out = [c64[offs[i]:offs[i + 1]] for i in range(nb)]
        return out, mask

    # -- K3 ------------------------------------------------------------------------------------
    def reconstruct_host(self, grids: np.ndarray, num_ring=160, num_sector=720, max_length=20.0,
                         full_height=True, height_offset=2.0):
        """grids (n,R,S) or (R,S) float32.  Returns list of (points (K,3) f64, bin_ids (K,) int32)."""
        self.set_tables(num_ring, num_sector, max_length)
        g = np.ascontiguousarray(grids, dtype=np.float32).reshape(-1, num_ring, num_sector)
        n = g.shape[0]
        offs = np.zeros(n + 1, np.int64)
        cap = int(np.count_nonzero(g)) if full_height else int(np.count_nonzero(g > 0))
        pts = np.empty((max(cap, 1), 3), np.float64)
        ids = np.empty((max(cap, 1),), np.int32)
        self._check(self._lib.dsc_reconstruct_batch(self._h, _ptr(g), n, int(bool(full_height)),
                                                    float(height_offset), _ffi.MEM_HOST, _ptr(offs), _ptr(pts),
                                                    _ptr(ids), cap, None))
        # views of the two result buffers (one allocation per call, no per-grid copies)
        return [(pts[offs[i]:offs[i + 1]], ids[offs[i]:offs[i + 1]]) for i in range(n)]

    def reconstruct_device(self, grids, pt_offsets, pts_out, bin_ids_out, num_ring=160, num_sector=720,
                           max_length=20.0, full_height=True, height_offset=2.0, stream: Optional[int] = None):
        """All torch CUDA tensors; capacity is pts_out.shape[0]; pt_offsets int64 [n+1] on device."""
        self.set_tables(num_ring, num_sector, max_length)
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
        n = int(grids.shape[0])
        self._check(self._lib.dsc_reconstruct_batch(self._h, _ptr(grids), n, int(bool(full_height)),
                                                    float(height_offset), _ffi.MEM_DEVICE, _ptr(pt_offsets),
                                                    _ptr(pts_out), _ptr(bin_ids_out), int(pts_out.shape[0]),
                                                    C.c_void_p(stream)))

    # -- K6 ------------------------------------------------------------------------------------
    def label_maps_host(self, cells: np.ndarray, cell_offsets: np.ndarray, box_weight: Optional[np.ndarray] = None,
                        box_frame: Optional[np.ndarray] = None, n_frames: int = 1, num_ring=160, num_sector=720,
                        mode: int = 0, want_weight: bool = True):
        """CSR covered cells of a batch of boxes -> (mask float32 (n_frames,R,S), weight float32 | None).
        mode 0: out-of-grid cells ignored (tiny_sc_pred_v2.py:81); mode 1: clip ring, wrap sector (train_sc_ped.py:181-182)."""
        c = np.ascontiguousarray(np.asarray(cells).reshape(-1, 2), dtype=np.int32)
        offs = np.ascontiguousarray(cell_offsets, dtype=np.int64)
        nb = offs.size - 1
        w = np.ascontiguousarray(box_weight, dtype=np.float32) if (want_weight and box_weight is not None) else None
        want_weight = want_weight and w is not None
        fr = np.ascontiguousarray(box_frame, dtype=np.int32) if box_frame is not None else None
        mask = np.empty((n_frames, num_ring, num_sector), np.float32)
        weight = np.empty_like(mask) if want_weight else None
        self._check(self._lib.dsc_label_maps(self._h, _ptr(c), _ptr(offs), _ptr(w), _ptr(fr), nb, int(n_frames),
                                             int(num_ring), int(num_sector), int(mode), _ffi.MEM_HOST, _ptr(mask),
                                             _ptr(weight), None))
        return mask, weight

    def label_heatmap_host(self, centers: np.ndarray, table: np.ndarray, center_frame: Optional[np.ndarray] = None,
                           n_frames: int = 1, num_ring=160, num_sector=720) -> np.ndarray:
        """(center_ring_idx, center_sector_idx) rows + the tabulated Gaussian -> float64 (n_frames,R,S) heat-maps."""
        c = np.ascontiguousarray(np.asarray(centers).reshape(-1, 2), dtype=np.int32)
        t = np.ascontiguousarray(table, dtype=np.float64)
        fr = np.ascontiguousarray(center_frame, dtype=np.int32) if center_frame is not None else None
        heat = np.empty((n_frames, num_ring, num_sector), np.float64)
        self._check(self._lib.dsc_label_heatmap(self._h, _ptr(c), _ptr(fr), c.shape[0], int(n_frames), int(num_ring),
                                                int(num_sector), _ptr(t), int(t.size), _ffi.MEM_HOST, _ptr(heat), None))
        return heat

    # -- K7 ------------------------------------------------------------------------------------
    def voxel_downsample_host(self, points: np.ndarray, voxel_size: float):
        """Voxel-grid filter (Open3D's published voxel_down_sample algorithm; parity unpinned, see include/dsc.h):
        (N,>=3) float32 or float64 points -> ((M,3) float64 voxel means sorted by voxel index, n_dropped non-finite)."""
        p = np.asarray(points)
        if p.ndim != 2 or p.shape[1] < 3:
            raise ValueError("points must be (N,>=3)")
        if p.dtype not in (np.float32, np.float64):
            p = p.astype(np.float64)
        if not p.flags.c_contiguous:
            p = np.ascontiguousarray(p)
        n = p.shape[0]
        if n == 0:
            return np.empty((0, 3), np.float64), 0
        out = np.empty((max(n, 1), 3), np.float64)
        n_out, n_drop = C.c_int64(0), C.c_int64(0)
        self._check(self._lib.dsc_voxel_downsample(self._h, _ptr(p), n, int(p.dtype == np.float64), int(p.strides[0]),
                                                   float(voxel_size), _ffi.MEM_HOST, _ptr(out), out.shape[0],
                                                   C.byref(n_out), C.byref(n_drop), None))
        return out[: n_out.value], int(n_drop.value)

    # -- K5 ------------------------------------------------------------------------------------
    def covered_mask_host(self, bin_ids: np.ndarray, cells: np.ndarray, num_ring=160, num_sector=720,
                          max_length=20.0) -> np.ndarray:
        self.set_tables(num_ring, num_sector, max_length)
        ids = np.ascontiguousarray(bin_ids, dtype=np.int32)
        c = np.ascontiguousarray(np.asarray(cells).reshape(-1, 2), dtype=np.int32)
        mask = np.zeros(ids.shape[0], np.uint8)
        self._check(self._lib.dsc_covered_mask(self._h, _ptr(ids), ids.shape[0], _ptr(c), c.shape[0],
                                               _ffi.MEM_HOST, _ptr(mask), None))
        return mask.astype(bool)

    def gather_cells_host(self, grid: np.ndarray, cells: np.ndarray, num_ring=160, num_sector=720,
                          max_length=20.0) -> np.ndarray:
        self.set_tables(num_ring, num_sector, max_length)
        g = np.ascontiguousarray(grid, dtype=np.float32)
        c = np.ascontiguousarray(np.asarray(cells).reshape(-1, 2), dtype=np.int32)
        pts = np.empty((c.shape[0], 3), np.float64)
        self._check(self._lib.dsc_gather_cells(self._h, _ptr(g), _ptr(c), c.shape[0], _ffi.MEM_HOST, _ptr(pts),
                                               None))
        return pts


_default = {}
_default_lock = threading.Lock()


def default_engine(device: int = 0) -> Engine:
    """Process-wide engine per device, created on first use."""
    with _default_lock:
        e = _default.get(device)
        if e is None or e._handle is None or e._pid != os.getpid():
            e = _default[device] = Engine(device)
        return e

"""On-disk formats either side of the hot path.

  in   KITTI Velodyne ``.bin`` files: float32 [x,y,z,i] rows  (scaner_version2.py:57-62, kitti_dataset.py:633-636)
  out  ``sc_XXXXXX.npy``  float32 (R,S) grids written with np.save  (kitti_dataset.py:474-475, scaner_version2.py:417-418)
       ``pedestrians_XXXXXX.json``  json.dump(list, indent=2)        (kitti_dataset.py:491-495)

The batch reader fills ONE page-locked buffer (the layout ``dsc_sc_build_batch`` takes) straight from the
files, so a directory of scans goes to the GPU with a single pipelined host-buffer call instead of one PCIe
round trip per file.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .engine import Engine, ScanContextParams, default_engine

_NPY_MAGIC = b"\x93NUMPY\x01\x00"


def list_bin_files(bin_dir: str) -> List[str]:
    """scaner_version2.py:11-19 (kitti_vlp_database): the sorted *.bin names of a directory."""
    return sorted(f for f in os.listdir(bin_dir) if f.endswith(".bin"))


def read_bin_batch(paths: Sequence[str], engine: Optional[Engine] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Read whole .bin files back to back into one (pinned, when an engine is given) float32 (N_total, 4)
    buffer.  Returns (points, offsets int64[n+1]).  A file whose size is not a multiple of 16 bytes is an error
    (np.fromfile(...).reshape(-1, 4) raises in the reference too)."""
    sizes = [os.path.getsize(p) for p in paths]
    for p, sz in zip(paths, sizes):
        if sz % 16:
            raise ValueError(f"cannot reshape array of size {sz // 4} into shape (4): {p}")
    offsets = np.zeros(len(paths) + 1, np.int64)
    np.cumsum([sz // 16 for sz in sizes], out=offsets[1:])
    total = int(offsets[-1])
    buf = engine.pinned_empty((total, 4), np.float32) if engine is not None else np.empty((total, 4), np.float32)
    flat = buf.reshape(-1).view(np.uint8)
    for p, a, sz in zip(paths, offsets[:-1], sizes):
        with open(p, "rb", buffering=0) as f:
            got = f.readinto(memoryview(flat[int(a) * 16: int(a) * 16 + sz]))
        if got != sz:
            raise IOError(f"short read on {p}: {got} of {sz} bytes")
    return buf, offsets


def npy_header(shape: Tuple[int, ...], dtype="<f4") -> bytes:
    """The NPY v1.0 header np.save writes for a C-ordered array of this shape (padded to a multiple of 64)."""
    d = "{'descr': '%s', 'fortran_order': False, 'shape': %s, }" % (dtype, repr(tuple(int(s) for s in shape)))
    pad = 64 - (len(_NPY_MAGIC) + 2 + len(d) + 1) % 64
    d = d + " " * (pad % 64) + "\n"
    return _NPY_MAGIC + len(d).to_bytes(2, "little") + d.encode("latin1")


def save_sc_npy(path: str, grid: np.ndarray) -> None:
    """np.save(path, grid.astype(np.float32)) without the intermediate copy: header + raw float32 bytes
    (byte-identical to the reference's files)."""
    g = np.ascontiguousarray(grid, dtype=np.float32)
    with open(path, "wb") as f:
        f.write(npy_header(g.shape))
        f.write(memoryview(g).cast("B"))


def save_sc_batch(out_dir: str, stems: Sequence[str], grids: np.ndarray) -> List[str]:
    """One ``sc_<stem>.npy`` per grid of an (n,R,S) batch."""
    os.makedirs(out_dir, exist_ok=True)
    hdr = npy_header(grids.shape[1:])
    g = np.ascontiguousarray(grids, dtype=np.float32)
    out = []
    for stem, a in zip(stems, g):
        p = os.path.join(out_dir, f"sc_{stem}.npy")
        with open(p, "wb") as f:
            f.write(hdr)
            f.write(memoryview(a).cast("B"))
        out.append(p)
    return out


def save_pedestrians_json(path: str, pedestrian_info: List[Dict]) -> None:
    """kitti_dataset.py:491-495."""
    with open(path, "w") as f:
        json.dump(pedestrian_info, f, indent=2)


def process_bin_dir(bin_dir: str, out_dir: str, params: ScanContextParams = ScanContextParams(),
                    device: int = 0, batch_bytes: int = 2 << 30, strict: bool = True) -> List[str]:
    """The numeric part of scaner_version2.py's __main__ + process_dataset_to_images (:430-487, :372-427) for a
    whole directory: every .bin -> sc_<stem>.npy, in batches of at most `batch_bytes` of points.
    strict: raise IndexError (naming the file) where the reference's variant A would -- NaN / Inf coordinates or an
    angle that rounds to 360.0 (its own sample_data/999999.bin does); strict=False drops such points silently."""
    eng = default_engine(device)
    names = list_bin_files(bin_dir)
    written: List[str] = []
    i = 0
    while i < len(names):
        j, tot = i, 0
        while j < len(names) and (j == i or tot + os.path.getsize(os.path.join(bin_dir, names[j])) <= batch_bytes):
            tot += os.path.getsize(os.path.join(bin_dir, names[j]))
            j += 1
        pts, offs = read_bin_batch([os.path.join(bin_dir, n) for n in names[i:j]])
        grids, _, stats = eng.build_host(pts, offs, params, 16, want_stats=strict)
        if strict and params.variant == 0:
            bad = np.flatnonzero((stats[:, 1] + stats[:, 2]) > 0)
            if bad.size:
                k = int(bad[0])
                raise IndexError(f"{names[i + k]}: {int(stats[k, 1])} point(s) with non-finite polar coordinates, "
                                 f"{int(stats[k, 2])} whose angle rounds to 360.0 (the reference raises IndexError here)")
        written += save_sc_batch(out_dir, [n.split(".")[0] for n in names[i:j]], grids)
        i = j
    return written

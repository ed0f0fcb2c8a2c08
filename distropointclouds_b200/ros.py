"""The live path's ingest without NumPy hops through sensor_msgs_py: PointCloud2 -> (N,3) float32.

Mirrors ``ScanContextNode._extract_xyz`` (scan_context_node.py:249-276).  ``rclpy`` / ``sensor_msgs`` are not
installable here, so the function takes any object with the PointCloud2 attributes (``fields`` with ``name`` /
``offset`` / ``datatype``, ``point_step``, ``row_step``, ``width``, ``height``, ``is_bigendian``, ``data``) -- a real
``sensor_msgs.msg.PointCloud2`` is one -- and decodes the byte buffer itself: one strided view per field instead of a
Python generator of tuples.  Like the reference's fast path (``read_points_numpy``) it keeps NaN points; the engine's
voxel filter and variant B drop non-finite points themselves.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

# sensor_msgs/PointField datatype codes
_DT = {1: np.int8, 2: np.uint8, 3: np.int16, 4: np.uint16, 5: np.int32, 6: np.uint32, 7: np.float32, 8: np.float64}
FLOAT32 = 7


@dataclass
class PointField:
    name: str
    offset: int
    datatype: int = FLOAT32
    count: int = 1


@dataclass
class PointCloud2:
    """Plain stand-in for sensor_msgs.msg.PointCloud2 (tests, offline replay of recorded buffers)."""
    width: int
    height: int = 1
    fields: List[PointField] = field(default_factory=list)
    is_bigendian: bool = False
    point_step: int = 16
    row_step: int = 0
    data: bytes = b""
    is_dense: bool = True


def make_xyzi_cloud(points: np.ndarray) -> PointCloud2:
    """(N,4) float32 x,y,z,intensity -> a Velodyne-style PointCloud2 (16-byte points), for tests and replay."""
    p = np.ascontiguousarray(points, dtype=np.float32)
    assert p.ndim == 2 and p.shape[1] == 4
    return PointCloud2(width=p.shape[0], height=1,
                       fields=[PointField("x", 0), PointField("y", 4), PointField("z", 8), PointField("intensity", 12)],
                       point_step=16, row_step=16 * p.shape[0], data=p.tobytes())


def extract_xyz(msg) -> np.ndarray:
    """scan_context_node.py:249-276 -- the x, y, z fields of a PointCloud2 as an (N,3) float32 array."""
    by_name = {f.name: f for f in msg.fields}
    for k in ("x", "y", "z"):
        if k not in by_name:
            raise ValueError(f"PointCloud2 has no field {k!r}")
    n_rows, n_cols = int(msg.height), int(msg.width)
    if n_rows * n_cols == 0:
        return np.empty((0, 3), dtype=np.float32)
    buf = np.frombuffer(bytes(msg.data) if not isinstance(msg.data, (bytes, bytearray, memoryview, np.ndarray)) else msg.data,
                        dtype=np.uint8)
    row_step = int(msg.row_step) or n_cols * int(msg.point_step)
    out = np.empty((n_rows * n_cols, 3), dtype=np.float32)
    for j, k in enumerate(("x", "y", "z")):
        f = by_name[k]
        dt = np.dtype(_DT[int(f.datatype)]).newbyteorder(">" if msg.is_bigendian else "<")
        col = np.ndarray((n_rows, n_cols), dtype=dt, buffer=buf, offset=int(f.offset),
                         strides=(row_step, int(msg.point_step)))
        out[:, j] = col.reshape(-1)
    return out

"""NumPy restatement of the reference's two scan-context generators -- the "reference NumPy path" that bench.py times
on the GPU box's host cores beside the GPU number (the reference itself is Python under
/root/reference and cannot travel to that box).

TEST INFRASTRUCTURE ONLY (same rules as the rest of oracle/).

* ``variant_a``  ScanContext.ptcloud2sc_vectorized   scaner_version2.py:64-115: the vectorised polar part followed by the
                 reference's own per-point Python loop (:105-108) -- that loop is 90 % of the reference's time and is
                 restated as a loop on purpose, so the timing is the reference's, not an optimised variant's.
* ``variant_b``  ScanContextComputer._scan_context   scan_context_node.py:102-157 (np.maximum.at).

``arctan2`` is pluggable: NumPy's own (the live flavour of the host: SVML on AVX-512 machines, libm elsewhere) or the
oracle's fdlibm restatement (``oracle.atan2f``), with which variant A reproduces the reference's golden grids bit for bit.
"""
from __future__ import annotations

import numpy as np


def variant_a(ptcloud, num_sector=np.int64(720), num_ring=np.int64(160), max_length=20, use_full_height_range=True,
              custom_height_offset=2.0, arctan2=np.arctan2):
    gap_ring = max_length / num_ring                       # :70-71
    gap_sector = 360 / num_sector
    x = ptcloud[:, 0]
    y = ptcloud[:, 1]
    z = ptcloud[:, 2] if use_full_height_range else ptcloud[:, 2] + custom_height_offset      # :77-82
    x = np.where(x == 0, 0.001, x)                          # :85-86
    y = np.where(y == 0, 0.001, y)
    theta = np.rad2deg(arctan2(y, x)) % 360                 # :89
    r = np.sqrt(x * x + y * y)                              # :90
    ring_idx = np.minimum(r // gap_ring, num_ring - 1).astype(np.int32)       # :93
    sector_idx = (theta // gap_sector).astype(np.int32)                        # :94
    sc = np.full((num_ring, num_sector), -np.inf) if use_full_height_range else np.zeros((num_ring, num_sector))
    for i in range(len(r)):                                 # :105-108 -- the reference's per-point loop
        if z[i] > sc[ring_idx[i], sector_idx[i]]:
            sc[ring_idx[i], sector_idx[i]] = z[i]
    if use_full_height_range:
        sc = np.where(sc == -np.inf, 0, sc)                 # :111-113
    return sc


def variant_b(pts_xyz, max_length=80.0, sector_res=720, ring_res=160, use_full_height_range=False,
              custom_height_offset=2.0, arctan2=np.arctan2):
    if pts_xyz.size == 0:                                   # :107-108
        return np.zeros((ring_res, sector_res), dtype=np.float32)
    x, y, z = pts_xyz[:, 0], pts_xyz[:, 1], pts_xyz[:, 2]
    m = np.isfinite(x) & np.isfinite(y) & np.isfinite(z)    # :115-117
    if not np.all(m):
        x, y, z = x[m], y[m], z[m]
    if use_full_height_range:                               # :120-127
        z_used, init_val = z, -np.inf
    else:
        z_used, init_val = z + custom_height_offset, 0.0
    r = np.sqrt(x * x + y * y)                              # :130-131
    theta = (np.rad2deg(arctan2(y, x)) + 360.0) % 360.0
    m = r <= max_length                                     # :134-137
    if not np.any(m):
        return np.zeros((ring_res, sector_res), dtype=np.float32)
    r, theta, z_used = r[m], theta[m], z_used[m]
    gap_ring = max_length / ring_res
    gap_sector = 360.0 / sector_res
    ring_idx = np.clip(np.floor(r / gap_ring).astype(np.int32), 0, ring_res - 1)          # :142-145
    sector_idx = np.clip(np.floor(theta / gap_sector).astype(np.int32), 0, sector_res - 1)
    flat_idx = ring_idx * sector_res + sector_idx
    sc_flat = np.full(ring_res * sector_res, init_val, dtype=np.float32)                  # :150-151
    np.maximum.at(sc_flat, flat_idx, z_used.astype(np.float32))
    if use_full_height_range:
        sc_flat[~np.isfinite(sc_flat)] = 0.0                # :154-155
    return sc_flat.reshape((ring_res, sector_res))


def _pool_job(args):
    pts, kw = args
    return variant_a(pts, **kw)


def variant_a_pool(scans, processes, **kw):
    """the same loop under multiprocessing.Pool(all host cores)."""
    import multiprocessing as mp
    with mp.get_context("fork").Pool(processes) as pool:
        return pool.map(_pool_job, [(s, kw) for s in scans])

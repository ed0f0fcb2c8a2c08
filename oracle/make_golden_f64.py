"""Generate tests/golden/ref_live_f64.npz: the UNMODIFIED reference run on FLOAT64 clouds (what Open3D's
voxel_down_sample hands it: scan_context_node.py:93-100 -> :161-168, scaner_version2.py:126-133).

    python -m oracle.make_golden_f64

Every case stores the float64 points, the float64 np.arctan2 values THIS host produced for them (NumPy's SVML loop
on AVX-512 hosts, libm elsewhere: the flavour is host dependent and upstream pins none), the configuration and the
grid the reference returned: float64 for variant A (ptcloud2sc_vectorized), float32 for variant B (_scan_context).
Tests then check  oracle(host arctan2 plugged in) == reference grid bit-for-bit  and
engine == oracle(fdlibm arctan2) == reference grid except where a counted boundary point says otherwise.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_shim as ref  # noqa: E402


def voxel_mean(points_f32: np.ndarray, voxel: float) -> np.ndarray:
    """float64 voxel means of a float32 cloud -- the kind of values a down-sampled cloud holds (input generator
    only; the engine's own voxel filter is tested separately)."""
    p = points_f32.astype(np.float64)
    key = np.floor((p - (p.min(axis=0) - voxel * 0.5)) / voxel).astype(np.int64)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    acc = np.zeros((cnt.size, 3))
    np.add.at(acc, inv.reshape(-1), p)
    return acc / cnt[:, None]


def synth_f64(rng, n, scale=25.0):
    p = rng.standard_normal((n, 3)) * [scale, scale, 1.5] - [0, 0, 0.5]
    k = max(8, n // 200)
    idx = rng.choice(n, 6 * k, replace=False)
    p[idx[:k], 0] = 0.0                                  # x == 0 (variant A: -> 0.001)
    p[idx[k:2 * k], 1] = 0.0                             # y == 0: theta exactly 0 or 180
    p[idx[2 * k:3 * k], 0] = p[idx[2 * k:3 * k], 1]      # exact diagonals: theta = 45 / 225 degrees, a sector boundary
    p[idx[3 * k:4 * k], 0] = -p[idx[3 * k:4 * k], 1]     # 135 / 315 degrees
    p[idx[4 * k:5 * k], 2] = 0.0
    p[idx[5 * k:5 * k + 6]] = [[20.0, 0.0, 1.0], [0.125, 0.0, 0.5], [12.0, 16.0, 0.25], [-16.0, 12.0, 0.75],
                                [0.0, -7.0, 0.1], [-3.0, 0.0, -0.2]]
    return p


def main():
    assert ref.available(), "needs /root/reference"
    rng = np.random.default_rng(20261018)
    scans = np.load(os.path.join(OUT, "scans.npz"))
    real = voxel_mean(np.ascontiguousarray(scans["s002095"].T[:, :3]), 0.2)        # ~ what the live node bins
    live = {}
    a_cases = [
        # name, points, num_ring, num_sector, max_length, full_height, offset
        ("a64_default", synth_f64(rng, 6000), np.int64(160), np.int64(720), 20, True, 2.0),
        ("a64_offset", synth_f64(rng, 6000), np.int64(160), np.int64(720), 20, False, 2.0),
        ("a64_rmax80", synth_f64(rng, 6000), np.int64(160), np.int64(720), 80, True, 2.0),
        ("a64_nondyadic", synth_f64(rng, 6000), np.int64(150), np.int64(700), 33, True, 2.0),
        ("a64_pyint", synth_f64(rng, 4000), 150, 700, 33, False, 0.1),
        ("a64_real_ds", real, np.int64(160), np.int64(720), 20, True, 2.0),
        ("a64_empty", np.zeros((0, 3)), np.int64(160), np.int64(720), 20, True, 2.0),
    ]
    for name, pts, R, S, ml, full, off in a_cases:
        x = np.where(pts[:, 0] == 0, 0.001, pts[:, 0])
        y = np.where(pts[:, 1] == 0, 0.001, pts[:, 1])
        g = np.asarray(ref.variant_a(pts, num_sector=S, num_ring=R, max_length=ml, full_height=full, offset=off))
        assert g.dtype == np.float64
        live[f"{name}__pts"] = pts
        live[f"{name}__atan2"] = np.arctan2(y, x)
        live[f"{name}__cfg"] = np.array([int(R), int(S), ml, int(full), off], np.float64)
        live[f"{name}__grid"] = g
    b_cases = [
        ("b64_default", synth_f64(rng, 6000, 40.0), 160, 720, 80.0, False, 2.0),
        ("b64_full", synth_f64(rng, 6000, 40.0), 160, 720, 80.0, True, 2.0),
        ("b64_rmax20", synth_f64(rng, 6000), 160, 720, 20.0, True, 2.0),
        ("b64_nondyadic", synth_f64(rng, 6000), 77, 1000, 33.3, False, 0.75),
        ("b64_real_ds", real, 160, 720, 80.0, False, 2.0),              # the live node's defaults on a down-sampled scan
        ("b64_real_ds_full", real, 160, 720, 80.0, True, 2.0),           # node parameter default (use_full_height_range=1)
    ]
    for name, pts, R, S, ml, full, off in b_cases:
        pts = pts.copy()
        if "real" not in name:
            k = rng.choice(len(pts), 16, replace=False)     # masked out by the finite test
            pts[k[:4], 0] = np.nan; pts[k[4:8], 1] = np.inf; pts[k[8:12], 2] = -np.inf
            pts[k[12]] = [0.0, 0.0, 1.0]                    # the origin
            pts[k[13]] = [3.0, 4.0, 1e39]                   # float32 overflow of the height -> +inf -> 0 (full) / inf (offset)
            pts[k[14]] = [-3.0, 4.0, -1e39]
            pts[k[15]] = [ml, 0.0, 0.3]                     # r == max_length exactly: kept, ring clipped to R-1
        with np.errstate(invalid="ignore", over="ignore"):
            live[f"{name}__atan2"] = np.arctan2(pts[:, 1], pts[:, 0])
            g = np.asarray(ref.variant_b(pts.copy(), max_length=ml, sector_res=S, ring_res=R, full_height=full, offset=off))
        assert g.dtype == np.float32
        live[f"{name}__pts"] = pts
        live[f"{name}__cfg"] = np.array([R, S, ml, int(full), off], np.float64)
        live[f"{name}__grid"] = g
    live["a_case_names"] = np.array([c[0] for c in a_cases])
    live["b_case_names"] = np.array([c[0] for c in b_cases])
    np.savez_compressed(os.path.join(OUT, "ref_live_f64.npz"), **live)
    print("wrote ref_live_f64.npz", os.path.getsize(os.path.join(OUT, "ref_live_f64.npz")) / 1e6, "MB;",
          "real down-sampled cloud:", real.shape)


if __name__ == "__main__":
    main()

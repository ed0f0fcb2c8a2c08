"""Generate tests/golden/* by running the UNMODIFIED reference (imported by path through
oracle/ref_shim.py) in the build container.  The reference is Python and cannot travel to the GPU
box, so its inputs/outputs are committed as fixtures together with this script.

    python -m oracle.make_golden

Fixtures
  scans.npz            the four KITTI scans of the reference's test_data/ (float32 [x,y,z,i], stored
                       column-major for compression) -- reference *data*, not source
  sc_golden.npz        the reference's own golden grids test_data/sc_*.npy
  pedestrians_*.json   the reference's own golden label files (52 boxes)
  ref_live.npz         outputs of the live reference on seeded inputs: variant A/B grids (with the
                       np.arctan2 values this host produced, because float32 arctan2 is host dependent:
                       SVML on AVX-512 hosts, libm elsewhere), rasteriser, reconstructor, to_tensor01
  tinycnn_logits.npz   TinyCNN(tinycnn_best.pth) fp32 CPU logits on to_tensor01(sc_000840)
  tinycnn_best.pth     the reference's TinyCNN state_dict (data, 583 KB), used by the config-4 tests and bench
"""
from __future__ import annotations

import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_shim as ref  # noqa: E402

SCANS = ["000840", "002073", "002095", "002160"]


def synth_points(rng, n, scale=25.0, special=True):
    """Seeded synthetic cloud with the edge cases SURVEY 8(a) lists (excluding the ones on which
    the reference raises)."""
    x = (rng.standard_normal(n) * scale).astype(np.float32)
    y = (rng.standard_normal(n) * scale).astype(np.float32)
    z = (rng.standard_normal(n) * 1.5 - 0.5).astype(np.float32)
    # mm quantisation like KITTI
    x, y, z = (np.round(v * 1000) / 1000 for v in (x, y, z))
    x, y, z = x.astype(np.float32), y.astype(np.float32), z.astype(np.float32)
    if special:
        k = max(8, n // 200)
        idx = rng.choice(n, 6 * k, replace=False)
        x[idx[:k]] = 0.0                      # x == 0 -> 0.001 substitution
        y[idx[k:2 * k]] = 0.0                 # y == 0 (positive x: theta 0.057 deg; negative x: ~180)
        x[idx[2 * k:3 * k]] = y[idx[2 * k:3 * k]]          # exact 45 / 225 degree diagonals
        z[idx[3 * k:4 * k]] = 0.0             # cells whose max is exactly 0
        z[idx[4 * k:5 * k]] = -0.0
        # exactly r == max_length and ring boundaries
        x[idx[5 * k:5 * k + 4]] = [20.0, 0.125, 12.0, 16.0]
        y[idx[5 * k:5 * k + 4]] = [0.0, 0.0, 16.0, 12.0]
    return np.stack([x, y, z], axis=1)


def main():
    assert ref.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    td = os.path.join(ref.REF_ROOT, "test_data")

    scans = {f"s{s}": np.ascontiguousarray(np.fromfile(os.path.join(td, f"{s}.bin"), dtype=np.float32).reshape(-1, 4).T)
             for s in SCANS}
    np.savez_compressed(os.path.join(OUT, "scans.npz"), **scans)
    np.savez_compressed(os.path.join(OUT, "sc_golden.npz"),
                        **{f"sc{s}": np.load(os.path.join(td, f"sc_{s}.npy")) for s in SCANS})
    for s in SCANS:
        shutil.copyfile(os.path.join(td, f"pedestrians_{s}.json"), os.path.join(OUT, f"pedestrians_{s}.json"))

    rng = np.random.default_rng(20251018)
    live = {}
    # ---- variant A ------------------------------------------------------------------------------
    a_cases = [
        # name, n, num_ring, num_sector, max_length, full_height, offset, python_int_res
        ("a_default", 6000, np.int64(160), np.int64(720), 20, True, 2.0),
        ("a_offset", 6000, np.int64(160), np.int64(720), 20, False, 2.0),
        ("a_rmax80", 6000, np.int64(160), np.int64(720), 80, True, 2.0),
        ("a_coarse", 4000, np.int64(20), np.int64(90), 20, True, 2.0),
        ("a_nondyadic", 6000, np.int64(150), np.int64(700), 33, True, 2.0),
        ("a_nondyadic_pyint", 6000, 150, 700, 33, False, 1.5),
        ("a_empty", 0, np.int64(160), np.int64(720), 20, True, 2.0),
        ("a_one", 1, np.int64(160), np.int64(720), 20, True, 2.0),
    ]
    for name, n, R, S, ml, full, off in a_cases:
        pts = synth_points(rng, n, special=n > 100) if n else np.zeros((0, 3), np.float32)
        x = np.where(pts[:, 0] == 0, 0.001, pts[:, 0])
        y = np.where(pts[:, 1] == 0, 0.001, pts[:, 1])
        live[f"{name}__pts"] = pts
        live[f"{name}__atan2"] = np.arctan2(y, x).astype(np.float32)
        live[f"{name}__cfg"] = np.array([int(R), int(S), ml, int(full), off, int(not isinstance(R, np.integer))], np.float64)
        live[f"{name}__grid"] = np.asarray(ref.variant_a(pts, num_sector=S, num_ring=R, max_length=ml,
                                                         full_height=full, offset=off), dtype=np.float64)
    # real-scan prefix with non-default parameters
    real = scans["s002095"].T[:30000, :3].copy()
    for name, R, S, ml, full in [("a_real_r40", np.int64(40), np.int64(180), 20, True),
                                 ("a_real_offset80", np.int64(160), np.int64(720), 80, False)]:
        x = np.where(real[:, 0] == 0, 0.001, real[:, 0]); y = np.where(real[:, 1] == 0, 0.001, real[:, 1])
        live[f"{name}__pts"] = real
        live[f"{name}__atan2"] = np.arctan2(y, x).astype(np.float32)
        live[f"{name}__cfg"] = np.array([int(R), int(S), ml, int(full), 2.0, 0], np.float64)
        live[f"{name}__grid"] = np.asarray(ref.variant_a(real, num_sector=S, num_ring=R, max_length=ml, full_height=full))
    # ---- variant B ------------------------------------------------------------------------------
    b_cases = [
        ("b_default", 6000, 160, 720, 80.0, False, 2.0),
        ("b_full", 6000, 160, 720, 80.0, True, 2.0),
        ("b_rmax20", 6000, 160, 720, 20.0, True, 2.0),
        ("b_nondyadic", 6000, 77, 1000, 33.3, False, 0.75),
        ("b_empty", 0, 160, 720, 80.0, False, 2.0),
    ]
    for name, n, R, S, ml, full, off in b_cases:
        pts = synth_points(rng, n) if n else np.zeros((0, 3), np.float32)
        if n:
            k = rng.choice(n, 12, replace=False)     # variant B masks non-finite points out
            pts[k[:4], 0] = np.nan; pts[k[4:8], 1] = np.inf; pts[k[8:], 2] = -np.inf
            pts[k[0], 0] = 0.0; pts[k[0], 1] = 0.0; pts[k[0], 2] = 1.0     # the origin: theta = 0, r = 0
        live[f"{name}__pts"] = pts
        with np.errstate(invalid="ignore"):
            live[f"{name}__atan2"] = np.arctan2(pts[:, 1], pts[:, 0]).astype(np.float32)
            live[f"{name}__grid"] = np.asarray(ref.variant_b(pts.copy(), max_length=ml, sector_res=S, ring_res=R,
                                                             full_height=full, offset=off), dtype=np.float32)
        live[f"{name}__cfg"] = np.array([R, S, ml, int(full), off, 0], np.float64)
    live["a_case_names"] = np.array([c[0] for c in a_cases] + ["a_real_r40", "a_real_offset80"])
    live["b_case_names"] = np.array([c[0] for c in b_cases])

    # ---- rasteriser: random boxes through the live reference ----------------------------------------
    nb = 48
    boxes = np.stack([rng.uniform(-24, 24, nb), rng.uniform(-24, 24, nb), rng.uniform(0.3, 4.5, nb),
                      rng.uniform(0.3, 2.5, nb), rng.uniform(-np.pi, np.pi, nb)], axis=1)
    boxes[0] = [0.0, 0.0, 1.0, 1.0, 0.0]            # box over the sensor origin
    boxes[1] = [19.9, 0.1, 2.0, 2.0, 0.3]           # straddles max range
    boxes[2] = [5.0, 0.0, 0.8, 0.8, 0.0]            # straddles the 0/360 seam
    boxes[3] = [30.0, 30.0, 1.0, 1.0, 1.0]          # beyond range: empty
    offs, cells = [0], []
    for b in boxes:
        c = ref.cells_covered_by_box(x=float(b[0]), y=float(b[1]), l=float(b[2]), w=float(b[3]), yaw_lidar=float(b[4]))
        cells.append(np.asarray(c, dtype=np.int64).reshape(-1, 2))
        offs.append(offs[-1] + cells[-1].shape[0])
    live["raster__boxes"] = boxes
    live["raster__offsets"] = np.array(offs, np.int64)
    live["raster__cells"] = np.concatenate(cells, axis=0)
    # a non-default grid
    c = ref.cells_covered_by_box(x=7.5, y=-3.25, l=3.9, w=1.6, yaw_lidar=0.77, num_ring=80, num_sector=360, max_length=40.0)
    live["raster_alt__cells"] = np.asarray(c, np.int64)
    # centre indices (float32 label path and float64 calib path)
    cx = rng.uniform(-25, 25, 64); cy = rng.uniform(-25, 25, 64)
    live["centre__xy"] = np.stack([cx, cy], 1)
    live["centre__f64"] = np.array([ref.pedestrian_to_scan_context_coords(float(a), float(b)) for a, b in zip(cx, cy)], np.int64)
    live["centre__f32"] = np.array([ref.pedestrian_to_scan_context_coords(np.float32(a), np.float32(b)) for a, b in zip(cx, cy)], np.int64)

    # ---- reconstructor on the golden grids ---------------------------------------------------------------
    for s in ("000840", "002160"):
        g = np.load(os.path.join(td, f"sc_{s}.npy"))
        P, Cc, ids = ref.reconstruct(g, full_height=True)
        live[f"recon_{s}__pts"] = P
        live[f"recon_{s}__ids"] = ids
        assert np.all(Cc == 0.7)
    g = np.load(os.path.join(td, "sc_002095.npy"))
    P, Cc, ids = ref.reconstruct(g, full_height=False, offset=2.0)
    live["recon_002095_offset__pts"] = P
    live["recon_002095_offset__ids"] = ids
    # covered-cell gather + isin mask (reconstruct_pointcloud.py:104-130,153-164)
    rec = ref.reconstruct_pointcloud().ScanContextReconstructor()
    import contextlib, io
    g = np.load(os.path.join(td, "sc_000840.npy"))
    with contextlib.redirect_stdout(io.StringIO()):
        cov = rec.load_covered_cells_from_json(os.path.join(td, "pedestrians_000840.json"))
        gp, gc = rec.add_covered_cells_to_pointcloud(g, cov)
    P, Cc, ids = ref.reconstruct(g)
    cids = np.array(cov, dtype=int).reshape(-1, 2)
    cids[:, 0] = np.clip(cids[:, 0], 0, 159); cids[:, 1] = np.clip(cids[:, 1], 0, 719)
    live["covered_000840__gather_pts"] = gp
    live["covered_000840__mask"] = np.isin(ids, (cids[:, 0] * 720 + cids[:, 1]).astype(np.int32))

    # ---- to_tensor01 ----------------------------------------------------------------------------------------
    tiny = ref.tiny_sc_ped()
    live["norm01_002073"] = tiny.to_tensor01(np.load(os.path.join(td, "sc_002073.npy"))).numpy()[0]
    m = tiny.build_mask_from_json(160, 720, os.path.join(td, "pedestrians_002160.json"))
    live["mask_002160"] = m
    np.savez_compressed(os.path.join(OUT, "ref_live.npz"), **live)

    # ---- TinyCNN logits (config 4) ------------------------------------------------------------------------
    import torch
    torch.set_num_threads(8)
    model = tiny.TinyCNN()
    sd = torch.load(os.path.join(ref.REF_ROOT, "distro_model", "tinycnn_best.pth"), map_location="cpu", weights_only=True)
    model.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        x = tiny.to_tensor01(np.load(os.path.join(td, "sc_000840.npy"))).unsqueeze(0)
        logits = model(x)[0, 0].numpy()
    np.savez_compressed(os.path.join(OUT, "tinycnn_logits.npz"), logits=logits)
    # the weights are reference *data* (like the scans): config 4 needs them on the GPU box
    shutil.copyfile(os.path.join(ref.REF_ROOT, "distro_model", "tinycnn_best.pth"), os.path.join(OUT, "tinycnn_best.pth"))
    info = {"numpy": np.__version__, "torch": torch.__version__,
            "note": "float32 arctan2 flavour of this host is recorded per case in ref_live.npz (*__atan2)"}
    with open(os.path.join(OUT, "README.json"), "w") as f:
        json.dump(info, f, indent=1)
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("wrote", OUT, f"{tot/1e6:.1f} MB")


if __name__ == "__main__":
    main()

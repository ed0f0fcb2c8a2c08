"""The Python mirrors of the reference classes behave like the reference: same signatures, return
dtypes, error behaviour -- these read like the tests the reference never had."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, SCAN_IDS

pytestmark = pytest.mark.gpu


@pytest.fixture()
def bin_dir(tmp_path, golden_scans):
    for sid in SCAN_IDS[:2]:
        golden_scans[sid].tofile(tmp_path / f"{sid}.bin")
    return str(tmp_path) + os.sep


def test_scan_context_offline_entry_point(bin_dir, golden_grids):
    """scaner_version2.py __main__ loop body: ScanContext(bin_dir, name).scancontexts[0] -> np.save f32."""
    from distropointclouds_b200.scan_context import ScanContext
    for sid in SCAN_IDS[:2]:
        sc = ScanContext(bin_dir, f"{sid}.bin")
        g = sc.scancontexts[0]
        assert g.dtype == np.float64 and g.shape == (160, 720)
        assert np.array_equal(g.astype(np.float32), golden_grids[sid])
        assert ScanContext.last_stats["n_points"] > 100000


def test_scan_context_kitti_dataset_entry_point(golden_scans, golden_grids):
    """kitti_dataset.py:466-475: ScanContext(lidarData[:, :3]).scancontexts[0]."""
    from distropointclouds_b200.scan_context import ScanContext
    lidar = golden_scans["002160"]
    sc = ScanContext(lidar[:, :3])
    assert np.array_equal(sc.scancontexts[0].astype(np.float32), golden_grids["002160"])


def test_scan_context_raises_index_error_like_reference():
    from distropointclouds_b200.scan_context import ScanContext
    obj = ScanContext.__new__(ScanContext)
    ok = np.array([[1.0, 2.0, 0.5], [3.0, -1.0, 0.2]], np.float32)
    g = obj.ptcloud2sc_vectorized(ok, np.int64(720), np.int64(160), 20)
    assert g.dtype == np.float64 and np.count_nonzero(g) == 2
    for bad in ([[np.nan, 1.0, 0.0]], [[np.inf, 1.0, 0.0]], [[50.0, -1e-7, 0.0]], [[3e38, 3e38, 0.0]]):
        with pytest.raises(IndexError):
            obj.ptcloud2sc_vectorized(np.array(bad, np.float32), np.int64(720), np.int64(160), 20)
    ScanContext.strict = False
    try:
        g = obj.ptcloud2sc_vectorized(np.array([[np.nan, 1.0, 0.0], [1.0, 1.0, 7.0]], np.float32), np.int64(720), np.int64(160), 20)
        assert np.count_nonzero(g) == 1 and ScanContext.last_stats["n_nonfinite"] == 1
    finally:
        ScanContext.strict = True
    # the reference's alien-format sample makes it raise too; empty clouds work
    g = obj.ptcloud2sc_vectorized(np.zeros((0, 3), np.float32), np.int64(720), np.int64(160), 20)
    assert g.shape == (160, 720) and not g.any()


def test_scan_context_offset_mode_and_statics(golden_scans, orc):
    from distropointclouds_b200.scan_context import ScanContext
    pts = golden_scans["000840"][:50000, :3]
    old = (ScanContext.use_full_height_range, ScanContext.custom_height_offset, ScanContext.max_length)
    try:
        ScanContext.use_full_height_range, ScanContext.custom_height_offset, ScanContext.max_length = 0, 2.0, 80
        g = ScanContext(np.ascontiguousarray(pts)).scancontexts[0]
        o, _ = orc.sc_build(np.ascontiguousarray(pts), orc.Params(max_length=80.0, full_height=False))
        assert np.array_equal(g.astype(np.float32), o) and g.min() >= 0
    finally:
        ScanContext.use_full_height_range, ScanContext.custom_height_offset, ScanContext.max_length = old


def test_live_computer(tmp_path, golden_scans, orc):
    from distropointclouds_b200.scan_context import ScanContextComputer
    comp = ScanContextComputer(output_dir=str(tmp_path / "o"), dl_images_dir=str(tmp_path / "i"))
    pts = np.ascontiguousarray(golden_scans["002073"][:, :3])
    sc, _, npy = comp.compute_and_save(pts, "000001")
    # compute_and_save = voxel down-sample (float64 voxel means) -> float64 variant B  (scan_context_node.py:159-168)
    ds, _ = orc.voxel_downsample(pts, comp.voxel_size)
    o, _ = orc.sc_build_f64(ds, orc.Params(max_length=80.0, variant=orc.VARIANT_B, full_height=False))
    assert sc.dtype == np.float32 and np.array_equal(sc, o) and np.array_equal(np.load(npy), o)
    # the bare generator on the raw float32 cloud is the float32 chain
    o32, _ = orc.sc_build(pts, orc.Params(max_length=80.0, variant=orc.VARIANT_B, full_height=False))
    assert np.array_equal(comp._scan_context(pts), o32)
    assert not comp._scan_context(np.zeros((0, 3), np.float32)).any()


def test_label_projection_reproduces_golden_json(golden_peds):
    """find_pedestrians_in_scan_context on the float32 labels the JSON's first half was made from."""
    from distropointclouds_b200.scan_context import ScanContext
    for sid in SCAN_IDS:
        peds = golden_peds[sid]
        half = peds[: len(peds) // 2]
        labels = np.array([[p["class"], p["x"], p["y"], p["z"], p["h"], p["w"], p["l"], p["yaw"]] for p in half], np.float32)
        labels = np.concatenate([labels, np.array([[0, 5, 5, 0, 1, 1, 1, 0]], np.float32)])      # a non-pedestrian row
        got = ScanContext.find_pedestrians_in_scan_context(labels)
        assert len(got) == len(half)
        for a, b in zip(got, half):
            assert a["covered_cells"] == b["covered_cells"]
            assert (a["center_ring_idx"], a["center_sector_idx"]) == (b["center_ring_idx"], b["center_sector_idx"])
            for k in ("x", "y", "z", "h", "w", "l", "yaw", "distance"):
                assert a[k] == b[k], k
            assert abs(a["angle"] - b["angle"]) <= 1e-4
        c = ScanContext.cells_covered_by_box(x=half[0]["x"], y=half[0]["y"], l=half[0]["l"], w=half[0]["w"], yaw_lidar=half[0]["yaw"])
        assert c.dtype == np.int64 and c.tolist() == half[0]["covered_cells"]


def test_reconstructor_class(tmp_path, golden_grids, golden_peds, ref_live):
    from distropointclouds_b200.reconstruct import ScanContextReconstructor
    from distropointclouds_b200.labels import build_mask_from_json
    np.save(tmp_path / "sc_000840.npy", golden_grids["000840"])
    rec = ScanContextReconstructor(str(tmp_path))
    rec.pedestrian_json_dir = GOLDEN
    pts, colors, ids = rec.scan_context_to_pointcloud(golden_grids["000840"])
    assert np.array_equal(pts, ref_live["recon_000840__pts"]) and np.array_equal(ids, ref_live["recon_000840__ids"])
    assert colors.shape == pts.shape and (colors == 0.7).all()
    res = rec.process_all_files()
    p, c, i, red = res["sc_000840.npy"]
    assert red == int(ref_live["covered_000840__mask"].sum()) and (c[ref_live["covered_000840__mask"]] == [1, 0, 0]).all()
    cov = rec.load_covered_cells_from_json(os.path.join(GOLDEN, "pedestrians_000840.json"))
    gp, gc = rec.add_covered_cells_to_pointcloud(golden_grids["000840"], cov)
    assert np.array_equal(gp, ref_live["covered_000840__gather_pts"]) and (gc == [1.0, 0.0, 0.0]).all()
    m = build_mask_from_json(160, 720, os.path.join(GOLDEN, "pedestrians_002160.json"))
    assert m.dtype == np.uint8 and np.array_equal(m, ref_live["mask_002160"])


@pytest.mark.gpu
def test_add_covered_cells_offset_mode(golden_grids, golden_peds):
    """reconstruct_pointcloud.py:104-130 with use_full_height_range = 0: only cells with z > 0 are kept and the height
    offset is removed in the grid's dtype (float32), against the reference's own NumPy expressions; out-of-grid cell
    indices are clipped (:109-110)."""
    from distropointclouds_b200.reconstruct import ScanContextReconstructor
    sc = (golden_grids["002095"] + np.float32(2.0)) * (golden_grids["002095"] != 0)     # an offset-mode grid (z + 2, empty = 0)
    sc = sc.astype(np.float32)
    cov = [c for p in golden_peds["002095"] for c in p["covered_cells"]] + [[-3, 10], [200, 800], [159, 719], [0, 0]]

    class Rec(ScanContextReconstructor):
        use_full_height_range = 0
        custom_height_offset = 2.0
    gp, gc = Rec().add_covered_cells_to_pointcloud(sc, cov)
    idx = np.array(cov, dtype=int)
    idx[:, 0] = np.clip(idx[:, 0], 0, 159); idx[:, 1] = np.clip(idx[:, 1], 0, 719)
    r = (idx[:, 0] + 0.5) * (20 / 160)
    th = np.deg2rad((idx[:, 1] + 0.5) * (360.0 / 720))
    x, y, z = r * np.cos(th), r * np.sin(th), sc[idx[:, 0], idx[:, 1]]
    valid = z > 0
    ref = np.column_stack((x[valid], y[valid], z[valid] - 2.0))
    assert gp.dtype == np.float64 and gp.shape == ref.shape and np.array_equal(gp, ref)
    assert gc.shape == ref.shape and np.all(gc == [1.0, 0.0, 0.0])
    assert 0 < valid.sum() < len(cov)

"""K7 voxel-grid down-sampling and the live node's path around it.

Open3D -- the reference's dependency for this step (scan_context_node.py:93-100, scaner_version2.py:126-133) -- is not
installed and un-vendored: PARITY UNPINNED.  What is checked: the engine against the oracle's restatement of Open3D's
published algorithm (bit-exact, including the order-dependent float64 sums), the restatement against an independent
NumPy formulation, and the whole live chain PointCloud2 -> xyz -> voxel filter -> float64 scan context."""
import os
import tempfile

import numpy as np
import pytest


def numpy_voxel(points, voxel):
    """Independent formulation: np.unique over the voxel indices + np.add.at (which adds in index order)."""
    p = np.asarray(points, dtype=np.float64)[:, :3]
    p = p[np.isfinite(p).all(axis=1)]
    if len(p) == 0:
        return p
    idx = np.floor((p - (p.min(axis=0) - voxel * 0.5)) / voxel).astype(np.int64)
    _, inv, cnt = np.unique(idx, axis=0, return_inverse=True, return_counts=True)
    acc = np.zeros((cnt.size, 3))
    np.add.at(acc, inv.reshape(-1), p)
    return acc / cnt[:, None].astype(np.float64)


def test_oracle_voxel_filter_equals_numpy_formulation(orc, golden_scans):
    rng = np.random.default_rng(2)
    clouds = [golden_scans["002095"][:40000, :3], rng.standard_normal((5000, 3)) * [30, 30, 2],
              (rng.standard_normal((3000, 3)) * 1e-3).astype(np.float32), np.array([[1.0, 2.0, 3.0]])]
    for pts in clouds:
        for voxel in (0.2, 0.05, 1.7):
            got, dropped = orc.voxel_downsample(pts, voxel)
            ref = numpy_voxel(pts, voxel)
            assert dropped == 0 and got.shape == ref.shape
            assert np.array_equal(got, ref)                                   # same voxels, same order, same bits


@pytest.mark.gpu
def test_k7_equals_oracle_bit_for_bit(orc, golden_scans):
    from distropointclouds_b200.engine import Engine
    rng = np.random.default_rng(3)
    bad = rng.standard_normal((4000, 3)) * [20, 20, 2]
    bad[[5, 17, 900], 0] = np.nan
    bad[[6, 3999], 2] = np.inf
    clouds = [np.ascontiguousarray(golden_scans["000840"]),                  # (N,4) float32: 16-byte rows
              np.ascontiguousarray(golden_scans["002160"][:, :3]),            # (N,3) float32
              rng.standard_normal((20000, 3)) * [40, 40, 3],                  # float64
              bad, np.array([[1.0, 2.0, 3.0]]), np.zeros((7, 3)),             # non-finite points; one point; one voxel
              (rng.standard_normal((3000, 3)) * 300).astype(np.float32)]
    with Engine(0) as e:
        for pts in clouds:
            for voxel in (0.2, 0.5, 0.031):
                got, dropped = e.voxel_downsample_host(pts, voxel)
                ref, rdrop = orc.voxel_downsample(pts, voxel)
                assert dropped == rdrop and got.shape == ref.shape, (pts.shape, voxel)
                assert np.array_equal(got, ref), (pts.shape, voxel)
        out, _ = e.voxel_downsample_host(np.zeros((0, 3), np.float32), 0.2)
        assert out.shape == (0, 3)
        from distropointclouds_b200._ffi import DscError
        with pytest.raises(DscError):                                          # Open3D: "voxel_size is too small"
            e.voxel_downsample_host(np.array([[0.0, 0, 0], [1e6, 0, 0]]), 1e-3)
        with pytest.raises(DscError):
            e.voxel_downsample_host(np.ones((4, 3)), 0.0)


@pytest.mark.gpu
def test_live_node_chain_pointcloud2_to_scan_context(orc, golden_scans):
    """ScanContextNode._cb's numeric chain (scan_context_node.py:278-291): _extract_xyz -> compute_and_save
    (_downsample_voxel -> float64 _scan_context -> np.save) on the engine, against the oracle chain."""
    from distropointclouds_b200 import ros
    from distropointclouds_b200.scan_context import ScanContextComputer
    scan = golden_scans["002073"]
    msg = ros.make_xyzi_cloud(scan)
    xyz = ros.extract_xyz(msg)
    assert xyz.dtype == np.float32 and np.array_equal(xyz, scan[:, :3])
    with tempfile.TemporaryDirectory() as d:
        for full in (True, False):
            comp = ScanContextComputer(voxel_size=0.2, max_length=80.0, use_full_height_range=full,
                                       output_dir=os.path.join(d, "o"), dl_images_dir=os.path.join(d, "i"))
            sc, _, npy = comp.compute_and_save(xyz, "velodyne_1.000000001")
            ds, _ = orc.voxel_downsample(xyz, 0.2)
            ref, _ = orc.sc_build_f64(ds, orc.Params(160, 720, 80.0, 1, full, 2.0))
            assert sc.dtype == np.float32 and np.array_equal(sc, ref)
            assert np.array_equal(np.load(npy), ref)
            assert 5000 < len(ds) < len(xyz) // 2


@pytest.mark.gpu
def test_offline_generator_with_downsampling(orc, golden_scans):
    """scaner_version2.py:126-133: use_downsampling=1 -> voxel filter -> float64 variant A."""
    from distropointclouds_b200.scan_context import ScanContext
    pts = golden_scans["002095"][:, :3]

    class SC(ScanContext):
        use_downsampling = 1
        downcell_size = 0.2
    sc = SC(np.ascontiguousarray(pts)).scancontexts[0]
    ds, _ = orc.voxel_downsample(pts, 0.2)
    ref, _ = orc.sc_build_f64(ds, orc.Params())
    assert sc.dtype == np.float64 and np.array_equal(sc, ref)


@pytest.mark.gpu
def test_compat_patches_the_voxel_step_too(orc, golden_scans):
    import types
    from distropointclouds_b200 import compat
    node = types.ModuleType("fake_scan_context_node_voxel")

    class ScanContextComputer:
        def __init__(self):
            self.voxel_size, self.max_length, self.sector_res, self.ring_res = 0.2, 80.0, 720, 160
            self.use_full_height_range, self.custom_height_offset = True, 2.0

        @staticmethod
        def _downsample_voxel(pts_xyz, voxel):
            raise AssertionError("unpatched (would need open3d)")

        def _scan_context(self, pts_xyz):
            raise AssertionError("unpatched")

        def compute(self, pts_xyz):                        # the reference's compute_and_save :161-168
            return self._scan_context(self._downsample_voxel(pts_xyz, self.voxel_size))
    ScanContextComputer.__module__ = node.__name__
    node.ScanContextComputer = ScanContextComputer
    done = compat.install([node])
    try:
        assert done[node.__name__] == ["ScanContextComputer._scan_context", "ScanContextComputer._downsample_voxel"]
        xyz = np.ascontiguousarray(golden_scans["000840"][:, :3])
        sc = node.ScanContextComputer().compute(xyz)
        ds, _ = orc.voxel_downsample(xyz, 0.2)
        ref, _ = orc.sc_build_f64(ds, orc.Params(160, 720, 80.0, 1, True, 2.0))
        assert np.array_equal(sc, ref)
    finally:
        compat.uninstall()

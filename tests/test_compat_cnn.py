"""compat.install() (patching the engine into reference-shaped classes) and the config-4 CNN pipeline."""
import os
import sys
import types

import numpy as np
import pytest

from conftest import GOLDEN, SCAN_IDS

REF = "/root/reference"


def _fake_reference_modules():
    """Modules shaped like the reference's (class/method/static names only, no reference code): what
    compat.install() has to find on the GPU box, where /root/reference does not exist."""
    def boom(*a, **k):
        raise AssertionError("unpatched reference method was called")

    sv2 = types.ModuleType("fake_scaner_version2")

    class ScanContext:
        use_full_height_range = 1
        custom_height_offset = 2.0
        sector_res = np.array([720])
        ring_res = np.array([160])
        max_length = 20
        ptcloud2sc_vectorized = boom
        scan_context_to_pointcloud = boom
        cells_covered_by_box = staticmethod(boom)
    ScanContext.__module__ = sv2.__name__
    sv2.ScanContext = ScanContext

    node = types.ModuleType("fake_scan_context_node")

    class ScanContextComputer:
        def __init__(self):
            self.max_length, self.sector_res, self.ring_res = 80.0, 720, 160
            self.use_full_height_range, self.custom_height_offset = False, 2.0
        _scan_context = boom
    ScanContextComputer.__module__ = node.__name__
    node.ScanContextComputer = ScanContextComputer

    rec = types.ModuleType("fake_reconstruct_pointcloud")

    class ScanContextReconstructor:
        use_full_height_range = 1
        custom_height_offset = 2.0

        def __init__(self):
            self.sector_res, self.ring_res, self.max_length = 720, 160, 20
        scan_context_to_pointcloud = boom
        add_covered_cells_to_pointcloud = boom
    ScanContextReconstructor.__module__ = rec.__name__
    rec.ScanContextReconstructor = ScanContextReconstructor

    tiny = types.ModuleType("fake_tiny_sc_ped")
    tiny.TinyCNN = type("TinyCNN", (), {})
    tiny.to_tensor01 = types.FunctionType(boom.__code__, {"AssertionError": AssertionError}, "to_tensor01")
    tiny.to_tensor01.__module__ = tiny.__name__
    return sv2, node, rec, tiny


def test_install_and_uninstall_on_reference_shaped_modules():
    from distropointclouds_b200 import compat
    mods = _fake_reference_modules()
    before = [m.__dict__.copy() for m in mods]
    orig = mods[0].ScanContext.__dict__["ptcloud2sc_vectorized"]
    done = compat.install(mods)
    try:
        assert done["fake_scaner_version2"] == ["ScanContext.ptcloud2sc_vectorized",
                                                "ScanContext.scan_context_to_pointcloud",
                                                "ScanContext.cells_covered_by_box"]
        assert done["fake_scan_context_node"] == ["ScanContextComputer._scan_context"]
        assert len(done["fake_reconstruct_pointcloud"]) == 2 and done["fake_tiny_sc_ped"] == ["to_tensor01"]
        assert mods[0].ScanContext.__dict__["ptcloud2sc_vectorized"] is not orig
    finally:
        compat.uninstall()
    assert mods[0].ScanContext.__dict__["ptcloud2sc_vectorized"] is orig
    assert isinstance(mods[0].ScanContext.__dict__["cells_covered_by_box"], staticmethod)
    assert mods[3].to_tensor01 is before[3]["to_tensor01"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_install_finds_the_real_reference_classes():
    from oracle import ref_shim
    from distropointclouds_b200 import compat
    mods = [ref_shim.scaner_version2(), ref_shim.scan_context_node(), ref_shim.reconstruct_pointcloud(),
            ref_shim.kitti_dataset(), ref_shim.tiny_sc_ped()]
    originals = [mods[0].ScanContext.ptcloud2sc_vectorized, mods[3].ScanContext.cells_covered_by_box]
    done = compat.install(mods)
    try:
        patched = sorted(n for v in done.values() for n in v)
        assert patched.count("ScanContext.ptcloud2sc_vectorized") == 2          # scaner_version2 + kitti_dataset
        assert "ScanContext.cells_covered_by_box" in patched and "ScanContextComputer._scan_context" in patched
        assert "ScanContextReconstructor.add_covered_cells_to_pointcloud" in patched and "to_tensor01" in patched
    finally:
        compat.uninstall()
    assert mods[0].ScanContext.ptcloud2sc_vectorized is originals[0]
    assert mods[3].ScanContext.cells_covered_by_box is originals[1]


def test_tiny_cnn_rebuild_loads_reference_weights_and_matches_golden_logits(golden_grids):
    """The torch module handed the engine's grids is the reference's network: strict state_dict load and the
    reference's own fp32 CPU logits (tests/golden/tinycnn_logits.npz, made by oracle/make_golden.py)."""
    import torch
    from distropointclouds_b200.cnn import build_tiny_cnn
    m = build_tiny_cnn()
    sd = torch.load(os.path.join(GOLDEN, "tinycnn_best.pth"), map_location="cpu", weights_only=True)
    m.load_state_dict(sd, strict=True)
    m.eval()
    g = golden_grids["000840"]
    x = (g - g.min()) / (g.max() - g.min())
    with torch.no_grad():
        y = m(torch.from_numpy(x)[None, None])[0, 0].numpy()
    ref = np.load(os.path.join(GOLDEN, "tinycnn_logits.npz"))["logits"]
    np.testing.assert_allclose(y, ref, rtol=0, atol=1e-5)


def test_mask_to_detections():
    from distropointclouds_b200.cnn import mask_to_detections
    m = np.zeros((160, 720), np.uint8)
    m[10:15, 100:106] = 1          # 30 px
    m[50:52, 300:303] = 1          # 6 px: below min_pixels
    d = mask_to_detections(m)
    assert d == [{"center_ring_idx": 12.0, "center_sector_idx": 102.5, "area_pixels": 30}]


# ---- GPU ---------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_patched_reference_shaped_classes_reproduce_golden(golden_scans, golden_grids, golden_peds, ref_live):
    from distropointclouds_b200 import compat
    from oracle import oracle as orc
    sv2, node, rec, tiny = mods = _fake_reference_modules()
    compat.install(mods)
    try:
        sc = sv2.ScanContext()
        for sid in SCAN_IDS[:2]:
            g = sc.ptcloud2sc_vectorized(golden_scans[sid][:, :3], sv2.ScanContext.sector_res[0],
                                         sv2.ScanContext.ring_res[0], sv2.ScanContext.max_length)
            assert g.dtype == np.float64 and np.array_equal(g.astype(np.float32), golden_grids[sid])
        bad = np.array([[np.nan, 1, 1], [1, 2, 3]], np.float32)
        with pytest.raises(IndexError):
            sc.ptcloud2sc_vectorized(bad, np.int64(720), np.int64(160), 20)
        # statics are read from the patched class at call time
        sv2.ScanContext.use_full_height_range = 0
        try:
            g = sc.ptcloud2sc_vectorized(golden_scans["000840"][:, :3], np.int64(720), np.int64(160), 20)
            o, _ = orc.sc_build(golden_scans["000840"], orc.Params(160, 720, 20.0, 0, False, 2.0))
            assert np.array_equal(g.astype(np.float32), o)
        finally:
            sv2.ScanContext.use_full_height_range = 1
        ped = next(p for p in golden_peds["000840"] if p["covered_cells"])
        cells = sv2.ScanContext.cells_covered_by_box(ped["x"], ped["y"], ped["l"], ped["w"], ped["yaw"])
        assert cells.dtype == np.int64 and cells.tolist() == ped["covered_cells"]
        pts, colors = sc.scan_context_to_pointcloud(golden_grids["000840"])
        p_ref, ids_ref = orc.reconstruct(golden_grids["000840"])
        assert np.array_equal(pts, p_ref) and colors.shape == pts.shape and np.all(colors == 0.7)
        # live generator (variant B)
        comp = node.ScanContextComputer()
        gb = comp._scan_context(golden_scans["002073"][:, :3])
        ob, _ = orc.sc_build(golden_scans["002073"], orc.Params(160, 720, 80.0, 1, False, 2.0))
        assert gb.dtype == np.float32 and np.array_equal(gb, ob)
        assert comp._scan_context(np.zeros((0, 3), np.float32)).shape == (160, 720)
        # reconstructor
        r = rec.ScanContextReconstructor()
        pts3, col3, ids3 = r.scan_context_to_pointcloud(golden_grids["000840"])
        assert np.array_equal(pts3, p_ref) and np.array_equal(ids3, ids_ref) and ids3.dtype == np.int32
        cov = [c for p in golden_peds["000840"] for c in p["covered_cells"]]
        cp, cc = r.add_covered_cells_to_pointcloud(golden_grids["000840"], cov)
        assert np.array_equal(cp, ref_live["covered_000840__gather_pts"])
        assert np.all(cc == np.array([1.0, 0.0, 0.0]))
        # to_tensor01
        t = tiny.to_tensor01(golden_grids["002073"])
        assert tuple(t.shape) == (1, 160, 720) and np.array_equal(t.numpy()[0], ref_live["norm01_002073"])
    finally:
        compat.uninstall()


@pytest.mark.gpu
def test_cnn_pipeline_logits_match_reference(golden_scans, golden_grids):
    """Config 4: packed points on the device -> K1 -> K4 -> TinyCNN.  Grids bit-exact; fp32 logits within
    2e-4 absolute of the reference's CPU fp32 logits (cuDNN vs oneDNN summation order; logits span -4.4..-2.8);
    the opt-in fast modes (BN folded; tf32 / fp16 within 2e-2, bf16 within 1e-1)."""
    import torch
    from distropointclouds_b200.cnn import ScanContextCnnPipeline
    from distropointclouds_b200.engine import ScanContextParams
    ref = np.load(os.path.join(GOLDEN, "tinycnn_logits.npz"))["logits"]
    w = os.path.join(GOLDEN, "tinycnn_best.pth")
    scans = [golden_scans["000840"], golden_scans["002073"]]
    pipe = ScanContextCnnPipeline(w, 0, "fp32")
    pts = torch.from_numpy(np.concatenate(scans)).cuda()
    offs = np.array([0, len(scans[0]), len(scans[0]) + len(scans[1])], np.int64)
    grids, logits, mask = pipe.infer_device(pts, offs, ScanContextParams())
    assert np.array_equal(grids[0].cpu().numpy(), golden_grids["000840"])
    assert np.array_equal(grids[1].cpu().numpy(), golden_grids["002073"])
    np.testing.assert_allclose(logits[0].cpu().numpy(), ref, rtol=0, atol=2e-4)
    assert mask.dtype == torch.uint8 and int(mask[0].sum()) == int((1 / (1 + np.exp(-ref)) > 0.5).sum())
    g2, probs, m2 = pipe.infer_scans([s[:, :3] for s in scans])
    assert np.array_equal(g2[0], golden_grids["000840"]) and probs.shape == (2, 160, 720)
    for prec, tol in (("tf32", 2e-2), ("fp16", 2e-2), ("bf16", 1e-1)):
        fast = ScanContextCnnPipeline(w, 0, prec)
        _, lg, _ = fast.infer_device(pts, offs, ScanContextParams())
        np.testing.assert_allclose(lg[0].cpu().numpy(), ref, rtol=0, atol=tol)
    # a scan on which the reference's grid builder raises (NaN coordinate -> INT_MIN index, scaner_version2.py:93-107):
    # the host entry point raises too; strict=False drops the point and keeps going
    bad = scans[0][:1000, :3].copy()
    bad[17, 0] = np.nan
    with pytest.raises(IndexError):
        pipe.infer_scans([scans[1][:, :3], bad])
    g3, _, _ = pipe.infer_scans([scans[1][:, :3], bad], strict=False)
    assert np.array_equal(g3[0], golden_grids["002073"])


@pytest.mark.gpu
def test_engine_refuses_to_run_in_forked_child():
    from distropointclouds_b200.engine import Engine
    e = Engine(0)
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        try:
            try:
                e.kernel_launches
                os.write(w, b"no error")
            except RuntimeError as ex:
                os.write(w, str(ex).encode()[:200])
        finally:
            os._exit(0)
    os.waitpid(pid, 0)
    msg = os.read(r, 300).decode()
    e.close()
    assert "forked child" in msg

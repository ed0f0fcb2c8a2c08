"""The oracle pinned against the reference's own known answers (runs on CPU)."""
import ctypes
import ctypes.util
import os

import numpy as np
import pytest

from conftest import SCAN_IDS, case_params


def _glibc_atan2f():
    libm = ctypes.CDLL(ctypes.util.find_library("m"))
    libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]
    libm.atan2f.restype = ctypes.c_float
    return libm.atan2f


def test_atan2f_matches_glibc_bitwise(orc):
    """glibc's float atan2f is the flavour that produced the golden grids."""
    f = _glibc_atan2f()
    rng = np.random.default_rng(7)
    n = 60000
    sc = [1e-3, 1, 30, 1e4, 1e20, 1e-20]
    x = (rng.standard_normal(n) * rng.choice(sc, n)).astype(np.float32)
    y = (rng.standard_normal(n) * rng.choice(sc, n)).astype(np.float32)
    sp = np.array([0, -0.0, 1, -1, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3.4e38, 2 ** 25, 2 ** -29,
                   0.4375, 0.6875, 1.1875, 2.4375], np.float32)
    X, Y = np.meshgrid(sp, sp)
    x = np.concatenate([x, X.ravel()]); y = np.concatenate([y, Y.ravel()])
    a = orc.atan2f(y, x)
    b = np.array([f(float(p), float(q)) for p, q in zip(y, x)], np.float32)
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    assert same.all(), f"{(~same).sum()} mismatches"


@pytest.mark.parametrize("sid", SCAN_IDS)
def test_variant_a_reproduces_golden_grids_bit_exact(orc, golden_scans, golden_grids, sid):
    out, st = orc.sc_build(golden_scans[sid])
    g = golden_grids[sid]
    assert out.dtype == np.float32 and g.dtype == np.float32 and out.shape == (160, 720)
    assert np.array_equal(out.view(np.uint32) | (out == 0) * 0, g.view(np.uint32)) or np.array_equal(out, g)
    assert (out != g).sum() == 0
    assert st["n_nonfinite"] == 0 and st["n_sector_overflow"] == 0


def test_variant_a_live_reference_fixtures(orc, ref_live):
    """Every arithmetic step except arctan2 itself, against the live reference on this repo's
    seeded inputs; arctan2 values are the ones the generating host produced (host dependent)."""
    for name in ref_live["a_case_names"]:
        pts = ref_live[f"{name}__pts"]
        P = case_params(orc, ref_live[f"{name}__cfg"], orc.VARIANT_A)
        out, _ = orc.sc_build(pts, P, theta_override=ref_live[f"{name}__atan2"])
        assert np.array_equal(out.astype(np.float64), ref_live[f"{name}__grid"]), name


def test_variant_b_live_reference_fixtures(orc, ref_live):
    for name in ref_live["b_case_names"]:
        pts = ref_live[f"{name}__pts"]
        P = case_params(orc, ref_live[f"{name}__cfg"], orc.VARIANT_B)
        out, _ = orc.sc_build(pts, P, theta_override=ref_live[f"{name}__atan2"])
        assert np.array_equal(out, ref_live[f"{name}__grid"]), name


def _theta_deg(a):
    d = (a * np.float32(np.float32(180.0) / np.float32(np.pi))).astype(np.float32)
    return np.where(d < 0, (d + np.float32(360.0)).astype(np.float32), d)


def test_fdlibm_flavour_differs_only_at_sector_boundaries(orc, ref_live):
    """The oracle's fdlibm atan2f vs the generating host's arctan2 flavour (SVML on AVX-512): every
    point whose sector differs lies within 1 ulp (of its angle in degrees) of a sector boundary, and
    the grids differ in no more cells than such points can touch (north-star rule)."""
    for name in ["a_default", "a_real_r40", "a_rmax80"]:
        pts = ref_live[f"{name}__pts"]
        cfg = ref_live[f"{name}__cfg"]
        P = case_params(orc, cfg, orc.VARIANT_A)
        a, _ = orc.sc_build(pts, P)
        b, _ = orc.sc_build(pts, P, theta_override=ref_live[f"{name}__atan2"])
        x = np.where(pts[:, 0] == 0, np.float32(0.001), pts[:, 0])
        y = np.where(pts[:, 1] == 0, np.float32(0.001), pts[:, 1])
        t1, t2 = _theta_deg(orc.atan2f(y, x)), _theta_deg(ref_live[f"{name}__atan2"])
        gap = 360.0 / cfg[1]
        s1, s2 = np.floor(t1.astype(np.float64) / gap), np.floor(t2.astype(np.float64) / gap)
        moved = s1 != s2
        for t in (t1, t2):
            q = t[moved].astype(np.float64) / gap
            dist_deg = np.abs(q - np.round(q)) * gap
            assert (dist_deg <= np.spacing(t[moved]).astype(np.float64)).all()
        assert (a != b).sum() <= 2 * int(moved.sum())


def test_rasteriser_reproduces_golden_json(orc, golden_peds):
    n = 0
    for sid, peds in golden_peds.items():
        for ped in peds:
            c = orc.cells_covered_by_box(ped["x"], ped["y"], ped["l"], ped["w"], ped["yaw"])
            g = np.array(ped["covered_cells"], np.int64).reshape(-1, 2)
            assert c.shape == g.shape and (c == g).all()
            assert orc.pedestrian_to_scan_context_coords(ped["x"], ped["y"]) == \
                (ped["center_ring_idx"], ped["center_sector_idx"])
            n += 1
    assert n == 52


def test_rasteriser_live_reference_fixture(orc, ref_live):
    boxes, offs, cells = ref_live["raster__boxes"], ref_live["raster__offsets"], ref_live["raster__cells"]
    for i, b in enumerate(boxes):
        c = orc.cells_covered_by_box(*b)
        assert np.array_equal(c, cells[offs[i]:offs[i + 1]]), i
    c = orc.cells_covered_by_box(7.5, -3.25, 3.9, 1.6, 0.77, 80, 360, 40.0)
    assert np.array_equal(c, ref_live["raster_alt__cells"])


def test_reconstructor_live_reference_fixture(orc, golden_grids, ref_live):
    for sid in ("000840", "002160"):
        p, ids = orc.reconstruct(golden_grids[sid])
        assert np.array_equal(p, ref_live[f"recon_{sid}__pts"]) and np.array_equal(ids, ref_live[f"recon_{sid}__ids"])
    p, ids = orc.reconstruct(golden_grids["002095"], full_height=False, height_offset=2.0)
    assert np.array_equal(p, ref_live["recon_002095_offset__pts"])
    assert np.array_equal(ids, ref_live["recon_002095_offset__ids"])


def test_normalize01_live_reference_fixture(orc, golden_grids, ref_live):
    assert np.array_equal(orc.normalize01(golden_grids["002073"]), ref_live["norm01_002073"])


def test_batch_driver_matches_single(orc, golden_scans):
    scans = [golden_scans[s][:20000] for s in SCAN_IDS]
    pts = np.concatenate(scans)
    offs = np.cumsum([0] + [len(s) for s in scans]).astype(np.int64)
    out, st = orc.sc_build_batch(pts, offs, n_threads=3)
    for i, s in enumerate(scans):
        one, st1 = orc.sc_build(s)
        assert np.array_equal(out[i], one) and st[i] == st1


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_oracle_against_live_reference_when_present(orc, golden_scans):
    from oracle import ref_shim as ref
    pts = golden_scans["002160"][:25000, :3]
    x = np.where(pts[:, 0] == 0, 0.001, pts[:, 0]); y = np.where(pts[:, 1] == 0, 0.001, pts[:, 1])
    a = ref.variant_a(pts)
    b, _ = orc.sc_build(pts, theta_override=np.arctan2(y, x))
    assert np.array_equal(a, b.astype(np.float64))
    a = ref.variant_b(pts.copy(), max_length=80.0, full_height=False)
    b, _ = orc.sc_build(pts, orc.Params(max_length=80.0, variant=orc.VARIANT_B, full_height=False),
                        theta_override=np.arctan2(pts[:, 1], pts[:, 0]))
    assert np.array_equal(a, b)
    # the reference raises IndexError on its own alien sample and on angles that round to 360
    bad = np.array([[50.0, -1e-7, 0.0]], np.float32)
    with pytest.raises(IndexError):
        ref.variant_a(bad)
    _, st = orc.sc_build(bad)
    assert st["n_sector_overflow"] == 1
    with pytest.raises(IndexError):
        ref.variant_a(np.array([[np.nan, 1.0, 0.0]], np.float32))
    _, st = orc.sc_build(np.array([[np.nan, 1.0, 0.0]], np.float32))
    assert st["n_nonfinite"] == 1


def test_numpy_restatement_reproduces_golden_and_live_reference(orc, golden_scans, golden_grids, ref_live):
    """oracle/restated.py (the NumPy path bench.py times on the GPU box): variant A with the fdlibm arctan2 plugged in
    is the golden grid bit for bit; with the host's own arctan2 both variants equal the live-reference fixtures
    wherever the fixture host's arctan2 flavour agrees with this host's (checked through the C oracle)."""
    from oracle import restated
    sid = "002073"
    pts = golden_scans[sid][:, :3]
    g = restated.variant_a(pts, arctan2=lambda y, x: orc.atan2f(y, x))
    assert g.dtype == np.float64 and np.array_equal(g.astype(np.float32), golden_grids[sid])
    for name in ("a_default", "a_offset", "a_nondyadic"):
        p, cfg, th = ref_live[f"{name}__pts"], ref_live[f"{name}__cfg"], ref_live[f"{name}__atan2"]
        got = restated.variant_a(p, np.int64(cfg[1]), np.int64(cfg[0]), int(cfg[2]), bool(cfg[3]), float(cfg[4]),
                                 arctan2=lambda y, x, th=th: th)
        assert np.array_equal(got, ref_live[f"{name}__grid"]), name
    for name in ("b_default", "b_full", "b_nondyadic"):
        p, cfg, th = ref_live[f"{name}__pts"], ref_live[f"{name}__cfg"], ref_live[f"{name}__atan2"]
        with np.errstate(invalid="ignore"):
            fin = np.isfinite(p).all(axis=1)
            got = restated.variant_b(p.copy(), float(cfg[2]), int(cfg[1]), int(cfg[0]), bool(cfg[3]), float(cfg[4]),
                                     arctan2=lambda y, x, th=th[fin]: th)
        assert np.array_equal(got, ref_live[f"{name}__grid"]), name

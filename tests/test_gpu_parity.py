"""Parity of the CUDA engine (through the C ABI) against the oracle, the reference's golden vectors
and the committed live-reference fixtures.  Bit-exact for grids, cell indices, counters, bin ids and
(because the trig tables come from NumPy exactly as in the reference) the float64 coordinates."""
import numpy as np
import pytest

from conftest import SCAN_IDS, case_params, synth_scan

pytestmark = pytest.mark.gpu


def P(**kw):
    from distropointclouds_b200.engine import ScanContextParams
    return ScanContextParams(**kw)


def eng_params(orc_params, exact_only=False):
    return P(num_ring=orc_params.num_ring, num_sector=orc_params.num_sector, max_length=orc_params.max_length,
             variant=orc_params.variant, full_height=orc_params.full_height, height_offset=orc_params.height_offset,
             gap_f32=orc_params.gap_f32, exact_only=exact_only)


def stats_dict(row):
    from distropointclouds_b200._ffi import STATS_FIELDS
    return dict(zip(STATS_FIELDS, (int(v) for v in row)))


def assert_same_grid(a, b, what=""):
    assert a.shape == b.shape and a.dtype == b.dtype == np.float32
    neq = (a != b) & ~(np.isnan(a) & np.isnan(b))
    assert neq.sum() == 0, f"{what}: {int(neq.sum())} cells differ"
    nz = a != 0
    assert np.array_equal(a.view(np.uint32)[nz], b.view(np.uint32)[nz]), what     # bit-exact where non-zero


def check_against_oracle(engine, orc, scans, op, exact_only=False, what=""):
    grids, _, stats = engine.build_scans(scans, eng_params(op, exact_only))
    for i, s in enumerate(scans):
        o, st = orc.sc_build(np.ascontiguousarray(s), op)
        assert_same_grid(grids[i], o, f"{what}[{i}]")
        d = stats_dict(stats[i])
        for k in ("n_points", "n_nonfinite", "n_sector_overflow", "n_clamped_ring", "n_out_of_range"):
            assert d[k] == st[k], (what, i, k, d, st)
        if exact_only:
            assert d["n_slow_path"] == 0
    return grids, stats


# ---- K1: golden vectors ------------------------------------------------------------------------------
@pytest.mark.parametrize("exact_only", [False, True])
def test_golden_grids_bit_exact(engine, golden_scans, golden_grids, exact_only):
    """BASELINE configs[0]/[1]: all four test_data scans, one batch, bit-exact vs test_data/sc_*.npy."""
    scans = [golden_scans[s] for s in SCAN_IDS]
    grids, norm, stats = engine.build_scans(scans, P(exact_only=exact_only), want_norm=True)
    for i, sid in enumerate(SCAN_IDS):
        g = golden_grids[sid]
        assert (grids[i] != g).sum() == 0, sid
        assert np.array_equal(grids[i].view(np.uint32)[g != 0], g.view(np.uint32)[g != 0])
        d = stats_dict(stats[i])
        assert d["n_points"] == len(scans[i]) and d["n_nonfinite"] == 0 and d["n_sector_overflow"] == 0
        if not exact_only:
            assert 0 < d["n_slow_path"] < 0.01 * d["n_points"]


def test_golden_scans_through_file_view_and_stride12(engine, golden_scans, golden_grids):
    sid = "002095"
    scan = golden_scans[sid]
    view = scan[:, :-1]                                   # load_velo_scan's strided view (stride 16)
    g1, _, _ = engine.build_scans([view], P())
    g2, _, _ = engine.build_scans([np.ascontiguousarray(view)], P())        # packed xyz (stride 12)
    assert (g1[0] != golden_grids[sid]).sum() == 0 and (g2[0] != golden_grids[sid]).sum() == 0


# ---- K1: oracle sweep ------------------------------------------------------------------------------------
CFGS = [
    dict(), dict(full_height=False), dict(max_length=80.0), dict(max_length=80.0, full_height=False, height_offset=1.25),
    dict(num_ring=20, num_sector=90), dict(num_ring=320, num_sector=1440, max_length=40.0),
    dict(num_ring=150, num_sector=700, max_length=33.0), dict(num_ring=150, num_sector=700, max_length=33.0, gap_f32=True),
    dict(variant=1, max_length=80.0, full_height=False), dict(variant=1, max_length=80.0), dict(variant=1),
    dict(variant=1, num_ring=77, num_sector=1000, max_length=33.3, full_height=False, height_offset=0.75),
]


@pytest.mark.parametrize("cfg", CFGS)
def test_engine_equals_oracle_on_real_and_synthetic(engine, orc, golden_scans, cfg):
    op = orc.Params(**cfg)
    scans = [golden_scans["000840"][:40000], synth_scan(3, 30000), golden_scans["002160"][60000:],
             np.zeros((0, 4), np.float32), synth_scan(4, 5000)[:1]]
    check_against_oracle(engine, orc, scans, op, what=str(cfg))
    check_against_oracle(engine, orc, scans[:2], op, exact_only=True, what=str(cfg) + " exact")


def adversarial(seed=3, n=150000):
    rng = np.random.default_rng(seed)
    sc = [1e-3, 1, 30, 1e4, 1e19, 1e-20, 1e-38]
    x = (rng.standard_normal(n) * rng.choice(sc, n)).astype(np.float32)
    y = (rng.standard_normal(n) * rng.choice(sc, n)).astype(np.float32)
    z = rng.standard_normal(n).astype(np.float32)
    idx = rng.integers(0, n, 3000)
    x[idx[:1000]] = 0; y[idx[1000:2000]] = 0; x[idx[2000:2200]] = np.inf; y[idx[2200:2400]] = np.nan
    z[idx[2400:2600]] = np.nan; z[idx[2600:2800]] = -np.inf; x[idx[2800:2900]] = -0.0; y[idx[2900:]] = -0.0
    z[idx[100:200]] = np.inf
    x[:500] = 50; y[:500] = -1e-7 * rng.random(500)
    return np.stack([x, y, z, z], 1)


@pytest.mark.parametrize("variant", [0, 1])
def test_engine_equals_oracle_on_adversarial_points(engine, orc, variant):
    pts = adversarial()
    for cfg in (dict(), dict(max_length=80.0, full_height=False), dict(num_ring=90, num_sector=333, max_length=47.0)):
        op = orc.Params(variant=variant, **cfg)
        check_against_oracle(engine, orc, [pts, pts[::-1].copy()], op, what=f"adv v{variant} {cfg}")


def test_boundary_dense_points(engine, orc):
    """Points sprayed right across ring and sector boundaries: the screen must send them to the exact path."""
    rng = np.random.default_rng(9)
    k = rng.integers(0, 720, 200000)
    th = np.deg2rad(k * 0.5 + rng.normal(0, 2e-4, k.size))
    r = rng.integers(1, 170, k.size) * 0.125 + rng.normal(0, 2e-5, k.size)
    pts = np.stack([r * np.cos(th), r * np.sin(th), rng.normal(0, 1, k.size), np.zeros(k.size)], 1).astype(np.float32)
    for variant in (0, 1):
        _, stats = check_against_oracle(engine, orc, [pts], orc.Params(variant=variant), what="boundary")
        assert stats_dict(stats[0])["n_slow_path"] > 0.5 * len(pts)


def test_live_reference_fixtures_with_boundary_accounting(engine, orc, ref_live):
    """Against the *reference's* outputs: the engine (fdlibm atan2f) may differ from the fixture (whatever
    arctan2 flavour the generating host had) only through points within 1 ulp of a sector boundary; such
    points are counted (north-star rule).  Everything else must be bit-exact."""
    moved_total = 0
    for names, variant in ((ref_live["a_case_names"], 0), (ref_live["b_case_names"], 1)):
        for name in names:
            pts = ref_live[f"{name}__pts"]
            cfg = ref_live[f"{name}__cfg"]
            op = case_params(orc, cfg, variant)
            grids, _, _ = engine.build_scans([np.ascontiguousarray(pts)], eng_params(op))
            mine = grids[0]
            ref_grid = ref_live[f"{name}__grid"].astype(np.float32)
            same_flavour, _ = orc.sc_build(pts, op, theta_override=ref_live[f"{name}__atan2"])
            assert np.array_equal(same_flavour, ref_grid), name              # oracle == reference
            fdl, _ = orc.sc_build(pts, op)
            assert_same_grid(mine, fdl, name)                                # engine == oracle (fdlibm)
            ndiff = int((mine != ref_grid).sum())
            if ndiff:
                if variant == 0:
                    x = np.where(pts[:, 0] == 0, np.float32(0.001), pts[:, 0]); y = np.where(pts[:, 1] == 0, np.float32(0.001), pts[:, 1])
                else:
                    x, y = pts[:, 0], pts[:, 1]
                with np.errstate(invalid="ignore"):
                    ulps = np.abs(orc.atan2f(y, x).view(np.int32).astype(np.int64) - ref_live[f"{name}__atan2"].view(np.int32).astype(np.int64))
                moved = int((ulps > 0).sum())
                assert ndiff <= 2 * moved, (name, ndiff, moved)
                moved_total += ndiff
    print(f"cells differing from the live-reference fixtures through boundary points: {moved_total}")


# ---- K1: device path, sizes, properties ---------------------------------------------------------------------------
def test_device_path_equals_host_path(engine, orc):
    import torch
    scans = [synth_scan(10 + i, 20000 + 3000 * i) for i in range(5)]
    pts = np.concatenate(scans)
    offs = np.cumsum([0] + [len(s) for s in scans]).astype(np.int64)
    p = P()
    host, norm_h, st_h = engine.build_host(pts, offs, p, want_norm=True)
    pad = 7        # offsets[0] != 0: the batch may start anywhere in the buffer
    d_pts = torch.from_numpy(np.concatenate([np.zeros((pad, 4), np.float32), pts])).cuda()
    out = torch.empty((5, 160, 720), dtype=torch.float32, device="cuda")
    nrm = torch.empty_like(out)
    st = torch.zeros((5, 7), dtype=torch.int64, device="cuda")
    engine.build_device(d_pts, offs + pad, p, out, nrm, st)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), host) and np.array_equal(nrm.cpu().numpy(), norm_h)
    assert np.array_equal(st.cpu().numpy(), st_h)
    # twice in a row on the same buffers: the key scratch must come back clean
    engine.build_device(d_pts, offs + pad, p, out, None, None)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), host)


def test_batch_of_120k_scans_vs_oracle(engine, orc):
    """BASELINE configs[2] shape (64-beam ~120k-point scans), 48 scans here so the oracle stays in seconds."""
    scans = [synth_scan(100 + i, 120000) for i in range(48)]
    pts = np.concatenate(scans)
    offs = np.cumsum([0] + [len(s) for s in scans]).astype(np.int64)
    grids, _, stats = engine.build_host(pts, offs, P())
    ref, rst = orc.sc_build_batch(pts, offs)
    assert np.array_equal(grids, ref)
    assert [int(v) for v in stats[:, 3]] == [s["n_clamped_ring"] for s in rst]


def test_dense_2m_point_scan_vs_oracle(engine, orc):
    """BASELINE configs[4] shape: one 128-beam x 16384-step scan (2 097 152 points)."""
    s = synth_scan(7, 2097152, beams=128)
    assert len(s) == 2097152
    grids, _, _ = engine.build_scans([s, s[:1000]], P())
    o, _ = orc.sc_build(s)
    assert np.array_equal(grids[0], o)


def test_size_independent_properties(engine):
    """Properties that hold at any size: determinism, shard invariance (batched == one by one),
    permutation invariance (max is order independent), idempotence under duplication."""
    scans = [synth_scan(200 + i, 120000) for i in range(16)]
    a, _, _ = engine.build_scans(scans, P())
    b, _, _ = engine.build_scans(scans, P())
    assert np.array_equal(a, b)
    for i in (0, 7, 15):
        one, _, _ = engine.build_scans([scans[i]], P())
        assert np.array_equal(one[0], a[i])
    rng = np.random.default_rng(0)
    perm = [s[rng.permutation(len(s))] for s in scans[:4]]
    c, _, _ = engine.build_scans(perm, P())
    assert np.array_equal(c, a[:4])
    dup = [np.concatenate([s, s[::3]]) for s in scans[:2]]
    d, _, _ = engine.build_scans(dup, P())
    assert np.array_equal(d, a[:2])
    # occupancy == cells hit (minus exact-zero heights), checksum of checksums
    assert a.shape == (16, 160, 720) and np.isfinite(a).all()


def test_contention_and_skew_distributions(engine, orc):
    """SURVEY 8(d) adversarial distributions: all points in one cell; uniform over the 80 m disc."""
    rng = np.random.default_rng(1)
    one = np.tile(np.array([[3.0, 4.0, 0.0, 0.0]], np.float32), (100000, 1))
    one[:, 2] = rng.normal(0, 1, 100000)
    r = 80 * np.sqrt(rng.random(200000)); t = rng.random(200000) * 2 * np.pi
    disc = np.stack([r * np.cos(t), r * np.sin(t), rng.normal(0, 1, 200000), np.zeros(200000)], 1).astype(np.float32)
    check_against_oracle(engine, orc, [one, disc], orc.Params(), what="skew")
    g, _, _ = engine.build_scans([one], P())
    assert np.count_nonzero(g[0]) == 1 and g[0].max() == one[:, 2].max()


def test_error_codes(engine):
    from distropointclouds_b200._ffi import DscError
    pts = np.zeros((10, 4), np.float32)
    with pytest.raises(DscError):
        engine.build_host(pts, np.array([0, 10]), P(num_ring=0))
    with pytest.raises(DscError):
        engine.build_host(pts, np.array([5, 2]), P())
    with pytest.raises(DscError):
        engine.build_host(pts, np.array([0, 10]), P(max_length=float("nan")))
    with pytest.raises(DscError):
        engine.build_host(np.zeros((10, 5), np.float32), np.array([0, 10]), P(), stride_bytes=20)
    n0 = engine.kernel_launches
    g, _, _ = engine.build_host(pts, np.array([0, 10]), P())
    assert engine.kernel_launches > n0          # the CUDA path really ran
    assert g.shape == (1, 160, 720)


# ---- K4 ----------------------------------------------------------------------------------------------------------
def test_normalize01(engine, orc, golden_grids, ref_live):
    g = np.stack([golden_grids[s] for s in SCAN_IDS] + [np.zeros((160, 720), np.float32), np.full((160, 720), 2.5, np.float32)])
    out = engine.normalize01_host(g)
    for i in range(len(g)):
        assert np.array_equal(out[i], orc.normalize01(g[i])), i
    assert np.array_equal(out[1], ref_live["norm01_002073"])
    assert np.array_equal(out[4], g[4]) and np.array_equal(out[5], g[5])      # max == min: unchanged


# ---- K2 ----------------------------------------------------------------------------------------------------------
def test_rasteriser_golden_json(engine, golden_peds):
    boxes, want = [], []
    for sid in SCAN_IDS:
        for ped in golden_peds[sid]:
            boxes.append([ped["x"], ped["y"], ped["l"], ped["w"], ped["yaw"]])
            want.append(np.array(ped["covered_cells"], np.int64).reshape(-1, 2))
    assert len(boxes) == 52
    got, mask = engine.raster_boxes_host(np.array(boxes), want_mask=True)
    for g, w in zip(got, want):
        assert g.dtype == np.int64 and g.shape == w.shape and (g == w).all()
    m = np.zeros((160, 720), np.uint8)
    for w in want:
        m[w[:, 0], w[:, 1]] = 1
    assert np.array_equal(mask, m)
    one, _ = engine.raster_boxes_host(np.array(boxes[:1]))
    assert np.array_equal(one[0], want[0])
    none, _ = engine.raster_boxes_host(np.zeros((0, 5)))
    assert none == []


def test_rasteriser_live_reference_fixture(engine, orc, ref_live):
    boxes, offs, cells = ref_live["raster__boxes"], ref_live["raster__offsets"], ref_live["raster__cells"]
    got, _ = engine.raster_boxes_host(boxes)
    for i in range(len(boxes)):
        assert np.array_equal(got[i], cells[offs[i]:offs[i + 1]]), i
    alt, _ = engine.raster_boxes_host(np.array([[7.5, -3.25, 3.9, 1.6, 0.77]]), 80, 360, 40.0)
    assert np.array_equal(alt[0], ref_live["raster_alt__cells"])
    rng = np.random.default_rng(5)
    nb = 300
    rb = np.stack([rng.uniform(-22, 22, nb), rng.uniform(-22, 22, nb), rng.uniform(0.2, 6, nb), rng.uniform(0.2, 3, nb),
                   rng.uniform(-7, 7, nb)], 1)
    got, _ = engine.raster_boxes_host(rb)
    for i in range(0, nb, 7):
        assert np.array_equal(got[i], orc.cells_covered_by_box(*rb[i])), i


# ---- K3 / K5 -------------------------------------------------------------------------------------------------------
def test_reconstructor_fixtures(engine, orc, golden_grids, ref_live):
    out = engine.reconstruct_host(np.stack([golden_grids[s] for s in SCAN_IDS]))
    for i, sid in enumerate(SCAN_IDS):
        p, ids = out[i]
        op, oi = orc.reconstruct(golden_grids[sid])
        assert p.dtype == np.float64 and ids.dtype == np.int32
        assert np.array_equal(p, op) and np.array_equal(ids, oi), sid
        assert (np.diff(ids) > 0).all()
        # north-star tolerance (1e-5 relative) is met with margin: the match is bit-exact
    assert np.array_equal(out[0][0], ref_live["recon_000840__pts"]) and np.array_equal(out[0][1], ref_live["recon_000840__ids"])
    assert np.array_equal(out[3][0], ref_live["recon_002160__pts"])
    p, ids = engine.reconstruct_host(golden_grids["002095"], full_height=False, height_offset=2.0)[0]
    assert np.array_equal(p, ref_live["recon_002095_offset__pts"]) and np.array_equal(ids, ref_live["recon_002095_offset__ids"])
    e = engine.reconstruct_host(np.zeros((2, 160, 720), np.float32))
    assert e[0][0].shape == (0, 3) and e[1][1].shape == (0,)


def test_build_then_reconstruct_round_trip(engine, golden_scans):
    """grid -> points -> grid is the identity on occupied cells (every reconstructed point falls back
    into its own cell, carrying its height)."""
    g, _, _ = engine.build_scans([golden_scans["000840"]], P())
    p, ids = engine.reconstruct_host(g[0])[0]
    back, _, _ = engine.build_scans([np.concatenate([p, np.zeros((len(p), 1))], 1).astype(np.float32)], P())
    assert np.array_equal(back[0], g[0])
    assert np.array_equal(np.flatnonzero(g[0].ravel()).astype(np.int32), ids)


def test_covered_cells_fixtures(engine, golden_grids, golden_peds, ref_live):
    cells = [c for ped in golden_peds["000840"] for c in ped["covered_cells"]]
    g = golden_grids["000840"]
    pts = engine.gather_cells_host(g, np.array(cells + [[-3, 5], [400, 9999]]))
    assert np.array_equal(pts[:-2], ref_live["covered_000840__gather_pts"])
    assert pts[-2, 2] == g[0, 5] and pts[-1, 2] == g[159, 719]                # clipped like the reference
    _, ids = engine.reconstruct_host(g)[0]
    m = engine.covered_mask_host(ids, np.array(cells))
    assert m.dtype == bool and np.array_equal(m, ref_live["covered_000840__mask"])


# ---- added with the single-pass K3 / one-pass K4 / split-mode K1 ----------------------------------------------------
@pytest.mark.parametrize("shape", [(40, 90), (8, 1024), (3, 2000), (200, 36), (160, 722)])
def test_reconstructor_other_grids_vs_oracle(engine, orc, shape):
    """Vector path (S % 4 == 0, up to 8 warp rounds per row) and the generic three-kernel path (odd S, very long
    rows), dense / sparse / empty grids, both height modes -- bit-exact against the oracle."""
    R, S = shape
    rng = np.random.default_rng(R * 1000 + S)
    grids = np.zeros((5, R, S), np.float32)
    grids[0] = rng.standard_normal((R, S)).astype(np.float32)                         # dense
    m = rng.random((R, S)) < 0.07
    grids[1][m] = rng.uniform(-3, 3, m.sum()).astype(np.float32)                      # sparse
    grids[3][R - 1, S - 1] = 1.5                                                       # one cell, the very last
    grids[4] = np.where(rng.random((R, S)) < 0.5, rng.uniform(-1, 1, (R, S)), 0).astype(np.float32)
    for full in (True, False):
        out = engine.reconstruct_host(grids, R, S, 33.0, full_height=full, height_offset=0.75)
        for i in range(len(grids)):
            op, oi = orc.reconstruct(grids[i], R, S, 33.0, full, 0.75)
            assert np.array_equal(out[i][0], op) and np.array_equal(out[i][1], oi), (shape, full, i)


def test_reconstruct_device_capacity_and_offsets(engine, golden_grids):
    """Device form: CSR offsets on the device, writes clipped at the caller's capacity, many grids (look-back chain)."""
    import torch
    g = torch.from_numpy(np.stack([golden_grids[s] for s in SCAN_IDS] * 75)).cuda()      # 300 grids > #SMs
    n = g.shape[0]
    counts = np.array([np.count_nonzero(golden_grids[s]) for s in SCAN_IDS] * 75)
    offs = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    total = int(counts.sum())
    pts = torch.empty((total, 3), dtype=torch.float64, device="cuda")
    ids = torch.empty(total, dtype=torch.int32, device="cuda")
    engine.reconstruct_device(g, offs, pts, ids)
    torch.cuda.synchronize()
    assert np.array_equal(offs.cpu().numpy(), np.concatenate([[0], np.cumsum(counts)]))
    first = engine.reconstruct_host(golden_grids[SCAN_IDS[0]])[0]
    a, b = int(offs[296]), int(offs[297])
    assert np.array_equal(pts[a:b].cpu().numpy(), first[0]) and np.array_equal(ids[a:b].cpu().numpy(), first[1])
    # capacity smaller than the result: offsets still complete, nothing written past the end
    small = torch.full((1000, 3), -7.0, dtype=torch.float64, device="cuda")
    guard = torch.full((1000,), -7, dtype=torch.int32, device="cuda")
    engine.reconstruct_device(g[:2], offs[:3], small[:900], guard[:900])
    torch.cuda.synchronize()
    assert int(offs[2]) == int(counts[:2].sum()) and float(small[900:].max()) == -7.0 and int(guard[900:].max()) == -7
    assert np.array_equal(small[:900].cpu().numpy(), first[0][:900])


def test_normalize01_odd_sizes_and_constant_grids(engine, orc):
    rng = np.random.default_rng(5)
    for shape in [(3, 7, 9), (2, 160, 720), (5, 1, 4), (1, 33, 100)]:
        g = rng.standard_normal(shape).astype(np.float32) * 5
        g[0] = 2.5                                                     # max == min: returned unchanged
        got = engine.normalize01_host(g)
        for i in range(shape[0]):
            assert np.array_equal(got[i], orc.normalize01(g[i])), (shape, i)


@pytest.mark.parametrize("item_points", [16384, 50000, 1 << 20])
def test_split_items_are_invisible(orc, golden_scans, item_points):
    """K1 cuts big scans / small batches into work items; the grids must not depend on how."""
    from distropointclouds_b200.engine import Engine
    rng = np.random.default_rng(11)
    base = golden_scans["002095"]
    scans = [base, base[:1], base[:0], np.concatenate([base, golden_scans["000840"], base[::-1]]), base[:70001]]
    e = Engine(0)
    try:
        e.set_option("sc_split", 1)
        e.set_option("sc_item_points", item_points)
        got, _, stats = e.build_scans(scans, P())
    finally:
        e.close()
    for i, sc in enumerate(scans):
        ref, st = orc.sc_build(sc) if len(sc) else (np.zeros((160, 720), np.float32), None)
        assert np.array_equal(got[i], ref), i
        assert int(stats[i][0]) == len(sc)


def test_full_size_batch_properties(orc):
    """BASELINE configs[2] at full size (4096 x 120 000 points, device resident, the whole-scan path of K1):
    a sample of the grids against the oracle, and the whole batch against the same scans pushed through in
    small batches (split-item path) -- two different kernels paths must agree on every one of the 4096 grids."""
    import torch
    from distropointclouds_b200.engine import Engine
    from distropointclouds_b200.synth import synth_batch_device
    B, N = 4096, 120000
    pts = synth_batch_device(B, N, 64, seed=3, device="cuda")
    N = pts.shape[1]
    offs = np.arange(B + 1, dtype=np.int64) * N
    e = Engine(0)
    try:
        out = torch.empty((B, 160, 720), dtype=torch.float32, device="cuda")
        st = torch.zeros((B, 7), dtype=torch.int64, device="cuda")
        e.build_device(pts, offs, P(), out, None, st)
        again = torch.empty_like(out)
        e.build_device(pts, offs, P(), again, None, None)
        torch.cuda.synchronize()
        assert torch.equal(out, again)                                           # run-to-run determinism
        assert int(st[:, 0].min()) == N and int(st[:, 0].max()) == N and int(st[:, 1].sum()) == 0
        for i in (0, 1, 147, 148, 2047, 4094, 4095):                             # first/last wave, CTA boundaries
            ref, rst = orc.sc_build(pts[i].cpu().numpy())
            assert np.array_equal(out[i].cpu().numpy(), ref), i
            assert int(st[i, 3]) == rst["n_clamped_ring"], i
        small = torch.empty((64, 160, 720), dtype=torch.float32, device="cuda")
        for a in range(0, B, 64):                                                # 64 scans per call: split items
            e.build_device(pts[a:a + 64], offs[: 65], P(), small, None, None)
            assert torch.equal(small, out[a:a + 64]), a
    finally:
        e.close()


def test_full_size_screened_equals_exact_only():
    """The screened fast path against DSC_FLAG_EXACT_ONLY on every one of the 4096 x 120 000 = 491 M points of
    BASELINE configs[2] (real MUFU.RCP / MUFU.SQRT, not the host emulation): all 4096 grids and every counter but
    n_slow_path must be identical -- the screen's error budget (dsc_math.cuh) checked at device scale."""
    import torch
    from distropointclouds_b200.engine import Engine
    from distropointclouds_b200.synth import synth_batch_device
    B, N = 4096, 120000
    pts = synth_batch_device(B, N, 64, seed=5, device="cuda")
    N = pts.shape[1]
    offs = np.arange(B + 1, dtype=np.int64) * N
    e = Engine(0)
    try:
        for variant, ml, full in ((0, 20.0, True), (1, 80.0, False)):
            fast = torch.empty((B, 160, 720), dtype=torch.float32, device="cuda")
            exact = torch.empty_like(fast)
            st_f = torch.zeros((B, 7), dtype=torch.int64, device="cuda")
            st_e = torch.zeros_like(st_f)
            e.build_device(pts, offs, P(max_length=ml, variant=variant, full_height=full), fast, None, st_f)
            e.build_device(pts, offs, P(max_length=ml, variant=variant, full_height=full, exact_only=True), exact, None, st_e)
            torch.cuda.synchronize()
            assert torch.equal(fast, exact), variant
            assert torch.equal(st_f[:, :5], st_e[:, :5]), variant
            assert int(st_f[:, 5].sum()) > 0 and int(st_e[:, 5].sum()) == 0
            del fast, exact
    finally:
        e.close()


@pytest.mark.parametrize("n_scans", [70, 150, 300])
def test_many_small_scans_with_empty_ones(engine, orc, n_scans):
    """Batches of 65+ small scans with empty scans among them (an empty scan has no tile and no work item; round 1's
    tile scheduler could wait for one forever -- ADVICE r1)."""
    rng = np.random.default_rng(n_scans)
    scans = []
    for i in range(n_scans):
        n = 0 if i < 4 or i % 37 == 0 else int(rng.integers(3000, 7000))
        p = (rng.standard_normal((n, 4)) * [9, 9, 1.5, 1]).astype(np.float32)
        scans.append(p)
    got, _, stats = engine.build_scans(scans, P())
    for i, sc in enumerate(scans):
        ref = orc.sc_build(sc)[0] if len(sc) else np.zeros((160, 720), np.float32)
        assert np.array_equal(got[i], ref), i
        assert int(stats[i][0]) == len(sc)


def test_reconstructor_dense_grids_take_several_compaction_rounds(orc):
    """K3 compacts a grid's kept cells into shared memory (~33 k entries); denser grids are emitted in several rounds.
    Fully dense, 60 % dense, one dense ring band, and a sparse grid in one batch, both height modes."""
    from distropointclouds_b200.engine import Engine
    rng = np.random.default_rng(21)
    g = np.zeros((5, 160, 720), np.float32)
    g[0] = rng.standard_normal((160, 720)).astype(np.float32) + 0.01                     # every cell kept
    g[1] = np.where(rng.random((160, 720)) < 0.6, rng.standard_normal((160, 720)), 0).astype(np.float32)
    g[2, 40:110] = (rng.standard_normal((70, 720)) * 3).astype(np.float32)              # 50 400 kept cells in a band
    g[3] = np.where(rng.random((160, 720)) < 0.05, 1.5, 0).astype(np.float32)
    g[4, ::2, ::3] = -0.25
    with Engine(0) as e:
        for full in (True, False):
            out = e.reconstruct_host(g, full_height=full, height_offset=2.0)
            for i in range(len(g)):
                op, oi = orc.reconstruct(g[i], full_height=full, height_offset=2.0)
                assert np.array_equal(out[i][1], oi), (full, i)
                assert np.array_equal(out[i][0], op), (full, i)


@pytest.mark.parametrize("shape", [0, 1, 2])
@pytest.mark.parametrize("split", [0, 1])
def test_every_launch_shape_of_the_scan_kernel(orc, golden_scans, golden_grids, shape, split):
    """k1_scan ships three launch shapes for the default grid (768 threads x 6 loads per register set, 768 x 3,
    1024 x 2; "sc_cfg").  Each one, as whole-scan items and as split items, on the golden scans and on ragged
    synthetic scans whose tails exercise the masked last register sets (0, 1, a few hundred and a few thousand
    points past a multiple of the set size), both variants, both height modes: bit-exact against the oracle."""
    from distropointclouds_b200.engine import Engine
    rng = np.random.default_rng(100 + shape)
    sizes = [0, 1, 767, 768, 4607, 4608, 4609, 9216 + 300, 3 * 4608 + 2047, 31000]
    scans = [golden_scans["000840"], golden_scans["002073"]]
    scans += [(rng.standard_normal((n, 4)) * [11, 11, 1.5, 1]).astype(np.float32) for n in sizes]
    with Engine(0) as e:
        e.set_option("k1_impl", 3)
        e.set_option("sc_cfg", shape)
        e.set_option("sc_split", split)
        for variant in (0, 1):
            for full in (True, False):
                op = orc.Params(160, 720, 20.0 if variant == 0 else 80.0, variant, full, 2.0)
                check_against_oracle(e, orc, scans, op, what=f"shape{shape} split{split} v{variant} full{full}")
        if shape == 0 and split == 0:
            got, _, _ = e.build_scans(scans[:2], P())
            assert np.array_equal(got[0], golden_grids["000840"]) and np.array_equal(got[1], golden_grids["002073"])

"""The per-point binning code of the CUDA kernels (csrc/dsc_math.cuh) compiled for the host with
g++ -ffp-contract=off and compared with the oracle point by point -- catches arithmetic-order bugs
before any GPU time is spent."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, SCAN_IDS


@pytest.fixture(scope="module")
def hm():
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostmath.so")
    src = os.path.join(ROOT, "tests", "host_math.cpp")
    hdrs = [os.path.join(ROOT, "distropointclouds_b200", "csrc", h) for h in ("dsc_math.cuh", "dsc_math_f64.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-msse2",
                               "-DDSC_HAVE_FAST", "-o", so, src])
    L = C.CDLL(so)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.hm_bin_exact.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_float, C.c_int, fp,
                               C.c_longlong, C.c_int, ip, ip, fp]
    L.hm_bin_fast.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_float, fp, C.c_longlong, C.c_int,
                              fp, ip, ip, ip]
    L.hm_defgrid_mismatches.argtypes = [C.c_double, fp, C.c_longlong, C.c_int, fp]
    L.hm_defgrid_mismatches.restype = C.c_longlong
    L.hm_atan2f.argtypes = [C.c_float, C.c_float]
    L.hm_atan2f.restype = C.c_float
    L.hm_float_to_key.argtypes = [C.c_float]
    L.hm_float_to_key.restype = C.c_uint
    L.hm_key_to_float.argtypes = [C.c_uint]
    L.hm_key_to_float.restype = C.c_float
    return L


def bin_exact(L, pts, R=160, S=720, ml=20.0, var=0, full=1, off=2.0, g32=0):
    pts = np.ascontiguousarray(pts, np.float32)
    n = len(pts)
    cell = np.empty(n, np.int32); st = np.empty(n, np.int32); z = np.empty(n, np.float32)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.hm_bin_exact(R, S, ml, var, full, off, g32, pts.ctypes.data_as(fp), n, pts.shape[1], cell.ctypes.data_as(ip),
                   st.ctypes.data_as(ip), z.ctypes.data_as(fp))
    return cell, st, z


def bin_fast(L, pts, noise=None, R=160, S=720, ml=20.0, var=0, full=1, off=2.0):
    pts = np.ascontiguousarray(pts, np.float32)
    n = len(pts)
    cell = np.empty(n, np.int32); st = np.empty(n, np.int32); slow = np.empty(n, np.int32)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    nz = None if noise is None else np.ascontiguousarray(noise, np.float32)
    L.hm_bin_fast(R, S, ml, var, full, off, pts.ctypes.data_as(fp), n, pts.shape[1],
                  nz.ctypes.data_as(fp) if nz is not None else None, cell.ctypes.data_as(ip), st.ctypes.data_as(ip),
                  slow.ctypes.data_as(ip))
    return cell, st, slow


def grid_from(cell, st, z, R, S, full):
    g = np.full(R * S, -np.inf if full else 0.0, np.float32)
    ok = (st <= 1) & ~np.isnan(z)
    np.maximum.at(g, cell[ok], z[ok])
    g[g == -np.inf] = 0
    return g.reshape(R, S)


CFGS = [(160, 720, 20, 0, 1, 0), (160, 720, 80, 0, 0, 0), (150, 700, 33, 0, 1, 0), (150, 700, 33, 0, 1, 1),
        (160, 720, 80, 1, 0, 0), (160, 720, 20, 1, 1, 0), (77, 1000, 50.5, 1, 0, 0)]


@pytest.mark.parametrize("sid", SCAN_IDS[:2])
def test_exact_path_equals_oracle_on_real_scans(hm, orc, golden_scans, sid):
    pts = golden_scans[sid]
    for (R, S, ml, var, full, g32) in CFGS:
        cell, st, z = bin_exact(hm, pts, R, S, ml, var, full, 2.0, g32)
        o, stt = orc.sc_build(pts, orc.Params(R, S, ml, var, bool(full), 2.0, bool(g32)))
        assert np.array_equal(grid_from(cell, st, z, R, S, full), o)
        assert (st == 1).sum() == stt["n_clamped_ring"] and (st == 4).sum() == stt["n_out_of_range"]


def adversarial_points(seed=3, n=200000):
    rng = np.random.default_rng(seed)
    sc = [1e-3, 1, 30, 1e4, 1e19, 1e-20, 1e-38]
    x = (rng.standard_normal(n) * rng.choice(sc, n)).astype(np.float32)
    y = (rng.standard_normal(n) * rng.choice(sc, n)).astype(np.float32)
    z = rng.standard_normal(n).astype(np.float32)
    idx = rng.integers(0, n, 3000)
    x[idx[:1000]] = 0; y[idx[1000:2000]] = 0; x[idx[2000:2200]] = np.inf; y[idx[2200:2400]] = np.nan
    z[idx[2400:2600]] = np.nan; z[idx[2600:2800]] = -np.inf; x[idx[2800:2900]] = -0.0; y[idx[2900:]] = -0.0
    x[:500] = 50; y[:500] = -1e-7 * rng.random(500)      # tiny negative angles: theta rounds to 360.0
    return np.stack([x, y, z, z], 1)


def test_exact_path_equals_oracle_on_adversarial_points(hm, orc):
    pts = adversarial_points()
    for var in (0, 1):
        for (R, S, ml, full) in [(160, 720, 20.0, 1), (160, 720, 80.0, 0), (90, 333, 47.0, 1)]:
            cell, st, z = bin_exact(hm, pts, R, S, ml, var, full)
            o, stt = orc.sc_build(pts, orc.Params(R, S, ml, var, bool(full)))
            g = grid_from(cell, st, z, R, S, full)
            assert ((g != o) & ~(np.isnan(g) & np.isnan(o))).sum() == 0
            assert (st == 2).sum() == stt["n_nonfinite"] and (st == 3).sum() == stt["n_sector_overflow"]
            assert (st == 4).sum() == stt["n_out_of_range"] and (st == 1).sum() == stt["n_clamped_ring"]


def test_atan2f_device_source_equals_oracle(hm, orc):
    rng = np.random.default_rng(11)
    x = (rng.standard_normal(50000) * rng.choice([1e-3, 1, 30, 1e4], 50000)).astype(np.float32)
    y = (rng.standard_normal(50000) * rng.choice([1e-3, 1, 30, 1e4], 50000)).astype(np.float32)
    a = orc.atan2f(y, x)
    b = np.array([hm.hm_atan2f(float(p), float(q)) for p, q in zip(y, x)], np.float32)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_order_preserving_key_roundtrip(hm):
    vals = np.array([-np.inf, -3.4e38, -1.5, -1e-45, -0.0, 0.0, 1e-45, 0.25, 2.9, 3.4e38, np.inf], np.float32)
    keys = [hm.hm_float_to_key(float(v)) for v in vals]
    assert keys == sorted(keys) and len(set(keys)) == len(keys)
    assert keys[0] == 0x007FFFFF
    for v, k in zip(vals, keys):
        back = hm.hm_key_to_float(k)
        assert back == (0.0 if v == -np.inf else v)
    assert hm.hm_key_to_float(0) == 0.0


# ---- screened fast path -----------------------------------------------------------------------------
def _check_fast(hm, pts, R=160, S=720, ml=20.0, var=0, full=1, noise_mode="rand", seed=0):
    n = len(pts)
    rng = np.random.default_rng(seed)
    if noise_mode == "rand":
        noise = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    else:
        noise = np.full((n, 2), float(noise_mode), np.float32)
    ce, se, _ = bin_exact(hm, pts, R, S, ml, var, full)
    cf, sf, slow = bin_fast(hm, pts, noise, R, S, ml, var, full)
    acc = slow == 0
    bad = acc & ((se != sf) | ((se <= 1) & (ce != cf)))
    return int(bad.sum()), float(slow.mean())


@pytest.mark.parametrize("sid", SCAN_IDS)
def test_fast_path_never_disagrees_with_exact_on_real_scans(hm, golden_scans, sid):
    pts = golden_scans[sid]
    for var, ml, full in [(0, 20.0, 1), (0, 80.0, 0), (1, 80.0, 0), (1, 20.0, 1)]:
        for nm in ("rand", 1.0, -1.0):
            bad, slow = _check_fast(hm, pts, ml=ml, var=var, full=full, noise_mode=nm)
            assert bad == 0
            assert slow < 0.01, slow        # the screen must stay cheap: < 1 % of points fall back


def test_fast_path_on_synthetic_and_adversarial_points(hm):
    from conftest import synth_scan
    pts = synth_scan(1, n=60000)
    for (R, S, ml) in [(160, 720, 20.0), (80, 360, 40.0), (320, 1440, 40.0), (20, 90, 80.0)]:
        for var in (0, 1):
            bad, slow = _check_fast(hm, pts, R, S, ml, var, 1)
            assert bad == 0 and slow < 0.02
    adv = adversarial_points(n=120000)
    for var in (0, 1):
        for nm in ("rand", 1.0, -1.0):
            bad, _ = _check_fast(hm, adv, var=var, noise_mode=nm)
            assert bad == 0
    # dense sweep right across sector and ring boundaries
    rng = np.random.default_rng(9)
    k = rng.integers(0, 720, 100000)
    th = np.deg2rad(k * 0.5 + rng.normal(0, 2e-4, k.size))
    r = rng.integers(1, 170, k.size) * 0.125 + rng.normal(0, 2e-5, k.size)
    pts = np.stack([r * np.cos(th), r * np.sin(th), rng.normal(0, 1, k.size), np.zeros(k.size)], 1).astype(np.float32)
    for var in (0, 1):
        for nm in ("rand", 1.0, -1.0):
            bad, slow = _check_fast(hm, pts, var=var, noise_mode=nm)
            assert bad == 0
            assert slow > 0.5        # these points really are near boundaries: most must fall back


def test_default_grid_specialisation_is_bit_identical(hm, golden_scans):
    """screened_core_t<true> (constants baked in as immediates) == screened_core_t<false> (constant bank)."""
    fp = C.POINTER(C.c_float)
    rng = np.random.default_rng(4)
    for pts in (golden_scans["000840"], adversarial_points(n=100000)):
        pts = np.ascontiguousarray(pts, np.float32)
        noise = rng.uniform(-1, 1, (len(pts), 2)).astype(np.float32)
        for ml in (20.0, 80.0):
            assert hm.hm_defgrid_mismatches(ml, pts.ctypes.data_as(fp), len(pts), 4, noise.ctypes.data_as(fp)) == 0


def test_random_bit_patterns(hm, orc):
    """x, y, z drawn uniformly from ALL float32 bit patterns (every exponent, denormals, infinities, NaNs, both
    zeros): the exact path equals the oracle point for point, and the screened path never accepts a wrong bin."""
    rng = np.random.default_rng(2026)
    n = 300000
    bits = rng.integers(0, 2 ** 32, (n, 4), dtype=np.uint64).astype(np.uint32)
    pts = bits.view(np.float32)
    # half of the points get moderate magnitudes in x/y so that the in-range code paths are fuzzed too
    m = rng.random(n) < 0.5
    pts[m, 0] = (rng.standard_normal(m.sum()) * 15).astype(np.float32)
    pts[m, 1] = (rng.standard_normal(m.sum()) * 15).astype(np.float32)
    for var in (0, 1):
        for full in (1, 0):
            cell, st, z = bin_exact(hm, pts, 160, 720, 20.0, var, full)
            with np.errstate(all="ignore"):
                o, stt = orc.sc_build(pts, orc.Params(160, 720, 20.0, var, bool(full)))
            g = grid_from(cell, st, z, 160, 720, full)
            assert ((g != o) & ~(np.isnan(g) & np.isnan(o))).sum() == 0
            assert (st == 2).sum() == stt["n_nonfinite"] and (st == 3).sum() == stt["n_sector_overflow"]
            assert (st == 4).sum() == stt["n_out_of_range"] and (st == 1).sum() == stt["n_clamped_ring"]
            bad, _ = _check_fast(hm, pts, var=var, full=full, noise_mode="rand")
            assert bad == 0


def test_mixed_gap_dtypes_follow_numpy(hm, orc):
    """Non-dyadic gaps with only ONE of the two resolutions given as np.int64: the two floor-divides run in
    different dtypes.  The device code (host build) == the oracle == NumPy itself, point by point."""
    rng = np.random.default_rng(12)
    pts = (rng.standard_normal((20000, 3)) * [12, 12, 1]).astype(np.float32)
    pts[:, 0] = np.where(pts[:, 0] == 0, 0.001, pts[:, 0]); pts[:, 1] = np.where(pts[:, 1] == 0, 0.001, pts[:, 1])
    R, S, ml = 150, 700, 33
    th = np.rad2deg(orc.atan2f(pts[:, 1], pts[:, 0])) % 360
    r = np.sqrt(pts[:, 0] * pts[:, 0] + pts[:, 1] * pts[:, 1])
    for mask, nr, ns in ((2, np.int64(R), S), (1, R, np.int64(S))):
        ring = np.minimum(r // (ml / nr), R - 1).astype(np.int32)          # NumPy's own promotion decides
        sec = (th // (360 / ns)).astype(np.int32)
        cell, st, _ = bin_exact(hm, pts, R, S, float(ml), 0, 1, 2.0, 5 if mask == 1 else 2)   # 5 -> (4|1) is not 1: ring only
        ok = st <= 1
        assert np.array_equal(cell[ok], (ring * S + sec)[ok]), mask
        o, _ = orc.sc_build(pts, orc.Params(R, S, float(ml), 0, True, 2.0, mask))
        g = grid_from(cell, st, pts[:, 2].copy(), R, S, True)
        assert np.array_equal(g, o), mask

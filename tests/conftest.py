import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
SCAN_IDS = ["000840", "002073", "002095", "002160"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_scans():
    z = np.load(os.path.join(GOLDEN, "scans.npz"))
    return {s: np.ascontiguousarray(z[f"s{s}"].T) for s in SCAN_IDS}      # (N,4) float32


@pytest.fixture(scope="session")
def golden_grids():
    z = np.load(os.path.join(GOLDEN, "sc_golden.npz"))
    return {s: z[f"sc{s}"] for s in SCAN_IDS}


@pytest.fixture(scope="session")
def golden_peds():
    return {s: json.load(open(os.path.join(GOLDEN, f"pedestrians_{s}.json"))) for s in SCAN_IDS}


@pytest.fixture(scope="session")
def ref_live():
    return np.load(os.path.join(GOLDEN, "ref_live.npz"))


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def engine():
    from distropointclouds_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def case_params(orc_mod, cfg, variant):
    R, S, ml, full, off, pyint = cfg
    return orc_mod.Params(int(R), int(S), float(ml), variant, bool(full), float(off), bool(pyint))


def synth_scan(seed, n=120000, beams=64):
    """SURVEY 8(d) config-3 generator: 64-beam street-canyon scan, beam-major order, mm-quantised."""
    rng = np.random.default_rng(20251018 + seed)
    az_steps = max(1, n // beams)
    el = np.deg2rad(np.linspace(2.0, -24.8, beams))
    phi0 = rng.uniform(0, 2 * np.pi)
    phi = 2 * np.pi * np.arange(az_steps) / az_steps + phi0
    EL, PH = np.meshgrid(el, phi, indexing="ij")           # beam-major
    h = 1.73
    with np.errstate(divide="ignore"):
        ground = np.where(EL < 0, h / np.sin(-EL), np.inf)
    wall_y = rng.uniform(4, 12)
    dirx, diry = np.cos(EL) * np.cos(PH), np.cos(EL) * np.sin(PH)
    with np.errstate(divide="ignore", invalid="ignore"):
        wall = np.where(np.abs(diry) > 1e-6, wall_y / np.abs(diry), np.inf)
    rngd = np.minimum(np.minimum(ground, wall), 80.0)
    # a few boxes
    for _ in range(32):
        bd, ba, bs = rng.uniform(3, 40), rng.uniform(0, 2 * np.pi), rng.uniform(0.5, 4)
        hit = (np.abs(((PH - ba + np.pi) % (2 * np.pi)) - np.pi) < bs / (2 * bd)) & (rngd * np.cos(EL) > bd) \
            & (h + bd * np.tan(EL) < 2.0)
        rngd = np.where(hit, bd / np.cos(EL), rngd)
    rngd = rngd * (1 + rng.normal(0, 0.002, rngd.shape))
    x = rngd * dirx; y = rngd * diry; z = rngd * np.sin(EL)
    pts = np.stack([x.ravel(), y.ravel(), z.ravel(), rng.uniform(0, 1, x.size)], axis=1)
    pts[:, :3] = np.round(pts[:, :3] * 1000) / 1000
    pts[:, 3] = np.round(pts[:, 3] * 100) / 100
    pts = pts.astype(np.float32)
    k = rng.choice(pts.shape[0], max(1, pts.shape[0] // 10000), replace=False)
    pts[k, 0] = 0.0
    return pts

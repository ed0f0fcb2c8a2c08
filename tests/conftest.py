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


@pytest.fixture(scope="session", params=["auto", "l2keys", "scan", "scan_nosplit"])
def engine(request):
    """One engine per K1 path: the default (persistent scan kernel; small batches and big scans are split into
    work items), the generic L2-key kernels, the scan kernel forced on every shape it takes, and the scan kernel
    with whole-scan items only (spill lists + band passes even for one tiny scan)."""
    from distropointclouds_b200.engine import Engine
    e = Engine(0)
    e.set_option("k1_impl", {"auto": 0, "l2keys": 1, "scan": 3, "scan_nosplit": 3}[request.param])
    if request.param == "scan_nosplit":
        e.set_option("sc_split", 0)          # one CTA per scan even for tiny batches
    yield e
    e.close()


def case_params(orc_mod, cfg, variant):
    R, S, ml, full, off, pyint = cfg
    return orc_mod.Params(int(R), int(S), float(ml), variant, bool(full), float(off), bool(pyint))


from distropointclouds_b200.synth import synth_scan  # noqa: E402,F401

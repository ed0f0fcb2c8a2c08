"""Generate tests/golden/ref_labels.npz: label-side training tensors produced by the UNMODIFIED reference dataset
classes on its own test_data/ (sc_*.npy + pedestrians_*.json):

  ScanContextPedestrianDataset.__getitem__  distro_model/tiny_sc_pred_v2.py:38-98   -> 'mask', 'weight'
  ScanContextPedDataset.__getitem__         distro_model/train_sc_ped.py:154-211    -> 'y_heat', 'y_mask' (augment off)

plus one synthetic pedestrian list that exercises out-of-grid cells, negative / wrapped sectors, overlapping boxes and
centres outside the grid.

    python -m oracle.make_golden_labels
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_shim as ref  # noqa: E402

SCANS = ["000840", "002073", "002095", "002160"]


def synthetic_peds():
    rng = np.random.default_rng(20261019)
    peds = []
    for k in range(12):
        r0, s0 = int(rng.integers(-5, 170)), int(rng.integers(-30, 760))
        n = int(rng.integers(0, 160))
        cells = [[int(r0 + rng.integers(-4, 5)), int(s0 + rng.integers(-9, 10))] for _ in range(n)]
        peds.append({"center_ring_idx": r0, "center_sector_idx": s0, "covered_cells": cells,
                     "distance": float(rng.uniform(0.5, 45.0))})
    peds.append({"center_ring_idx": 159, "center_sector_idx": 719, "covered_cells": [[159, 719], [0, 0], [159, 0]]})   # no 'distance'
    peds.append({"center_ring_idx": 0, "center_sector_idx": 0, "covered_cells": [], "distance": 3.0})
    peds.append({"center_ring_idx": 80, "center_sector_idx": 360, "covered_cells": [[80, 360]] * 3, "distance": 39.9})
    return peds


def main():
    assert ref.available(), "needs /root/reference"
    ref._install_stubs()
    dm = os.path.join(ref.REF_ROOT, "distro_model")
    with contextlib.redirect_stdout(io.StringIO()):
        v2 = ref._load(os.path.join(dm, "tiny_sc_pred_v2.py"), "_ref_tiny_sc_pred_v2")
        tr = ref._load(os.path.join(dm, "train_sc_ped.py"), "_ref_train_sc_ped")
    td = os.path.join(ref.REF_ROOT, "test_data")
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for s in SCANS:
            shutil.copyfile(os.path.join(td, f"sc_{s}.npy"), os.path.join(d, f"sc_{s}.npy"))
            shutil.copyfile(os.path.join(td, f"pedestrians_{s}.json"), os.path.join(d, f"pedestrians_{s}.json"))
        # the synthetic frame
        np.save(os.path.join(d, "sc_999001.npy"), np.zeros((160, 720), np.float32))
        with open(os.path.join(d, "pedestrians_999001.json"), "w") as f:
            json.dump(synthetic_peds(), f)
        ids = [int(s) for s in SCANS] + [999001]
        with contextlib.redirect_stdout(io.StringIO()):
            ds = v2.ScanContextPedestrianDataset(d, ids)
        assert len(ds) == len(ids)
        for i, sid in enumerate(ids):
            it = ds[i]
            out[f"v2_{sid:06d}__mask"] = it["mask"].numpy()[0]
            out[f"v2_{sid:06d}__weight"] = it["weight"].numpy()[0]
        ds2 = tr.ScanContextPedDataset(d, augment=False, use_mask=True)
        for i, p in enumerate(ds2.sc_files):
            sid = os.path.basename(p)[3:9]
            it = ds2[i]
            out[f"tr_{sid}__heat"] = it["y_heat"].numpy()[0]
            out[f"tr_{sid}__mask"] = it["y_mask"].numpy()[0]
        out["sigma"] = np.array(tr.CFG.heat_sigma)
    out["ids"] = np.array([f"{i:06d}" for i in ids])
    np.savez_compressed(os.path.join(OUT, "ref_labels.npz"), **out)
    with open(os.path.join(OUT, "pedestrians_999001.json"), "w") as f:
        json.dump(synthetic_peds(), f)
    print("wrote ref_labels.npz", os.path.getsize(os.path.join(OUT, "ref_labels.npz")) / 1e6, "MB",
          {k: (v.dtype, v.shape) for k, v in out.items() if k.startswith("tr_000840") or k.startswith("v2_000840")})


if __name__ == "__main__":
    main()

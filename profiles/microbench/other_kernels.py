"""Device-time measurements of K2 (box rasteriser), K3 (reconstructor), K4 (to_tensor01) and K5 (covered mask)
against their algorithmic bytes.   python profiles/microbench/other_kernels.py [out.json]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from distropointclouds_b200.engine import Engine, ScanContextParams  # noqa: E402
from distropointclouds_b200.synth import synth_batch_device  # noqa: E402

R, S = 160, 720
dev = torch.device("cuda", 0)
eng = Engine(0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
st = torch.cuda.current_stream(dev)
res = {}


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


B = 2048
pts = synth_batch_device(B, 120000, 64, seed=0, device=str(dev))
offs = np.arange(B + 1, dtype=np.int64) * pts.shape[1]
grids = torch.empty((B, R, S), dtype=torch.float32, device=dev)
eng.build_device(pts, offs, ScanContextParams(), grids, None, None, 16, st.cuda_stream)
torch.cuda.synchronize()
del pts

# K4
norm = torch.empty_like(grids)
ms = timed(lambda: eng.normalize01_device(grids, norm, st.cuda_stream))
alg = B * 2 * 4 * R * S
res["K4 to_tensor01"] = {"ms": ms, "grids": B, "alg_bytes": alg, "GB/s": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / peak}

# K3
cap = B * R * S
po = torch.empty(B + 1, dtype=torch.int64, device=dev)
pout = torch.empty((cap, 3), dtype=torch.float64, device=dev)
ids = torch.empty(cap, dtype=torch.int32, device=dev)
ms = timed(lambda: eng.reconstruct_device(grids, po, pout, ids, R, S, 20.0, True, 2.0, st.cuda_stream))
K = int(po[-1])
alg = B * 4 * R * S + K * 28
res["K3 reconstruct"] = {"ms": ms, "grids": B, "points_out": K, "alg_bytes": alg, "GB/s": alg / ms / 1e6,
                         "frac_of_hbm_peak": alg / ms / 1e6 / peak}
del pout, ids

# K2 (host-buffer API: boxes up, cells down; the device work is the interesting part -> use last_kernel timing off)
rng = np.random.default_rng(1)
nb = 4096
boxes = np.stack([rng.uniform(-18, 18, nb), rng.uniform(-18, 18, nb), rng.uniform(0.4, 1.2, nb), rng.uniform(0.4, 1.0, nb),
                  rng.uniform(-np.pi, np.pi, nb)], axis=1)
import time
eng.raster_boxes_host(boxes)
t0 = time.perf_counter()
for _ in range(5):
    cells, _ = eng.raster_boxes_host(boxes)
dt = (time.perf_counter() - t0) / 5
res["K2 raster (host buffers, incl. PCIe + CSR split in Python)"] = {"ms": dt * 1e3, "boxes": nb, "boxes_per_s": nb / dt,
                                                                       "cells_out": int(sum(len(c) for c in cells))}

# K5
g0 = grids[0].cpu().numpy()
recs = eng.reconstruct_host(g0)
idsh = recs[0][1]
cellsh = np.concatenate([c for c in cells[:64] if len(c)])
eng.covered_mask_host(idsh, cellsh)
t0 = time.perf_counter()
for _ in range(20):
    eng.covered_mask_host(idsh, cellsh)
res["K5 covered mask (host buffers)"] = {"ms": (time.perf_counter() - t0) / 20 * 1e3, "ids": int(len(idsh)), "cells": int(len(cellsh))}

for k, v in res.items():
    print(k, json.dumps(v))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)

"""K3 alone on 2048 real-density grids (for ncu):  python profiles/microbench/k3_only.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from distropointclouds_b200.engine import Engine, ScanContextParams  # noqa: E402
from distropointclouds_b200.synth import synth_batch_device  # noqa: E402

R, S, B = 160, 720, 2048
dev = torch.device("cuda", 0)
eng = Engine(0)
st = torch.cuda.current_stream(dev)
pts = synth_batch_device(B, 120000, 64, seed=0, device=str(dev))
offs = np.arange(B + 1, dtype=np.int64) * pts.shape[1]
grids = torch.empty((B, R, S), dtype=torch.float32, device=dev)
eng.build_device(pts, offs, ScanContextParams(), grids, None, None, 16, st.cuda_stream)
torch.cuda.synchronize()
del pts
cap = B * R * S
po = torch.empty(B + 1, dtype=torch.int64, device=dev)
pout = torch.empty((cap, 3), dtype=torch.float64, device=dev)
ids = torch.empty(cap, dtype=torch.int32, device=dev)
for _ in range(4):
    eng.reconstruct_device(grids, po, pout, ids, R, S, 20.0, True, 2.0, st.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(10):
    eng.reconstruct_device(grids, po, pout, ids, R, S, 20.0, True, 2.0, st.cuda_stream)
e1.record(st)
torch.cuda.synchronize()
print("K3 ms", e0.elapsed_time(e1) / 10, "points", int(po[-1]))
eng.close()

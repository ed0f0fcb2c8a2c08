"""K1 on the four real KITTI scans of tests/golden (replicated to a 4096-scan ragged batch, ~7.8 GB): device time and
algorithmic GB/s, checked bit-exact against the golden grids.   python profiles/microbench/real_scans.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from distropointclouds_b200.engine import Engine, ScanContextParams  # noqa: E402

dev = torch.device("cuda", 0)
z = np.load(os.path.join(ROOT, "tests", "golden", "scans.npz"))
gold = np.load(os.path.join(ROOT, "tests", "golden", "sc_golden.npz"))
ids = ["000840", "002073", "002095", "002160"]
scans = [torch.from_numpy(np.ascontiguousarray(z["s" + i].T)).to(dev) for i in ids]
B = 4096
pts = torch.cat([scans[i % 4] for i in range(B)])
offs = np.concatenate([[0], np.cumsum([scans[i % 4].shape[0] for i in range(B)])]).astype(np.int64)
out = torch.empty((B, 160, 720), dtype=torch.float32, device=dev)
eng = Engine(0)
st = torch.cuda.current_stream(dev)
f = lambda: eng.build_device(pts, offs, ScanContextParams(), out, None, None, 16, st.cuda_stream)
for _ in range(3):
    f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(10):
    f()
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
ok = all(np.array_equal(out[i].cpu().numpy(), gold["sc" + ids[i % 4]]) for i in (0, 1, 2, 3, 4093, 4094, 4095))
alg = 16 * int(offs[-1]) + B * 4 * 160 * 720
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(json.dumps({"workload": "4 real KITTI scans x 1024 (ragged, 117-126 k points)", "ms": ms, "scans_per_s": B / ms * 1e3,
                  "mpoints_per_s": int(offs[-1]) / ms / 1e3, "alg_GBps": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / peak,
                  "bit_exact_vs_golden": ok}))

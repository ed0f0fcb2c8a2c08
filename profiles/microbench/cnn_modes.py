"""A/B of ways to run the reference's TinyCNN (stock PyTorch/cuDNN) on B200: dtype, memory format, BN folding,
chunk size.  Prints ms per scan.   python profiles/microbench/cnn_modes.py"""
import itertools
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from distropointclouds_b200.cnn import build_tiny_cnn, fold_batchnorm  # noqa: E402

dev = torch.device("cuda", 0)
sd = torch.load(os.path.join(ROOT, "tests", "golden", "tinycnn_best.pth"), map_location="cpu", weights_only=True)
x_all = torch.rand(256, 1, 160, 720, device=dev)


def run(fold, dtype, cl, tf32, chunk, bench):
    torch.backends.cudnn.benchmark = bench
    torch.backends.cudnn.allow_tf32 = tf32
    m = build_tiny_cnn()
    m.load_state_dict(sd)
    m.eval()
    if fold:
        m = fold_batchnorm(m)
    m = m.to(dev).to(dtype)
    if cl:
        m = m.to(memory_format=torch.channels_last)
    x = x_all.to(dtype)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        for _ in range(2):
            for a in range(0, x.shape[0], chunk):
                y = m(x[a:a + chunk])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for a in range(0, x.shape[0], chunk):
            y = m(x[a:a + chunk])
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / x.shape[0] * 1e3, y


ref = None
for fold, (dtype, tf32), cl, chunk, bench in itertools.product(
        (False, True), ((torch.float32, False), (torch.float32, True), (torch.bfloat16, True), (torch.float16, True)),
        (False, True), (16, 64), (False, True)):
    try:
        ms, y = run(fold, dtype, cl, tf32, chunk, bench)
        if ref is None:
            ref = y.float()
        err = float((y.float() - ref).abs().max())
        print(f"fold={int(fold)} dtype={str(dtype)[6:]:9s} tf32={int(tf32)} channels_last={int(cl)} chunk={chunk:3d} "
              f"cudnn.benchmark={int(bench)}: {ms:.4f} ms/scan  ({1e3 / ms:.0f} scans/s)  max|dlogit| vs first {err:.2e}", flush=True)
    except Exception as e:  # noqa: BLE001
        print("failed", fold, dtype, cl, chunk, bench, repr(e)[:120], flush=True)

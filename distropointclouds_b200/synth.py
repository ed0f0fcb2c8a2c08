"""Synthetic LiDAR workloads of SURVEY 8(d): 64-beam ~120k-point street-canyon scans (config 3) and
128-beam ~2M-point dense scans (config 5), beam-major point order, mm-quantised like KITTI files."""
from __future__ import annotations

import numpy as np


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


def synth_batch_device(n_scans: int, n_points: int = 120000, beams: int = 64, seed: int = 0, n_base: int = 16,
                       device: str = "cuda"):
    """A device-resident batch (n_scans, n_points, 4) float32 built from `n_base` CPU-generated scans by
    per-scan random yaw rotation + re-quantisation to 1 mm on the GPU (torch used as plumbing only: this
    is workload generation, outside any timed region)."""
    import torch
    base = np.stack([synth_scan(seed * 1000 + i, n_points, beams) for i in range(n_base)])
    n_pts = base.shape[1]
    base_d = torch.from_numpy(base).to(device)
    out = torch.empty((n_scans, n_pts, 4), dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(20251018 + seed)
    yaw = torch.rand(n_scans, generator=g, device=device, dtype=torch.float64) * (2 * np.pi)
    for s0 in range(0, n_scans, 128):
        s1 = min(n_scans, s0 + 128)
        idx = torch.arange(s0, s1, device=device) % n_base
        b = base_d[idx].to(torch.float64)
        c, s = torch.cos(yaw[s0:s1])[:, None], torch.sin(yaw[s0:s1])[:, None]
        x = c * b[:, :, 0] - s * b[:, :, 1]
        y = s * b[:, :, 0] + c * b[:, :, 1]
        out[s0:s1, :, 0] = (torch.round(x * 1000) / 1000).to(torch.float32)
        out[s0:s1, :, 1] = (torch.round(y * 1000) / 1000).to(torch.float32)
        out[s0:s1, :, 2] = b[:, :, 2].to(torch.float32)
        out[s0:s1, :, 3] = b[:, :, 3].to(torch.float32)
    return out

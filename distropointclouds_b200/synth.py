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


def adversarial_batch_device(kind: str, n_scans: int, n_points: int = 120000, seed: int = 0, device: str = "cuda"):
    """The two adversarial distributions of SURVEY 8(d), generated on the device (workload generation, outside any
    timed region):  'onecell'   every point of a scan falls into ONE grid cell (maximum atomic contention);
                    'uniform80' uniform over the 80 m disc (~94 % of the points beyond 20 m: clamp-ring pile-up)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(20251018 + 7919 * seed + (1 if kind == "onecell" else 2))
    out = torch.empty((n_scans, n_points, 4), dtype=torch.float32, device=device)
    for s0 in range(0, n_scans, 256):
        s1 = min(n_scans, s0 + 256)
        shape = (s1 - s0, n_points)
        u = torch.rand(shape, generator=g, device=device, dtype=torch.float64)
        v = torch.rand(shape, generator=g, device=device, dtype=torch.float64)
        w = torch.rand(shape, generator=g, device=device, dtype=torch.float64)
        if kind == "onecell":
            # cell (ring 40, sector 90): r in [5.0, 5.125), theta in [45.0, 45.5) degrees, well inside
            r = 5.02 + 0.08 * u
            th = np.deg2rad(45.05) + np.deg2rad(0.4) * v
        elif kind == "uniform80":
            r = 80.0 * torch.sqrt(u)
            th = 2 * np.pi * v
        else:
            raise ValueError(kind)
        out[s0:s1, :, 0] = (torch.round(r * torch.cos(th) * 1000) / 1000).to(torch.float32)
        out[s0:s1, :, 1] = (torch.round(r * torch.sin(th) * 1000) / 1000).to(torch.float32)
        out[s0:s1, :, 2] = (torch.round((w * 3.0 - 2.0) * 1000) / 1000).to(torch.float32)
        out[s0:s1, :, 3] = 0.5
    return out


def adversarial_scan(kind: str, seed: int, n: int = 120000):
    """Host (NumPy) flavour of the adversarial distributions, for the CPU arm."""
    rng = np.random.default_rng(20251018 + 7919 * seed + (1 if kind == "onecell" else 2))
    u, v, w = rng.random(n), rng.random(n), rng.random(n)
    if kind == "onecell":
        r, th = 5.02 + 0.08 * u, np.deg2rad(45.05) + np.deg2rad(0.4) * v
    elif kind == "uniform80":
        r, th = 80.0 * np.sqrt(u), 2 * np.pi * v
    else:
        raise ValueError(kind)
    pts = np.stack([np.round(r * np.cos(th) * 1000) / 1000, np.round(r * np.sin(th) * 1000) / 1000,
                    np.round((w * 3.0 - 2.0) * 1000) / 1000, np.full(n, 0.5)], axis=1)
    return pts.astype(np.float32)

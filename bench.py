#!/usr/bin/env python
"""bench.py -- scan-context build throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scans B] [--points P]

A "step" is one pass of the hot path (K1: packed float32 x,y,z,i scans -> 160x720 scan-context grids)
over one batch of synthetic 64-beam ~120k-point scans (BASELINE configs[2], batch 4096 per GPU; weak
scaling: every rank owns its own batch, there is no data-path collective -- SURVEY 8e).

  value      whole-job scans/s with the batch resident in HBM when the timed region starts
  e2e        the same metric through the C-ABI call with HOST (pinned) buffers: H2D of the points and
             D2H of the grids are inside the timed region
  roofline   algorithmic bytes (16*N + 4*R*S per scan) / device time of the step, against the
             measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline   the oracle's C port of the reference algorithm on the host cores (bounded sample)

--impl reference times that CPU port alone (all host threads) on bounded samples of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scan_context_build_scans_per_s"
R, S, MAXLEN = 160, 720, 20.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # median of the upper half: samples taken under load
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def run_reference(args):
    """The reference algorithm's CPU port (oracle/dsc_oracle.c) with every host thread, on bounded samples."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import oracle as orc
    from distropointclouds_b200.synth import synth_scan
    cores = orc.max_threads()
    sample = min(args.scans, max(cores * 8, 256))
    base = [synth_scan(i, args.points) for i in range(16)]
    pts = np.concatenate([base[i % 16] for i in range(sample)])
    offs = (np.arange(sample + 1, dtype=np.int64) * len(base[0]))
    p = orc.Params(R, S, MAXLEN)
    for _ in range(max(1, min(args.warmup, 1))):
        orc.sc_build_batch(pts[: offs[cores]], offs[: cores + 1], p, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.sc_build_batch(pts, offs, p, n_threads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"synthetic 64-beam {args.points}-point scans, batch {args.scans} per GPU (BASELINE configs[2])",
                   "num_ring": R, "num_sector": S, "max_length": MAXLEN, "variant": "A (scaner_version2.py:64-115)"},
        "mpoints_per_s": v * args.points / 1e6,
        "cpu_baseline": {"value": v, "unit": "scans/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} scans of the workload per step, C port of the reference algorithm "
                                   f"(oracle/dsc_oracle.c), {cores} pthreads"},
        "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from distropointclouds_b200.engine import Engine, ScanContextParams
    from distropointclouds_b200.synth import synth_batch_device

    rank, local_rank, world = dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    eng = Engine(local_rank)
    eng.set_option("k1_impl", args.k1_impl)
    eng.set_option("hg_cfg", args.hg_cfg)
    if args.pf_dist >= 0:
        eng.set_option("pf_dist", args.pf_dist)
    eng.set_option("st_cfg", args.st_cfg)
    eng.set_option("sc_cfg", args.sc_cfg)
    if args.l2_hints >= 0:
        eng.set_option("l2_hints", args.l2_hints)
    eng.set_option("l2_persist", args.l2_persist)
    eng.set_option("no_defgrid", args.no_defgrid)
    eng.set_option("sc_smem_rows", args.sc_smem_rows)
    if args.st_slots > 0:
        eng.set_option("st_slots", args.st_slots)
    params = ScanContextParams(R, S, MAXLEN)
    B, Pn = args.scans, args.points

    pts = synth_batch_device(B, Pn, 64, seed=rank, device=str(dev))          # (B, P, 4) resident in HBM
    Pn = pts.shape[1]
    offs = np.arange(B + 1, dtype=np.int64) * Pn
    out = torch.empty((B, R, S), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    alg_bytes = B * (16 * Pn + 4 * R * S)
    l2_mb = torch.cuda.get_device_properties(dev).L2_cache_size / 2 ** 20

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step():
        eng.build_device(pts, offs, params, out, None, None, 16, stream.cuda_stream)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = eng.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = eng.kernel_launches - n0
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host (pinned) buffers through the C ABI, H2D + D2H inside the timed region -------------------
    e2e_steps = max(1, min(args.steps, 3))
    e2e_s, ok, h_pts, h_out = float("nan"), None, None, None
    if not args.skip_e2e:
        h_pts = eng.pinned_empty((B * Pn, 4), np.float32)
        h_out = eng.pinned_empty((B, R, S), np.float32)
        torch.from_numpy(h_pts).copy_(pts.view(-1, 4))
        torch.cuda.synchronize(dev)
        eng.build_host(h_pts, offs, params, 16, want_stats=False, out=h_out)        # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.build_host(h_pts, offs, params, 16, want_stats=False, out=h_out)
        torch.cuda.synchronize(dev)
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        ok = bool(np.array_equal(h_out[:4], out[:4].cpu().numpy()))

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])

    if rank == 0:
        peak, peak_src = peaks()
        value = world * B / (ms_max * 1e-3)
        achieved = alg_bytes / (ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic 64-beam {Pn}-point scans, batch {B} per GPU (BASELINE configs[2])",
                       "num_ring": R, "num_sector": S, "max_length": MAXLEN, "variant": "A (scaner_version2.py:64-115)",
                       "l2_policy": f"inputs larger than L2: {B * Pn * 16 / 2**20:.0f} MiB of points per step vs {l2_mb:.0f} MiB L2",
                       "points_per_scan": Pn},
            "mpoints_per_s": value * Pn / 1e6,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_step": alg_bytes,
                         "note": "achieved = (16*N + 4*R*S)*B / device time of the whole step (all engine kernels) on rank 0"},
            "e2e": {"value": world * B / (e2e_ms_max * 1e-3), "unit": "scans/s", "h2d_bytes_per_step": int(B * Pn * 16),
                    "d2h_bytes_per_step": int(B * R * S * 4), "steps": e2e_steps, "host_memory": "pinned",
                    "matches_device_path": ok},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if args.skip_e2e:
            line["e2e"] = None
        line["config"]["k1_impl"] = {0: "auto", 1: "l2keys", 2: "halfgrid", 3: "scan", 4: "tile"}[args.k1_impl]
        if world == 1 and not args.no_cpu_baseline and h_pts is not None:
            from oracle import oracle as orc
            cores = orc.max_threads()
            sample = min(B, 1024)
            hp = h_pts[: sample * Pn]
            t0 = time.perf_counter()
            ref, _ = orc.sc_build_batch(hp, offs[: sample + 1], orc.Params(R, S, MAXLEN), n_threads=cores)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample / dt, "unit": "scans/s", "cores": cores, "kind": "port",
                                    "sample": f"first {sample} scans of the step's batch, C port of the reference "
                                              f"algorithm (oracle/dsc_oracle.c), {cores} pthreads, {dt:.1f} s",
                                    "matches_gpu": bool(np.array_equal(ref, h_out[:sample]))}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scans", type=int, default=4096, help="scans per GPU per step")
    ap.add_argument("--points", type=int, default=120000, help="points per scan")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--l2-hints", type=int, default=-1, help="L2 eviction hints (A/B runs)")
    ap.add_argument("--sc-smem-rows", type=int, default=-1, help="scan kernel: resident ring rows (A/B runs)")
    ap.add_argument("--no-defgrid", type=int, default=0, help="disable the default-grid specialisation (A/B runs)")
    ap.add_argument("--l2-persist", type=int, default=0, help="L2 persisting window for key slots (A/B runs)")
    ap.add_argument("--sc-cfg", type=int, default=0, help="scan kernel shape (A/B runs)")
    ap.add_argument("--st-cfg", type=int, default=0, help="streaming kernel shape (A/B runs)")
    ap.add_argument("--st-slots", type=int, default=0, help="key-grid slots of the streaming kernel (A/B runs)")
    ap.add_argument("--pf-dist", type=int, default=-1, help="bulk L2 prefetch distance in steps (A/B runs)")
    ap.add_argument("--hg-cfg", type=int, default=0, help="half-grid launch shape (A/B runs)")
    ap.add_argument("--k1-impl", type=int, default=0, help="0 streaming kernel (default), 1 chunked L2-key kernels, 2 half-grid kernel (A/B runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- scan-context build throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload kitti120k|dense2m|cnn|onecell|uniform80] [--scans B] [--points P]

A "step" is one pass of the hot path over one batch of synthetic scans; every rank owns its own batch
(weak scaling, no data-path collective -- SURVEY 8e).

  kitti120k  (default; BASELINE configs[2], the configuration the metric is quoted on)
             K1: 4096 packed 64-beam 120 000-point scans per GPU -> 160x720 scan-context grids
  dense2m    (configs[4]) K1 over 64 128-beam 2 097 152-point scans per GPU, then the K3 inverse
             (grid -> points) on the grids just produced
  cnn        (configs[3]) K1 + K4 (to_tensor01) -> the reference's TinyCNN (tinycnn_best.pth) -> masks,
             1024 scans per GPU, grids never leave the device

  value      whole-job scans/s with the batch resident in HBM when the timed region starts (CUDA events,
             max over ranks)
  e2e        the same metric through the host-buffer call (C ABI, DSC_MEM_HOST) from pinned host memory:
             H2D of the points and D2H of the results are inside the timed region
  roofline   the K1 launch: algorithmic bytes (16*N + 4*R*S per scan) / its duration measured
             with CUDA events around each launch; `traffic` is the ncu dram__bytes of the same launch
             (profiles/k1_traffic.json, from the committed `ncu --set full` capture)
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
WORKLOADS = {
    # name: (scans per GPU, points per scan, beams, n_base, description)
    "kitti120k": (4096, 120000, 64, 16, "synthetic 64-beam {P}-point scans, batch {B} per GPU (BASELINE configs[2])"),
    "dense2m": (64, 2097152, 128, 4, "synthetic 128-beam dense {P}-point scans, batch {B} per GPU, K1 build + K3 "
                                      "inverse (BASELINE configs[4])"),
    "onecell": (4096, 120000, 1, 0, "adversarial: all {P} points of a scan in ONE cell, batch {B} per GPU (SURVEY 8d)"),
    "uniform80": (4096, 120000, 1, 0, "adversarial: {P} points uniform over the 80 m disc (94 % clamp into the last ring), "
                                       "batch {B} per GPU (SURVEY 8d)"),
    "cnn": (1024, 120000, 64, 16, "synthetic 64-beam {P}-point scans, batch {B} per GPU, K1 + K4 -> TinyCNN "
                                   "(tinycnn_best.pth) -> masks (BASELINE configs[3])"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, B, Pn):
    """dram bytes per K1 launch from the committed ncu --set full capture of this very workload, else None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))[workload]
        if t["scans"] == B and t["points"] == Pn:
            return {"bytes": t["dram_bytes_read"] + t["dram_bytes_write"], "source": t["source"]}
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi sampled every 100 ms during the timed region (B200_PROFILING.md clocks line)."""
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
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            time.sleep(0.3)
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
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
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []     # upper half: samples taken under load
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def workload_cfg(args):
    B, Pn, beams, n_base, desc = WORKLOADS[args.workload]
    B = args.scans or B
    Pn = args.points or Pn
    Pn = (Pn // beams) * beams
    return B, Pn, beams, n_base, desc.format(B=B, P=Pn)


def base_line(args, world, desc, Pn):
    return {"metric": METRIC, "unit": "scans/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "num_ring": R, "num_sector": S, "max_length": MAXLEN,
                       "variant": "A (scaner_version2.py:64-115)", "points_per_scan": Pn}}


# ---- reference arm ---------------------------------------------------------------------------------------

def run_reference(args):
    """The reference algorithm's CPU port (oracle/dsc_oracle.c) with every host thread, on bounded samples."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import oracle as orc
    from distropointclouds_b200.synth import adversarial_scan, synth_scan
    B, Pn, beams, n_base, desc = workload_cfg(args)
    cores = orc.max_threads()
    # ~256 scans of 120k points per step (~0.1 s on 16 threads); the same number of points for the dense workload
    sample = max(cores, min(B, int(256 * 120000 / Pn)))
    if args.workload in ("onecell", "uniform80"):
        base = [adversarial_scan(args.workload, i, Pn) for i in range(4)]
    else:
        base = [synth_scan(i, Pn, beams) for i in range(min(n_base, 4))]
    pts = np.concatenate([base[i % len(base)] for i in range(sample)])
    offs = np.arange(sample + 1, dtype=np.int64) * len(base[0])
    p = orc.Params(R, S, MAXLEN)

    def step():
        grids, _ = orc.sc_build_batch(pts, offs, p, n_threads=cores)
        if args.workload == "dense2m":
            for g in grids[: max(1, sample)]:
                orc.reconstruct(g)
        return grids

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = sample / dt
    line = base_line(args, args.gpus, desc, Pn)
    line.update({"impl": "reference", "value": v, "ms_per_step": dt * 1e3, "warmup": args.warmup,
                 "mpoints_per_s": v * Pn / 1e6,
                 "cpu_baseline": {"value": v, "unit": "scans/s", "cores": cores, "kind": "port",
                                  "sample": f"{sample} scans of the workload per step, C port of the reference "
                                            f"algorithm (oracle/dsc_oracle.c), {cores} pthreads"},
                 "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    if args.workload == "cnn":
        line["cpu_baseline"]["sample"] += "; grids only (the CNN stage is the reference's own PyTorch module on both arms)"
    print(json.dumps(line), flush=True)


# ---- our arm ---------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    from distropointclouds_b200.engine import Engine, ScanContextParams
    from distropointclouds_b200.synth import adversarial_batch_device, synth_batch_device

    rank, local_rank, world = dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    eng = Engine(local_rank)
    for name in ("k1_impl", "sc_cfg", "no_defgrid", "sc_smem_rows", "sc_split", "sc_item_points", "sc_window"):
        v = getattr(args, name)
        if v is not None:
            eng.set_option(name, v)
    params = ScanContextParams(R, S, MAXLEN)
    B, Pn, beams, n_base, desc = workload_cfg(args)
    wl = args.workload

    if wl in ("onecell", "uniform80"):
        pts = adversarial_batch_device(wl, B, Pn, seed=rank, device=str(dev))
    else:
        pts = synth_batch_device(B, Pn, beams, seed=rank, n_base=n_base, device=str(dev))   # (B, P, 4) resident in HBM
    Pn = pts.shape[1]
    offs = np.arange(B + 1, dtype=np.int64) * Pn
    out = torch.empty((B, R, S), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    alg_bytes = B * (16 * Pn + 4 * R * S)
    l2_mb = torch.cuda.get_device_properties(dev).L2_cache_size / 2 ** 20
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if B * Pn * 16 < 4 * l2_mb * 2 ** 20 else None

    # workload-specific device state
    rec = pipe = None
    if wl == "dense2m":
        cap = B * R * S
        rec = (torch.empty(B + 1, dtype=torch.int64, device=dev), torch.empty((cap, 3), dtype=torch.float64, device=dev),
               torch.empty(cap, dtype=torch.int32, device=dev))
    if wl == "cnn":
        from distropointclouds_b200.cnn import ScanContextCnnPipeline
        w = os.path.join(ROOT, "tests", "golden", "tinycnn_best.pth")
        pipe = ScanContextCnnPipeline(w if os.path.exists(w) else None, local_rank, args.cnn_precision, engine=eng)
        norm = torch.empty_like(out)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def k1():
        eng.build_device(pts, offs, params, out, norm if wl == "cnn" else None, None, 16, stream.cuda_stream)

    def step():
        if flush is not None:
            flush.zero_()                       # inputs are not >> L2: evict them between iterations
        k1()
        if wl == "dense2m":
            eng.reconstruct_device(out, rec[0], rec[1], rec[2], R, S, MAXLEN, True, 2.0, stream.cuda_stream)
        elif wl == "cnn":
            logits = pipe.logits_from_norm(norm)
            return (torch.sigmoid(logits) > 0.5).to(torch.uint8)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = eng.kernel_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 2)]
    barrier()
    ev[-2].record(stream)
    for i in range(args.steps):
        step()
    ev[-1].record(stream)
    barrier()
    ms = ev[-2].elapsed_time(ev[-1]) / args.steps
    launches = eng.kernel_launches - n0
    # the K1 launch alone, CUDA events around each launch on the launching stream (roofline.achieved)
    for i in range(args.steps):
        if flush is not None:
            flush.zero_()
        ev[2 * i].record(stream)
        k1()
        ev[2 * i + 1].record(stream)
    barrier()
    k1_ms = float(np.mean([ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.steps)]))
    clocks = sampler.stop() if rank == 0 else None
    if wl == "cnn":
        k1_ms_note = "K1 + K4 (the build call with out_norm)"
        alg_k1 = alg_bytes + B * 2 * 4 * R * S
    else:
        k1_ms_note = "K1"
        alg_k1 = alg_bytes

    # ---- e2e: host (pinned) buffers through the public host-buffer call, H2D + D2H inside the timed region -----
    e2e_steps = max(1, min(args.steps, 3))
    e2e_s, ok, h_pts, h_out, d2h = float("nan"), None, None, None, 0
    if not args.skip_e2e:
        h_pts = eng.pinned_empty((B * Pn, 4), np.float32)
        torch.from_numpy(h_pts).copy_(pts.view(-1, 4))
        torch.cuda.synchronize(dev)
        if wl in ("kitti120k", "onecell", "uniform80"):
            h_out = eng.pinned_empty((B, R, S), np.float32)
            d2h = B * R * S * 4

            def e2e_step():
                eng.build_host(h_pts, offs, params, 16, want_stats=False, out=h_out)
        elif wl == "dense2m":
            h_out = eng.pinned_empty((B, R, S), np.float32)

            def e2e_step():
                eng.build_host(h_pts, offs, params, 16, want_stats=False, out=h_out)
                return eng.reconstruct_host(h_out, R, S, MAXLEN)
        else:
            h_mask = torch.empty((B, R, S), dtype=torch.uint8).pin_memory()
            d_stage = torch.empty_like(pts)
            d2h = B * R * S

            def e2e_step():
                d_stage.view(-1, 4).copy_(torch.from_numpy(h_pts), non_blocking=True)
                _, _, m = pipe.infer_device(d_stage, offs, params)
                h_mask.copy_(m, non_blocking=True)
                torch.cuda.synchronize(dev)
        r = e2e_step()                                           # warm-up (allocates staging)
        if wl == "dense2m":
            d2h = B * R * S * 4 + sum(p.nbytes + i.nbytes for p, i in r)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize(dev)
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        if h_out is not None:
            ok = bool(np.array_equal(h_out[:4], out[:4].cpu().numpy()))

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])

    if rank == 0:
        peak, peak_src = peaks()
        value = world * B / (ms_max * 1e-3)
        achieved = alg_k1 / (k1_ms * 1e-3) / 1e9
        traffic = ncu_traffic(wl, B, Pn)
        line = base_line(args, world, desc, Pn)
        line["config"]["l2_policy"] = (f"inputs larger than L2: {B * Pn * 16 / 2**20:.0f} MiB of points per step vs "
                                       f"{l2_mb:.0f} MiB L2" if flush is None else "256 MiB flush buffer written between steps")
        line["config"]["k1_impl"] = {None: "auto", 0: "auto", 1: "l2keys", 3: "scan"}[args.k1_impl]
        if wl == "cnn":
            line["config"]["cnn_precision"] = args.cnn_precision
        line.update({
            "value": value, "ms_per_step": ms_max, "mpoints_per_s": value * Pn / 1e6,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic["bytes"] if traffic else None, "peak_source": peak_src,
                         "kernel": k1_ms_note, "kernel_ms": k1_ms, "kernel_share_of_step": k1_ms / ms,
                         "algorithmic_bytes_per_launch": alg_k1,
                         "traffic_source": traffic["source"] if traffic else None,
                         "note": "achieved = algorithmic bytes of one launch / its duration (CUDA events around each "
                                 "launch, mean over the steps, rank 0)"},
            "e2e": None if args.skip_e2e else {
                "value": world * B / (e2e_ms_max * 1e-3), "unit": "scans/s", "h2d_bytes_per_step": int(B * Pn * 16),
                "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "host_memory": "pinned", "matches_device_path": ok},
            "gpu_launches": int(launches),
            "clocks": clocks,
        })
        if world == 1 and not args.no_cpu_baseline and h_pts is not None:
            from oracle import oracle as orc
            cores = orc.max_threads()
            sample = max(1, min(B, int(1024 * 120000 / Pn)))
            hp = h_pts[: sample * Pn]
            t0 = time.perf_counter()
            ref, _ = orc.sc_build_batch(hp, offs[: sample + 1], orc.Params(R, S, MAXLEN), n_threads=cores)
            if wl == "dense2m":
                for g in ref:
                    orc.reconstruct(g)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample / dt, "unit": "scans/s", "cores": cores, "kind": "port",
                                    "sample": f"first {sample} scans of the step's batch, C port of the reference "
                                              f"algorithm (oracle/dsc_oracle.c), {cores} pthreads, {dt:.1f} s",
                                    "matches_gpu": bool(np.array_equal(ref, out[:sample].cpu().numpy()))}
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
    ap.add_argument("--workload", default="kitti120k", choices=sorted(WORKLOADS))
    ap.add_argument("--scans", type=int, default=0, help="scans per GPU per step (0: the workload's)")
    ap.add_argument("--points", type=int, default=0, help="points per scan (0: the workload's)")
    ap.add_argument("--cnn-precision", default="tf32", choices=["fp32", "tf32", "fp16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    # A/B knobs of the engine (see dsc_api.cu); None = engine default
    for name in ("k1-impl", "sc-cfg", "no-defgrid", "sc-smem-rows", "sc-split", "sc-item-points", "sc-window"):
        ap.add_argument("--" + name, type=int, default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

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

The ONE JSON line of the default invocation is the kitti120k measurement; short dense2m and cnn legs run after it in
the same process and are attached under "extra" (so configs[3] and configs[4] show up in the driver's records at
every N).

  value      whole-job scans/s with the batch resident in HBM when the timed region starts (CUDA events,
             max over ranks)
  e2e        the same metric through the host-buffer call (C ABI, DSC_MEM_HOST) from pinned host memory:
             H2D of the points and D2H of the results are inside the timed region
  roofline   the K1 launch: algorithmic bytes (16*N + 4*R*S per scan) / its duration, measured with
             CUDA events around each K1 call INSIDE the timed loop; `traffic` is the ncu dram__bytes of the same
             launch (profiles/k1_traffic.json, from the committed `ncu --set full` capture)
  cpu_baseline   the oracle's C port of the reference algorithm on the host cores (bounded sample);
             cpu_baseline_numpy: the reference's NumPy path (oracle/restated.py: variant A with its per-point Python
             loop on 1 core and on all cores, variant B np.maximum.at on 1 core), BASELINE.md section 5

--impl reference times the CPU port alone (all host threads) on bounded samples of the same workload.
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


def workload_cfg(name, scans=0, points=0):
    B, Pn, beams, n_base, desc = WORKLOADS[name]
    B = scans or B
    Pn = points or Pn
    Pn = (Pn // beams) * beams
    return B, Pn, beams, n_base, desc.format(B=B, P=Pn)


def base_line(args, world, desc, Pn):
    # `config` is identical on both arms (the driver compares them); engine / run details go under "run"
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
    B, Pn, beams, n_base, desc = workload_cfg(args.workload, args.scans, args.points)
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


def numpy_baseline(h_pts, Pn, budget_s=12.0):
    """the reference's NumPy path (oracle/restated.py) on this box's host cores.
    (1) variant A with its per-point Python loop, one process; (2) the same under Pool(all cores);
    (3) variant B (np.maximum.at), one process.  Bounded: a few scans each."""
    from oracle import restated
    cores = os.cpu_count() or 1
    xyz = [np.ascontiguousarray(h_pts[i * Pn:(i + 1) * Pn, :3]) for i in range(max(2, min(2 * cores, 32)))]
    out = {"numpy": np.__version__, "cores": cores, "unit": "scans/s", "points_per_scan": Pn,
           "source": "oracle/restated.py (restatement of scaner_version2.py:64-115 incl. its Python loop, and of "
                     "scan_context_node.py:102-157); the reference itself cannot travel to the GPU box"}
    t0 = time.perf_counter(); n = 0
    while n < 2 or (time.perf_counter() - t0 < budget_s / 4 and n < len(xyz)):
        restated.variant_a(xyz[n % len(xyz)]); n += 1
    out["variant_a_loop_1core"] = n / (time.perf_counter() - t0)
    try:
        t0 = time.perf_counter()
        restated.variant_a_pool(xyz[: min(len(xyz), 2 * cores)], cores)
        out["variant_a_loop_allcores"] = min(len(xyz), 2 * cores) / (time.perf_counter() - t0)
    except Exception as e:          # noqa: BLE001
        out["variant_a_loop_allcores"] = None
        out["pool_error"] = str(e)[:80]
    t0 = time.perf_counter(); n = 0
    while n < 4 or (time.perf_counter() - t0 < budget_s / 6 and n < 200):
        restated.variant_b(xyz[n % len(xyz)], 80.0); n += 1
    out["variant_b_maximum_at_1core"] = n / (time.perf_counter() - t0)
    return out


# ---- our arm ---------------------------------------------------------------------------------------------

def measure(args, eng, dev, rank, world, wl, scans, points, steps, warmup, do_e2e, sampler=None):
    """One workload on this rank: returns a dict of rank-local numbers (times in ms) + workload facts."""
    import torch
    import torch.distributed as dist
    from distropointclouds_b200.engine import ScanContextParams
    from distropointclouds_b200.synth import adversarial_batch_device, synth_batch_device

    params = ScanContextParams(R, S, MAXLEN)
    B, Pn, beams, n_base, desc = workload_cfg(wl, scans, points)
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

    rec = pipe = norm = None
    if wl == "dense2m":
        cap = B * R * S
        rec = (torch.empty(B + 1, dtype=torch.int64, device=dev), torch.empty((cap, 3), dtype=torch.float64, device=dev),
               torch.empty(cap, dtype=torch.int32, device=dev))
    if wl == "cnn":
        from distropointclouds_b200.cnn import ScanContextCnnPipeline
        w = os.path.join(ROOT, "tests", "golden", "tinycnn_best.pth")
        pipe = ScanContextCnnPipeline(w if os.path.exists(w) else None, dev.index, args.cnn_precision, engine=eng)
        norm = torch.empty_like(out)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    k3ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]

    def step(i=None):
        if flush is not None:
            flush.zero_()                       # inputs are not >> L2: evict them between iterations
        if i is not None:
            kev[i][0].record(stream)
        eng.build_device(pts, offs, params, out, norm, None, 16, stream.cuda_stream)
        if i is not None:
            kev[i][1].record(stream)
        if wl == "dense2m":
            if i is not None:
                k3ev[i][0].record(stream)
            eng.reconstruct_device(out, rec[0], rec[1], rec[2], R, S, MAXLEN, True, 2.0, stream.cuda_stream)
            if i is not None:
                k3ev[i][1].record(stream)
        elif wl == "cnn":
            logits = pipe.logits_from_norm(norm)
            return (torch.sigmoid(logits) > 0.5).to(torch.uint8)

    for _ in range(max(warmup, 3)):
        step()
    barrier()
    if sampler is not None:
        sampler.start()
    n0 = eng.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(steps):
        step(i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / steps
    launches = eng.kernel_launches - n0
    # K1 alone: events around each K1 call inside the timed loop above (so its share of the step is <= 1)
    k1_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    res = {"wl": wl, "B": B, "Pn": Pn, "desc": desc, "ms": ms, "k1_ms": k1_ms, "launches": int(launches),
           "alg_k1": alg_bytes + (B * 2 * 4 * R * S if wl == "cnn" else 0),
           "k1_note": "K1 + K4 (the build call with out_norm)" if wl == "cnn" else "K1",
           "l2_policy": (f"inputs larger than L2: {B * Pn * 16 / 2**20:.0f} MiB of points per step vs {l2_mb:.0f} MiB L2"
                         if flush is None else "256 MiB flush buffer written between steps"),
           "offs": offs, "out": out, "pts": pts}
    if wl == "dense2m":
        k3_ms = float(np.mean([a.elapsed_time(b) for a, b in k3ev]))
        n_pts = int(rec[0][B].item())
        res.update({"k3_ms": k3_ms, "k3_alg": B * 4 * R * S + 28 * n_pts, "k3_points": n_pts})

    # ---- e2e: host (pinned) buffers through the public host-buffer call, H2D + D2H inside the timed region -----
    res.update({"e2e_ms": float("nan"), "e2e_ok": None, "d2h": 0, "e2e_steps": 0, "h_pts": None})
    if do_e2e:
        e2e_steps = max(1, min(steps, 3))
        h_pts = eng.pinned_empty((B * Pn, 4), np.float32)
        torch.from_numpy(h_pts).copy_(pts.view(-1, 4))
        torch.cuda.synchronize(dev)
        h_out, d2h = None, 0
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
        res["e2e_ms"] = (time.perf_counter() - t0) / e2e_steps * 1e3
        res["e2e_steps"], res["d2h"], res["h_pts"] = e2e_steps, d2h, h_pts
        if h_out is not None:
            res["e2e_ok"] = bool(np.array_equal(h_out[:4], out[:4].cpu().numpy()))
    return res


def cpu_port_baseline(res, sample_scans):
    from oracle import oracle as orc
    B, Pn = res["B"], res["Pn"]
    cores = orc.max_threads()
    sample = max(1, min(B, sample_scans))
    hp = res["h_pts"][: sample * Pn] if res["h_pts"] is not None else res["pts"][:sample].reshape(-1, 4).cpu().numpy()
    t0 = time.perf_counter()
    ref, _ = orc.sc_build_batch(hp, res["offs"][: sample + 1], orc.Params(R, S, MAXLEN), n_threads=cores)
    if res["wl"] == "dense2m":
        for g in ref:
            orc.reconstruct(g)
    dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": "scans/s", "cores": cores, "kind": "port",
            "sample": f"first {sample} scans of the step's batch, C port of the reference algorithm "
                      f"(oracle/dsc_oracle.c), {cores} pthreads, {dt:.1f} s",
            "matches_gpu": bool(np.array_equal(ref, res["out"][:sample].cpu().numpy()))}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from distropointclouds_b200.engine import Engine

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
    peak, peak_src = peaks()

    def reduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    sampler = ClockSampler(local_rank) if rank == 0 else None
    wl = args.workload
    res = measure(args, eng, dev, rank, world, wl, args.scans, args.points, args.steps, args.warmup,
                  not args.skip_e2e, sampler)
    clocks = sampler.stop() if sampler is not None else None
    ms_max, e2e_ms_max = reduce_max([res["ms"], res["e2e_ms"]])
    line = None
    if rank == 0:
        B, Pn = res["B"], res["Pn"]
        value = world * B / (ms_max * 1e-3)
        achieved = res["alg_k1"] / (res["k1_ms"] * 1e-3) / 1e9
        traffic = ncu_traffic(wl, B, Pn)
        line = base_line(args, world, res["desc"], Pn)
        line["run"] = {"l2_policy": res["l2_policy"], "k1_impl": {None: "auto", 0: "auto", 1: "l2keys", 3: "scan"}[args.k1_impl]}
        if wl == "cnn":
            line["run"]["cnn_precision"] = args.cnn_precision
        line.update({
            "value": value, "ms_per_step": ms_max, "mpoints_per_s": value * Pn / 1e6,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic["bytes"] if traffic else None, "peak_source": peak_src,
                         "kernel": res["k1_note"], "kernel_ms": res["k1_ms"], "kernel_share_of_step": res["k1_ms"] / res["ms"],
                         "algorithmic_bytes_per_launch": res["alg_k1"],
                         "traffic_source": traffic["source"] if traffic else None,
                         "frac_of_read_stream_ceiling": achieved / 7400.0,
                         "note": "achieved = algorithmic bytes of one launch / its duration (CUDA events around each K1 "
                                 "call inside the timed loop, mean over the steps, rank 0); read-stream ceiling 7.4 TB/s = "
                                 "the loads-only microbenchmark (profiles/r1_r3l_microbench_k1_primitives.txt)"},
            "e2e": None if args.skip_e2e else {
                "value": world * B / (e2e_ms_max * 1e-3), "unit": "scans/s", "h2d_bytes_per_step": int(B * Pn * 16),
                "d2h_bytes_per_step": int(res["d2h"]), "steps": res["e2e_steps"], "host_memory": "pinned",
                "matches_device_path": res["e2e_ok"]},
            "gpu_launches": res["launches"],
            "clocks": clocks,
        })
        if wl == "dense2m":
            a3 = res["k3_alg"] / (res["k3_ms"] * 1e-3) / 1e9
            line["roofline_k3"] = {"kernel": "K3 k3_fused", "kernel_ms": res["k3_ms"], "achieved": a3, "frac": a3 / peak,
                                   "points_out": res["k3_points"]}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_port_baseline(res, int(1024 * 120000 / Pn))
            if wl == "kitti120k" and res["h_pts"] is not None:
                line["cpu_baseline_numpy"] = numpy_baseline(res["h_pts"], Pn)
    del res
    torch.cuda.empty_cache()

    # ---- extra legs: BASELINE configs[4] (dense2m: K1 split items + K3 inverse) and configs[3] (cnn) -----------
    if wl == "kitti120k" and not args.no_extra and not args.scans and not args.points:
        extra = {}
        for name, steps in (("dense2m", 5), ("cnn", 3)):
            try:
                r = measure(args, eng, dev, rank, world, name, 0, 0, steps, 3, False)
                (ms_x,) = reduce_max([r["ms"]])
                if rank == 0:
                    ach = r["alg_k1"] / (r["k1_ms"] * 1e-3) / 1e9
                    e = {"workload": r["desc"], "value": world * r["B"] / (ms_x * 1e-3), "unit": "scans/s",
                         "ms_per_step": ms_x, "steps": steps, "mpoints_per_s": world * r["B"] / (ms_x * 1e-3) * r["Pn"] / 1e6,
                         "kernel_ms": r["k1_ms"], "roofline": {"kernel": r["k1_note"], "achieved": ach, "frac": ach / peak,
                                                              "algorithmic_bytes_per_launch": r["alg_k1"]},
                         "gpu_launches": r["launches"]}
                    if name == "dense2m":
                        a3 = r["k3_alg"] / (r["k3_ms"] * 1e-3) / 1e9
                        e["k3"] = {"kernel_ms": r["k3_ms"], "achieved": a3, "frac": a3 / peak, "points_out": r["k3_points"]}
                        if world == 1 and not args.no_cpu_baseline:
                            e["matches_gpu"] = cpu_port_baseline(r, 2)["matches_gpu"]
                    else:
                        e["cnn_precision"] = args.cnn_precision
                    extra[name] = e
                del r
                torch.cuda.empty_cache()
            except Exception as ex:      # noqa: BLE001 -- an extra leg must never take the contract line down
                if rank == 0:
                    extra[name] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
        if rank == 0:
            line["extra"] = extra
    if rank == 0:
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
    ap.add_argument("--no-extra", action="store_true", help="skip the dense2m / cnn legs of the default invocation")
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

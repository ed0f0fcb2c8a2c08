"""Copy-only microbenchmark of the host path (VERDICT r1 task 6): what can N processes move over PCIe at the same
time, with no kernel at all?  One process per GPU (python, or torchrun for N > 1), pinned host buffers, H2D and D2H
on separate streams in the byte ratio of the e2e leg (7.86 GB up : 1.89 GB down per 4096 scans).

    python profiles/microbench/host_path.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/microbench/host_path.py

Prints one JSON line: per-rank and aggregate GB/s for H2D alone, D2H alone and both together, and the e2e scans/s
ceiling they imply (2 380 800 B per scan: 1 920 000 up + 460 800 down)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
CH = int(os.environ.get("CHUNK_MB", 64)) << 20
h_up = torch.empty(16 * CH, dtype=torch.uint8).pin_memory()
h_dn = torch.empty(4 * CH, dtype=torch.uint8).pin_memory()
d_up = torch.empty(2 * CH, dtype=torch.uint8, device=dev)
d_dn = torch.empty(2 * CH, dtype=torch.uint8, device=dev)
s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(up, dn, seconds=2.0):
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    nu = nd = 0
    i = 0
    while time.perf_counter() - t0 < seconds:
        if up:
            with torch.cuda.stream(s_up):
                for k in range(4):
                    d_up[(k % 2) * CH:(k % 2 + 1) * CH].copy_(h_up[((i + k) % 16) * CH:((i + k) % 16 + 1) * CH], non_blocking=True)
            nu += 4
        if dn:
            with torch.cuda.stream(s_dn):
                h_dn[(i % 4) * CH:(i % 4 + 1) * CH].copy_(d_dn[(i % 2) * CH:(i % 2 + 1) * CH], non_blocking=True)
            nd += 1
        i += 4
        s_up.synchronize(); s_dn.synchronize()
    dt = time.perf_counter() - t0
    return nu * CH / dt / 1e9, nd * CH / dt / 1e9


res = {}
for name, (u, d) in {"h2d_only": (True, False), "d2h_only": (False, True), "both_4to1": (True, True)}.items():
    run(u, d, 0.5)
    gu, gd = run(u, d)
    t = torch.tensor([gu, gd], dtype=torch.float64, device=dev)
    mn = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    res[name] = {"h2d_GBps_sum": float(t[0]), "d2h_GBps_sum": float(t[1]), "h2d_GBps_min_rank": float(mn[0]), "d2h_GBps_min_rank": float(mn[1])}
if rank == 0:
    both = res["both_4to1"]
    res["n_gpus"] = world
    res["chunk_MB"] = CH >> 20
    # e2e ceiling: the up-stream has to carry 1.92 MB per scan, the down-stream 0.4608 MB
    res["e2e_scans_per_s_ceiling"] = min(both["h2d_GBps_sum"] * 1e9 / 1.92e6, both["d2h_GBps_sum"] * 1e9 / 0.4608e6 if both["d2h_GBps_sum"] else 1e30)
    res["e2e_scans_per_s_ceiling_h2d_only"] = res["h2d_only"]["h2d_GBps_sum"] * 1e9 / 1.92e6
    print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

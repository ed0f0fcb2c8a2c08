#!/bin/bash
# round 2, GPU session i (8 GPUs): K3 check, the contract bench at N = 1/2/4/8, copy-only host-path ceiling at N = 1/2/4/8
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "recon" > gpurun_out/r2i_pytest_k3.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_pytest_k3.log
python profiles/microbench/k3_only.py > gpurun_out/r2i_k3.log 2>&1
python bench.py > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n > gpurun_out/r2i_bench_n$n.json 2> gpurun_out/r2i_bench_n$n.err
done
python profiles/microbench/host_path.py > gpurun_out/r2i_host_path_n1.json 2>/dev/null
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 profiles/microbench/host_path.py > gpurun_out/r2i_host_path_n$n.json 2> gpurun_out/r2i_host_path_n$n.err
done
nvidia-smi topo -m > gpurun_out/r2i_topo.txt 2>&1
tail -2 gpurun_out/r2i_pytest_k3.log; cat gpurun_out/r2i_k3.log
for n in 1 2 4 8; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2i_bench_n$n.json") if x.startswith("{")][-1]; d=json.loads(l)
    print("N=$n value %.0f ms %.3f frac %.3f e2e %.0f dense %s cnn %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"], d.get("extra",{}).get("dense2m",{}).get("value"), d.get("extra",{}).get("cnn",{}).get("value")))
except Exception as e: print("N=$n", e)
PY
done
cat gpurun_out/r2i_host_path_n*.json | cut -c1-700

#!/bin/bash
# round 2, GPU session g: whole GPU suite, K3 under ncu, host-path copy microbench at N=1
python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
python profiles/microbench/k3_only.py > gpurun_out/r2g_k3_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k3_fused -s 5 -c 1 -o gpurun_out/r2g_ncu_k3 python profiles/microbench/k3_only.py > gpurun_out/r2g_k3_ncu.log 2>&1
python profiles/microbench/host_path.py > gpurun_out/r2g_host_path_n1.json 2> gpurun_out/r2g_host_path_n1.err
CHUNK_MB=16 python profiles/microbench/host_path.py > gpurun_out/r2g_host_path_n1_16mb.json 2>/dev/null
tail -5 gpurun_out/r2g_pytest.log
cat gpurun_out/r2g_k3_plain.log gpurun_out/r2g_host_path_n1.json gpurun_out/r2g_host_path_n1_16mb.json

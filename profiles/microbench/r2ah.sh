#!/bin/bash
# round 2, GPU session ah: ncu --set full of K3, committed build
timeout 120 python profiles/microbench/k3_only.py > gpurun_out/r2ah_k3_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k3_fused -s 5 -c 1 -o gpurun_out/r2ah_ncu_k3 python profiles/microbench/k3_only.py > gpurun_out/r2ah_k3_ncu.log 2>&1
cat gpurun_out/r2ah_k3_plain.log

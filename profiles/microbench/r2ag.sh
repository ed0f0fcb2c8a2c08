#!/bin/bash
# round 2, GPU session ag: ncu --set full of K1, committed build
P="timeout 300 python bench.py --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
$P > gpurun_out/r2ag_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_scan -s 4 -c 1 -o gpurun_out/r2ag_ncu_k1 $P > gpurun_out/r2ag_ncu_k1.log 2>&1
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2ag_plain.log

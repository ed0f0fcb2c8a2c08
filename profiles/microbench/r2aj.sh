#!/bin/bash
# round 2, GPU session aj: DRAM traffic of K4 (ncu), committed build
P="timeout 100 python bench.py --workload cnn --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline --no-extra"
$P > gpurun_out/r2aj_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k4_fused -c 2 --csv --log-file gpurun_out/r2aj_k4.csv $P > gpurun_out/r2aj_ncu.log 2>&1
cat gpurun_out/r2aj_k4.csv | tail -8

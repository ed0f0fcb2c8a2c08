#!/bin/bash
# round 2, GPU session y: ncu launch list of the bench step (kernel shares), final build
P="timeout 300 python bench.py --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline"
$P > gpurun_out/r2y_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2y_launches.csv $P > gpurun_out/r2y_launches.log 2>&1
grep -c . gpurun_out/r2y_launches.csv

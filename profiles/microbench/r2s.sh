#!/bin/bash
# round 2, GPU session s: five and six loads per thread per register set
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
S="env DSC_LIB=$PWD/variants/libdsc_shapes.so $B"
for rep in 1 2; do
$B > gpurun_out/r2s_main_$rep.log 2>&1
for c in 3 4 5 6 7; do $S --sc-cfg $c > gpurun_out/r2s_shape${c}_$rep.log 2>&1; done
done
for r in 56 58 60; do $S --sc-cfg 3 --sc-smem-rows $r > gpurun_out/r2s_shape3_rows$r.log 2>&1; $S --sc-cfg 6 --sc-smem-rows $r > gpurun_out/r2s_shape6_rows$r.log 2>&1; done
DSC_LIB=$PWD/variants/libdsc_shapes.so python profiles/microbench/real_scans.py > gpurun_out/r2s_real_default.log 2>&1
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2s_*.log

#!/bin/bash
# round 2, GPU session aa: finish-phase knobs re-measured with the 768 x 6 kernel
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2; do
$B > gpurun_out/r2aa_main_$rep.log 2>&1
for v in fu2 lu3 lu6 q32; do DSC_LIB=$PWD/variants/libdsc_$v.so $B > gpurun_out/r2aa_${v}_$rep.log 2>&1; done
done
for r in 62 64; do DSC_LIB=$PWD/variants/libdsc_q32.so $B --sc-smem-rows $r > gpurun_out/r2aa_q32_rows$r.log 2>&1; done
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2aa_*.log

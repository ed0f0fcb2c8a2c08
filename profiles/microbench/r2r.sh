#!/bin/bash
# round 2, GPU session r: four and five loads per thread per register set
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
S="env DSC_LIB=$PWD/variants/libdsc_shapes.so $B"
for rep in 1 2; do
$B > gpurun_out/r2r_main_$rep.log 2>&1
for c in 3 4 5 6 7; do $S --sc-cfg $c > gpurun_out/r2r_shape${c}_$rep.log 2>&1; done
done
for c in 3 4 6; do
$S --sc-cfg $c --workload uniform80 > gpurun_out/r2r_shape${c}_uniform80.log 2>&1
$S --sc-cfg $c --workload onecell > gpurun_out/r2r_shape${c}_onecell.log 2>&1
$S --sc-cfg $c --workload dense2m > gpurun_out/r2r_shape${c}_dense2m.log 2>&1
done
DSC_LIB=$PWD/variants/libdsc_shapes.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2r_pytest_shapes.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest_shapes.log
tail -2 gpurun_out/r2r_pytest_shapes.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2r_*.log

#!/bin/bash
# round 2, GPU session z: self-cleaning probe (no per-call memsets, parallel window search): parity + A/B
python -m pytest tests -m gpu -q -x > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2 3; do
DSC_LIB=$PWD/variants/libdsc_r2x.so $B > gpurun_out/r2z_r2x_$rep.log 2>&1
$B > gpurun_out/r2z_main_$rep.log 2>&1
done
$B --workload dense2m > gpurun_out/r2z_dense2m.log 2>&1
python profiles/microbench/real_scans.py > gpurun_out/r2z_real.log 2>&1
tail -3 gpurun_out/r2z_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2z_*.log
grep -H -o "\"kernel_ms\": [0-9.]*" gpurun_out/r2z_main_1.log gpurun_out/r2z_r2x_1.log
tail -1 gpurun_out/r2z_real.log

#!/bin/bash
# round 2, GPU session ae: sources with the measured dead ends removed: parity + timing against the previous build
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ae_pytest.log
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2 3; do
DSC_LIB=$PWD/variants/libdsc_nopipe.so $B > gpurun_out/r2ae_prev_$rep.log 2>&1
$B > gpurun_out/r2ae_main_$rep.log 2>&1
done
$B --workload dense2m > gpurun_out/r2ae_dense2m.log 2>&1
timeout 120 python profiles/microbench/real_scans.py > gpurun_out/r2ae_real.log 2>&1
tail -3 gpurun_out/r2ae_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2ae_*.log
tail -1 gpurun_out/r2ae_real.log

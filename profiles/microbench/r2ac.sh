#!/bin/bash
# round 2, GPU session ac: band passes with pipelined list loads: parity + A/B
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ac_pytest.log
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2 3; do
DSC_LIB=$PWD/variants/libdsc_nopipe.so $B > gpurun_out/r2ac_nopipe_$rep.log 2>&1
$B > gpurun_out/r2ac_main_$rep.log 2>&1
done
timeout 120 python profiles/microbench/real_scans.py > gpurun_out/r2ac_real.log 2>&1
DSC_LIB=$PWD/variants/libdsc_nopipe.so timeout 120 python profiles/microbench/real_scans.py > gpurun_out/r2ac_real_nopipe.log 2>&1
tail -3 gpurun_out/r2ac_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2ac_*.log
tail -1 gpurun_out/r2ac_real.log; tail -1 gpurun_out/r2ac_real_nopipe.log

#!/bin/bash
# round 2, GPU session ab: pooled end-of-scan drain of the exact-path queues (groups of 4 warps): parity + A/B
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2ab_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ab_pytest.log
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2 3; do
DSC_LIB=$PWD/variants/libdsc_r2z.so $B > gpurun_out/r2ab_r2z_$rep.log 2>&1
$B > gpurun_out/r2ab_main_$rep.log 2>&1
done
for w in uniform80 dense2m; do $B --workload $w > gpurun_out/r2ab_$w.log 2>&1; done
timeout 120 python profiles/microbench/real_scans.py > gpurun_out/r2ab_real.log 2>&1
tail -3 gpurun_out/r2ab_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2ab_*.log
tail -1 gpurun_out/r2ab_real.log

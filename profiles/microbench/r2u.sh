#!/bin/bash
# round 2, GPU session u: one vote per point for "spills or failed the screen": parity + A/B
python -m pytest tests -m gpu -q -x > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2 3; do
DSC_LIB=$PWD/variants/libdsc_r2t.so $B > gpurun_out/r2u_r2t_$rep.log 2>&1
$B > gpurun_out/r2u_main_$rep.log 2>&1
done
DSC_LIB=$PWD/variants/libdsc_head.so $B > gpurun_out/r2u_head.log 2>&1
for w in uniform80 onecell dense2m; do $B --workload $w > gpurun_out/r2u_$w.log 2>&1; done
python profiles/microbench/real_scans.py > gpurun_out/r2u_real.log 2>&1
tail -3 gpurun_out/r2u_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2u_*.log
tail -1 gpurun_out/r2u_real.log

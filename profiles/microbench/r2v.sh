#!/bin/bash
# round 2, GPU session v: screen without the ring clamp and the overflow test: parity + A/B
python -m pytest tests -m gpu -q -x > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2 3; do
DSC_LIB=$PWD/variants/libdsc_r2u.so $B > gpurun_out/r2v_r2u_$rep.log 2>&1
$B > gpurun_out/r2v_main_$rep.log 2>&1
done
for w in uniform80 onecell dense2m; do $B --workload $w > gpurun_out/r2v_$w.log 2>&1; done
python profiles/microbench/real_scans.py > gpurun_out/r2v_real.log 2>&1
tail -3 gpurun_out/r2v_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2v_*.log
tail -1 gpurun_out/r2v_real.log

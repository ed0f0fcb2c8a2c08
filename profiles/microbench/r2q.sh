#!/bin/bash
# round 2, GPU session q: plain write-out loop (SC_FU=1) as default; launch shapes re-measured now that nothing spills at 64 registers
python -m pytest tests -m gpu -q -x > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2; do
DSC_LIB=$PWD/variants/libdsc_head.so $B > gpurun_out/r2q_head_$rep.log 2>&1
$B > gpurun_out/r2q_main_$rep.log 2>&1
$B --sc-cfg 1 > gpurun_out/r2q_768_$rep.log 2>&1
$B --sc-cfg 2 > gpurun_out/r2q_1024x2_$rep.log 2>&1
done
for v in lu2 lu3; do DSC_LIB=$PWD/variants/libdsc_$v.so $B > gpurun_out/r2q_$v.log 2>&1; done
S="env DSC_LIB=$PWD/variants/libdsc_shapes.so $B"
for c in 3 4 5 6 7; do $S --sc-cfg $c > gpurun_out/r2q_shape$c.log 2>&1; done
for r in 56 58; do $S --sc-cfg 6 --sc-smem-rows $r > gpurun_out/r2q_shape6_rows$r.log 2>&1; done
for r in 58 60 62; do $B --sc-smem-rows $r > gpurun_out/r2q_main_rows$r.log 2>&1; done
$B --workload dense2m > gpurun_out/r2q_dense2m.log 2>&1
$B --workload uniform80 > gpurun_out/r2q_uniform80.log 2>&1
python profiles/microbench/real_scans.py > gpurun_out/r2q_real.log 2>&1
tail -3 gpurun_out/r2q_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2q_*.log
tail -1 gpurun_out/r2q_real.log

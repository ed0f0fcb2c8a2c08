#!/bin/bash
# round 2, GPU session p: write-out with 4 groups in flight, branch-free band entries, one loop counter (no spills left at 64 regs):
# parity, A/B, 1024 x 3 re-measured with row sweeps
python -m pytest tests -m gpu -q -x > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2; do
DSC_LIB=$PWD/variants/libdsc_head.so $B > gpurun_out/r2p_head_$rep.log 2>&1
$B > gpurun_out/r2p_main_$rep.log 2>&1
for v in fu1 fu2 brn lu6; do
DSC_LIB=$PWD/variants/libdsc_$v.so $B > gpurun_out/r2p_${v}_$rep.log 2>&1
done
done
$B --sc-cfg 1 > gpurun_out/r2p_768.log 2>&1
S="env DSC_LIB=$PWD/variants/libdsc_shapes.so $B"
$S --sc-cfg 6 > gpurun_out/r2p_1024x3.log 2>&1
for r in 44 52 56 58; do $S --sc-cfg 6 --sc-smem-rows $r > gpurun_out/r2p_1024x3_rows$r.log 2>&1; done
$S --sc-cfg 2 > gpurun_out/r2p_1024x2.log 2>&1
$S --sc-cfg 2 --sc-smem-rows 58 > gpurun_out/r2p_1024x2_rows58.log 2>&1
$S --sc-cfg 7 > gpurun_out/r2p_832.log 2>&1
$B --workload dense2m > gpurun_out/r2p_dense2m.log 2>&1
python profiles/microbench/real_scans.py > gpurun_out/r2p_real.log 2>&1
tail -3 gpurun_out/r2p_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2p_*.log
tail -1 gpurun_out/r2p_real.log

#!/bin/bash
# round 2, GPU session k: per-set screening (3 interleaved chains) + magic-number rounding (no FRND/F2I): parity, same-box A/B, ncu
python -m pytest tests -m gpu -q -x > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
B="python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2; do
DSC_LIB=$PWD/variants/libdsc_head.so $B > gpurun_out/r2k_head_$rep.log 2>&1
$B > gpurun_out/r2k_main_$rep.log 2>&1
DSC_LIB=$PWD/variants/libdsc_batch_xu.so $B > gpurun_out/r2k_batchxu_$rep.log 2>&1
done
for c in 0 3 4 5; do
DSC_LIB=$PWD/variants/libdsc_shapes.so $B --sc-cfg $c > gpurun_out/r2k_shape$c.log 2>&1
done
$B --workload uniform80 > gpurun_out/r2k_uniform80.log 2>&1
python profiles/microbench/real_scans.py > gpurun_out/r2k_real.log 2>&1
P="python bench.py --steps 2 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
$P > gpurun_out/r2k_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_scan -s 5 -c 1 -o gpurun_out/r2k_ncu_k1 $P > gpurun_out/r2k_ncu_k1.log 2>&1
tail -3 gpurun_out/r2k_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2k_*.log
tail -1 gpurun_out/r2k_real.log

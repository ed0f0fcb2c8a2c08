#!/bin/bash
# round 2, GPU session d: instruction-trimmed k1_scan (predicated smem max, STATS template, register-resident base/ltmask)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_f64.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
B="python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline"
$B > gpurun_out/r2d_default.log 2>&1
$B --sc-smem-rows 56 > gpurun_out/r2d_rows56.log 2>&1
$B --sc-cfg 1 > gpurun_out/r2d_cfg1.log 2>&1
DSC_LIB=$PWD/variants/libdsc_slot.so $B > gpurun_out/r2d_slot.log 2>&1
$B --workload dense2m > gpurun_out/r2d_dense.log 2>&1
$B --workload uniform80 > gpurun_out/r2d_uniform80.log 2>&1
$B --workload onecell > gpurun_out/r2d_onecell.log 2>&1
python profiles/microbench/real_scans.py > gpurun_out/r2d_real.log 2>&1
P="python bench.py --steps 2 --warmup 3 --skip-e2e --no-cpu-baseline"
$P > gpurun_out/r2d_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_scan -s 5 -c 1 -o gpurun_out/r2d_ncu $P > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2d_*.log
tail -1 gpurun_out/r2d_real.log

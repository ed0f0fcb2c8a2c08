#!/bin/bash
# round 2, GPU session b: spill-list k1_scan with the probed resident window; where does the time go
python -m pytest tests/test_gpu_parity.py tests/test_gpu_f64.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
B="python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline"
$B > gpurun_out/r2b_default.log 2>&1
$B --sc-window 0 > gpurun_out/r2b_window0.log 2>&1
for w in 16 24 30; do $B --sc-window $w > gpurun_out/r2b_window$w.log 2>&1; done
for rows in 54 48 66 70; do $B --sc-smem-rows $rows > gpurun_out/r2b_rows$rows.log 2>&1; done
for v in lu4 lu6 lu8 lu4_nopass2 lu4_nospill; do DSC_LIB=$PWD/variants/libdsc_$v.so $B > gpurun_out/r2b_v_$v.log 2>&1; done
$B --workload dense2m > gpurun_out/r2b_dense.log 2>&1
python profiles/microbench/real_scans.py > gpurun_out/r2b_real.log 2>&1
tail -3 gpurun_out/r2b_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2b_*.log
tail -2 gpurun_out/r2b_real.log

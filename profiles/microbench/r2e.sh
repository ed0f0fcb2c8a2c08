#!/bin/bash
# round 2, GPU session e: same-box A/B of the round-1 library, the spill-list kernel (WAW fix, 32-bit list offset) and the slot form
python -m pytest tests/test_gpu_parity.py tests/test_gpu_f64.py tests/test_labels.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
B="python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline"
for rep in 1 2; do
DSC_LIB=$PWD/variants/libdsc_r1.so $B > gpurun_out/r2e_r1_$rep.log 2>&1
$B > gpurun_out/r2e_main_$rep.log 2>&1
DSC_LIB=$PWD/variants/libdsc_slot.so $B > gpurun_out/r2e_slot_$rep.log 2>&1
DSC_LIB=$PWD/variants/libdsc_opt2.so $B > gpurun_out/r2e_opt2_$rep.log 2>&1
done
$B --sc-smem-rows 56 > gpurun_out/r2e_main_rows56.log 2>&1
$B --sc-smem-rows 66 > gpurun_out/r2e_main_rows66.log 2>&1
$B --sc-window 0 > gpurun_out/r2e_main_w0.log 2>&1
$B --workload dense2m > gpurun_out/r2e_dense.log 2>&1
DSC_LIB=$PWD/variants/libdsc_r1.so $B --workload dense2m > gpurun_out/r2e_dense_r1.log 2>&1
P="python bench.py --steps 2 --warmup 3 --skip-e2e --no-cpu-baseline"
$P > gpurun_out/r2e_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_scan -s 5 -c 1 -o gpurun_out/r2e_ncu $P > gpurun_out/r2e_ncu.log 2>&1
tail -3 gpurun_out/r2e_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2e_*.log

#!/bin/bash
# round 2, GPU session c: round-1-style L2 key slots + probed window (A/B build) vs spill lists; ncu of both
B="python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline"
S=$PWD/variants/libdsc_slot.so
DSC_LIB=$S python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or oracle_sweep or full_size or adversarial" > gpurun_out/r2c_pytest_slot.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest_slot.log
DSC_LIB=$S $B > gpurun_out/r2c_slot.log 2>&1
DSC_LIB=$S $B --sc-window 0 > gpurun_out/r2c_slot_w0.log 2>&1
DSC_LIB=$S $B --sc-smem-rows 56 > gpurun_out/r2c_slot_rows56.log 2>&1
$B > gpurun_out/r2c_spill.log 2>&1
P="python bench.py --steps 2 --warmup 3 --skip-e2e --no-cpu-baseline"
$P > gpurun_out/r2c_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_scan -s 5 -c 1 -o gpurun_out/r2c_ncu_spill $P > gpurun_out/r2c_ncu_spill.log 2>&1
DSC_LIB=$S $P > gpurun_out/r2c_plain_slot.log 2>&1 && DSC_LIB=$S ncu --set full --clock-control none --import-source on -k regex:k1_scan -s 5 -c 1 -o gpurun_out/r2c_ncu_slot $P > gpurun_out/r2c_ncu_slot.log 2>&1
tail -2 gpurun_out/r2c_pytest_slot.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2c_*.log
ls -la gpurun_out/*.ncu-rep

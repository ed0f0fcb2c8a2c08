#!/bin/bash
# round 2, GPU session m: cp.async staged stream (3/4/5 groups), 32-entry exact queues, 896 x 3 shape
for v in async4 q32; do
DSC_LIB=$PWD/variants/libdsc_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2m_pytest_$v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest_$v.log
done
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2; do
$B > gpurun_out/r2m_main_$rep.log 2>&1
for v in async3 async4 async5 q32; do
DSC_LIB=$PWD/variants/libdsc_$v.so $B > gpurun_out/r2m_${v}_$rep.log 2>&1
done
done
DSC_LIB=$PWD/variants/libdsc_shapes.so $B --sc-cfg 6 > gpurun_out/r2m_shape6.log 2>&1
DSC_LIB=$PWD/variants/libdsc_q32.so $B --sc-smem-rows 64 > gpurun_out/r2m_q32_rows64.log 2>&1
DSC_LIB=$PWD/variants/libdsc_async4.so $B --sc-smem-rows 56 > gpurun_out/r2m_async4_rows56.log 2>&1
for f in gpurun_out/r2m_pytest*.log; do echo $f; tail -2 $f; done
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2m_*.log

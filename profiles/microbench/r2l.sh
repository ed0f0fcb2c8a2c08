#!/bin/bash
# round 2, GPU session l: floor-magic screen (no FRND/F2I), raw-bit keys in offset mode, 32-bit spill index; A/B of the
# per-set vote and of dropping the LDS pre-check
python -m pytest tests -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
for v in vs np vsnp; do
DSC_LIB=$PWD/variants/libdsc_$v.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2l_pytest_$v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest_$v.log
done
B="python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2; do
DSC_LIB=$PWD/variants/libdsc_head.so $B > gpurun_out/r2l_head_$rep.log 2>&1
$B > gpurun_out/r2l_main_$rep.log 2>&1
for v in vs np vsnp; do
DSC_LIB=$PWD/variants/libdsc_$v.so $B > gpurun_out/r2l_${v}_$rep.log 2>&1
done
done
$B --workload uniform80 > gpurun_out/r2l_uniform80.log 2>&1
python profiles/microbench/real_scans.py > gpurun_out/r2l_real.log 2>&1
P="python bench.py --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
$P > gpurun_out/r2l_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_scan -s 4 -c 1 -o gpurun_out/r2l_ncu_k1 $P > gpurun_out/r2l_ncu_k1.log 2>&1
for f in gpurun_out/r2l_pytest*.log; do echo $f; tail -2 $f; done
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2l_*.log
tail -1 gpurun_out/r2l_real.log

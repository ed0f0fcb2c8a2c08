#!/bin/bash
# round 2, GPU session w: K3 with the polar tables in shared memory: parity + timing against the previous library
python -m pytest tests -m gpu -q -x > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log
for rep in 1 2 3; do
DSC_LIB=$PWD/variants/libdsc_r2u.so timeout 120 python profiles/microbench/k3_only.py > gpurun_out/r2w_k3_r2u_$rep.log 2>&1
timeout 120 python profiles/microbench/k3_only.py > gpurun_out/r2w_k3_main_$rep.log 2>&1
done
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
$B --workload dense2m > gpurun_out/r2w_dense2m.log 2>&1
tail -3 gpurun_out/r2w_pytest.log
grep -H "K3 ms" gpurun_out/r2w_k3*.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2w_*.log

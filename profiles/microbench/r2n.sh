#!/bin/bash
# round 2, GPU session n: L2 prefetch of the next work item during the write-out (K1) / of the next grid during pass C (K3);
# launch shapes 1024x3, 1024x2, 896x3, 832x3 re-measured
DSC_LIB=$PWD/variants/libdsc_pf.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shims.py -m gpu -q -x > gpurun_out/r2n_pytest_pf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_pf.log
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2; do
for kb in 0 256 512 1024 1920; do
$B --sc-prefetch-kb $kb > gpurun_out/r2n_pf${kb}_$rep.log 2>&1
done
done
for c in 1 2 6 7; do
DSC_LIB=$PWD/variants/libdsc_shapes.so $B --sc-cfg $c > gpurun_out/r2n_shape$c.log 2>&1
DSC_LIB=$PWD/variants/libdsc_shapes.so $B --sc-cfg $c --sc-prefetch-kb 512 > gpurun_out/r2n_shape${c}_pf512.log 2>&1
done
$B --workload dense2m > gpurun_out/r2n_dense2m_pf0.log 2>&1
$B --workload dense2m --sc-prefetch-kb 512 > gpurun_out/r2n_dense2m_pf512.log 2>&1
for rep in 1 2; do
K3_PREFETCH=0 timeout 120 python profiles/microbench/k3_only.py > gpurun_out/r2n_k3_pf0_$rep.log 2>&1
K3_PREFETCH=1 timeout 120 python profiles/microbench/k3_only.py > gpurun_out/r2n_k3_pf1_$rep.log 2>&1
done
for f in gpurun_out/r2n_pytest*.log; do echo $f; tail -2 $f; done
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2n_*.log
grep -H "K3 ms" gpurun_out/r2n_k3*.log

#!/bin/bash
# round 2, GPU session af: K3 pass A with two rows ping-pong per warp (768 / 1024 threads) against the committed build
for v in k3pp768 k3pp1024; do
DSC_LIB=$PWD/variants/libdsc_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shims.py -m gpu -q -x -k "reconstruct or K3 or k3 or cloud or inverse" > gpurun_out/r2af_pytest_$v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2af_pytest_$v.log
done
for rep in 1 2; do
timeout 120 python profiles/microbench/k3_only.py > gpurun_out/r2af_k3_main_$rep.log 2>&1
for v in k3pp768 k3pp1024; do DSC_LIB=$PWD/variants/libdsc_$v.so timeout 120 python profiles/microbench/k3_only.py > gpurun_out/r2af_k3_${v}_$rep.log 2>&1; done
done
for f in gpurun_out/r2af_pytest*.log; do tail -2 $f; done
grep -H "K3 ms" gpurun_out/r2af_k3*.log

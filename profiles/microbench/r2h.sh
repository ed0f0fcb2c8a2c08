#!/bin/bash
# round 2, GPU session h: whole GPU suite after the K3 overlap / K7 fixes; K3 + other kernels timing
python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
python profiles/microbench/k3_only.py > gpurun_out/r2h_k3.log 2>&1
python profiles/microbench/other_kernels.py > gpurun_out/r2h_other_kernels.json 2> gpurun_out/r2h_other_kernels.err
python bench.py --workload dense2m --steps 10 --skip-e2e --no-cpu-baseline > gpurun_out/r2h_dense.log 2>&1
tail -4 gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_k3.log; head -3 gpurun_out/r2h_other_kernels.json
grep -o "\"ms_per_step\": [0-9.]*\|\"roofline_k3\": {[^}]*}" gpurun_out/r2h_dense.log

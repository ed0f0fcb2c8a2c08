#!/bin/bash
# round 2, GPU session ai: K4 with the head of the grid kept in shared memory: parity + K1+K4 time of the cnn leg against the previous build
timeout 200 python -m pytest tests -m gpu -q -x -k "norm or tensor01 or cnn or golden_grids" > gpurun_out/r2ai_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ai_pytest.log
B="timeout 100 python bench.py --workload cnn --steps 3 --warmup 2 --skip-e2e --no-cpu-baseline --no-extra"
DSC_LIB=$PWD/variants/libdsc_final0.so $B > gpurun_out/r2ai_cnn_prev.log 2>&1
$B > gpurun_out/r2ai_cnn_main.log 2>&1
tail -2 gpurun_out/r2ai_pytest.log
grep -H -o "\"kernel_ms\": [0-9.]*" gpurun_out/r2ai_cnn_*.log

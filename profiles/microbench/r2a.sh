#!/bin/bash
# round 2, GPU session a: parity of the spill-list k1_scan + first timings (resident-row sweep)
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
for rows in -1 54 48 40; do python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --sc-smem-rows $rows > gpurun_out/r2a_bench_rows$rows.log 2>&1; done
python bench.py --steps 10 --skip-e2e --no-cpu-baseline --sc-cfg 1 > gpurun_out/r2a_bench_cfg1.log 2>&1
python bench.py --steps 10 --skip-e2e --no-cpu-baseline --workload dense2m > gpurun_out/r2a_bench_dense.log 2>&1
tail -3 gpurun_out/r2a_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*\|\"frac\": [0-9.]*" gpurun_out/r2a_bench_*.log

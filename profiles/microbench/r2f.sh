#!/bin/bash
# round 2, GPU session f: whole GPU suite (K3 compaction, K6, K7, f64), the new bench line with its extra legs, K3/K4 microbench
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
( time python bench.py ) > gpurun_out/r2f_bench.log 2>&1
( time python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r2f_bench_ref.log 2>&1
python profiles/microbench/other_kernels.py > gpurun_out/r2f_other_kernels.json 2> gpurun_out/r2f_other_kernels.err
B="python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
$B --sc-cfg 1 > gpurun_out/r2f_cfg1.log 2>&1
$B --sc-cfg 2 > gpurun_out/r2f_cfg2.log 2>&1
python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1
tail -4 gpurun_out/r2f_pytest.log
tail -c 3000 gpurun_out/r2f_bench.log
tail -3 gpurun_out/r2f_bench_ref.log | cut -c1-400
cat gpurun_out/r2f_other_kernels.json | head -30
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2f_cfg*.log
tail -2 gpurun_out/r2f_smoke.log

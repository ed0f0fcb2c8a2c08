#!/bin/bash
# round 2, GPU session o: 896 x 3 as the default launch shape: full parity, same-box A/B against 768 x 3 and the round-start library, contract bench, ncu
python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2; do
DSC_LIB=$PWD/variants/libdsc_head.so $B > gpurun_out/r2o_head_$rep.log 2>&1
$B > gpurun_out/r2o_main_$rep.log 2>&1
$B --sc-cfg 1 > gpurun_out/r2o_768_$rep.log 2>&1
done
for w in uniform80 onecell dense2m; do
$B --workload $w > gpurun_out/r2o_$w.log 2>&1
$B --workload $w --sc-cfg 1 > gpurun_out/r2o_${w}_768.log 2>&1
done
$B --sc-smem-rows 58 > gpurun_out/r2o_rows58.log 2>&1
$B --sc-smem-rows 60 > gpurun_out/r2o_rows60.log 2>&1
python profiles/microbench/real_scans.py > gpurun_out/r2o_real.log 2>&1
timeout 600 python bench.py > gpurun_out/r2o_bench_default.json 2> gpurun_out/r2o_bench_default.err
P="timeout 300 python bench.py --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
$P > gpurun_out/r2o_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_scan -s 4 -c 1 -o gpurun_out/r2o_ncu_k1 $P > gpurun_out/r2o_ncu_k1.log 2>&1
tail -3 gpurun_out/r2o_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2o_*.log
tail -1 gpurun_out/r2o_real.log
cat gpurun_out/r2o_bench_default.json

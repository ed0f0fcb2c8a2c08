#!/bin/bash
# round 2, GPU session t: 768 x 6 as the default shape, masked two-load sets for the ragged tail: parity, A/B, contract bench, ncu, launch list
python -m pytest tests -m gpu -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2; do
DSC_LIB=$PWD/variants/libdsc_head.so $B > gpurun_out/r2t_head_$rep.log 2>&1
$B > gpurun_out/r2t_main_$rep.log 2>&1
$B --sc-cfg 1 > gpurun_out/r2t_768x3_$rep.log 2>&1
DSC_LIB=$PWD/variants/libdsc_ut1.so $B > gpurun_out/r2t_ut1_$rep.log 2>&1
done
for w in uniform80 onecell dense2m; do $B --workload $w > gpurun_out/r2t_$w.log 2>&1; done
$B --points 123000 > gpurun_out/r2t_main_123k.log 2>&1
DSC_LIB=$PWD/variants/libdsc_ut1.so $B --points 123000 > gpurun_out/r2t_ut1_123k.log 2>&1
$B --sc-cfg 1 --points 123000 > gpurun_out/r2t_768x3_123k.log 2>&1
python profiles/microbench/real_scans.py > gpurun_out/r2t_real.log 2>&1
DSC_LIB=$PWD/variants/libdsc_ut1.so python profiles/microbench/real_scans.py > gpurun_out/r2t_real_ut1.log 2>&1
timeout 600 python bench.py > gpurun_out/r2t_bench_default.json 2> gpurun_out/r2t_bench_default.err
P="timeout 300 python bench.py --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
$P > gpurun_out/r2t_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_scan -s 4 -c 1 -o gpurun_out/r2t_ncu_k1 $P > gpurun_out/r2t_ncu_k1.log 2>&1
tail -3 gpurun_out/r2t_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2t_*.log
tail -1 gpurun_out/r2t_real.log; tail -1 gpurun_out/r2t_real_ut1.log
cat gpurun_out/r2t_bench_default.json | cut -c1-1500

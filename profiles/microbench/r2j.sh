#!/bin/bash
# round 2, GPU session j: lane-private spill lists + one screen vote per set: parity, same-box A/B, ncu of K1 and K3, launch list
python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
B="python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2; do
DSC_LIB=$PWD/variants/libdsc_r1.so $B > gpurun_out/r2j_r1_$rep.log 2>&1
$B > gpurun_out/r2j_main_$rep.log 2>&1
done
$B --sc-smem-rows 58 > gpurun_out/r2j_rows58.log 2>&1
$B --workload uniform80 > gpurun_out/r2j_uniform80.log 2>&1
$B --workload onecell > gpurun_out/r2j_onecell.log 2>&1
python profiles/microbench/real_scans.py > gpurun_out/r2j_real.log 2>&1
P="python bench.py --steps 2 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
$P > gpurun_out/r2j_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_scan -s 5 -c 1 -o gpurun_out/r2j_ncu_k1 $P > gpurun_out/r2j_ncu_k1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2j_launches.csv $P > gpurun_out/r2j_launches.log 2>&1
python profiles/microbench/k3_only.py > gpurun_out/r2j_k3_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k3_fused -s 5 -c 1 -o gpurun_out/r2j_ncu_k3 python profiles/microbench/k3_only.py > gpurun_out/r2j_k3_ncu.log 2>&1
tail -3 gpurun_out/r2j_pytest.log
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2j_*.log
tail -1 gpurun_out/r2j_real.log; cat gpurun_out/r2j_k3_plain.log

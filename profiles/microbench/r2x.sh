#!/bin/bash
# round 2, GPU session x: final state -- full parity, contract bench (default invocation), A/B against the round-start library, ncu of K1
python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2x_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/r2x_bench_default.json 2> gpurun_out/r2x_bench_default.err
timeout 600 python bench.py --impl reference > gpurun_out/r2x_bench_reference.json 2> gpurun_out/r2x_bench_reference.err
B="timeout 120 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
for rep in 1 2; do
DSC_LIB=$PWD/variants/libdsc_head.so $B > gpurun_out/r2x_head_$rep.log 2>&1
$B > gpurun_out/r2x_main_$rep.log 2>&1
done
for w in uniform80 onecell dense2m cnn; do $B --workload $w > gpurun_out/r2x_$w.log 2>&1; done
python profiles/microbench/real_scans.py > gpurun_out/r2x_real.log 2>&1
python profiles/microbench/k3_only.py > gpurun_out/r2x_k3.log 2>&1
P="timeout 300 python bench.py --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra"
$P > gpurun_out/r2x_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_scan -s 4 -c 1 -o gpurun_out/r2x_ncu_k1 $P > gpurun_out/r2x_ncu_k1.log 2>&1
tail -3 gpurun_out/r2x_pytest.log; cat gpurun_out/r2x_smoke.log | tail -2
grep -H -o "\"ms_per_step\": [0-9.]*" gpurun_out/r2x_*.log
tail -1 gpurun_out/r2x_real.log; cat gpurun_out/r2x_k3.log
cut -c1-700 gpurun_out/r2x_bench_default.json; echo; cut -c1-600 gpurun_out/r2x_bench_reference.json

#!/bin/bash
# round 2, GPU session ad: the committed build, end to end -- build(), full parity, smoke, contract bench (both arms)
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/r2ad_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ad_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2ad_smoke.log 2>&1
timeout 600 python bench.py --impl reference > gpurun_out/r2ad_bench_reference.json 2> gpurun_out/r2ad_bench_reference.err
timeout 600 python bench.py > gpurun_out/r2ad_bench_default.json 2> gpurun_out/r2ad_bench_default.err
tail -2 gpurun_out/r2ad_build.log; tail -3 gpurun_out/r2ad_pytest.log; tail -2 gpurun_out/r2ad_smoke.log
cut -c1-200 gpurun_out/r2ad_bench_reference.json; echo
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ad_bench_default.json').read().strip().splitlines()[-1])
r=d["roofline"]; print("value",d["value"],"ms",d["ms_per_step"],"frac",r["frac"],"kernel_ms",r["kernel_ms"],"traffic",r["traffic"],"e2e",d["e2e"]["value"],"launches",d["gpu_launches"],d["clocks"])
x=d["extra"]; print("dense2m",x["dense2m"]["value"],x["dense2m"]["roofline"]["frac"],"cnn",x["cnn"]["value"])
PY

#!/bin/bash
# GPU round T: full verification -- parity suite, smoke, judged bench, launch list
set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout 120 --durations=8 > gpurun_out/t_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/t_pytest.txt
grep -E "FAIL|passed|failed|exit|Error" gpurun_out/t_pytest.txt | cut -c1-250 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t_smoke.txt 2>&1; tail -2 gpurun_out/t_smoke.txt
timeout 900 python bench.py > gpurun_out/t_bench.txt 2>&1
tail -1 gpurun_out/t_bench.txt | cut -c1-2600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/t_launches.csv python bench.py --steps 5 --warmup 3 --cpu-seconds 1 > gpurun_out/t_ncu_bench.log 2>&1
grep -c ot_fused gpurun_out/t_launches.csv

#!/bin/bash
# Final verification of round 1: parity suite, smoke, judged bench (default flags), reference arm, launch list, full ncu
set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/fin_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/fin_pytest.txt
grep -E "FAIL|passed|failed|exit|Error" gpurun_out/fin_pytest.txt | cut -c1-250 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.txt 2>&1; tail -1 gpurun_out/fin_smoke.txt
timeout 900 python bench.py > gpurun_out/fin_bench.txt 2>&1
tail -1 gpurun_out/fin_bench.txt | cut -c1-2700
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin_bench_ref.txt 2>&1
tail -1 gpurun_out/fin_bench_ref.txt | cut -c1-400
timeout 300 python tools/quick_bench.py 1000 64000 256000 1024000 > gpurun_out/fin_quick.txt 2>&1; cat gpurun_out/fin_quick.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 5 --warmup 3 --cpu-seconds 1 > gpurun_out/fin_ncu_bench.log 2>&1
grep -c ot_fused gpurun_out/fin_launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ot_fused_v7 -s 4 -c 1 -o gpurun_out/fin_fused python tools/quick_bench.py 256000 > gpurun_out/fin_ncu.log 2>&1
tail -1 gpurun_out/fin_ncu.log

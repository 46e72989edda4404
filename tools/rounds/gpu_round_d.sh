#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -v --timeout 150 > gpurun_out/d_pytest.txt 2>&1; echo "exit $?" >> gpurun_out/d_pytest.txt
grep -E "FAIL|ERROR|exit|Timeout|passed|failed|Error|assert " gpurun_out/d_pytest.txt | cut -c1-300 | tail -30
timeout 600 python tools/quick_bench.py 1000 8000 64000 256000 > gpurun_out/d_quick.txt 2>&1
cat gpurun_out/d_quick.txt
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/d_bench.txt 2>&1
tail -2 gpurun_out/d_bench.txt | cut -c1-1500
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ot_fused -s 4 -c 1 -o gpurun_out/d_fused python tools/quick_bench.py 64000 > gpurun_out/d_ncu.log 2>&1
tail -3 gpurun_out/d_ncu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/d_smoke.txt 2>&1; tail -2 gpurun_out/d_smoke.txt

#!/bin/bash
# GPU round E: full parity suite, micro-bench, judged bench, launch list, one --set full capture of the fused kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/e_gpu.txt
timeout 700 python -m pytest tests -m gpu -v --timeout 120 --durations=15 > gpurun_out/e_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/e_pytest.txt
grep -E "FAIL|ERROR|exit|Timeout|passed|failed|Error|assert " gpurun_out/e_pytest.txt | cut -c1-300 | tail -30
timeout 300 python tools/quick_bench.py 1000 8000 64000 256000 > gpurun_out/e_quick.txt 2>&1
cat gpurun_out/e_quick.txt
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/e_bench.txt 2>&1
tail -2 gpurun_out/e_bench.txt | cut -c1-2500
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ot_fused -s 4 -c 1 -o gpurun_out/e_fused python tools/quick_bench.py 64000 > gpurun_out/e_ncu.log 2>&1
tail -3 gpurun_out/e_ncu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/e_smoke.txt 2>&1; tail -2 gpurun_out/e_smoke.txt

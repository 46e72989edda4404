#!/bin/bash
# GPU round r2x: sort-free single-pass top-k + selecting merge; top-k tests; durations
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_misc_gpu.py tests/test_parity_l2max_gpu.py -m gpu -q --timeout 300 -k "topk or rank_corpus or config3 or shard" > gpurun_out/r2x_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2x_pytest.txt
grep -E "FAIL|passed|failed|exit|Error|assert" gpurun_out/r2x_pytest.txt | cut -c1-250 | tail -10
timeout 200 python tools/side_bench.py topk > gpurun_out/r2x_side.txt 2>&1; cat gpurun_out/r2x_side.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:topk -c 4 --csv --log-file gpurun_out/r2x_topk.csv python tools/side_bench.py topk > /dev/null 2>&1
grep -E "topk" gpurun_out/r2x_topk.csv | cut -d, -f5,15- | cut -c1-140

#!/bin/bash
# round 2, call A: first run of the one-kernel long-document path (ot_varlen.cu): its parity tests, the old tests that
# now route through it, and the config-5 developer bench (new path, then the two-kernel path for comparison)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_varlen_gpu.py tests/test_parity_ot_gpu.py tests/test_parity_l2max_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/r2a_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.txt
grep -E "FAIL|passed|failed|exit|Error|assert" gpurun_out/r2a_pytest.txt | cut -c1-250 | tail -12
timeout 300 python tools/side_bench.py varlen > gpurun_out/r2a_side_varlen.txt 2>&1; cat gpurun_out/r2a_side_varlen.txt

#!/bin/bash
# round 2, call I: swizzled 128-byte ring rows (LDGSTS wavefronts), sorted l2max, full GPU suite, ncu
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2i_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest.txt
grep -E "FAIL|passed|failed|exit|Error|assert" gpurun_out/r2i_pytest.txt | cut -c1-250 | tail -12
timeout 300 python tools/side_bench.py varlen > gpurun_out/r2i_side_varlen.txt 2>&1; cat gpurun_out/r2i_side_varlen.txt
ASP_VARLEN_B=20000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ot_varlen -s 3 -c 1 -o gpurun_out/r2i_varlen python tools/side_bench.py varlen > gpurun_out/r2i_ncu_log.txt 2>&1; tail -3 gpurun_out/r2i_ncu_log.txt

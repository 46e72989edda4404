#!/bin/bash
# GPU round r4n: short final check after the last GEMM change: parity suite, smoke, judged bench (default flags)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r4n_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r4n_pytest.txt
grep -E "FAIL|passed|failed|exit|Error" gpurun_out/r4n_pytest.txt | cut -c1-250 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4n_smoke.txt 2>&1; tail -1 gpurun_out/r4n_smoke.txt
timeout 600 python bench.py > gpurun_out/r4n_bench.txt 2>&1
tail -1 gpurun_out/r4n_bench.txt | cut -c1-600

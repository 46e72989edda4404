#!/bin/bash
# GPU round r2m: single-pass top-k + packed gather, spec-size parity tests, side-kernel section of bench.py, reference mirror arm
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --durations=8 > gpurun_out/r2m_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2m_pytest.txt
grep -E "FAIL|passed|failed|exit|Error" gpurun_out/r2m_pytest.txt | cut -c1-250 | tail -8
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.txt 2>&1; tail -1 gpurun_out/r2m_smoke.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_bench_ref.txt 2>&1; tail -1 gpurun_out/r2m_bench_ref.txt | cut -c1-600
timeout 900 python bench.py > gpurun_out/r2m_bench.txt 2>&1
tail -1 gpurun_out/r2m_bench.txt | cut -c1-3000

#!/bin/bash
# GPU round ZZ: last verification of round 1 (192-wide GEMM tiles, CTA-pair option): parity suite, smoke, encoder and
# GEMM developer benches, judged bench with default flags
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/zz_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/zz_pytest.txt
grep -E "FAIL|passed|failed|exit|Error" gpurun_out/zz_pytest.txt | cut -c1-250 | tail -6
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/zz_smoke.txt 2>&1; tail -1 gpurun_out/zz_smoke.txt
timeout 200 python tools/encoder_bench.py > gpurun_out/zz_encoder_bench.txt 2>&1; tail -4 gpurun_out/zz_encoder_bench.txt
timeout 100 python tools/gemm_bench.py --tokens 8192 2048 --modes 0 3 > gpurun_out/zz_gemm_bench.txt 2>&1
timeout 100 python tools/gemm_bench.py --tokens 8192 --modes 0 3 --x3 >> gpurun_out/zz_gemm_bench.txt 2>&1
timeout 100 python tools/gemm_bench.py --tokens 8192 --modes 3 --pairs 1 2 >> gpurun_out/zz_gemm_bench.txt 2>&1; cat gpurun_out/zz_gemm_bench.txt
timeout 600 python bench.py > gpurun_out/zz_bench.txt 2>&1
tail -1 gpurun_out/zz_bench.txt | cut -c1-900

#!/bin/bash
# GPU round Y: re-verification after the encoder work (persistent GEMM, cp.async attention, dependent launch):
# full parity suite, smoke, judged bench with default flags, encoder + GEMM developer benches, ncu of the encoder kernels
set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/y_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/y_pytest.txt
grep -E "FAIL|passed|failed|exit|Error" gpurun_out/y_pytest.txt | cut -c1-250 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/y_smoke.txt 2>&1; tail -1 gpurun_out/y_smoke.txt
timeout 900 python bench.py > gpurun_out/y_bench.txt 2>&1
tail -1 gpurun_out/y_bench.txt | cut -c1-2700
timeout 300 python tools/encoder_bench.py > gpurun_out/y_encoder_bench.txt 2>&1; tail -4 gpurun_out/y_encoder_bench.txt
timeout 200 python tools/gemm_bench.py --tokens 8192 2048 --modes 0 3 > gpurun_out/y_gemm_bench.txt 2>&1; cat gpurun_out/y_gemm_bench.txt
timeout 200 python tools/gemm_bench.py --tokens 8192 --modes 0 3 --x3 >> gpurun_out/y_gemm_bench.txt 2>&1; tail -4 gpurun_out/y_gemm_bench.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn|attention_kernel|ln_kernel" -s 50 -c 7 -o gpurun_out/y_encoder python tools/encoder_bench.py --quick > gpurun_out/y_ncu.log 2>&1; tail -1 gpurun_out/y_ncu.log

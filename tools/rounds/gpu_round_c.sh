#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -v --timeout 150 > gpurun_out/c_pytest.txt 2>&1; echo "exit $?" >> gpurun_out/c_pytest.txt
grep -E "FAIL|ERROR|exit|Timeout|passed|failed|Error" gpurun_out/c_pytest.txt | tail -30
timeout 600 python tools/quick_bench.py 1000 8000 64000 256000 > gpurun_out/c_quick.txt 2>&1
cat gpurun_out/c_quick.txt
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/c_bench.txt 2>&1
tail -2 gpurun_out/c_bench.txt
timeout 600 python tools/encoder_bench.py > gpurun_out/c_encoder_bench.txt 2>&1
tail -6 gpurun_out/c_encoder_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ot_fused -s 4 -c 1 -o gpurun_out/c_fused python tools/quick_bench.py 64000 > gpurun_out/c_ncu.log 2>&1
tail -3 gpurun_out/c_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 30 -c 3 -o gpurun_out/c_gemm python tools/encoder_bench.py --quick > gpurun_out/c_ncu_gemm.log 2>&1
tail -3 gpurun_out/c_ncu_gemm.log

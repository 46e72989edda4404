#!/bin/bash
# GPU-box visit B: tcgen05 GEMM + encoder tests first (isolated, short timeouts), then the full GPU suite, dev bench, ncu.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -v --timeout 120 -x > gpurun_out/b_gemm.txt 2>&1; echo "exit $?" >> gpurun_out/b_gemm.txt
tail -25 gpurun_out/b_gemm.txt
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -v --timeout 200 > gpurun_out/b_encoder.txt 2>&1; echo "exit $?" >> gpurun_out/b_encoder.txt
tail -25 gpurun_out/b_encoder.txt
timeout 900 python -m pytest tests -m gpu -v --timeout 120 --deselect tests/test_gemm_gpu.py --deselect tests/test_encoder_gpu.py > gpurun_out/b_pytest.txt 2>&1; echo "exit $?" >> gpurun_out/b_pytest.txt
grep -E "PASS|FAIL|ERROR|exit|Timeout|passed|failed" gpurun_out/b_pytest.txt | tail -70
timeout 600 python tools/quick_bench.py 1000 8000 64000 256000 > gpurun_out/b_quick.txt 2>&1
cat gpurun_out/b_quick.txt
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/b_bench.txt 2>&1
tail -3 gpurun_out/b_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ot_fused -s 4 -c 1 -o gpurun_out/b_fused python tools/quick_bench.py 64000 > gpurun_out/b_ncu.log 2>&1
tail -3 gpurun_out/b_ncu.log

#!/bin/bash
# GPU round r2n: first run of the tcgen05 all-pairs otAspire kernel (parity tests, then the side bench)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_ot_gpu.py -m gpu -q --timeout 120 -k "allpairs" > gpurun_out/r2n_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2n_pytest.txt
grep -E "FAIL|passed|failed|exit|Error|assert" gpurun_out/r2n_pytest.txt | cut -c1-250 | tail -12
timeout 300 python tools/side_bench.py otallpairs > gpurun_out/r2n_side.txt 2>&1; cat gpurun_out/r2n_side.txt | tail -5

#!/bin/bash
# GPU round r2t: tensor-core pool kernel (ot_fused_tc.cu): parity, then sustained A/B against the FFMA2 kernel
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2t_ab.txt
timeout 120 python -m pytest tests/test_parity_ot_gpu.py -m gpu -q --timeout 60 -x -k "pool_kernel" > gpurun_out/r2t_pytest.txt 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/r2t_pytest.txt
grep -E "FAIL|passed|failed|exit|Error|assert" gpurun_out/r2t_pytest.txt | cut -c1-300 | tail -12
if [ $rc -ne 0 ]; then exit 0; fi
for m in 1 0; do
echo "== ot_fused_tc=$m" >> gpurun_out/r2t_ab.txt
ASP_TC=$m timeout 120 python tools/sustained_ab.py >> gpurun_out/r2t_ab.txt 2>&1
done
cat gpurun_out/r2t_ab.txt

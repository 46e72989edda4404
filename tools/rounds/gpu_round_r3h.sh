#!/bin/bash
# GPU round r3h: persistent tcgen05 attention: parity, then encoder bench modes 2 / 1 / 0
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_encoder_gpu.py -m gpu -q --timeout 90 -x -k "tcgen05_attention" > gpurun_out/r3h_pytest.txt 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/r3h_pytest.txt
grep -E "FAIL|passed|failed|exit|Error|assert" gpurun_out/r3h_pytest.txt | cut -c1-300 | tail -12
if [ $rc -ne 0 ]; then exit 0; fi
rm -f gpurun_out/r3h_enc.txt
for m in 2 1 0; do
  echo "== attn_tc=$m" >> gpurun_out/r3h_enc.txt
  timeout 200 python tools/encoder_bench.py --quick --opt=attn_tc=$m 2>&1 | tail -1 >> gpurun_out/r3h_enc.txt
done
echo "== attn_tc=2, all shapes" >> gpurun_out/r3h_enc.txt
timeout 200 python tools/encoder_bench.py --opt=attn_tc=2 2>&1 | tail -4 >> gpurun_out/r3h_enc.txt
cat gpurun_out/r3h_enc.txt

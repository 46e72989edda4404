#!/bin/bash
# GPU round r3a: tcgen05 attention: parity inside the encoder, then encoder bench with / without it
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_encoder_gpu.py -m gpu -q --timeout 100 -x -k "tcgen05_attention" > gpurun_out/r3a_pytest.txt 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/r3a_pytest.txt
grep -E "FAIL|passed|failed|exit|Error|assert" gpurun_out/r3a_pytest.txt | cut -c1-300 | tail -12
if [ $rc -ne 0 ]; then exit 0; fi
timeout 200 python tools/encoder_bench.py --opt=attn_tc=1 > gpurun_out/r3a_enc_tc.txt 2>&1; tail -6 gpurun_out/r3a_enc_tc.txt
timeout 200 python tools/encoder_bench.py --quick --opt=attn_tc=0 > gpurun_out/r3a_enc_mma.txt 2>&1; tail -3 gpurun_out/r3a_enc_mma.txt
timeout 300 python -m pytest tests/test_encoder_gpu.py -m gpu -q --timeout 200 2>&1 | tail -2

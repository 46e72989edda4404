#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_ot_gpu.py -m gpu -q --timeout 200 -k "structured" 2>&1 | tail -8

#!/bin/bash
# GPU round r4l: LayerNorm on read through the other GEMM kernels
timeout 300 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu -k layernorm 2>&1 | grep -E "passed|failed|Error|assert" | head

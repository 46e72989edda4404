#!/bin/bash
timeout 100 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu -k pair_auto 2>&1 | grep -E "passed|failed|Error|assert" | head -5

#!/bin/bash
# GPU round r4j: the encoder tests four times over (timing-dependent failures would show)
for i in 1 2 3 4; do timeout 300 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu 2>&1 | grep -E "passed|failed"; done

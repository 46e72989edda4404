#!/bin/bash
# GPU round r4f: many-tiles-per-CTA test of the pipelined attention kernel + the whole encoder test file
set -x
timeout 600 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu -k exact 2>&1 | tail -4

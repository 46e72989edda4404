#!/bin/bash
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 100 python -m pytest tests/test_kernels_misc_gpu.py -x -q -m gpu -k "span" 2>&1 | grep -E "passed|failed|Error" | head -3

#!/bin/bash
# GPU round r3e: edge-case test of the all-pairs kernel; synccheck with a larger barrier table; racecheck on the round-2 kernels
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_parity_ot_gpu.py -m gpu -q --timeout 100 -k "empty_documents" 2>&1 | tail -3
timeout 900 compute-sanitizer --tool synccheck --num-cuda-barriers 65536 python tools/sanitize_small.py > gpurun_out/r3e_san_synccheck.txt 2>&1
echo "== synccheck: $(grep -E 'ERROR SUMMARY|sanitize_small|Warning' gpurun_out/r3e_san_synccheck.txt | tr '\n' ' ')"
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/r3e_san_racecheck.txt 2>&1
echo "== racecheck: $(grep -E 'RACECHECK SUMMARY|sanitize_small' gpurun_out/r3e_san_racecheck.txt | tr '\n' ' ')"
grep -E "Error: Race" gpurun_out/r3e_san_racecheck.txt | cut -c1-170 | sort | uniq -c | sort -rn | head -12

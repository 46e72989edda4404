#!/bin/bash
# GPU round r4b: synccheck on every kernel but the var-len one (whose two false reports abort the tool's process)
set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool synccheck --num-cuda-barriers 4194304 python tools/sanitize_small.py > gpurun_out/r4b_san_synccheck.txt 2>&1
echo "== synccheck: $(grep -E 'ERROR SUMMARY|sanitize_small' gpurun_out/r4b_san_synccheck.txt | tr '\n' ' ' | cut -c1-400)"
grep -E "Barrier error|error detected|Missing|Divergent" gpurun_out/r4b_san_synccheck.txt | sort | uniq -c | head
grep -E "^=========     at " gpurun_out/r4b_san_synccheck.txt | sort | uniq -c | head

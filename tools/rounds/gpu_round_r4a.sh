#!/bin/bash
# GPU round r4a: synccheck (with room for the mbarriers) and racecheck on the small invocations, final build
set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool synccheck --num-cuda-barriers 65536 python tools/sanitize_small.py > gpurun_out/r4a_san_synccheck.txt 2>&1
echo "== synccheck: $(grep -E 'ERROR SUMMARY|sanitize_small' gpurun_out/r4a_san_synccheck.txt | tr '\n' ' ')"
grep -E "Barrier error|error detected|Missing|Divergent" gpurun_out/r4a_san_synccheck.txt | sort | uniq -c | head
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/r4a_san_racecheck.txt 2>&1
echo "== racecheck: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small' gpurun_out/r4a_san_racecheck.txt | tr '\n' ' ')"
grep -E "hazard detected" gpurun_out/r4a_san_racecheck.txt | sed 's/at .* in //' | sort | uniq -c | sort -rn | head -20

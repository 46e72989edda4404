#!/bin/bash
# GPU round r4i: racecheck on the encoder alone; the encoder tests five times over (timing-dependent failures would show)
set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_encoder.py > gpurun_out/r4i_race_encoder.txt 2>&1
echo "== racecheck: $(grep -E 'RACECHECK SUMMARY|sanitize_encoder' gpurun_out/r4i_race_encoder.txt | tr '\n' ' ')"
grep -E "^========= Error: Race reported between" gpurun_out/r4i_race_encoder.txt | sed -E 's/.*Write access at (void |bool |float |int )?(asp::)?([a-zA-Z0-9_:]+).*/\3/' | sort | uniq -c | sort -rn | head
for i in 1 2 3 4 5; do timeout 300 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu 2>&1 | tail -1; done

#!/bin/bash
# Final verification of round 2: parity suite, smoke, sanitizers on small invocations of every kernel, judged bench (default
# flags) + reference arm, launch list of the bench command, ncu --set full of the dominant kernel at the bench's launch size
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --durations=6 > gpurun_out/r2fin_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2fin_pytest.txt
grep -E "FAIL|passed|failed|exit|Error" gpurun_out/r2fin_pytest.txt | cut -c1-250 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2fin_smoke.txt 2>&1; tail -1 gpurun_out/r2fin_smoke.txt
for tool in memcheck synccheck; do
  # (synccheck needs room for the kernels' mbarriers: without --num-cuda-barriers its tracking table overflows and a later
  #  launch fails under the tool)
  extra=""; [ $tool = synccheck ] && extra="--num-cuda-barriers 65536"
  timeout 900 compute-sanitizer --tool $tool $extra python tools/sanitize_small.py > gpurun_out/r2fin_san_$tool.txt 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|sanitize_small' gpurun_out/r2fin_san_$tool.txt | tr '\n' ' ')"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2fin_bench_ref.txt 2>&1
tail -1 gpurun_out/r2fin_bench_ref.txt | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2fin_bench.txt 2>&1
tail -1 gpurun_out/r2fin_bench.txt | cut -c1-2500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2fin_launches.csv python bench.py --steps 5 --warmup 3 --cpu-seconds 1 --no-side > gpurun_out/r2fin_ncu_bench.log 2>&1
grep -c ot_fused gpurun_out/r2fin_launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ot_fused_v7 -s 4 -c 1 -o gpurun_out/r2fin_fused python tools/quick_bench.py 256000 > gpurun_out/r2fin_ncu.log 2>&1
ncu -i gpurun_out/r2fin_fused.ncu-rep --page raw --csv > gpurun_out/r2fin_fused_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r2fin_fused_raw.csv | head -8
timeout 200 python tools/encoder_bench.py > gpurun_out/r2fin_encoder_bench.txt 2>&1; cat gpurun_out/r2fin_encoder_bench.txt

#!/bin/bash
# GPU round r3c: full GPU suite, smoke, sanitizers on small invocations of every kernel, judged bench (both arms)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --durations=5 > gpurun_out/r3c_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r3c_pytest.txt
grep -E "FAIL|passed|failed|exit|Error" gpurun_out/r3c_pytest.txt | cut -c1-250 | tail -6
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3c_smoke.txt 2>&1; tail -1 gpurun_out/r3c_smoke.txt
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/r3c_san_$tool.txt 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|sanitize_small' gpurun_out/r3c_san_$tool.txt | tr '\n' ' ')"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3c_bench_ref.txt 2>&1; tail -1 gpurun_out/r3c_bench_ref.txt | cut -c1-300
timeout 900 python bench.py > gpurun_out/r3c_bench.txt 2>&1
tail -1 gpurun_out/r3c_bench.txt | cut -c1-1200

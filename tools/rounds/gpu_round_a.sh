#!/bin/bash
# One GPU-box visit: ISA microbench, GPU parity tests, developer bench, ncu capture of the fused kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
tools/ubench > gpurun_out/a_ubench.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/a_pytest.txt
timeout 600 python tools/quick_bench.py 1000 8000 64000 256000 > gpurun_out/a_quick.txt 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/a_bench.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ot_fused -s 4 -c 1 -o gpurun_out/a_fused python tools/quick_bench.py 64000 > gpurun_out/a_ncu.log 2>&1
tail -5 gpurun_out/a_pytest.txt; cat gpurun_out/a_ubench.txt gpurun_out/a_quick.txt; tail -2 gpurun_out/a_bench.txt

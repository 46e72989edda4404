#!/bin/bash
# GPU round r3w: probe of tcgen05.mma with A in tensor memory
mkdir -p gpurun_out
timeout 60 tools/ubench4 | tee gpurun_out/r3w_ts_probe.txt

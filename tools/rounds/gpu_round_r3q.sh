#!/bin/bash
# GPU round r3q: LDS.128 / LDS.64 wavefront cost by lane -> address map
mkdir -p gpurun_out
timeout 120 tools/ubench3 | tee gpurun_out/r3q_lds_wavefronts.txt

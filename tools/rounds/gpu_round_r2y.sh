#!/bin/bash
# GPU round r2y: kernel durations of the top-k path (ncu launch list)
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,dram__bytes_read.sum,smsp__inst_executed.sum --clock-control none -k regex:topk -c 8 --csv --log-file gpurun_out/r2y_topk.csv python tools/side_bench.py topk > gpurun_out/r2y_log.txt 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2y_topk.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]
for r in rows[hdr+1:]:
    d=dict(zip(h,r))
    print(d['Kernel Name'][:40], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
PY

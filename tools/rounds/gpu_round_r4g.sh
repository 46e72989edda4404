#!/bin/bash
# GPU round r4g (as r3m, final build): OT parity after the step-table change; per-kernel launch list of the encoder at B=128 L=256 (bf16)
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4g_enc_launches.csv python tools/encoder_bench.py --shape=128,256 --prec=bf16 > gpurun_out/r4g_enc_log.txt 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r4g_enc_launches.csv")) if len(r) > 10 and r[0].isdigit()]
n = len(rows)
# the last forward = the last 1/7 of launches (2 warm + 5 timed)
per = n // 7
last = rows[-per:]
agg = collections.OrderedDict()
for r in last:
    k = r[4][:70]
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r[-1]) / 1e3
tot = sum(v[1] for v in agg.values())
print("launches", n, "per forward", per, "total us", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {v[0]:4d} {v[1]:10.1f} us {v[1]/tot:6.3f}")
# per-launch list of one middle layer
for r in last[len(last)//2 : len(last)//2 + 12]:
    print(r[4][:60], r[-1])
PY

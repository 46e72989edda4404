"""Summarise an `ncu --page source --csv` export: instruction mix by opcode and the hottest SASS ranges.
usage: python tools/ncu_src_summary.py src.csv [kernel-substring]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
# several kernels may be concatenated: split on "Kernel Name"
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    if len(sys.argv) > 2 and sys.argv[2] not in b["name"]:
        continue
    hdr = b["rows"][0]; data = b["rows"][1:]
    ix = {h: i for i, h in enumerate(hdr)}
    tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
    tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
    print("kernel", b["name"][:80], "SASS lines", len(data), "warp-inst", tot_inst, "samples", tot_samp)
    byop = collections.Counter(); sop = collections.Counter()
    for r in data:
        op = r[ix["Source"]].split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
        op = op.split(".")[0]
        byop[op] += int(r[ix["Instructions Executed"]]); sop[op] += int(r[ix["# Samples"]])
    print("  opcode           warp-inst   %inst  %samples")
    for op, n in byop.most_common(28):
        print(f"  {op:14s} {n:12d}  {100*n/tot_inst:5.1f}  {100*sop[op]/max(tot_samp,1):5.1f}")
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    st = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stall_cols}
    print("  stalls:", ", ".join(f"{k[6:]}={100*v/max(tot_samp,1):.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    # hottest lines
    top = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]
    for r in top:
        print(f"   {r[ix['# Samples']]:>6s} smp  {r[ix['Instructions Executed']]:>9s} inst  {r[ix['Source']].strip()[:90]}")

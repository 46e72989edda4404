"""Per-kernel SASS opcode histogram of the in-tree library (cuobjdump -sass): the instructions that prove which machine
features a kernel uses -- UTCHMMA (tcgen05.mma), UTMALDG (TMA load), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit),
LDGSTS (cp.async), HMMA (legacy mma.sync), FFMA2 / FADD2 (packed fp32), MUFU.  usage: python tools/sass_opcodes.py [lib.so]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "aspire_b200/libaspire_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTMALDG", "UBLKCP", "UTCBAR", "LDTM", "STTM", "LDGSTS", "HMMA", "FFMA2", "FADD2", "MUFU", "SYNCS", "USETMAXREG"]
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        hist[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        hist[cur]["total"] += 1
        for k in KEYS:
            if op.startswith(k):
                hist[cur][k] += 1
print(f"{'kernel':70s} " + " ".join(f"{k:>8s}" for k in ["total"] + KEYS))
for name, h in hist.items():
    print(f"{name[:70]:70s} " + " ".join(f"{h[k]:8d}" for k in ["total"] + KEYS))

"""Per-kernel SASS instruction counts of the built library -> profiles/sass_r02.txt

    python profiles/sass_counts.py > profiles/sass_r02.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "revs-admm_b200", "librevs_admm.so")
PICK = ("MUFU", "DFMA", "DADD", "DMUL", "REDUX", "LDGSTS", "DMMA", "SYNCS", "UTMALDG", "LDTM", "UTCHMMA", "UTCBAR", "HMMA", "CCTL")

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, cur = collections.OrderedDict(), None
for line in out.split("\n"):
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        kern[cur]["_total"] += 1
        op = m.group(1).split(".")[0]
        if op in PICK:
            kern[cur][op] += 1
names = subprocess.run(["cu++filt"] + list(kern), capture_output=True, text=True).stdout.split("\n")
whole = collections.Counter()
for c in kern.values():
    whole.update({k: v for k, v in c.items() if k != "_total"})
print("cuobjdump -sass revs-admm_b200/librevs_admm.so (sm_100a), instruction counts per kernel: total, selected mnemonics")
print("UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld (TMEM), DMMA = FP64 mma.sync, REDUX = warp reduce, LDGSTS = cp.async")
print("whole library:", dict(whole))
print()
for (k, c), n in sorted(zip(kern.items(), names), key=lambda x: -x[0][1]["_total"]):
    print(f"{c['_total']:7d} instr  {dict((a, b) for a, b in c.items() if a != '_total')}  {n[:110]}")

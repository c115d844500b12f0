#!/usr/bin/env python
"""Per-kernel SASS instruction histogram of qfa_b200/libqfa_b200.so (evidence that the production kernels use the
Blackwell tensor path: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, UTMALDG = TMA tensor load,
UTCBAR = tcgen05.commit; STL/LDL = local-memory spills).   python scripts/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "qfa_b200", "libqfa_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
ops = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU", "LDG", "STG", "LDS", "STS",
       "STL", "LDL", "HMMA", "DMMA", "DFMA", "FFMA", "BAR"]
cur, hist, order, total = None, collections.defaultdict(collections.Counter), [], collections.Counter()
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        full = next(it)
        mm = re.search(r"((?:aux::)?k_\w+)(<[^>]*>)?", full)
        cur = (mm.group(1) + (mm.group(2) or "")).replace("(int)", "").replace("(bool)", "") if mm else full[:58]
        while cur in hist or cur in order:
            cur += "'"
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for o in ops:
            if op == o or op.startswith(o + "."):
                hist[cur][o] += 1
print("SASS instruction histogram of", os.path.relpath(lib, ROOT), "(cuobjdump -sass, sm_100a)")
print("%-58s %7s " % ("kernel", "instr") + " ".join("%7s" % o for o in ops))
for k in order:
    if not re.search(r"k_|aux", k):
        continue
    print("%-58s %7d " % (k[-58:], total[k]) + " ".join("%7d" % hist[k][o] for o in ops))

"""Summarise `ncu --page source --csv` output: per kernel, stall-reason totals and instruction-class mix (SASS view)."""
import csv
import gzip
import re
import sys
from collections import Counter, OrderedDict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
kernels = OrderedDict()
cur, hdr = None, None
for row in csv.reader(f):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = row[1]
        n = 2
        while cur in kernels:
            cur = row[1] + " #%d" % n
            n += 1
        kernels[cur] = []
        hdr = None
    elif row[0] == "Address":
        hdr = row
    elif cur is not None and hdr is not None:
        kernels[cur].append(dict(zip(hdr, row)))
for name, rows in kernels.items():
    tot_inst = sum(int(r["Instructions Executed"] or 0) for r in rows)
    samples = sum(int(r["# Samples"] or 0) for r in rows)
    stalls = Counter()
    for r in rows:
        for k, v in r.items():
            if k.startswith("stall_") and "Not Issued" not in k and v:
                stalls[k] += int(v)
    mix = Counter()
    for r in rows:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r["Source"])
        op = m.group(2) if m else "?"
        mix[op] += int(r["Instructions Executed"] or 0)
    print("==", name[:110])
    print("   warp-instructions %d, samples %d" % (tot_inst, samples))
    print("   stalls:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(1, sum(stalls.values()))) for k, v in stalls.most_common(8)))
    print("   mix   :", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, tot_inst)) for k, v in mix.most_common(top)))
    hot = sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:top]
    for r in hot:
        print("      %6s samples  %-60s" % (r["# Samples"], r["Source"].strip()[:60]))

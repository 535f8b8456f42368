#!/usr/bin/env python
"""Per-instruction stall samples of an .ncu-rep (source page): totals per stall reason and the top instructions of a reason.

    python tools/ncu_hot.py REP [reason] [top]
"""
import csv, subprocess, sys
rep = sys.argv[1]
reason = (sys.argv[2] or None) if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(text.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
head = rows[hi]
data = [dict(zip(head, r)) for r in rows[hi + 1:] if len(r) == len(head)]
stalls = [h for h in head if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(int(d[s] or 0) for d in data) for s in stalls}
allsamp = sum(int(d["# Samples"] or 0) for d in data)
print("samples", allsamp, "instructions executed", sum(int(d["Instructions Executed"] or 0) for d in data))
for s, v in sorted(tot.items(), key=lambda x: -x[1]):
    if v: print("  %-24s %8d %5.1f%%" % (s, v, 100.0 * v / max(allsamp, 1)))
key = reason or "# Samples"
print("top by", key)
for i, d in sorted(enumerate(data), key=lambda x: -int(x[1][key] or 0))[:top]:
    print("  #%-5d %6s  exec %9s  %s" % (i, d[key], d["Instructions Executed"], d["Source"].strip()[:90]))

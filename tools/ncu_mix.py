#!/usr/bin/env python
"""Executed-instruction mix (per SASS opcode) of the kernel in an .ncu-rep captured with --import-source on.

    python tools/ncu_mix.py REP [top]
"""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(text.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
head = rows[hi]
data = [dict(zip(head, r)) for r in rows[hi + 1:] if len(r) == len(head)]
tot = sum(int(d["Instructions Executed"] or 0) for d in data)
ops = collections.Counter()
for d in data:
    src = d["Source"].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = '.'.join((m.group(2) if m else src[:10]).split('.')[:2])
    ops[op] += int(d["Instructions Executed"] or 0)
print("warp instructions executed", tot)
for op, n in ops.most_common(top):
    print("%-22s %11d %5.1f%%" % (op, n, 100 * n / tot))

#!/usr/bin/env python
"""Opcode mix of the instructions whose executed count equals (or is within tol of) the hottest loop's trip count.

    python tools/ncu_loop.py REP [exec_count]"""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(text.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
head = rows[hi]
data = [dict(zip(head, r)) for r in rows[hi + 1:] if len(r) == len(head)]
cnt = collections.Counter(int(d["Instructions Executed"] or 0) for d in data)
print("most common exec counts:", sorted(cnt.items(), key=lambda x: -x[0] * x[1])[:12])
target = int(sys.argv[2]) if len(sys.argv) > 2 else max(cnt.items(), key=lambda x: x[0] * x[1])[0]
sel = [d for d in data if int(d["Instructions Executed"] or 0) == target]
ops = collections.Counter()
for d in sel:
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', d["Source"].strip())
    ops['.'.join(m.group(2).split('.')[:2])] += 1
print("exec", target, ":", len(sel), "static instructions")
print(", ".join(f"{k}:{v}" for k, v in ops.most_common()))

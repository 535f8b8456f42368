#!/usr/bin/env python
"""Executed warp instructions grouped by per-instruction execution count (each group = one loop / code region).

    python tools/ncu_groups.py REP [top]"""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(text.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
head = rows[hi]
data = [dict(zip(head, r)) for r in rows[hi + 1:] if len(r) == len(head)]
tot = sum(int(d["Instructions Executed"] or 0) for d in data)
g = collections.defaultdict(list)
for i, d in enumerate(data):
    g[int(d["Instructions Executed"] or 0)].append(i)
print("total", tot)
for cnt, idx in sorted(g.items(), key=lambda x: -x[0] * len(x[1]))[:top]:
    ops = collections.Counter()
    for i in idx:
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', data[i]["Source"].strip())
        ops[m.group(2).split('.')[0]] += 1
    print("exec %9d x %4d = %6.1fM (%4.1f%%) idx %d..%d  %s" % (cnt, len(idx), cnt * len(idx) / 1e6, 100.0 * cnt * len(idx) / tot, idx[0], idx[-1],
                                                      ", ".join(f"{k}:{v}" for k, v in ops.most_common(8))))

#!/usr/bin/env python
"""Opcode histogram of every ring kernel in libj2kb200.so (cuobjdump -sass): the record that shows which memory / math
instructions the kernels are made of (UBLKCP = 1-D TMA bulk copies, SYNCS = mbarrier, FFMA2 / FADD2 / FMUL2 = packed f32x2,
no UTMALDG / UTMASTG / tcgen05 by design).  python tools/sass_hist.py [lib] > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "go-dicom-codec_b200", "csrc", "build", "libj2kb200.so")
text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
demangle = subprocess.run(["c++filt"] + list(hist), capture_output=True, text=True).stdout.splitlines()
print("# opcode histogram per kernel:", os.path.relpath(lib, ROOT), "(cuobjdump -sass, sm_100a)")
tot = collections.Counter()
for (k, h), name in zip(hist.items(), demangle):
    if "ring_kernel" not in k and "inv3w" not in k:
        continue
    n = sum(h.values())
    tot.update(h)
    print(f"\n{name}\n  {n} instructions: " + ", ".join(f"{op}:{c}" for op, c in h.most_common(28)))
print("\nall ring kernels: " + ", ".join(f"{op}:{c}" for op, c in tot.most_common(60)))
for op in ("UBLKCP", "SYNCS", "UTMALDG", "UTMASTG", "FFMA2", "FADD2", "FMUL2", "DMUL", "DADD", "ELECT", "STG", "LDG", "LDS", "STS", "SHFL", "F2I", "I2FP", "PRMT"):
    print(f"  {op}: {tot.get(op, 0)}")

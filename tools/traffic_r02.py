#!/usr/bin/env python
"""profiles/traffic.json from the round's --set full captures: DRAM bytes (read + write) per launch of the dominant kernel of
every config, next to the algorithmic bytes of the same launch.  bench.py reads `roofline.traffic` and the per-config
`fwd_traffic` / `inv_traffic` from it."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def dram(name):
    """DRAM bytes (read + write) of the one launch in profiles/ncu_<tag>_<name>.json (written on the GPU box by tools/final_r02.sh)"""
    path = f"{ROOT}/profiles/ncu_{tag}_{name}.json"
    if not os.path.exists(path):
        return None
    k = json.load(open(path))["kernels"][0]
    tot = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(k[key]["value"]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[k[key]["unit"]]
    return tot


out = {"note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (ncu --set full, tools/final_r02.sh); "
               "algorithmic = SURVEY 8(d) bytes of the same launch"}
f, i = dram("fwd_ring"), dram("inv_ring")
if f:
    out["fwd_ring_dram_bytes_per_frame"] = f / 32
if i:
    out["inv_ring_dram_bytes_per_frame"] = i / 32
out["c2_algorithmic_bytes_per_frame"] = bench.alg_bytes_per_frame()
names = {"C1": "c1", "C3i": "rgb97", "C3ii": "rgb53", "C4": "c4", "C5": "c5", "DX": "dx", "CR": "cr"}
cfgs = {}
for cfg in bench.OTHER_CONFIGS:
    key, _, w, h, c, bits, _, L, _, frames, _ = cfg
    alg = bench.config_alg_bytes(frames * w * h * c, 1 if bits <= 8 else 2, L)
    cfgs[key] = {"fwd": dram("fwd_" + names[key]),
                 "inv": dram("inv_" + names[key]), "algorithmic": alg}
out["configs"] = cfgs
json.dump(out, open(f"{ROOT}/profiles/traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))

#!/usr/bin/env python
"""Resident throughput of the BASELINE.json configs other than the bench workload: one JSON line per config with forward /
inverse Mpixel/s and the achieved fraction of the HBM roofline (SURVEY 8d algorithmic bytes).  The same legs run inside
`bench.py` (`configs` in its JSON line); this driver is for kernel experiments:
    python tools/config_bench.py [--steps 10] [--only C3i]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
import j2kb200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default="", help="substring of the config key / name to run")
    ap.add_argument("--frames", type=int, default=0, help="override the frames per launch")
    a = ap.parse_args()
    peak, _ = bench.measured_peak()
    with j2kb200.Context(devices=[0]) as ctx:
        extra = [("C2", "C2 x16: 4096x4096 12-bit mono, 9/7 L6", 4096, 4096, 1, 12, False, 6, False, 16, (0, 0))]
        for cfg in bench.OTHER_CONFIGS + extra:
            if a.only not in cfg[0] + " " + cfg[1]:
                continue
            if a.frames:
                cfg = cfg[:9] + (a.frames,) + cfg[10:]
            d = bench.run_config(ctx, torch, cfg, a.steps, peak)
            d["fwd_frac_hbm"], d["inv_frac_hbm"] = d["fwd_frac"], d["inv_frac"]
            print(json.dumps(d), flush=True)


if __name__ == "__main__":
    main()

import sys, os, json
sys.path[:0] = ["/root/repo", "/root/repo/go-dicom-codec_b200", "/root/repo/tests", "/root/repo/tools"]
import config_bench as cb, j2kb200
with j2kb200.Context(devices=[0]) as ctx:
    for w in (4096, 4160, 4032, 2056, 1000):
        cb.run_config(ctx, "w%d" % w, w, 4096, 1, 12, False, 6, False, 32, steps=10)

#!/usr/bin/env python
"""The copy kernel MEASURED_PEAKS.json is defined by (torch b.copy_(a), 1 Gi bf16 elements, read + write bytes), held for
seconds: does the HBM copy peak itself move when the board sits at its power limit?  Prints GB/s per 0.5 s window with the
SM clock, board power and throttle reasons, then the widening copy (1 read : 2 write, the mix of forward level 1)."""
import json
import time

import pynvml
import torch

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
N = 1 << 30
a = torch.empty(N, dtype=torch.bfloat16, device="cuda"); b = torch.empty(N, dtype=torch.bfloat16, device="cuda")
a16 = torch.empty(N // 2, dtype=torch.int16, device="cuda"); b32 = torch.empty(N // 2, dtype=torch.int32, device="cuda")
a.normal_(); a16.random_(0, 1000)
for name, fn, nbytes in (("copy_1r_1w", lambda: b.copy_(a), N * 4), ("widen_1r_2w", lambda: b32.copy_(a16), (N // 2) * 6)):
    time.sleep(2.0)
    t0 = time.time()
    while time.time() - t0 < 5.0:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(400):
            fn()
        e1.record()
        time.sleep(0.15)
        mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); w = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        torch.cuda.synchronize()
        print(json.dumps({"kernel": name, "t": round(time.time() - t0, 1), "GBps": round(nbytes * 400 / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1),
                          "sm_mhz": mhz, "watts": round(w), "reasons": hex(r)}), flush=True)

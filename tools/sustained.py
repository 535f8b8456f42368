#!/usr/bin/env python
"""Sustained-load behaviour of the two ring kernels on C2 (32 frames): ms per step in consecutive 0.5 s windows together
with the SM clock, board power and throttle reasons NVML reports (what the 100-step bench region does not see).
    python tools/sustained.py [seconds per direction]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import pynvml  # noqa: E402
import torch  # noqa: E402

import j2kb200  # noqa: E402
from j2kb200 import abi  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
W = H = 4096; L = 6; bits = 12; B = 32
enc, _ = j2kb200.openjpeg_quant_params(L, bits)
es, ds = j2kb200.runtime_quant_steps(enc, L, bits), j2kb200.decode_quant_steps(enc, L, bits, False)
fp = abi.fwd_params(W, H, 1, bits, False, 0, 0, L, False, False, abi.MCT_NONE, es)
ip = abi.inv_params(W, H, 1, bits, False, 0, 0, L, False, False, abi.MCT_NONE, ds)
fb = W * H * 2
pynvml.nvmlInit()
hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
ctx = j2kb200.Context(devices=[0])
g = torch.Generator(device="cuda").manual_seed(7)
d_in = torch.randint(0, 256, (B, fb), dtype=torch.uint8, device="cuda", generator=g)
d_in.view(B, -1, 2)[:, :, 1] &= 15
d_co = [torch.empty((B, W * H), dtype=torch.int32, device="cuda") for _ in range(2)]
d_px = [torch.empty((B, fb), dtype=torch.uint8, device="cuda") for _ in range(2)]
st = [torch.cuda.Stream() for _ in range(2)]
ctx.forward_device(fp, B, d_in.data_ptr(), fb, d_co[0].data_ptr(), stream=st[0].cuda_stream)
torch.cuda.synchronize()


def window(fn, steps=500):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st[0]); st[1].wait_event(e0)
    for i in range(steps):
        fn(i)
    ev = torch.cuda.Event(); ev.record(st[1]); st[0].wait_event(ev)
    e1.record(st[0])
    time.sleep(0.2)  # sample while the queued steps run
    mhz = pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM)
    watts = pynvml.nvmlDeviceGetPowerUsage(hnd) / 1000.0
    reasons = pynvml.nvmlDeviceGetCurrentClocksEventReasons(hnd) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)
    temp = pynvml.nvmlDeviceGetTemperature(hnd, pynvml.NVML_TEMPERATURE_GPU)
    torch.cuda.synchronize()
    return {"ms_per_step": round(e0.elapsed_time(e1) / steps, 4), "sm_mhz": mhz, "watts": round(watts), "reasons": hex(reasons), "temp_c": temp}


for name, fn in (("forward", lambda i: ctx.forward_device(fp, B, d_in.data_ptr(), fb, d_co[i % 2].data_ptr(), stream=st[i % 2].cuda_stream)),
                 ("inverse", lambda i: ctx.inverse_device(ip, B, d_co[0].data_ptr(), d_px[i % 2].data_ptr(), fb, stream=st[i % 2].cuda_stream))):
    t0 = time.time()
    time.sleep(2.0)  # start each direction from an idle GPU
    while time.time() - t0 < secs + 2.0:
        print(json.dumps({"dir": name, "t": round(time.time() - t0 - 2.0, 1), **window(fn)}), flush=True)
ctx.close()

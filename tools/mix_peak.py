#!/usr/bin/env python
"""What plain streaming kernels reach on this GPU for the read:write mixes of the DWT levels (library kernels, best of 10):
a same-size copy (1:1, the MEASURED_PEAKS.json definition), a widening copy int16 -> int32 (1 read : 2 write, the
mix of forward level 1), a narrowing copy int32 -> int16 (2:1, inverse level 1) and a pure fill."""
import json
import torch

def best(fn, nbytes, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return nbytes / (min(t) * 1e-3) / 1e9

N = 1 << 29  # elements
a32 = torch.empty(N, dtype=torch.int32, device="cuda"); b32 = torch.empty(N, dtype=torch.int32, device="cuda")
a16 = torch.empty(N, dtype=torch.int16, device="cuda"); b16 = torch.empty(N, dtype=torch.int16, device="cuda")
a32.random_(0, 1000); a16.random_(0, 1000)
out = {
    "copy_1r_1w_GBps": best(lambda: b32.copy_(a32), N * 8),
    "widen_1r_2w_GBps": best(lambda: b32.copy_(a16), N * 6),
    "narrow_2r_1w_GBps": best(lambda: b16.copy_(a32), N * 6),
    "fill_0r_1w_GBps": best(lambda: b32.fill_(7), N * 4),
    "elements": N,
}
print(json.dumps(out))

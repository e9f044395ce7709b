#!/usr/bin/env python
"""Write-only bandwidth of this B200 (developer tool): what a store-only kernel such as pp_dubins_fill_kernel can hope
for.  MEASURED_PEAKS.json's HBM figure is a COPY (read + write bytes); a pure store stream has its own ceiling.
  gpurun -- 'python tools/store_peak.py > gpurun_out/store_peak.json'"""
import json

import torch

n = 1_233_000_000 // 8  # the fill kernel's 1.23 GB of samples
x = torch.empty(n, dtype=torch.float64, device="cuda")
y = torch.empty(n, dtype=torch.float64, device="cuda")
out = {}
for name, fn in [("fill_", lambda: x.fill_(1.0)), ("zero_", lambda: x.zero_()), ("copy_", lambda: y.copy_(x))]:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    bytes_moved = n * 8 * (2 if name == "copy_" else 1)
    out[name] = {"ms": best, "GB_per_s": bytes_moved / best / 1e6}
print(json.dumps(out))

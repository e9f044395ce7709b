"""Dubins evaluate: device-resident time per call against the batch size (run under gpurun):
python tools/eval_scaling.py  ->  gpurun_out/eval_scaling.json"""
import json
import sys

sys.path.insert(0, ".")
import torch

import __graft_entry__ as g

pp = g.import_package()
ctx = pp.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
N = 1 << 24
arrs = [torch.from_numpy(a).to(dev) for a in pp.synth.dubins_pairs(N)]
cost = torch.empty(N, dtype=torch.float64, device=dev)
word = torch.empty(N, dtype=torch.uint8, device=dev)
rows = []
for lg in range(8, 25, 2):
    n = 1 << lg
    reps = max(10, min(2000, (1 << 26) // n))
    with torch.cuda.stream(stream):
        for _ in range(5):
            ctx.dubins_eval_dev(n, *arrs, 1.0, cost, word)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            ctx.dubins_eval_dev(n, *arrs, 1.0, cost, word)
        e1.record(stream)
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    rows.append({"pairs": n, "us_per_call": us, "pairs_per_s": n / (us * 1e-6)})
    print(rows[-1], flush=True)
json.dump(rows, open("gpurun_out/eval_scaling.json", "w"))

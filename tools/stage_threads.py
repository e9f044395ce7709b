"""Sweep of PP_STAGE_THREADS for the staged (pageable-memory) path of pp_dubins_eval: one context per setting (the pool
is sized when a context first stages), 2^24 pairs of the C3 workload in plain numpy arrays, best of 3 calls.
Usage (GPU box): python tools/stage_threads.py > gpurun_out/stage_threads.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

pp = graft.import_package()
n = 1 << 24
host = pp.synth.dubins_pairs(n, "mixed")
cost, word = np.empty(n, np.float64), np.empty(n, np.uint8)
res = {"host_threads": os.cpu_count(), "pairs": n, "runs": []}
ref = None
for t in (2, 3, 4, 6, 8, 12, 16):
    os.environ["PP_STAGE_THREADS"] = str(t)
    ctx = pp._ffi.Context(0)
    ctx.dubins_eval(*host, radius=1.0, want_tpq=False, out=(cost, word, None))  # warm-up: ring + pool + first touch
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.dubins_eval(*host, radius=1.0, want_tpq=False, out=(cost, word, None))
        best = min(best, time.perf_counter() - t0)
    chk = (float(cost.sum()), int(word.astype(np.int64).sum()))
    ref = ref or chk
    res["runs"].append({"threads": t, "ms": best * 1e3, "pairs_per_s": n / best, "same_result": chk == ref})
    ctx.close()
print(json.dumps(res))

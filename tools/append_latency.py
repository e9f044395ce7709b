#!/usr/bin/env python
"""Append + query latency against the tree size (developer tool).

The reference inserts ONE node per iteration (src/rrt.rs:586-589) and queries the tree right after (:378-391).  With
the incremental node grid an append only extends the linearly scanned tail; the O(n) rebuild runs once per 4 096
appended nodes.  For n = 2^12 ... 2^20: wall-clock microseconds per (append 1 + query 1) iteration over 1 024
iterations, and per (append 512 + query 512) round over 32 rounds, with the number of grid rebuilds each took.
  gpurun -- 'python tools/append_latency.py > gpurun_out/append_latency.json'"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

pp = graft.import_package()
ctx = pp.Context(0)
rng = np.random.default_rng(1)
rows = []
for lg in range(12, 21, 2):
    n = 1 << lg
    x, y = rng.uniform(0, 1000, n + 40_000), rng.uniform(0, 1000, n + 40_000)
    ctx.tree_upload(x[:n], y[:n])
    ctx.nn(x[:200], y[:200], want_d2=False)  # builds the grid
    q = rng.uniform(0, 1000, (2, 2048))
    b0 = ctx.nn_grid_builds
    t0 = time.perf_counter()
    for k in range(1024):
        ctx.tree_append(x[n + k:n + k + 1], y[n + k:n + k + 1])
        ctx.nn(q[0][k:k + 1], q[1][k:k + 1], want_d2=False)
    t1 = time.perf_counter()
    b1 = ctx.nn_grid_builds
    base = n + 1024
    for r in range(32):
        ctx.tree_append(x[base + 512 * r:base + 512 * (r + 1)], y[base + 512 * r:base + 512 * (r + 1)])
        ctx.nn(q[0][:512], q[1][:512], want_d2=False)
    t2 = time.perf_counter()
    b2 = ctx.nn_grid_builds
    rows.append({"nodes": n, "us_per_append1_query1": (t1 - t0) / 1024 * 1e6, "grid_rebuilds_in_1024_appends": b1 - b0,
                 "us_per_append512_query512": (t2 - t1) / 32 * 1e6, "grid_rebuilds_in_32_rounds": b2 - b1})
print(json.dumps({"what": "host wall clock through the ctypes binding (includes ~10 us of Python per call)", "rows": rows}))

"""Small end-to-end pass over the kernels that use shared-memory staging / device-side work lists (sample fill, plan,
Dubins verify, extend step) for compute-sanitizer:

  compute-sanitizer --tool memcheck  python tools/sanitize_probe.py
  compute-sanitizer --tool racecheck python tools/sanitize_probe.py

Sizes are tiny (the tools slow kernels down 10-100x).  (compute-sanitizer is closed on the pool this was written on:
the script then simply serves as a quick cross-check -- batch vs scalar entry points, culled vs exhaustive verdicts.)"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pp = g.import_package()
ctx = pp.Context(0)
rng = np.random.default_rng(5)
n = 700
sx, sy, syaw, ex, ey, eyaw = pp.synth.dubins_pairs(n, "mixed", seed=3)
ex[:50] = sx[:50] + 40.0  # a few long paths: several rows, the coefficient table refresh
out, offsets, counts = pp.dubins.batch_paths(sx, sy, syaw, ex, ey, eyaw, 1.0, 0.05, ctx=ctx)
assert int(counts.sum()) == out.shape[0] and np.isfinite(out).all()
bounds, rings = pp.synth.circle_world(300, world=120.0, rmin=0.5, rmax=2.0)
ctx.obstacles_upload(bounds, rings)
qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(3000, 5000, world=120.0)
ctx.tree_upload(nx, ny, nyaw)
idx, yaw, ok = ctx.rrt_extend_dubins(qx, qy, 0.8, 0.1)
idx2, yaw2, ok2 = ctx.rrt_extend(qx, qy)
assert np.array_equal(idx, idx2)
e = pp.synth.dubins_edges(600, world=120.0, reach=40.0)
a = ctx.collide_dubins(*e, 1.0, 0.05)
b = ctx.collide_dubins(*e, 1.0, 0.05, flags=1)
assert np.array_equal(a, b)
c = ctx.collide_dubins(qx, qy, yaw, nx[idx], ny[idx], nyaw[idx], 0.8, 0.1)
assert np.array_equal(c, ok)
print("probe ok: samples", out.shape[0], "extend free", float(ok.mean()), "long edges free", float(a.mean()))
ctx.close()

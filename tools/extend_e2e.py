"""The host-pointer extend step (pp_rrt_extend) on the C4 workload: the chunk-pipelined default route against the
single-stream route (selected here by asking for the scan verify), same answers, and the end-to-end time per call on
pinned and on pageable arrays.  Usage (GPU box): python tools/extend_e2e.py > gpurun_out/extend_e2e.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

pp = graft.import_package()
F = pp._ffi
m = n_nodes = 1 << 20
qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(m, n_nodes)
bounds, rings = pp.synth.circle_world(10_000)
ctx = F.Context(0)
ctx.obstacles_upload(bounds, rings)
ctx.tree_upload(nx, ny, nyaw)
ref = ctx.rrt_extend(qx, qy, collide_flags=F.COLLIDE_SCAN)  # one stream: upload, grid NN, binned scan verify, download
res = {"queries": m, "nodes": n_nodes, "rings": 10_000, "free_fraction": float(ref[2].mean())}


def leg(aq, out, reps=10):
    ctx.rrt_extend(aq[0], aq[1], out=out)
    best, t_all = 1e9, time.perf_counter()
    for _ in range(reps):
        t0 = time.perf_counter()
        ctx.rrt_extend(aq[0], aq[1], out=out)
        best = min(best, time.perf_counter() - t0)
    avg = (time.perf_counter() - t_all) / reps
    same = bool(np.array_equal(out[0], ref[0]) and np.array_equal(out[2], ref[2])
                and np.array_equal(out[1], ref[1], equal_nan=True))
    return {"ms_avg": avg * 1e3, "ms_best": best * 1e3, "steps_per_s": m / avg, "same_as_single_stream_route": same}


pins = [F.PinnedArray(m, np.float64) for _ in range(2)]
pins[0].array[:] = qx
pins[1].array[:] = qy
po = (F.PinnedArray(m, np.uint32), F.PinnedArray(m, np.float64), F.PinnedArray(m, np.uint8))
res["pinned"] = leg([p.array for p in pins], tuple(p.array for p in po))
res["pageable"] = leg([qx, qy], (np.empty(m, np.uint32), np.empty(m, np.float64), np.empty(m, np.uint8)))
# ragged sizes around the chunk boundaries (2^17 queries per chunk; the pipelined route starts at 2 chunks)
ragged = {}
for k in ((1 << 18) - 1, 1 << 18, (1 << 18) + 1, 3 * (1 << 17) + 777, (1 << 19) + 5):
    o = ctx.rrt_extend(qx[:k], qy[:k])
    ragged[str(k)] = bool(np.array_equal(o[0], ref[0][:k]) and np.array_equal(o[2], ref[2][:k])
                          and np.array_equal(o[1], ref[1][:k], equal_nan=True))
res["ragged_sizes_same"] = ragged
# developer sweep of the chunk size (PP_EXTEND_CHUNK_LOG2 is read per call); 21 = one chunk = the single-stream route
sweep = {}
for lg in (15, 16, 17, 18, 19, 21):
    os.environ["PP_EXTEND_CHUNK_LOG2"] = str(lg)
    sweep[str(lg)] = leg([p.array for p in pins], tuple(p.array for p in po), reps=20)
os.environ.pop("PP_EXTEND_CHUNK_LOG2")
res["chunk_log2_sweep_pinned"] = sweep
print(json.dumps(res))

import sys, json, time
sys.path.insert(0, '.')
import numpy as np, torch
import __graft_entry__ as g
pp = g.import_package()
ctx = pp.Context(0)
dev = torch.device('cuda:0')
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
bounds, rings = pp.synth.circle_world(10_000)
ctx.obstacles_upload(bounds, rings)
m = 1 << 20
qx = torch.from_numpy(pp.synth.uniform(pp.synth.SEED_C4_Q, 0, m, 0.0, 1000.0)).to(dev)
qy = torch.from_numpy(pp.synth.uniform(pp.synth.SEED_C4_Q, 1, m, 0.0, 1000.0)).to(dev)
idx = torch.empty(m, dtype=torch.int32, device=dev); yaw = torch.empty(m, dtype=torch.float64, device=dev); ok = torch.empty(m, dtype=torch.uint8, device=dev)
rows = []
for lg in (12, 14, 16, 18, 20, 22, 24):
    n = 1 << lg
    _, _, nx, ny, nyaw = pp.synth.extend_inputs(1, n)
    t = torch.from_numpy(np.stack([nx, ny, nyaw])).to(dev)
    ctx.tree_upload_dev(n, t[0], t[1], t[2])
    ctx.timing_enable(True)
    with torch.cuda.stream(stream):
        ctx.rrt_extend_dev(m, qx, qy, idx, yaw, ok); torch.cuda.synchronize()
        b = ctx.timing_get("nn_grid_build"); ctx.timing_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10): ctx.rrt_extend_dev(m, qx, qy, idx, yaw, ok)
        e1.record(stream); torch.cuda.synchronize()
    nn = ctx.timing_get("nn_grid"); cs = ctx.timing_get("collide_segments_grid"); ctx.timing_enable(False)
    rows.append({"nodes": n, "step_ms": e0.elapsed_time(e1) / 10, "nn_ms": nn[0] / max(nn[1], 1), "collide_ms": cs[0] / max(cs[1], 1), "grid_build_ms": b[0] / max(b[1], 1), "free": float(ok.float().mean())})
    print(rows[-1], flush=True)
json.dump(rows, open('gpurun_out/nn_scaling.json', 'w'))

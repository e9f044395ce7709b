"""GPU: the multi-GPU part of the C-ABI (include/pathplanning_b200.h "multi-GPU"): pp_group (one host
process, one context + worker per device), pp_ctx_comm_init (one process per GPU), replication of tree / appended tail /
obstacles by ncclBroadcast, contiguous slicing of batches.  SURVEY section 4: outputs must be BYTE-IDENTICAL for every
device count.  Multi-device cases skip on a 1-GPU box (gpurun --gpus 2 runs them)."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs(pp, n_pairs=200_001, m=50_003, n_nodes=30_000, e=4001):
    pairs = pp.synth.dubins_pairs(n_pairs)
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(m, n_nodes, world=300.0)
    bounds, rings = pp.synth.circle_world(900, world=300.0)
    edges = pp.synth.dubins_edges(e, world=300.0, reach=15.0)
    return pairs, (qx, qy, nx, ny, nyaw), (bounds, rings), edges


def _reference_answers(pp, pairs, ext, world, edges):
    ctx = pp.Context(0)
    qx, qy, nx, ny, nyaw = ext
    ctx.tree_upload(nx, ny, nyaw)
    ctx.obstacles_upload(*world)
    out = {
        "eval": ctx.dubins_eval(*pairs, radius=1.0),
        "nn": ctx.nn(qx, qy),
        "extend": ctx.rrt_extend(qx, qy),
        "extend_dubins": ctx.rrt_extend_dubins(qx[:9001], qy[:9001], 0.8, 0.1),
        "collide_dubins": ctx.collide_dubins(*edges, 1.0, 0.05),
        "collide_segments": ctx.collide_segments(qx, qy, qx + 3.0, qy - 2.0),
    }
    ctx.close()
    return out


def _same(a, b):
    if isinstance(a, tuple):
        return all(_same(x, y) for x, y in zip(a, b))
    if a is None or b is None:
        return a is b
    return np.array_equal(a, b, equal_nan=True)


def _check_group(pp, devices, pairs, ext, world, edges, want):
    g = pp.Group(devices)
    assert len(g) == len(devices)
    qx, qy, nx, ny, nyaw = ext
    g.tree_upload(nx, ny, nyaw)
    g.obstacles_upload(*world)
    for i in range(len(g)):
        assert g.ctx(i).tree_size == nx.size
    assert _same(g.dubins_eval(*pairs, radius=1.0), want["eval"])
    assert _same(g.nn(qx, qy), want["nn"])
    assert _same(g.rrt_extend(qx, qy), want["extend"])
    assert _same(g.rrt_extend_dubins(qx[:9001], qy[:9001], 0.8, 0.1), want["extend_dubins"])
    assert _same(g.collide_dubins(*edges, 1.0, 0.05), want["collide_dubins"])
    assert _same(g.collide_segments(qx, qy, qx + 3.0, qy - 2.0), want["collide_segments"])
    # tiny batches: fewer items than devices, and an empty one
    assert _same(g.nn(qx[:1], qy[:1]), (want["nn"][0][:1], want["nn"][1][:1]))
    assert g.nn(qx[:0], qy[:0])[0].size == 0
    g.close()


def test_group_of_one_device_equals_the_plain_context(pp):
    pairs, ext, world, edges = _inputs(pp)
    want = _reference_answers(pp, pairs, ext, world, edges)
    _check_group(pp, [0], pairs, ext, world, edges, want)


def test_group_outputs_are_byte_identical_for_every_device_count(pp):
    n = pp.device_count()
    if n < 2:
        pytest.skip("needs >= 2 B200s (gpurun --gpus 2)")
    pairs, ext, world, edges = _inputs(pp)
    want = _reference_answers(pp, pairs, ext, world, edges)
    for g in sorted({2, n}):
        _check_group(pp, list(range(g)), pairs, ext, world, edges, want)


def test_group_tree_append_broadcasts_the_tail(pp, O):
    """src/rrt.rs:586-589: inserts reach every replica (tail-only ncclBroadcast), NN stays bit-exact on every device"""
    n = pp.device_count()
    devs = list(range(min(n, 4)))
    g = pp.Group(devs)
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(4000, 20_000, world=100.0)
    par = np.maximum(np.arange(nx.size, dtype=np.int32) - 1, -1)
    g.tree_upload(nx[:5000], ny[:5000], nyaw[:5000], par[:5000])
    for a, b in [(5000, 5001), (5001, 5513), (5513, 9000), (9000, 20_000)]:
        g.tree_append(nx[a:b], ny[a:b], nyaw[a:b], par[a:b])
        want = O.nn_brute(nx[:b], ny[:b], qx, qy)[0]
        assert np.array_equal(g.nn(qx, qy, want_d2=False), want)
        for i in range(len(g)):  # every replica answers the whole batch alone, identically
            assert g.ctx(i).tree_size == b
            assert np.array_equal(g.ctx(i).nn(qx[:700], qy[:700], want_d2=False), want[:700])
    g.close()


def _rank_main(rank, world, comm_id, outdir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as graft
    pp = graft.import_package()
    ctx = pp.Context(rank)
    ctx.comm_init(comm_id, world, rank)
    assert (ctx.comm_rank, ctx.comm_size) == (rank, world)
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(30_001, 12_000, world=200.0)
    bounds, rings = pp.synth.circle_world(400, world=200.0)
    if rank == 0:  # only the root holds the tree and the world before the broadcasts
        ctx.tree_upload_bcast(0, 10_000, nx[:10_000], ny[:10_000], nyaw[:10_000])
        ctx.tree_append_bcast(0, 2000, nx[10_000:], ny[10_000:], nyaw[10_000:])
        ctx.obstacles_upload_bcast(0, bounds, rings)
    else:
        ctx.tree_upload_bcast(0, 10_000)
        ctx.tree_append_bcast(0, 2000)
        ctx.obstacles_upload_bcast(0)
    lo, hi = pp.slice_bounds(qx.size, world, rank)
    idx, yaw, ok = ctx.rrt_extend(qx[lo:hi], qy[lo:hi])
    np.savez(os.path.join(outdir, f"r{rank}.npz"), idx=idx, yaw=yaw, ok=ok)
    ctx.close()


def test_one_process_per_gpu_replicates_by_broadcast(pp, tmp_path):
    """the torchrun shape: each rank owns one device and joins the communicator with the id rank 0 made"""
    n = pp.device_count()
    if n < 2:
        pytest.skip("needs >= 2 B200s (gpurun --gpus 2)")
    import multiprocessing as mp
    world = 2
    comm_id = pp.comm_unique_id()
    assert len(comm_id) == 128
    mpctx = mp.get_context("spawn")
    procs = [mpctx.Process(target=_rank_main, args=(r, world, comm_id, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(30_001, 12_000, world=200.0)
    bounds, rings = pp.synth.circle_world(400, world=200.0)
    ctx = pp.Context(0)
    ctx.tree_upload(nx, ny, nyaw)
    ctx.obstacles_upload(bounds, rings)
    idx, yaw, ok = ctx.rrt_extend(qx, qy)
    ctx.close()
    assert np.array_equal(np.concatenate([p["idx"] for p in parts]), idx)
    assert np.array_equal(np.concatenate([p["ok"] for p in parts]), ok)
    assert np.array_equal(np.concatenate([p["yaw"] for p in parts]), yaw)


def test_incremental_node_grid(pp, ctx, O):
    """src/rrt.rs:586-589 inserts one node per iteration: the node grid must not be rebuilt per append.  A tree grown by
    1-node and 512-node appends up to 2^20 nodes: NN bit-exact against the oracle along the way, and the number of
    O(n) rebuilds stays at one per ~4 096 appended nodes."""
    rng = np.random.default_rng(5)
    total = 1 << 20
    nx, ny = rng.uniform(0, 1000, total), rng.uniform(0, 1000, total)
    qx, qy = rng.uniform(-5, 1005, 96), rng.uniform(-5, 1005, 96)
    n = 6000
    ctx.tree_upload(nx[:n], ny[:n])
    b0 = ctx.nn_grid_builds
    for k in range(3000):  # the scalar plan_one loop: one append, one query
        ctx.tree_append(nx[n:n + 1], ny[n:n + 1])
        n += 1
        got = ctx.nn(qx[k % 96:k % 96 + 1], qy[k % 96:k % 96 + 1], want_d2=False)
        if k % 250 == 0 or k == 2999:
            assert got[0] == O.nn_brute(nx[:n], ny[:n], qx[k % 96:k % 96 + 1], qy[k % 96:k % 96 + 1])[0][0], k
    assert ctx.nn_grid_builds - b0 <= 2, ctx.nn_grid_builds - b0  # 3 000 appends: one initial build + at most one more
    # jump to half a million nodes, then rounds of 512 up to 2^20
    ctx.tree_append(nx[n:1 << 19], ny[n:1 << 19])
    n = 1 << 19
    b1 = ctx.nn_grid_builds
    rounds = 0
    while n < total:
        ctx.tree_append(nx[n:n + 512], ny[n:n + 512])
        n += 512
        rounds += 1
        got = ctx.nn(qx, qy, want_d2=False)
        if rounds % 128 == 1 or n == total:
            assert np.array_equal(got, O.nn_brute(nx[:n], ny[:n], qx, qy)[0]), n
    assert ctx.tree_size == total and rounds == 1024
    builds = ctx.nn_grid_builds - b1
    assert builds <= rounds * 512 // 4096 + 2, builds  # one rebuild per 4 096 appended nodes, not one per append
    # many queries against a freshly appended tail take the rebuild when the tail would cost more
    ctx.tree_append(nx[:2000] + 0.25, ny[:2000])
    m = 1 << 18
    bq = rng.uniform(0, 1000, m), rng.uniform(0, 1000, m)
    got = ctx.nn(bq[0], bq[1], want_d2=False)
    sub = np.arange(0, m, 4001)
    full_x, full_y = np.concatenate([nx, nx[:2000] + 0.25]), np.concatenate([ny, ny[:2000]])
    assert np.array_equal(got[sub], O.nn_brute(full_x, full_y, bq[0][sub], bq[1][sub])[0])

"""CPU: host-side logic -- synthetic input generator, the module mirror's pure-host pieces, and the
N > 1 sharding path with the gloo backend (world_size 2)."""
import math
import os
import socket

import numpy as np
import pytest


def test_synth_matches_oracle_generator(pp, O):
    for seed, stream in [(0xD0B10003, 0), (0xD0B10004, 1), (12345, 7)]:
        a = pp.synth.uniform(seed, stream, 4097, -3.5, 9.25)
        assert np.array_equal(a, O.uniform(seed, stream, 4097, -3.5, 9.25))
        assert a.min() >= -3.5 and a.max() < 9.25
    full = pp.synth.uniform(99, 3, 1000)
    assert np.array_equal(full[300:700], pp.synth.uniform(99, 3, 400, first=300))  # element i is position-independent
    assert O.uniform(1, 0, 3)[0] == O.lib().ppo_uniform(1, 0, 0)


def test_synth_workloads(pp):
    sx, sy, syaw, ex, ey, eyaw = pp.synth.dubins_pairs(1000)
    assert np.abs(sx).max() < 2 and np.abs(syaw).max() <= math.pi
    bounds, rings = pp.synth.circle_world(50, world=100.0)
    assert len(rings) == 50 and all(8 <= len(r[0]) <= 21 for r in rings)
    _, far = pp.synth.circle_world(50, world=100.0, shift=5000.0)
    assert min(r[0].min() for r in far) > 4000
    e = pp.synth.dubins_edges(100)
    assert np.allclose(e[2], np.arctan2(e[4] - e[1], e[3] - e[0]))  # child yaw aims at the parent (src/rrt.rs:267-271)


def test_module_mirror_host_pieces(pp):
    r, d = pp.rrt, pp.dubins
    assert d.WORD_MODES[2] == (d.L, d.S, d.R) and d.Mode.S.value == 1
    assert d.DubinsConfig(1, 1, 0.5, -3, -3, -0.5, turn_radius=1.0, step_size=0.1).turn_radius == 1.0
    root = r.Node.new_root((0.0, 0.0), 0.3)
    a = r.Node((3.0, 4.0), root)
    b = r.Node((5.0, 5.0), a)
    assert a.get_yaw() == math.atan2(-4.0, -3.0)  # heading toward the parent
    assert [n.point for n in r.NodeIter(b)] == [(5.0, 5.0), (3.0, 4.0), (0.0, 0.0)]
    assert [n.point for n in b.get_above()] == [(3.0, 4.0), (0.0, 0.0)]
    g = r.Node.new_goal((9.0, 9.0), b, 1.0)
    assert g.get_yaw() == 1.0 and g.get_parent() is b
    rx, ry = r.create_circle((5.0, 5.0), 1.0)
    assert len(rx) == 8 + 1 - 1 or len(rx) in (8, 9)
    assert r.euclidean_length((np.array([0.0, 3.0, 3.0]), np.array([0.0, 4.0, 5.0]))) == 6.0
    assert r.Robot(1.0, 2.0, 0.8).get_steer() == 0.8


def test_slice_bounds(pp):
    """pp_slice_bounds (the C-ABI's slicing rule, used by pp_group_* and by bench.py's ranks): contiguous, covering,
    [g*n/G, (g+1)*n/G)"""
    for n, world in [(16, 1), (17, 2), (1 << 24, 8), (5, 8), (0, 4), ((1 << 40) + 3, 7)]:
        cuts = [pp.slice_bounds(n, world, r) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
        assert cuts == [((r * n) // world, ((r + 1) * n) // world) for r in range(world)]
    assert pp.slice_bounds(10, 2, 5) == (5, 10)  # out-of-range part is clamped, never out of bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    import sys
    import torch
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as graft
    pp = graft.import_package()
    O = graft.import_oracle()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n_nodes, m_total = 3000, 4001
    # replicated tree: only rank 0 holds the data before the broadcast
    if rank == 0:
        _, _, nx, ny, nyaw = pp.synth.extend_inputs(1, n_nodes, world=100.0)
        tree = torch.from_numpy(np.stack([nx, ny, nyaw]))
    else:
        tree = torch.zeros((3, n_nodes), dtype=torch.float64)
    dist.broadcast(tree, 0)  # on the GPU box: pp_tree_upload_bcast (ncclBroadcast inside the library)
    # the launcher's channel also carries the 128-byte NCCL id from rank 0 to the others (pp_ctx_comm_init)
    ident = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ident, 0)
    assert ident[0] == bytes(range(128))
    lo, hi = pp.slice_bounds(m_total, world, rank)
    qx = pp.synth.uniform(pp.synth.SEED_C4_Q, 0, hi - lo, 0.0, 100.0, first=lo)
    qy = pp.synth.uniform(pp.synth.SEED_C4_Q, 1, hi - lo, 0.0, 100.0, first=lo)
    # the GPU kernels cannot run here; the oracle stands in for the per-rank compute so that the host-side
    # slicing / replication / gathering logic is what is under test
    idx, _ = O.nn_brute(tree[0].numpy(), tree[1].numpy(), qx, qy)
    tt = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)  # multi-GPU timings are the max over ranks
    t = float(tt.item())
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(idx, bucket, dst=0)
    allidx = np.concatenate(bucket) if rank == 0 else None
    if rank == 0:
        np.save(os.path.join(outdir, "idx.npy"), allidx)
        np.save(os.path.join(outdir, "t.npy"), np.array([t]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(pp, O, tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "idx.npy")
    assert np.load(tmp_path / "t.npy")[0] == 11.0  # max over ranks
    _, _, nx, ny, _ = pp.synth.extend_inputs(1, 3000, world=100.0)
    qx = pp.synth.uniform(pp.synth.SEED_C4_Q, 0, 4001, 0.0, 100.0)
    qy = pp.synth.uniform(pp.synth.SEED_C4_Q, 1, 4001, 0.0, 100.0)
    want, _ = O.nn_brute(nx, ny, qx, qy)
    assert np.array_equal(got, want)  # sharded run is byte-identical to the single-process run


class _FakeCtx:
    """stands in for the GPU context: verdicts are a deterministic function of each edge's coordinates, so the batched
    and the node-by-node code paths must see the same answers (pure host-logic test, no device)"""

    def __init__(self):
        self.calls = 0

    def obstacles_upload(self, *a):
        pass

    def tree_upload(self, *a):
        pass

    def tree_append(self, *a):
        pass

    def collide_dubins(self, sx, sy, syaw, ex, ey, eyaw, radius, step):
        self.calls += 1
        v = np.asarray(sx) * 12.9898 + np.asarray(sy) * 78.233 + np.asarray(ex) * 37.719 + np.asarray(ey) * 4.581
        return ((np.sin(v) * 43758.5453) % 1.0 < 0.55).astype(np.uint8)


def test_batched_optimize_matches_recursive_optimize_on_the_host(pp):
    """RRT._optimize_many (all nodes of a round, level by level) picks what RRT.optimize picks node by node"""
    r = pp.rrt
    ctx = _FakeCtx()
    bounds = (np.array([0.0, 0.0, 50.0, 50.0, 0.0]), np.array([0.0, 50.0, 50.0, 0.0, 0.0]))
    space = r.Space.from_inflated(bounds, r.Robot(1.0, 1.0, 0.8), [], ctx=ctx, seed=3)
    planner = r.RRT((1.0, 1.0), 0.0, (49.0, 49.0), 0.0, 100, 0.1, space)
    rng = np.random.default_rng(5)
    nodes = [planner.nodes[0]]
    for _ in range(120):  # a random tree: every new node hangs off a random earlier one
        nodes.append(r.Node((float(rng.uniform(0, 50)), float(rng.uniform(0, 50))), nodes[int(rng.integers(0, len(nodes)))]))
    picks = nodes[1:]
    many = planner._optimize_many(picks)
    batched_calls = ctx.calls
    assert batched_calls <= r.RECURSION_LIMIT
    found = 0
    for n, a in zip(picks, many):
        b = planner.optimize(n, 0)
        assert (a is None) == (b is None)
        if a is not None:
            found += 1
            assert [v.point for v in r.NodeIter(a)] == [v.point for v in r.NodeIter(b)]
    assert found > 10 and ctx.calls > batched_calls * 5
    assert planner._optimize_many([]) == []

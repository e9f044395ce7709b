"""CPU: the Python mirror's planner logic (plan_one, plan_rounds, check_finish_many, optimize) run end to end with the
GPU context replaced by a stand-in that answers every batch call from the CPU oracle.  This is test scaffolding
only -- the product has no such back end (tests/test_abi.py::test_no_cpu_fallback) -- and it lets the host-side
bookkeeping (tree slots, span arithmetic, prefix-ANDs, chunk order, min_by) be checked without a device."""
import math

import numpy as np
import pytest


class OracleCtx:
    """the subset of rs-pathplanning_b200._ffi.Context that rrt.py uses, answered by oracle/pp_oracle.c"""

    def __init__(self, O, pp):
        self.O, self.ffi = O, pp._ffi
        self.world = None
        self.x, self.y, self.yaw, self.parent = [], [], [], []
        self.calls = {}
        self._paths = None

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def obstacles_upload(self, bounds_xy, rings_xy):
        self.world = self.O.OracleWorld(bounds_xy, list(rings_xy))

    def tree_upload(self, x, y, yaw=None, parent=None):
        self.x, self.y = list(np.atleast_1d(x)), list(np.atleast_1d(y))
        self.yaw, self.parent = list(np.atleast_1d(yaw)), list(np.atleast_1d(parent))

    def tree_append(self, x, y, yaw=None, parent=None):
        self._count("tree_append")
        self.x += list(np.atleast_1d(x)); self.y += list(np.atleast_1d(y))
        self.yaw += list(np.atleast_1d(yaw)); self.parent += list(np.atleast_1d(parent))

    def nn(self, qx, qy, flags=0, want_d2=True):
        self._count("nn")
        idx, d2 = self.O.nn_brute(self.x, self.y, np.atleast_1d(qx), np.atleast_1d(qy))
        return (idx, d2) if want_d2 else idx

    def collide_dubins(self, sx, sy, syaw, ex, ey, eyaw, radius, step, flags=0):
        self._count("collide_dubins")
        return self.world.verify_dubins_edges(sx, sy, syaw, ex, ey, eyaw, radius, step)

    def verify_polylines(self, lines, flags=0):
        self._count("verify_polylines")
        return np.array([self.world.verify(lx, ly) for lx, ly in lines], np.uint8)

    def dubins_sample_count(self, sx, sy, syaw, ex, ey, eyaw, radius, step, from_origin=False):
        self._count("dubins_sample_count")
        n = len(sx)
        plan = np.zeros(n * self.ffi.PLAN_BYTES, np.uint8)
        counts = np.zeros(n, np.uint32)
        self._paths = []
        for i in range(n):
            p = self.O.dubins_path(sx[i], sy[i], syaw[i], ex[i], ey[i], eyaw[i], radius, step, from_origin)
            self._paths.append(p)
            plan[i * self.ffi.PLAN_BYTES + 104] = self.ffi.WORD_NONE if p is None else p.word
            counts[i] = 0 if p is None else p.x.size
        return counts, plan

    def dubins_sample_fill(self, plan, counts):
        offsets = np.zeros(len(counts), np.uint64)
        if len(counts):
            np.cumsum(counts[:-1].astype(np.uint64), out=offsets[1:])
        rows = [np.stack([p.x, p.y, p.yaw], 1) for p in self._paths if p is not None]
        out = np.concatenate(rows) if rows else np.zeros((0, 3))
        return out, offsets

    def rrt_extend_dubins(self, qx, qy, radius, step, nn_flags=0, collide_flags=0):
        self._count("rrt_extend_dubins")
        qx, qy = np.atleast_1d(qx), np.atleast_1d(qy)
        idx, _ = self.O.nn_brute(self.x, self.y, qx, qy)
        nx, ny, nyaw = (np.asarray(v)[idx] for v in (self.x, self.y, self.yaw))
        yaw = np.array([self.O.compute_yaw(qx[i], qy[i], nx[i], ny[i]) for i in range(qx.size)])  # Node::new
        ok = self.world.verify_dubins_edges(qx, qy, yaw, nx, ny, nyaw, radius, step)
        return idx, yaw, ok


def _planner(O, pp, seed, max_iter=240, obstacles=True):
    r = pp.rrt
    ctx = OracleCtx(O, pp)
    bounds = (np.array([0.0, 0.0, 40.0, 40.0]), np.array([0.0, 40.0, 40.0, 0.0]))
    rings = [r.create_circle((20.0, 20.0), 5.0), r.create_circle((10.0, 28.0), 3.0),
             r.create_circle((30.0, 12.0), 3.0)] if obstacles else []
    space = r.Space.from_inflated(bounds, r.Robot(1.0, 1.0, 2.0), rings, ctx=ctx, seed=seed)
    return r.RRT((3.0, 3.0), 0.0, (36.0, 36.0), 0.0, max_iter, 0.25, space), ctx


def _chain_ok(O, ctx, planner, slot):
    lx, ly = O.line_to_origin(ctx.x, ctx.y, ctx.yaw, ctx.parent, slot, planner.space.get_steer(), planner.step_size)
    return ctx.world.verify(lx, ly)


def test_plan_rounds_keeps_the_tree_invariant_and_returns_a_verified_path(O, pp):
    planner, ctx = _planner(O, pp, seed=11)
    path = planner.plan_rounds(batch=40)
    n = len(planner.nodes)
    assert n > 40 and n == len(ctx.x) == len(ctx.parent)
    # host objects and the flat mirror describe the same tree
    for slot, node in enumerate(planner.nodes):
        assert (ctx.x[slot], ctx.y[slot], ctx.yaw[slot]) == (node.point[0], node.point[1], node.yaw)
        assert ctx.parent[slot] == (-1 if node.parent is None else planner._slot[id(node.parent)])
        assert ctx.parent[slot] < slot
    # src/rrt.rs:583-589: only nodes whose whole chain verifies are inserted
    for slot in range(1, n, 7):
        assert _chain_ok(O, ctx, planner, slot)
    # one extend call, at most one append and one goal-visibility launch per round
    rounds = math.ceil(240 / 40)
    assert ctx.calls["rrt_extend_dubins"] == rounds and ctx.calls["tree_append"] <= rounds
    assert path is not None, "this world is easy: 240 samples must reach the goal"
    px, py = path
    assert ctx.world.verify(px, py)
    assert math.hypot(px[-1] - 36.0, py[-1] - 36.0) < 1e-9  # finalize reverses: ... -> goal
    assert math.hypot(px[0] - 3.0, py[0] - 3.0) < 0.3  # first sample of the edge that ends at the root


def test_check_finish_many_returns_what_check_finish_returns(O, pp):
    planner, ctx = _planner(O, pp, seed=5, max_iter=120)
    planner.plan_rounds(batch=30)
    picks = planner.nodes[1::3]
    many = planner.check_finish_many(picks)
    found = 0
    for node, a in zip(picks, many):
        b = planner.check_finish(node)
        assert (a is None) == (b is None)
        if a is not None:
            found += 1
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert found >= 1
    assert planner.check_finish_many([]) == []


def test_plan_rounds_picks_the_shortest_candidate(O, pp):
    """min_by euclidean_length (src/rrt.rs:611-617): the returned line is no longer than any node's goal line"""
    planner, ctx = _planner(O, pp, seed=2, max_iter=90, obstacles=False)
    best = planner.plan_rounds(batch=30)
    assert best is not None
    r = pp.rrt
    lines = [ln for ln in planner.check_finish_many(planner.nodes[1:]) if ln is not None]
    assert lines and r.euclidean_length(best) <= min(r.euclidean_length(ln) for ln in lines) + 1e-12


def test_plan_one_inserts_exactly_the_verified_samples(O, pp):
    planner, ctx = _planner(O, pp, seed=8, max_iter=60)
    inserted = 0
    for _ in range(60):
        before = len(planner.nodes)
        planner.plan_one()
        inserted += len(planner.nodes) - before
        assert len(planner.nodes) - before in (0, 1)
    assert inserted == len(planner.nodes) - 1 == ctx.calls.get("tree_append", 0)
    assert 0 < inserted < 60  # some samples land in obstacles or cannot be connected
    for slot in range(1, len(planner.nodes)):
        assert _chain_ok(O, ctx, planner, slot)


def test_finalize_panics_where_the_reference_does_and_line_to_origin_falls_back(O, pp):
    """an edge without a feasible word (non-finite pose, Q5): line_to_origin yields the edge's start point
    (src/rrt.rs:313), finalize's copy of the loop panics "Should plan dubins curve" (src/rrt.rs:529)"""
    planner, ctx = _planner(O, pp, seed=3, max_iter=10, obstacles=False)
    r = pp.rrt
    root = planner.nodes[0]
    # nodes outside the bounds: no shortcut verifies, so optimize() returns None and finalize walks this very chain
    mid = r.Node((55.0, 54.0), root)
    bad = r.Node.new_goal((58.0, 58.0), mid, float("nan"))  # a pose with a NaN heading has no Dubins word
    leaf = r.Node((59.0, 59.5), bad)
    lx, ly = r.line_to_origin(leaf, planner.space.get_steer(), planner.step_size, ctx)
    assert (58.0, 58.0) in set(zip(lx.tolist(), ly.tolist()))  # the fallback point of the edge bad -> mid
    assert (lx[-1], ly[-1]) == root.point
    goal = r.Node.new_goal(planner.goal, leaf, planner.goal_yaw)
    with pytest.raises(RuntimeError, match="Should plan dubins curve"):
        planner.finalize(goal)
    with pytest.raises(RuntimeError, match="Should plan dubins curve"):
        planner.check_finish_many([leaf])

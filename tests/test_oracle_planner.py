"""CPU: the oracle's restatement of RRT::optimize / optimize_from_goal / finalize / check_finish
(/root/reference/src/rrt.rs:428-540, RECURSION_LIMIT :14) and its verdict margins.

Three independent routes must agree exactly on the same tree:
  (1) oracle/pp_oracle.c  ppo_optimize / ppo_check_finish (flat tree + node arena),
  (2) a literal object-graph transliteration of the Rust source in this file (Node objects, the reference's
      candidate-by-candidate loop, whole-chain line_to_origin + verify per candidate),
  (3) the product's Python mirror (rrt.py: batched candidates, per-edge decomposition, suffix-ANDs) run on the
      oracle-backed stand-in context of tests/test_planner_on_oracle.py.
"""
import math

import numpy as np
import pytest

from test_planner_on_oracle import OracleCtx

RECURSION_LIMIT = 16


class RNode:  # src/rrt.rs:161-214
    def __init__(self, x, y, yaw, parent):
        self.x, self.y, self.yaw, self.parent = x, y, yaw, parent

    @staticmethod
    def new(O, x, y, parent):  # :169-175
        return RNode(x, y, O.compute_yaw(x, y, parent.x, parent.y), parent)

    def chain(self):  # NodeIter :248-265
        out, c = [], self
        while c is not None:
            out.append(c)
            c = c.parent
        return out


def r_line_to_origin(O, node, radius, step):  # :291-321, node -> root chunk order
    xs, ys = [], []
    for n in node.chain():
        if n.parent is None:
            xs.append([n.x]); ys.append([n.y])
        else:
            p = O.dubins_path(n.x, n.y, n.yaw, n.parent.x, n.parent.y, n.parent.yaw, radius, step)
            if p is None:
                xs.append([n.x]); ys.append([n.y])
            else:
                xs.append(p.x); ys.append(p.y)
    return np.concatenate(xs), np.concatenate(ys)


def r_optimize(O, W, node, i, radius, step):  # :463-487
    if i >= RECURSION_LIMIT:
        return None
    for to_node in reversed(node.chain()):
        new_node = RNode.new(O, node.x, node.y, to_node)
        if W.verify(*r_line_to_origin(O, new_node, radius, step)):
            deeper = r_optimize(O, W, to_node, i + 1, radius, step)
            return RNode.new(O, node.x, node.y, deeper) if deeper is not None else new_node
    return None


def r_check_finish(O, W, node, goal, goal_yaw, radius, step):  # :428-438, :489-540
    goal_node = RNode(goal[0], goal[1], goal_yaw, node)
    opt = r_optimize(O, W, node, 0, radius, step)
    top = RNode(goal[0], goal[1], goal_yaw, opt) if opt is not None else goal_node
    xs, ys = [], []
    for n in top.chain():
        if n.parent is not None:
            p = O.dubins_path(n.x, n.y, n.yaw, n.parent.x, n.parent.y, n.parent.yaw, radius, step)
            assert p is not None
            xs.append(p.x); ys.append(p.y)
    lx, ly = np.concatenate(xs)[::-1], np.concatenate(ys)[::-1]
    return (lx, ly), bool(W.verify(lx, ly)), top


def _poses(chain_nodes):
    return np.array([(n.x, n.y, n.yaw) for n in chain_nodes])


def _grown_planner(O, pp, seed, iters, obstacles=True, start=(3.0, 3.0)):
    r = pp.rrt
    ctx = OracleCtx(O, pp)
    bounds = (np.array([0.0, 0.0, 40.0, 40.0]), np.array([0.0, 40.0, 40.0, 0.0]))
    rings = [r.create_circle((20.0, 20.0), 5.0), r.create_circle((10.0, 28.0), 3.0),
             r.create_circle((30.0, 12.0), 3.0), r.create_circle((27.0, 30.0), 2.5)] if obstacles else []
    space = r.Space(bounds, r.Robot(0.0, 1.0, 2.0), rings, ctx=ctx, seed=seed)
    planner = r.RRT(start, 0.3, (36.0, 36.0), 0.5, iters, 0.25, space)
    planner.plan_rounds(batch=16, max_iter=iters)
    return planner, ctx


def _rtree(ctx):
    nodes = []
    for x, y, yaw, par in zip(ctx.x, ctx.y, ctx.yaw, ctx.parent):
        nodes.append(RNode(float(x), float(y), float(yaw), nodes[int(par)] if par >= 0 else None))
    return nodes


@pytest.mark.parametrize("seed", [3, 17])
def test_optimize_three_routes_agree(O, pp, seed):
    planner, ctx = _grown_planner(O, pp, seed, 160)
    W = ctx.world
    rnodes = _rtree(ctx)
    depth = [len(n.chain()) for n in rnodes]
    picks = sorted(range(1, len(rnodes)), key=lambda i: -depth[i])[:10] + list(range(1, len(rnodes), 9))
    assert max(depth) >= 4
    compared = 0
    for i in picks:
        chain, flags, verifies = O.optimize(W, ctx.x, ctx.y, ctx.yaw, ctx.parent, i, 2.0, 0.25)
        lit = r_optimize(O, W, rnodes[i], 0, 2.0, 0.25)
        assert (chain is None) == (lit is None)
        mir = planner.optimize(planner.nodes[i], 0)
        assert (mir is None) == (lit is None)
        if lit is None:
            continue
        assert np.array_equal(chain, _poses(lit.chain())), i  # same arithmetic, same decisions: bit-equal
        mp = np.array([(n.point[0], n.point[1], n.yaw) for n in pp.rrt.NodeIter(mir)])
        assert np.array_equal(mp, chain), i
        assert verifies >= 1
        compared += 1
    assert compared >= 10
    # the batched form picks the same chains
    many = planner._optimize_many([planner.nodes[i] for i in picks])
    for i, m in zip(picks, many):
        chain, _, _ = O.optimize(W, ctx.x, ctx.y, ctx.yaw, ctx.parent, i, 2.0, 0.25)
        assert (m is None) == (chain is None)
        if m is not None:
            assert np.array_equal(np.array([(n.point[0], n.point[1], n.yaw) for n in pp.rrt.NodeIter(m)]), chain)


def test_optimize_root_nesting_quirk(O, pp):
    """when the shortcut reaches the root the reference recurses on the root itself: Node::new(root, root) has
    yaw atan2(0, 0) = 0 and nests until RECURSION_LIMIT; the result chain then carries 16 copies of the root point"""
    planner, ctx = _grown_planner(O, pp, 5, 40, obstacles=False, start=(20.0, 20.0))  # room for the turn on the spot
    W = ctx.world
    i = len(ctx.x) - 1
    chain, flags, verifies = O.optimize(W, ctx.x, ctx.y, ctx.yaw, ctx.parent, i, 2.0, 0.25)
    lit = r_optimize(O, W, _rtree(ctx)[i], 0, 2.0, 0.25)
    assert chain is not None and np.array_equal(chain, _poses(lit.chain()))
    root = (ctx.x[0], ctx.y[0])
    copies = int(((chain[:, 0] == root[0]) & (chain[:, 1] == root[1])).sum())
    assert copies == RECURSION_LIMIT  # new nodes at recursion levels 1..15 + the root itself
    assert len(chain) == RECURSION_LIMIT + 1
    assert np.all(chain[1:-1, 2] == 0.0)  # atan2(0, 0)


@pytest.mark.parametrize("seed", [3, 17])
def test_check_finish_three_routes_agree(O, pp, seed):
    planner, ctx = _grown_planner(O, pp, seed, 160)
    W = ctx.world
    rnodes = _rtree(ctx)
    goal, gyaw = planner.goal, planner.goal_yaw
    picks = list(range(1, len(rnodes), 5))
    mirror_many = planner.check_finish_many([planner.nodes[i] for i in picks])
    found = 0
    for i, mm in zip(picks, mirror_many):
        fin = O.check_finish(W, ctx.x, ctx.y, ctx.yaw, ctx.parent, i, goal, gyaw, 2.0, 0.25)
        (lx, ly), ok, top = r_check_finish(O, W, rnodes[i], goal, gyaw, 2.0, 0.25)
        assert fin.ok == ok
        assert np.array_equal(fin.line[0], lx) and np.array_equal(fin.line[1], ly)
        assert np.array_equal(fin.chain, _poses(top.chain()))
        m1 = planner.check_finish(planner.nodes[i])
        assert (m1 is None) == (not ok) and (mm is None) == (not ok)
        if ok:
            found += 1
            for m in (m1, mm):
                assert np.array_equal(m[0], lx) and np.array_equal(m[1], ly)
            assert lx[-1] == goal[0] and ly[-1] == goal[1]  # finalize reverses: the goal's own sample comes last
    assert found >= 2


def test_verify_margin_is_a_safe_radius(O, pp):
    """moving every vertex by less than the margin never changes the verdict (free: clearance; blocked: depth)"""
    bounds, rings = pp.synth.circle_world(120, world=60.0, rmin=1.0, rmax=3.0)
    W = O.OracleWorld(bounds, rings)
    rng = np.random.default_rng(4)
    seen = {True: 0, False: 0}
    for k in range(300):
        n = int(rng.choice([1, 2, 3, 12, 40]))
        x0, y0 = rng.uniform(-1, 61, 2)
        lx, ly = x0 + np.cumsum(rng.normal(0, 0.5, n)), y0 + np.cumsum(rng.normal(0, 0.5, n))
        v, m = W.verify_margin(lx, ly)
        assert v == W.verify(lx, ly) and m >= 0.0
        seen[v] += 1
        if not (m > 1e-6) or not math.isfinite(m):
            continue
        for _ in range(6):
            th = rng.uniform(0, 2 * math.pi, n)
            rad = rng.uniform(0, 0.99 * m, n)
            assert W.verify(lx + rad * np.cos(th), ly + rad * np.sin(th)) == v, (k, v, m)
    assert seen[True] > 20 and seen[False] > 20
    # a grazing line: margin ~ 0 on both sides of the contact
    rx, ry = pp.rrt.create_circle((30.0, 30.0), 2.0)
    W = O.OracleWorld(bounds, [(rx, ry)])
    k = int(np.argmax(ry))
    cx, top = rx[k], ry[k]
    for dy in (1e-10, -1e-10):  # just above / just below the ring's top vertex
        v, m = W.verify_margin(np.array([cx - 0.2, cx + 0.2]), np.array([top + dy, top + dy]))
        assert m < 1e-6


def test_count_and_feasibility_flags(O):
    """PPO_FLAG_NEAR_COUNT fires exactly on knife-edge sample counts; the word margin changes sign at infeasibility"""
    # straight LSL path of length exactly 5 steps of 1.0: pd reaches |l| exactly -> knife edge
    n, fl = O.dubins_path_flags(0.0, 0.0, 0.0, 5.0, 0.0, 0.0, 1.0, 1.0)
    assert fl & O.FLAG_NEAR_COUNT
    n2, fl2 = O.dubins_path_flags(0.0, 0.0, 0.0, 5.3, 0.0, 0.0, 1.0, 1.0)
    assert not (fl2 & O.FLAG_NEAR_COUNT) and n2 >= 5
    rng = np.random.default_rng(2)
    flagged = 0
    for _ in range(400):
        p = rng.uniform(-5, 5, 4)
        a, b = rng.uniform(-math.pi, math.pi, 2)
        flagged += bool(O.dubins_path_flags(p[0], p[1], a, p[2], p[3], b, 1.0, 0.1)[1] & O.FLAG_NEAR_COUNT)
    assert flagged <= 1  # random paths are not knife edges
    assert O.dubins_path_flags(0.0, 0.0, 2.0e6, 3.0, 1.0, 0.0, 1.0, 0.1)[1] & O.FLAG_HUGE_ANGLE
    for w in range(6):
        for _ in range(200):
            al, be, d = rng.uniform(0, 2 * math.pi), rng.uniform(0, 2 * math.pi), rng.uniform(0, 6)
            m = O.dubins_word_margin(w, al, be, d)
            feas = O.dubins_word(w, al, be, d) is not None
            assert feas == (m >= 0.0) or abs(m) < 1e-12

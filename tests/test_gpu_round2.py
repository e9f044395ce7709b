"""GPU: round-2 parity additions -- full-size exhaustive cross-check of the culls, the adversarial near-parallel set,
extend steps with unanswerable queries, and planners living side by side (one context per Space)."""
import math

import numpy as np
import pytest

from conftest import check_dubins_verdicts, exactly_free, near_parallel_edges

pytestmark = pytest.mark.gpu

NN_DEFAULT, NN_PLAIN, NN_GRID, NN_UNSORTED, NN_SCAN = 0, 1, 2, 4, 8
DEFAULT, NO_CULL, USE_GRID, UNSORTED, SCAN, FUSED = 0, 1, 2, 4, 8, 16


def test_c4_full_size_default_equals_exhaustive_loop(ctx, pp, O):
    """Space::verify (src/rrt.rs:124-137) on ALL 2^20 C4 edges x 10 k rings: the default (obstacle-grid cull) against
    PP_COLLIDE_NO_CULL, geo's exhaustive segment-pair loop with no rejection at all -- byte-equal flags.  A difference
    would have to be rounding noise of a near-parallel pair (free in exact rational geometry)."""
    m = n_nodes = 1 << 20
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(m, n_nodes)
    bounds, rings = pp.synth.circle_world(10_000)
    ctx.tree_upload(nx, ny, nyaw)
    ctx.obstacles_upload(bounds, rings)
    idx, _, ok = ctx.rrt_extend(qx, qy)
    ex, ey = nx[idx], ny[idx]
    ok_all = ctx.collide_segments(qx, qy, ex, ey, flags=NO_CULL)
    bad = np.nonzero(ok != ok_all)[0]
    if bad.size:
        W = O.OracleWorld(bounds, rings)
        assert bad.size < 8
        for i in bad:
            assert ok_all[i] == 0 and ok[i] == 1 and exactly_free(W, qx[i], qy[i], ex[i], ey[i]), i
    assert 0.8 < ok.mean() < 0.9
    # the no-hit obstacle set (rings moved outside the world): everything inside the bounds is free on both paths
    far = [(rx + 5000.0, ry) for rx, ry in rings[:2000]]
    ctx.obstacles_upload(bounds, far)
    sub = slice(0, 1 << 17)
    a = ctx.collide_segments(qx[sub], qy[sub], ex[sub], ey[sub])
    b = ctx.collide_segments(qx[sub], qy[sub], ex[sub], ey[sub], flags=NO_CULL)
    assert np.array_equal(a, b) and a.mean() > 0.99


def test_c5_edges_default_equals_exhaustive_loop(ctx, pp, O):
    """2^16 C5 Dubins edges (step 0.05, ~785 samples each) against the 10 k-ring obstacle set: default culls vs the
    exhaustive loop on the SAME device-generated samples -- byte-equal flags (5e7 line segments x 1.4e5 ring segments)."""
    e = 1 << 16
    bounds, rings = pp.synth.circle_world(10_000)
    ctx.obstacles_upload(bounds, rings)
    sx, sy, syaw, ex, ey, eyaw = pp.synth.dubins_edges(e)
    ok = ctx.collide_dubins(sx, sy, syaw, ex, ey, eyaw, 1.0, 0.05)
    ok_all = ctx.collide_dubins(sx, sy, syaw, ex, ey, eyaw, 1.0, 0.05, flags=NO_CULL)
    assert np.array_equal(ok, ok_all)
    assert 0.01 < ok.mean() < 0.9
    # and against the oracle (its own samples, 1e-9 apart) on a sub-sample, differences classified by its margins
    W = O.OracleWorld(bounds, rings)
    sub = np.arange(0, e, 331)
    check_dubins_verdicts(O, W, ok[sub], sx[sub], sy[sub], syaw[sub], ex[sub], ey[sub], eyaw[sub], 1.0, 0.05)


@pytest.mark.parametrize("radius,step,n_rings", [(1.0, 0.1, 10_000), (0.4, 0.05, 20_000), (3.0, 0.2, 2_000)])
def test_short_dubins_edges_path_box_equals_exhaustive_loop(ctx, pp, O, radius, step, n_rings):
    """the verify kernel dismisses a whole Dubins edge when nothing is registered under the path's bounding box
    (pp_path_box / pp_path_box_free): exercised where it fires -- extend-step edges (query -> nearest tree node, a few
    turn radii long) -- against PP_COLLIDE_NO_CULL on the same device-generated samples (byte-equal), against the
    oracle on a sub-sample, and on edges hugging the bounds and pointing out of them."""
    m = 1 << 16
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(m, 1 << 18)
    bounds, rings = pp.synth.circle_world(n_rings)
    ctx.tree_upload(nx, ny, nyaw)
    ctx.obstacles_upload(bounds, rings)
    idx = ctx.nn(qx, qy)[0]
    sx, sy = qx, qy                                   # new node -> its parent, yaw aimed at the parent (src/rrt.rs:267-271)
    ex, ey, eyaw = nx[idx], ny[idx], nyaw[idx]
    syaw = np.arctan2(ey - sy, ex - sx)
    ok = ctx.collide_dubins(sx, sy, syaw, ex, ey, eyaw, radius, step)
    ok_all = ctx.collide_dubins(sx, sy, syaw, ex, ey, eyaw, radius, step, flags=NO_CULL)
    assert np.array_equal(ok, ok_all)
    assert 0.2 < ok.mean() < 0.999
    W = O.OracleWorld(bounds, rings)
    sub = np.arange(0, m, 1021)
    check_dubins_verdicts(O, W, ok[sub], sx[sub], sy[sub], syaw[sub], ex[sub], ey[sub], eyaw[sub], radius, step)
    # edges whose loops leave the bounds: start poses within a turn radius of the border, heading outward
    lo_x, hi_x, lo_y, hi_y = bounds[0].min(), bounds[0].max(), bounds[1].min(), bounds[1].max()
    rng = np.random.default_rng(77)
    k = 4096
    t = rng.uniform(0.0, 1.0, k)
    side = rng.integers(0, 4, k)
    off = rng.uniform(0.0, 2.5 * radius, k)
    sx2 = np.where(side == 0, lo_x + off, np.where(side == 1, hi_x - off, lo_x + t * (hi_x - lo_x)))
    sy2 = np.where(side == 2, lo_y + off, np.where(side == 3, hi_y - off, lo_y + t * (hi_y - lo_y)))
    syaw2 = rng.uniform(-math.pi, math.pi, k)
    ex2 = np.clip(sx2 + rng.uniform(-2, 2, k) * radius, lo_x + 1e-3, hi_x - 1e-3)
    ey2 = np.clip(sy2 + rng.uniform(-2, 2, k) * radius, lo_y + 1e-3, hi_y - 1e-3)
    eyaw2 = rng.uniform(-math.pi, math.pi, k)
    a = ctx.collide_dubins(sx2, sy2, syaw2, ex2, ey2, eyaw2, radius, step)
    b = ctx.collide_dubins(sx2, sy2, syaw2, ex2, ey2, eyaw2, radius, step, flags=NO_CULL)
    assert np.array_equal(a, b) and 0.02 < a.mean() < 0.98
    sub = np.arange(0, k, 64)
    check_dubins_verdicts(O, W, a[sub], sx2[sub], sy2[sub], syaw2[sub], ex2[sub], ey2[sub], eyaw2[sub], radius, step)
    # degenerate words: goals straight ahead with the same heading.  A zero-length arc makes the reference's index
    # arithmetic put samples up to five steps off a segment's ends, even BEHIND the start pose (src/dubins.rs:228-237):
    # such words must not take the path-level shortcuts.  Also goals one full turn circle away, identical poses, and
    # long straight-ahead edges (the coarse pass's length)
    sx3 = rng.uniform(lo_x + 20, hi_x - 20, k)
    sy3 = rng.uniform(lo_y + 20, hi_y - 20, k)
    syaw3 = rng.choice([0.0, math.pi / 2, -math.pi / 2, math.pi, 0.3, -2.1], k)
    dist = rng.choice([0.0, 0.5 * step * radius, 3.3 * radius, 2 * math.pi * radius, 7.0, 45.0, 300 * step * radius], k)
    ex3, ey3 = sx3 + dist * np.cos(syaw3), sy3 + dist * np.sin(syaw3)
    eyaw3 = np.where(rng.uniform(size=k) < 0.7, syaw3, syaw3 + rng.choice([math.pi, 1e-9, -1e-9], k))
    a = ctx.collide_dubins(sx3, sy3, syaw3, ex3, ey3, eyaw3, radius, step)
    b = ctx.collide_dubins(sx3, sy3, syaw3, ex3, ey3, eyaw3, radius, step, flags=NO_CULL)
    assert np.array_equal(a, b), np.nonzero(a != b)[0][:10]
    assert 0.05 < a.mean() < 0.999
    sub = np.arange(0, k, 64)
    check_dubins_verdicts(O, W, a[sub], sx3[sub], sy3[sub], syaw3[sub], ex3[sub], ey3[sub], eyaw3[sub], radius, step,
                          max_fragile=None)  # (axis-aligned lattices sit on mod2pi wraps: many are flagged)


def test_near_parallel_extensions(ctx, pp, O):
    """the adversarial class of DESIGN section 3 (ii): PP_COLLIDE_NO_CULL must reproduce geo's exhaustive answer bit
    for bit (rounding noise included), every culled path must reproduce the oracle's culled answer, and the two differ
    only where exact rational geometry says the segment is free."""
    bounds, rings = pp.synth.circle_world(24, world=100.0, rmin=1.0, rmax=3.0)
    ctx.obstacles_upload(bounds, rings)
    W = O.OracleWorld(bounds, rings)
    ax, ay, bx, by = near_parallel_edges(W.rings())
    plain = W.verify_segments(ax, ay, bx, by, culled=False)
    culled = W.verify_segments(ax, ay, bx, by, culled=True)
    assert np.array_equal(ctx.collide_segments(ax, ay, bx, by, flags=NO_CULL), plain)
    for flags in (DEFAULT, USE_GRID, UNSORTED, SCAN):
        assert np.array_equal(ctx.collide_segments(ax, ay, bx, by, flags=flags), culled), flags
    bad = np.nonzero(plain != culled)[0]
    assert bad.size > 0
    for i in bad:
        assert plain[i] == 0 and exactly_free(W, ax[i], ay[i], bx[i], by[i])
    # the same segments as 2-point polylines through the polyline kernel
    lines = [(np.array([ax[i], bx[i]]), np.array([ay[i], by[i]])) for i in range(0, ax.size, 97)]
    sel = np.arange(0, ax.size, 97)
    assert np.array_equal(ctx.verify_polylines(lines, flags=NO_CULL), plain[sel])
    assert np.array_equal(ctx.verify_polylines(lines), culled[sel])


def test_extend_with_unanswerable_queries(ctx, pp, O):
    """pp_nn answers 0xFFFFFFFF for NaN / Inf queries (get_random_node would return None, src/rrt.rs:408-411): the
    extend steps must report ok = 0 and yaw = NaN for them instead of reading node[0xFFFFFFFF], and leave the context
    usable (no sticky CUDA error)"""
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(12_000, 6000, world=100.0)  # > 8 192: PP_COLLIDE_FUSED bins the queries
    bounds, rings = pp.synth.circle_world(100, world=100.0)
    ctx.tree_upload(nx, ny, nyaw)
    ctx.obstacles_upload(bounds, rings)
    nan, inf = float("nan"), float("inf")
    badpos = np.array([0, 1, 2, 31, 32, 1000, 11_999])
    qx[badpos] = [nan, inf, -inf, 5.0, nan, 1e300, 50.0]
    qy[badpos] = [1.0, 2.0, 3.0, nan, nan, 1e300, inf]
    oidx, _ = O.nn_brute(nx, ny, qx, qy)
    none = oidx == 0xFFFFFFFF
    assert none[badpos].all() and none.sum() == badpos.size
    W = O.OracleWorld(bounds, rings)
    good = ~none
    want = np.zeros(qx.size, np.uint8)
    want[good] = W.verify_segments(qx[good], qy[good], nx[oidx[good]], ny[oidx[good]])
    for nnf, cf in [(NN_DEFAULT, DEFAULT), (NN_GRID, USE_GRID), (NN_DEFAULT, FUSED), (NN_SCAN, SCAN), (NN_PLAIN, UNSORTED),
                    (NN_DEFAULT, NO_CULL)]:
        idx, yaw, ok = ctx.rrt_extend(qx, qy, nn_flags=nnf, collide_flags=cf)
        assert np.array_equal(idx, oidx), (nnf, cf)
        assert np.array_equal(ok, want), (nnf, cf)
        assert np.isnan(yaw[none]).all() and np.isfinite(yaw[good]).all()
    idx, yaw, ok = ctx.rrt_extend_dubins(qx, qy, 0.8, 0.1)
    assert np.array_equal(idx, oidx) and not ok[none].any() and np.isnan(yaw[none]).all()
    wyaw = np.arctan2(ny[oidx[good]] - qy[good], nx[oidx[good]] - qx[good])
    check_dubins_verdicts(O, W, ok[good], qx[good], qy[good], wyaw, nx[oidx[good]], ny[oidx[good]], nyaw[oidx[good]],
                          0.8, 0.1)
    few = ctx.rrt_extend(qx[:3000], qy[:3000], collide_flags=FUSED)  # < 8 192 queries: the fused kernel in the caller's order
    assert np.array_equal(few[0], oidx[:3000]) and np.array_equal(few[2], want[:3000])
    one = ctx.rrt_extend([nan], [0.0])  # scalar call
    assert one[0][0] == 0xFFFFFFFF and one[2][0] == 0 and math.isnan(one[1][0])
    # the context is still healthy
    assert np.array_equal(ctx.nn(qx[good][:100], qy[good][:100], want_d2=False), oidx[good][:100])


def test_two_planners_side_by_side(pp, O):
    """every Space owns its context (world + tree), as an RRT owns its Space and RTree in the reference
    (src/rrt.rs:325-356; benches/all.rs builds one per bench): two planners on different worlds, interleaved, must not
    see each other's obstacles or nodes"""
    r = pp.rrt
    b1 = (np.array([0.0, 0.0, 30.0, 30.0]), np.array([0.0, 30.0, 30.0, 0.0]))
    rings1 = [r.create_circle((15.0, 15.0), 6.0)]
    b2 = (np.array([100.0, 100.0, 140.0, 140.0]), np.array([100.0, 140.0, 140.0, 100.0]))
    rings2 = [r.create_circle((120.0, 110.0), 3.0), r.create_circle((112.0, 128.0), 4.0)]
    s1 = r.Space(b1, r.Robot(0.0, 1.0, 0.8), rings1, seed=1)
    s2 = r.Space(b2, r.Robot(0.0, 1.0, 0.8), rings2, seed=2)
    assert s1.ctx is not s2.ctx
    p1 = r.RRT((3.0, 3.0), 0.5, (27.0, 27.0), 0.0, 100, 0.1, s1)
    p2 = r.RRT((104.0, 104.0), 0.2, (135.0, 136.0), 0.0, 100, 0.1, s2)
    for _ in range(60):
        p1.plan_one()
        p2.plan_one()
    p1.plan_rounds(batch=32, max_iter=64)
    p2.plan_rounds(batch=32, max_iter=64)
    for p, s, bounds, rings in ((p1, s1, b1, rings1), (p2, s2, b2, rings2)):
        assert s.ctx.tree_size == len(p.nodes) > 10
        W = O.OracleWorld(bounds, rings)
        nx = np.array([n.point[0] for n in p.nodes]); ny = np.array([n.point[1] for n in p.nodes])
        nyaw = np.array([n.yaw for n in p.nodes])
        par = np.array([p._slot[id(n.parent)] if n.parent is not None else -1 for n in p.nodes], np.int32)
        assert nx.min() >= bounds[0].min() and nx.max() <= bounds[0].max()
        for i in range(1, len(p.nodes), 3):  # every inserted chain verifies in ITS world
            lx, ly = O.line_to_origin(nx, ny, nyaw, par, i, 0.8, 0.1)
            assert W.verify(lx, ly), i
        q = (float(nx.mean()), float(ny.mean()))
        assert p.get_nearest_node(q) is p.nodes[int(O.nn_brute(nx, ny, [q[0]], [q[1]])[0][0])]
    # the plain constructor refuses to drop the robot's safety margin silently (Space::new inflates by width / 2)
    with pytest.raises(pp.PathPlanningError):
        r.Space(b1, r.Robot(1.0, 1.0, 0.8), rings1)
    s1.close()
    s2.close()


def test_host_extend_step_pipelined_chunks_equal_the_single_stream_route(ctx, pp, O):
    """pp_rrt_extend on host arrays cuts a large batch on its default route into 2^18-query chunks that rotate over
    three streams; the answers must be the bytes of the single-stream route (selected here by asking for the scan
    verify, which returns the same flags) at sizes around the chunk boundaries, and the oracle's on a sub-sample."""
    n_nodes, chunk = 1 << 16, 1 << 18
    m = 3 * chunk + 777
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(m, n_nodes, world=400.0)
    bounds, rings = pp.synth.circle_world(1500, world=400.0)
    ctx.tree_upload(nx, ny, nyaw)
    ctx.obstacles_upload(bounds, rings)
    ref = ctx.rrt_extend(qx, qy, collide_flags=SCAN)
    for k in (2 * chunk - 1, 2 * chunk, 2 * chunk + 1, m):  # the first is below the pipelined route's threshold
        idx, yaw, ok = ctx.rrt_extend(qx[:k], qy[:k])
        assert np.array_equal(idx, ref[0][:k]) and np.array_equal(ok, ref[2][:k])
        assert np.array_equal(yaw, ref[1][:k], equal_nan=True)
    sub = np.arange(0, m, 997)
    oidx, _ = O.nn_brute(nx, ny, qx[sub], qy[sub])
    assert np.array_equal(ref[0][sub], oidx)
    W = O.OracleWorld(bounds, rings)
    assert np.array_equal(ref[2][sub], W.verify_segments(qx[sub], qy[sub], nx[oidx], ny[oidx], culled=True))
    assert 0.5 < ref[2].mean() < 0.99

"""CPU: the geo 0.12.2 predicate restatement, create_circle, exact NN and the exactness of the AABB culls
('property-test against the plain form in the oracle')."""
import json
import math
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

SQ = (np.array([0.0, 0.0, 10.0, 10.0, 0.0]), np.array([0.0, 10.0, 10.0, 0.0, 0.0]))


def test_create_circle(O, pp):
    for r, n in [(1.0, 7), (2.0, 13), (0.5, 4), (3.0, 19)]:
        x, y = O.create_circle(5.0, 5.0, r)
        assert len(x) in (n + 1, n + 2)  # n+1 points, +1 when the ring is not bit-closed
        assert x[0] == x[-1] and y[0] == y[-1]
        assert np.allclose(np.hypot(x - 5.0, y - 5.0), r)
        gx, gy = pp.synth.create_circle(5.0, 5.0, r)  # the product's generator must build identical rings
        assert np.array_equal(x, gx) and np.array_equal(y, gy)
    # bench circles (benches/all.rs:12-17): 4 of the 6 are not bit-closed before Polygon::new closes them
    extra = sum(len(O.create_circle(cx, cy, r)[0]) == math.ceil(2 * math.pi * r) + 2
                for (cx, cy, r) in [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2)])
    assert extra == 4


def test_point_position_semantics(O):
    rx, ry = SQ
    assert O.point_position(rx, ry, 5.0, 5.0) == 1
    assert O.point_position(rx, ry, 15.0, 5.0) == 0 and O.point_position(rx, ry, -1.0, 5.0) == 0
    assert O.point_position(rx, ry, 0.0, 5.0) == 2 and O.point_position(rx, ry, 10.0, 10.0) == 2  # boundary
    assert O.point_position(rx, ry, 5.0, 0.0) == 2
    assert O.point_position(rx, ry, float("nan"), 5.0) == 0
    # concave ring
    cx = np.array([0.0, 4.0, 4.0, 2.0, 2.0, 0.0, 0.0]); cy = np.array([0.0, 0.0, 4.0, 4.0, 2.0, 2.0, 0.0])
    assert O.point_position(cx, cy, 3.0, 3.0) == 1 and O.point_position(cx, cy, 1.0, 3.0) == 0


def test_lines_intersect_semantics(O):
    rx, ry = SQ
    assert O.lines_intersect(rx, ry, [5.0, 15.0], [5.0, 5.0])       # crosses the right edge
    assert not O.lines_intersect(rx, ry, [2.0, 8.0], [5.0, 5.0])    # strictly inside: rings do not meet it
    assert not O.lines_intersect(rx, ry, [0.0, 0.0], [2.0, 8.0])    # collinear with an edge: u_b == 0 skipped
    assert O.lines_intersect(rx, ry, [10.0, 12.0], [5.0, 5.0])      # touches at a parameter of exactly 0
    assert not O.lines_intersect(rx, ry, [5.0], [5.0])              # a 1-point line has no segments


def test_verify_semantics(O):
    ring = O.create_circle(5.0, 5.0, 1.0)
    W = O.OracleWorld(SQ, [ring])
    assert W.verify([1.0, 2.0], [1.0, 2.0])
    assert not W.verify([1.0, 9.0], [5.0, 5.0])        # passes through the obstacle
    assert not W.verify([5.0, 5.1], [5.0, 5.1])        # inside the obstacle (no ring crossing)
    assert not W.verify([1.0, 11.0], [1.0, 1.0])       # leaves the bounds
    assert not W.verify([0.0, 1.0], [1.0, 1.0])        # a point on the bounds ring is not contained
    assert W.verify([], []) and W.verify([1.0], [1.0])  # empty / single point
    assert not W.verify([5.0], [5.0])
    # contains() tests only the vertices against the bounds: a segment may leave and re-enter
    L = (np.array([0.0, 10.0, 10.0, 6.0, 6.0, 4.0, 4.0, 0.0, 0.0]), np.array([0.0, 0.0, 10.0, 10.0, 3.0, 3.0, 10.0, 10.0, 0.0]))
    assert O.OracleWorld(L, []).verify([2.0, 8.0], [8.0, 8.0])


def _random_world(O, rng, n_rings, world):
    rings = [O.create_circle(rng.uniform(0, world), rng.uniform(0, world), rng.uniform(0.5, 3.0)) for _ in range(n_rings)]
    b = (np.array([0.0, 0.0, world, world, 0.0]), np.array([0.0, world, world, 0.0, 0.0]))
    return O.OracleWorld(b, rings)


def test_culled_equals_plain_random(O):
    rng = np.random.default_rng(17)
    W = _random_world(O, rng, 150, 60.0)
    m = 20000
    ax, ay = rng.uniform(-1, 61, m), rng.uniform(-1, 61, m)
    ln, th = rng.choice([0.05, 1.0, 8.0], m), rng.uniform(-math.pi, math.pi, m)
    bx, by = ax + ln * np.cos(th), ay + ln * np.sin(th)
    plain = W.verify_segments(ax, ay, bx, by, culled=False)
    culled = W.verify_segments(ax, ay, bx, by, culled=True)
    assert np.array_equal(plain, culled) and 0 < plain.sum() < m


def test_culled_equals_plain_adversarial(O):
    """points and segments placed on / next to ring vertices, ring edges and AABB faces"""
    rng = np.random.default_rng(23)
    W = _random_world(O, rng, 40, 30.0)
    pts = []
    for rx, ry in W.rings():
        for i in range(len(rx) - 1):
            for t in (0.0, 0.5, 1.0):
                px, py = rx[i] + t * (rx[i + 1] - rx[i]), ry[i] + t * (ry[i + 1] - ry[i])
                for k in (-2, -1, 0, 1, 2):
                    pts.append((np.nextafter(px, px + k) if k else px, py))
                    pts.append((px, np.nextafter(py, py + k) if k else py))
        for px in (rx.min(), rx.max()):
            for py in (ry.min(), ry.max(), 0.5 * (ry.min() + ry.max())):
                for k in (-1e-13, 0.0, 1e-13):
                    pts.append((px + k, py))
    pts = np.array(pts)
    n = len(pts)
    for dx, dy in [(0.0, 0.0), (0.7, 0.0), (0.0, -0.7), (0.31, 0.53), (-4.0, 3.0)]:
        bx, by = pts[:, 0] + dx, pts[:, 1] + dy
        plain = W.verify_segments(pts[:, 0], pts[:, 1], bx, by, culled=False)
        culled = W.verify_segments(pts[:, 0], pts[:, 1], bx, by, culled=True)
        assert np.array_equal(plain, culled), (dx, dy, int((plain != culled).sum()), n)


@settings(max_examples=300, deadline=None)
@given(st.floats(-5, 35), st.floats(-5, 35), st.floats(-6, 6), st.floats(-6, 6), st.integers(0, 2 ** 31))
def test_culled_equals_plain_hypothesis(O, x, y, dx, dy, seed):
    W = _random_world(O, np.random.default_rng(seed % 7), 25, 30.0)
    lx, ly = [x, x + dx, x + dx * 0.5], [y, y + dy, y - dy]
    assert W.verify(lx, ly) == W.verify(lx, ly, culled=True)


def test_division_free_interval_test_counterexample(O):
    """0 <= x/y <= 1 is NOT equivalent to the sign/compare form when x/y underflows to -0.0;
    the oracle (and the kernels) therefore keep geo's division form"""
    x, y = -1e-300, 1e30
    assert 0.0 <= x / y <= 1.0 and not (0.0 <= x <= y)


def test_nn_brute_and_grid(O):
    rng = np.random.default_rng(4)
    for n, m, world in [(1, 5, 10.0), (50, 200, 10.0), (20000, 3000, 1000.0)]:
        nx, ny = rng.uniform(0, world, n), rng.uniform(0, world, n)
        qx, qy = rng.uniform(-0.1 * world, 1.1 * world, m), rng.uniform(-0.1 * world, 1.1 * world, m)
        i1, d1, dis = O.nn_brute(nx, ny, qx, qy, check_hypot=True)
        i2, d2 = O.nn_grid(nx, ny, qx, qy)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
        ref = np.argmin((nx[None, :] - qx[:, None]) ** 2 + (ny[None, :] - qy[:, None]) ** 2, axis=1) if n * m < 5e7 else None
        if ref is not None:
            assert np.array_equal(i1, ref.astype(np.uint32))
        assert dis <= m // 100  # argmin(hypot) (the metric written at src/rrt.rs:244) agrees with argmin(d2)
    # ties -> lowest index; empty tree -> 0xFFFFFFFF
    i, _ = O.nn_brute([1.0, 3.0, 1.0, 3.0], [0.0, 0.0, 0.0, 0.0], [2.0], [0.0])
    assert i[0] == 0
    assert O.nn_brute([], [], [1.0], [1.0])[0][0] == 0xFFFFFFFF and O.nn_grid([], [], [1.0], [1.0])[0][0] == 0xFFFFFFFF


def test_edge_polyline_and_line_to_origin(O):
    nx = np.array([0.0, 4.0, 7.0, 9.0]); ny = np.array([0.0, 1.0, 5.0, 9.0])
    parent = np.array([-1, 0, 1, 2], np.int32)
    nyaw = np.array([0.3] + [O.compute_yaw(nx[i], ny[i], nx[parent[i]], ny[parent[i]]) for i in range(1, 4)])
    lx, ly = O.line_to_origin(nx, ny, nyaw, parent, 3, 0.8, 0.1)
    chunks = [O.dubins_path(nx[i], ny[i], nyaw[i], nx[i - 1], ny[i - 1], nyaw[i - 1], 0.8, 0.1) for i in (3, 2, 1)]
    assert len(lx) == sum(len(c.x) for c in chunks) + 1 and (lx[-1], ly[-1]) == (0.0, 0.0)  # root's own point last
    assert (lx[0], ly[0]) == (9.0, 9.0)
    # per-edge decomposition used by the GPU path: samples ++ [parent point] (Q6/Q12)
    ex, ey = O.dubins_edge_polyline(nx[3], ny[3], nyaw[3], nx[2], ny[2], nyaw[2], 0.8, 0.1)
    assert np.array_equal(ex[:-1], chunks[0].x) and (ex[-1], ey[-1]) == (7.0, 5.0)
    # whole-chain verify == AND of per-edge verifies
    ring = O.create_circle(5.5, 3.0, 1.0)
    for rings in ([], [ring]):
        W = O.OracleWorld((np.array([-5.0, -5.0, 15.0, 15.0, -5.0]), np.array([-5.0, 15.0, 15.0, -5.0, -5.0])), rings)
        per_edge = W.verify_dubins_edges(nx[[3, 2, 1]], ny[[3, 2, 1]], nyaw[[3, 2, 1]], nx[[2, 1, 0]], ny[[2, 1, 0]],
                                         nyaw[[2, 1, 0]], 0.8, 0.1)
        assert W.verify(lx, ly) == bool(per_edge.all())
    # no feasible word -> fallback [(sx, sy), parent] (src/rrt.rs:313)
    fx, fy = O.dubins_edge_polyline(float("nan"), 0.0, 0.0, 1.0, 1.0, 0.0, 1.0, 0.1)
    assert len(fx) == 2 and fx[1] == 1.0


def test_transit_fixture_loads(O):
    conf = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "transit_world.json")))
    assert len(conf["bounds_x"]) == 20 and [len(r["x"]) for r in conf["rings"]] == [17, 10, 98]
    W = O.OracleWorld((conf["bounds_x"], conf["bounds_y"]), [(r["x"], r["y"]) for r in conf["rings"]])
    s, g = conf["start"], conf["goal"]
    assert W.verify([s[0]], [s[1]]) and W.verify([g[0]], [g[1]])  # start and goal are free points


def test_culls_on_near_parallel_extensions_differ_only_by_rounding_noise(O, pp):
    """DESIGN section 3 exactness (ii): the one class where the AABB culls can change geo's answer -- a line segment on
    the extension of a ring segment and parallel to it to within ~2^-45 rad: geo's denominator and numerators are
    rounding noise there and the exhaustive loop may report an intersection between segments that are far apart.
    Every such difference must be (a) exhaustive = blocked, culled = free and (b) free in EXACT rational geometry."""
    from conftest import exactly_free, near_parallel_edges
    bounds, rings = pp.synth.circle_world(24, world=100.0, rmin=1.0, rmax=3.0)
    W = O.OracleWorld(bounds, rings)
    ax, ay, bx, by = near_parallel_edges(W.rings())
    assert ax.size > 200_000
    plain = W.verify_segments(ax, ay, bx, by, culled=False)
    culled = W.verify_segments(ax, ay, bx, by, culled=True)
    bad = np.nonzero(plain != culled)[0]
    assert 0 < bad.size < ax.size // 1000  # the class exists, and it is tiny even on this adversarial set
    for i in bad:
        assert plain[i] == 0 and culled[i] == 1
        assert exactly_free(W, ax[i], ay[i], bx[i], by[i]), i

"""GPU parity: nearest neighbour, Space::verify on segments / polylines / Dubins edges and the batched
extend step against the CPU oracle, through the C-ABI.  NN indices and straight-edge flags are bit-exact."""
import json
import math
import os

import numpy as np
import pytest

from conftest import check_dubins_verdicts

pytestmark = pytest.mark.gpu

NN_DEFAULT, NN_PLAIN, NN_GRID, NN_UNSORTED, NN_SCAN = 0, 1, 2, 4, 8
DEFAULT, NO_CULL, USE_GRID, UNSORTED, SCAN = 0, 1, 2, 4, 8


def _bench_world(pp):
    """benches/all.rs:8-30 (no geo-offset inflation)"""
    rings = [pp.rrt.create_circle(c, r) for c, r in
             [((5.0, 5.0), 1.0), ((3.0, 6.0), 2.0), ((3.0, 8.0), 2.0), ((3.0, 10.0), 2.0), ((7.0, 5.0), 2.0), ((9.0, 5.0), 2.0)]]
    bounds = (np.array([-6.0, -6.0, 15.0, 15.0, -6.0]), np.array([-6.0, 15.0, 15.0, -6.0, -6.0]))
    return bounds, rings


@pytest.mark.parametrize("n_nodes,m", [(1, 10), (7, 300), (1000, 5000), (1024, 2048), (4096, 129), (4097, 70_001),
                                        (50_000, 20_000), (3000, 3)])
@pytest.mark.parametrize("flags", [NN_DEFAULT, NN_PLAIN, NN_GRID, NN_UNSORTED, NN_SCAN])
def test_nn_bit_exact(ctx, O, pp, n_nodes, m, flags):
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(m, n_nodes, world=1000.0)
    ctx.tree_upload(nx, ny, nyaw)
    idx, d2 = ctx.nn(qx, qy, flags=flags)
    oidx, od2 = O.nn_brute(nx, ny, qx, qy)
    assert np.array_equal(idx, oidx)
    assert np.array_equal(d2, od2)  # same non-fused arithmetic -> same bits


@pytest.mark.parametrize("flags", [NN_DEFAULT, NN_PLAIN, NN_GRID, NN_UNSORTED, NN_SCAN])
def test_nn_ties_lowest_index(ctx, O, flags):
    # lattice nodes duplicated 3x: every query has exact ties, some at equal distance to 4 lattice points
    g = np.arange(0, 20, dtype=np.float64)
    gx, gy = np.meshgrid(g, g)
    nx = np.tile(gx.ravel(), 3)
    ny = np.tile(gy.ravel(), 3)
    qx = np.concatenate([gx.ravel()[:150] + 0.5, gx.ravel()[:150], np.array([-5.0, 30.0, 9.5])])
    qy = np.concatenate([gy.ravel()[:150] + 0.5, gy.ravel()[:150] + 0.25, np.array([-5.0, 30.0, 9.5])])
    ctx.tree_upload(nx, ny)
    idx, d2 = ctx.nn(qx, qy, flags=flags)
    oidx, od2 = O.nn_brute(nx, ny, qx, qy)
    assert np.array_equal(idx, oidx) and np.array_equal(d2, od2)
    assert idx.max() < gx.size  # always the first copy


def test_nn_empty_tree_and_append(ctx, O, pp):
    ctx.tree_upload(np.zeros(0), np.zeros(0))
    idx = ctx.nn([1.0, 2.0], [3.0, 4.0], want_d2=False)
    assert np.all(idx == 0xFFFFFFFF)
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(500, 3000, world=50.0)
    ctx.tree_upload(nx[:10], ny[:10], nyaw[:10])
    for a, b in [(10, 11), (11, 1024), (1024, 1030), (1030, 3000)]:
        ctx.tree_append(nx[a:b], ny[a:b], nyaw[a:b], np.arange(a, b) - 1)
        assert ctx.tree_size == b
        for flags in (NN_DEFAULT, NN_GRID, NN_SCAN):
            idx = ctx.nn(qx, qy, flags=flags, want_d2=False)
            assert np.array_equal(idx, O.nn_brute(nx[:b], ny[:b], qx, qy)[0])
    one = ctx.nn([qx[0]], [qy[0]], want_d2=False)  # scalar call -> wide kernel
    assert one[0] == O.nn_brute(nx, ny, qx[:1], qy[:1])[0][0]


def test_nn_grid_clustered_tree(ctx, O):
    """device-built node grid on a hostile distribution: tight clusters (thousands of nodes per cell), exact duplicates,
    far outliers that stretch the bounding box, a non-finite node; queries inside, between and far outside"""
    rng = np.random.default_rng(11)
    parts_x, parts_y = [], []
    for cx, cy, s, k in [(10.0, 10.0, 1e-3, 6000), (-400.0, 250.0, 0.5, 3000), (10.0, 10.002, 1e-6, 2000)]:
        parts_x.append(cx + s * rng.standard_normal(k))
        parts_y.append(cy + s * rng.standard_normal(k))
    parts_x += [np.array([1e6, -1e6, np.nan, 3.0]), np.full(500, 77.0)]
    parts_y += [np.array([1e6, -1e6, 1.0, np.inf]), np.full(500, -12.5)]
    nx, ny = np.concatenate(parts_x), np.concatenate(parts_y)
    qx = np.concatenate([10.0 + 2e-3 * rng.standard_normal(300), rng.uniform(-500, 500, 300), [5e6, -3e5, 77.0, 0.0]])
    qy = np.concatenate([10.0 + 2e-3 * rng.standard_normal(300), rng.uniform(-500, 500, 300), [-5e6, 1e5, -12.5, 0.0]])
    ctx.tree_upload(nx, ny)
    oidx, od2 = O.nn_brute(nx, ny, qx, qy)
    for flags in (NN_DEFAULT, NN_GRID, NN_SCAN):
        idx, d2 = ctx.nn(qx, qy, flags=flags)
        assert np.array_equal(idx, oidx) and np.array_equal(d2, od2), flags
    # appending invalidates the grid; the next default call rebuilds it on the device
    ctx.tree_append(qx[:50] + 1e-4, qy[:50], np.zeros(50), np.zeros(50, dtype=np.int32))
    nx2, ny2 = np.concatenate([nx, qx[:50] + 1e-4]), np.concatenate([ny, qy[:50]])
    idx = ctx.nn(qx, qy, want_d2=False)
    assert np.array_equal(idx, O.nn_brute(nx2, ny2, qx, qy)[0])
    one = ctx.nn(qx[:3], qy[:3], want_d2=False)  # few queries on a current grid: grid path again
    assert np.array_equal(one, idx[:3])


def test_nn_grid_degenerate_geometries(ctx, O):
    """node sets that stress the device-built grid's geometry: all nodes identical, collinear (zero height), a 1e6 : 1
    aspect ratio, the 4 096-node switch point, and growth by appends past the allocated capacity"""
    rng = np.random.default_rng(17)
    q = (rng.uniform(-2, 12, 700), rng.uniform(-2, 12, 700))
    sets = {
        "identical": (np.full(5000, 3.25), np.full(5000, -1.5)),
        "collinear_x": (rng.uniform(0, 10, 6000), np.full(6000, 4.0)),
        "collinear_y": (np.full(4096, 7.0), rng.uniform(0, 10, 4096)),
        "thin": (rng.uniform(0, 1e6, 9000), rng.uniform(0, 1.0, 9000)),
        "switch": (rng.uniform(0, 10, 4096), rng.uniform(0, 10, 4096)),
        "below_switch": (rng.uniform(0, 10, 4095), rng.uniform(0, 10, 4095)),
    }
    for name, (nx, ny) in sets.items():
        qx, qy = (q[0] * (1e5 if name == "thin" else 1.0), q[1] * (0.1 if name == "thin" else 1.0))
        ctx.tree_upload(nx, ny)
        oidx, od2 = O.nn_brute(nx, ny, qx, qy)
        for flags in (NN_DEFAULT, NN_GRID):
            idx, d2 = ctx.nn(qx, qy, flags=flags)
            assert np.array_equal(idx, oidx) and np.array_equal(d2, od2), (name, flags)
    # growth: 5 000 -> 40 000 nodes in uneven appends, default method after every step
    nx, ny = rng.uniform(0, 10, 40_000), rng.uniform(0, 10, 40_000)
    ctx.tree_upload(nx[:5000], ny[:5000])
    for a, b in [(5000, 5001), (5001, 9000), (9000, 20_000), (20_000, 40_000)]:
        ctx.tree_append(nx[a:b], ny[a:b], np.zeros(b - a), np.zeros(b - a, dtype=np.int32))
        idx = ctx.nn(q[0], q[1], want_d2=False)
        assert np.array_equal(idx, O.nn_brute(nx[:b], ny[:b], q[0], q[1])[0]), b


def test_nn_nonfinite(ctx, O):
    nx = np.array([0.0, np.nan, 5.0, np.inf, 1.0])
    ny = np.array([0.0, 1.0, np.nan, 2.0, 1.0])
    qx = np.array([0.9, np.nan, 100.0])
    qy = np.array([0.9, 0.0, 100.0])
    ctx.tree_upload(nx, ny)
    for flags in (NN_DEFAULT, NN_PLAIN, NN_GRID, NN_UNSORTED, NN_SCAN):
        idx = ctx.nn(np.tile(qx, 40), np.tile(qy, 40), flags=flags, want_d2=False)  # > 64 queries: scan kernels
        assert np.array_equal(idx, np.tile(O.nn_brute(nx, ny, qx, qy)[0], 40)), flags
        idx = ctx.nn(qx, qy, flags=flags, want_d2=False)
        assert np.array_equal(idx, O.nn_brute(nx, ny, qx, qy)[0]), flags


@pytest.mark.parametrize("flags", [DEFAULT, NO_CULL, USE_GRID, UNSORTED, SCAN])
def test_collide_segments_random_world(ctx, O, pp, flags):
    bounds, rings = pp.synth.circle_world(300, world=100.0, rmin=1.0, rmax=3.0)
    ctx.obstacles_upload(bounds, rings)
    W = O.OracleWorld(bounds, rings)
    m = 20_000 if flags != NO_CULL else 4000
    rng = np.random.default_rng(5)
    ax, ay = rng.uniform(-2, 102, m), rng.uniform(-2, 102, m)
    ln = rng.choice([0.3, 2.0, 15.0, 60.0], m)  # 60: boxes that cover more grid cells than there are rings
    th = rng.uniform(-math.pi, math.pi, m)
    bx, by = ax + ln * np.cos(th), ay + ln * np.sin(th)
    ok = ctx.collide_segments(ax, ay, bx, by, flags=flags)
    ook = W.verify_segments(ax, ay, bx, by)
    assert np.array_equal(ok, ook)
    assert 0 < ok.sum() < m  # both outcomes present


def test_collide_segments_bench_world_and_boundary_cases(ctx, O, pp):
    bounds, rings = _bench_world(pp)
    ctx.obstacles_upload(bounds, rings)
    W = O.OracleWorld(bounds, rings)
    rx, ry = rings[1]
    ax = np.array([-5.0, -6.0, 0.0, rx[0], rx[2], 3.0, 3.0, -5.9999999, 14.0, 5.0, 0.0, -7.0, 3.0 + 1e-13])
    ay = np.array([-5.0, 0.0, 0.0, ry[0], ry[2], 6.0, 6.0, -5.0, 14.0, 3.9, 15.0, 0.0, 6.0])
    bx = np.array([-4.0, -5.0, 14.0, rx[0], rx[3], 3.1, 12.0, -5.0, 14.5, 5.0, 0.0, 0.0, 3.0 + 1e-13])
    by = np.array([-4.0, 0.0, 0.0, ry[0], ry[3], 6.1, 6.0, -5.0, 14.5, 4.1, 14.0, 0.0, 6.0])
    for flags in (DEFAULT, NO_CULL, USE_GRID, UNSORTED, SCAN):
        ok = ctx.collide_segments(ax, ay, bx, by, flags=flags)
        assert np.array_equal(ok, W.verify_segments(ax, ay, bx, by)), flags
    # degenerate segments (a == b) and points exactly on the bounds ring are not contained
    on = ctx.collide_segments([-6.0, 15.0, 0.0], [0.0, 3.0, -6.0], [-6.0, 15.0, 0.0], [0.0, 3.0, -6.0])
    assert not on.any()


def test_collide_transit_fixture(ctx, O):
    """the reference's only shipped world (examples/rrt/transit.debug.json), re-encoded as tests/golden/transit_world.json"""
    path = os.path.join(os.path.dirname(__file__), "golden", "transit_world.json")
    conf = json.load(open(path))
    b = np.stack([np.array(conf["bounds_x"]), np.array(conf["bounds_y"])], axis=1)
    rings = [(np.array(o["x"]), np.array(o["y"])) for o in conf["rings"]]
    bounds = (b[:, 0].copy(), b[:, 1].copy())
    ctx.obstacles_upload(bounds, rings)
    W = O.OracleWorld(bounds, rings)
    rng = np.random.default_rng(9)
    m = 30_000
    ax = rng.uniform(b[:, 0].min() - 2, b[:, 0].max() + 2, m)
    ay = rng.uniform(b[:, 1].min() - 2, b[:, 1].max() + 2, m)
    th = rng.uniform(-math.pi, math.pi, m)
    ln = rng.choice([0.1, 3.0, 30.0], m)
    bx, by = ax + ln * np.cos(th), ay + ln * np.sin(th)
    want = W.verify_segments(ax, ay, bx, by)
    for flags in (DEFAULT, USE_GRID, UNSORTED, SCAN):
        assert np.array_equal(ctx.collide_segments(ax, ay, bx, by, flags=flags), want), flags
    assert 0 < want.sum() < m
    # start -> goal as one straight line, and as the reference's start/goal poses
    s, g = conf["start"], conf["goal"]
    assert ctx.collide_segments([s[0]], [s[1]], [g[0]], [g[1]])[0] == W.verify_segments([s[0]], [s[1]], [g[0]], [g[1]])[0]


def test_verify_polylines(ctx, O, pp):
    bounds, rings = pp.synth.circle_world(200, world=100.0)
    ctx.obstacles_upload(bounds, rings)
    W = O.OracleWorld(bounds, rings)
    rng = np.random.default_rng(11)
    lines = []
    for k in range(400):
        npts = int(rng.choice([0, 1, 2, 3, 31, 32, 33, 64, 100]))
        x0, y0 = rng.uniform(5, 95, 2)
        stepx, stepy = rng.normal(0, 0.4, npts), rng.normal(0, 0.4, npts)
        lines.append((x0 + np.cumsum(stepx), y0 + np.cumsum(stepy)))
    want = np.array([W.verify(lx, ly) for lx, ly in lines], np.uint8)
    for flags in (DEFAULT, NO_CULL):
        assert np.array_equal(ctx.verify_polylines(lines, flags=flags), want), flags
    assert 0 < want.sum() < len(lines)


def test_collide_nonfinite_coordinates(ctx, O, pp):
    """NaN / infinite end points and huge finite ones: never `contained` by the bounds, whatever the integer cell
    arithmetic makes of them (saturating conversions, NaN -> cell 0)"""
    bounds, rings = pp.synth.circle_world(300, world=100.0, rmin=1.0, rmax=3.0)
    ctx.obstacles_upload(bounds, rings)
    W = O.OracleWorld(bounds, rings)
    nan, inf = float("nan"), float("inf")
    bad = [nan, inf, -inf, 1e300, -1e300, 1e15, -4e9]
    ax, ay, bx, by = [], [], [], []
    for v in bad:
        for pos in range(4):
            p = [50.0, 50.0, 50.5, 50.5]
            p[pos] = v
            ax.append(p[0]); ay.append(p[1]); bx.append(p[2]); by.append(p[3])
    ax += [50.0] * 3; ay += [50.0] * 3; bx += [50.5, 50.0, 99.9]; by += [50.5, 50.0, 0.1]  # a few ordinary ones
    ax, ay, bx, by = (np.array(v) for v in (ax, ay, bx, by))
    want = W.verify_segments(ax, ay, bx, by)
    assert not want[: 4 * len(bad)].any()
    for flags in (DEFAULT, SCAN, UNSORTED, NO_CULL):
        assert np.array_equal(ctx.collide_segments(ax, ay, bx, by, flags=flags), want), flags
    lines = [(np.array([50.0, v, 51.0]), np.array([50.0, 50.0, 50.0])) for v in bad]
    lines += [(np.array([50.0, 50.2, 50.4]), np.array([v, 50.0, 50.0])) for v in bad]
    assert not ctx.verify_polylines(lines).any()
    # NN queries with non-finite coordinates through the grid path (tree large enough for the grid)
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(200, 5000, world=100.0)
    ctx.tree_upload(nx, ny, nyaw)
    qx[:7], qy[7:14] = bad, bad
    for flags in (NN_DEFAULT, NN_GRID, NN_SCAN):
        idx = ctx.nn(qx, qy, flags=flags, want_d2=False)
        assert np.array_equal(idx, O.nn_brute(nx, ny, qx, qy)[0]), flags
    # Dubins edges with a non-finite pose component: no feasible word -> [(sx, sy), parent] -> not contained
    e = 7 * 6
    sx, sy, syaw, ex, ey, eyaw = (np.full(e, v) for v in (40.0, 40.0, 0.3, 43.0, 41.0, -0.2))
    for k, v in enumerate(bad):
        for pos, arr in enumerate((sx, sy, syaw, ex, ey, eyaw)):
            arr[6 * k + pos] = v
    ok = ctx.collide_dubins(sx, sy, syaw, ex, ey, eyaw, 1.0, 0.1)
    want = W.verify_dubins_edges(sx, sy, syaw, ex, ey, eyaw, 1.0, 0.1)
    finite_yaw_only = np.zeros(e, bool)
    finite_yaw_only[[6 * k + p for k in range(3, 7) for p in (2, 5)]] = True  # huge but finite yaws are legal poses
    declined = want == 255  # the oracle refuses to replay a path of ~1e16 samples (positions at 1e15); the GPU reports blocked
    assert not ok[~finite_yaw_only].any()
    assert np.array_equal(ok[~finite_yaw_only & ~declined], want[~finite_yaw_only & ~declined])
    # huge but finite yaws (1e15 ... 1e300 rad): the last digits of the argument reduction decide -- the oracle flags them
    f = finite_yaw_only
    check_dubins_verdicts(O, W, ok[f], sx[f], sy[f], syaw[f], ex[f], ey[f], eyaw[f], 1.0, 0.1, want=want[f])


def test_world_without_obstacles_and_degenerate_rings(ctx, O, pp):
    """no rings at all; then rings of 0, 1 and 2 points and a zero-area ring next to ordinary ones"""
    bounds = (np.array([0.0, 0.0, 50.0, 50.0, 0.0]), np.array([0.0, 50.0, 50.0, 0.0, 0.0]))
    rng = np.random.default_rng(31)
    m = 5000
    ax, ay = rng.uniform(-1, 51, m), rng.uniform(-1, 51, m)
    th, ln = rng.uniform(-math.pi, math.pi, m), rng.choice([0.2, 4.0, 30.0], m)
    bx, by = ax + ln * np.cos(th), ay + ln * np.sin(th)
    sx, sy, syaw, ex, ey, eyaw = pp.synth.dubins_edges(600, world=50.0, reach=8.0)
    worlds = [[], [(np.zeros(0), np.zeros(0)), (np.array([10.0]), np.array([10.0])),
                   (np.array([20.0, 24.0]), np.array([20.0, 23.0])),
                   (np.array([30.0, 34.0, 30.0]), np.array([30.0, 30.0, 30.0])),  # zero area
                   pp.rrt.create_circle((25.0, 25.0), 3.0), pp.rrt.create_circle((12.0, 40.0), 1.0)]]
    for rings in worlds:
        ctx.obstacles_upload(bounds, rings)
        W = O.OracleWorld(bounds, rings)
        want = W.verify_segments(ax, ay, bx, by)
        for flags in (DEFAULT, SCAN, UNSORTED, NO_CULL):
            assert np.array_equal(ctx.collide_segments(ax, ay, bx, by, flags=flags), want), (len(rings), flags)
        ok = ctx.collide_dubins(sx, sy, syaw, ex, ey, eyaw, 1.0, 0.1)
        check_dubins_verdicts(O, W, ok, sx, sy, syaw, ex, ey, eyaw, 1.0, 0.1)
        assert 0 < want.sum() < m


def test_collide_dubins_zero_length_and_tiny_paths(ctx, O, pp):
    """identical poses (every sample trimmed: the polyline is the parent point alone), pure turns on the spot and
    paths shorter than one step, in free space, inside an obstacle and outside the bounds"""
    bounds, rings = pp.synth.circle_world(150, world=60.0, rmin=1.0, rmax=3.0)
    ctx.obstacles_upload(bounds, rings)
    W = O.OracleWorld(bounds, rings)
    rng = np.random.default_rng(41)
    n = 1500
    sx, sy = rng.uniform(-2, 62, n), rng.uniform(-2, 62, n)
    syaw = rng.uniform(-math.pi, math.pi, n)
    kind = rng.integers(0, 4, n)
    ex = sx + np.where(kind == 2, rng.uniform(-0.02, 0.02, n), 0.0)
    ey = sy + np.where(kind == 2, rng.uniform(-0.02, 0.02, n), 0.0)
    eyaw = np.where(kind == 0, syaw, np.where(kind == 1, syaw + rng.uniform(-3, 3, n), rng.uniform(-math.pi, math.pi, n)))
    ex = np.where(kind == 3, sx + rng.uniform(-0.3, 0.3, n), ex)
    for radius, step in [(1.0, 0.1), (0.5, 0.05)]:
        ok = ctx.collide_dubins(sx, sy, syaw, ex, ey, eyaw, radius, step)
        want = W.verify_dubins_edges(sx, sy, syaw, ex, ey, eyaw, radius, step)
        # identical / coincident poses are near-ties by construction: no bound on the fragile share here
        check_dubins_verdicts(O, W, ok, sx, sy, syaw, ex, ey, eyaw, radius, step, want=want, max_fragile=None)
        assert 0 < want.sum() < n


def test_many_vertex_rings(ctx, O, pp):
    """rings of 3 ... 700 points (stars, non-convex): the grouped narrow phase strides a ring eight segments at a time,
    so rings longer than one pass, longer than a warp, and the tiny ones all take different trip counts"""
    rng = np.random.default_rng(23)
    rings = []
    for k in range(60):
        n = int(rng.choice([3, 4, 8, 9, 16, 17, 33, 64, 65, 130, 700]))
        cx, cy = rng.uniform(8, 92, 2)
        th = np.sort(rng.uniform(0, 2 * math.pi, n))
        rad = rng.uniform(1.0, 4.0) * (1.0 + 0.45 * np.sin(7 * th + rng.uniform(0, 6)))  # star-ish, non-convex
        rings.append((cx + rad * np.cos(th), cy + rad * np.sin(th)))
    bounds = (np.array([0.0, 0.0, 100.0, 100.0, 0.0]), np.array([0.0, 100.0, 100.0, 0.0, 0.0]))
    ctx.obstacles_upload(bounds, rings)
    W = O.OracleWorld(bounds, rings)
    m = 40_000
    ax, ay = rng.uniform(-1, 101, m), rng.uniform(-1, 101, m)
    ln = rng.choice([0.05, 0.5, 3.0, 25.0], m)
    th = rng.uniform(-math.pi, math.pi, m)
    bx, by = ax + ln * np.cos(th), ay + ln * np.sin(th)
    want = W.verify_segments(ax, ay, bx, by)
    for flags in (DEFAULT, SCAN, UNSORTED):
        assert np.array_equal(ctx.collide_segments(ax, ay, bx, by, flags=flags), want), flags
    assert 0.05 < want.mean() < 0.95
    lines = []
    for k in range(300):
        npts = int(rng.choice([2, 5, 40, 90]))
        x0, y0 = rng.uniform(5, 95, 2)
        lines.append((x0 + np.cumsum(rng.normal(0, 0.3, npts)), y0 + np.cumsum(rng.normal(0, 0.3, npts))))
    wantl = np.array([W.verify(lx, ly) for lx, ly in lines], np.uint8)
    assert np.array_equal(ctx.verify_polylines(lines), wantl) and 0 < wantl.sum() < len(lines)
    e = 2000
    sx, sy, syaw, ex, ey, eyaw = pp.synth.dubins_edges(e, world=100.0, reach=10.0)
    ok = ctx.collide_dubins(sx, sy, syaw, ex, ey, eyaw, 1.0, 0.05)
    check_dubins_verdicts(O, W, ok, sx, sy, syaw, ex, ey, eyaw, 1.0, 0.05)


@pytest.mark.parametrize("radius,step", [(1.0, 0.05), (0.8, 0.1)])
def test_collide_dubins_edges(ctx, O, pp, radius, step):
    bounds, rings = pp.synth.circle_world(400, world=100.0, rmin=0.5, rmax=1.5)
    ctx.obstacles_upload(bounds, rings)
    W = O.OracleWorld(bounds, rings)
    e = 3000
    sx, sy, syaw, ex, ey, eyaw = pp.synth.dubins_edges(e, world=100.0, reach=12.0)
    ok = ctx.collide_dubins(sx, sy, syaw, ex, ey, eyaw, radius, step)
    want = W.verify_dubins_edges(sx, sy, syaw, ex, ey, eyaw, radius, step)
    # sample coordinates agree to 1e-9, so a flag can only differ where the oracle's margin is below that
    check_dubins_verdicts(O, W, ok, sx, sy, syaw, ex, ey, eyaw, radius, step, want=want)
    assert 0 < want.sum() < e
    ok2 = ctx.collide_dubins(sx[:300], sy[:300], syaw[:300], ex[:300], ey[:300], eyaw[:300], radius, step, flags=NO_CULL)
    check_dubins_verdicts(O, W, ok2, sx[:300], sy[:300], syaw[:300], ex[:300], ey[:300], eyaw[:300], radius, step,
                          want=want[:300])
    assert np.array_equal(ok2, ok[:300])  # the culls never change a verdict of this library's own samples


def test_extend_step(ctx, O, pp):
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(30_000, 20_000, world=200.0)
    bounds, rings = pp.synth.circle_world(500, world=200.0)
    ctx.tree_upload(nx, ny, nyaw)
    ctx.obstacles_upload(bounds, rings)
    W = O.OracleWorld(bounds, rings)
    oidx, _ = O.nn_brute(nx, ny, qx, qy)
    want = W.verify_segments(qx, qy, nx[oidx], ny[oidx])
    wyaw = np.arctan2(ny[oidx] - qy, nx[oidx] - qx)
    # 16 = PP_COLLIDE_FUSED: queries binned by node-grid block, then NN + yaw + verify in one cell-coherent launch
    for nnf, cf in [(NN_DEFAULT, DEFAULT), (NN_GRID, USE_GRID), (NN_DEFAULT, 16), (NN_PLAIN, DEFAULT), (NN_SCAN, SCAN)]:
        idx, yaw, ok = ctx.rrt_extend(qx, qy, nn_flags=nnf, collide_flags=cf)
        assert np.array_equal(idx, oidx) and np.array_equal(ok, want)
        assert np.abs(yaw - wyaw).max() < 1e-12
    assert 0 < want.sum() < want.size


def test_rrt_planner_drop_in(pp, ctx, O):
    """benches/all.rs:6-46 world through the module mirror: plan_one keeps tree and GPU mirror in step"""
    r = pp.rrt
    bounds, rings = _bench_world(pp)
    space = r.Space.from_inflated(bounds, r.Robot(1.0, 1.0, 0.8), rings, ctx=ctx, seed=1234)
    planner = r.RRT((-5.0, -5.0), math.radians(-45.0), (6.0, 10.0), math.radians(45.0), 8000, 0.1, space)
    W = O.OracleWorld(bounds, rings)
    for _ in range(40):
        planner.plan_one()
    assert ctx.tree_size == len(planner.nodes) >= 2
    # every node in the tree has a chain the oracle also accepts (same polyline, same verify)
    nx = np.array([n.point[0] for n in planner.nodes]); ny = np.array([n.point[1] for n in planner.nodes])
    nyaw = np.array([n.yaw for n in planner.nodes])
    par = np.array([planner.nodes.index(n.parent) if n.parent is not None else -1 for n in planner.nodes], np.int32)
    for i in range(1, min(len(planner.nodes), 12)):
        lx, ly = O.line_to_origin(nx, ny, nyaw, par, i, 0.8, 0.1)
        gx, gy = r.line_to_origin(planner.nodes[i], 0.8, 0.1, ctx)
        assert len(lx) == len(gx) and np.abs(lx - gx).max() < 1e-9 and np.abs(ly - gy).max() < 1e-9
        assert W.verify(lx, ly)
    q = (1.0, 1.0)
    assert planner.get_nearest_node(q) is planner.nodes[int(O.nn_brute(nx, ny, [q[0]], [q[1]])[0][0])]


def _sequential_optimize(planner, r, node, i):
    """the reference's loop (src/rrt.rs:463-487) with one verify_node call per candidate"""
    if i >= r.RECURSION_LIMIT:
        return None
    for to_node in reversed(list(r.NodeIter(node))):
        new_node = r.Node(node.get_coord(), to_node)
        if planner.verify_node(new_node):
            deeper = _sequential_optimize(planner, r, to_node, i + 1)
            return r.Node(node.get_coord(), deeper) if deeper is not None else new_node
    return None


def _chain_points(r, node):
    return [n.point for n in r.NodeIter(node)]


def _flat_tree(planner):
    nx = np.array([n.point[0] for n in planner.nodes]); ny = np.array([n.point[1] for n in planner.nodes])
    nyaw = np.array([n.yaw for n in planner.nodes])
    par = np.array([planner._slot[id(n.parent)] if n.parent is not None else -1 for n in planner.nodes], np.int32)
    return nx, ny, nyaw, par


def _chain_poses(r, node):
    return np.array([(n.point[0], n.point[1], n.yaw) for n in r.NodeIter(node)])


def _assert_optimize_matches_oracle(O, W, r, planner, nodes, results, radius, step):
    """GPU-backed optimize() results against the oracle's restatement of src/rrt.rs:463-487 on the same flat tree: same
    chain (points and yaws bit-equal: both sides copy the points and take glibc's atan2) unless the oracle marks one of
    the verify decisions it took as fragile at the 1e-9 sample tolerance.  Returns (compared, fragile)."""
    nx, ny, nyaw, par = _flat_tree(planner)
    compared = fragile = 0
    for node, got in zip(nodes, results):
        chain, flags, _ = O.optimize(W, nx, ny, nyaw, par, planner._slot[id(node)], radius, step)
        same = (got is None) == (chain is None) and (got is None or np.array_equal(_chain_poses(r, got), chain))
        if flags:
            fragile += 1
            continue
        assert same, (planner._slot[id(node)], None if got is None else _chain_poses(r, got), chain)
        compared += 1
    return compared, fragile


def test_batched_shortcutting_matches_the_sequential_loop(pp, ctx, O):
    """optimize() verifies all shortcut candidates of a level in one fused launch and must pick
    exactly what the reference's candidate-by-candidate loop picks"""
    r = pp.rrt
    bounds, rings = _bench_world(pp)
    space = r.Space.from_inflated(bounds, r.Robot(1.0, 1.0, 0.8), rings, ctx=ctx, seed=99)
    planner = r.RRT((-5.0, -5.0), math.radians(-45.0), (6.0, 10.0), math.radians(45.0), 8000, 0.1, space)
    for _ in range(150):
        planner.plan_one()
    deep = sorted(planner.nodes, key=lambda n: -len(list(r.NodeIter(n))))[:6]
    assert len(list(r.NodeIter(deep[0]))) >= 3
    got = []
    for node in deep:
        a = planner.optimize(node, 0)
        b = _sequential_optimize(planner, r, node, 0)
        assert (a is None) == (b is None)
        if a is not None:
            assert _chain_points(r, a) == _chain_points(r, b)
            # (no "shorter than before" assertion: when the shortcut reaches the root, the reference recurses on
            # the root itself and nests Node(root, root) until RECURSION_LIMIT -- a quirk both versions reproduce)
        got.append(a)
    # ... and what the ORACLE's optimize (oracle/pp_oracle.c ppo_optimize, the literal candidate loop with whole-chain
    # line_to_origin + verify per candidate) picks on the same tree
    W = O.OracleWorld(bounds, rings)
    more = planner.nodes[1:40]
    compared, fragile = _assert_optimize_matches_oracle(O, W, r, planner, deep + more,
                                                        got + [planner.optimize(n, 0) for n in more], 0.8, 0.1)
    assert compared >= 20 and fragile <= compared


def test_check_finish_many_matches_check_finish(pp, ctx, O):
    """the round-level goal check (all nodes' optimize recursions level by level in one launch each, one batched
    line_to_origin, one verify launch) returns exactly the lines of the per-node check_finish"""
    r = pp.rrt
    bounds, rings = _bench_world(pp)
    space = r.Space.from_inflated(bounds, r.Robot(1.0, 1.0, 0.8), rings, ctx=ctx, seed=7)
    planner = r.RRT((-5.0, -5.0), math.radians(-45.0), (6.0, 10.0), math.radians(45.0), 8000, 0.1, space)
    for _ in range(250):
        planner.plan_one()
    nodes = planner.nodes[1:60] + sorted(planner.nodes, key=lambda n: -len(list(r.NodeIter(n))))[:12]
    many = planner.check_finish_many(nodes)
    assert len(many) == len(nodes)
    found = 0
    for n, got in zip(nodes, many):
        want = planner.check_finish(n)
        assert (got is None) == (want is None)
        if got is not None:
            found += 1
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert found >= 1
    opt_many = planner._optimize_many(nodes)
    for n, a in zip(nodes, opt_many):
        b = planner.optimize(n, 0)
        assert (a is None) == (b is None)
        if a is not None:
            assert _chain_points(r, a) == _chain_points(r, b)
    assert planner.check_finish_many([]) == []
    # the oracle's check_finish (src/rrt.rs:428-438 through finalize :503-540 and optimize_from_goal :489-501) on the
    # same flat tree: same verdict, same optimised chain, same line to 1e-9 -- unless the oracle flags a decision
    W = O.OracleWorld(bounds, rings)
    nx, ny, nyaw, par = _flat_tree(planner)
    compared, _ = _assert_optimize_matches_oracle(O, W, r, planner, nodes, opt_many, 0.8, 0.1)
    assert compared >= len(nodes) // 2
    checked = some = exact = 0
    for n, got in zip(nodes, many):
        fin = O.check_finish(W, nx, ny, nyaw, par, planner._slot[id(n)], planner.goal, planner.goal_yaw, 0.8, 0.1)
        if fin.flags:
            continue
        checked += 1
        assert (got is not None) == fin.ok, planner._slot[id(n)]
        if got is not None:
            some += 1
            tol = 1e-9 * 15.0  # 1e-9 relative, coordinates up to 15
            if fin.line_flags == 0:  # sample for sample
                assert len(got[0]) == len(fin.line[0])
                assert np.abs(got[0] - fin.line[0]).max() < tol and np.abs(got[1] - fin.line[1]).max() < tol
                exact += 1
            else:
                # an edge of the line has a knife-edge sample count (typically the loop edge of the root-nesting
                # quirk, whose end point returns to local x = 0 up to rounding: the trim at src/dubins.rs:281-288 then
                # keeps or drops one more sample).  Every GPU sample must still be an oracle sample, and at most one
                # sample per chain edge may be missing on either side.
                from scipy.spatial import cKDTree
                go, orc = np.stack(got, 1), np.stack(fin.line, 1)
                d_go, _ = cKDTree(orc).query(go)
                d_or, _ = cKDTree(go).query(orc)
                edges = len(fin.chain) - 1
                assert (d_go > tol).sum() <= edges and (d_or > tol).sum() <= edges
                assert abs(len(go) - len(orc)) <= edges
    assert checked >= len(nodes) // 2 and some >= 1


def test_plan_rounds_keeps_the_tree_invariant(pp, ctx, O):
    """batched rounds; every inserted node's chain must verify under the oracle, the device
    mirror must track the host tree, and a returned path must verify"""
    r = pp.rrt
    bounds, rings = _bench_world(pp)
    space = r.Space.from_inflated(bounds, r.Robot(1.0, 1.0, 0.8), rings, ctx=ctx, seed=4242)
    planner = r.RRT((-5.0, -5.0), math.radians(-45.0), (6.0, 10.0), math.radians(45.0), 8000, 0.1, space)
    path = planner.plan_rounds(batch=128, max_iter=1024)
    assert ctx.tree_size == len(planner.nodes) > 20
    W = O.OracleWorld(bounds, rings)
    nx = np.array([n.point[0] for n in planner.nodes]); ny = np.array([n.point[1] for n in planner.nodes])
    nyaw = np.array([n.yaw for n in planner.nodes])
    par = np.array([planner._slot[id(n.parent)] if n.parent is not None else -1 for n in planner.nodes], np.int32)
    for i in list(range(1, 15)) + list(range(len(planner.nodes) - 15, len(planner.nodes))):
        lx, ly = O.line_to_origin(nx, ny, nyaw, par, i, 0.8, 0.1)
        assert W.verify(lx, ly), i
    # NN on the grown device tree still agrees with the oracle
    q = np.array([[0.0, 0.0], [5.0, 12.0], [-4.0, 9.0]])
    assert np.array_equal(ctx.nn(q[:, 0], q[:, 1], want_d2=False), O.nn_brute(nx, ny, q[:, 0], q[:, 1])[0])
    if path is not None:
        assert W.verify(path[0], path[1])
        assert abs(path[0][-1] - 6.0) < 0.2 and abs(path[1][-1] - 10.0) < 0.2  # ends next to the goal (Q6)


def test_extend_step_with_dubins_edges(ctx, O, pp):
    """pp_rrt_extend_dubins = get_random_node + verify of the Dubins edge new -> nearest (src/rrt.rs:406-426)"""
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(6000, 4000, world=120.0)
    bounds, rings = pp.synth.circle_world(150, world=120.0, rmin=0.5, rmax=2.0)
    ctx.tree_upload(nx, ny, nyaw)
    ctx.obstacles_upload(bounds, rings)
    oidx, _ = O.nn_brute(nx, ny, qx, qy)
    wyaw = np.array([O.compute_yaw(qx[i], qy[i], nx[oidx[i]], ny[oidx[i]]) for i in range(qx.size)])
    W = O.OracleWorld(bounds, rings)
    want = W.verify_dubins_edges(qx, qy, wyaw, nx[oidx], ny[oidx], nyaw[oidx], 0.8, 0.1)
    for nnf, cf in [(NN_DEFAULT, DEFAULT), (NN_GRID, DEFAULT), (NN_SCAN, DEFAULT)]:
        idx, yaw, ok = ctx.rrt_extend_dubins(qx, qy, 0.8, 0.1, nn_flags=nnf, collide_flags=cf)
        assert np.array_equal(idx, oidx)
        assert np.abs(yaw - wyaw).max() < 1e-12
        check_dubins_verdicts(O, W, ok, qx, qy, wyaw, nx[oidx], ny[oidx], nyaw[oidx], 0.8, 0.1, want=want)
    assert 0 < want.sum() < want.size
    one = ctx.rrt_extend_dubins([qx[0]], [qy[0]], 0.8, 0.1)  # scalar use
    assert one[0][0] == oidx[0] and one[2][0] == want[0]

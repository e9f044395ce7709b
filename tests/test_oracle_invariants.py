"""CPU: pins of the oracle that do not depend on anybody's transcription of the reference's formulas.  The reference
ships no tests, so besides the golden vectors (a second, independent transliteration) the oracle is held to the
mathematics it claims to restate: every feasible word, integrated forward as a unicycle, must reach the goal pose; the
selected word must be the cheapest feasible one (first wins ties); and a sampled path must start at the start pose,
advance by at most one step in arc length and heading per sample, and stop within one step of the goal."""
import math

import numpy as np

MODES = ["LSL", "RSR", "LSR", "RSL", "RLR", "LRL"]  # ALL_PLANNERS order, src/dubins.rs:291


def _integrate(word, tpq, alpha):
    """unit-radius unicycle from (0, 0, alpha) through the three segments of the word"""
    x = y = 0.0
    th = alpha
    for m, l in zip(MODES[word], tpq):
        if m == "S":
            x += l * math.cos(th)
            y += l * math.sin(th)
        elif m == "L":
            x += math.sin(th + l) - math.sin(th)
            y += -math.cos(th + l) + math.cos(th)
            th += l
        else:
            x += -math.sin(th - l) + math.sin(th)
            y += math.cos(th - l) - math.cos(th)
            th -= l
    return x, y, th


def test_every_feasible_word_reaches_the_goal(O):
    """normalised problem of src/dubins.rs:333-338: start (0, 0, alpha), goal (d, 0, beta)"""
    rng = np.random.default_rng(3)
    feasible = [0] * 6
    for _ in range(4000):
        alpha, beta = rng.uniform(0, 2 * math.pi, 2)
        d = rng.uniform(0, 8) if rng.random() < 0.8 else rng.uniform(8, 120)
        for w in range(6):
            tpq = O.dubins_word(w, alpha, beta, d)
            if tpq is None:
                continue
            feasible[w] += 1
            assert all(0.0 <= v <= 2 * math.pi or MODES[w][i] == "S" for i, v in enumerate(tpq))  # arcs are mod2pi'd
            x, y, th = _integrate(w, tpq, alpha)
            err = max(abs(x - d), abs(y), abs(math.remainder(th - beta, 2 * math.pi)))
            assert err < 1e-9 * max(1.0, d), (MODES[w], alpha, beta, d, err)
    assert feasible[0] == feasible[1] == 4000 and min(feasible) > 1000  # LSL / RSR always exist


def test_selection_is_the_cheapest_feasible_word_first_wins(O):
    rng = np.random.default_rng(5)
    for _ in range(3000):
        span = 2.0 if rng.random() < 0.7 else 50.0
        sx, sy, ex, ey = rng.uniform(-span, span, 4)
        syaw, eyaw = rng.uniform(-math.pi, math.pi, 2)
        radius = float(rng.choice([0.5, 1.0, 2.5]))
        w, cost, tpq, _ = O.dubins_eval(sx, sy, syaw, ex, ey, eyaw, radius)
        # the same normalisation as the reference (src/dubins.rs:401-408, 333-338), from the oracle's own primitives
        dx, dy = ex - sx, ey - sy
        lex = math.cos(syaw) * dx + math.sin(syaw) * dy
        ley = -math.sin(syaw) * dx + math.cos(syaw) * dy
        d = math.hypot(lex, ley) / radius
        theta = O.mod2pi(math.atan2(ley, lex))
        alpha, beta = O.mod2pi(-theta), O.mod2pi((eyaw - syaw) - theta)
        best, best_w = math.inf, 0xFF
        for k in range(6):
            r = O.dubins_word(k, alpha, beta, d)
            if r is not None and abs(r[0]) + abs(r[1]) + abs(r[2]) < best:  # strict: the earliest word keeps a tie
                best, best_w = abs(r[0]) + abs(r[1]) + abs(r[2]), k
        assert w == best_w and (cost == best or abs(cost - best) <= 1e-12 * best)
        assert abs(tpq[0]) + abs(tpq[1]) + abs(tpq[2]) == cost  # (not sum(): Python 3.12 compensates float sums)


def test_sampled_paths_are_unit_speed_curves_from_start_to_goal(O):
    rng = np.random.default_rng(4)
    seen = 0
    for _ in range(1500):
        span = float(rng.choice([2.0, 50.0]))
        sx, sy, ex, ey = rng.uniform(-span, span, 4)
        syaw, eyaw = rng.uniform(-math.pi, math.pi, 2)
        radius, step = float(rng.choice([0.5, 1.0, 2.5])), float(rng.choice([0.05, 0.1, 0.3]))
        p = O.dubins_path(sx, sy, syaw, ex, ey, eyaw, radius, step)
        if p is None or p.x.size < 2:
            continue
        seen += 1
        assert p.x[0] == sx and p.y[0] == sy and abs(math.remainder(p.yaw[0] - syaw, 2 * math.pi)) < 1e-12
        chord = np.hypot(np.diff(p.x), np.diff(p.y))
        assert chord.max() <= step * radius * (1 + 1e-9)  # step is arc length in units of the turn radius
        dyaw = np.abs(np.remainder(np.diff(p.yaw) + math.pi, 2 * math.pi) - math.pi)
        assert dyaw.max() <= step * (1 + 1e-9)  # curvature never exceeds 1 / radius
        # the reference drops the final end point: the last sample is less than one step short of the goal
        assert math.hypot(p.x[-1] - ex, p.y[-1] - ey) <= step * radius * (1 + 1e-9)
        assert abs(math.remainder(p.yaw[-1] - eyaw, 2 * math.pi)) <= step * (1 + 1e-9)
        assert abs((p.x.size - 1) * step - p.cost) <= step * (1 + 1e-9)  # cost is the normalised length
    assert seen > 1400


def _word_modes(word):
    return {0: "LSL", 1: "RSR", 2: "LSR", 3: "RSL", 4: "RLR", 5: "LRL"}[word]


def _path_box(sx, sy, syaw, ex, ey, tpq, word, r):
    """Python restatement of pp_path_box (csrc/dubins_device.cuh): start, segment origins, end, parent point, and per arc
    the axis extremes of its turn circle that fall inside the swept angle"""
    xs, ys = [sx, ex], [sy, ey]
    x, y, th = sx, sy, syaw
    for mode, ln in zip(_word_modes(word), tpq):
        xs.append(x); ys.append(y)
        if mode == "S":
            x, y = x + ln * r * math.cos(th), y + ln * r * math.sin(th)
        else:
            left = mode == "L"
            cx, cy = (x - r * math.sin(th), y + r * math.cos(th)) if left else (x + r * math.sin(th), y - r * math.cos(th))
            phi0 = th - math.pi / 2 if left else th + math.pi / 2
            for k in range(4):  # axis k at k pi/2
                a = (k * math.pi / 2 - phi0) % (2 * math.pi) if left else (phi0 - k * math.pi / 2) % (2 * math.pi)
                if a <= ln + 1e-6 or a >= 2 * math.pi - 1e-6:
                    xs.append(cx + r * math.cos(k * math.pi / 2)); ys.append(cy + r * math.sin(k * math.pi / 2))
            th = th + ln if left else th - ln
            x, y = (cx + r * math.sin(th), cy - r * math.cos(th)) if left else (cx - r * math.sin(th), cy + r * math.cos(th))
        xs.append(x); ys.append(y)
    pad = 1e-9 * (abs(min(xs)) + abs(max(xs)) + abs(min(ys)) + abs(max(ys)) + r + 1.0)
    return min(xs) - pad, min(ys) - pad, max(xs) + pad, max(ys) + pad


def test_path_level_shortcut_geometry(O):
    """The two geometric claims behind the verify kernel's hierarchy, checked on the ORACLE's samples:
    for words with three positive lengths (i) every sample lies inside the path box, (ii) consecutive samples are at
    most one step of path length apart, so every point of a 32-point chunk lies in the chunk's end-point box grown by
    sqrt(S^2 - c^2) / 2 with S = 31.5 steps.  And a word with a zero-length segment violates both (why the GPU path
    excludes it)."""
    rng = np.random.default_rng(2024)
    checked = 0
    for _ in range(5000):
        span = rng.choice([1.5, 6.0, 40.0])
        r = float(rng.choice([0.5, 1.0, 3.0]))
        step = float(rng.choice([0.05, 0.1, 0.3]))
        sx, sy, ex, ey = (float(v) for v in rng.uniform(-span, span, 4))
        syaw, eyaw = (float(v) for v in rng.uniform(-math.pi, math.pi, 2))
        p = O.dubins_path(sx, sy, syaw, ex, ey, eyaw, r, step)
        w, cost, tpq, fl = O.dubins_eval(sx, sy, syaw, ex, ey, eyaw, r)
        if p is None or not all(v > 0.0 for v in tpq) or len(p.x) < 2:
            continue
        checked += 1
        x0, y0, x1, y1 = _path_box(sx, sy, syaw, ex, ey, tpq, w, r)
        px, py = np.asarray(p.x), np.asarray(p.y)
        assert px.min() >= x0 and px.max() <= x1 and py.min() >= y0 and py.max() <= y1, (sx, sy, syaw, ex, ey, eyaw, r, step)
        gaps = np.hypot(np.diff(px), np.diff(py))
        assert gaps.max() <= step * r * (1 + 1e-9), gaps.max() / (step * r)
        S = 31.5 * step * r
        for k0 in range(0, len(px) - 32, 31):  # chunks of 32 points, one point of overlap
            ax, ay, bx, by = px[k0], py[k0], px[k0 + 31], py[k0 + 31]
            h = 0.5 * math.sqrt(max(S * S - (bx - ax) ** 2 - (by - ay) ** 2, 0.0)) + 1e-9 * (abs(ax) + abs(ay) + 1.0)
            cx, cy = px[k0:k0 + 32], py[k0:k0 + 32]
            assert cx.min() >= min(ax, bx) - h and cx.max() <= max(ax, bx) + h
            assert cy.min() >= min(ay, by) - h and cy.max() <= max(ay, by) + h
    assert checked > 3500
    # the counter-example: goal straight ahead (t = q = 0): a sample BEHIND the start, outside the geometric path's box
    p = O.dubins_path(0.0, 0.0, 0.0, 9.9, 0.0, 0.0, 3.0, 0.2)
    assert min(p.x) < -0.5 and 0.0 in O.dubins_eval(0.0, 0.0, 0.0, 9.9, 0.0, 0.0, 3.0)[2]


def test_path_box_free_implies_oracle_free(O):
    """the implication the plan kernel's path-level test rests on, on the oracle's side: a Dubins edge (three positive
    lengths) whose path box lies inside the bounds and meets no ring's bounding box is free for the oracle's full
    Space::verify of the sampled line (src/rrt.rs:124-137, 291-321)"""
    rng = np.random.default_rng(99)
    world = 120.0
    cx, cy, rr = rng.uniform(0, world, 250), rng.uniform(0, world, 250), rng.uniform(0.5, 2.5, 250)
    rings = [O.create_circle(float(a), float(b), float(c)) for a, b, c in zip(cx, cy, rr)]
    bounds = (np.array([0.0, 0.0, world, world, 0.0]), np.array([0.0, world, world, 0.0, 0.0]))
    W = O.OracleWorld(bounds, rings)
    rb = np.array([(rx.min(), ry.min(), rx.max(), ry.max()) for rx, ry in rings])
    n = 4000
    sx, sy = rng.uniform(-2, world + 2, n), rng.uniform(-2, world + 2, n)
    ex, ey = sx + rng.uniform(-2.5, 2.5, n), sy + rng.uniform(-2.5, 2.5, n)
    syaw = np.arctan2(ey - sy, ex - sx)  # the new node aims at its parent (src/rrt.rs:267-271)
    eyaw = rng.uniform(-math.pi, math.pi, n)
    r, step = 0.8, 0.1
    want = W.verify_dubins_edges(sx, sy, syaw, ex, ey, eyaw, r, step)
    by_box = 0
    for i in range(n):
        w, cost, tpq, fl = O.dubins_eval(sx[i], sy[i], syaw[i], ex[i], ey[i], eyaw[i], r)
        if w == O.NONE or not all(v > 0.0 for v in tpq):
            continue
        x0, y0, x1, y1 = _path_box(sx[i], sy[i], syaw[i], ex[i], ey[i], tpq, w, r)
        inside = x0 > 1e-6 and y0 > 1e-6 and x1 < world - 1e-6 and y1 < world - 1e-6
        meets = np.any(~((rb[:, 2] < x0 - 1e-6) | (rb[:, 0] > x1 + 1e-6) | (rb[:, 3] < y0 - 1e-6) | (rb[:, 1] > y1 + 1e-6)))
        if inside and not meets:
            by_box += 1
            assert want[i] == 1, (i, sx[i], sy[i], ex[i], ey[i], eyaw[i])
    assert by_box > 0.3 * n and 0.3 < want.mean() < 0.95  # the shortcut answers a large share, and not everything is free

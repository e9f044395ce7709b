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

"""CPU: the oracle's geo 0.12.2 predicates (float formulas restated from SURVEY B.1) against EXACT rational geometry
(fractions.Fraction) on random general-position inputs.  For such inputs the float predicates must give the
mathematically true answer -- a wrong sign, a swapped operand or a mis-ordered comparison in the restatement shows up
here without any reference to geo's source.  Degenerate inputs (on-boundary points, collinear overlaps), where geo's
answers are conventions rather than geometry, are pinned separately in tests/test_oracle_geo.py."""
from fractions import Fraction as F

import numpy as np


def _star_ring(rng, cx, cy, n):
    """random star-shaped (possibly concave) closed ring"""
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    rad = rng.uniform(1.0, 4.0, n)
    x, y = cx + rad * np.cos(ang), cy + rad * np.sin(ang)
    return np.append(x, x[0]), np.append(y, y[0])


def _exact_inside(rx, ry, px, py):
    """crossing number with exact arithmetic; returns None when the point is on the ring"""
    px, py = F(px), F(py)
    inside = False
    for i in range(len(rx) - 1):
        x0, y0, x1, y1 = F(rx[i]), F(ry[i]), F(rx[i + 1]), F(ry[i + 1])
        cross = (x1 - x0) * (py - y0) - (y1 - y0) * (px - x0)
        if cross == 0 and min(x0, x1) <= px <= max(x0, x1) and min(y0, y1) <= py <= max(y0, y1):
            return None
        if (y0 > py) != (y1 > py):
            xi = x0 + (py - y0) * (x1 - x0) / (y1 - y0)
            if xi > px:
                inside = not inside
    return inside


def _orient(ax, ay, bx, by, cx, cy):
    v = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)
    return (v > 0) - (v < 0)


def _exact_segments_meet(a, b):
    """closed segments a = (x0, y0, x1, y1), b likewise; None when any three end points are collinear (degenerate)"""
    ax0, ay0, ax1, ay1 = map(F, a)
    bx0, by0, bx1, by1 = map(F, b)
    o = (_orient(ax0, ay0, ax1, ay1, bx0, by0), _orient(ax0, ay0, ax1, ay1, bx1, by1),
         _orient(bx0, by0, bx1, by1, ax0, ay0), _orient(bx0, by0, bx1, by1, ax1, ay1))
    if 0 in o:
        return None
    return o[0] != o[1] and o[2] != o[3]


def test_point_in_ring_matches_exact_geometry(O):
    rng = np.random.default_rng(21)
    seen = {0: 0, 1: 0}
    for _ in range(300):
        rx, ry = _star_ring(rng, rng.uniform(-50, 50), rng.uniform(-50, 50), int(rng.integers(3, 24)))
        for _ in range(20):
            px, py = rng.uniform(rx.min() - 1, rx.max() + 1), rng.uniform(ry.min() - 1, ry.max() + 1)
            truth = _exact_inside(rx, ry, px, py)
            if truth is None:
                continue
            got = O.point_position(rx, ry, float(px), float(py))  # 0 outside, 1 inside, 2 boundary
            assert got == int(truth), (rx.tolist(), ry.tolist(), px, py)
            assert not O.ring_has_point(rx, ry, float(px), float(py))  # geo's on-boundary test: the point is off the ring
            seen[int(truth)] += 1
    assert min(seen.values()) > 1000


def test_ring_vs_segment_matches_exact_geometry(O):
    rng = np.random.default_rng(22)
    hits = misses = 0
    for _ in range(150):
        rx, ry = _star_ring(rng, rng.uniform(-50, 50), rng.uniform(-50, 50), int(rng.integers(3, 24)))
        for _ in range(20):
            a = rng.uniform([rx.min() - 2, ry.min() - 2], [rx.max() + 2, ry.max() + 2])
            b = a + rng.uniform(-4, 4, 2)
            truth, degenerate = False, False
            for i in range(len(rx) - 1):
                m = _exact_segments_meet((rx[i], ry[i], rx[i + 1], ry[i + 1]), (a[0], a[1], b[0], b[1]))
                if m is None:
                    degenerate = True
                    break
                truth |= m
            if degenerate:
                continue
            assert O.lines_intersect(rx, ry, [a[0], b[0]], [a[1], b[1]]) == truth
            hits += truth
            misses += not truth
    assert hits > 400 and misses > 400


def test_verify_matches_exact_geometry(O):
    """Space::verify on 2-point lines = both ends strictly inside the bounds, no ring crossed, no end inside a ring"""
    rng = np.random.default_rng(23)
    free = blocked = 0
    for _ in range(24):
        rings = [_star_ring(rng, rng.uniform(5, 95), rng.uniform(5, 95), int(rng.integers(3, 16))) for _ in range(12)]
        bx, by = np.array([0.0, 0.0, 100.0, 100.0, 0.0]), np.array([0.0, 100.0, 100.0, 0.0, 0.0])
        W = O.OracleWorld((bx, by), rings)
        for _ in range(32):
            a = rng.uniform(-3, 103, 2)
            b = a + rng.uniform(-8, 8, 2)
            ok, degenerate = True, False
            for p in (a, b):
                ins = _exact_inside(bx, by, p[0], p[1])
                degenerate |= ins is None
                ok &= bool(ins)
            for rx, ry in rings:
                if degenerate:
                    break
                for p in (a, b):
                    ins = _exact_inside(rx, ry, p[0], p[1])
                    degenerate |= ins is None
                    ok &= not ins
                for i in range(len(rx) - 1):
                    m = _exact_segments_meet((rx[i], ry[i], rx[i + 1], ry[i + 1]), (a[0], a[1], b[0], b[1]))
                    if m is None:
                        degenerate = True
                        break
                    ok &= not m
            if degenerate:
                continue
            assert W.verify([a[0], b[0]], [a[1], b[1]]) == ok
            assert W.verify([a[0], b[0]], [a[1], b[1]], culled=True) == ok
            free += ok
            blocked += not ok
    assert free > 100 and blocked > 100


def test_circle_filter_classes_are_exact_statements(O):
    """the second cull of the culled loop (circle filter): class 0 must mean that the closed segment
    and the closed polygon have NO point in common, class 1 that both end points lie strictly inside the polygon --
    checked in exact rational arithmetic on random star-shaped (concave) rings and create_circle polygons, for segments
    of all lengths including ones that graze the outer / inner circle"""
    rng = np.random.default_rng(77)
    seen = {0: 0, 1: 0, 2: 0}
    for trial in range(60):
        if trial % 2:
            rx, ry = _star_ring(rng, rng.uniform(-50, 50), rng.uniform(-50, 50), int(rng.integers(3, 24)))
        else:
            rx, ry = O.create_circle(rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(0.6, 4.0))
        cx, cy = 0.5 * (rx.min() + rx.max()), 0.5 * (ry.min() + ry.max())
        rout = np.hypot(rx - cx, ry - cy).max()
        for _ in range(120):
            kind = rng.integers(0, 4)
            if kind == 0:    # anywhere around
                a = np.array([cx, cy]) + rng.uniform(-3 * rout, 3 * rout, 2)
                b = a + rng.normal(0, rout, 2)
            elif kind == 1:  # short, near the centre
                a = np.array([cx, cy]) + rng.normal(0, 0.3 * rout, 2)
                b = a + rng.normal(0, 0.1 * rout, 2)
            elif kind == 2:  # tangent-ish to the outer circle
                th = rng.uniform(0, 2 * np.pi)
                rad = rout * (1.0 + rng.choice([1e-9, 1e-6, 1e-3, 0.05]) * rng.choice([-1, 1]))
                mid = np.array([cx + rad * np.cos(th), cy + rad * np.sin(th)])
                tang = np.array([-np.sin(th), np.cos(th)]) * rng.uniform(0.1, 3.0) * rout
                a, b = mid - tang, mid + tang
            else:            # a single vertex (a == b), as the last point of a polyline
                a = np.array([cx, cy]) + rng.uniform(-1.5 * rout, 1.5 * rout, 2)
                b = a.copy()
            cls = O.circle_class(rx, ry, a[0], a[1], b[0], b[1])
            seen[cls] += 1
            if cls == 0:
                # no contact: neither an intersection with a ring segment nor an end point inside
                for i in range(len(rx) - 1):
                    meet = _exact_segments_touch((a[0], a[1], b[0], b[1]), (rx[i], ry[i], rx[i + 1], ry[i + 1]))
                    assert not meet, (trial, i)
                assert _exact_inside(rx, ry, a[0], a[1]) is False and _exact_inside(rx, ry, b[0], b[1]) is False
            elif cls == 1:
                assert _exact_inside(rx, ry, a[0], a[1]) is True and _exact_inside(rx, ry, b[0], b[1]) is True
    assert seen[0] > 500 and seen[1] > 200 and seen[2] > 500, seen


def _exact_segments_touch(a, b):
    """closed segments share at least one point (exact, degenerate cases included)"""
    ax0, ay0, ax1, ay1 = map(F, a)
    bx0, by0, bx1, by1 = map(F, b)

    def on(px, py, qx, qy, rx_, ry_):
        return min(px, qx) <= rx_ <= max(px, qx) and min(py, qy) <= ry_ <= max(py, qy)
    o1, o2 = _orient(ax0, ay0, ax1, ay1, bx0, by0), _orient(ax0, ay0, ax1, ay1, bx1, by1)
    o3, o4 = _orient(bx0, by0, bx1, by1, ax0, ay0), _orient(bx0, by0, bx1, by1, ax1, ay1)
    if (ax0, ay0) == (ax1, ay1):  # a is a single point
        return _orient(bx0, by0, bx1, by1, ax0, ay0) == 0 and on(bx0, by0, bx1, by1, ax0, ay0)
    if o1 != o2 and o3 != o4:
        return True
    return ((o1 == 0 and on(ax0, ay0, ax1, ay1, bx0, by0)) or (o2 == 0 and on(ax0, ay0, ax1, ay1, bx1, by1)) or
            (o3 == 0 and on(bx0, by0, bx1, by1, ax0, ay0)) or (o4 == 0 and on(bx0, by0, bx1, by1, ax1, ay1)))

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def O():
    """the CPU oracle (checker)"""
    o = graft.import_oracle()
    o.build()
    return o


@pytest.fixture(scope="session")
def pp():
    """the product package; importing it needs the built .so but no GPU"""
    return graft.import_package()


@pytest.fixture(scope="session")
def ctx(pp):
    if pp.device_count() == 0:
        pytest.fail("GPU test selected but no sm_100 device is visible: the CUDA path must run, there is no fallback")
    c = pp.Context(0)
    yield c
    c.close()


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))


def check_dubins_verdicts(O, W, ok, sx, sy, syaw, ex, ey, eyaw, radius, step, want=None, max_fragile=0.02):
    """GPU verdicts of Dubins edges against the oracle with NO numeric slack: every verdict that differs must be one
    the oracle itself classifies as undecidable at the 1e-9 sample tolerance (PPO_FLAG_*: near wrap / tie /
    feasibility of the word, knife-edge sample count, huge angles, or a verdict margin -- clearance of a free
    line, penetration depth of a blocked one -- below GRAZE_TOL).  Returns the number of classified differences.
    max_fragile bounds the share of fragile edges in a sub-sample so that the classification cannot be vacuous."""
    arrs = [np.ascontiguousarray(a, np.float64) for a in (sx, sy, syaw, ex, ey, eyaw)]
    ok = np.asarray(ok)
    culled = len(W.off) > 2001  # worlds of >= 2 000 rings: the oracle's culled loop (same verdicts, O(rings) per edge)
    if want is None:
        want = W.verify_dubins_edges(*arrs, radius, step, culled=culled)
    bad = np.nonzero(ok != want)[0]
    if bad.size:
        ok2, fl, mg = W.verify_dubins_edges_flags(*[a[bad] for a in arrs], radius, step, culled=culled)
        assert np.array_equal(ok2, want[bad])
        unclassified = bad[fl == 0]
        assert unclassified.size == 0, ("verdicts differ on edges the oracle calls robust", unclassified[:8],
                                        mg[fl == 0][:8])
    if max_fragile is not None and ok.size >= 200:
        sub = np.arange(0, ok.size, max(1, ok.size // 400))
        _, fl, _ = W.verify_dubins_edges_flags(*[a[sub] for a in arrs], radius, step, culled=culled)
        assert (fl != 0).mean() <= max_fragile, (fl != 0).mean()
    return int(bad.size)


def check_sample_counts(O, counts, sx, sy, syaw, ex, ey, eyaw, radius, step, ocounts=None):
    """sample counts against the oracle with no slack: a count may differ only where the oracle flags the path"""
    if ocounts is None:
        ocounts = O.dubins_count_batch(sx, sy, syaw, ex, ey, eyaw, radius, step)
    bad = np.nonzero(np.asarray(counts).astype(np.int64) != ocounts)[0]
    for i in bad:
        n, fl = O.dubins_path_flags(sx[i], sy[i], syaw[i], ex[i], ey[i], eyaw[i], radius, step)
        assert fl != 0, ("sample count differs on a path the oracle calls robust", int(i), int(counts[i]), n)
    # the classification cannot be vacuous: the oracle flags only a small share of a sub-sample
    sub = range(0, len(ocounts), max(1, len(ocounts) // 300))
    fragile = sum(O.dubins_path_flags(sx[i], sy[i], syaw[i], ex[i], ey[i], eyaw[i], radius, step)[1] != 0 for i in sub)
    assert fragile <= 0.02 * len(sub) + 1, fragile
    return int(bad.size)


def near_parallel_edges(rings, gaps=(1e-9, 1e-6, 1e-3, 0.5, 5.0, 40.0), lengths=(0.1, 1.0, 10.0)):
    """adversarial straight edges for the AABB culls (exactness (ii)): segments that start on the
    EXTENSION of a ring segment, beyond either end, and run (anti)parallel to it up to a rotation of 0 or +-2^-k rad,
    k = 20 .. 52.  For such pairs geo's denominator and both numerators are rounding noise."""
    import math
    ax, ay, bx, by = [], [], [], []
    for rx, ry in rings:
        for i in range(len(rx) - 1):
            x0, y0, x1, y1 = rx[i], ry[i], rx[i + 1], ry[i + 1]
            L = math.hypot(x1 - x0, y1 - y0)
            if L == 0.0:
                continue
            ux, uy = (x1 - x0) / L, (y1 - y0) / L
            for gap in gaps:
                for ln in lengths:
                    for eps in [0.0] + [s * 2.0 ** -k for k in range(20, 54, 3) for s in (1, -1)]:
                        c, s = math.cos(eps), math.sin(eps)
                        vx, vy = c * ux - s * uy, s * ux + c * uy
                        px, py = x1 + gap * ux, y1 + gap * uy
                        ax.append(px); ay.append(py); bx.append(px + ln * vx); by.append(py + ln * vy)
                        px, py = x0 - gap * ux, y0 - gap * uy
                        ax.append(px); ay.append(py); bx.append(px - ln * vx); by.append(py - ln * vy)
    return tuple(np.array(v) for v in (ax, ay, bx, by))


def exactly_free(W, ax, ay, bx, by):
    """EXACT rational geometry: True iff the closed segment a-b has no point in common with any ring boundary, both end
    points lie strictly inside the bounds ring and strictly outside every obstacle ring.  A `blocked` verdict of
    the float predicates on such a segment is rounding noise (near-parallel pairs), not geometry."""
    from fractions import Fraction as F

    def orient(px, py, qx, qy, rx, ry):
        v = (qx - px) * (ry - py) - (qy - py) * (rx - px)
        return (v > 0) - (v < 0)

    def on_seg(px, py, qx, qy, rx, ry):  # r on closed segment pq, given collinear
        return min(px, qx) <= rx <= max(px, qx) and min(py, qy) <= ry <= max(py, qy)

    def meet(a, b):
        o1, o2 = orient(*a, b[0], b[1]), orient(*a, b[2], b[3])
        o3, o4 = orient(*b, a[0], a[1]), orient(*b, a[2], a[3])
        if o1 != o2 and o3 != o4:
            return True
        return ((o1 == 0 and on_seg(*a, b[0], b[1])) or (o2 == 0 and on_seg(*a, b[2], b[3])) or
                (o3 == 0 and on_seg(*b, a[0], a[1])) or (o4 == 0 and on_seg(*b, a[2], a[3])))

    def inside(rx, ry, px, py):  # strict; on-boundary was excluded by meet()
        c = False
        for i in range(len(rx) - 1):
            x0, y0, x1, y1 = rx[i], ry[i], rx[i + 1], ry[i + 1]
            if (y0 > py) != (y1 > py) and x0 + (py - y0) * (x1 - x0) / (y1 - y0) > px:
                c = not c
        return c

    a = (F(ax), F(ay), F(bx), F(by))
    rings = [(W.bx, W.by)] + W.rings()
    for k, (rx, ry) in enumerate(rings):
        rx, ry = [F(v) for v in rx], [F(v) for v in ry]
        for i in range(len(rx) - 1):
            if meet(a, (rx[i], ry[i], rx[i + 1], ry[i + 1])):
                return False
        for px, py in ((a[0], a[1]), (a[2], a[3])):
            if inside(rx, ry, px, py) != (k == 0):  # inside the bounds, outside every obstacle
                return False
    return True

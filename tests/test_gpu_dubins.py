"""GPU parity: Dubins evaluate / words / sampling against the CPU oracle (through the C-ABI).
Tolerance (BASELINE.json north_star): 1e-9 relative on costs and samples; word selection identical
except pairs the oracle flags as near-wrap (Q3), near-tie (Q4) or near-infeasible."""
import math

import numpy as np
import pytest

from conftest import check_sample_counts, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9
R45 = 45.0 * (math.pi / 180.0)

KAT = [  # SURVEY.md Appendix C (restatement-derived)
    ("bench", (1, 1, R45, -3, -3, -R45, 1.0, 0.1), "LSL", 9.47540184001621, 95),
    ("conf1", (1, 1, R45, -3, -3, -R45, 0.5, 0.01), "LSL", 15.074463241942095, 1508),
    ("conf2", (-3, -3, -R45, 1, 1, R45, 0.5, 0.01), "LSR", 11.93317573386152, 1194),
    ("ccc", (0, 0, 0, 0.5, 0.5, math.pi, 1.0, 0.1), "RLR", 6.660418079530395, 67),
    ("straight", (0, 0, 0, 5, 0, 0, 1.0, 0.1), "LSL", 5.0, 53),
    ("same", (2, 3, 0.7, 2, 3, 0.7, 1.0, 0.1), "LSL", 0.0, 0),
    ("rrt", (10, 10, 2.356194490192345, 5, 15, 1.0, 0.8, 0.1), "LSR", 9.257602488965489, 93),
]


def _check_eval(O, cost, word, tpq, sx, sy, syaw, ex, ey, eyaw, radius):
    ocost, oword, otpq, oflags = O.dubins_eval_batch(sx, sy, syaw, ex, ey, eyaw, radius)
    clean = oflags == 0
    none = oword == O.NONE
    assert np.array_equal(word[clean], oword[clean])
    fin = clean & ~none
    assert rel_err(cost[fin], ocost[fin]).max(initial=0) < TOL
    assert rel_err(tpq[fin], otpq[fin]).max(initial=0) < TOL
    assert np.all(np.isinf(cost[clean & none]))
    # flagged pairs: the GPU answer must still be one of the words, evaluated consistently:
    # its (t,p,q) must match the oracle's value of THAT word up to a 2*pi wrap of t or q
    for i in np.nonzero(~clean)[0]:
        if word[i] == O.NONE:
            continue
        lex = math.cos(syaw[i]) * (ex[i] - sx[i]) + math.sin(syaw[i]) * (ey[i] - sy[i])
        ley = -math.sin(syaw[i]) * (ex[i] - sx[i]) + math.cos(syaw[i]) * (ey[i] - sy[i])
        d = math.hypot(lex, ley) / radius
        theta = O.mod2pi(math.atan2(ley, lex))
        r = O.dubins_word(int(word[i]), O.mod2pi(-theta), O.mod2pi(eyaw[i] - syaw[i] - theta), d)
        if r is None:
            continue  # near-infeasible word: GPU found it (just) feasible
        # p must agree; t and q may each be off by a 2*pi wrap (Q3), and when the straight / middle
        # segment has (near) zero length the split between t and q is rounding noise (atan2 of two
        # ~0 arguments): only t + q (mod 2*pi) is determined
        assert abs(tpq[i, 1] - r[1]) < 1e-6 or min(r[1], 2 * math.pi - r[1]) < 1e-6, (i, tpq[i], r)
        dsum = abs((tpq[i, 0] + tpq[i, 2]) - (r[0] + r[2])) % (2 * math.pi)
        assert min(dsum, 2 * math.pi - dsum) < 1e-6, (i, tpq[i], r)
    return int((~clean).sum())


def test_kat_paths(ctx, O):
    for name, c, w, cost, n in KAT:
        c = [float(v) for v in c]
        r = ctx.dubins_path(*c)
        assert r is not None, name
        px, py, pyaw, word, gcost = r
        p = O.dubins_path(*c)
        assert O.WORDS[word] == w and word == p.word, name
        assert abs(gcost - cost) <= TOL * max(1.0, cost), name
        assert len(px) == n == len(p.x), (name, len(px), n)
        if n:
            assert np.abs(px - p.x).max() < TOL * 10 and np.abs(py - p.y).max() < TOL * 10, name
            dyaw = np.abs(pyaw - p.yaw)
            assert np.minimum(dyaw, np.abs(dyaw - 2 * math.pi)).max() < TOL * 10, name
            assert px[0] == c[0] and py[0] == c[1]  # sample 0 is exactly the start pose


def test_mod2pi_bit_exact(ctx, O):
    rng = np.random.default_rng(7)
    x = np.concatenate([
        rng.uniform(-50, 50, 20000), rng.uniform(-1e4, 1e4, 2000),
        np.array([0.0, -0.0, 2 * math.pi, -2 * math.pi, 4 * math.pi, -1e-300, 1e-300, -1e-17, math.pi, -math.pi,
                  6.283185307179586, 6.283185307179585, 6.283185307179587, 12.566370614359172, 1e6, -1e6, 1e15]),
        2 * math.pi * np.arange(-40, 40) + rng.uniform(-1e-14, 1e-14, 80)])
    got = ctx.mod2pi(x)
    want = np.array([O.mod2pi(float(v)) for v in x])
    assert np.array_equal(got, want)
    got2 = ctx.mod2pi(x, pi_2_pi=True)
    want2 = np.array([O.pi_2_pi(float(v)) for v in x])
    assert np.array_equal(got2, want2)


def test_words_against_oracle(ctx, O):
    rng = np.random.default_rng(3)
    n = 4000
    alpha, beta = rng.uniform(0, 2 * math.pi, n), rng.uniform(0, 2 * math.pi, n)
    d = np.concatenate([rng.uniform(0, 6, n // 2), rng.uniform(0, 60, n - n // 2)])
    tpq, feas = ctx.dubins_words(alpha, beta, d)
    for i in range(n):
        for w in range(6):
            r = O.dubins_word(w, float(alpha[i]), float(beta[i]), float(d[i]))
            if (r is not None) != bool(feas[i, w]):
                # only legitimate right at the feasibility boundary: the oracle's margin of that test must be ~0
                assert abs(O.dubins_word_margin(w, float(alpha[i]), float(beta[i]), float(d[i]))) <= 1e-9, (i, w)
                continue
            if r is None:
                continue
            for k in range(3):
                diff = abs(tpq[i, w, k] - r[k])
                near_wrap = min(r[k], 2 * math.pi - r[k]) < 1e-9
                assert diff < 1e-9 * max(1.0, abs(r[k])) or (near_wrap and abs(diff - 2 * math.pi) < 1e-8), (i, w, k)


@pytest.mark.parametrize("dist,radius", [("mixed", 1.0), ("far", 1.0), ("mixed", 0.5), ("far", 2.5)])
def test_eval_random(ctx, O, pp, dist, radius):
    n = 200_000
    sx, sy, syaw, ex, ey, eyaw = pp.synth.dubins_pairs(n, dist)
    cost, word, tpq = ctx.dubins_eval(sx, sy, syaw, ex, ey, eyaw, radius=radius)
    flagged = _check_eval(O, cost, word, tpq, sx, sy, syaw, ex, ey, eyaw, radius)
    assert flagged < n * 1e-3


def test_eval_radius_array_and_no_tpq(ctx, O, pp):
    n = 50_000
    sx, sy, syaw, ex, ey, eyaw = pp.synth.dubins_pairs(n, seed=11)
    rad = pp.synth.uniform(12, 0, n, 0.3, 3.0)
    cost, word, tpq = ctx.dubins_eval(sx, sy, syaw, ex, ey, eyaw, radius_arr=rad)
    ocost, oword, _, oflags = O.dubins_eval_batch(sx, sy, syaw, ex, ey, eyaw, radius_arr=rad)
    clean = oflags == 0
    assert np.array_equal(word[clean], oword[clean])
    assert rel_err(cost[clean], ocost[clean]).max() < TOL
    cost2, word2, none = ctx.dubins_eval(sx, sy, syaw, ex, ey, eyaw, radius_arr=rad, want_tpq=False)
    assert none is None and np.array_equal(cost, cost2) and np.array_equal(word, word2)


def test_eval_start_yaw_routes(ctx, O, pp):
    """the frame change takes theta = atan2(dy, dx) - syaw for |syaw| < 64 and the reference's rotation beyond (and for
    coincident positions): yaws across the switch, far beyond it, and whole turns added to both yaws"""
    n = 60_000
    sx, sy, syaw, ex, ey, eyaw = pp.synth.dubins_pairs(n, seed=21)
    rng = np.random.default_rng(3)
    kind = rng.integers(0, 6, n)
    syaw = np.where(kind == 0, rng.uniform(-64.0, 64.0, n), syaw)
    syaw = np.where(kind == 1, rng.choice([63.999, -63.999, 64.0, -64.0, 64.001, 70.0], n), syaw)
    syaw = np.where(kind == 2, rng.uniform(-1e4, 1e4, n), syaw)
    turns = np.where(kind == 3, rng.integers(-9, 10, n) * 2.0 * math.pi, 0.0)
    syaw, eyaw = syaw + turns, eyaw + np.where(kind == 4, rng.uniform(-500.0, 500.0, n), turns)
    same = kind == 5  # coincident positions: only the yaws differ
    ex, ey = np.where(same, sx, ex), np.where(same, sy, ey)
    cost, word, tpq = ctx.dubins_eval(sx, sy, syaw, ex, ey, eyaw, radius=1.0)
    flagged = _check_eval(O, cost, word, tpq, sx, sy, syaw, ex, ey, eyaw, 1.0)
    assert flagged < n * 0.2  # coincident positions are flagged near-ties by construction
    # the plan kernel goes through the same frame function: same words, same lengths
    counts, plans = ctx.dubins_sample_count(sx[:2000], sy[:2000], syaw[:2000], ex[:2000], ey[:2000], eyaw[:2000], 1.0, 0.25)
    pw = plans.reshape(-1, 112)[:, 104]  # PP_DUBINS_PLAN_BYTES, word at byte 104
    assert np.array_equal(pw, word[:2000])


def test_eval_structured_axis_aligned(ctx, O):
    """the 1296 axis-aligned grid poses of SURVEY A.3 Q3: many sit on a mod2pi wrap"""
    vals = [-2.0, -1.0, 0.0, 1.0, 2.0, 3.0]
    yaws = [0.0, math.pi / 2, math.pi, -math.pi / 2, math.pi / 4, -3 * math.pi / 4]
    P = np.array([(x, y, a, b) for x in vals for y in vals for a in yaws for b in yaws])
    n = len(P)
    sx, sy = np.zeros(n), np.zeros(n)
    cost, word, tpq = ctx.dubins_eval(sx, sy, P[:, 2].copy(), P[:, 0].copy(), P[:, 1].copy(), P[:, 3].copy(), radius=1.0)
    flagged = _check_eval(O, cost, word, tpq, sx, sy, P[:, 2].copy(), P[:, 0].copy(), P[:, 1].copy(), P[:, 3].copy(), 1.0)
    assert flagged < n  # informational: how many sit on a wrap / tie
    # un-flagged structured cases must be exact matches of word; exact ties go to the earliest word (Q4)
    c, w, _ = ctx.dubins_eval([0.0], [0.0], [0.0], [5.0], [0.0], [0.0], radius=1.0)
    assert w[0] == 0 and c[0] == 5.0


def test_eval_edge_cases(ctx, O):
    nan, inf = float("nan"), float("inf")
    sx = np.array([0.0, nan, 0.0, 2.0, 0.0])
    sy = np.array([0.0, 0.0, 0.0, 3.0, 0.0])
    syaw = np.array([0.0, 0.0, nan, 0.7, 0.0])
    ex = np.array([0.0, 1.0, 1.0, 2.0, inf])
    ey = np.array([0.0, 1.0, 1.0, 3.0, 0.0])
    eyaw = np.array([0.0, 0.0, 0.0, 0.7, 0.0])
    cost, word, tpq = ctx.dubins_eval(sx, sy, syaw, ex, ey, eyaw, radius=1.0)
    ocost, oword, _, _ = O.dubins_eval_batch(sx, sy, syaw, ex, ey, eyaw, 1.0)
    assert np.array_equal(word, oword)  # NaN poses -> None (Q5); identical poses -> LSL cost 0 (Q7)
    assert word[1] == O.NONE and word[2] == O.NONE and word[0] == 0 and cost[0] == 0.0 and cost[3] == 0.0
    # empty batch and bad radius
    c, w, t = ctx.dubins_eval([], [], [], [], [], [], radius=1.0)
    assert c.size == 0 and w.size == 0
    with pytest.raises(Exception):
        ctx.dubins_eval([0.0], [0.0], [0.0], [1.0], [0.0], [0.0], radius=0.0)
    # the word formulas' domain is the reference's own call pattern (alpha, beta out of mod2pi): anything else is refused
    for a, b, dd in [(7.0, 0.0, 1.0), (0.0, -0.1, 1.0), (1.0, 1.0, -1.0), (1.0, 1.0, 2e5)]:
        with pytest.raises(Exception):
            ctx.dubins_words([a], [b], [dd])
    t6, f6 = ctx.dubins_words([nan], [1.0], [1.0])  # NaN passes the `p_squared < 0` / `|tmp| > 1` tests: Some(NaN) (Q5)
    assert np.isnan(t6).all()


@pytest.mark.parametrize("radius,step,dist", [(1.0, 0.1, "mixed"), (0.5, 0.01, "mixed"), (1.0, 0.05, "far"), (2.0, 0.3, "far")])
def test_sampling_random(ctx, O, pp, radius, step, dist):
    n = 3000 if dist == "mixed" else 600
    sx, sy, syaw, ex, ey, eyaw = pp.synth.dubins_pairs(n, dist, seed=21)
    out, offsets, counts = pp.dubins.batch_paths(sx, sy, syaw, ex, ey, eyaw, radius, step, ctx=ctx)
    ocounts = O.dubins_count_batch(sx, sy, syaw, ex, ey, eyaw, radius, step)
    _, _, _, oflags = O.dubins_eval_batch(sx, sy, syaw, ex, ey, eyaw, radius)
    clean = oflags == 0
    # step-boundary knife edges: a count may differ only where the oracle flags the path (PPO_FLAG_NEAR_COUNT, ...)
    check_sample_counts(O, counts, sx, sy, syaw, ex, ey, eyaw, radius, step, ocounts)
    assert int(counts.sum()) == out.shape[0]
    worst = 0.0
    for i in np.nonzero(clean)[0][:: max(1, n // 300)]:
        p = O.dubins_path(sx[i], sy[i], syaw[i], ex[i], ey[i], eyaw[i], radius, step)
        if counts[i] != len(p.x):
            continue
        seg = out[int(offsets[i]):int(offsets[i]) + int(counts[i])]
        scale = max(1.0, np.abs(p.x).max(initial=0), np.abs(p.y).max(initial=0))
        worst = max(worst, np.abs(seg[:, 0] - p.x).max(initial=0) / scale, np.abs(seg[:, 1] - p.y).max(initial=0) / scale)
        dyaw = np.abs(seg[:, 2] - p.yaw)
        worst = max(worst, np.minimum(dyaw, np.abs(dyaw - 2 * math.pi)).max(initial=0))
    assert worst < TOL


def test_sampling_from_origin_and_none(ctx, O):
    p = ctx.dubins_path(0.0, 0.0, 0.0, 3.0, 2.0, 1.0, 0.8, 0.1, from_origin=True)
    q = O.dubins_path(0.0, 0.0, 0.0, 3.0, 2.0, 1.0, 0.8, 0.1, from_origin=True)
    assert p is not None and len(p[0]) == len(q.x) and p[3] == q.word
    assert np.abs(p[0] - q.x).max() < TOL and np.abs(p[1] - q.y).max() < TOL and np.abs(p[2] - q.yaw).max() < TOL
    assert ctx.dubins_path(float("nan"), 0.0, 0.0, 3.0, 2.0, 1.0, 0.8, 0.1) is None  # reference: None


@pytest.mark.parametrize("radius,step,span", [(1.0, 0.1, 4.0), (0.5, 0.05, 2.0), (2.5, 0.2, 40.0)])
def test_from_origin_batch(ctx, O, pp, radius, step, span):
    """dubins_path_planning_from_origin (src/dubins.rs:326-399) as a batch: goals given in the start frame, NO
    rotation and NO translation of the samples, yaw left un-normalised -- counts, words and every sample against the
    oracle's own from_origin route, plus the scalar entry point on the same goals."""
    rng = np.random.default_rng(4242)
    n = 400
    ex, ey = rng.uniform(-span, span, n), rng.uniform(-span, span, n)
    eyaw = rng.uniform(-math.pi, math.pi, n)
    z = np.zeros(n)
    junk = rng.uniform(-9, 9, n)  # start poses are ignored when from_origin is set
    out, offsets, counts = pp.dubins.batch_paths(junk, junk, junk, ex, ey, eyaw, radius, step, from_origin=True, ctx=ctx)
    assert int(counts.sum()) == out.shape[0]
    worst, compared = 0.0, 0
    for i in range(n):
        q = O.dubins_path(0.0, 0.0, 0.0, ex[i], ey[i], eyaw[i], radius, step, from_origin=True)
        cnt, fl = O.dubins_path_flags(0.0, 0.0, 0.0, ex[i], ey[i], eyaw[i], radius, step, from_origin=True)
        if q is None:
            assert counts[i] == 0
            continue
        if counts[i] != len(q.x):
            assert fl != 0, ("from_origin count differs on a path the oracle calls robust", i, int(counts[i]), len(q.x))
            continue
        seg = out[int(offsets[i]):int(offsets[i]) + int(counts[i])]
        if fl == 0:
            scale = max(1.0, np.abs(q.x).max(initial=0), np.abs(q.y).max(initial=0))
            worst = max(worst, np.abs(seg[:, 0] - q.x).max(initial=0) / scale, np.abs(seg[:, 1] - q.y).max(initial=0) / scale,
                        np.abs(seg[:, 2] - q.yaw).max(initial=0))  # yaw is NOT wrapped on this route: plain difference
            compared += 1
        if i % 40 == 0:  # the scalar entry point walks the same plan
            p = ctx.dubins_path(0.0, 0.0, 0.0, ex[i], ey[i], eyaw[i], radius, step, from_origin=True)
            assert p is not None and len(p[0]) == counts[i] and p[3] == q.word
            assert np.array_equal(p[0], seg[:, 0]) and np.array_equal(p[1], seg[:, 1]) and np.array_equal(p[2], seg[:, 2])
    assert compared > 0.9 * n and worst < TOL
    # the same goals through the world-frame entry with a zero start pose: identical positions (rotation by 0,
    # translation by 0 are exact), yaw wrapped by pi_2_pi there and raw here
    out0, off0, cnt0 = pp.dubins.batch_paths(z, z, z, ex, ey, eyaw, radius, step, from_origin=False, ctx=ctx)
    same = cnt0 == counts
    assert same.mean() > 0.95
    for i in np.nonzero(same)[0][::20]:
        a = out[int(offsets[i]):int(offsets[i]) + int(counts[i])]
        b = out0[int(off0[i]):int(off0[i]) + int(cnt0[i])]
        assert np.abs(a[:, :2] - b[:, :2]).max(initial=0) < TOL


def test_drop_in_module_api(pp, ctx, O):
    """the mirror of `pathplanning::dubins` reads like the reference's own call sites (benches/all.rs:100-115)"""
    d = pp.dubins
    conf = d.DubinsConfig(sx=1.0, sy=1.0, syaw=R45, ex=-3.0, ey=-3.0, eyaw=-R45, turn_radius=1.0, step_size=0.1)
    px, py, pyaw, mode, cost = d.dubins_path_planning(conf)
    assert mode == d.LSL_MODE and len(px) == 95 and abs(cost - 9.47540184001621) < 1e-9
    assert d.mod2pi(-1e-300) == 2 * math.pi  # Q1
    t, p, q, mode = d.lsl(math.pi, math.pi / 2, 5.65685424949238)
    assert abs((t + p + q) - 9.47540184001621) < 1e-9
    assert d.rlr(math.pi, math.pi / 2, 5.65685424949238)[0] is None  # infeasible for the bench pose

"""GPU: BASELINE.json's full-size configurations, checked through size-independent properties (the oracle
only sees sub-samples): C3 = 2^24 pose pairs, C4 = 2^20 queries x 2^20 nodes x 10 k rings,
C5 slice = 2^19 Dubins edges x 100 k rings."""
import math

import numpy as np
import pytest

from conftest import check_dubins_verdicts, check_sample_counts, rel_err

pytestmark = pytest.mark.gpu


def test_c3_full_batch_properties(ctx, pp, O):
    n = 1 << 24
    sx, sy, syaw, ex, ey, eyaw = pp.synth.dubins_pairs(n)
    cost, word, _ = ctx.dubins_eval(sx, sy, syaw, ex, ey, eyaw, radius=1.0, want_tpq=False)
    # every pair has a feasible word; the histogram is the survey's mix (LSL .25 RSR .25 LSR .12 RSL .12 RLR .13 LRL .13)
    hist = np.bincount(word, minlength=256)
    assert hist[:6].sum() == n
    frac = hist[:6] / n
    assert np.allclose(frac, [0.247, 0.247, 0.122, 0.121, 0.131, 0.131], atol=0.004)
    # a Dubins path is never shorter than the straight line (cost is radius-normalised, radius = 1)
    dist = np.hypot(ex - sx, ey - sy)
    assert np.all(cost >= dist * (1.0 - 1e-12)) and np.all(np.isfinite(cost)) and cost.max() < dist.max() + 4 * math.pi + 1e-9
    # oracle on a strided sub-sample of the full batch
    sub = slice(0, n, 997)
    ocost, oword, _, oflags = O.dubins_eval_batch(sx[sub], sy[sub], syaw[sub], ex[sub], ey[sub], eyaw[sub], 1.0)
    clean = oflags == 0
    assert clean.mean() > 0.999
    assert np.array_equal(word[sub][clean], oword[clean]) and rel_err(cost[sub][clean], ocost[clean]).max() < 1e-9
    # rigid-motion invariance: rotating and translating both poses leaves the cost unchanged
    m = 1 << 20
    th, tx, ty = 0.7315, 123.25, -77.5
    c, s = math.cos(th), math.sin(th)
    rot = lambda x, y: (c * x - s * y + tx, s * x + c * y + ty)
    sx2, sy2 = rot(sx[:m], sy[:m])
    ex2, ey2 = rot(ex[:m], ey[:m])
    cost2, word2, _ = ctx.dubins_eval(sx2, sy2, syaw[:m] + th, ex2, ey2, eyaw[:m] + th, radius=1.0, want_tpq=False)
    same_word = word2 == word[:m]
    assert same_word.mean() > 0.9999  # a moved pose can flip a near-tie
    assert rel_err(cost2[same_word], cost[:m][same_word]).max() < 1e-9
    # scaling invariance: positions and radius scaled by k -> same normalised cost
    cost3, word3, _ = ctx.dubins_eval(3.0 * sx[:m], 3.0 * sy[:m], syaw[:m], 3.0 * ex[:m], 3.0 * ey[:m], eyaw[:m], radius=3.0,
                                      want_tpq=False)
    same = word3 == word[:m]
    assert same.mean() > 0.9999 and rel_err(cost3[same], cost[:m][same]).max() < 1e-9


def test_c4_full_extend_step_properties(ctx, pp, O):
    m = n_nodes = 1 << 20
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(m, n_nodes)
    bounds, rings = pp.synth.circle_world(10_000)
    ctx.tree_upload(nx, ny, nyaw)
    ctx.obstacles_upload(bounds, rings)
    idx, yaw, ok = ctx.rrt_extend(qx, qy, nn_flags=8, collide_flags=8)  # tiled brute-force scans (PP_NN_SCAN, PP_COLLIDE_SCAN)
    idx_g, yaw_g, ok_g = ctx.rrt_extend(qx, qy)  # default: device-built node grid, then obstacle grid (two launches per 2^18-query chunk, chunks pipelined over three streams)
    assert np.array_equal(idx, idx_g) and np.array_equal(ok, ok_g) and np.array_equal(yaw, yaw_g)
    idx_s, yaw_s, ok_s = ctx.rrt_extend(qx, qy, collide_flags=16)  # PP_COLLIDE_FUSED: binned queries, one fused launch
    assert np.array_equal(idx, idx_s) and np.array_equal(ok, ok_s) and np.array_equal(yaw, yaw_s)
    # the reported neighbour is at least as close as 32 random nodes, and as the oracle's on a sub-sample
    d2 = (nx[idx] - qx) ** 2 + (ny[idx] - qy) ** 2
    rng = np.random.default_rng(1)
    for _ in range(32):
        r = rng.integers(0, n_nodes, m)
        assert np.all(d2 <= (nx[r] - qx) ** 2 + (ny[r] - qy) ** 2)
    sub = np.arange(0, m, 4099)
    oidx, _ = O.nn_brute(nx, ny, qx[sub], qy[sub])
    assert np.array_equal(idx[sub], oidx)
    W = O.OracleWorld(bounds, rings)
    sub2 = np.arange(0, m, 257)
    assert np.array_equal(ok[sub2], W.verify_segments(qx[sub2], qy[sub2], nx[idx[sub2]], ny[idx[sub2]], culled=True))
    assert 0.8 < ok.mean() < 0.9
    # order independence: a permuted batch gives the permuted answers (no cross-talk between queries)
    perm = rng.permutation(m)
    idx_p, _, ok_p = ctx.rrt_extend(qx[perm], qy[perm])
    assert np.array_equal(idx_p, idx[perm]) and np.array_equal(ok_p, ok[perm])


def test_c5_slice_properties(ctx, pp, O):
    e = 1 << 19
    bounds, rings = pp.synth.circle_world(100_000, rmin=0.5, rmax=1.5)
    ctx.obstacles_upload(bounds, rings)
    sx, sy, syaw, ex, ey, eyaw = pp.synth.dubins_edges(e)
    ok = ctx.collide_dubins(sx, sy, syaw, ex, ey, eyaw, 1.0, 0.05)
    W = O.OracleWorld(bounds, rings)
    sub = np.arange(0, e, 1531)
    want = W.verify_dubins_edges(sx[sub], sy[sub], syaw[sub], ex[sub], ey[sub], eyaw[sub], 1.0, 0.05, culled=True)
    check_dubins_verdicts(O, W, ok[sub], sx[sub], sy[sub], syaw[sub], ex[sub], ey[sub], eyaw[sub], 1.0, 0.05, want=want)
    # free edges are rare in this world (31 % of the area is covered); an edge outside the bounds is never free
    assert 0.002 < ok.mean() < 0.05
    outside = (ex < 0) | (ex > 1000) | (ey < 0) | (ey > 1000)
    assert outside.any() and not ok[outside].any()
    # order independence
    perm = np.random.default_rng(2).permutation(e)[: 1 << 17]
    okp = ctx.collide_dubins(sx[perm], sy[perm], syaw[perm], ex[perm], ey[perm], eyaw[perm], 1.0, 0.05)
    assert np.array_equal(okp, ok[perm])
    # sample counts of the same edges: count + fill agree with the plan the verify kernel walks
    counts, plan = ctx.dubins_sample_count(sx[:4096], sy[:4096], syaw[:4096], ex[:4096], ey[:4096], eyaw[:4096], 1.0, 0.05)
    ocnt = O.dubins_count_batch(sx[:4096], sy[:4096], syaw[:4096], ex[:4096], ey[:4096], eyaw[:4096], 1.0, 0.05)
    check_sample_counts(O, counts, sx[:4096], sy[:4096], syaw[:4096], ex[:4096], ey[:4096], eyaw[:4096], 1.0, 0.05, ocnt)
    assert 600 < counts.mean() < 1000  # ~785 samples per edge

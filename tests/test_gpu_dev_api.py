"""GPU: the `_dev` twins (device pointers, enqueue-only) give the same bits as the host-pointer calls, the
context survives concurrent callers (the reference calls its RRT methods from 4 rayon workers,
src/rrt.rs:600-609) and repeated create / destroy."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_dev_entry_points_match_host_calls(ctx, pp, O):
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    n = 100_000
    host = pp.synth.dubins_pairs(n, "mixed", seed=77)
    d_in = [torch.from_numpy(a).to(dev) for a in host]
    cost = torch.empty(n, dtype=torch.float64, device=dev)
    word = torch.empty(n, dtype=torch.uint8, device=dev)
    tpq = torch.empty(n * 3, dtype=torch.float64, device=dev)
    ctx.dubins_eval_dev(n, *d_in, 1.0, cost, word, tpq)
    ctx.sync()
    hc, hw, ht = ctx.dubins_eval(*host, radius=1.0)
    assert np.array_equal(cost.cpu().numpy(), hc) and np.array_equal(word.cpu().numpy(), hw)
    assert np.array_equal(tpq.cpu().numpy().reshape(n, 3), ht)

    # count -> device prefix sum -> fill, against the host-pointer path
    m = 2000
    counts = torch.empty(m, dtype=torch.int32, device=dev)
    plan = torch.empty(m * pp._ffi.PLAN_BYTES, dtype=torch.uint8, device=dev)
    offsets = torch.empty(m, dtype=torch.int64, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    sub = [t[:m].contiguous() for t in d_in]
    ctx.dubins_sample_count_dev(m, *sub, 1.0, 0.1, counts, plan)
    ctx.exclusive_scan_u32_dev(m, counts, offsets, total)
    ctx.sync()
    tot = int(total.item())
    out = torch.empty(tot * 3, dtype=torch.float64, device=dev)
    ctx.dubins_sample_fill_dev(m, plan, offsets, tot, out)
    ctx.sync()
    hout, hoff, hcnt = pp.dubins.batch_paths(*[a[:m] for a in host], 1.0, 0.1, ctx=ctx)
    assert np.array_equal(counts.cpu().numpy().astype(np.uint32), hcnt)
    assert np.array_equal(offsets.cpu().numpy().astype(np.uint64), hoff) and tot == hout.shape[0]
    assert np.array_equal(out.cpu().numpy().reshape(tot, 3), hout)

    # tree / NN / verify / extend on device pointers
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(20_000, 30_000, world=300.0)
    bounds, rings = pp.synth.circle_world(800, world=300.0)
    ctx.obstacles_upload(bounds, rings)
    tree = torch.from_numpy(np.stack([nx, ny, nyaw])).to(dev)
    ctx.tree_upload_dev(nx.size, tree[0], tree[1], tree[2])
    dqx, dqy = torch.from_numpy(qx).to(dev), torch.from_numpy(qy).to(dev)
    idx = torch.empty(qx.size, dtype=torch.int32, device=dev)
    d2 = torch.empty(qx.size, dtype=torch.float64, device=dev)
    yaw = torch.empty(qx.size, dtype=torch.float64, device=dev)
    ok = torch.empty(qx.size, dtype=torch.uint8, device=dev)
    oidx, od2 = O.nn_brute(nx, ny, qx, qy)
    for flags in (0, 1, 2, 4, 8):
        idx.zero_()
        ctx.nn_dev(qx.size, dqx, dqy, idx, d2, flags=flags)
        ctx.sync()
        assert np.array_equal(idx.cpu().numpy().astype(np.uint32), oidx) and np.array_equal(d2.cpu().numpy(), od2)
    ctx.rrt_extend_dev(qx.size, dqx, dqy, idx, yaw, ok)
    ctx.sync()
    W = O.OracleWorld(bounds, rings)
    assert np.array_equal(ok.cpu().numpy(), W.verify_segments(qx, qy, nx[oidx], ny[oidx]))
    bxy = torch.from_numpy(np.stack([nx[oidx], ny[oidx]])).to(dev)
    for flags in (0, 2, 4, 8):
        ok.zero_()
        ctx.collide_segments_dev(qx.size, dqx, dqy, bxy[0], bxy[1], ok, flags=flags)
        ctx.sync()
        assert np.array_equal(ok.cpu().numpy(), W.verify_segments(qx, qy, nx[oidx], ny[oidx])), flags


def test_caller_owned_stream_and_kernel_timing(ctx, pp):
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    n = 1 << 18
    d_in = [torch.from_numpy(a).to(dev) for a in pp.synth.dubins_pairs(n)]
    cost = torch.empty(n, dtype=torch.float64, device=dev)
    word = torch.empty(n, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream(device=dev)
    ctx.set_stream(s.cuda_stream)
    try:
        ctx.timing_enable(True)
        ctx.timing_reset()
        l0 = ctx.launch_count
        with torch.cuda.stream(s):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ctx.dubins_eval_dev(n, *d_in, 1.0, cost, word)
            e1.record()
        s.synchronize()
        ms, launches = ctx.timing_get("dubins_eval")
        assert launches == 5 and ctx.launch_count - l0 == 5
        assert 0 < ms <= e0.elapsed_time(e1) * 1.05  # the events of the caller's stream bracket our kernels
    finally:
        ctx.timing_enable(False)
        ctx.set_stream(None)


def test_concurrent_callers_and_context_lifecycle(ctx, pp, O):
    qx, qy, nx, ny, nyaw = pp.synth.extend_inputs(4000, 5000, world=100.0)
    ctx.tree_upload(nx, ny, nyaw)
    want_idx = O.nn_brute(nx, ny, qx, qy)[0]
    pairs = pp.synth.dubins_pairs(20_000, seed=5)
    want_cost = ctx.dubins_eval(*pairs, radius=1.0)[0]
    errs = []

    def worker(k):
        try:
            for _ in range(6):
                if k % 2:
                    assert np.array_equal(ctx.nn(qx, qy, want_d2=False), want_idx)
                else:
                    assert np.array_equal(ctx.dubins_eval(*pairs, radius=1.0)[0], want_cost)
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    for _ in range(10):  # create / use / destroy repeatedly (device mirrors and tables are per context)
        c = pp.Context(0)
        c.tree_upload(nx[:100], ny[:100])
        assert np.array_equal(c.nn(qx[:70], qy[:70], want_d2=False), O.nn_brute(nx[:100], ny[:100], qx[:70], qy[:70])[0])
        assert np.array_equal(c.dubins_eval(*[a[:50] for a in pairs], radius=1.0)[0], want_cost[:50])
        c.close()


def test_error_paths(ctx, pp):
    c = pp.Context(0)
    with pytest.raises(pp.PathPlanningError):  # obstacles not uploaded
        c.collide_segments([0.0], [0.0], [1.0], [1.0])
    with pytest.raises(pp.PathPlanningError):  # empty tree
        c.obstacles_upload((np.array([0.0, 0, 1, 1, 0]), np.array([0.0, 1, 1, 0, 0])), [])
        c.rrt_extend([0.5], [0.5])
    with pytest.raises(pp.PathPlanningError):
        c.dubins_sample_count([0.0], [0.0], [0.0], [1.0], [1.0], [0.0], 1.0, -0.1)  # step must be positive
    with pytest.raises(pp.PathPlanningError):
        pp.Context(99)  # no such device
    c.close()


def test_pageable_batches_take_the_staged_path_and_give_the_same_bits(ctx, pp):
    """pp_dubins_eval on ordinary (pageable) arrays of >= 2 chunks goes through the pinned ring and the copy threads
    (csrc/pp_stage.hpp); pinned arrays take the direct path.  Same bits either way, with a ragged last chunk, with
    per-pair radii and t, p, q requested, and on the slot-reuse boundaries (exactly 3 and 4 chunks)."""
    chunk = 1 << 20
    for n, with_radii, want_tpq in ((3 * chunk + 12_345, True, True), (3 * chunk, False, False), (4 * chunk, False, True),
                                    (2 * chunk, False, False)):
        host = pp.synth.dubins_pairs(n, "mixed", seed=4242 + n % 7)
        radii = (0.5 + (np.arange(n) % 5) * 0.25) if with_radii else None
        pins = [pp.PinnedArray(n, np.float64) for _ in range(7 if with_radii else 6)]
        for p, a in zip(pins, list(host) + ([radii] if with_radii else [])):
            p.array[:] = a
        pc, pw = pp.PinnedArray(n, np.float64), pp.PinnedArray(n, np.uint8)
        pt = pp.PinnedArray(3 * n, np.float64) if want_tpq else None
        out_pin = (pc.array, pw.array, pt.array.reshape(n, 3) if want_tpq else None)
        ctx.dubins_eval(*[p.array for p in pins[:6]], radius=1.0, radius_arr=pins[6].array if with_radii else None,
                        want_tpq=want_tpq, out=out_pin)
        cost, word, tpq = ctx.dubins_eval(*host, radius=1.0, radius_arr=radii, want_tpq=want_tpq)  # pageable numpy
        assert np.array_equal(cost, pc.array, equal_nan=True) and np.array_equal(word, pw.array)
        if want_tpq:
            assert np.array_equal(tpq, out_pin[2], equal_nan=True)
        assert np.bincount(word, minlength=6)[:6].min() > n // 20  # every word occurs: nothing was left unwritten
        del pins, pc, pw, pt, out_pin

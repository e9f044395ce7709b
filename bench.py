#!/usr/bin/env python
"""bench.py -- headline measurement of the hot path (BASELINE.json: "Dubins pairs/s; RRT extend steps/s").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--skip-secondary]

One JSON line on stdout (rank 0).  Primary metric: batched Dubins pose-pair evaluations per second on
config C3 (2^24 pairs per GPU, radius 1.0, shortest-word selection + length only).  The same line carries
the RRT extend-step (C4) and Dubins-edge verify (C5 slice) numbers under "workloads".
N > 1 is launched by torch.distributed.run, one rank per GPU; the path shards with no data-path collective
(weak scaling: every GPU takes its own 2^24-pair slice); the tree / obstacle buffers of the secondary
workloads are replicated with one NCCL broadcast outside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

N_PAIRS = 1 << 24          # C3, per GPU
W_INSTR_PER_PAIR = 1100.0  # fixed yard-stick of SURVEY.md Appendix D (FP64-pipe thread-instructions per pair)
BYTES_PER_PAIR = 57.0      # 48 B read + 8 B cost + 1 B word
NCU_FP64_INSTR_PER_PAIR = 490.0  # DFMA + DADD + DMUL + DSETP executed per pair (ncu source page, same capture)
NCU_DRAM_BYTES_PER_LAUNCH = 949.3e6  # measured once per kernel change by ncu (see profiles/r01_summary.md)
NCU_NN_GRID_L2_BYTES = 25431908 * 32.0  # lts__t_sectors.sum x 32 B, pp_nn_grid_kernel on 2^20 queries / 2^20 nodes
FP64_PEAK_NOMINAL = 148 * 64 * 1.965e9  # lanes * clock: used only if the live DFMA measurement fails
C4_M, C4_NODES, C4_RINGS = 1 << 20, 1 << 20, 10_000
C5_EDGES, C5_RINGS = 1 << 19, 100_000  # the per-GPU slice of config 5 (2^22 edges over 8 GPUs)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None
        self.lines = []
        if self.p:
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:  # region shorter than the sampling period: take the nearest samples
            for ts, ln in self.lines[-3:]:
                try:
                    f = [x.strip() for x in ln.split(",")]
                    sm.append(float(f[0]))
                    mx = float(f[1])
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """the reference's CPU implementation of the path = the C restatement (oracle port; no rustc here),
    all host threads, bounded sample per step"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    O = graft.import_oracle()
    O.build()
    pp_synth = graft.import_package().synth
    threads = O.max_threads()
    n = 1 << 21
    sx, sy, syaw, ex, ey, eyaw = pp_synth.dubins_pairs(n)
    for _ in range(max(args.warmup, 1)):
        O.dubins_eval_batch(sx[: n // 8], sy[: n // 8], syaw[: n // 8], ex[: n // 8], ey[: n // 8], eyaw[: n // 8], 1.0,
                            want_flags=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.dubins_eval_batch(sx, sy, syaw, ex, ey, eyaw, 1.0, want_flags=False)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"{args.steps} steps x 2^21 pairs of the C3 'mixed' distribution (seed 0xD0B10003), OpenMP static partition"
    line = {
        "impl": "reference", "metric": "dubins_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "c3_batched_dubins: 2^24 random pose pairs per GPU, radius 1.0, shortest-word selection + "
                               "length only (reference arm: bounded 2^21-pair sample per step on the host cores)"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement of src/dubins.rs (oracle/pp_oracle.c), not rustc output: no Rust toolchain in the image",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ own arm
def bind_to_gpu_numa(local):
    """pin this rank's host threads (hence its first-touch pinned buffers) to the NUMA node its GPU hangs off:
    with 8 ranks streaming 50 GB/s each, remote-socket staging halves the end-to-end rate"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml pads the domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # pragma: no cover
        return {"error": str(e)[:80]}
    return None


def time_steps(torch, dist, fn, steps, warmup, world):
    """W warm-up steps, then K steps bracketed by barrier + synchronize, CUDA events on the launching stream;
    returns max-over-ranks milliseconds for the K steps"""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), t0, t1


def run_own(args):
    import torch
    import torch.distributed as dist
    pp = graft.import_package()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = pp.Context(local)
    # a dedicated non-default stream shared by torch (events, copies) and the library's launches, so that
    # torch.cuda.Event brackets exactly our kernels (the legacy default stream's handle 0 means "own stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    peaks = measured_peaks()
    hbm_peak = (peaks or {}).get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"

    # ---- FP64 pipe peak, measured live
    try:
        fp64_peak, _ = ctx.measure_fp64_peak(4096)
        fp64_src = "measured live: pp_measure_fp64_peak DFMA micro-benchmark"
    except Exception as e:  # pragma: no cover
        fp64_peak, fp64_src = FP64_PEAK_NOMINAL, f"nominal 148x64x1.965GHz ({e})"

    # ---- C3 inputs: this rank's contiguous 2^24-pair slice of the global batch
    n = N_PAIRS
    t_gen = time.time()
    host = pp.synth.dubins_pairs(n, "mixed", first=rank * n)
    d_in = [torch.from_numpy(a).to(dev) for a in host]
    d_cost = torch.empty(n, dtype=torch.float64, device=dev)
    d_word = torch.empty(n, dtype=torch.uint8, device=dev)
    log(f"[rank {rank}] inputs ready in {time.time() - t_gen:.1f}s")

    def step():
        ctx.dubins_eval_dev(n, *d_in, 1.0, d_cost, d_word)

    sampler = ClockSampler(local) if rank == 0 else None
    ctx.timing_enable(True)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    # ~1 s of identical untimed steps directly before the timed region: the 100 ms nvidia-smi samples then
    # describe the clocks under this load even when K steps last only a few milliseconds
    t_load = time.time()
    while time.time() - t_load < (0.0 if args.profile else 1.0):
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    ctx.timing_reset()
    l0 = ctx.launch_count
    ms, t0, t1 = time_steps(torch, dist, step, args.steps, 0, world)
    launches = ctx.launch_count - l0
    k_ms, k_n = ctx.timing_get("dubins_eval")
    ctx.timing_enable(False)
    clocks = sampler.stop(t_load + 0.3, t1) if sampler else None
    if clocks is not None:
        clocks["window"] = "~1 s of identical untimed steps immediately before the timed region + the timed region"
    value = world * n * args.steps / (ms * 1e-3)
    k_avg_ms = k_ms / max(k_n, 1)
    pairs_per_s_kernel = n / (k_avg_ms * 1e-3)

    # sanity of what was computed (never a fallback): word histogram on rank 0
    hist = torch.bincount(d_word.to(torch.int64), minlength=256)[:6].tolist()

    # ---- e2e through the C-ABI with pinned HOST buffers (H2D + kernel + D2H inside the timed region)
    e2e_steps = max(1, min(args.steps, 3))
    pins = [pp.PinnedArray(n, np.float64) for _ in range(6)]
    for p, a in zip(pins, host):
        p.array[:] = a
    pcost, pword = pp.PinnedArray(n, np.float64), pp.PinnedArray(n, np.uint8)
    ctx.dubins_eval(*[p.array for p in pins], radius=1.0, want_tpq=False, out=(pcost.array, pword.array, None))  # warm-up
    if world > 1:
        dist.barrier()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.dubins_eval(*[p.array for p in pins], radius=1.0, want_tpq=False, out=(pcost.array, pword.array, None))
    e2e_s = time.perf_counter() - te
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(e2e_t.item())
    same = bool(np.array_equal(pcost.array, d_cost.cpu().numpy()) and np.array_equal(pword.array, d_word.cpu().numpy()))
    e2e_launches = 16 * (e2e_steps + 1)
    del pins, host

    workloads = {}
    if not args.skip_secondary:
        # C3's secondary "far" distribution (positions U[-50,50)^2, CSC words ~99.9 %): same kernel, same
        # buffers, other data -- the evaluation is branch-free, so the time should not depend on the word mix
        far = pp.synth.dubins_pairs(n, "far", first=rank * n)
        for d, a in zip(d_in, far):
            d.copy_(torch.from_numpy(a))
        del far
        for _ in range(3):
            step()
        far_steps = max(1, min(args.steps, 10))
        l0 = ctx.launch_count
        far_ms, _, _ = time_steps(torch, dist, step, far_steps, 0, world)
        far_hist = torch.bincount(d_word.to(torch.int64), minlength=256)[:6].tolist()
        workloads["dubins_far"] = {
            "metric": "dubins_pairs_per_s", "value": world * n * far_steps / (far_ms * 1e-3), "unit": "pairs/s",
            "ms_per_step": far_ms / far_steps, "steps": far_steps, "gpu_launches": ctx.launch_count - l0,
            "config": {"workload": "c3 'far': 2^24 pose pairs per GPU, positions U[-50,50)^2, yaws U[-pi,pi), radius 1.0",
                       "word_hist_rank0": far_hist}}
        workloads.update(secondary(args, torch, dist, pp, ctx, dev, rank, world, fp64_peak, hbm_peak))

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu_baseline = cpu_baseline_leg(pp)

    if rank == 0:
        achieved = pairs_per_s_kernel * W_INSTR_PER_PAIR / 1e9
        line = {
            "metric": "dubins_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "c3_batched_dubins: 2^24 random pose pairs per GPU ('mixed' U[-2,2)^2 positions, "
                                   "U[-pi,pi) yaws, seed 0xD0B10003), radius 1.0, shortest-word selection + length only",
                       "pairs_per_gpu": n, "l2": "inputs (768 MiB) larger than L2, no flush needed",
                       "sharding": "contiguous slice per rank, no data-path collective", "word_hist_rank0": hist},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": 48 * n, "d2h_bytes_per_step": 9 * n,
                    "steps": e2e_steps, "matches_device_run": same,
                    "how": "pp_dubins_eval on pinned host buffers: 16 chunks over 3 streams, copies inside the timed region",
                    "numa_binding_rank0": numa},
            "gpu_launches": launches + e2e_launches + sum(w.get("gpu_launches", 0) for w in workloads.values()),
            "gpu_launches_primary_timed_region": launches,
            "clocks": clocks,
            "roofline": {
                "kernel": "pp_dubins_eval_kernel", "bound": "fp64", "achieved": achieved, "peak": fp64_peak / 1e9,
                "unit": "Ginstr/s", "frac": achieved * 1e9 / fp64_peak,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "traffic_source": "ncu --set full, dram__bytes_read.sum + "
                "dram__bytes_write.sum per launch of 2^24 pairs (profiles/r01_dubins_eval_final4_raw.csv); algorithmic 956.3e6",
                "fp64_pipe_busy_ncu": 0.728, "fp64_instr_per_pair_ncu": NCU_FP64_INSTR_PER_PAIR,
                # `frac` follows the contract (SURVEY 8d yard-stick W = 1100 per pair) and exceeds 1 because the
                # kernel executes only ~490 FP64-pipe instructions per pair (ncu source page); with the executed count the same
                # timing gives the pipe utilisation ncu reports
                "frac_of_executed_fp64_work": pairs_per_s_kernel * NCU_FP64_INSTR_PER_PAIR / fp64_peak,
                "per_unit": f"W = {W_INSTR_PER_PAIR:.0f} FP64-pipe thread-instructions per pair (fixed yard-stick, SURVEY App. D)",
                "peak_source": fp64_src, "kernel_ms_avg": k_avg_ms, "kernel_launches_timed": k_n,
                "hbm_view": {"bound": "hbm", "achieved": pairs_per_s_kernel * BYTES_PER_PAIR / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": pairs_per_s_kernel * BYTES_PER_PAIR / 1e9 / hbm_peak,
                             "per_unit": "57 B per pair (48 read + 9 written)", "peak_source": hbm_src},
            },
            "cpu_baseline": cpu_baseline,
            "workloads": workloads,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def cpu_baseline_leg(pp):
    """oracle port timed on the host cores, bounded sample of the same workload (rank 0, N = 1 only)"""
    O = graft.import_oracle()
    O.build()
    threads = O.max_threads()
    n1 = 1 << 20
    a = pp.synth.dubins_pairs(n1)
    O.dubins_eval_batch(*[v[: 1 << 16] for v in a], 1.0, want_flags=False, nthreads=1)
    t = time.perf_counter()
    O.dubins_eval_batch(*a, 1.0, want_flags=False, nthreads=1)
    single = n1 / (time.perf_counter() - t)
    nall = 1 << 22
    b = pp.synth.dubins_pairs(nall)
    O.dubins_eval_batch(*[v[: 1 << 18] for v in b], 1.0, want_flags=False)
    best = 0.0
    for _ in range(3):
        t = time.perf_counter()
        O.dubins_eval_batch(*b, 1.0, want_flags=False)
        best = max(best, nall / (time.perf_counter() - t))
    # the second metric on the CPU: exact grid NN (the honest comparator for the reference's R-tree) + the
    # culled straight-edge verify (per edge: every ring's box, then geo's predicates -- the reference has no broad phase)
    mq = 1 << 16
    qx, qy, nx, ny, _ = pp.synth.extend_inputs(mq, C4_NODES)
    bounds, rings = pp.synth.circle_world(C4_RINGS)
    W = O.OracleWorld(bounds, rings)
    t = time.perf_counter()
    oidx, _ = O.nn_grid(nx, ny, qx, qy)
    t_nn = time.perf_counter() - t
    W.verify_segments(qx, qy, nx[oidx], ny[oidx], culled=True)
    t_ext = time.perf_counter() - t
    m1 = 1 << 13  # the same step on one thread
    t = time.perf_counter()
    i1, _ = O.nn_grid(nx, ny, qx[:m1], qy[:m1], nthreads=1)
    W.verify_segments(qx[:m1], qy[:m1], nx[i1], ny[i1], culled=True, nthreads=1)
    t_single = time.perf_counter() - t
    mb = 1 << 11  # exact brute force, what the tiled GPU scans do: 2^11 x 2^20 pair evaluations
    t = time.perf_counter()
    bidx, _ = O.nn_brute(nx, ny, qx[:mb], qy[:mb])
    t_brute = time.perf_counter() - t
    extend = {"value": mq / t_ext, "unit": "steps/s", "cores": threads, "kind": "port", "nn_share": t_nn / t_ext,
              "single_thread_value": m1 / t_single,
              "nn_grid_queries_per_s": mq / t_nn, "nn_brute_queries_per_s": mb / t_brute,
              "nn_brute_matches_grid": bool(np.array_equal(bidx, oidx[:mb])),
              "sample": "2^16 queries of the C4 workload vs the 2^20-node tree and 10 k rings: exact grid NN + culled "
                        "Space::verify of the straight edge, all host threads; nn_brute: the first 2^11 of those queries by "
                        "exact brute force (SURVEY 8d asks for both NN comparators)"}
    return {"value": best, "unit": "pairs/s", "cores": threads, "kind": "port", "extend": extend,
            "sample": "2^22 pairs of the C3 workload, all host threads (OpenMP static), best of 3; single thread on 2^20 pairs",
            "single_thread_value": single,
            "note": "C restatement of src/dubins.rs (oracle/pp_oracle.c, gcc -O3 -ffp-contract=off -fno-fast-math), not rustc output"}


def secondary(args, torch, dist, pp, ctx, dev, rank, world, fp64_peak, hbm_peak):
    """C4 (RRT extend step: NN + straight-edge verify) and the per-GPU slice of C5 (Dubins-edge verify)"""
    out = {}
    steps = max(1, min(args.steps, 10))      # the millisecond-scale workloads
    slow_steps = max(1, min(args.steps, 2))  # the two deliberately slow yard-stick scans (0.36 / 0.48 s per step)
    # ---- C4: replicated tree + obstacles (rank 0 generates, one NCCL broadcast), queries sharded
    m = C4_M
    if rank == 0:
        _, _, nx, ny, nyaw = pp.synth.extend_inputs(1, C4_NODES)
        tree = torch.from_numpy(np.stack([nx, ny, nyaw])).to(dev)
    else:
        tree = torch.empty((3, C4_NODES), dtype=torch.float64, device=dev)
    if world > 1:
        dist.broadcast(tree, 0)  # NVLink/NVSwitch; the only collective on the path, outside the timed region
    ctx.tree_upload_dev(C4_NODES, tree[0], tree[1], tree[2])
    bounds, rings = pp.synth.circle_world(C4_RINGS)
    ctx.obstacles_upload(bounds, rings)
    qx = torch.from_numpy(pp.synth.uniform(pp.synth.SEED_C4_Q, 0, m, 0.0, 1000.0, first=rank * m)).to(dev)
    qy = torch.from_numpy(pp.synth.uniform(pp.synth.SEED_C4_Q, 1, m, 0.0, 1000.0, first=rank * m)).to(dev)
    idx = torch.empty(m, dtype=torch.int32, device=dev)
    yaw = torch.empty(m, dtype=torch.float64, device=dev)
    ok = torch.empty(m, dtype=torch.uint8, device=dev)
    ref_idx = None
    grid_build_ms = None
    # "extend" is the call a user makes (default flags: the library picks the exact grid search for a tree this size);
    # the *_scan* rows force the tiled brute-force kernels of the north-star design; all rows must agree bit for bit
    for name, nnf, cf, kname in [("extend_scan", 8, 8, "nn_scan"), ("extend", 0, 0, "nn_grid"),
                                 ("extend_scan_plain_f64", 1, 8, "nn_scan_f64"),
                                 ("extend_scan_unsorted", 4, 4, "nn_scan_unsorted")]:
        ctx.timing_enable(True)
        fn = lambda: ctx.rrt_extend_dev(m, qx, qy, idx, yaw, ok, nn_flags=nnf, collide_flags=cf)  # noqa: E731
        fn()
        torch.cuda.synchronize()
        if kname == "nn_grid" and grid_build_ms is None:
            b_ms, b_n = ctx.timing_get("nn_grid_build")  # the first grid call after the upload built the node grid
            grid_build_ms = b_ms / max(b_n, 1)
        ctx.timing_reset()
        l0 = ctx.launch_count
        ksteps = slow_steps if nnf in (1, 4) else steps
        ms, _, _ = time_steps(torch, dist, fn, ksteps, 0, world)
        nn_ms, nn_n = ctx.timing_get(kname)
        c_ms, c_n = ctx.timing_get({0: "collide_segments_grid", 8: "collide_segments", 4: "collide_segments_unsorted"}[cf])
        ctx.timing_enable(False)
        if ref_idx is None:
            ref_idx, ref_ok = idx.clone(), ok.clone()
        agree = bool(torch.equal(ref_idx, idx) and torch.equal(ref_ok, ok))
        pair_evals = float(m) * C4_NODES
        nn_s = nn_ms / max(nn_n, 1) * 1e-3
        out[name] = {
            "metric": "rrt_extend_steps_per_s", "value": world * m * ksteps / (ms * 1e-3), "unit": "steps/s",
            "ms_per_step": ms / ksteps, "steps": ksteps, "gpu_launches": ctx.launch_count - l0,
            "config": {"workload": f"c4: {m} queries/GPU vs {C4_NODES}-node tree, straight edge vs {C4_RINGS} create_circle rings",
                       "free_fraction_rank0": float(ok.float().mean().item())},
            "nn_kernel_ms": nn_ms / max(nn_n, 1), "collide_kernel_ms": c_ms / max(c_n, 1),
            "matches_scan": agree,
            "roofline": ({"kernel": kname, "bound": "fp64", "achieved": pair_evals * 6.0 / nn_s / 1e9, "peak": fp64_peak / 1e9,
                          "unit": "Ginstr/s", "frac": pair_evals * 6.0 / nn_s / fp64_peak,
                          "per_unit": "6 FP64-pipe instructions per (query, node) pair (SURVEY 8d yard-stick)",
                          "hbm_view": {"achieved": 36.0 * 2 ** 20 / nn_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                       "per_unit": "36 MiB algorithmic bytes per launch"},
                          # SURVEY 8d "report both": node tiles streamed L2 -> shared memory by TMA, m/Q x N x tile
                          # bytes per node with Q = 512 queries per CTA (20 B per node: x, y, fl32(x); the bucketed
                          # kernel streams only the 8 B fp32 copies)
                          "streamed_view": {"bytes_per_launch": m / 512.0 * C4_NODES * (8.0 if kname == "nn_scan" else 20.0),
                                            "achieved": m / 512.0 * C4_NODES * (8.0 if kname == "nn_scan" else 20.0) / nn_s / 1e9,
                                            "unit": "GB/s"}} if kname != "nn_grid" else
                         {"kernel": kname, "bound": "hbm", "achieved": 36.0 * 2 ** 20 / nn_s / 1e9, "peak": hbm_peak,
                          "unit": "GB/s", "frac": 36.0 * 2 ** 20 / nn_s / 1e9 / hbm_peak,
                          "per_unit": "36 MiB algorithmic bytes per launch (queries + nodes + indices); the search is a "
                                      "latency-bound gather, not a stream",
                          # what actually bounds it (ncu --set full, profiles/r01_rrt_kernels_final6_raw.csv): 25.4 M L2
                          # sectors per launch of 2^20 queries, l1tex throughput 78 %, lts throughput 60 %
                          "l2_view": {"l2_bytes_per_launch_ncu": NCU_NN_GRID_L2_BYTES, "unit": "GB/s",
                                      "achieved": NCU_NN_GRID_L2_BYTES / nn_s / 1e9, "l1tex_throughput_ncu": 0.777,
                                      "lts_throughput_ncu": 0.602}}),
        }
        if kname == "nn_grid":
            out[name]["nn_grid_build_ms_after_upload"] = grid_build_ms
            # the same step end to end through the host C-ABI call (pp_rrt_extend on pinned host buffers: 16 B in and
            # 13 B out per query cross PCIe inside the timed region); wall clock, max over ranks
            hq = [pp.PinnedArray(m, np.float64) for _ in range(2)]
            hq[0].array[:] = qx.cpu().numpy()
            hq[1].array[:] = qy.cpu().numpy()
            ho = (pp.PinnedArray(m, np.uint32), pp.PinnedArray(m, np.float64), pp.PinnedArray(m, np.uint8))
            outs = tuple(h.array for h in ho)
            ctx.rrt_extend(hq[0].array, hq[1].array, out=outs)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                ctx.rrt_extend(hq[0].array, hq[1].array, out=outs)
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            out[name]["e2e"] = {"value": world * m * steps / float(dt.item()), "unit": "steps/s",
                                "h2d_bytes_per_step": 16 * m, "d2h_bytes_per_step": 13 * m,
                                "matches_device_run": bool(np.array_equal(outs[0].astype(np.int32), idx.cpu().numpy())
                                                           and np.array_equal(outs[2], ok.cpu().numpy()))}
    # the same step with the reference's real edge geometry: Dubins curve new node -> nearest node, sampled at 0.1
    fn = lambda: ctx.rrt_extend_dubins_dev(m, qx, qy, 0.8, 0.1, idx, yaw, ok)  # noqa: E731
    ctx.timing_enable(True)
    fn()
    torch.cuda.synchronize()
    ctx.timing_reset()
    l0 = ctx.launch_count
    ms, _, _ = time_steps(torch, dist, fn, steps, 0, world)
    parts = {k: ctx.timing_get(k) for k in ("nn_grid", "dubins_plan", "collide_dubins")}
    ctx.timing_enable(False)
    out["extend_dubins"] = {
        "metric": "rrt_extend_steps_per_s", "value": world * m * steps / (ms * 1e-3), "unit": "steps/s",
        "ms_per_step": ms / steps, "steps": steps, "gpu_launches": ctx.launch_count - l0,
        "kernel_ms": {k: v[0] / max(v[1], 1) for k, v in parts.items()},
        "config": {"workload": f"c4 with Dubins edges: {m} samples/GPU vs {C4_NODES}-node tree, NN + Node::new yaw + fused "
                               f"sample-and-verify of the Dubins edge (turn radius 0.8, step 0.1) vs {C4_RINGS} rings",
                   "free_fraction_rank0": float(ok.float().mean().item())},
    }
    # no-hit obstacle set: same rings translated outside the world, so early exit cannot flatter the number
    bounds2, rings2 = pp.synth.circle_world(C4_RINGS, shift=5000.0)
    ctx.obstacles_upload(bounds2, rings2)
    for name, cf, kname, what in [("collide_scan_nohit", 8, "collide_segments", "tiled fp32 box scan over all 10k rings"),
                                  ("collide_grid_nohit", 0, "collide_segments_grid", "uniform obstacle grid, the default")]:
        fn = lambda: ctx.rrt_extend_dev(m, qx, qy, idx, yaw, ok, nn_flags=2, collide_flags=cf)  # noqa: E731
        ctx.timing_enable(True)
        fn()
        torch.cuda.synchronize()
        ctx.timing_reset()
        ms, _, _ = time_steps(torch, dist, fn, steps, 0, world)
        c_ms, c_n = ctx.timing_get(kname)
        ctx.timing_enable(False)
        out[name] = {"collide_kernel_ms": c_ms / max(c_n, 1), "edges_per_s": m / (c_ms / max(c_n, 1) * 1e-3),
                     "free_fraction_rank0": float(ok.float().mean().item()),
                     "config": {"workload": f"c4 edges vs the no-hit ring set ({what})"}}
    # ---- C5 slice: Dubins edges sampled at 0.05 and verified against 100k rings
    e = C5_EDGES
    bounds5, rings5 = pp.synth.circle_world(C5_RINGS, rmin=0.5, rmax=1.5)
    ctx.obstacles_upload(bounds5, rings5)
    edges = [torch.from_numpy(a).to(dev) for a in pp.synth.dubins_edges(e, first=rank * e)]
    ok5 = torch.empty(e, dtype=torch.uint8, device=dev)
    fn = lambda: ctx.collide_dubins_dev(e, *edges, 1.0, 0.05, ok5)  # noqa: E731
    ctx.timing_enable(True)
    fn()
    torch.cuda.synchronize()
    ctx.timing_reset()
    l0 = ctx.launch_count
    ms, _, _ = time_steps(torch, dist, fn, steps, 0, world)
    p_ms, p_n = ctx.timing_get("dubins_plan")
    v_ms, v_n = ctx.timing_get("collide_dubins")
    ctx.timing_enable(False)
    out["dubins_rrt"] = {
        "metric": "dubins_edges_verified_per_s", "value": world * e * steps / (ms * 1e-3), "unit": "edges/s",
        "ms_per_step": ms / steps, "steps": steps, "gpu_launches": ctx.launch_count - l0,
        "plan_kernel_ms": p_ms / max(p_n, 1), "verify_kernel_ms": v_ms / max(v_n, 1),
        "config": {"workload": f"c5 slice: {e} Dubins edges/GPU (2^22 over 8 GPUs), step 0.05, radius 1.0, vs {C5_RINGS} rings",
                   "free_fraction_rank0": float(ok5.float().mean().item())},
    }
    # same edges against the no-hit ring set: no early exit, every sample segment is generated and tested
    _, rings6 = pp.synth.circle_world(C5_RINGS, rmin=0.5, rmax=1.5, shift=5000.0)
    bx = np.array([-100.0, -100.0, 1100.0, 1100.0, -100.0])
    ctx.obstacles_upload((bx, np.array([-100.0, 1100.0, 1100.0, -100.0, -100.0])), rings6)
    ctx.timing_enable(True)
    fn()
    torch.cuda.synchronize()
    ctx.timing_reset()
    ms, _, _ = time_steps(torch, dist, fn, steps, 0, world)
    v_ms, v_n = ctx.timing_get("collide_dubins")
    ctx.timing_enable(False)
    out["dubins_rrt_nohit"] = {
        "metric": "dubins_edges_verified_per_s", "value": world * e * steps / (ms * 1e-3), "unit": "edges/s",
        "ms_per_step": ms / steps, "verify_kernel_ms": v_ms / max(v_n, 1),
        "config": {"workload": "c5 slice against rings translated outside the (enlarged) bounds: every sample is tested",
                   "free_fraction_rank0": float(ok5.float().mean().item())},
    }
    # ---- sample materialisation (config-1 style paths in bulk): count -> prefix sum -> fill
    ns = 1 << 16
    se = [torch.from_numpy(a).to(dev) for a in pp.synth.dubins_edges(ns, first=rank * ns)]
    counts = torch.empty(ns, dtype=torch.int32, device=dev)
    plan = torch.empty(ns * 112, dtype=torch.uint8, device=dev)
    offsets = torch.empty(ns, dtype=torch.int64, device=dev)
    total_d = torch.zeros(1, dtype=torch.int64, device=dev)
    ctx.dubins_sample_count_dev(ns, *se, 1.0, 0.05, counts, plan)
    ctx.exclusive_scan_u32_dev(ns, counts, offsets, total_d)
    torch.cuda.synchronize()
    total = int(total_d.item())
    samples = torch.empty(total * 3, dtype=torch.float64, device=dev)

    def fill_step():
        ctx.dubins_sample_count_dev(ns, *se, 1.0, 0.05, counts, plan)
        ctx.exclusive_scan_u32_dev(ns, counts, offsets, total_d)
        ctx.dubins_sample_fill_dev(ns, plan, offsets, total, samples)

    ctx.timing_enable(True)
    fill_step()
    torch.cuda.synchronize()
    ctx.timing_reset()
    l0 = ctx.launch_count
    ms, _, _ = time_steps(torch, dist, fill_step, steps, 0, world)
    f_ms, f_n = ctx.timing_get("dubins_fill")
    ctx.timing_enable(False)
    f_s = f_ms / max(f_n, 1) * 1e-3
    out["dubins_sample"] = {
        "metric": "dubins_samples_per_s", "value": world * total * steps / (ms * 1e-3), "unit": "samples/s",
        "ms_per_step": ms / steps, "fill_kernel_ms": f_ms / max(f_n, 1), "gpu_launches": ctx.launch_count - l0,
        "config": {"workload": f"{ns} Dubins paths/GPU (c5 edge distribution), step 0.05: count + scan + fill of {total} samples"},
        "roofline": {"kernel": "pp_dubins_fill_kernel", "bound": "hbm", "achieved": total * 24.0 / f_s / 1e9, "peak": hbm_peak,
                     "unit": "GB/s", "frac": total * 24.0 / f_s / 1e9 / hbm_peak, "traffic": None,
                     "per_unit": "24 B written per sample (+112 B plan record per path)"},
    }
    # ---- C2 (functional, host-driven): the examples/rrt shape -- 100 x 100 world, 50 create_circle obstacles, 10 k
    # iterations, fixed seed -- through the Python mirror's batched rounds (NN -> yaw -> fused Dubins verify ->
    # append -> goal connection per round).  Wall clock of the whole planner incl. the Python bookkeeping and every
    # host<->device copy: a statement about the drop-in API, not a kernel benchmark.  Every rank runs its own copy.
    try:
        rng = np.random.default_rng(0xC2)
        rings = [pp.rrt.create_circle((float(cx), float(cy)), float(r))
                 for cx, cy, r in zip(rng.uniform(10, 90, 50), rng.uniform(10, 90, 50), rng.uniform(1.0, 3.0, 50))]
        bounds_ring = (np.array([0.0, 0.0, 100.0, 100.0, 0.0]), np.array([0.0, 100.0, 100.0, 0.0, 0.0]))
        space = pp.rrt.Space.from_inflated(bounds_ring, pp.rrt.Robot(1.8, 3.0, 0.8), rings, ctx=ctx, seed=7)
        planner = pp.rrt.RRT((4.0, 4.0), 0.0, (96.0, 96.0), 0.0, 10_000, 0.1, space)
        l0 = ctx.launch_count
        t0 = time.perf_counter()
        path = planner.plan_rounds(batch=512)
        dt = time.perf_counter() - t0
        out["c2_rrt_plan_rounds"] = {
            "metric": "rrt_iterations_per_s", "value": 10_000 / dt, "unit": "iterations/s", "wall_s": dt,
            "tree_nodes": len(planner.nodes), "path_found": path is not None,
            "path_points": 0 if path is None else int(len(path[0])), "gpu_launches": ctx.launch_count - l0,
            "config": {"workload": "c2: RRT.plan_rounds(batch=512), 10 000 iterations, 100 x 100 world, 50 create_circle "
                                   "obstacles (r in [1, 3)), Robot(1.8, 3.0, 0.8), step 0.1, seed 7; host-driven, per rank"},
        }
    except Exception as e:  # the functional leg must never take the benchmark line down
        out["c2_rrt_plan_rounds"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    return out


def main():
    # libraries (NCCL's version banner, ...) may print to fd 1: keep the real stdout for the one JSON line
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--skip-secondary", action="store_true", help="only the primary C3 metric")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--profile", action="store_true", help="for ncu runs: no 1 s clock-sampling load phase")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()

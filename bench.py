#!/usr/bin/env python
"""bench.py -- headline measurement of the hot path (BASELINE.json: "Dubins pairs/s; RRT extend steps/s").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--skip-secondary]

One JSON line on stdout (rank 0).  Primary metric: batched Dubins pose-pair evaluations per second on
config C3 (2^24 pairs per GPU, radius 1.0, shortest-word selection + length only).  The same line carries
the RRT extend-step (C4) and Dubins-edge verify (C5 slice) numbers under "workloads".
N > 1 is launched by torch.distributed.run, one rank per GPU; the path shards with no data-path collective
(weak scaling: every GPU takes its own 2^24-pair slice; the "strong" block splits ONE 2^24-pair and ONE 2^22-edge
batch over the ranks); tree / obstacle buffers are replicated by the library's own ncclBroadcast
(pp_tree_upload_bcast, pp_obstacles_upload_bcast), and the 512-node tail broadcast is timed on its own.
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
PASSES_PER_STEP = 128      # a step = 128 passes of the kernel over the 2^24-pair batch (inputs 768 MiB > L2, so a
                           # pass never finds its data in cache): K = 10 steps then last 0.8 s and the clock samples,
                           # the throttle flags and the driver's own clock around the run describe the timed region
W_INSTR_PER_PAIR = 1100.0  # fixed yard-stick of SURVEY.md Appendix D (FP64-pipe thread-instructions per pair)
BYTES_PER_PAIR = 57.0      # 48 B read + 8 B cost + 1 B word
FP64_PEAK_NOMINAL = 148 * 64 * 1.965e9  # lanes * clock: used only if the live DFMA measurement fails
C4_M, C4_NODES, C4_RINGS = 1 << 20, 1 << 20, 10_000
C5_EDGES, C5_RINGS = 1 << 19, 100_000  # the per-GPU slice of config 5 (2^22 edges over 8 GPUs)
C5_EDGES_TOTAL = 1 << 22
# the workload both arms are measured on (the driver compares the two `config` objects)
C3_CONFIG = {"workload": "c3_batched_dubins: 2^24 random pose pairs per GPU ('mixed' U[-2,2)^2 positions, U[-pi,pi) yaws, "
                         "seed 0xD0B10003), radius 1.0, shortest-word selection + length only",
             "pairs_per_gpu": N_PAIRS, "l2": "inputs (768 MiB) larger than L2, no flush needed",
             "sharding": "contiguous slice per rank, no data-path collective"}


def kernel_facts():
    """per-kernel figures read from the committed ncu captures (tools/ncu_facts.py -> profiles/kernel_facts.json):
    executed FP64-pipe instructions and DRAM bytes per launch.  Nothing here is a constant of this file."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "kernel_facts.json")))
    except Exception:
        return {}


def fact(facts, prefix, key, default=None):
    for name, f in facts.items():
        if name.startswith(prefix) and f.get(key) is not None:
            return f[key], f.get("source")
    return default, None


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
def load_synth_standalone():
    """the synthetic-input generator WITHOUT importing the product package (whose __init__ loads the CUDA library):
    the reference arm must not map libpathplanning_b200.so at all"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("pp_synth_standalone",
                                                  os.path.join(ROOT, "rs-pathplanning_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_reference(args):
    """the reference's CPU implementation of the path = the C restatement (oracle port; no rustc here) on all host
    threads, on the SAME workload: every step is one pass over the full 2^24-pair batch"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    O = graft.import_oracle()
    O.build()
    synth = load_synth_standalone()
    assert "rs_pathplanning_b200" not in sys.modules
    threads = O.max_threads()
    n = N_PAIRS
    sx, sy, syaw, ex, ey, eyaw = synth.dubins_pairs(n)
    cost, word = None, None
    for _ in range(max(min(args.warmup, 2), 1)):
        k = n // 8
        O.dubins_eval_batch(sx[:k], sy[:k], syaw[:k], ex[:k], ey[:k], eyaw[:k], 1.0, want_flags=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cost, word, _, _ = O.dubins_eval_batch(sx, sy, syaw, ex, ey, eyaw, 1.0, want_flags=False)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"{args.steps} steps x the full 2^24-pair C3 batch (seed 0xD0B10003), OpenMP static partition over {threads} threads"
    line = {
        "impl": "reference", "metric": "dubins_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": C3_CONFIG,
        "step": {"passes_per_step": 1, "pairs_per_step": n},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "word_hist": np.bincount(word, minlength=6)[:6].tolist(),
        "note": "C restatement of src/dubins.rs (oracle/pp_oracle.c), not rustc output: no Rust toolchain in the image; "
                "parity of this oracle is UNPINNED by the reference (it ships no tests)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ own arm
def bind_to_gpu_numa(local):
    """pin this rank's host threads (hence its first-touch pinned buffers) to the NUMA node its GPU hangs off: with 8
    ranks streaming 50 GB/s each, remote-socket staging halves the end-to-end rate.  Always returns a dict that says
    what was done or WHY nothing could be done."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml pads the domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")] \
            if os.path.isdir("/sys/devices/system/node") else []
        path = f"/sys/bus/pci/devices/{bus}/numa_node"
        if not os.path.exists(path):
            return {"bound": False, "reason": f"{path} does not exist", "host_numa_nodes": len(nodes)}
        node = int(open(path).read().strip())
        if node < 0:
            return {"bound": False, "reason": "sysfs numa_node = -1 for the GPU's PCI device: the (virtualised) host "
                    "exposes no GPU-to-node affinity", "host_numa_nodes": len(nodes)}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"bound": False, "reason": f"no allowed CPU on node {node}", "host_numa_nodes": len(nodes)}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "numa_node": node, "cpus": len(cpus), "host_numa_nodes": len(nodes)}
    except Exception as e:  # pragma: no cover
        return {"bound": False, "reason": f"{type(e).__name__}: {e}"[:120]}


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
    numa = bind_to_gpu_numa(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    # copy threads of the staged (pageable-memory) path: the ranks of one host share its cores, so each takes its
    # share of half the hardware threads (the library's own default is what a single process would use)
    stage_threads = max(2, min(8, len(os.sched_getaffinity(0)) // (2 * world)))
    os.environ.setdefault("PP_STAGE_THREADS", str(stage_threads))
    ctx = pp.Context(local)
    if world > 1:
        # the library's own communicator (ncclBroadcast of tree / tail / obstacles inside the C-ABI): rank 0 makes the
        # id, the launcher's channel (torch.distributed) hands the 128 bytes to the other ranks
        ident = [pp.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, 0)
        ctx.comm_init(ident[0], world, rank)
    # a dedicated non-default stream shared by torch (events, copies) and the library's launches, so that
    # torch.cuda.Event brackets exactly our kernels (the legacy default stream's handle 0 means "own stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    peaks = measured_peaks()
    hbm_peak = (peaks or {}).get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    facts = kernel_facts()

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

    def one_pass():
        ctx.dubins_eval_dev(n, *d_in, 1.0, d_cost, d_word)

    def step():
        for _ in range(PASSES_PER_STEP):
            one_pass()

    sampler = ClockSampler(local) if rank == 0 else None
    ctx.timing_enable(True)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    ctx.timing_reset()
    l0 = ctx.launch_count
    ms, t0, t1 = time_steps(torch, dist, step, args.steps, 0, world)
    launches = ctx.launch_count - l0
    k_ms, k_n = ctx.timing_get("dubins_eval")
    ctx.timing_enable(False)
    clocks = sampler.stop(t0, t1) if sampler else None
    if clocks is not None:
        clocks["window"] = f"the timed region itself ({(t1 - t0):.2f} s of wall clock)"
    value = world * n * PASSES_PER_STEP * args.steps / (ms * 1e-3)
    k_avg_ms = k_ms / max(k_n, 1)
    pairs_per_s_kernel = n / (k_avg_ms * 1e-3)

    # sanity of what was computed (never a fallback): word histogram on rank 0
    hist = torch.bincount(d_word.to(torch.int64), minlength=256)[:6].tolist()

    # ---- e2e through the C-ABI with HOST buffers (H2D + kernel + D2H inside the timed region): pinned buffers of the
    # library (pp_host_alloc), and the plain pageable arrays a caller holding Vec<f64> / numpy memory passes
    e2e_steps = max(1, min(args.steps, 3))
    d_cost_host, d_word_host = d_cost.cpu().numpy(), d_word.cpu().numpy()

    def e2e_leg(arrays, out):
        ctx.dubins_eval(*arrays, radius=1.0, want_tpq=False, out=out)  # warm-up
        if world > 1:
            dist.barrier()
        te = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.dubins_eval(*arrays, radius=1.0, want_tpq=False, out=out)
        t = torch.tensor([time.perf_counter() - te], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        same = bool(np.array_equal(out[0], d_cost_host) and np.array_equal(out[1], d_word_host))
        return world * n * e2e_steps / float(t.item()), same

    pins = [pp.PinnedArray(n, np.float64) for _ in range(6)]
    for p, a in zip(pins, host):
        p.array[:] = a
    pcost, pword = pp.PinnedArray(n, np.float64), pp.PinnedArray(n, np.uint8)
    e2e_value, same = e2e_leg([p.array for p in pins], (pcost.array, pword.array, None))
    del pins, pcost, pword
    pg_cost, pg_word = np.empty(n, np.float64), np.empty(n, np.uint8)
    e2e_pageable, same_pg = e2e_leg(list(host), (pg_cost, pg_word, None))
    del pg_cost, pg_word
    e2e_launches = 2 * 16 * (e2e_steps + 1)
    # copy-only ceilings of THIS box, same bytes, no kernel (all ranks at once: the host side is shared)
    if world > 1:
        dist.barrier()
    copy_ms = {k: ctx.measure_copy(48 * n, 9 * n, pinned=(k == "pinned")) for k in ("pinned", "pageable")}
    for k in copy_ms:
        t = torch.tensor([copy_ms[k]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        copy_ms[k] = float(t.item())
    ceil_pinned = world * n / (copy_ms["pinned"] * 1e-3)
    ceil_pageable = world * n / (copy_ms["pageable"] * 1e-3)
    del host

    strong = strong_block(args, torch, dist, pp, ctx, dev, rank, world)

    workloads = {}
    if not args.skip_secondary:
        # C3's secondary "far" distribution (positions U[-50,50)^2, CSC words ~99.9 %): same kernel, same
        # buffers, other data -- the evaluation is branch-free, so the time should not depend on the word mix
        far = pp.synth.dubins_pairs(n, "far", first=rank * n)
        for d, a in zip(d_in, far):
            d.copy_(torch.from_numpy(a))
        del far
        for _ in range(3):
            one_pass()
        far_steps = max(1, min(args.steps, 10))
        l0 = ctx.launch_count
        far_ms, _, _ = time_steps(torch, dist, one_pass, far_steps, 0, world)
        far_hist = torch.bincount(d_word.to(torch.int64), minlength=256)[:6].tolist()
        workloads["dubins_far"] = {
            "metric": "dubins_pairs_per_s", "value": world * n * far_steps / (far_ms * 1e-3), "unit": "pairs/s",
            "ms_per_step": far_ms / far_steps, "steps": far_steps, "gpu_launches": ctx.launch_count - l0,
            "config": {"workload": "c3 'far': 2^24 pose pairs per GPU, positions U[-50,50)^2, yaws U[-pi,pi), radius 1.0",
                       "word_hist_rank0": far_hist}}
        del d_in
        workloads.update(secondary(args, torch, dist, pp, ctx, dev, rank, world, fp64_peak, hbm_peak, facts))

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu_baseline = cpu_baseline_leg(pp)

    if rank == 0:
        # executed FP64-pipe work per pair: from the committed ncu capture of this kernel (profiles/kernel_facts.json),
        # not a constant of this file; `frac` = live pairs/s x that / the live DFMA peak = the pipe utilisation
        ex_instr, ex_src = fact(facts, "pp_dubins_eval_kernel", "fp64_thread_instr_per_unit")
        traffic, tr_src = fact(facts, "pp_dubins_eval_kernel", "dram_bytes")
        busy_ncu, _ = fact(facts, "pp_dubins_eval_kernel", "fp64_pipe_pct_active")
        achieved = pairs_per_s_kernel * ex_instr / 1e9 if ex_instr else None
        line = {
            "metric": "dubins_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": C3_CONFIG,
            "step": {"passes_per_step": PASSES_PER_STEP, "pairs_per_step": n * PASSES_PER_STEP, "word_hist_rank0": hist,
                     "timed_region_s": ms * 1e-3,
                     "why": "one pass over the 2^24-pair batch lasts 0.6 ms; a step repeats it so that the timed region is "
                            ">= 0.5 s (inputs exceed L2: no pass finds its data cached)"},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": 48 * n, "d2h_bytes_per_step": 9 * n,
                    "steps": e2e_steps, "matches_device_run": same,
                    "how": "pp_dubins_eval on pinned host buffers (pp_host_alloc): 16 chunks over 3 streams, copies inside "
                           "the timed region",
                    "copy_ceiling": {"value": ceil_pinned, "unit": "pairs/s", "ms": copy_ms["pinned"],
                                     "what": "the same 805 MB up + 151 MB down per rank, both directions at once from pinned "
                                             "memory, all ranks together, NO kernel (pp_measure_copy, best of 3)"},
                    "frac_of_copy_ceiling": e2e_value / ceil_pinned,
                    "numa_binding_rank0": numa},
            "e2e_pageable": {"value": e2e_pageable, "unit": "pairs/s", "h2d_bytes_per_step": 48 * n,
                             "d2h_bytes_per_step": 9 * n, "steps": e2e_steps, "matches_device_run": same_pg,
                             "how": "the same call on ordinary pageable arrays (what a Rust caller holding Vec<f64> passes): "
                                    "the library stages each chunk through its own pinned ring with a few copy threads "
                                    "(csrc/pp_stage.hpp) while the previous chunks are on the wire, instead of leaving "
                                    "the staging to the driver's single-threaded bounce-buffer path",
                             "driver_pageable_copy": {"value": ceil_pageable, "unit": "pairs/s", "ms": copy_ms["pageable"],
                                                      "what": "plain cudaMemcpyAsync of the same bytes from / to pageable "
                                                              "memory, no kernel: what the call cost before the staged path"},
                             "vs_driver_pageable_copy": e2e_pageable / ceil_pageable,
                             "copy_threads_per_rank": int(os.environ["PP_STAGE_THREADS"]),
                             "frac_of_copy_ceiling": e2e_pageable / ceil_pinned},
            "gpu_launches": launches + e2e_launches + sum(w.get("gpu_launches", 0) for w in workloads.values())
            + strong.get("gpu_launches", 0),
            "gpu_launches_primary_timed_region": launches,
            "clocks": clocks,
            "roofline": {
                "kernel": "pp_dubins_eval_kernel", "bound": "fp64", "achieved": achieved, "peak": fp64_peak / 1e9,
                "unit": "Ginstr/s", "frac": (achieved * 1e9 / fp64_peak) if achieved else None,
                "per_unit": f"{ex_instr:.0f} executed FP64-pipe thread-instructions per pair, from {ex_src}" if ex_instr else None,
                "fp64_pipe_busy_ncu": busy_ncu / 100.0 if busy_ncu else None,
                "traffic": traffic, "traffic_source": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per "
                f"launch of 2^24 pairs ({tr_src}); algorithmic {BYTES_PER_PAIR * n:.4g}",
                # the SURVEY 8d yard-stick (W = 1100 instructions per pair, fixed before any kernel existed) is 2.2x the
                # work this kernel executes, so the fraction against it exceeds 1; kept for continuity only
                "yardstick_frac": pairs_per_s_kernel * W_INSTR_PER_PAIR / fp64_peak,
                "yardstick_per_unit": f"W = {W_INSTR_PER_PAIR:.0f} FP64-pipe thread-instructions per pair (SURVEY App. D)",
                "peak_source": fp64_src, "kernel_ms_avg": k_avg_ms, "kernel_launches_timed": k_n,
                "hbm_view": {"bound": "hbm", "achieved": pairs_per_s_kernel * BYTES_PER_PAIR / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": pairs_per_s_kernel * BYTES_PER_PAIR / 1e9 / hbm_peak,
                             "per_unit": "57 B per pair (48 read + 9 written)", "peak_source": hbm_src},
            },
            "cpu_baseline": cpu_baseline,
            "strong": strong,
            "workloads": workloads,
            "parity": "oracle port, UNPINNED by the reference (it ships no tests and cannot be compiled here)",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def strong_block(args, torch, dist, pp, ctx, dev, rank, world):
    """fixed-TOTAL work split over the ranks (north_star: "16M pairs (1/2/4/8 B200)", "4M edges sharded across 8"):
    C3 = ONE 2^24-pair batch, C5 = ONE 2^22-edge batch; rank r takes the contiguous slice [r*n/W, (r+1)*n/W)
    (pp_slice_bounds).  Device-resident times are CUDA events on the launching stream, max over ranks; the end-to-end
    C3 figure goes through the host entry point with pinned buffers.  Also the latency of the only collective on the
    path: a 512-node pp_tree_append_bcast (H2D on the root + ncclBroadcast of the tail + the finishing kernels)."""
    out = {"scaling": "strong", "n_gpus": world}
    steps = max(1, min(args.steps, 10))
    l_start = ctx.launch_count
    # ---- C3, 2^24 pairs in total
    lo, hi = pp.slice_bounds(N_PAIRS, world, rank)
    cnt = hi - lo
    host = pp.synth.dubins_pairs(cnt, "mixed", first=lo)
    d_in = [torch.from_numpy(a).to(dev) for a in host]
    d_cost = torch.empty(cnt, dtype=torch.float64, device=dev)
    d_word = torch.empty(cnt, dtype=torch.uint8, device=dev)
    fn = lambda: ctx.dubins_eval_dev(cnt, *d_in, 1.0, d_cost, d_word)  # noqa: E731
    ms, _, _ = time_steps(torch, dist, fn, steps * 8, 3, world)
    pins = [pp.PinnedArray(cnt, np.float64) for _ in range(6)]
    for p_, a in zip(pins, host):
        p_.array[:] = a
    pc, pw = pp.PinnedArray(cnt, np.float64), pp.PinnedArray(cnt, np.uint8)
    arrays, outs = [p_.array for p_ in pins], (pc.array, pw.array, None)
    ctx.dubins_eval(*arrays, radius=1.0, want_tpq=False, out=outs)
    if world > 1:
        dist.barrier()
    te = time.perf_counter()
    for _ in range(3):
        ctx.dubins_eval(*arrays, radius=1.0, want_tpq=False, out=outs)
    t = torch.tensor([time.perf_counter() - te], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # checksum of the gathered result (the same for every device count: pure slicing)
    chk = torch.tensor([float(np.sum(pc.array)), float(np.sum(pw.array.astype(np.float64)))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(chk, op=dist.ReduceOp.SUM)
    out["c3"] = {"metric": "dubins_pairs_per_s", "total_pairs": N_PAIRS, "value": N_PAIRS * steps * 8 / (ms * 1e-3),
                 "unit": "pairs/s", "ms_per_batch": ms / (steps * 8),
                 "e2e": {"value": N_PAIRS * 3 / float(t.item()), "unit": "pairs/s", "ms_per_batch": float(t.item()) / 3 * 1e3},
                 "checksum": {"sum_cost": float(chk[0].item()), "sum_word": float(chk[1].item())}}
    del pins, pc, pw, d_in, d_cost, d_word, host
    # ---- C5, 2^22 Dubins edges in total vs 100 k rings (replicated by the library's broadcast)
    e_lo, e_hi = pp.slice_bounds(C5_EDGES_TOTAL, world, rank)
    e = e_hi - e_lo
    if rank == 0:
        bounds5, rings5 = pp.synth.circle_world(C5_RINGS, rmin=0.5, rmax=1.5)
        ctx.obstacles_upload_bcast(0, bounds5, rings5)
    else:
        ctx.obstacles_upload_bcast(0)
    edges = [torch.from_numpy(a).to(dev) for a in pp.synth.dubins_edges(e, first=e_lo)]
    ok5 = torch.empty(e, dtype=torch.uint8, device=dev)
    fn = lambda: ctx.collide_dubins_dev(e, *edges, 1.0, 0.05, ok5)  # noqa: E731
    ms, _, _ = time_steps(torch, dist, fn, steps, 2, world)
    free = ok5.sum().to(torch.float64)
    if world > 1:
        dist.all_reduce(free, op=dist.ReduceOp.SUM)
    out["c5"] = {"metric": "dubins_edges_verified_per_s", "total_edges": C5_EDGES_TOTAL,
                 "value": C5_EDGES_TOTAL * steps / (ms * 1e-3), "unit": "edges/s", "ms_per_batch": ms / steps,
                 "free_edges": int(free.item())}
    del edges, ok5
    # ---- the collective: 512-node append, tail broadcast
    rng = np.random.default_rng(3)
    base = 1 << 16
    bx, by = rng.uniform(0, 1000, base), rng.uniform(0, 1000, base)
    if rank == 0:
        ctx.tree_upload_bcast(0, base, bx, by)
    else:
        ctx.tree_upload_bcast(0, base)
    k, reps = 512, 40
    tx, ty = rng.uniform(0, 1000, k), rng.uniform(0, 1000, k)
    for _ in range(3):
        ctx.tree_append_bcast(0, k, tx, ty) if rank == 0 else ctx.tree_append_bcast(0, k)
    if world > 1:
        dist.barrier()
    te = time.perf_counter()
    for _ in range(reps):
        ctx.tree_append_bcast(0, k, tx, ty) if rank == 0 else ctx.tree_append_bcast(0, k)
    t = torch.tensor([time.perf_counter() - te], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sizes = torch.tensor([ctx.tree_size], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(sizes, op=dist.ReduceOp.MIN)
    out["tree_append_512"] = {"ms": float(t.item()) / reps * 1e3, "reps": reps, "bytes_broadcast": k * 28,
                              "tree_size_all_ranks": int(sizes.item()),
                              "what": "pp_tree_append_bcast of 512 nodes: H2D on the root, one fused ncclBroadcast of the "
                                      "tail (x, y, yaw, parent) over NVLink, fp32 copies + padding kernels, stream sync; "
                                      "wall clock, max over ranks (N = 1: no collective)"}
    out["gpu_launches"] = ctx.launch_count - l_start
    return out


C1_POSE = (1.0, 1.0, np.pi / 4.0, -3.0, -3.0, -np.pi / 4.0, 1.0, 0.1)  # benches/all.rs:102-111 (c read as turn_radius)


def single_path_latency(fn, reps=2000, warm=50):
    """C1 (first row): wall clock per scalar dubins_path_planning call on the bench pose.
    fn(*C1_POSE) -> (px, py, pyaw, word, cost) or an object with .x / .word / .cost"""
    for _ in range(warm):
        fn(*C1_POSE)
    t0 = time.perf_counter()
    for _ in range(reps):
        p = fn(*C1_POSE)
    dt = time.perf_counter() - t0
    px, word, cost = (p[0], p[3], p[4]) if isinstance(p, tuple) else (p.x, p.word, p.cost)
    return {"value": dt / reps * 1e6, "unit": "us/path", "higher_is_better": False, "calls": reps,
            "samples": int(len(px)), "word": int(word), "cost": float(cost)}


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
    try:  # C1: the scalar call on one thread (through ctypes, as the GPU figure in workloads.c1_single_path is)
        c1 = single_path_latency(O.dubins_path)
        c1["what"] = ("oracle port, one thread: evaluate + sample the bench pose (95 samples); value = one call at a time "
                      "through ctypes (what workloads.c1_single_path pays too), in_c_loop_us = the same call repeated "
                      "inside one C loop (no Python), the figure to hold against criterion's")
        reps = 20_000
        pose = [np.full(reps, v) for v in C1_POSE[:6]]
        O.dubins_count_batch(*[v[:256] for v in pose], C1_POSE[6], C1_POSE[7], nthreads=1)
        t = time.perf_counter()
        cnt = O.dubins_count_batch(*pose, C1_POSE[6], C1_POSE[7], nthreads=1)
        c1["in_c_loop_us"] = (time.perf_counter() - t) / reps * 1e6
        c1["in_c_loop_samples"] = int(cnt[0])
    except Exception as e:
        c1 = {"error": f"{type(e).__name__}: {e}"[:200]}
    return {"value": best, "unit": "pairs/s", "cores": threads, "kind": "port", "extend": extend, "c1_single_path": c1,
            "sample": "2^22 pairs of the C3 workload, all host threads (OpenMP static), best of 3; single thread on 2^20 pairs",
            "single_thread_value": single,
            "note": "C restatement of src/dubins.rs (oracle/pp_oracle.c, gcc -O3 -ffp-contract=off -fno-fast-math), not rustc output"}


def secondary(args, torch, dist, pp, ctx, dev, rank, world, fp64_peak, hbm_peak, facts):
    """C4 (RRT extend step: NN + straight-edge verify) and the per-GPU slice of C5 (Dubins-edge verify)"""
    out = {}
    steps = max(1, min(args.steps, 10))      # the millisecond-scale workloads
    slow_steps = max(1, min(args.steps, 2))  # the two deliberately slow yard-stick scans (0.36 / 0.48 s per step)
    # ---- C4: tree + obstacles replicated INSIDE the library (rank 0 holds them, pp_tree_upload_bcast /
    # pp_obstacles_upload_bcast = ncclBroadcast over NVLink, outside the timed region), queries sharded
    m = C4_M
    if rank == 0:
        _, _, nx, ny, nyaw = pp.synth.extend_inputs(1, C4_NODES)
        bounds, rings = pp.synth.circle_world(C4_RINGS)
        ctx.tree_upload_bcast(0, C4_NODES, nx, ny, nyaw)
        ctx.obstacles_upload_bcast(0, bounds, rings)
    else:
        ctx.tree_upload_bcast(0, C4_NODES)
        ctx.obstacles_upload_bcast(0)
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
                                 ("extend_fused", 0, 16, "extend_fused"),
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
        c_ms, c_n = ctx.timing_get({16: "extend_sort", 0: "collide_segments_grid", 8: "collide_segments",
                                    4: "collide_segments_unsorted"}[cf])
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
            **({"fused_kernel_ms": nn_ms / max(nn_n, 1), "sort_kernels_ms": c_ms / max(c_n, 1)} if kname == "extend_fused"
               else {"nn_kernel_ms": nn_ms / max(nn_n, 1), "collide_kernel_ms": c_ms / max(c_n, 1)}),
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
                                            "unit": "GB/s"}} if kname not in ("nn_grid", "extend_fused") else
                         {"kernel": kname, "bound": "hbm", "achieved": 36.0 * 2 ** 20 / nn_s / 1e9, "peak": hbm_peak,
                          "unit": "GB/s", "frac": 36.0 * 2 ** 20 / nn_s / 1e9 / hbm_peak,
                          "per_unit": ("36 MiB (queries + nodes + indices) + 9 MiB (yaw + ok) algorithmic bytes per fused launch"
                                       if kname == "extend_fused" else "36 MiB algorithmic bytes per launch (queries + nodes + "
                                       "indices)") + "; the search is a latency-bound gather, not a stream",
                          # what actually bounds it (ncu --set full, profiles/r01_rrt_kernels_final6_raw.csv): 25.4 M L2
                          # sectors per launch of 2^20 queries, l1tex throughput 78 %, lts throughput 60 %
                          "l2_view": (lambda sec, src: {"l2_bytes_per_launch_ncu": sec * 32.0 if sec else None, "unit": "GB/s",
                                                        "achieved": sec * 32.0 / nn_s / 1e9 if sec else None,
                                                        "source": src})(*fact(facts, "pp_rrt_extend_fused_kernel" if kname == "extend_fused"
                                                                              else "pp_nn_grid_kernel", "l2_sectors"))}),
        }
        if kname == "nn_grid":
            out[name]["nn_grid_build_ms_after_upload"] = grid_build_ms
        if kname == "nn_grid" and not args.profile:
            # the same step end to end through the host C-ABI call (pp_rrt_extend on pinned host buffers: 16 B in and
            # 13 B out per query cross PCIe inside the timed region, in 2^18-query chunks rotating over three streams
            # so that uploads, the two kernels and downloads overlap); wall clock, max over ranks
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
    parts = {k: ctx.timing_get(k) for k in ("nn_grid", "extend_gather", "dubins_plan", "collide_dubins")}
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
        out[name] = {"collide_kernel_ms": c_ms / max(c_n, 1), "edges_per_s": m / max(c_ms / max(c_n, 1) * 1e-3, 1e-12),
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
    # same edges against the no-hit ring set: no early exit; every sample segment of a path whose bounding box is too
    # large for the path-level test (pp_path_box_free: most C5 edges) is generated and tested
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
        "config": {"workload": "c5 slice against rings translated outside the (enlarged) bounds: no early exit; every sample of the paths "
                                "too long for the path-level box test is generated and tested",
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
                     "unit": "GB/s", "frac": total * 24.0 / f_s / 1e9 / hbm_peak,
                     "traffic": fact(facts, "pp_dubins_fill_kernel", "dram_bytes")[0],
                     "traffic_source": fact(facts, "pp_dubins_fill_kernel", "dram_bytes")[1],
                     "per_unit": "24 B written per sample (+112 B plan record per path)"},
    }
    # ---- C1: one scalar dubins_path_planning call (examples/dubins, benches/all.rs "Dubins::dubins_path_planning"): a
    # batch of one through the host C-ABI -- one launch writing header + samples into mapped pinned memory, one stream
    # synchronisation.  Expected SLOWER than the CPU's ~2 us: the value of the GPU path is in
    # the batch entry points; this line states what the drop-in costs a caller who keeps calling it one pose at a time.
    try:
        l0 = ctx.launch_count
        c1 = single_path_latency(ctx.dubins_path)
        c1.update({"metric": "dubins_path_latency_us", "gpu_launches": ctx.launch_count - l0,
                   "config": {"workload": "c1: pose (1, 1, 45 deg) -> (-3, -3, -45 deg), radius 1.0, step 0.1, one call "
                                          "at a time through pp_dubins_path (ctypes call included)"}})
        out["c1_single_path"] = c1
    except Exception as e:
        out["c1_single_path"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    # ---- the reference's own criterion benches (benches/all.rs: RRT::plan_one, RRT::plan_10,
    # Dubins::dubins_path_planning on its bench world / pose) through the C++ mirror on this GPU: host/bench_all, a
    # plain timing loop (warm-up 0.2 s, measure 1 s per bench; criterion's are 5 s / 15 s).  Scalar calls, one launch
    # or a few per iteration: latency figures of the drop-in API, not throughput.  Rank 0 only, separate process.
    if rank == 0:
        try:
            import re
            exe = os.path.join(ROOT, "rs-pathplanning_b200", "host", "bench_all")
            r = subprocess.run([exe, "0.2", "1.0"], capture_output=True, text=True, timeout=90)
            rows = {m.group(1): float(m.group(2))
                    for m in re.finditer(r"^(\S+)\s+time:\s+([0-9.]+) us/iter", r.stdout, re.M)}
            if r.returncode != 0 or not rows:
                raise RuntimeError(f"rc {r.returncode}: {(r.stderr or r.stdout)[-160:]}")
            out["benches_all_cpp"] = {"unit": "us/iter", "higher_is_better": False, **rows,
                                      "config": {"workload": "benches/all.rs entry points through host/pathplanning.hpp "
                                                             "(bench world: 6 create_circle obstacles, 21 x 21 bounds, "
                                                             "turn radius 0.8, step 0.1; bench pose: radius 1.0, step 0.1); "
                                                             "the tree persists across iterations, as under criterion"}}
        except Exception as e:
            out["benches_all_cpp"] = {"error": f"{type(e).__name__}: {e}"[:240]}
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
    ap.add_argument("--profile", action="store_true", help="for ncu runs: one pass per step (a number from such a run is never a bench value)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    if args.profile:  # under ncu every launch is replayed ~40 times: one pass per step is enough for a capture
        global PASSES_PER_STEP
        PASSES_PER_STEP = 1
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()

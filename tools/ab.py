"""A/B harness for kernel variants (developer tool, not product).

  python tools/ab.py build occ5="-DPP_EVAL_MIN_BLOCKS=5" poly8="-DPP_POLY_MIN_BLOCKS=8"
      -> build_ab/lib_<name>.so per variant (csrc/Makefile with EXTRA=<flags>), then the default library is rebuilt,
         so the tree ends in its normal state.  Runs on the CPU-only build container.
  gpurun --timeout 300 -- 'python tools/ab.py run > gpurun_out/ab.txt'
      -> one `bench.py --skip-cpu` per library (PP_B200_LIB selects it; the default library first), one table of the
         headline numbers side by side.  ~20 s of GPU time per library.

Every variant must still pass `pytest -m gpu` before it replaces the default: PP_B200_LIB=build_ab/lib_x.so pytest ...
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "rs-pathplanning_b200", "csrc")
OUT = os.path.join(ROOT, "build_ab")

ROWS = [  # (label, path into the bench JSON line, scale, unit)
    ("eval 2^24 pairs (kernel avg)", ("roofline", "kernel_ms_avg"), 1.0, "ms"),
    ("eval far", ("workloads", "dubins_far", "ms_per_step"), 1.0, "ms"),
    ("extend (default)", ("workloads", "extend", "ms_per_step"), 1.0, "ms"),
    ("  nn_grid", ("workloads", "extend", "nn_kernel_ms"), 1.0, "ms"),
    ("  collide grid", ("workloads", "extend", "collide_kernel_ms"), 1.0, "ms"),
    ("extend (PP_COLLIDE_FUSED)", ("workloads", "extend_fused", "ms_per_step"), 1.0, "ms"),
    ("  fused kernel", ("workloads", "extend_fused", "fused_kernel_ms"), 1.0, "ms"),
    ("  binning (3 launches)", ("workloads", "extend_fused", "sort_kernels_ms"), 1.0, "ms"),
    ("extend_scan", ("workloads", "extend_scan", "ms_per_step"), 1.0, "ms"),
    ("extend_dubins", ("workloads", "extend_dubins", "ms_per_step"), 1.0, "ms"),
    ("c5 slice", ("workloads", "dubins_rrt", "ms_per_step"), 1.0, "ms"),
    ("c5 no-hit", ("workloads", "dubins_rrt_nohit", "ms_per_step"), 1.0, "ms"),
    ("sample fill", ("workloads", "dubins_sample", "fill_kernel_ms"), 1.0, "ms"),
    ("e2e eval", ("e2e", "value"), 1e-9, "Gpairs/s"),
    ("sm clock", ("clocks", "sm_mhz"), 1.0, "MHz"),
]


def build(variants):
    os.makedirs(OUT, exist_ok=True)
    for spec in variants:
        name, _, flags = spec.partition("=")
        target = os.path.join(OUT, f"lib_{name}.so")
        print(f"[ab] {name}: EXTRA={flags!r}", flush=True)
        subprocess.run(["make", "-C", CSRC, "-B", f"TARGET={target}", f"EXTRA={flags}"], check=True, capture_output=True)
    print("[ab] rebuilding the default library", flush=True)
    subprocess.run(["make", "-C", CSRC, "-B"], check=True, capture_output=True)


def dig(d, path):
    for k in path:
        if not isinstance(d, dict) or k not in d:
            return None
        d = d[k]
    return d


def run(extra_args):
    libs = [("default", "")] + [(os.path.basename(p)[4:-3], p) for p in sorted(glob.glob(os.path.join(OUT, "lib_*.so")))]
    results = {}
    for name, path in libs:
        env = dict(os.environ)
        if path:
            env["PP_B200_LIB"] = path
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--skip-cpu"] + extra_args,
                           capture_output=True, text=True, env=env, cwd=ROOT)
        line = next((ln for ln in reversed(r.stdout.splitlines()) if ln.startswith("{")), None)
        if r.returncode != 0 or line is None:
            print(f"[ab] {name}: bench failed (rc {r.returncode})\n{r.stderr[-800:]}", flush=True)
            continue
        results[name] = json.loads(line)
    names = list(results)
    print("| workload | " + " | ".join(names) + " |")
    print("|---|" + "---:|" * len(names))
    for label, path, scale, unit in ROWS:
        cells = []
        for n in names:
            v = dig(results[n], path)
            cells.append("-" if v is None else f"{v * scale:.4g}")
        print(f"| {label} ({unit}) | " + " | ".join(cells) + " |")
    return 0 if results else 1


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "build":
        build(sys.argv[2:])
    elif len(sys.argv) >= 2 and sys.argv[1] == "run":
        sys.exit(run(sys.argv[2:]))
    else:
        print(__doc__)
        sys.exit(2)

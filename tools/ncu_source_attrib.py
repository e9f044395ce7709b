#!/usr/bin/env python
"""Stall / instruction attribution per source FUNCTION from an ncu report's source page.

  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_source.csv
  python tools/ncu_source_attrib.py X_source.csv [--md]

Every correlated source line (file, line) is mapped to the function that encloses it in the repo's own sources (the
last `name(` definition at brace depth 0 above the line), and the sampled warp stalls / executed instructions are summed
per function.  Used for the per-routine table in profiles/ (VERDICT r1 #10)."""
import csv
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEF = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:static\s+|inline\s+|__device__\s+|__host__\s+|__forceinline__\s+|__global__\s+|PP_HD\s+|"
                 r"constexpr\s+)*[\w:<>\*&\s]+?\b(\w+)\s*\([^;]*$")


def function_map(path):
    """line number -> enclosing function name (crude but adequate for these sources)"""
    out, cur, depth = {}, "?", 0
    try:
        lines = open(path, errors="ignore").read().split("\n")
    except OSError:
        return {}
    for i, ln in enumerate(lines, 1):
        code = ln.split("//")[0]
        if depth == 0:
            m = DEF.match(code)
            if m and m.group(1) not in ("if", "for", "while", "switch", "return", "asm", "defined"):
                cur = m.group(1)
        out[i] = cur
        depth += code.count("{") - code.count("}")
        depth = max(depth, 0)
    return out


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    agg = defaultdict(lambda: defaultdict(float))
    cur_file, hdr, fmap = None, None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            rel = cur_file.split("/csrc/")[-1] if "/csrc/" in cur_file else os.path.basename(cur_file)
            local = os.path.join(ROOT, "rs-pathplanning_b200", "csrc", rel)
            fmap = function_map(local if os.path.exists(local) else cur_file)
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or r[0] in ("Function Name", "") or not r[0].isdigit():
            continue
        line = int(r[0])
        fn = fmap.get(line, "?")
        key = (os.path.basename(cur_file or "?"), fn)

        def val(name):
            i = hdr.get(name)
            try:
                return float(r[i]) if i is not None and i < len(r) and r[i] not in ("-", "") else 0.0
            except ValueError:
                return 0.0
        a = agg[key]
        a["samples"] += val("# Samples")
        a["inst"] += val("Instructions Executed")
        for s in ("stall_long_sb", "stall_math", "stall_wait", "stall_short_sb", "stall_not_selected", "stall_selected",
                  "stall_dispatch", "stall_no_inst", "stall_branch_resolving", "stall_lg", "stall_mio", "stall_barrier"):
            a[s] += val(s)
    tot_s = sum(a["samples"] for a in agg.values()) or 1.0
    tot_i = sum(a["inst"] for a in agg.values()) or 1.0
    cols = ["stall_math", "stall_wait", "stall_not_selected", "stall_selected", "stall_long_sb", "stall_short_sb", "stall_dispatch",
            "stall_no_inst", "stall_branch_resolving", "stall_lg", "stall_mio"]
    print("| file | function | samples % | warp-instr % | " + " | ".join(c.replace("stall_", "") for c in cols) + " |")
    print("|---|---|---|---|" + "---|" * len(cols))
    for (f, fn), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
        if a["samples"] / tot_s < 0.004 and a["inst"] / tot_i < 0.004:
            continue
        print(f"| {f} | {fn} | {100 * a['samples'] / tot_s:.1f} | {100 * a['inst'] / tot_i:.1f} | " +
              " | ".join(f"{100 * a[c] / max(a['samples'], 1):.0f}" for c in cols) + " |")
    print(f"\ntotal samples {tot_s:.0f}, warp instructions {tot_i:.0f}; stall columns are % of the function's own samples")


if __name__ == "__main__":
    main()

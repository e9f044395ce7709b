#!/usr/bin/env python
"""Extract the per-kernel facts bench.py quotes from `ncu --page raw --csv` exports into profiles/kernel_facts.json.

  python tools/ncu_facts.py profiles/r02_dubins_eval_raw.csv[:units_per_launch] ... -o profiles/kernel_facts.json

For every kernel row: duration, DRAM bytes read / written, L2 sectors, FP64-pipe utilisation, and the number of
FP64-pipe THREAD instructions the launch executed, derived from the pipe counter itself:
    warp instructions on the pipe = pct_of_peak_sustained_active/100 x 2 per cycle per SM x sm__cycles_active.avg x #SM
    thread instructions           = x 32 x (threads per warp instruction on average)
(the FP64 pipe of one sm_100 SM retires 64 lanes = 2 warp instructions per cycle).  With `:units` given (pose pairs,
queries, samples per launch) the per-unit figures are added.  bench.py reads the JSON instead of carrying constants.
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import re


def _num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6,
        "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def facts_from_csv(path, units=None):
    rows = list(csv.reader(open(path)))
    hdr, unit_row, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, scale_units=True):
        i = col.get(name)
        if i is None:
            return None
        v = _num(r[i])
        if v is None:
            return None
        return v * UNIT.get(unit_row[i], 1.0) if scale_units else v

    out = []
    for r in data:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").strip()
        sms = 148
        cyc_active = get(r, "sm__cycles_active.avg", False)
        pct = get(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", False)
        tpi = get(r, "smsp__thread_inst_executed_per_inst_executed.ratio", False) or 32.0
        f = {
            "kernel": name, "source": os.path.relpath(path, os.path.join(os.path.dirname(__file__), "..")),
            "grid": get(r, "launch__grid_size", False), "block": get(r, "launch__block_size", False),
            "registers": get(r, "launch__registers_per_thread", False),
            "duration_us_under_ncu": get(r, "gpu__time_duration.sum"),
            "dram_bytes_read": get(r, "dram__bytes_read.sum"), "dram_bytes_write": get(r, "dram__bytes_write.sum"),
            "l2_sectors": get(r, "lts__t_sectors.sum", False),
            "fp64_pipe_pct_active": pct,
            "issue_pct": get(r, "sm__inst_executed.avg.pct_of_peak_sustained_elapsed", False),
            "warps_active_pct": get(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False),
            "threads_per_inst": tpi,
            "inst_executed": get(r, "smsp__inst_executed.sum", False),
        }
        if f["dram_bytes_read"] is not None and f["dram_bytes_write"] is not None:
            f["dram_bytes"] = f["dram_bytes_read"] + f["dram_bytes_write"]
        if pct is not None and cyc_active is not None:
            f["fp64_thread_instr"] = pct / 100.0 * 2.0 * cyc_active * sms * tpi
        if units:
            f["units_per_launch"] = units
            for k in ("fp64_thread_instr", "dram_bytes", "inst_executed", "l2_sectors"):
                if f.get(k) is not None:
                    f[k + "_per_unit"] = f[k] / units * (32.0 if k == "inst_executed" else 1.0)
        out.append(f)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csvs", nargs="+", help="raw csv[:units_per_launch[:key suffix]]")
    ap.add_argument("-o", "--out", default="profiles/kernel_facts.json")
    a = ap.parse_args()
    facts = {}
    if os.path.exists(a.out):
        facts = json.load(open(a.out))
    for spec in a.csvs:
        path, _, rest = spec.partition(":")
        units, _, suffix = rest.partition(":")
        for f in facts_from_csv(path, float(eval(units)) if units else None):  # noqa: S307 (own command line)
            key = f["kernel"] + (f" [{suffix}]" if suffix else "")
            if key in facts and facts[key].get("source") != f["source"]:
                pass  # newer capture replaces the older one
            facts[key] = f
    json.dump(facts, open(a.out, "w"), indent=1, sort_keys=True)
    print(f"{len(facts)} kernels -> {a.out}")


if __name__ == "__main__":
    main()

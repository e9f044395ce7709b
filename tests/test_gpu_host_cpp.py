"""GPU: the C++ host mirror (rs-pathplanning_b200/host) and the reference's entry points rebuilt on it
(examples/dubins, examples/rrt, benches/all.rs) run through the C-ABI and agree with the oracle."""
import json
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "rs-pathplanning_b200", "host")


def _bin(name):
    p = os.path.join(HOST, name)
    if not os.path.exists(p):
        subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    return p


def test_example_dubins(ctx, O):
    out = subprocess.run([_bin("example_dubins")], capture_output=True, text=True, check=True).stdout
    m = re.findall(r"conf(\d): word=(\w+) cost=([\d.e+-]+) samples=(\d+)", out)
    assert [(w, int(n)) for _, w, _, n in m] == [("LSL", 1508), ("LSR", 1194)]
    assert abs(float(m[0][2]) - 15.074463241942095) < 1e-8 and abs(float(m[1][2]) - 11.93317573386152) < 1e-8


def test_bench_all_entry_points(ctx):
    out = subprocess.run([_bin("bench_all"), "0.05", "0.2", "--batch"], capture_output=True, text=True, check=True).stdout
    for name in ("RRT::plan_one", "RRT::plan_10", "Dubins::dubins_path_planning", "dubins::batch::eval"):
        assert name in out
    assert "cost 9.47540184001621" in out[:2000] or "9.4754018400162" in out
    assert "95 samples" in out


def test_example_rrt_on_reference_world_format(ctx, tmp_path):
    conf = json.load(open(os.path.join(ROOT, "tests", "golden", "transit_world.json")))
    world = {"bounds": list(map(list, zip(conf["bounds_x"], conf["bounds_y"]))),
             "obstacles": [list(map(list, zip(r["x"], r["y"]))) for r in conf["rings"]],
             "path": [], "start": conf["start"], "goal": conf["goal"]}  # examples/rrt/src/main.rs:13-20
    p = tmp_path / "world.json"
    p.write_text(json.dumps(world))
    out = subprocess.run([_bin("example_rrt"), str(p), "300", "7"], capture_output=True, text=True, check=True).stdout
    assert "bounds 20 pts, 3 obstacles" in out
    nodes = int(re.search(r"tree nodes: (\d+)", out).group(1))
    assert nodes > 1
    if "Path generated!" in out:
        assert "verify(path) = 1" in out
    # the round-based planner of the C++ mirror (plan_rounds + check_finish_many) and its node-by-node self-check
    out = subprocess.run([_bin("example_rrt"), str(p), "2000", "7", "256"], capture_output=True, text=True,
                         check=True).stdout
    m = re.search(r"check_finish_many agrees on (\d+) of (\d+) nodes", out)
    assert m and m.group(1) == m.group(2) and int(m.group(2)) > 5, out[-400:]
    assert int(re.search(r"tree nodes: (\d+)", out).group(1)) > 50
    if "Path generated!" in out:
        assert "verify(path) = 1" in out


def test_example_group_drives_every_device_from_one_process(ctx):
    """a compiled host using pp_group: replication by broadcast (upload + 512-node append), sliced extend steps,
    byte-identical to one context for 1, 2, 4 ... devices (1 on a single-GPU box)"""
    out = subprocess.run([_bin("example_group"), str(1 << 16)], capture_output=True, text=True, check=True).stdout
    assert "group ok" in out and "DIFFERS" not in out
    assert "group of 1 device(s): extend byte-identical, extend_dubins byte-identical" in out

"""CPU: the C++ host mirror (rs-pathplanning_b200/host/pathplanning.hpp) and its examples, linked against a TEST
DOUBLE of the C-ABI (tests/fake_abi/fake_pp.cpp, answered by the CPU oracle).  Covers the mirror's host logic -- node
graph, tree slots, batched optimize / check_finish, plan_rounds, the JSON world reader -- without a device.  The GPU
suite runs the same binaries against the real library (tests/test_gpu_host_cpp.py)."""
import json
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "rs-pathplanning_b200", "host")


@pytest.fixture(scope="module")
def fake_bins(O, tmp_path_factory):
    """example_dubins / example_rrt built against the fake ABI (never against libpathplanning_b200.so)"""
    out = tmp_path_factory.mktemp("fake_abi")
    oracle_dir = os.path.join(ROOT, "oracle")
    bins = {}
    for name in ("example_dubins", "example_rrt"):
        exe = str(out / name)
        cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", oracle_dir, "-I", HOST,
               os.path.join(HOST, name + ".cpp"), os.path.join(ROOT, "tests", "fake_abi", "fake_pp.cpp"),
               "-o", exe, "-L", oracle_dir, "-lpp_oracle", "-Wl,-rpath," + oracle_dir, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
        assert "pathplanning_b200" not in ldd and "cudart" not in ldd
        bins[name] = exe
    return bins


def test_example_dubins_known_answers(fake_bins):
    out = subprocess.run([fake_bins["example_dubins"]], capture_output=True, text=True, check=True).stdout
    m = re.findall(r"conf(\d): word=(\w+) cost=([\d.e+-]+) samples=(\d+)", out)
    assert [(w, int(n)) for _, w, _, n in m] == [("LSL", 1508), ("LSR", 1194)]
    assert abs(float(m[0][2]) - 15.074463241942095) < 1e-12 and abs(float(m[1][2]) - 11.93317573386152) < 1e-12


def _world(tmp_path):
    conf = json.load(open(os.path.join(ROOT, "tests", "golden", "transit_world.json")))
    world = {"bounds": list(map(list, zip(conf["bounds_x"], conf["bounds_y"]))),
             "obstacles": [list(map(list, zip(r["x"], r["y"]))) for r in conf["rings"]],
             "path": [], "start": conf["start"], "goal": conf["goal"]}  # examples/rrt/src/main.rs:13-20
    p = tmp_path / "world.json"
    p.write_text(json.dumps(world))
    return str(p)


def test_example_rrt_plan_loop(fake_bins, tmp_path):
    out = subprocess.run([fake_bins["example_rrt"], _world(tmp_path), "150", "7"], capture_output=True, text=True,
                         check=True, timeout=600).stdout
    assert "bounds 20 pts, 3 obstacles" in out
    assert int(re.search(r"tree nodes: (\d+)", out).group(1)) > 1
    if "Path generated!" in out:
        assert "verify(path) = 1" in out


def test_example_rrt_round_planner_and_batched_goal_check(fake_bins, tmp_path):
    out = subprocess.run([fake_bins["example_rrt"], _world(tmp_path), "600", "7", "100"], capture_output=True, text=True,
                         check=True, timeout=600).stdout
    m = re.search(r"check_finish_many agrees on (\d+) of (\d+) nodes", out)
    assert m and m.group(1) == m.group(2) and int(m.group(2)) > 5, out[-400:]
    assert int(re.search(r"tree nodes: (\d+)", out).group(1)) > 20
    if "Path generated!" in out:
        assert "verify(path) = 1" in out

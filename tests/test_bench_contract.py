"""CPU: the parts of bench.py that run without a GPU -- the reference arm (`--impl reference`: the oracle port on the
host cores), the cpu_baseline leg of the default run, and the own arm's refusal to run without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "dubins_pairs_per_s" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 1e4 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "2^24-pair" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_never_loads_the_product_library():
    """hygiene: the reference arm must not map libpathplanning_b200.so -- with the library path pointing
    nowhere the product package cannot even be imported, and the arm must still run; both arms print the same `config`"""
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"], env={"PP_B200_LIB": "/nonexistent/libpp.so"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.C3_CONFIG and d["step"]["pairs_per_step"] == 1 << 24


def test_reference_arm_other_ranks_exit_quietly():
    """under torchrun (N > 1) rank 0 alone runs and prints; the other ranks exit 0 without work"""
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_own_arm_fails_loudly_without_a_gpu(pp):
    """no CPU fallback: without an sm_100 device the bench must die, not print a number"""
    if pp.device_count() > 0:
        return  # on the GPU box the driver runs the real thing
    r = _run(["--steps", "1", "--warmup", "1", "--skip-secondary", "--skip-cpu"])
    assert r.returncode != 0
    assert not any(ln.lstrip().startswith("{") for ln in r.stdout.splitlines())


def test_cpu_baseline_leg_has_the_contract_keys(pp):
    sys.path.insert(0, ROOT)
    import bench

    cb = bench.cpu_baseline_leg(pp)
    assert cb["kind"] == "port" and cb["unit"] == "pairs/s" and cb["cores"] >= 1 and cb["value"] > cb["single_thread_value"] / 2
    assert "sample" in cb and cb["single_thread_value"] > 1e4
    ext = cb["extend"]
    assert ext["unit"] == "steps/s" and ext["value"] > 0 and ext["nn_brute_matches_grid"] is True
    assert ext["nn_grid_queries_per_s"] > ext["nn_brute_queries_per_s"] > 0
    assert 0 < ext["single_thread_value"] <= ext["value"] * 2


def test_kernel_facts_point_at_committed_captures():
    """bench.py quotes per-kernel figures (executed FP64 instructions per pair, DRAM bytes per launch) from
    profiles/kernel_facts.json: every entry must name an ncu raw-page export that is committed under profiles/, and
    re-extracting that file with tools/ncu_facts.py must reproduce the figures bench.py reads"""
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    facts = json.load(open(os.path.join(root, "profiles", "kernel_facts.json")))
    assert "pp_dubins_eval_kernel<0, 0>" in facts and "pp_dubins_fill_kernel" in facts
    spec = importlib.util.spec_from_file_location("ncu_facts", os.path.join(root, "tools", "ncu_facts.py"))
    nf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(nf)
    for key, f in facts.items():
        path = os.path.join(root, f["source"])
        assert os.path.exists(path), (key, f["source"])
        again = [g for g in nf.facts_from_csv(path, f.get("units_per_launch")) if g["kernel"] == f["kernel"]]
        assert any(abs(g["inst_executed"] - f["inst_executed"]) < 0.5 and
                   abs(g["dram_bytes"] - f["dram_bytes"]) < 0.5 for g in again), key
    ev = facts["pp_dubins_eval_kernel<0, 0>"]
    assert 400 < ev["fp64_thread_instr_per_unit"] < 600 and 0.9e9 < ev["dram_bytes"] < 1.0e9

"""CPU: accuracy of the hand-written f64 sincos / atan2 / acos (rs-pathplanning_b200/csrc/pp_math.cuh), compiled
as plain C++ and compared with glibc's long-double functions on 10^6 samples (tools/math_accuracy.cpp).
This is product code under test (the same header the kernels include), not the oracle."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_math_kernels_accuracy(tmp_path):
    exe = str(tmp_path / "math_accuracy")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off",
                    os.path.join(ROOT, "tools", "math_accuracy.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe, "1000000"], capture_output=True, text=True, check=True).stdout
    v = {k: float(x) for k, x in (ln.split() for ln in out.strip().splitlines())}
    assert v["sin_ulp"] < 2.0 and v["cos_ulp"] < 2.0 and v["sincos_big_ulp"] < 2.0, v
    assert v["atan2_ulp"] < 2.0 and v["acos_err"] < 2.5, v
    assert v["batch_mismatch"] == 0 and v["special_bad"] == 0, v


def test_division_free_predicates_are_exact(tmp_path):
    """geo_predicates.cuh decides geo's `0 <= a/b <= 1` tests without dividing (pp_quot_in01): same answers as the IEEE
    quotient on random, near-equal, tiny / huge / zero / non-finite operands, and the ring predicates built on it agree
    with a restatement that performs every division (tools/predicate_check.cpp)"""
    exe = str(tmp_path / "predicate_check")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off",
                    os.path.join(ROOT, "tools", "predicate_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe, "2000000"], capture_output=True, text=True, check=True).stdout
    v = {k: int(x) for k, x in (ln.split() for ln in out.strip().splitlines())}
    assert v == {"bad_random": 0, "bad_close": 0, "bad_special": 0, "bad_ring": 0}, v


def test_sample_loop_fast_forward_is_exact(tmp_path):
    """pp_replay.cuh jumps through generate_local_course's `while pd.abs() <= l.abs() { pd += d }` loop
    (src/dubins.rs:239-255) binade by binade: iteration count and final pd must equal the literal loop's, bit for bit,
    on the reference's call pattern, arbitrary steps of either sign, limits placed on iterates / binade tops / their
    neighbours, tie binades, steps that vanish against pd, zeros, subnormals, infinities and NaN
    (tools/replay_check.cpp)"""
    exe = str(tmp_path / "replay_check")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off",
                    os.path.join(ROOT, "tools", "replay_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe, "400000"], capture_output=True, text=True, check=True).stdout
    v = {k: int(x) for k, x in (ln.split() for ln in out.strip().splitlines())}
    assert v == {"bad_call_pattern": 0, "bad_steps": 0, "bad_random": 0, "bad_special": 0, "bad_tiny": 0}, v


def test_tables_are_reproducible(tmp_path):
    """the committed coefficient tables are what tools/gen_math_tables.py generates"""
    import shutil
    inc = os.path.join(ROOT, "rs-pathplanning_b200", "csrc", "pp_math_tables.inc")
    keep = str(tmp_path / "tables.inc")
    shutil.copy(inc, keep)
    try:
        subprocess.run(["python", os.path.join(ROOT, "tools", "gen_math_tables.py")], check=True, capture_output=True)
        assert open(inc).read() == open(keep).read()
    finally:
        shutil.copy(keep, inc)

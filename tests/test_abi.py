"""CPU: the C-ABI library loads, exports every symbol include/pathplanning_b200.h declares, the Python
binding covers exactly that set, and the product path neither falls back to the CPU nor touches oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pathplanning_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(pp):
    names = header_functions()
    assert len(names) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", pp._ffi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (pp_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared but not exported: {missing}"


def test_binding_covers_header(pp):
    assert sorted(pp._ffi.SIGNATURES) == header_functions()
    for n in pp._ffi.SIGNATURES:
        assert getattr(pp._ffi.lib, n) is not None


def test_constants_match_header(pp):
    src = open(HEADER).read()
    assert int(re.search(r"#define PP_DUBINS_PLAN_BYTES (\d+)", src).group(1)) == pp._ffi.PLAN_BYTES
    assert int(re.search(r"#define PP_ABI_VERSION (\d+)", src).group(1)) == pp._ffi.lib.pp_abi_version()
    assert pp._ffi.lib.pp_status_string(-2).decode().startswith("no sm_100")
    assert pp._ffi.WORDS == ("LSL", "RSR", "LSR", "RSL", "RLR", "LRL")  # ALL_PLANNERS order, src/dubins.rs:291
    # method flags: the binding's constants are the header's enumerators
    enums = dict(re.findall(r"\b(PP_(?:NN|COLLIDE)_[A-Z0-9_]+)\s*=\s*(\d+)", src))
    for name, value in enums.items():
        assert getattr(pp._ffi, name[3:]) == int(value), name
    assert {"PP_NN_DEFAULT", "PP_NN_GRID", "PP_NN_SCAN", "PP_COLLIDE_DEFAULT", "PP_COLLIDE_SCAN"} <= set(enums)


def test_no_cpu_fallback(pp):
    h = ctypes.c_void_p()
    assert pp._ffi.lib.pp_ctx_create(0, None) == pp._ffi.PP_ERR_INVALID
    if pp.device_count() == 0:
        assert pp._ffi.lib.pp_ctx_create(0, ctypes.byref(h)) == pp._ffi.PP_ERR_NO_DEVICE and not h.value
        with pytest.raises(pp.PathPlanningError):
            pp.Context(0)
        with pytest.raises(pp.PathPlanningError):  # the scalar drop-in API fails loudly too
            pp.dubins.mod2pi(1.0)
    assert pp._ffi.lib.pp_host_free(None) == 0
    assert pp._ffi.lib.pp_launch_count(None) == 0 and pp._ffi.lib.pp_tree_size(None) == 0


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "rs-pathplanning_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".rs")) or f == "Makefile":
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"pp_oracle|libpp_oracle|from oracle|import oracle|oracle/", txt):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
    out = subprocess.run(["ldd", os.path.join(pkg, "libpathplanning_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libcudart" in out


def test_kernels_are_sm_100a_with_tma(pp):
    """the scan kernels really use the TMA engine: UBLKCP in the SASS of the built library"""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    r = subprocess.run([cuobjdump, "-sass", pp._ffi.LIB_PATH], capture_output=True, text=True)
    assert "sm_100a" in r.stdout
    assert "UBLKCP" in r.stdout and "SYNCS" in r.stdout
    assert r.stdout.count("DFMA") > 100

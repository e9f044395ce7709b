"""CPU: static checks of the source-only Rust shim (there is no rustc in this image, so nothing here compiles it).
The extern "C" block of rust/src/ffi.rs must declare header functions only, with the header's parameter list
(count, order, pointer-ness, constness and scalar width), and the constants it copies must be the header's."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pathplanning_b200.h")
RUST = os.path.join(ROOT, "rs-pathplanning_b200", "rust", "src")

C_SCALARS = {"int": "c_int", "double": "c_double", "size_t": "usize", "uint8_t": "u8", "uint32_t": "u32",
             "uint64_t": "u64", "int32_t": "i32", "void": "c_void", "char": "c_char", "pp_ctx": "pp_ctx",
             "pp_group": "pp_group"}


def _split_args(s):
    return [a.strip() for a in s.split(",") if a.strip() and a.strip() != "void"]


def header_prototypes():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"([A-Za-z_][\w \*]*?)\b(pp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        sig = []
        for a in _split_args(args):
            const = "const " in a
            stars = a.count("*")
            base = re.sub(r"\bconst\b", "", a).replace("*", " ").split()[0]
            sig.append(("*const " if const else "*mut ") * stars + C_SCALARS[base])
        protos[name] = sig
    return protos


def rust_externs():
    src = open(os.path.join(RUST, "ffi.rs")).read()
    block = re.search(r'extern "C" \{(.*?)\n\}', src, flags=re.S).group(1)
    out = {}
    for name, args in re.findall(r"pub fn (pp_[a-z0-9_]+)\s*\(([^)]*)\)", block):
        out[name] = [re.sub(r"\s+", " ", a.split(":", 1)[1].strip()) for a in _split_args(args)]
    return out


def test_extern_block_matches_the_header():
    protos, externs = header_prototypes(), rust_externs()
    assert len(protos) >= 40 and len(externs) >= 20
    for name, sig in externs.items():
        assert name in protos, f"{name} is not declared in the header"
        assert sig == protos[name], f"{name}: rust {sig} != header {protos[name]}"


def test_every_ffi_call_in_the_shim_is_declared():
    externs = rust_externs()
    for f in ("dubins.rs", "rrt.rs", "group.rs"):
        src = open(os.path.join(RUST, f)).read()
        for name in set(re.findall(r"ffi::(pp_[a-z0-9_]+)\s*\(", src)):
            assert name in externs, f"{f} calls ffi::{name}, which ffi.rs does not declare"


def test_copied_constants_match_the_header():
    hdr = open(HEADER).read()
    ffi = open(os.path.join(RUST, "ffi.rs")).read()
    plan_bytes = int(re.search(r"#define PP_DUBINS_PLAN_BYTES (\d+)", hdr).group(1))
    assert int(re.search(r"PP_DUBINS_PLAN_BYTES: usize = (\d+)", ffi).group(1)) == plan_bytes
    assert int(re.search(r"ABI version (\d+)", ffi).group(1)) == int(re.search(r"#define PP_ABI_VERSION (\d+)", hdr).group(1))
    assert re.search(r"PP_WORD_NONE: c_int = 0xFF", ffi) and re.search(r"PP_WORD_NONE\s*=\s*(0xFF|255)", hdr)
    assert re.search(r"PP_ERR_OVERFLOW: c_int = (-\d+)", ffi).group(1) == re.search(r"PP_ERR_OVERFLOW\s*=\s*(-\d+)", hdr).group(1)


def test_sources_are_balanced_and_mirror_the_reference_api():
    for f in os.listdir(RUST):
        src = open(os.path.join(RUST, f)).read()
        code = re.sub(r"//[^\n]*", "", src)
        code = re.sub(r'"(?:\\.|[^"\\])*"', '""', code)
        for o, c in ("{}", "()", "[]"):
            assert code.count(o) == code.count(c), f"{f}: unbalanced {o}{c}"
    rrt = open(os.path.join(RUST, "rrt.rs")).read()
    # the reference's public surface (src/rrt.rs) plus the batched forms of SURVEY 8f
    for fn in ("create_circle", "line_to_origin", "verify", "rand_point", "get_nearest_node", "get_random_node",
               "verify_node", "check_finish", "optimize", "optimize_from_goal", "finalize", "plan_one", "plan",
               "verify_many", "optimize_many", "check_finish_many", "plan_rounds"):
        assert re.search(r"pub fn %s\b" % fn, rrt), fn
    dub = open(os.path.join(RUST, "dubins.rs")).read()
    for fn in ("mod2pi", "pi_2_pi", "dubins_path_planning", "dubins_path_planning_from_origin"):
        assert re.search(r"pub fn %s\b" % fn, dub), fn

"""ctypes binding of the CPU oracle (oracle/libpp_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.  Parity is UNPINNED by the reference (no tests upstream,
no rustc here); see oracle/pp_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpp_oracle.so")

WORDS = ("LSL", "RSR", "LSR", "RSL", "RLR", "LRL")
NONE = 0xFF
FLAG_NEAR_WRAP, FLAG_NEAR_TIE, FLAG_NEAR_FEAS, FLAG_NEAR_COUNT, FLAG_NEAR_GRAZE, FLAG_HUGE_ANGLE = 1, 2, 4, 8, 16, 32
GRAZE_TOL = 1e-8  # margin below GRAZE_TOL * max(1, |coordinates|) => a 1e-9-accurate implementation may flip the verdict

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("pp_oracle.c", "pp_oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    r = subprocess.run(["make", "-C", _HERE, "-B"], capture_output=True, text=True)
    if r.returncode != 0:
        # fall back to a build without OpenMP (single-threaded baseline) rather than no oracle at all
        r2 = subprocess.run(["gcc", "-O3", "-std=c11", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared",
                             "-o", _LIB_PATH, os.path.join(_HERE, "pp_oracle.c"), "-lm"],
                            capture_output=True, text=True)
        if r2.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stderr + r2.stderr)
    return _LIB_PATH


class World(C.Structure):
    _fields_ = [("bx", _dp), ("by", _dp), ("nb", C.c_size_t), ("ox", _dp), ("oy", _dp),
                ("ring_off", _u32p), ("n_rings", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    d, sz = C.c_double, C.c_size_t
    L.ppo_mod2pi.restype = d
    L.ppo_mod2pi.argtypes = [d]
    L.ppo_pi_2_pi.restype = d
    L.ppo_pi_2_pi.argtypes = [d]
    L.ppo_dubins_word.restype = C.c_int
    L.ppo_dubins_word.argtypes = [C.c_int, d, d, d, _dp]
    L.ppo_dubins_eval.restype = C.c_int
    L.ppo_dubins_eval.argtypes = [d] * 7 + [_dp, _dp, _u32p]
    L.ppo_dubins_eval_batch.restype = None
    L.ppo_dubins_eval_batch.argtypes = [sz] + [_dp] * 7 + [d, _dp, _u8p, _dp, _u32p, C.c_int]
    L.ppo_dubins_path.restype = C.c_long
    L.ppo_dubins_path.argtypes = [d] * 8 + [C.c_int, _dp, _dp, _dp, sz, C.POINTER(C.c_int), _dp,
                                            C.POINTER(C.c_long)]
    L.ppo_dubins_word_margin.restype = d
    L.ppo_dubins_word_margin.argtypes = [C.c_int, d, d, d]
    L.ppo_dubins_path_flags.restype = C.c_long
    L.ppo_dubins_path_flags.argtypes = [d] * 8 + [C.c_int, _dp, _dp, _dp, sz, C.POINTER(C.c_int), _dp, _u32p]
    L.ppo_dubins_count_batch.restype = None
    L.ppo_dubins_count_batch.argtypes = [sz] + [_dp] * 6 + [d, d, _i64p, C.c_int]
    L.ppo_create_circle.restype = C.c_long
    L.ppo_create_circle.argtypes = [d, d, d, _dp, _dp, sz]
    L.ppo_compute_yaw.restype = d
    L.ppo_compute_yaw.argtypes = [d] * 4
    L.ppo_nn_brute.restype = None
    L.ppo_nn_brute.argtypes = [sz, _dp, _dp, sz, _dp, _dp, _u32p, _dp, C.POINTER(sz), C.c_int]
    L.ppo_nn_grid.restype = None
    L.ppo_nn_grid.argtypes = [sz, _dp, _dp, sz, _dp, _dp, _u32p, _dp, C.c_int]
    L.ppo_ring_has_point.restype = C.c_int
    L.ppo_ring_has_point.argtypes = [_dp, _dp, sz, d, d]
    L.ppo_point_position.restype = C.c_int
    L.ppo_point_position.argtypes = [_dp, _dp, sz, d, d]
    L.ppo_lines_intersect.restype = C.c_int
    L.ppo_lines_intersect.argtypes = [_dp, _dp, sz, _dp, _dp, sz]
    wp = C.POINTER(World)
    L.ppo_verify.restype = C.c_int
    L.ppo_verify.argtypes = [wp, _dp, _dp, sz]
    L.ppo_verify_culled.restype = C.c_int
    L.ppo_verify_culled.argtypes = [wp, _dp, _dp, sz]
    L.ppo_circle_class.restype = C.c_int
    L.ppo_circle_class.argtypes = [_dp, _dp, sz, d, d, d, d]
    L.ppo_verify_segments.restype = None
    L.ppo_verify_segments.argtypes = [wp, sz, _dp, _dp, _dp, _dp, _u8p, C.c_int, C.c_int]
    L.ppo_dubins_edge_polyline.restype = C.c_long
    L.ppo_dubins_edge_polyline.argtypes = [d] * 8 + [_dp, _dp, sz]
    L.ppo_verify_dubins_edges.restype = None
    L.ppo_verify_dubins_edges.argtypes = [wp, sz] + [_dp] * 6 + [d, d, _u8p, C.c_int, C.c_int]
    L.ppo_verify_margin.restype = C.c_int
    L.ppo_verify_margin.argtypes = [wp, _dp, _dp, sz, _dp]
    L.ppo_verify_dubins_edges_flags.restype = None
    L.ppo_verify_dubins_edges_flags.argtypes = [wp, sz] + [_dp] * 6 + [d, d, d, _u8p, _u32p, _dp, C.c_int, C.c_int]
    L.ppo_optimize.restype = C.c_long
    L.ppo_optimize.argtypes = [wp, sz, _dp, _dp, _dp, _i32p, C.c_uint32, d, d, d, _dp, _dp, _dp, sz, _u32p,
                               C.POINTER(C.c_long)]
    L.ppo_check_finish.restype = C.c_long
    L.ppo_check_finish.argtypes = [wp, sz, _dp, _dp, _dp, _i32p, C.c_uint32, d, d, d, d, d, d, _dp, _dp, sz,
                                   C.POINTER(C.c_long), _dp, _dp, _dp, sz, C.POINTER(C.c_long), _u32p, _u32p,
                                   C.POINTER(C.c_long)]
    L.ppo_line_to_origin.restype = C.c_long
    L.ppo_line_to_origin.argtypes = [_dp, _dp, _dp, _i32p, C.c_uint32, d, d, _dp, _dp, sz]
    L.ppo_uniform.restype = d
    L.ppo_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    L.ppo_fill_uniform.restype = None
    L.ppo_fill_uniform.argtypes = [C.c_uint64, C.c_uint64, sz, d, d, _dp]
    L.ppo_max_threads.restype = C.c_int
    _lib = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=_dp):
    return a.ctypes.data_as(t)


def max_threads() -> int:
    return int(lib().ppo_max_threads())


def mod2pi(x: float) -> float:
    return float(lib().ppo_mod2pi(x))


def pi_2_pi(x: float) -> float:
    return float(lib().ppo_pi_2_pi(x))


def dubins_word(word: int, alpha: float, beta: float, d: float):
    o = (C.c_double * 3)()
    ok = lib().ppo_dubins_word(word, alpha, beta, d, o)
    return (o[0], o[1], o[2]) if ok else None


def dubins_eval(sx, sy, syaw, ex, ey, eyaw, radius):
    cost = C.c_double()
    o = (C.c_double * 3)()
    fl = C.c_uint32()
    w = lib().ppo_dubins_eval(sx, sy, syaw, ex, ey, eyaw, radius, C.byref(cost), o, C.byref(fl))
    return w, cost.value, (o[0], o[1], o[2]), fl.value


def dubins_eval_batch(sx, sy, syaw, ex, ey, eyaw, radius=1.0, radius_arr=None, nthreads=0, want_flags=True):
    sx, sy, syaw, ex, ey, eyaw = map(_f64, (sx, sy, syaw, ex, ey, eyaw))
    n = sx.size
    cost = np.empty(n, np.float64)
    word = np.empty(n, np.uint8)
    tpq = np.empty((n, 3), np.float64)
    flags = np.zeros(n, np.uint32) if want_flags else None
    ra = _f64(radius_arr) if radius_arr is not None else None
    lib().ppo_dubins_eval_batch(n, _p(sx), _p(sy), _p(syaw), _p(ex), _p(ey), _p(eyaw),
                                _p(ra) if ra is not None else None, float(radius), _p(cost), _p(word, _u8p),
                                _p(tpq), _p(flags, _u32p) if want_flags else None, int(nthreads))
    return cost, word, tpq, flags


@dataclass
class Path:
    x: np.ndarray
    y: np.ndarray
    yaw: np.ndarray
    word: int
    cost: float
    n_point: int


def dubins_path(sx, sy, syaw, ex, ey, eyaw, radius, step, from_origin=False):
    """dubins_path_planning (or _from_origin): returns Path or None (no feasible word)."""
    cap = 4096
    while True:
        px, py, pyaw = (np.empty(cap, np.float64) for _ in range(3))
        w, c, npnt = C.c_int(), C.c_double(), C.c_long()
        n = lib().ppo_dubins_path(sx, sy, syaw, ex, ey, eyaw, radius, step, int(from_origin), _p(px), _p(py),
                                  _p(pyaw), cap, C.byref(w), C.byref(c), C.byref(npnt))
        if n == -2:
            cap *= 4
            continue
        if n == -1:
            return None
        if n < 0:
            raise RuntimeError("reference would panic (index out of n_point buffer)")
        return Path(px[:n].copy(), py[:n].copy(), pyaw[:n].copy(), w.value, c.value, npnt.value)


def dubins_word_margin(word: int, alpha: float, beta: float, d: float) -> float:
    return float(lib().ppo_dubins_word_margin(word, alpha, beta, d))


def dubins_path_flags(sx, sy, syaw, ex, ey, eyaw, radius, step, from_origin=False):
    """(sample count or -1 for None, harness flags) of one path"""
    cap = 4096
    while True:
        px, py = np.empty(cap), np.empty(cap)
        fl = C.c_uint32()
        n = lib().ppo_dubins_path_flags(sx, sy, syaw, ex, ey, eyaw, radius, step, int(from_origin), _p(px), _p(py),
                                        None, cap, None, None, C.byref(fl))
        if n == -2:
            cap *= 4
            continue
        return int(n), int(fl.value)


def dubins_count_batch(sx, sy, syaw, ex, ey, eyaw, radius, step, nthreads=0):
    sx, sy, syaw, ex, ey, eyaw = map(_f64, (sx, sy, syaw, ex, ey, eyaw))
    n = sx.size
    counts = np.empty(n, np.int64)
    lib().ppo_dubins_count_batch(n, _p(sx), _p(sy), _p(syaw), _p(ex), _p(ey), _p(eyaw), float(radius),
                                 float(step), _p(counts, _i64p), int(nthreads))
    return counts


def create_circle(cx, cy, radius):
    cap = int(np.ceil(2 * np.pi * radius)) + 8
    rx, ry = np.empty(cap), np.empty(cap)
    n = lib().ppo_create_circle(cx, cy, radius, _p(rx), _p(ry), cap)
    if n < 0:
        raise ValueError("bad circle")
    return rx[:n].copy(), ry[:n].copy()


def compute_yaw(fx, fy, tx, ty):
    return float(lib().ppo_compute_yaw(fx, fy, tx, ty))


def nn_brute(nx, ny, qx, qy, nthreads=0, check_hypot=False):
    nx, ny, qx, qy = map(_f64, (nx, ny, qx, qy))
    m = qx.size
    idx = np.empty(m, np.uint32)
    d2 = np.empty(m, np.float64)
    dis = C.c_size_t(0)
    lib().ppo_nn_brute(nx.size, _p(nx), _p(ny), m, _p(qx), _p(qy), _p(idx, _u32p), _p(d2),
                       C.byref(dis) if check_hypot else None, int(nthreads))
    return (idx, d2, dis.value) if check_hypot else (idx, d2)


def nn_grid(nx, ny, qx, qy, nthreads=0):
    nx, ny, qx, qy = map(_f64, (nx, ny, qx, qy))
    m = qx.size
    idx = np.empty(m, np.uint32)
    d2 = np.empty(m, np.float64)
    lib().ppo_nn_grid(nx.size, _p(nx), _p(ny), m, _p(qx), _p(qy), _p(idx, _u32p), _p(d2), int(nthreads))
    return idx, d2


def close_ring(x, y):
    """Polygon::new's ring closing (geo-types 0.4): append the first point if last != first."""
    x, y = _f64(x).copy(), _f64(y).copy()
    if x.size and (x[0] != x[-1] or y[0] != y[-1]):
        x = np.append(x, x[0])
        y = np.append(y, y[0])
    return x, y


class OracleWorld:
    """bounds ring + obstacle rings, as Space{bounds, obstacles} after Space::new (no inflation here)."""

    def __init__(self, bounds_xy, rings_xy):
        bx, by = close_ring(*bounds_xy)
        self.bx, self.by = bx, by
        xs, ys, off = [], [], [0]
        for (rx, ry) in rings_xy:
            rx, ry = close_ring(rx, ry)
            xs.append(rx)
            ys.append(ry)
            off.append(off[-1] + rx.size)
        self.ox = np.concatenate(xs) if xs else np.zeros(0)
        self.oy = np.concatenate(ys) if ys else np.zeros(0)
        self.off = np.asarray(off, np.uint32)
        self.w = World(_p(self.bx), _p(self.by), self.bx.size, _p(self.ox), _p(self.oy), _p(self.off, _u32p),
                       len(rings_xy))

    def rings(self):
        return [(self.ox[self.off[i]:self.off[i + 1]], self.oy[self.off[i]:self.off[i + 1]])
                for i in range(len(self.off) - 1)]

    def verify(self, lx, ly, culled=False):
        lx, ly = _f64(lx), _f64(ly)
        f = lib().ppo_verify_culled if culled else lib().ppo_verify
        return bool(f(C.byref(self.w), _p(lx), _p(ly), lx.size))

    def verify_margin(self, lx, ly):
        """(verdict, margin): how far the vertices may move without changing the verdict (harness, not reference)"""
        lx, ly = _f64(lx), _f64(ly)
        m = C.c_double()
        v = lib().ppo_verify_margin(C.byref(self.w), _p(lx), _p(ly), lx.size, C.byref(m))
        return bool(v), m.value

    def verify_dubins_edges_flags(self, sx, sy, syaw, ex, ey, eyaw, radius, step, graze_tol=GRAZE_TOL, culled=False,
                                  nthreads=0):
        """(ok, flags, margins): verdicts with the classification flags of the parity harness"""
        sx, sy, syaw, ex, ey, eyaw = map(_f64, (sx, sy, syaw, ex, ey, eyaw))
        ok = np.empty(sx.size, np.uint8)
        flags = np.zeros(sx.size, np.uint32)
        margins = np.empty(sx.size, np.float64)
        lib().ppo_verify_dubins_edges_flags(C.byref(self.w), sx.size, _p(sx), _p(sy), _p(syaw), _p(ex), _p(ey),
                                            _p(eyaw), float(radius), float(step), float(graze_tol), _p(ok, _u8p),
                                            _p(flags, _u32p), _p(margins), int(culled), int(nthreads))
        return ok, flags, margins

    def verify_segments(self, ax, ay, bx, by, culled=False, nthreads=0):
        ax, ay, bx, by = map(_f64, (ax, ay, bx, by))
        ok = np.empty(ax.size, np.uint8)
        lib().ppo_verify_segments(C.byref(self.w), ax.size, _p(ax), _p(ay), _p(bx), _p(by), _p(ok, _u8p),
                                  int(culled), int(nthreads))
        return ok

    def verify_dubins_edges(self, sx, sy, syaw, ex, ey, eyaw, radius, step, culled=False, nthreads=0):
        sx, sy, syaw, ex, ey, eyaw = map(_f64, (sx, sy, syaw, ex, ey, eyaw))
        ok = np.empty(sx.size, np.uint8)
        lib().ppo_verify_dubins_edges(C.byref(self.w), sx.size, _p(sx), _p(sy), _p(syaw), _p(ex), _p(ey), _p(eyaw),
                                      float(radius), float(step), _p(ok, _u8p), int(culled), int(nthreads))
        return ok


def _tree(nx, ny, nyaw, parent):
    return _f64(nx), _f64(ny), _f64(nyaw), np.ascontiguousarray(parent, np.int32)


def optimize(world, nx, ny, nyaw, parent, node, radius, step, graze_tol=GRAZE_TOL):
    """RRT::optimize(node, 0) over a flat tree: (chain poses [(x, y, yaw)] new node -> root or None, flags, verifies)"""
    nx, ny, nyaw, parent = _tree(nx, ny, nyaw, parent)
    cap = 4096
    while True:
        cx, cy, cyaw = np.empty(cap), np.empty(cap), np.empty(cap)
        fl, nv = C.c_uint32(), C.c_long()
        n = lib().ppo_optimize(C.byref(world.w), nx.size, _p(nx), _p(ny), _p(nyaw), _p(parent, _i32p), int(node),
                               radius, step, graze_tol, _p(cx), _p(cy), _p(cyaw), cap, C.byref(fl), C.byref(nv))
        if n == -2:
            cap *= 4
            continue
        if n < 0:
            raise RuntimeError("oracle optimize failed")
        chain = None if n == 0 else np.stack([cx[:n], cy[:n], cyaw[:n]], 1)
        return chain, int(fl.value), int(nv.value)


@dataclass
class Finish:
    line: tuple       # (x, y) of the finalized line, start -> goal, whether or not it verifies
    ok: bool          # check_finish returned Some(line)
    chain: np.ndarray  # optimised chain goal -> root as (x, y, yaw) rows
    flags: int        # fragile decisions (chain choice, verdict)
    line_flags: int   # raw path flags of the final line's edges (knife-edge counts change samples, not decisions)
    verifies: int


def check_finish(world, nx, ny, nyaw, parent, node, goal, goal_yaw, radius, step, graze_tol=GRAZE_TOL):
    """RRT::check_finish(node) over a flat tree (src/rrt.rs:428-438 via finalize / optimize_from_goal)"""
    nx, ny, nyaw, parent = _tree(nx, ny, nyaw, parent)
    cap, ccap = 1 << 16, 4096
    while True:
        lx, ly = np.empty(cap), np.empty(cap)
        cx, cy, cyaw = np.empty(ccap), np.empty(ccap), np.empty(ccap)
        fl, lfl, nv, ln, cn = C.c_uint32(), C.c_uint32(), C.c_long(), C.c_long(), C.c_long()
        r = lib().ppo_check_finish(C.byref(world.w), nx.size, _p(nx), _p(ny), _p(nyaw), _p(parent, _i32p), int(node),
                                   float(goal[0]), float(goal[1]), float(goal_yaw), radius, step, graze_tol, _p(lx),
                                   _p(ly), cap, C.byref(ln), _p(cx), _p(cy), _p(cyaw), ccap, C.byref(cn),
                                   C.byref(fl), C.byref(lfl), C.byref(nv))
        if r == -2:
            cap *= 4
            ccap *= 4
            continue
        if r == -3:
            raise RuntimeError("reference would panic (finalize: no Dubins word)")
        n, c = ln.value, cn.value
        return Finish((lx[:n].copy(), ly[:n].copy()), r >= 0, np.stack([cx[:c], cy[:c], cyaw[:c]], 1),
                      int(fl.value), int(lfl.value), int(nv.value))


def circle_class(rx, ry, ax, ay, bx, by):
    """class of segment a-b against one (closed) ring under the circle filter of the culled loop: 0 skip, 1 blocked, 2 exact"""
    rx, ry = _f64(rx), _f64(ry)
    return int(lib().ppo_circle_class(_p(rx), _p(ry), rx.size, ax, ay, bx, by))


def ring_has_point(rx, ry, px, py):
    rx, ry = _f64(rx), _f64(ry)
    return bool(lib().ppo_ring_has_point(_p(rx), _p(ry), rx.size, px, py))


def point_position(rx, ry, px, py):
    rx, ry = _f64(rx), _f64(ry)
    return int(lib().ppo_point_position(_p(rx), _p(ry), rx.size, px, py))


def lines_intersect(ax, ay, bx, by):
    ax, ay, bx, by = map(_f64, (ax, ay, bx, by))
    return bool(lib().ppo_lines_intersect(_p(ax), _p(ay), ax.size, _p(bx), _p(by), bx.size))


def dubins_edge_polyline(sx, sy, syaw, ex, ey, eyaw, radius, step):
    cap = 4096
    while True:
        lx, ly = np.empty(cap), np.empty(cap)
        n = lib().ppo_dubins_edge_polyline(sx, sy, syaw, ex, ey, eyaw, radius, step, _p(lx), _p(ly), cap)
        if n == -2:
            cap *= 4
            continue
        if n < 0:
            raise RuntimeError("reference would panic")
        return lx[:n].copy(), ly[:n].copy()


def line_to_origin(nx, ny, nyaw, parent, node, radius, step):
    nx, ny, nyaw = map(_f64, (nx, ny, nyaw))
    parent = np.ascontiguousarray(parent, np.int32)
    cap = 1 << 16
    while True:
        lx, ly = np.empty(cap), np.empty(cap)
        n = lib().ppo_line_to_origin(_p(nx), _p(ny), _p(nyaw), _p(parent, _i32p), int(node), radius, step,
                                     _p(lx), _p(ly), cap)
        if n == -2:
            cap *= 4
            continue
        if n < 0:
            raise RuntimeError("reference would panic")
        return lx[:n].copy(), ly[:n].copy()


def uniform(seed, stream, n, lo=0.0, hi=1.0):
    out = np.empty(n, np.float64)
    lib().ppo_fill_uniform(seed, stream, n, lo, hi, _p(out))
    return out

"""ctypes binding of libpathplanning_b200.so (the C-ABI of include/pathplanning_b200.h).

This is the only place the Python host mirror touches native code.  There is no
CPU fallback: importing works anywhere (so CPU-only tests can check the exported
symbols), but creating a Context without a B200 raises, and a missing shared
library raises at import.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PP_B200_LIB") or os.path.join(_PKG, "libpathplanning_b200.so")  # env: A/B builds

PP_OK, PP_ERR_INVALID, PP_ERR_NO_DEVICE, PP_ERR_CUDA, PP_ERR_NOMEM, PP_ERR_STATE, PP_ERR_OVERFLOW, PP_ERR_COMM = 0, -1, -2, -3, -4, -5, -6, -7
COMM_ID_BYTES = 128
WORDS = ("LSL", "RSR", "LSR", "RSL", "RLR", "LRL")
WORD_NONE = 0xFF
COLLIDE_DEFAULT, COLLIDE_NO_CULL, COLLIDE_USE_GRID, COLLIDE_UNSORTED, COLLIDE_SCAN, COLLIDE_FUSED = 0, 1, 2, 4, 8, 16
NN_DEFAULT, NN_PLAIN_F64, NN_GRID, NN_UNSORTED, NN_SCAN = 0, 1, 2, 4, 8
PLAN_BYTES = 112

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a). There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p
_sz = C.c_size_t
_d = C.c_double
_i = C.c_int

# name -> (restype, argtypes); must list every symbol include/pathplanning_b200.h declares
SIGNATURES = {
    "pp_abi_version": (_i, []),
    "pp_status_string": (C.c_char_p, [_i]),
    "pp_device_count": (_i, []),
    "pp_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "pp_ctx_destroy": (None, [_vp]),
    "pp_last_error": (C.c_char_p, [_vp]),
    "pp_ctx_device": (_i, [_vp]),
    "pp_ctx_sm_count": (_i, [_vp]),
    "pp_ctx_stream": (_vp, [_vp]),
    "pp_ctx_set_stream": (_i, [_vp, _vp]),
    "pp_sync": (_i, [_vp]),
    "pp_launch_count": (C.c_uint64, [_vp]),
    "pp_host_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "pp_host_free": (_i, [_vp]),
    "pp_mod2pi": (_i, [_vp, _sz, _vp, _vp, _i]),
    "pp_dubins_words": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "pp_dubins_eval": (_i, [_vp, _sz] + [_vp] * 7 + [_d, _vp, _vp, _vp]),
    "pp_dubins_eval_dev": (_i, [_vp, _sz] + [_vp] * 7 + [_d, _vp, _vp, _vp]),
    "pp_dubins_sample_count": (_i, [_vp, _sz] + [_vp] * 6 + [_d, _d, _i, _vp, _vp]),
    "pp_dubins_sample_count_dev": (_i, [_vp, _sz] + [_vp] * 6 + [_d, _d, _i, _vp, _vp]),
    "pp_dubins_sample_fill": (_i, [_vp, _sz, _vp, _vp, C.c_uint64, _vp]),
    "pp_dubins_sample_fill_dev": (_i, [_vp, _sz, _vp, _vp, C.c_uint64, _vp]),
    "pp_exclusive_scan_u32_dev": (_i, [_vp, _sz, _vp, _vp, _vp]),
    "pp_dubins_path": (_i, [_vp] + [_d] * 8 + [_i, _vp, _vp, _vp, _sz, C.POINTER(_sz), C.POINTER(_i), C.POINTER(_d)]),
    "pp_tree_upload": (_i, [_vp, _sz, _vp, _vp, _vp, _vp]),
    "pp_tree_upload_dev": (_i, [_vp, _sz, _vp, _vp, _vp, _vp]),
    "pp_tree_append": (_i, [_vp, _sz, _vp, _vp, _vp, _vp]),
    "pp_tree_size": (_sz, [_vp]),
    "pp_obstacles_upload": (_i, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _sz]),
    "pp_nn": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _i]),
    "pp_nn_dev": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _i]),
    "pp_collide_segments": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _vp, _i]),
    "pp_collide_segments_dev": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _vp, _i]),
    "pp_verify_polylines": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _i]),
    "pp_collide_dubins": (_i, [_vp, _sz] + [_vp] * 6 + [_d, _d, _vp, _i]),
    "pp_collide_dubins_dev": (_i, [_vp, _sz] + [_vp] * 6 + [_d, _d, _vp, _i]),
    "pp_rrt_extend": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _vp, _i, _i]),
    "pp_rrt_extend_dev": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _vp, _i, _i]),
    "pp_rrt_extend_dubins": (_i, [_vp, _sz, _vp, _vp, _d, _d, _vp, _vp, _vp, _i, _i]),
    "pp_rrt_extend_dubins_dev": (_i, [_vp, _sz, _vp, _vp, _d, _d, _vp, _vp, _vp, _i, _i]),
    "pp_nn_grid_builds": (C.c_uint64, [_vp]),
    "pp_slice_bounds": (None, [_sz, _i, _i, C.POINTER(_sz), C.POINTER(_sz)]),
    "pp_comm_unique_id": (_i, [_vp]),
    "pp_ctx_comm_init": (_i, [_vp, _vp, _i, _i]),
    "pp_ctx_comm_rank": (_i, [_vp]),
    "pp_ctx_comm_size": (_i, [_vp]),
    "pp_tree_upload_bcast": (_i, [_vp, _i, _sz, _vp, _vp, _vp, _vp]),
    "pp_tree_append_bcast": (_i, [_vp, _i, _sz, _vp, _vp, _vp, _vp]),
    "pp_obstacles_upload_bcast": (_i, [_vp, _i, _vp, _vp, _sz, _vp, _vp, _vp, _sz]),
    "pp_group_create": (_i, [_vp, _i, C.POINTER(_vp)]),
    "pp_group_destroy": (None, [_vp]),
    "pp_group_size": (_i, [_vp]),
    "pp_group_ctx": (_vp, [_vp, _i]),
    "pp_group_last_error": (C.c_char_p, [_vp]),
    "pp_group_tree_upload": (_i, [_vp, _sz, _vp, _vp, _vp, _vp]),
    "pp_group_tree_append": (_i, [_vp, _sz, _vp, _vp, _vp, _vp]),
    "pp_group_obstacles_upload": (_i, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _sz]),
    "pp_group_dubins_eval": (_i, [_vp, _sz] + [_vp] * 7 + [_d, _vp, _vp, _vp]),
    "pp_group_nn": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _i]),
    "pp_group_collide_segments": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _vp, _i]),
    "pp_group_collide_dubins": (_i, [_vp, _sz] + [_vp] * 6 + [_d, _d, _vp, _i]),
    "pp_group_rrt_extend": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _vp, _i, _i]),
    "pp_group_rrt_extend_dubins": (_i, [_vp, _sz, _vp, _vp, _d, _d, _vp, _vp, _vp, _i, _i]),
    "pp_measure_copy": (_i, [_vp, _sz, _sz, _i, C.POINTER(_d)]),
    "pp_measure_fp64_peak": (_i, [_vp, _i, C.POINTER(_d), C.POINTER(_d)]),
    "pp_timing_enable": (_i, [_vp, _i]),
    "pp_timing_reset": (_i, [_vp]),
    "pp_timing_get": (_i, [_vp, C.c_char_p, C.POINTER(_d), C.POINTER(C.c_uint64)]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)  # AttributeError here = header and library out of sync
    _f.restype = _res
    _f.argtypes = _args


class PathPlanningError(RuntimeError):
    def __init__(self, status, msg=""):
        self.status = status
        super().__init__(f"pathplanning_b200 status {status} ({lib.pp_status_string(status).decode()}) {msg}")


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    """host numpy array, torch tensor (host or device) or None -> void*"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


def device_count() -> int:
    return int(lib.pp_device_count())


def slice_bounds(n: int, parts: int, part: int):
    """contiguous slice [lo, hi) of device `part` out of `parts` for a batch of n (pp_slice_bounds; no GPU needed)"""
    lo, hi = _sz(), _sz()
    lib.pp_slice_bounds(int(n), int(parts), int(part), C.byref(lo), C.byref(hi))
    return int(lo.value), int(hi.value)


def comm_unique_id() -> bytes:
    """128-byte NCCL id made by rank 0 and handed to the other ranks by the launcher's own channel"""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = lib.pp_comm_unique_id(buf)
    if rc:
        raise PathPlanningError(rc, "pp_comm_unique_id")
    return buf.raw


def _csr(rings_xy):
    off = np.zeros(len(rings_xy) + 1, np.uint32)
    for i, (rx, _) in enumerate(rings_xy):
        off[i + 1] = off[i] + len(rx)
    ox = _np(np.concatenate([np.asarray(r[0], np.float64) for r in rings_xy]) if rings_xy else np.zeros(0), np.float64)
    oy = _np(np.concatenate([np.asarray(r[1], np.float64) for r in rings_xy]) if rings_xy else np.zeros(0), np.float64)
    return ox, oy, off


class PinnedArray:
    """numpy view of pinned host memory from pp_host_alloc (freed with the object)."""

    def __init__(self, n, dtype):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(n) * self.dtype.itemsize
        p = _vp()
        rc = lib.pp_host_alloc(max(self.nbytes, 1), C.byref(p))
        if rc:
            raise PathPlanningError(rc, "pp_host_alloc")
        self._p = p
        buf = (C.c_char * max(self.nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(n))

    def __del__(self):
        p = getattr(self, "_p", None)
        if p is not None and p.value:
            self.array = None
            lib.pp_host_free(p)
            self._p = None


class Context:
    """One pp_ctx = one B200 + one stream.  Host-pointer calls take numpy arrays, `_dev` calls
    take torch CUDA tensors (or raw device pointers)."""

    def __init__(self, device: int = 0, _borrowed=None):
        self._owned = _borrowed is None
        if _borrowed is not None:  # a group's per-device context: owned by the group
            self._h = _vp(_borrowed)
            self.device = int(lib.pp_ctx_device(self._h))
            return
        h = _vp()
        rc = lib.pp_ctx_create(int(device), C.byref(h))
        if rc:
            raise PathPlanningError(rc, "pp_ctx_create: a B200 (sm_100) is required; there is no CPU fallback")
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            if self._owned:
                lib.pp_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what=""):
        if rc:
            raise PathPlanningError(rc, f"{what}: {lib.pp_last_error(self._h).decode()}")

    # ---- context
    @property
    def sm_count(self):
        return int(lib.pp_ctx_sm_count(self._h))

    @property
    def stream(self):
        return lib.pp_ctx_stream(self._h)

    def set_stream(self, cuda_stream):
        self._ck(lib.pp_ctx_set_stream(self._h, cuda_stream), "set_stream")

    def sync(self):
        self._ck(lib.pp_sync(self._h), "sync")

    @property
    def launch_count(self):
        return int(lib.pp_launch_count(self._h))

    def timing_enable(self, on=True):
        self._ck(lib.pp_timing_enable(self._h, int(on)))

    def timing_reset(self):
        self._ck(lib.pp_timing_reset(self._h))

    def timing_get(self, kernel: str):
        ms, n = _d(), C.c_uint64()
        self._ck(lib.pp_timing_get(self._h, kernel.encode(), C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    @property
    def nn_grid_builds(self):
        return int(lib.pp_nn_grid_builds(self._h))

    def measure_copy(self, h2d_bytes, d2h_bytes, pinned=True):
        """milliseconds for h2d_bytes up and d2h_bytes down at once, no kernel (the end-to-end ceiling of this box)"""
        ms = _d()
        self._ck(lib.pp_measure_copy(self._h, int(h2d_bytes), int(d2h_bytes), int(bool(pinned)), C.byref(ms)), "measure_copy")
        return ms.value

    # ---- one process per GPU: join an NCCL communicator, replicate tree / obstacles by broadcast
    def comm_init(self, comm_id: bytes, n_ranks: int, rank: int):
        buf = C.create_string_buffer(bytes(comm_id), COMM_ID_BYTES)
        self._ck(lib.pp_ctx_comm_init(self._h, buf, int(n_ranks), int(rank)), "comm_init")

    @property
    def comm_rank(self):
        return int(lib.pp_ctx_comm_rank(self._h))

    @property
    def comm_size(self):
        return int(lib.pp_ctx_comm_size(self._h))

    def tree_upload_bcast(self, root, n, x=None, y=None, yaw=None, parent=None):
        a = [None if v is None else _np(v, np.float64) for v in (x, y, yaw)]
        par = None if parent is None else _np(parent, np.int32)
        self._ck(lib.pp_tree_upload_bcast(self._h, int(root), int(n), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(par)),
                 "tree_upload_bcast")

    def tree_append_bcast(self, root, k, x=None, y=None, yaw=None, parent=None):
        a = [None if v is None else _np(np.atleast_1d(v), np.float64) for v in (x, y, yaw)]
        par = None if parent is None else _np(np.atleast_1d(parent), np.int32)
        self._ck(lib.pp_tree_append_bcast(self._h, int(root), int(k), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(par)),
                 "tree_append_bcast")

    def obstacles_upload_bcast(self, root, bounds_xy=None, rings_xy=None):
        if bounds_xy is None:
            self._ck(lib.pp_obstacles_upload_bcast(self._h, int(root), None, None, 0, None, None, None, 0),
                     "obstacles_upload_bcast")
            return
        bx, by = _np(bounds_xy[0], np.float64), _np(bounds_xy[1], np.float64)
        ox, oy, off = _csr(rings_xy)
        self._ck(lib.pp_obstacles_upload_bcast(self._h, int(root), _ptr(bx), _ptr(by), bx.size, _ptr(ox), _ptr(oy),
                                               _ptr(off), len(rings_xy)), "obstacles_upload_bcast")

    def measure_fp64_peak(self, iters=2048):
        v, ms = _d(), _d()
        self._ck(lib.pp_measure_fp64_peak(self._h, int(iters), C.byref(v), C.byref(ms)), "fp64 peak")
        return v.value, ms.value

    # ---- Dubins (host)
    def mod2pi(self, x, pi_2_pi=False):
        x = _np(x, np.float64)
        out = np.empty_like(x)
        self._ck(lib.pp_mod2pi(self._h, x.size, _ptr(x), _ptr(out), int(pi_2_pi)), "mod2pi")
        return out

    def dubins_words(self, alpha, beta, d):
        alpha, beta, d = (_np(v, np.float64) for v in (alpha, beta, d))
        n = alpha.size
        tpq = np.empty((n, 6, 3), np.float64)
        feas = np.empty((n, 6), np.uint8)
        self._ck(lib.pp_dubins_words(self._h, n, _ptr(alpha), _ptr(beta), _ptr(d), _ptr(tpq), _ptr(feas)), "words")
        return tpq, feas

    def dubins_eval(self, sx, sy, syaw, ex, ey, eyaw, radius=1.0, radius_arr=None, want_tpq=True, out=None):
        a = [_np(v, np.float64) for v in (sx, sy, syaw, ex, ey, eyaw)]
        n = a[0].size
        ra = _np(radius_arr, np.float64) if radius_arr is not None else None
        if out is None:
            cost = np.empty(n, np.float64)
            word = np.empty(n, np.uint8)
            tpq = np.empty((n, 3), np.float64) if want_tpq else None
        else:
            cost, word, tpq = out
        self._ck(lib.pp_dubins_eval(self._h, n, *[_ptr(v) for v in a], _ptr(ra), float(radius), _ptr(cost),
                                    _ptr(word), _ptr(tpq)), "dubins_eval")
        return cost, word, tpq

    def dubins_eval_dev(self, n, sx, sy, syaw, ex, ey, eyaw, radius, cost, word, tpq=None, radius_arr=None):
        self._ck(lib.pp_dubins_eval_dev(self._h, int(n), _ptr(sx), _ptr(sy), _ptr(syaw), _ptr(ex), _ptr(ey),
                                        _ptr(eyaw), _ptr(radius_arr), float(radius), _ptr(cost), _ptr(word),
                                        _ptr(tpq)), "dubins_eval_dev")

    def dubins_sample_count(self, sx, sy, syaw, ex, ey, eyaw, radius, step, from_origin=False):
        a = [_np(v, np.float64) for v in (sx, sy, syaw, ex, ey, eyaw)]
        n = a[0].size
        counts = np.empty(n, np.uint32)
        plan = np.empty(n * PLAN_BYTES, np.uint8)
        self._ck(lib.pp_dubins_sample_count(self._h, n, *[_ptr(v) for v in a], float(radius), float(step),
                                            int(from_origin), _ptr(counts), _ptr(plan)), "sample_count")
        return counts, plan

    def dubins_sample_fill(self, plan, counts):
        counts = _np(counts, np.uint32)
        c64 = np.where(counts == 0xFFFFFFFF, 0, counts).astype(np.uint64)
        offsets = np.zeros(counts.size, np.uint64)
        if counts.size:
            np.cumsum(c64[:-1], out=offsets[1:])
        total = int(c64.sum())
        out = np.empty((total, 3), np.float64)
        self._ck(lib.pp_dubins_sample_fill(self._h, counts.size, _ptr(plan), _ptr(offsets), total, _ptr(out)),
                 "sample_fill")
        return out, offsets

    def dubins_sample_count_dev(self, n, sx, sy, syaw, ex, ey, eyaw, radius, step, counts, plan, from_origin=False):
        self._ck(lib.pp_dubins_sample_count_dev(self._h, int(n), _ptr(sx), _ptr(sy), _ptr(syaw), _ptr(ex), _ptr(ey),
                                                _ptr(eyaw), float(radius), float(step), int(from_origin),
                                                _ptr(counts), _ptr(plan)), "sample_count_dev")

    def dubins_sample_fill_dev(self, n, plan, offsets, total, out):
        self._ck(lib.pp_dubins_sample_fill_dev(self._h, int(n), _ptr(plan), _ptr(offsets), int(total), _ptr(out)),
                 "sample_fill_dev")

    def exclusive_scan_u32_dev(self, n, counts, offsets, total):
        self._ck(lib.pp_exclusive_scan_u32_dev(self._h, int(n), _ptr(counts), _ptr(offsets), _ptr(total)), "scan")

    def dubins_path(self, sx, sy, syaw, ex, ey, eyaw, radius, step, from_origin=False):
        """-> (px, py, pyaw, word, cost) or None when no word is feasible (reference: None)."""
        cap = 4096
        while True:
            px, py, pyaw = (np.empty(cap, np.float64) for _ in range(3))
            n, w, c = _sz(), _i(), _d()
            rc = lib.pp_dubins_path(self._h, sx, sy, syaw, ex, ey, eyaw, radius, step, int(from_origin), _ptr(px),
                                    _ptr(py), _ptr(pyaw), cap, C.byref(n), C.byref(w), C.byref(c))
            if rc == PP_ERR_OVERFLOW and n.value > cap:
                cap = int(n.value)
                continue
            self._ck(rc, "dubins_path")
            if w.value == WORD_NONE:
                return None
            k = int(n.value)
            return px[:k].copy(), py[:k].copy(), pyaw[:k].copy(), int(w.value), float(c.value)

    # ---- tree / obstacles
    def tree_upload(self, x, y, yaw=None, parent=None):
        x, y = _np(x, np.float64), _np(y, np.float64)
        yaw = _np(yaw, np.float64) if yaw is not None else None
        parent = _np(parent, np.int32) if parent is not None else None
        self._ck(lib.pp_tree_upload(self._h, x.size, _ptr(x), _ptr(y), _ptr(yaw), _ptr(parent)), "tree_upload")

    def tree_upload_dev(self, n, x, y, yaw=None, parent=None):
        self._ck(lib.pp_tree_upload_dev(self._h, int(n), _ptr(x), _ptr(y), _ptr(yaw), _ptr(parent)), "tree_upload_dev")

    def tree_append(self, x, y, yaw=None, parent=None):
        x, y = _np(np.atleast_1d(x), np.float64), _np(np.atleast_1d(y), np.float64)
        yaw = _np(np.atleast_1d(yaw), np.float64) if yaw is not None else None
        parent = _np(np.atleast_1d(parent), np.int32) if parent is not None else None
        self._ck(lib.pp_tree_append(self._h, x.size, _ptr(x), _ptr(y), _ptr(yaw), _ptr(parent)), "tree_append")

    @property
    def tree_size(self):
        return int(lib.pp_tree_size(self._h))

    def obstacles_upload(self, bounds_xy, rings_xy):
        bx, by = _np(bounds_xy[0], np.float64), _np(bounds_xy[1], np.float64)
        ox, oy, off = _csr(rings_xy)
        self._ck(lib.pp_obstacles_upload(self._h, _ptr(bx), _ptr(by), bx.size, _ptr(ox), _ptr(oy), _ptr(off),
                                         len(rings_xy)), "obstacles_upload")

    def obstacles_upload_csr(self, bx, by, ox, oy, off):
        bx, by, ox, oy = (_np(v, np.float64) for v in (bx, by, ox, oy))
        off = _np(off, np.uint32)
        self._ck(lib.pp_obstacles_upload(self._h, _ptr(bx), _ptr(by), bx.size, _ptr(ox), _ptr(oy), _ptr(off),
                                         off.size - 1), "obstacles_upload")

    # ---- NN / verify (host)
    def nn(self, qx, qy, flags=NN_DEFAULT, want_d2=True):
        qx, qy = _np(np.atleast_1d(qx), np.float64), _np(np.atleast_1d(qy), np.float64)
        idx = np.empty(qx.size, np.uint32)
        d2 = np.empty(qx.size, np.float64) if want_d2 else None
        self._ck(lib.pp_nn(self._h, qx.size, _ptr(qx), _ptr(qy), _ptr(idx), _ptr(d2), int(flags)), "nn")
        return (idx, d2) if want_d2 else idx

    def nn_dev(self, m, qx, qy, idx, d2=None, flags=NN_DEFAULT):
        self._ck(lib.pp_nn_dev(self._h, int(m), _ptr(qx), _ptr(qy), _ptr(idx), _ptr(d2), int(flags)), "nn_dev")

    def collide_segments(self, ax, ay, bx, by, flags=COLLIDE_DEFAULT):
        a = [_np(np.atleast_1d(v), np.float64) for v in (ax, ay, bx, by)]
        ok = np.empty(a[0].size, np.uint8)
        self._ck(lib.pp_collide_segments(self._h, a[0].size, *[_ptr(v) for v in a], _ptr(ok), int(flags)),
                 "collide_segments")
        return ok

    def collide_segments_dev(self, m, ax, ay, bx, by, ok, flags=COLLIDE_DEFAULT):
        self._ck(lib.pp_collide_segments_dev(self._h, int(m), _ptr(ax), _ptr(ay), _ptr(bx), _ptr(by), _ptr(ok),
                                             int(flags)), "collide_segments_dev")

    def verify_polylines(self, lines, flags=COLLIDE_DEFAULT):
        """lines: list of (x[], y[]) polylines -> uint8[n] (Space::verify per line)"""
        off = np.zeros(len(lines) + 1, np.uint32)
        for i, (lx, _) in enumerate(lines):
            off[i + 1] = off[i] + len(lx)
        px = _np(np.concatenate([np.asarray(l[0], np.float64) for l in lines]) if lines else np.zeros(0), np.float64)
        py = _np(np.concatenate([np.asarray(l[1], np.float64) for l in lines]) if lines else np.zeros(0), np.float64)
        ok = np.empty(len(lines), np.uint8)
        self._ck(lib.pp_verify_polylines(self._h, len(lines), _ptr(px), _ptr(py), _ptr(off), _ptr(ok), int(flags)),
                 "verify_polylines")
        return ok

    def collide_dubins(self, sx, sy, syaw, ex, ey, eyaw, radius, step, flags=COLLIDE_DEFAULT):
        a = [_np(np.atleast_1d(v), np.float64) for v in (sx, sy, syaw, ex, ey, eyaw)]
        ok = np.empty(a[0].size, np.uint8)
        self._ck(lib.pp_collide_dubins(self._h, a[0].size, *[_ptr(v) for v in a], float(radius), float(step),
                                       _ptr(ok), int(flags)), "collide_dubins")
        return ok

    def collide_dubins_dev(self, m, sx, sy, syaw, ex, ey, eyaw, radius, step, ok, flags=COLLIDE_DEFAULT):
        self._ck(lib.pp_collide_dubins_dev(self._h, int(m), _ptr(sx), _ptr(sy), _ptr(syaw), _ptr(ex), _ptr(ey),
                                           _ptr(eyaw), float(radius), float(step), _ptr(ok), int(flags)),
                 "collide_dubins_dev")

    def rrt_extend(self, qx, qy, nn_flags=NN_DEFAULT, collide_flags=COLLIDE_DEFAULT, out=None):
        qx, qy = _np(np.atleast_1d(qx), np.float64), _np(np.atleast_1d(qy), np.float64)
        m = qx.size
        if out is None:
            idx, yaw, ok = np.empty(m, np.uint32), np.empty(m, np.float64), np.empty(m, np.uint8)
        else:
            idx, yaw, ok = out
        self._ck(lib.pp_rrt_extend(self._h, m, _ptr(qx), _ptr(qy), _ptr(idx), _ptr(yaw), _ptr(ok), int(nn_flags),
                                   int(collide_flags)), "rrt_extend")
        return idx, yaw, ok

    def rrt_extend_dev(self, m, qx, qy, idx, yaw, ok, nn_flags=NN_DEFAULT, collide_flags=COLLIDE_DEFAULT):
        self._ck(lib.pp_rrt_extend_dev(self._h, int(m), _ptr(qx), _ptr(qy), _ptr(idx), _ptr(yaw), _ptr(ok),
                                       int(nn_flags), int(collide_flags)), "rrt_extend_dev")

    def rrt_extend_dubins(self, qx, qy, radius, step, nn_flags=NN_DEFAULT, collide_flags=COLLIDE_DEFAULT):
        """nearest node, new node's yaw and verify of the Dubins edge new -> nearest, per sample point"""
        qx, qy = _np(np.atleast_1d(qx), np.float64), _np(np.atleast_1d(qy), np.float64)
        m = qx.size
        idx, yaw, ok = np.empty(m, np.uint32), np.empty(m, np.float64), np.empty(m, np.uint8)
        self._ck(lib.pp_rrt_extend_dubins(self._h, m, _ptr(qx), _ptr(qy), float(radius), float(step), _ptr(idx),
                                          _ptr(yaw), _ptr(ok), int(nn_flags), int(collide_flags)), "rrt_extend_dubins")
        return idx, yaw, ok

    def rrt_extend_dubins_dev(self, m, qx, qy, radius, step, idx, yaw, ok, nn_flags=NN_DEFAULT,
                              collide_flags=COLLIDE_DEFAULT):
        self._ck(lib.pp_rrt_extend_dubins_dev(self._h, int(m), _ptr(qx), _ptr(qy), float(radius), float(step), _ptr(idx),
                                              _ptr(yaw), _ptr(ok), int(nn_flags), int(collide_flags)),
                 "rrt_extend_dubins_dev")


class Group:
    """pp_group: ONE host process driving several B200s -- one context + worker thread per device, tree and obstacles
    replicated by ncclBroadcast, batches cut into contiguous slices (same call names and results as Context)."""

    def __init__(self, devices):
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = _vp()
        rc = lib.pp_group_create(devs, len(devices), C.byref(h))
        if rc:
            raise PathPlanningError(rc, "pp_group_create (B200s + libnccl.so.2 required; there is no CPU fallback)")
        self._h = h
        self.devices = [int(d) for d in devices]

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.pp_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what=""):
        if rc:
            raise PathPlanningError(rc, f"{what}: {lib.pp_group_last_error(self._h).decode()}")

    def __len__(self):
        return int(lib.pp_group_size(self._h))

    def ctx(self, i) -> Context:
        return Context(_borrowed=lib.pp_group_ctx(self._h, int(i)))

    def tree_upload(self, x, y, yaw=None, parent=None):
        x, y = _np(x, np.float64), _np(y, np.float64)
        yaw = _np(yaw, np.float64) if yaw is not None else None
        parent = _np(parent, np.int32) if parent is not None else None
        self._ck(lib.pp_group_tree_upload(self._h, x.size, _ptr(x), _ptr(y), _ptr(yaw), _ptr(parent)), "group_tree_upload")

    def tree_append(self, x, y, yaw=None, parent=None):
        x, y = _np(np.atleast_1d(x), np.float64), _np(np.atleast_1d(y), np.float64)
        yaw = _np(np.atleast_1d(yaw), np.float64) if yaw is not None else None
        parent = _np(np.atleast_1d(parent), np.int32) if parent is not None else None
        self._ck(lib.pp_group_tree_append(self._h, x.size, _ptr(x), _ptr(y), _ptr(yaw), _ptr(parent)), "group_tree_append")

    def obstacles_upload(self, bounds_xy, rings_xy):
        bx, by = _np(bounds_xy[0], np.float64), _np(bounds_xy[1], np.float64)
        ox, oy, off = _csr(rings_xy)
        self._ck(lib.pp_group_obstacles_upload(self._h, _ptr(bx), _ptr(by), bx.size, _ptr(ox), _ptr(oy), _ptr(off),
                                               len(rings_xy)), "group_obstacles_upload")

    def dubins_eval(self, sx, sy, syaw, ex, ey, eyaw, radius=1.0, radius_arr=None, want_tpq=True, out=None):
        a = [_np(v, np.float64) for v in (sx, sy, syaw, ex, ey, eyaw)]
        n = a[0].size
        ra = _np(radius_arr, np.float64) if radius_arr is not None else None
        if out is None:
            cost, word = np.empty(n, np.float64), np.empty(n, np.uint8)
            tpq = np.empty((n, 3), np.float64) if want_tpq else None
        else:
            cost, word, tpq = out
        self._ck(lib.pp_group_dubins_eval(self._h, n, *[_ptr(v) for v in a], _ptr(ra), float(radius), _ptr(cost),
                                          _ptr(word), _ptr(tpq)), "group_dubins_eval")
        return cost, word, tpq

    def nn(self, qx, qy, flags=NN_DEFAULT, want_d2=True):
        qx, qy = _np(np.atleast_1d(qx), np.float64), _np(np.atleast_1d(qy), np.float64)
        idx = np.empty(qx.size, np.uint32)
        d2 = np.empty(qx.size, np.float64) if want_d2 else None
        self._ck(lib.pp_group_nn(self._h, qx.size, _ptr(qx), _ptr(qy), _ptr(idx), _ptr(d2), int(flags)), "group_nn")
        return (idx, d2) if want_d2 else idx

    def collide_segments(self, ax, ay, bx, by, flags=COLLIDE_DEFAULT):
        a = [_np(np.atleast_1d(v), np.float64) for v in (ax, ay, bx, by)]
        ok = np.empty(a[0].size, np.uint8)
        self._ck(lib.pp_group_collide_segments(self._h, a[0].size, *[_ptr(v) for v in a], _ptr(ok), int(flags)),
                 "group_collide_segments")
        return ok

    def collide_dubins(self, sx, sy, syaw, ex, ey, eyaw, radius, step, flags=COLLIDE_DEFAULT, out=None):
        a = [_np(np.atleast_1d(v), np.float64) for v in (sx, sy, syaw, ex, ey, eyaw)]
        ok = np.empty(a[0].size, np.uint8) if out is None else out
        self._ck(lib.pp_group_collide_dubins(self._h, a[0].size, *[_ptr(v) for v in a], float(radius), float(step),
                                             _ptr(ok), int(flags)), "group_collide_dubins")
        return ok

    def rrt_extend(self, qx, qy, nn_flags=NN_DEFAULT, collide_flags=COLLIDE_DEFAULT, out=None):
        qx, qy = _np(np.atleast_1d(qx), np.float64), _np(np.atleast_1d(qy), np.float64)
        m = qx.size
        idx, yaw, ok = (np.empty(m, np.uint32), np.empty(m, np.float64), np.empty(m, np.uint8)) if out is None else out
        self._ck(lib.pp_group_rrt_extend(self._h, m, _ptr(qx), _ptr(qy), _ptr(idx), _ptr(yaw), _ptr(ok), int(nn_flags),
                                         int(collide_flags)), "group_rrt_extend")
        return idx, yaw, ok

    def rrt_extend_dubins(self, qx, qy, radius, step, nn_flags=NN_DEFAULT, collide_flags=COLLIDE_DEFAULT):
        qx, qy = _np(np.atleast_1d(qx), np.float64), _np(np.atleast_1d(qy), np.float64)
        m = qx.size
        idx, yaw, ok = np.empty(m, np.uint32), np.empty(m, np.float64), np.empty(m, np.uint8)
        self._ck(lib.pp_group_rrt_extend_dubins(self._h, m, _ptr(qx), _ptr(qy), float(radius), float(step), _ptr(idx),
                                                _ptr(yaw), _ptr(ok), int(nn_flags), int(collide_flags)),
                 "group_rrt_extend_dubins")
        return idx, yaw, ok

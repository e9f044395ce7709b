"""Host mirror of the crate's `dubins` module (/root/reference/src/dubins.rs) over the C-ABI.

Same names, argument order and return conventions as the Rust API:
    mod2pi, pi_2_pi                         src/dubins.rs:18, 22
    lsl, rsr, lsr, rsl, rlr, lrl            src/dubins.rs:27-153  -> (t, p, q, mode) with None for infeasible
    DubinsConfig                            src/dubins.rs:315-324
    dubins_path_planning_from_origin        src/dubins.rs:326     -> (px, py, pyaw, mode, cost) or None
    dubins_path_planning                    src/dubins.rs:401     -> (px, py, pyaw, mode, cost) or None
Every call runs on the GPU (a batch of one for the scalar functions); `batch_eval` / `batch_paths`
are the batched entry points the hot path is built for.  There is no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum

import numpy as np

from . import _ffi


class Mode(Enum):  # src/dubins.rs:4-9
    L = 0
    S = 1
    R = 2


L, S, R = Mode.L, Mode.S, Mode.R
# src/dubins.rs:26,50,73,94,115,135
LSL_MODE, RSR_MODE, LSR_MODE = (L, S, L), (R, S, R), (L, S, R)
RSL_MODE, RLR_MODE, LRL_MODE = (R, S, L), (R, L, R), (L, R, L)
WORD_MODES = (LSL_MODE, RSR_MODE, LSR_MODE, RSL_MODE, RLR_MODE, LRL_MODE)


def _ctx():
    from . import default_context
    return default_context()


def mod2pi(theta: float) -> float:
    return float(_ctx().mod2pi([theta])[0])


def pi_2_pi(angle: float) -> float:
    return float(_ctx().mod2pi([angle], pi_2_pi=True)[0])


def _word(w, alpha, beta, d):
    tpq, feas = _ctx().dubins_words([alpha], [beta], [d])
    if not feas[0, w]:
        return None, None, None, WORD_MODES[w]
    t, p, q = (float(v) for v in tpq[0, w])
    return t, p, q, WORD_MODES[w]


def lsl(alpha, beta, d):
    return _word(0, alpha, beta, d)


def rsr(alpha, beta, d):
    return _word(1, alpha, beta, d)


def lsr(alpha, beta, d):
    return _word(2, alpha, beta, d)


def rsl(alpha, beta, d):
    return _word(3, alpha, beta, d)


def rlr(alpha, beta, d):
    return _word(4, alpha, beta, d)


def lrl(alpha, beta, d):
    return _word(5, alpha, beta, d)


@dataclass
class DubinsConfig:  # src/dubins.rs:315-324
    sx: float
    sy: float
    syaw: float
    ex: float
    ey: float
    eyaw: float
    turn_radius: float
    step_size: float


def dubins_path_planning_from_origin(dx: float, dy: float, eyaw: float, c: float, step_size: float):
    r = _ctx().dubins_path(0.0, 0.0, 0.0, dx, dy, eyaw, 1.0 / c, step_size, from_origin=True)
    if r is None:
        return None
    px, py, pyaw, w, cost = r
    return px, py, pyaw, WORD_MODES[w], cost


def dubins_path_planning(conf: DubinsConfig):
    r = _ctx().dubins_path(conf.sx, conf.sy, conf.syaw, conf.ex, conf.ey, conf.eyaw, conf.turn_radius,
                           conf.step_size, from_origin=False)
    if r is None:
        return None
    px, py, pyaw, w, cost = r
    return px, py, pyaw, WORD_MODES[w], cost


# ---- batched entry points (what the GPU path is for) ------------------------------------------------
def batch_eval(sx, sy, syaw, ex, ey, eyaw, turn_radius=1.0, want_tpq=True, ctx=None):
    """cost (radius-normalised), word (0..5 / 0xFF), (t,p,q) for n pose pairs; turn_radius scalar or array"""
    ctx = ctx or _ctx()
    if np.ndim(turn_radius) == 0:
        return ctx.dubins_eval(sx, sy, syaw, ex, ey, eyaw, radius=float(turn_radius), want_tpq=want_tpq)
    return ctx.dubins_eval(sx, sy, syaw, ex, ey, eyaw, radius_arr=turn_radius, want_tpq=want_tpq)


def batch_paths(sx, sy, syaw, ex, ey, eyaw, turn_radius, step_size, from_origin=False, ctx=None):
    """samples of n paths: returns (xyyaw[total,3], offsets[n], counts[n]); path i = rows offsets[i] .. +counts[i]"""
    ctx = ctx or _ctx()
    counts, plan = ctx.dubins_sample_count(sx, sy, syaw, ex, ey, eyaw, turn_radius, step_size, from_origin)
    out, offsets = ctx.dubins_sample_fill(plan, counts)
    return out, offsets, counts

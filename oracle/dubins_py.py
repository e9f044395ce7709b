"""Independent pure-Python transliteration of /root/reference/src/dubins.rs (lines 14-428).

TEST INFRASTRUCTURE ONLY.  Its single purpose is to pin oracle/pp_oracle.c: the two
restatements were written separately (this one mirrors the Rust data flow with
Python lists standing in for Vec, including the pop-based trim loop) and must agree
bit for bit on every test input.  All libm calls go to glibc through ctypes so that
sin/cos/atan2/acos/hypot/fmod are the functions rustc's f64 methods reach on
x86_64-unknown-linux-gnu (CPython's own math.hypot differs from glibc by 1 ulp in places).
Python float arithmetic is IEEE binary64 without contraction, like rustc's.
"""
import ctypes
import ctypes.util

_m = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("sin", "cos", "acos", "sqrt", "floor", "trunc", "fabs"):
    getattr(_m, _n).restype = ctypes.c_double
    getattr(_m, _n).argtypes = [ctypes.c_double]
for _n in ("atan2", "hypot", "fmod"):
    getattr(_m, _n).restype = ctypes.c_double
    getattr(_m, _n).argtypes = [ctypes.c_double, ctypes.c_double]

sin, cos, acos, sqrt, floor, trunc = _m.sin, _m.cos, _m.acos, _m.sqrt, _m.floor, _m.trunc
atan2, hypot, fmod = _m.atan2, _m.hypot, _m.fmod

PI = 3.141592653589793
INF = float("inf")
L, S, R = "L", "S", "R"


def fmodr(x, y):  # dubins.rs:14
    return x - y * floor(x / y)


def mod2pi(theta):  # dubins.rs:18
    return fmodr(theta, 2.0 * PI)


def pi_2_pi(angle):  # dubins.rs:22
    return fmod(angle + PI, 2.0 * PI) - PI


def _trig(alpha, beta):
    return sin(alpha), sin(beta), cos(alpha), cos(beta), cos(alpha - beta)


def lsl(alpha, beta, d):  # dubins.rs:27
    sa, sb, ca, cb, c_ab = _trig(alpha, beta)
    tmp0 = d + sa - sb
    p_squared = 2.0 + (d * d) - (2.0 * c_ab) + (2.0 * d * (sa - sb))
    if p_squared < 0.0:
        return None, None, None, (L, S, L)
    tmp1 = atan2(cb - ca, tmp0)
    t = mod2pi(-alpha + tmp1)
    p = sqrt(p_squared)
    q = mod2pi(beta - tmp1)
    return t, p, q, (L, S, L)


def rsr(alpha, beta, d):  # dubins.rs:51
    sa, sb, ca, cb, c_ab = _trig(alpha, beta)
    tmp0 = d - sa + sb
    p_squared = 2.0 + (d * d) - (2.0 * c_ab) + (2.0 * d * (sb - sa))
    if p_squared < 0.0:
        return None, None, None, (R, S, R)
    tmp1 = atan2(ca - cb, tmp0)
    t = mod2pi(alpha - tmp1)
    p = sqrt(p_squared)
    q = mod2pi(-beta + tmp1)
    return t, p, q, (R, S, R)


def lsr(alpha, beta, d):  # dubins.rs:74
    sa, sb, ca, cb, c_ab = _trig(alpha, beta)
    p_squared = -2.0 + (d * d) + (2.0 * c_ab) + (2.0 * d * (sa + sb))
    if p_squared < 0.0:
        return None, None, None, (L, S, R)
    p = sqrt(p_squared)
    tmp = atan2(-ca - cb, d + sa + sb) - atan2(-2.0, p)
    t = mod2pi(-alpha + tmp)
    q = mod2pi(-mod2pi(beta) + tmp)
    return t, p, q, (L, S, R)


def rsl(alpha, beta, d):  # dubins.rs:95
    sa, sb, ca, cb, c_ab = _trig(alpha, beta)
    p_squared = -2.0 + (d * d) + (2.0 * c_ab) - (2.0 * d * (sa + sb))
    if p_squared < 0.0:
        return None, None, None, (R, S, L)
    p = sqrt(p_squared)
    tmp = atan2(ca + cb, d - sa - sb) - atan2(2.0, p)
    t = mod2pi(alpha - tmp)
    q = mod2pi(beta - tmp)
    return t, p, q, (R, S, L)


def rlr(alpha, beta, d):  # dubins.rs:116
    sa, sb, ca, cb, c_ab = _trig(alpha, beta)
    tmp_rlr = (6.0 - d * d + 2.0 * c_ab + 2.0 * d * (sa - sb)) / 8.0
    if abs(tmp_rlr) > 1.0:
        return None, None, None, (R, L, R)
    p = mod2pi(2.0 * PI - acos(tmp_rlr))
    t = mod2pi(alpha - atan2(ca - cb, d - sa + sb) + mod2pi(p / 2.0))
    q = mod2pi(alpha - beta - t + mod2pi(p))
    return t, p, q, (R, L, R)


def lrl(alpha, beta, d):  # dubins.rs:136
    sa, sb, ca, cb, c_ab = _trig(alpha, beta)
    tmp_lrl = (6.0 - d * d + 2.0 * c_ab + 2.0 * d * (-sa + sb)) / 8.0
    if abs(tmp_lrl) > 1.0:
        return None, None, None, (L, R, L)
    p = mod2pi(2.0 * PI - acos(tmp_lrl))
    t = mod2pi(-alpha - atan2(ca - cb, d + sa - sb) + p / 2.0)
    q = mod2pi(mod2pi(beta) - alpha - t + mod2pi(p))
    return t, p, q, (L, R, L)


ALL_PLANNERS = (lsl, rsr, lsr, rsl, rlr, lrl)  # dubins.rs:291
WORD_NAMES = ("LSL", "RSR", "LSR", "RSL", "RLR", "LRL")


def interpolate(ind, length, mode, max_curvature, ox, oy, oyaw, path_x, path_y, path_yaw, directions):  # :155
    if mode == S:
        path_x[ind] = ox + length / max_curvature * cos(oyaw)
        path_y[ind] = oy + length / max_curvature * sin(oyaw)
        path_yaw[ind] = oyaw
    else:
        ldx = sin(length) / max_curvature
        ldy = 0.0
        if mode == L:
            ldy = (1.0 - cos(length)) / max_curvature
        elif mode == R:
            ldy = (1.0 - cos(length)) / -max_curvature
        gdx = cos(-oyaw) * ldx + sin(-oyaw) * ldy
        gdy = -sin(-oyaw) * ldx + cos(-oyaw) * ldy
        path_x[ind] = ox + gdx
        path_y[ind] = oy + gdy
    if mode == L:
        path_yaw[ind] = oyaw + length
    elif mode == R:
        path_yaw[ind] = oyaw - length
    directions[ind] = 1 if length > 0.0 else -1


def generate_local_course(lengths, mode, max_curvature, step_size, path_x, path_y, path_yaw, directions):  # :200
    ind = 1
    directions[0] = 1 if lengths[0] > 0.0 else -1
    ll = 0.0
    for i in range(3):
        m, l = mode[i], lengths[i]
        d = step_size if l > 0.0 else -step_size
        ox, oy, oyaw = path_x[ind], path_y[ind], path_yaw[ind]
        ind -= 1
        if i >= 1 and (lengths[i - 1] * lengths[i]) > 0.0:
            pd = -d - ll
        else:
            pd = d - ll
        while abs(pd) <= abs(l):
            ind += 1
            interpolate(ind, pd, m, max_curvature, ox, oy, oyaw, path_x, path_y, path_yaw, directions)
            pd += d
        ll = l - pd - d
        ind += 1
        interpolate(ind, l, m, max_curvature, ox, oy, oyaw, path_x, path_y, path_yaw, directions)
    if len(path_x) <= 1:
        path_x.clear(), path_y.clear(), path_yaw.clear(), directions.clear()
    last = path_x[len(path_x) - 1]
    while len(path_x) >= 1 and last == 0.0:
        last = path_x[len(path_x) - 1]
        path_x.pop(), path_y.pop(), path_yaw.pop(), directions.pop()


def dubins_path_planning_from_origin(dx, dy, eyaw, c, step_size):  # :326
    hyp = hypot(dx, dy)
    d = hyp * c
    theta = mod2pi(atan2(dy, dx))
    alpha = mod2pi(-theta)
    beta = mod2pi(eyaw - theta)
    bcost = INF
    bt = bp = bq = bmode = None
    bword = None
    for wi, planner in enumerate(ALL_PLANNERS):
        t, p, q, mode = planner(alpha, beta, d)
        if t is not None and p is not None and q is not None:
            cost = abs(t) + abs(p) + abs(q)
            if bcost > cost:
                bt, bp, bq, bmode, bword = t, p, q, mode, wi
                bcost = cost
    if bt is None:
        return None
    lengths = [bt, bp, bq]
    total_length = 0.0
    for v in lengths:
        total_length = total_length + v
    n_point = int(trunc(total_length / step_size)) + len(lengths) + 4
    px, py, pyaw = [0.0] * n_point, [0.0] * n_point, [0.0] * n_point
    directions = [0] * n_point
    generate_local_course(lengths, bmode, c, step_size, px, py, pyaw, directions)
    return px, py, pyaw, bword, bcost, (bt, bp, bq), n_point


def dubins_path_planning(sx, sy, syaw, ex, ey, eyaw, turn_radius, step_size):  # :401
    ex_ = ex - sx
    ey_ = ey - sy
    c = 1.0 / turn_radius
    lex = cos(syaw) * ex_ + sin(syaw) * ey_
    ley = -(sin(syaw)) * ex_ + cos(syaw) * ey_
    leyaw = eyaw - syaw
    r = dubins_path_planning_from_origin(lex, ley, leyaw, c, step_size)
    if r is None:
        return None
    lpx, lpy, lpyaw, word, clen, tpq, n_point = r
    px = [cos(-syaw) * x + sin(-syaw) * y + sx for x, y in zip(lpx, lpy)]
    py = [-sin(-syaw) * x + cos(-syaw) * y + sy for x, y in zip(lpx, lpy)]
    pyaw = [pi_2_pi(iyaw + syaw) for iyaw in lpyaw]
    return px, py, pyaw, word, clen, tpq, n_point

"""Counter-based synthetic inputs: element i of any array is
u = splitmix64_finalise(seed + stream*2^56 + (i+1)*0x9E3779B97F4A7C15), uniform = (u >> 11) * 2^-53,
U[a,b) = a + (b-a)*uniform (non-fused); one `stream` id per array, so element i is position-independent."""
from __future__ import annotations

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def uniform(seed: int, stream: int, n: int, lo: float = 0.0, hi: float = 1.0, first: int = 0) -> np.ndarray:
    with np.errstate(over="ignore"):
        i = np.arange(first + 1, first + n + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.uint64((stream << 56) & 0xFFFFFFFFFFFFFFFF) + i * GOLDEN
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    u = (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return lo + (hi - lo) * u


SEED_C3, SEED_C4_Q, SEED_C4_N, SEED_C4_OBS, SEED_C5 = 0xD0B10003, 0xD0B10004, 0xD0B10005, 0xD0B10006, 0xD0B10007


def dubins_pairs(n: int, dist: str = "mixed", seed: int = SEED_C3, first: int = 0):
    """config C3: positions U[-2,2) ('mixed') or U[-50,50) ('far'), yaws U[-pi,pi)"""
    span = 2.0 if dist == "mixed" else 50.0
    sx = uniform(seed, 0, n, -span, span, first)
    sy = uniform(seed, 1, n, -span, span, first)
    syaw = uniform(seed, 2, n, -np.pi, np.pi, first)
    ex = uniform(seed, 3, n, -span, span, first)
    ey = uniform(seed, 4, n, -span, span, first)
    eyaw = uniform(seed, 5, n, -np.pi, np.pi, first)
    return sx, sy, syaw, ex, ey, eyaw


def create_circle(cx: float, cy: float, radius: float):
    """rrt::create_circle (src/rrt.rs:43-60) + Polygon::new ring closing.  math.cos/sin are glibc's."""
    import math
    n = math.ceil(2.0 * math.pi * radius / 1.0)
    cnt = int(n + 1.0)
    xs = [(math.cos(2.0 * math.pi / n * float(k)) * radius) + cx for k in range(cnt)]
    ys = [(math.sin(2.0 * math.pi / n * float(k)) * radius) + cy for k in range(cnt)]
    if cnt and (xs[0] != xs[-1] or ys[0] != ys[-1]):
        xs.append(xs[0])
        ys.append(ys[0])
    return np.asarray(xs, np.float64), np.asarray(ys, np.float64)


def circle_world(n_rings: int, world: float = 1000.0, rmin: float = 1.0, rmax: float = 3.0,
                 seed: int = SEED_C4_OBS, shift: float = 0.0):
    """config C4/C5 obstacle set: create_circle rings at U[0,world)^2, r ~ U[rmin,rmax); bounds = [0,world]^2.
    shift != 0 translates every ring outside the world (the no-hit set)."""
    cx = uniform(seed, 0, n_rings, 0.0, world) + shift
    cy = uniform(seed, 1, n_rings, 0.0, world) + shift
    rr = uniform(seed, 2, n_rings, rmin, rmax)
    rings = [create_circle(float(cx[i]), float(cy[i]), float(rr[i])) for i in range(n_rings)]
    bounds = (np.array([0.0, 0.0, world, world, 0.0]), np.array([0.0, world, world, 0.0, 0.0]))
    return bounds, rings


def extend_inputs(m: int, n_nodes: int, world: float = 1000.0):
    """config C4: m sample points and an n_nodes tree, both U[0,world)^2; node yaw U[-pi,pi)"""
    qx = uniform(SEED_C4_Q, 0, m, 0.0, world)
    qy = uniform(SEED_C4_Q, 1, m, 0.0, world)
    nx = uniform(SEED_C4_N, 0, n_nodes, 0.0, world)
    ny = uniform(SEED_C4_N, 1, n_nodes, 0.0, world)
    nyaw = uniform(SEED_C4_N, 2, n_nodes, -np.pi, np.pi)
    return qx, qy, nx, ny, nyaw


def dubins_edges(e: int, world: float = 1000.0, reach: float = 50.0, seed: int = SEED_C5, first: int = 0):
    """config C5: child ~ U[0,world)^2, parent = child + U[-reach,reach)^2, child yaw = atan2(parent - child)
    (src/rrt.rs:267-271), parent yaw U[-pi,pi)"""
    sx = uniform(seed, 0, e, 0.0, world, first)
    sy = uniform(seed, 1, e, 0.0, world, first)
    ex = sx + uniform(seed, 2, e, -reach, reach, first)
    ey = sy + uniform(seed, 3, e, -reach, reach, first)
    syaw = np.arctan2(ey - sy, ex - sx)
    eyaw = uniform(seed, 4, e, -np.pi, np.pi, first)
    return sx, sy, syaw, ex, ey, eyaw

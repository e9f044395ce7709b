"""Host mirror of the crate's `rrt` module (/root/reference/src/rrt.rs) over the C-ABI.

Kept from the reference: Robot (:17-40), create_circle (:43-60), Space (:70-159; verify on the GPU),
Node / NodeIter (:161-265), line_to_origin (:291-321, node->root chunk order), RRT (:325-619) with
get_nearest_node, get_random_node, verify_node, check_finish, optimize, optimize_from_goal, finalize,
plan_one and plan.  The tree lives twice: Python Node objects (parent links, as Arc<Node>) and a flat SoA
mirror on the GPU that the NN kernel scans.
Every Space owns its own GPU context (one world + one tree), as an RRT owns its Space and its RTree in the reference
(:325-356): any number of Space / RRT pairs can live side by side.
Out of scope: Space::new's geo-offset inflation -- the geo-offset crate's source is not
available, so Space(...) REFUSES a robot of non-zero width instead of silently dropping the safety margin, and
geometry that already carries it goes through Space.from_inflated(...); rand_point uses a seedable numpy Generator
instead of thread_rng.
"""
from __future__ import annotations

import itertools
import math
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi, synth
from .dubins import DubinsConfig

RECURSION_LIMIT = 16  # src/rrt.rs:14
SHOULD_PLAN = "Should plan dubins curve"  # src/rrt.rs:529: finalize panics when an edge has no feasible word
Ring = Tuple[np.ndarray, np.ndarray]


class Robot:  # src/rrt.rs:17-40
    def __init__(self, width: float, height: float, max_steer: float):
        self.width, self.height, self.max_steer = width, height, max_steer

    def get_width(self):
        return self.width

    def get_steer(self):
        return self.max_steer


def create_circle(center: Tuple[float, float], radius: float) -> Ring:  # src/rrt.rs:43-60
    return synth.create_circle(float(center[0]), float(center[1]), float(radius))


def _closed(ring: Ring) -> Ring:
    x, y = np.asarray(ring[0], np.float64), np.asarray(ring[1], np.float64)
    if x.size and (x[0] != x[-1] or y[0] != y[-1]):  # Polygon::new closes rings
        x, y = np.append(x, x[0]), np.append(y, y[0])
    return x, y


class Space:  # src/rrt.rs:70-159
    def __init__(self, bounds: Ring, robot: Robot, obstacle_list: Sequence[Ring], ctx=None, seed=None, device: int = 0,
                 _inflated: bool = False):
        if not _inflated and robot.get_width() != 0:
            # src/rrt.rs:81-111 shrinks the bounds and inflates the obstacles by width / 2 through geo-offset
            raise _ffi.PathPlanningError(
                _ffi.PP_ERR_INVALID,
                "Space::new: geo-offset inflation by Robot.width / 2 is not available in this mirror; pass "
                "pre-inflated bounds / obstacles through Space.from_inflated(...)")
        # one context per Space (world + tree live in it); a caller-supplied ctx is used as is and must not be
        # shared with another live Space / RRT
        self._own_ctx = ctx is None
        self.ctx = _ffi.Context(device) if ctx is None else ctx
        self.bounds = _closed(bounds)
        self.robot = robot
        self.obstacles = [_closed(o) for o in obstacle_list]
        bx, by = self.bounds
        self.minx, self.maxx, self.miny, self.maxy = float(bx.min()), float(bx.max()), float(by.min()), float(by.max())
        self._rng = np.random.default_rng(seed)
        self.ctx.obstacles_upload(self.bounds, self.obstacles)

    @classmethod
    def from_inflated(cls, bounds: Ring, robot: Robot, obstacle_list: Sequence[Ring], ctx=None, seed=None,
                      device: int = 0) -> "Space":
        """bounds already shrunk and obstacles already inflated by robot.width / 2 (what Space::new computes)"""
        return cls(bounds, robot, obstacle_list, ctx=ctx, seed=seed, device=device, _inflated=True)

    def close(self):
        if self._own_ctx and self.ctx is not None:
            self.ctx.close()
        self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def verify(self, line: Ring) -> bool:  # src/rrt.rs:124-137
        return bool(self.ctx.verify_polylines([line])[0])

    def verify_many(self, lines: Sequence[Ring]) -> np.ndarray:
        return self.ctx.verify_polylines(list(lines)).astype(bool)

    def rand_point(self) -> Tuple[float, float]:  # src/rrt.rs:139-146
        return (float(self._rng.uniform(self.minx, self.maxx)), float(self._rng.uniform(self.miny, self.maxy)))

    def get_steer(self):
        return self.robot.get_steer()

    def get_obs(self):
        return list(self.obstacles)

    def get_bounds(self):
        return self.bounds


def compute_yaw(frm: Tuple[float, float], to: Tuple[float, float]) -> float:  # src/rrt.rs:267-271
    return math.atan2(to[1] - frm[1], to[0] - frm[0])


class Node:  # src/rrt.rs:161-214
    __slots__ = ("point", "parent", "yaw")

    def __init__(self, point: Tuple[float, float], parent: "Node"):
        self.point = (float(point[0]), float(point[1]))
        self.parent = parent
        self.yaw = compute_yaw(self.point, parent.get_point())

    @classmethod
    def new_root(cls, point, yaw) -> "Node":
        n = object.__new__(cls)
        n.point, n.parent, n.yaw = (float(point[0]), float(point[1])), None, float(yaw)
        return n

    @classmethod
    def new_goal(cls, point, parent: "Node", yaw) -> "Node":
        n = object.__new__(cls)
        n.point, n.parent, n.yaw = (float(point[0]), float(point[1])), parent, float(yaw)
        return n

    def get_parent(self) -> Optional["Node"]:
        return self.parent

    def get_above(self) -> "NodeIter":
        return NodeIter(self.parent)

    def get_point(self):
        return self.point

    def get_coord(self):
        return self.point

    def get_yaw(self):
        return self.yaw


class NodeIter:  # src/rrt.rs:248-265
    def __init__(self, curr: Optional[Node]):
        self.curr = curr

    def __iter__(self) -> Iterator[Node]:
        return self

    def __next__(self) -> Node:
        if self.curr is None:
            raise StopIteration
        c = self.curr
        self.curr = c.get_parent()
        return c


def _chain_edges(node: Node):
    """(child, parent) pose pairs along node -> root"""
    sx, sy, syaw, ex, ey, eyaw = [], [], [], [], [], []
    for n in NodeIter(node):
        p = n.get_parent()
        if p is None:
            break
        sx.append(n.point[0]); sy.append(n.point[1]); syaw.append(n.yaw)
        ex.append(p.point[0]); ey.append(p.point[1]); eyaw.append(p.yaw)
    return sx, sy, syaw, ex, ey, eyaw


def line_to_origin(node: Node, turn_radius: float, step_size: float, ctx=None, _strict: bool = False) -> Ring:
    """src/rrt.rs:291-321.  Polyline node -> root: per-edge Dubins samples (one batched GPU call), then the root's point.
    An edge without a feasible word contributes its start point (:313); finalize's copy of the loop panics instead
    (:529), which `_strict` reproduces"""
    from . import default_context
    ctx = ctx or default_context()
    sx, sy, syaw, ex, ey, eyaw = _chain_edges(node)
    xs: List[np.ndarray] = []
    ys: List[np.ndarray] = []
    if sx:
        counts, plan = ctx.dubins_sample_count(sx, sy, syaw, ex, ey, eyaw, turn_radius, step_size)
        out, offsets = ctx.dubins_sample_fill(plan, counts)
        words = np.frombuffer(plan, np.uint8).reshape(-1, _ffi.PLAN_BYTES)[:, 104]
        for i in range(len(sx)):
            if words[i] == _ffi.WORD_NONE:  # src/rrt.rs:313
                if _strict:
                    raise RuntimeError(SHOULD_PLAN)
                xs.append(np.array([sx[i]])); ys.append(np.array([sy[i]]))
            else:
                o, c = int(offsets[i]), int(counts[i])
                xs.append(out[o:o + c, 0]); ys.append(out[o:o + c, 1])
    root = node
    while root.get_parent() is not None:
        root = root.get_parent()
    xs.append(np.array([root.point[0]])); ys.append(np.array([root.point[1]]))
    return np.concatenate(xs), np.concatenate(ys)


def euclidean_length(line: Ring) -> float:
    x, y = line
    return float(np.hypot(np.diff(x), np.diff(y)).sum()) if len(x) > 1 else 0.0


class RRT:  # src/rrt.rs:325-619
    def __init__(self, start, start_yaw, goal, goal_yaw, max_iter: int, step_size: float, space: Space):
        self.goal, self.goal_yaw = (float(goal[0]), float(goal[1])), float(goal_yaw)
        self.max_iter, self.step_size, self.space = int(max_iter), float(step_size), space
        self.ctx = space.ctx
        root = Node.new_root(start, start_yaw)
        self.nodes: List[Node] = [root]  # index i <-> device tree slot i (replaces the RTree, :345-346)
        self._slot = {id(root): 0}
        self.ctx.tree_upload([root.point[0]], [root.point[1]], [root.yaw], [-1])

    # -- src/rrt.rs:378-391
    def get_nearest_node(self, point) -> Optional[Node]:
        idx = self.ctx.nn([point[0]], [point[1]], want_d2=False)
        return None if idx[0] == 0xFFFFFFFF else self.nodes[int(idx[0])]

    # -- src/rrt.rs:406-412
    def get_random_node(self) -> Optional[Node]:
        point = self.space.rand_point()
        nearest = self.get_nearest_node(point)
        return None if nearest is None else Node(point, nearest)

    # -- src/rrt.rs:414-426 : verify of the whole chain = AND over its edges (each = samples ++ [parent])
    def verify_node(self, node: Node) -> bool:
        sx, sy, syaw, ex, ey, eyaw = _chain_edges(node)
        if not sx:
            return self.space.verify((np.array([node.point[0]]), np.array([node.point[1]])))
        ok = self.ctx.collide_dubins(sx, sy, syaw, ex, ey, eyaw, self.space.get_steer(), self.step_size)
        return bool(ok.all())

    def _insert(self, node: Node):  # src/rrt.rs:586-589
        self.nodes.append(node)
        self._slot[id(node)] = len(self.nodes) - 1
        par = self._slot[id(node.parent)] if node.parent is not None else -1
        self.ctx.tree_append([node.point[0]], [node.point[1]], [node.yaw], [par])

    # -- src/rrt.rs:428-438
    def check_finish(self, node: Node) -> Optional[Ring]:
        goal_node = Node.new_goal(self.goal, node, self.goal_yaw)
        line = self.finalize(goal_node)
        return line if self.space.verify(line) else None

    # -- src/rrt.rs:463-487.  Same candidates in the same (root-first) order and the same verdicts as the
    # reference's loop, but ALL shortcut candidates of a level are verified in ONE fused launch:
    # verify(line_to_origin(new_node)) = verify(edge new_node -> to_node) AND verify(chain of to_node), and the
    # chain verdicts are prefix-ANDs over the ancestors' own edges, so 2*depth edges replace depth^2.
    def optimize(self, node: Node, i: int) -> Optional[Node]:
        if i >= RECURSION_LIMIT:
            return None
        nodes_vec = list(NodeIter(node))  # node, parent, ..., root
        cands = [Node(node.get_coord(), to_node) for to_node in nodes_vec]
        valid = self._verify_shortcuts(nodes_vec, cands)
        for k in range(len(nodes_vec) - 1, -1, -1):  # .rev(): root first
            if valid[k]:
                deeper = self.optimize(nodes_vec[k], i + 1)
                return Node(node.get_coord(), deeper) if deeper is not None else cands[k]
        return None

    def _verify_shortcuts(self, chain: List[Node], cands: List[Node]) -> np.ndarray:
        """valid[k] = verify_node(cands[k]) where cands[k].parent is chain[k] (chain = node ... root)"""
        n = len(chain)
        sx = [c.point[0] for c in cands] + [a.point[0] for a in chain[:-1]]
        sy = [c.point[1] for c in cands] + [a.point[1] for a in chain[:-1]]
        syaw = [c.yaw for c in cands] + [a.yaw for a in chain[:-1]]
        ex = [a.point[0] for a in chain] + [a.parent.point[0] for a in chain[:-1]]
        ey = [a.point[1] for a in chain] + [a.parent.point[1] for a in chain[:-1]]
        eyaw = [a.yaw for a in chain] + [a.parent.yaw for a in chain[:-1]]
        ok = self.ctx.collide_dubins(sx, sy, syaw, ex, ey, eyaw, self.space.get_steer(), self.step_size).astype(bool)
        first, own = ok[:n], ok[n:]  # own[j]: edge chain[j] -> chain[j+1]
        # chain_ok[k] = AND of own[k:]; the root's own point is the last point of every edge that ends at the root
        chain_ok = np.ones(n, bool)
        if n > 1:
            chain_ok[: n - 1] = np.logical_and.accumulate(own[::-1])[::-1]
        return first & chain_ok

    # -- the same recursion for MANY start nodes at once: all candidates of all nodes of a recursion level go into one
    # fused launch, so a round's goal checks cost RECURSION_LIMIT launches instead of RECURSION_LIMIT per node.
    # Per node the candidates, their order and the verdicts are those of optimize() (tests compare them).
    def _optimize_many(self, nodes: List[Node], i: int = 0) -> List[Optional[Node]]:
        if i >= RECURSION_LIMIT or not nodes:
            return [None] * len(nodes)
        chains = [list(NodeIter(n)) for n in nodes]
        sx, sy, syaw, ex, ey, eyaw, spans = [], [], [], [], [], [], []
        atan2 = math.atan2
        for n, ch in zip(nodes, chains):
            a, k = len(sx), len(ch)
            x, y = n.point
            cx = [v.point[0] for v in ch]
            cy = [v.point[1] for v in ch]
            cyaw = [v.yaw for v in ch]
            # candidate k = Node::new(node.coord, chain[k]): same point, yaw aimed at chain[k] (src/rrt.rs:169-175);
            # only the winning candidate is materialised as a Node below
            sx += [x] * k + cx[:-1]
            sy += [y] * k + cy[:-1]
            syaw += [atan2(py - y, px - x) for px, py in zip(cx, cy)] + cyaw[:-1]
            ex += cx + cx[1:]  # own edges: chain[j] -> chain[j + 1]
            ey += cy + cy[1:]
            eyaw += cyaw + cyaw[1:]
            spans.append((a, k))
        ok = self.ctx.collide_dubins(sx, sy, syaw, ex, ey, eyaw, self.space.get_steer(), self.step_size).astype(bool)
        picks: List[int] = []
        for (a, n), ch in zip(spans, chains):
            first, own = ok[a:a + n], ok[a + n:a + 2 * n - 1]
            chain_ok = np.ones(n, bool)
            if n > 1:
                chain_ok[: n - 1] = np.logical_and.accumulate(own[::-1])[::-1]
            valid = np.nonzero(first & chain_ok)[0]
            picks.append(int(valid[-1]) if valid.size else -1)  # .rev(): the candidate closest to the root wins
        live = [j for j, k in enumerate(picks) if k >= 0]
        deeper = self._optimize_many([chains[j][picks[j]] for j in live], i + 1)
        out: List[Optional[Node]] = [None] * len(nodes)
        for j, d in zip(live, deeper):
            out[j] = Node(nodes[j].get_coord(), d if d is not None else chains[j][picks[j]])
        return out

    # -- check_finish (src/rrt.rs:428-438) for many nodes: batched optimize, one batched line_to_origin over all the
    # chains' edges, one verify launch.  Same lines as [check_finish(n) for n in nodes].
    def check_finish_many(self, nodes: List[Node]) -> List[Optional[Ring]]:
        if not nodes:
            return []
        goals = [Node.new_goal(self.goal, n, self.goal_yaw) for n in nodes]
        opt = self._optimize_many(list(nodes), 0)
        tops = [Node.new_goal(g.get_coord(), o, self.goal_yaw) if o is not None else g for g, o in zip(goals, opt)]
        edges = [_chain_edges(t) for t in tops]
        flat = [list(itertools.chain.from_iterable(e[c] for e in edges)) for c in range(6)]
        steer = self.space.get_steer()
        lines: List[Ring] = []
        if flat[0]:
            counts, plan = self.ctx.dubins_sample_count(*flat, steer, self.step_size)
            out, offsets = self.ctx.dubins_sample_fill(plan, counts)
            words = np.frombuffer(plan, np.uint8).reshape(-1, _ffi.PLAN_BYTES)[:, 104]
        pos = 0
        for e in edges:
            xs, ys = [], []
            for k in range(len(e[0])):
                i = pos + k
                if words[i] == _ffi.WORD_NONE:  # src/rrt.rs:529 (finalize's loop, not line_to_origin's :313)
                    raise RuntimeError(SHOULD_PLAN)
                else:
                    o, c = int(offsets[i]), int(counts[i])
                    xs.append(out[o:o + c, 0]); ys.append(out[o:o + c, 1])
            pos += len(e[0])
            # line_to_origin appends the root's point and finalize drops it again (:532): the samples alone, reversed
            lx = np.concatenate(xs) if xs else np.zeros(0)
            ly = np.concatenate(ys) if ys else np.zeros(0)
            lines.append((lx[::-1].copy(), ly[::-1].copy()))
        good = self.space.verify_many(lines)
        return [ln if g else None for ln, g in zip(lines, good)]

    # -- src/rrt.rs:489-501
    def optimize_from_goal(self, goal_node: Node) -> Node:
        parent = goal_node.get_parent()
        if parent is None:
            return goal_node
        n = self.optimize(parent, 0)
        return Node.new_goal(goal_node.get_coord(), n, self.goal_yaw) if n is not None else goal_node

    # -- src/rrt.rs:503-540 : root contributes nothing, result reversed (start -> goal)
    def finalize(self, goal_node: Node) -> Ring:
        top = self.optimize_from_goal(goal_node)
        lx, ly = line_to_origin(top, self.space.get_steer(), self.step_size, self.ctx, _strict=True)
        lx, ly = lx[:-1], ly[:-1]  # drop the root's own point (None => vec![] at :532)
        return lx[::-1].copy(), ly[::-1].copy()

    # -- src/rrt.rs:583-597
    def plan_one(self) -> Optional[Ring]:
        rnd = self.get_random_node()
        if rnd is not None and self.verify_node(rnd):
            self._insert(rnd)
            return self.check_finish(rnd)
        return None

    # -- the reference runs max_iter independent plan_one iterations on 4 racy workers that all see
    # a slightly stale tree (src/rrt.rs:600-609).  Here a ROUND processes `batch` samples against one tree
    # snapshot: one NN launch, one fused Dubins verify launch for the new edges (the parents' chains are already
    # verified, that is the tree invariant), one batched append, and check_finish_many for the round's fresh nodes
    # (one fused launch per optimize level).  min_by euclidean_length stays on the host (:611-617).
    def plan_rounds(self, batch: int = 256, max_iter: Optional[int] = None) -> Optional[Ring]:
        budget = self.max_iter if max_iter is None else int(max_iter)
        best, best_len = None, math.inf
        steer = self.space.get_steer()
        while budget > 0:
            b = min(batch, budget)
            budget -= b
            pts = [self.space.rand_point() for _ in range(b)]
            px, py = np.array([p[0] for p in pts]), np.array([p[1] for p in pts])
            # one call: NN -> Node::new yaw -> fused Dubins sample-and-verify of the new edges
            idx, _, ok = self.ctx.rrt_extend_dubins(px, py, steer, self.step_size)
            # idx = 0xFFFFFFFF: no nearest node (get_random_node returns None, src/rrt.rs:408-411); ok is 0 there
            fresh = [Node(p, self.nodes[int(i)]) for p, i, good in zip(pts, idx, ok.astype(bool))
                     if good and i != 0xFFFFFFFF]
            if not fresh:
                continue
            self.ctx.tree_append([c.point[0] for c in fresh], [c.point[1] for c in fresh], [c.yaw for c in fresh],
                                 [self._slot[id(c.parent)] for c in fresh])
            for c in fresh:
                self._slot[id(c)] = len(self.nodes)
                self.nodes.append(c)
            # check_finish for every fresh node, as plan_one does (:591): no goal-visibility pre-filter -- the goal
            # connects to the OPTIMIZED node, whose yaw differs from the fresh node's, so goal -> fresh being blocked
            # does not imply that the reference finds nothing
            for line in self.check_finish_many(fresh):
                if line is not None:
                    length = euclidean_length(line)
                    if length < best_len:
                        best, best_len = line, length
        return best

    # -- src/rrt.rs:599-619 (the reference's 4 racy workers become sequential iterations)
    def plan(self) -> Optional[Ring]:
        best, best_len = None, math.inf
        for _ in range(self.max_iter):
            r = self.plan_one()
            if r is not None:
                l = euclidean_length(r)
                if l < best_len:
                    best, best_len = r, l
        return best

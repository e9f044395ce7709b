//! Public API of the reference's src/rrt.rs, unchanged.  What moved to the GPU: nearest neighbour
//! (get_nearest_node), the per-edge Dubins polylines of line_to_origin / finalize, Space::verify and
//! verify_node (fused sample-and-verify per chain edge).  What stays on the host: the Arc<Node> graph, the
//! geo-offset inflation of Space::new, thread_rng sampling and the planner loops.
use crate::ffi::{self, CTX};
use geo::algorithm::euclidean_length::EuclideanLength;
use geo::{Coordinate, LineString, Point, Polygon};
use geo_offset::Offset;
use rand::{thread_rng, Rng};
use std::f64::consts::PI;
use std::sync::{Arc, Mutex};

const RECURSION_LIMIT: usize = 16;

#[allow(dead_code)]
pub struct Robot {
    width: f64,
    height: f64,
    max_steer: f64,
}

impl Robot {
    pub fn new(width: f64, height: f64, max_steer: f64) -> Robot { Robot { width, height, max_steer } }
    pub fn get_width(&self) -> f64 { self.width }
    pub fn get_steer(&self) -> f64 { self.max_steer }
}

pub fn create_circle(center: Point<f64>, radius: f64) -> Polygon<f64> {
    let (cx, cy) = center.x_y();
    let n = (2.0 * PI * radius / 1.0).ceil();
    let i = (n + 1.0) as usize;
    let points: Vec<(f64, f64)> = (0..i)
        .map(|x| {
            let x = x as f64;
            (((2.0 * PI / n * x).cos() * radius) + cx, ((2.0 * PI / n * x).sin() * radius) + cy)
        })
        .collect();
    Polygon::new(points.into(), vec![])
}

fn buffer_poly(poly: &Polygon<f64>, buffer: f64) -> Polygon<f64> {
    poly.offset(buffer).expect("polygon to set a buffer").into_iter().last().expect("should get buffered polygon")
}

fn ring_soa(p: &Polygon<f64>) -> (Vec<f64>, Vec<f64>) { p.exterior().points_iter().map(|q| q.x_y()).unzip() }

pub struct Space {
    /// every Space owns its GPU context (one world + one tree), as an RRT owns its Space and its RTree in the
    /// reference (src/rrt.rs:325-356): any number of Space / RRT pairs can live side by side
    ctx: Arc<ffi::Ctx>,
    bounds: Polygon<f64>,
    robot: Robot,
    obstacles: Vec<Polygon<f64>>,
    minx: f64,
    maxx: f64,
    miny: f64,
    maxy: f64,
}

impl Space {
    pub fn new(bounds: Polygon<f64>, robot: Robot, obstacle_list: Vec<Polygon<f64>>) -> Space {
        let ctx = ffi::Ctx::new();
        let width = robot.get_width() / 2.0;
        let bounds = buffer_poly(&bounds, -width);
        let (bx, by) = ring_soa(&bounds);
        let (mut minx, mut maxx, mut miny, mut maxy) = (bx[0], bx[0], by[0], by[0]);
        for (&x, &y) in bx.iter().zip(by.iter()) {
            if x < minx { minx = x; }
            if x > maxx { maxx = x; }
            if y < miny { miny = y; }
            if y > maxy { maxy = y; }
        }
        let obstacles: Vec<Polygon<f64>> = obstacle_list.iter().map(|o| buffer_poly(o, width)).collect();
        // device mirror of {bounds, obstacles}: rings in CSR form
        let (mut ox, mut oy, mut off) = (Vec::new(), Vec::new(), vec![0u32]);
        for o in &obstacles {
            let (x, y) = ring_soa(o);
            ox.extend(x);
            oy.extend(y);
            off.push(ox.len() as u32);
        }
        ffi::check(
            unsafe {
                ffi::pp_obstacles_upload(ctx.0, bx.as_ptr(), by.as_ptr(), bx.len(), ox.as_ptr(), oy.as_ptr(),
                                         off.as_ptr(), obstacles.len())
            },
            "obstacles_upload",
        );
        Space { ctx, bounds, robot, obstacles, minx, maxx, miny, maxy }
    }

    pub fn verify(&self, line: &LineString<f64>) -> bool {
        let (px, py): (Vec<f64>, Vec<f64>) = line.points_iter().map(|p| p.x_y()).unzip();
        let off = [0u32, px.len() as u32];
        let mut ok = 0u8;
        ffi::check(
            unsafe { ffi::pp_verify_polylines(self.ctx.0, 1, px.as_ptr(), py.as_ptr(), off.as_ptr(), &mut ok, 0) },
            "verify",
        );
        ok != 0
    }

    /// `verify` for many lines in one launch (one warp per line); same verdicts as mapping `verify`
    pub fn verify_many(&self, lines: &[LineString<f64>]) -> Vec<bool> {
        if lines.is_empty() {
            return vec![];
        }
        let (mut px, mut py, mut off) = (Vec::<f64>::new(), Vec::<f64>::new(), vec![0u32]);
        for l in lines {
            for p in l.points_iter() {
                let (x, y) = p.x_y();
                px.push(x);
                py.push(y);
            }
            off.push(px.len() as u32);
        }
        let mut ok = vec![0u8; lines.len()];
        ffi::check(
            unsafe {
                ffi::pp_verify_polylines(self.ctx.0, lines.len(), px.as_ptr(), py.as_ptr(), off.as_ptr(), ok.as_mut_ptr(), 0)
            },
            "verify_many",
        );
        ok.iter().map(|&v| v != 0).collect()
    }

    pub fn rand_point(&self) -> Point<f64> {
        let mut rng = thread_rng();
        Point::new(rng.gen_range(self.minx, self.maxx), rng.gen_range(self.miny, self.maxy))
    }

    pub fn get_steer(&self) -> f64 { self.robot.get_steer() }
    pub fn get_obs(&self) -> Vec<Polygon<f64>> { self.obstacles.clone() }
    pub fn get_bounds(&self) -> Polygon<f64> { self.bounds.clone() }
}

pub struct Node {
    point: Point<f64>,
    parent: Option<Arc<Node>>,
    yaw: f64,
    slot: Mutex<i64>, // index in the GPU tree mirror, set on insertion
}

impl Node {
    pub fn new(point: Point<f64>, parent: Arc<Node>) -> Node {
        Node { point, parent: Some(parent.clone()), yaw: compute_yaw(&point, parent.get_point()), slot: Mutex::new(-1) }
    }
    pub fn new_root(point: Point<f64>, yaw: f64) -> Node { Node { point, parent: None, yaw, slot: Mutex::new(-1) } }
    pub fn new_goal(point: Point<f64>, parent: Arc<Node>, yaw: f64) -> Node {
        Node { point, parent: Some(parent), yaw, slot: Mutex::new(-1) }
    }
    pub fn get_parent(&self) -> Option<Arc<Node>> { self.parent.clone() }
    pub fn get_above(&self) -> NodeIter { NodeIter { curr: self.parent.clone() } }
    pub fn get_point(&self) -> &Point<f64> { &self.point }
    pub fn get_coord(&self) -> Coordinate<f64> { self.point.into() }
    pub fn get_yaw(&self) -> f64 { self.yaw }
}

pub struct NodeIter {
    curr: Option<Arc<Node>>,
}

impl Iterator for NodeIter {
    type Item = Arc<Node>;
    fn next(&mut self) -> Option<Arc<Node>> {
        match self.curr.clone() {
            Some(current) => {
                self.curr = current.get_parent();
                Some(current)
            }
            None => None,
        }
    }
}

fn compute_yaw(from: &Point<f64>, to: &Point<f64>) -> f64 {
    let (fx, fy) = from.x_y();
    let (tx, ty) = to.x_y();
    (ty - fy).atan2(tx - fx)
}

/// SoA of the (child, parent) pose pairs along node -> root
fn chain_edges(node: &Arc<Node>) -> [Vec<f64>; 6] {
    let mut e: [Vec<f64>; 6] = Default::default();
    for n in (NodeIter { curr: Some(node.clone()) }) {
        if let Some(p) = n.get_parent() {
            let (sx, sy) = n.get_point().x_y();
            let (ex, ey) = p.get_point().x_y();
            e[0].push(sx); e[1].push(sy); e[2].push(n.get_yaw());
            e[3].push(ex); e[4].push(ey); e[5].push(p.get_yaw());
        }
    }
    e
}

fn push_edge(e: &mut [Vec<f64>; 6], from: &Node, to: &Node) {
    let (sx, sy) = from.get_point().x_y();
    let (ex, ey) = to.get_point().x_y();
    e[0].push(sx); e[1].push(sy); e[2].push(from.get_yaw());
    e[3].push(ex); e[4].push(ey); e[5].push(to.get_yaw());
}

/// fused Dubins sample-and-verify of m independent edges: ok[i] = Space::verify(samples of edge i ++ [its end point])
fn collide_dubins(ctx: &ffi::Ctx, e: &[Vec<f64>; 6], turn_radius: f64, step_size: f64) -> Vec<u8> {
    let m = e[0].len();
    let mut ok = vec![0u8; m];
    if m > 0 {
        ffi::check(
            unsafe {
                ffi::pp_collide_dubins(ctx.0, m, e[0].as_ptr(), e[1].as_ptr(), e[2].as_ptr(), e[3].as_ptr(),
                                       e[4].as_ptr(), e[5].as_ptr(), turn_radius, step_size, ok.as_mut_ptr(), 0)
            },
            "collide_dubins",
        );
    }
    ok
}

/// per-edge Dubins samples of the chain in node -> root order: one count pass + one fill pass on the GPU.
/// An edge without a feasible word yields its start point (line_to_origin, src/rrt.rs:313) or, with `strict`,
/// the panic of finalize's copy of the loop (src/rrt.rs:529)
fn chain_samples(e: &[Vec<f64>; 6], turn_radius: f64, step_size: f64, strict: bool) -> Vec<Vec<(f64, f64)>> {
    let m = e[0].len();
    if m == 0 {
        return vec![];
    }
    let mut counts = vec![0u32; m];
    let mut plan = vec![0u8; m * ffi::PP_DUBINS_PLAN_BYTES];
    ffi::check(
        unsafe {
            ffi::pp_dubins_sample_count(CTX.0, m, e[0].as_ptr(), e[1].as_ptr(), e[2].as_ptr(), e[3].as_ptr(),
                                        e[4].as_ptr(), e[5].as_ptr(), turn_radius, step_size, 0, counts.as_mut_ptr(),
                                        plan.as_mut_ptr() as *mut _)
        },
        "sample_count",
    );
    let mut offsets = vec![0u64; m];
    let mut total = 0u64;
    for i in 0..m {
        offsets[i] = total;
        total += counts[i] as u64;
    }
    let mut out = vec![0.0f64; 3 * total as usize];
    ffi::check(
        unsafe { ffi::pp_dubins_sample_fill(CTX.0, m, plan.as_ptr() as *const _, offsets.as_ptr(), total, out.as_mut_ptr()) },
        "sample_fill",
    );
    (0..m)
        .map(|i| {
            if plan[i * ffi::PP_DUBINS_PLAN_BYTES + ffi::PP_PLAN_WORD_OFFSET] as i32 == ffi::PP_WORD_NONE {
                if strict {
                    panic!("Should plan dubins curve"); // src/rrt.rs:529
                }
                vec![(e[0][i], e[1][i])] // src/rrt.rs:313 fallback
            } else {
                (offsets[i] as usize..(offsets[i] + counts[i] as u64) as usize).map(|k| (out[3 * k], out[3 * k + 1])).collect()
            }
        })
        .collect()
}

pub fn line_to_origin(node: Arc<Node>, turn_radius: f64, step_size: f64) -> LineString<f64> {
    let e = chain_edges(&node);
    let mut l: Vec<(f64, f64)> = chain_samples(&e, turn_radius, step_size, false).into_iter().flatten().collect();
    let mut root = node;
    while let Some(p) = root.get_parent() {
        root = p;
    }
    l.push(root.get_coord().x_y());
    l.into()
}

pub struct RRT {
    goal: Coordinate<f64>,
    goal_yaw: f64,
    max_iter: usize,
    step_size: f64,
    space: Space,
    nodes: Arc<Mutex<Vec<Arc<Node>>>>, // slot i <-> GPU tree slot i (replaces Arc<Mutex<RTree>>)
}

impl RRT {
    pub fn new(start: Coordinate<f64>, start_yaw: f64, goal: Coordinate<f64>, goal_yaw: f64, max_iter: usize,
               step_size: f64, space: Space) -> RRT {
        let root = Arc::new(Node::new_root(start.into(), start_yaw));
        *root.slot.lock().unwrap() = 0;
        let par = -1i32;
        ffi::check(unsafe { ffi::pp_tree_upload(space.ctx.0, 1, &start.x, &start.y, &start_yaw, &par) }, "tree_upload");
        RRT { goal, goal_yaw, max_iter, step_size, space, nodes: Arc::new(Mutex::new(vec![root])) }
    }

    pub fn get_nearest_node(&self, point: &Point<f64>) -> Option<Arc<Node>> {
        let (x, y) = point.x_y();
        let mut idx = 0xFFFF_FFFFu32;
        // the tree may grow between the query and the lookup (as under the reference's Mutex); indices stay valid
        ffi::check(unsafe { ffi::pp_nn(self.space.ctx.0, 1, &x, &y, &mut idx, std::ptr::null_mut(), 0) }, "nn");
        if idx == 0xFFFF_FFFF { None } else { Some(self.nodes.lock().unwrap()[idx as usize].clone()) }
    }

    pub fn get_random_node(&self) -> Option<Arc<Node>> {
        let point = self.space.rand_point();
        match self.get_nearest_node(&point) {
            Some(nearest_node) => Some(Arc::new(Node::new(point, nearest_node))),
            None => None,
        }
    }

    pub fn verify_node(&self, node: Arc<Node>) -> bool {
        let e = chain_edges(&node);
        let m = e[0].len();
        if m == 0 {
            return self.space.verify(&vec![node.get_coord().x_y()].into());
        }
        collide_dubins(&self.space.ctx, &e, self.space.get_steer(), self.step_size).iter().all(|&v| v != 0)
    }

    pub fn check_finish(&self, node: Arc<Node>) -> Option<LineString<f64>> {
        let goal_node = Arc::new(Node::new_goal(self.goal.into(), node, self.goal_yaw));
        let line = self.finalize(goal_node);
        if self.space.verify(&line) { Some(line) } else { None }
    }

    pub fn optimize(&self, node: Arc<Node>, i: usize) -> Option<Arc<Node>> {
        if i >= RECURSION_LIMIT {
            return None;
        }
        let nodes_vec: Vec<Arc<Node>> = (NodeIter { curr: Some(node.clone()) }).collect();
        for to_node in nodes_vec.into_iter().rev() {
            let new_node = Arc::new(Node::new(node.get_coord().into(), to_node.clone()));
            if self.verify_node(new_node.clone()) {
                match self.optimize(to_node, i + 1) {
                    Some(to_node) => return Some(Arc::new(Node::new(node.get_coord().into(), to_node))),
                    None => return Some(new_node),
                }
            }
        }
        None
    }

    /// `optimize` (src/rrt.rs:463-487) for MANY start nodes: the shortcut candidates of all nodes of a recursion
    /// level go into one fused launch.  verify(line_to_origin(candidate k)) = verify(edge candidate -> chain[k])
    /// AND verify(chain of chain[k]), and the chain verdicts are suffix-ANDs over the ancestors' own edges, so
    /// 2*depth - 1 edges per node replace depth^2.  Per node: the same candidates, root-first order and verdicts.
    pub fn optimize_many(&self, nodes: &[Arc<Node>], i: usize) -> Vec<Option<Arc<Node>>> {
        if i >= RECURSION_LIMIT || nodes.is_empty() {
            return vec![None; nodes.len()];
        }
        let chains: Vec<Vec<Arc<Node>>> =
            nodes.iter().map(|n| (NodeIter { curr: Some(n.clone()) }).collect()).collect(); // node, parent, ..., root
        let cands: Vec<Vec<Arc<Node>>> = nodes
            .iter()
            .zip(chains.iter())
            .map(|(n, ch)| ch.iter().map(|to| Arc::new(Node::new(n.get_coord().into(), to.clone()))).collect())
            .collect();
        let mut e: [Vec<f64>; 6] = Default::default();
        let mut spans: Vec<(usize, usize)> = Vec::with_capacity(nodes.len());
        for (ch, cd) in chains.iter().zip(cands.iter()) {
            spans.push((e[0].len(), ch.len()));
            for (c, to) in cd.iter().zip(ch.iter()) {
                push_edge(&mut e, c, to); // candidate -> chain[k]
            }
            for v in ch.iter().take(ch.len() - 1) {
                let p = v.get_parent().expect("only the root has no parent");
                push_edge(&mut e, v, &p); // chain[k] -> chain[k + 1]
            }
        }
        let ok = collide_dubins(&self.space.ctx, &e, self.space.get_steer(), self.step_size);
        let mut picks: Vec<Option<usize>> = Vec::with_capacity(nodes.len());
        for &(a, n) in spans.iter() {
            // .rev(): the valid candidate closest to the root wins; walking down from the root, the chain below
            // candidate k is good while every own edge k.. verified
            let mut pick = None;
            for k in (0..n).rev() {
                if k + 1 < n && ok[a + n + k] == 0 {
                    break;
                }
                if ok[a + k] != 0 {
                    pick = Some(k);
                    break;
                }
            }
            picks.push(pick);
        }
        let live: Vec<usize> = (0..nodes.len()).filter(|&j| picks[j].is_some()).collect();
        let next: Vec<Arc<Node>> = live.iter().map(|&j| chains[j][picks[j].unwrap()].clone()).collect();
        let deeper = self.optimize_many(&next, i + 1);
        let mut out: Vec<Option<Arc<Node>>> = vec![None; nodes.len()];
        for (&j, d) in live.iter().zip(deeper.into_iter()) {
            out[j] = Some(match d {
                Some(to_node) => Arc::new(Node::new(nodes[j].get_coord().into(), to_node)),
                None => cands[j][picks[j].unwrap()].clone(),
            });
        }
        out
    }

    /// `check_finish` (src/rrt.rs:428-438) for many nodes: batched optimize, ONE count + fill over the edges of all
    /// final chains, ONE verify launch.  Same lines as mapping `check_finish`.
    pub fn check_finish_many(&self, nodes: &[Arc<Node>]) -> Vec<Option<LineString<f64>>> {
        if nodes.is_empty() {
            return vec![];
        }
        let opt = self.optimize_many(nodes, 0);
        let tops: Vec<Arc<Node>> = nodes
            .iter()
            .zip(opt.into_iter())
            .map(|(n, o)| Arc::new(Node::new_goal(self.goal.into(), o.unwrap_or_else(|| n.clone()), self.goal_yaw)))
            .collect();
        let mut e: [Vec<f64>; 6] = Default::default();
        let mut spans: Vec<(usize, usize)> = Vec::with_capacity(tops.len());
        for t in tops.iter() {
            let te = chain_edges(t);
            spans.push((e[0].len(), te[0].len()));
            for c in 0..6 {
                e[c].extend_from_slice(&te[c]);
            }
        }
        let per_edge = chain_samples(&e, self.space.get_steer(), self.step_size, true);
        let lines: Vec<LineString<f64>> = spans
            .iter()
            .map(|&(a, n)| {
                let mut l: Vec<(f64, f64)> = per_edge[a..a + n].iter().flatten().cloned().collect();
                l.reverse(); // the root contributes nothing (None => vec![], src/rrt.rs:532); start -> goal
                l.into()
            })
            .collect();
        let good = self.space.verify_many(&lines);
        lines.into_iter().zip(good.into_iter()).map(|(l, g)| if g { Some(l) } else { None }).collect()
    }

    pub fn optimize_from_goal(&self, goal_node: Arc<Node>) -> Arc<Node> {
        match goal_node.get_parent() {
            Some(parent) => match self.optimize(parent, 0) {
                Some(n) => Arc::new(Node::new_goal(goal_node.get_coord().into(), n, self.goal_yaw)),
                None => goal_node,
            },
            None => goal_node,
        }
    }

    pub fn finalize(&self, goal_node: Arc<Node>) -> LineString<f64> {
        let top = self.optimize_from_goal(goal_node);
        let e = chain_edges(&top);
        // the root contributes nothing here (None => vec![]); a chain edge without a feasible word panics as the
        // reference does (src/rrt.rs:529)
        let mut l: Vec<(f64, f64)> = chain_samples(&e, self.space.get_steer(), self.step_size, true).into_iter().flatten().collect();
        l.reverse();
        l.into()
    }

    pub fn plan_one(&self) -> Option<LineString<f64>> {
        if let Some(rnd_node) = self.get_random_node() {
            if self.verify_node(rnd_node.clone()) {
                {
                    let mut nodes = self.nodes.lock().unwrap();
                    let slot = nodes.len() as i64;
                    *rnd_node.slot.lock().unwrap() = slot;
                    let (x, y) = rnd_node.get_point().x_y();
                    let yaw = rnd_node.get_yaw();
                    let par = rnd_node.get_parent().map(|p| *p.slot.lock().unwrap() as i32).unwrap_or(-1);
                    ffi::check(unsafe { ffi::pp_tree_append(self.space.ctx.0, 1, &x, &y, &yaw, &par) }, "tree_append");
                    nodes.push(rnd_node.clone());
                }
                if let Some(finish) = self.check_finish(rnd_node) {
                    return Some(finish);
                }
            }
        }
        None
    }

    /// SURVEY 8f-3.  The reference runs max_iter independent plan_one iterations on 4 racy workers that all see a
    /// slightly stale tree (src/rrt.rs:600-609).  Here a ROUND takes `batch` samples against one tree snapshot: one
    /// call for NN + Node::new yaw + fused Dubins sample-and-verify of the new edges (the parents' chains are
    /// verified already: the tree invariant), one batched append, one fused launch for the goal connections and
    /// the batched goal check for the nodes that see the goal.  min_by length stays on the host (:611-617).
    pub fn plan_rounds(&self, batch: usize) -> Option<LineString<f64>> {
        let steer = self.space.get_steer();
        let mut best: Option<(f64, LineString<f64>)> = None;
        let mut budget = self.max_iter;
        while budget > 0 {
            let b = batch.max(1).min(budget);
            budget -= b;
            let pts: Vec<Point<f64>> = (0..b).map(|_| self.space.rand_point()).collect();
            let (px, py): (Vec<f64>, Vec<f64>) = pts.iter().map(|p| p.x_y()).unzip();
            let (mut idx, mut yaw, mut ok) = (vec![0u32; b], vec![0f64; b], vec![0u8; b]);
            ffi::check(
                unsafe {
                    ffi::pp_rrt_extend_dubins(self.space.ctx.0, b, px.as_ptr(), py.as_ptr(), steer, self.step_size,
                                              idx.as_mut_ptr(), yaw.as_mut_ptr(), ok.as_mut_ptr(), 0, 0)
                },
                "rrt_extend_dubins",
            );
            let mut nodes = self.nodes.lock().unwrap();
            let fresh: Vec<Arc<Node>> = (0..b)
                .filter(|&k| ok[k] != 0 && idx[k] != 0xFFFF_FFFF)
                .map(|k| Arc::new(Node::new(pts[k], nodes[idx[k] as usize].clone())))
                .collect();
            if fresh.is_empty() {
                continue;
            }
            let (mut fx, mut fy, mut fyaw, mut fpar) =
                (Vec::<f64>::new(), Vec::<f64>::new(), Vec::<f64>::new(), Vec::<i32>::new());
            for c in fresh.iter() {
                let (x, y) = c.get_point().x_y();
                fx.push(x);
                fy.push(y);
                fyaw.push(c.get_yaw());
                fpar.push(c.get_parent().map(|p| *p.slot.lock().unwrap() as i32).unwrap_or(-1));
            }
            ffi::check(
                unsafe { ffi::pp_tree_append(self.space.ctx.0, fresh.len(), fx.as_ptr(), fy.as_ptr(), fyaw.as_ptr(), fpar.as_ptr()) },
                "tree_append",
            );
            for c in fresh.iter() {
                *c.slot.lock().unwrap() = nodes.len() as i64;
                nodes.push(c.clone());
            }
            drop(nodes);
            // check_finish for every fresh node, as plan_one does (src/rrt.rs:591): the goal connects to the OPTIMIZED
            // node, whose yaw differs from the fresh node's, so a blocked goal -> fresh edge decides nothing
            let reach: Vec<Arc<Node>> = fresh.clone();
            for line in self.check_finish_many(&reach).into_iter().flatten() {
                let len = line.euclidean_length();
                if best.as_ref().map_or(true, |cur| len < cur.0) {
                    best = Some((len, line));
                }
            }
        }
        best.map(|found| found.1)
    }

    pub fn plan(&self) -> Option<LineString<f64>> {
        // sequential, one sample per iteration as in the reference; plan_rounds() is the batched form
        (0..self.max_iter)
            .filter_map(|_| self.plan_one())
            .min_by(|a, b| a.euclidean_length().partial_cmp(&b.euclidean_length()).expect("should compared route costs"))
    }
}

/*
 * pp_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the hot path of tsturzl/rs-pathplanning
 * (crate `pathplanning` v0.1.2): src/dubins.rs in full and the NN / edge /
 * verify pieces of src/rrt.rs, plus the geo 0.12.2 predicates the latter
 * calls.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path
 * (libpathplanning_b200.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 *  and cannot be compiled here (no rustc/cargo), and
 * geo/rstar are not vendored.  The restatement is pinned only by (i) an
 * independent pure-Python transliteration (oracle/dubins_py.py) that must
 * agree bit-for-bit, and (ii) the restatement-derived known answers of
 * SURVEY.md Appendix C.
 *
 * Build: gcc -O3 -ffp-contract=off -fno-fast-math -fopenmp (see oracle/Makefile).
 * glibc libm is the same libm Rust's f64::{sin,cos,atan2,acos,hypot} reach on
 * x86_64-unknown-linux-gnu, and no a*b+c is ever contracted, as in rustc.
 */
#ifndef PP_ORACLE_H
#define PP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* word ids follow ALL_PLANNERS order, src/dubins.rs:291 */
enum { PPO_LSL = 0, PPO_RSR = 1, PPO_LSR = 2, PPO_RSL = 3, PPO_RLR = 4, PPO_LRL = 5, PPO_NONE = 0xFF };

/* flags reported by ppo_dubins_eval for the parity harness */
enum {
    PPO_FLAG_NEAR_WRAP = 1, /* a mod2pi result of a contending word lies within PPO_WRAP_EPS of 0 or 2pi */
    PPO_FLAG_NEAR_TIE = 2,  /* runner-up cost within PPO_TIE_REL (relative) of the best cost */
    PPO_FLAG_NEAR_FEAS = 4, /* a feasibility test (p^2 >= 0, |tmp| <= 1) is within 1e-9 of flipping */
    /* sampled paths / edges (ppo_dubins_path_flags, ppo_verify_dubins_edges_flags, ppo_optimize, ppo_check_finish) */
    PPO_FLAG_NEAR_COUNT = 8,  /* a count-deciding comparison of generate_local_course (the loop test at
                                 src/dubins.rs:239, the zero test of the trim at :281-288) is within
                                 PPO_COUNT_REL of flipping */
    PPO_FLAG_NEAR_GRAZE = 16, /* the verify verdict's margin (ppo_verify_margin) is below the tolerance */
    PPO_FLAG_HUGE_ANGLE = 32  /* |yaw| > PPO_HUGE_ANGLE: one ulp of the angle moves samples by more than 1e-10 */
};
#define PPO_COUNT_REL 4e-9
#define PPO_HUGE_ANGLE 1048576.0
#define PPO_WRAP_EPS 1e-9
#define PPO_TIE_REL 1e-9

double ppo_mod2pi(double theta);  /* src/dubins.rs:14-20 */
double ppo_pi_2_pi(double angle); /* src/dubins.rs:22-24 */

/* one word: returns 1 if feasible and writes tpq[3]; src/dubins.rs:27-153 */
int ppo_dubins_word(int word, double alpha, double beta, double d, double tpq[3]);

/* HARNESS: margin of the word's feasibility test relative to 1 + d^2 (negative = infeasible) */
double ppo_dubins_word_margin(int word, double alpha, double beta, double d);

/* normalisation + six words + selection (src/dubins.rs:401-408, 333-363).
 * Returns the word id (0..5) or PPO_NONE.  cost is radius-normalised. */
int ppo_dubins_eval(double sx, double sy, double syaw, double ex, double ey, double eyaw, double radius,
                    double *cost, double tpq[3], uint32_t *flags);

/* batch (n pairs, SoA); radius_arr may be NULL -> scalar radius.  nthreads<=0: all cores */
void ppo_dubins_eval_batch(size_t n, const double *sx, const double *sy, const double *syaw, const double *ex,
                           const double *ey, const double *eyaw, const double *radius_arr, double radius,
                           double *cost, uint8_t *word, double *tpq, uint32_t *flags, int nthreads);

/* full path, world frame: src/dubins.rs:401-428.  from_origin != 0 -> src/dubins.rs:326-399 semantics
 * with (ex,ey,eyaw) the local goal and sx=sy=syaw ignored, c = 1/radius.
 * Returns the number of samples (>= 0), -1 if no feasible word (None), -2 if cap too small,
 * -3 if the reference would index out of its n_point buffer (panic).
 * n_point_out (optional) receives the reference's buffer size (src/dubins.rs:369). */
long ppo_dubins_path(double sx, double sy, double syaw, double ex, double ey, double eyaw, double radius,
                     double step, int from_origin, double *px, double *py, double *pyaw, size_t cap,
                     int *word, double *cost, long *n_point_out);

/* the same path together with the harness flags (PPO_FLAG_*) of its evaluation and its sample count */
long ppo_dubins_path_flags(double sx, double sy, double syaw, double ex, double ey, double eyaw, double radius,
                           double step, int from_origin, double *px, double *py, double *pyaw, size_t cap,
                           int *word, double *cost, uint32_t *flags);

/* sample counts only (same loop semantics), batch */
void ppo_dubins_count_batch(size_t n, const double *sx, const double *sy, const double *syaw, const double *ex,
                            const double *ey, const double *eyaw, double radius, double step, int64_t *counts,
                            int nthreads);

/* create_circle, src/rrt.rs:43-60, followed by Polygon::new's ring closing (geo-types 0.4).
 * Returns the number of ring points written (<= cap) or -1. */
long ppo_create_circle(double cx, double cy, double radius, double *rx, double *ry, size_t cap);

/* compute_yaw, src/rrt.rs:267-271 */
double ppo_compute_yaw(double fx, double fy, double tx, double ty);

/* exact nearest neighbour: argmin_i (dx*dx + dy*dy), non-fused, lowest index wins ties
 * (the intended contract of src/rrt.rs:378-391).  Also reports, in
 * hypot_disagree (optional), how many queries have argmin(hypot) != argmin(d2). */
void ppo_nn_brute(size_t n_nodes, const double *nx, const double *ny, size_t m, const double *qx,
                  const double *qy, uint32_t *idx, double *d2, size_t *hypot_disagree, int nthreads);
/* same answer through a uniform grid (the fair CPU comparator for rstar's O(log N) query) */
void ppo_nn_grid(size_t n_nodes, const double *nx, const double *ny, size_t m, const double *qx, const double *qy,
                 uint32_t *idx, double *d2, int nthreads);

/* ---- geo 0.12.2 predicates on closed rings given as point arrays ---- */
/* LineString::contains(&Point) */
int ppo_ring_has_point(const double *rx, const double *ry, size_t n, double px, double py);
/* get_position: 0 outside, 1 inside, 2 on boundary */
int ppo_point_position(const double *rx, const double *ry, size_t n, double px, double py);
/* LineString::intersects(&LineString): a = ring (outer loop), b = line */
int ppo_lines_intersect(const double *ax, const double *ay, size_t na, const double *bx, const double *by,
                        size_t nb);

/* world = bounds ring + obstacle rings (CSR).  No interior rings (the reference never builds any). */
typedef struct {
    const double *bx, *by;
    size_t nb; /* bounds exterior ring, closed */
    const double *ox, *oy;
    const uint32_t *ring_off; /* n_rings+1 offsets into ox/oy */
    size_t n_rings;
} ppo_world;

/* Space::verify, src/rrt.rs:124-137, on one polyline (n >= 0 points) */
int ppo_verify(const ppo_world *w, const double *lx, const double *ly, size_t n);
/* same, with the exactness-preserving AABB culls of SURVEY B.1 (must agree with ppo_verify) */
int ppo_verify_culled(const ppo_world *w, const double *lx, const double *ly, size_t n);

/* HARNESS: the circle-filter class of segment a-b against one ring in the culled loop: 0 = the ring is skipped,
 * 1 = blocked (both end points inside the ring's inner circle), 2 = the exact predicates decide */
int ppo_circle_class(const double *rx, const double *ry, size_t n, double ax, double ay, double bx, double by);

/* straight 2-point edges a->b, batch */
void ppo_verify_segments(const ppo_world *w, size_t m, const double *ax, const double *ay, const double *bx,
                         const double *by, uint8_t *ok, int culled, int nthreads);

/* Dubins edge polyline = samples(child->parent) ++ [parent point]  (SURVEY A.3 Q6/Q12);
 * fallback [(sx,sy), parent] when no word is feasible (src/rrt.rs:313). Returns point count or <0. */
long ppo_dubins_edge_polyline(double sx, double sy, double syaw, double ex, double ey, double eyaw,
                              double radius, double step, double *lx, double *ly, size_t cap);
/* verify of Dubins edges, batch; min_clear (optional) is not computed here */
void ppo_verify_dubins_edges(const ppo_world *w, size_t m, const double *sx, const double *sy, const double *syaw,
                             const double *ex, const double *ey, const double *eyaw, double radius, double step,
                             uint8_t *ok, int culled, int nthreads);

/* HARNESS (not in the reference): Space::verify plus a lower bound on how far the line's vertices can move
 * without changing the verdict (free: clearance to every ring boundary; blocked: penetration depth). */
int ppo_verify_margin(const ppo_world *w, const double *lx, const double *ly, size_t n, double *margin);
/* verify of Dubins edges with DECISION flags: word-level path flags (wrap / tie / feasibility / huge angle);
 * PPO_FLAG_NEAR_GRAZE when margin < graze_tol * max(1, max |coordinate|); PPO_FLAG_NEAR_COUNT only when the
 * margin is also below the sagitta one added / dropped sample can move the polyline by.  margins may be NULL. */
void ppo_verify_dubins_edges_flags(const ppo_world *w, size_t m, const double *sx, const double *sy,
                                   const double *syaw, const double *ex, const double *ey, const double *eyaw,
                                   double radius, double step, double graze_tol, uint8_t *ok, uint32_t *flags,
                                   double *margins, int culled, int nthreads);

/* RRT::optimize (src/rrt.rs:463-487, RECURSION_LIMIT :14) of tree node `node` over a flat tree.  Writes the chain
 * of the returned node (new node ... root) as poses and returns its length; 0 = None; -2 cap too small; -3 error.
 * flags: OR of the harness flags of every verify decision taken; verifies: number of Space::verify calls. */
long ppo_optimize(const ppo_world *w, size_t n_nodes, const double *nx, const double *ny, const double *nyaw,
                  const int32_t *parent, uint32_t node, double radius, double step, double graze_tol, double *cx,
                  double *cy, double *cyaw, size_t cap, uint32_t *flags, long *verifies);
/* RRT::check_finish (src/rrt.rs:428-438) through finalize / optimize_from_goal (:489-540).  Returns the point
 * count of the line (start -> goal), -1 = None (line written, does not verify), -2 capacity, -3 panic.
 * flags: decisions (which chain, which verdict) that are fragile; line_flags: raw path flags of the final line's
 * edges (a knife-edge count there changes the number of samples, not the decision). */
long ppo_check_finish(const ppo_world *w, size_t n_nodes, const double *nx, const double *ny, const double *nyaw,
                      const int32_t *parent, uint32_t node, double gx, double gy, double gyaw, double radius,
                      double step, double graze_tol, double *lx, double *ly, size_t cap, long *line_len, double *cx,
                      double *cy, double *cyaw, size_t ccap, long *chain_len, uint32_t *flags, uint32_t *line_flags,
                      long *verifies);

/* line_to_origin over a flat tree (src/rrt.rs:291-321, node->root chunk order). Returns points or <0. */
long ppo_line_to_origin(const double *nx, const double *ny, const double *nyaw, const int32_t *parent,
                        uint32_t node, double radius, double step, double *lx, double *ly, size_t cap);

/* counter-based generator of SURVEY 8(d): uniform in [0,1) for (seed, stream, i) */
double ppo_uniform(uint64_t seed, uint64_t stream, uint64_t i);
void ppo_fill_uniform(uint64_t seed, uint64_t stream, size_t n, double lo, double hi, double *out);

int ppo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif

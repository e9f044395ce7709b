/*
 * pathplanning_b200.h -- C-ABI of libpathplanning_b200.so
 *
 * B200-native (sm_100a CUDA) replacement for the data-parallel hot path of the Rust
 * crate `pathplanning` v0.1.2 (tsturzl/rs-pathplanning): batched Dubins shortest-path
 * evaluation / sampling and the RRT extend step (nearest neighbour + edge-vs-obstacle
 * verification).  The reference has no FFI of its own; each entry point below names
 * the reference function (file:line under the crate root) whose arithmetic it
 * replaces and that the Rust shim (rs-pathplanning_b200/rust/) binds it under.
 *
 * Conventions
 *   - plain C types only; every buffer is caller-allocated and caller-owned; the
 *     library owns only pp_ctx and its device mirrors of tree / obstacles.
 *   - functions without a suffix take HOST pointers, copy in/out and are synchronous.
 *     `_dev` twins take DEVICE pointers (on the ctx's device), enqueue on the ctx
 *     stream and return without synchronising (use pp_sync or the stream itself).
 *   - return value: 0 = PP_OK, negative = error (never aborts, never unwinds);
 *     pp_last_error(ctx) gives a message for the last failure on that ctx.
 *   - there is NO CPU fallback: pp_ctx_create fails with PP_ERR_NO_DEVICE when no
 *     sm_100 GPU is present.
 *   - a pp_ctx may be used from several host threads (calls are serialised by an
 *     internal mutex, matching the `&self` + rayon use at src/rrt.rs:600-609).
 *   - numeric contract: NN indices and straight-edge verify flags are bit-exact with
 *     the oracle (non-fused f64, lowest index on ties); Dubins costs / samples agree
 *     to 1e-9 relative (the kernels' own f64 sincos / atan2 / acos are within 1.5 ulp of glibc's).
 */
#ifndef PATHPLANNING_B200_H
#define PATHPLANNING_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_ABI_VERSION 2

typedef struct pp_ctx pp_ctx;
typedef struct pp_group pp_group; /* one host process driving several B200s (section "multi-GPU" below) */

enum pp_status {
    PP_OK = 0,
    PP_ERR_INVALID = -1,   /* bad argument (null pointer, non-positive radius/step, ...) */
    PP_ERR_NO_DEVICE = -2, /* no CUDA device of compute capability 10.x */
    PP_ERR_CUDA = -3,      /* a CUDA runtime call or kernel failed; see pp_last_error */
    PP_ERR_NOMEM = -4,     /* device or pinned-host allocation failed */
    PP_ERR_STATE = -5,     /* tree / obstacles not uploaded yet */
    PP_ERR_OVERFLOW = -6,  /* an output capacity given by the caller is too small */
    PP_ERR_COMM = -7       /* NCCL is unavailable (libnccl.so.2 not loadable) or a collective failed */
};

/* word ids: ALL_PLANNERS order, src/dubins.rs:291 (LSL, RSR, LSR, RSL, RLR, LRL) */
enum pp_word { PP_LSL = 0, PP_RSR = 1, PP_LSR = 2, PP_RSL = 3, PP_RLR = 4, PP_LRL = 5, PP_WORD_NONE = 0xFF };
/* segment modes of a word, src/dubins.rs:4-9 */
enum pp_mode { PP_MODE_L = 0, PP_MODE_S = 1, PP_MODE_R = 2 };

/* flags for the collision entry points */
enum pp_collide_flags {
    PP_COLLIDE_DEFAULT = 0,  /* straight edges: broad phase through the uniform obstacle grid (fastest; same flags) */
    PP_COLLIDE_NO_CULL = 1,  /* exhaustive segment-pair loop exactly as geo's, no AABB rejection */
    PP_COLLIDE_USE_GRID = 2, /* the obstacle grid, explicitly */
    PP_COLLIDE_UNSORTED = 4, /* tiled AABB scan with one edge per thread in the caller's order (no binning) */
    PP_COLLIDE_SCAN = 8,     /* tiled scan of ALL ring boxes through shared memory (TMA tiles), edges binned by start
                                point, 32 boxes per instruction against the warp's union box, hit flags by ballot */
    PP_COLLIDE_FUSED = 16    /* pp_rrt_extend only: NN + yaw + verify in ONE cell-coherent launch, the queries binned by
                                node-grid block first (three small launches).  Same bits as the default pair of
                                launches; pays off when the binning is amortised, see DESIGN.md section 3 */
};

/* flags for pp_nn */
/* Every method returns the same bit-exact argmin of dx*dx + dy*dy with the lowest index on ties. */
enum pp_nn_flags {
    PP_NN_DEFAULT = 0,   /* automatic: the grid search for trees of >= PP_NN_GRID_MIN_NODES nodes, otherwise the
                            brute-force scans */
    PP_NN_PLAIN_F64 = 1, /* tiled brute-force scan, every pair in f64 (the yard-stick kernel) */
    PP_NN_GRID = 2,      /* exact uniform-grid search.  The grid is incremental: nodes appended since the last build
                            form a tail that every query scans linearly; the O(n) device-side rebuild runs once per
                            4 096 appended nodes (or when queries x tail outweighs it), never once per append */
    PP_NN_UNSORTED = 4,  /* tiled scan with per-thread fp32 pre-rejection, queries in the caller's order */
    PP_NN_SCAN = 8       /* tiled brute-force scan over ALL nodes, queries binned into a G x G grid, warp-wide exact
                            fp32 pre-rejection (node-parallel variant for <= 64 queries); no index on the tree */
};
#define PP_NN_GRID_MIN_NODES 4096

/* bytes of the opaque per-path plan record produced by pp_dubins_sample_count */
#define PP_DUBINS_PLAN_BYTES 112

/* ------------------------------------------------------------------ context */
int pp_abi_version(void);
const char *pp_status_string(int status);
/* number of usable (sm_100) devices; 0 when none (no CPU fallback exists) */
int pp_device_count(void);
/* crate-level lazy static in the Rust shim.  One context = one device = one stream. */
int pp_ctx_create(int device, pp_ctx **out);
void pp_ctx_destroy(pp_ctx *ctx);
const char *pp_last_error(pp_ctx *ctx);
int pp_ctx_device(pp_ctx *ctx);
int pp_ctx_sm_count(pp_ctx *ctx);
/* the cudaStream_t all _dev calls are enqueued on (for event timing by the caller) */
void *pp_ctx_stream(pp_ctx *ctx);
/* run the _dev calls on a caller-owned cudaStream_t instead (NULL restores the ctx's own stream).  A context keeps
 * small device-side scratch (last-block tickets, the work counter of the sample fill): calls of ONE context must stay
 * ordered -- when switching streams, order the new stream after the old one (event / pp_sync) before the next call. */
int pp_ctx_set_stream(pp_ctx *ctx, void *cuda_stream);
int pp_sync(pp_ctx *ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
uint64_t pp_launch_count(pp_ctx *ctx);
/* number of O(n) rebuilds of the node grid so far (the incremental grid keeps this far below the append count) */
uint64_t pp_nn_grid_builds(pp_ctx *ctx);
/* pinned host memory for fast, overlappable transfers through the host entry points */
int pp_host_alloc(size_t bytes, void **out);
int pp_host_free(void *p);

/* ------------------------------------------------------------------ Dubins */
/* mod2pi / pi_2_pi on the device, element-wise (src/dubins.rs:14-24); test/diagnostic entry */
int pp_mod2pi(pp_ctx *ctx, size_t n, const double *x, double *out, int pi_2_pi);

/* all six words for explicit (alpha, beta, d): src/dubins.rs:27-153 (lsl, rsr, lsr, rsl, rlr, lrl).
 * tpq: n*6*3 doubles (NaN where infeasible); feasible: n*6 bytes.
 * Domain: alpha, beta in [0, 2*pi] and 0 <= d < 1e5 -- what the reference's own caller passes (both angles come out of
 * mod2pi, src/dubins.rs:336-338); values outside are refused with PP_ERR_INVALID (NaN is accepted: infeasible). */
int pp_dubins_words(pp_ctx *ctx, size_t n, const double *alpha, const double *beta, const double *d, double *tpq,
                    uint8_t *feasible);

/* dubins_path_planning's evaluation half (src/dubins.rs:401-408 + 326-363): frame change, six words,
 * strict-< in-order minimum.  The frame change computes theta = atan2(dy, dx) - syaw (a few ulp from the reference's
 * rotate-then-atan2), and the kernels' sincos / atan2 / acos are within 1.5 ulp of glibc's: for a pair whose alpha,
 * t or q sits within ~1e-9 of a mod2pi wrap (RRT edges always have alpha at the 0 / 2*pi wrap), whose two best words
 * tie to 1e-9, or whose word is within 1e-9 of infeasible, the chosen word -- hence the sample count -- may differ
 * from the reference's; everywhere else word, cost and samples agree to 1e-9 relative.  radius_arr may be NULL -> scalar `radius` (DubinsConfig.turn_radius).
 * cost is radius-normalised as in the reference (src/dubins.rs:395); word = pp_word or PP_WORD_NONE
 * (reference returns None, src/dubins.rs:397; cost = +inf).  tpq (n*3) may be NULL.
 * Host memory of any kind: pinned buffers (pp_host_alloc) are copied from / to directly; batches of >= 2^21 pairs in
 * ordinary pageable memory (Vec<f64>, malloc) are staged through the context's pinned ring by a few copy threads
 * (PP_STAGE_THREADS overrides their number), about 3x the rate of handing such memory to cudaMemcpyAsync. */
int pp_dubins_eval(pp_ctx *ctx, size_t n, const double *sx, const double *sy, const double *syaw,
                   const double *ex, const double *ey, const double *eyaw, const double *radius_arr,
                   double radius, double *cost, uint8_t *word, double *tpq);
int pp_dubins_eval_dev(pp_ctx *ctx, size_t n, const double *sx, const double *sy, const double *syaw,
                       const double *ex, const double *ey, const double *eyaw, const double *radius_arr,
                       double radius, double *cost, uint8_t *word, double *tpq);

/* generate_local_course + trim (src/dubins.rs:200-289, 365-395) split in two passes.
 * count: evaluates each pair, walks the three segments with the reference's accumulated `pd`
 * loop and writes counts[i] = number of samples the reference returns (0 when no word is
 * feasible or the path is empty) plus an opaque plan record per path.
 * from_origin != 0 selects dubins_path_planning_from_origin semantics (src/dubins.rs:326):
 * (ex,ey,eyaw) is the local goal, output stays in the start frame. */
int pp_dubins_sample_count(pp_ctx *ctx, size_t n, const double *sx, const double *sy, const double *syaw,
                           const double *ex, const double *ey, const double *eyaw, double radius, double step,
                           int from_origin, uint32_t *counts, void *plan /* n*PP_DUBINS_PLAN_BYTES */);
int pp_dubins_sample_count_dev(pp_ctx *ctx, size_t n, const double *sx, const double *sy, const double *syaw,
                               const double *ex, const double *ey, const double *eyaw, double radius, double step,
                               int from_origin, uint32_t *counts, void *plan);
/* fill: writes interleaved (x, y, yaw) f64 triples of path i at out[3*offsets[i] ...]
 * (interpolate + back-transform, src/dubins.rs:155-198, 412-422). offsets = exclusive prefix sum of counts. */
int pp_dubins_sample_fill(pp_ctx *ctx, size_t n, const void *plan, const uint64_t *offsets, uint64_t total,
                          double *out_xyyaw);
int pp_dubins_sample_fill_dev(pp_ctx *ctx, size_t n, const void *plan, const uint64_t *offsets, uint64_t total,
                              double *out_xyyaw);
/* exclusive prefix sum of counts on the device (helper for the two passes) */
int pp_exclusive_scan_u32_dev(pp_ctx *ctx, size_t n, const uint32_t *counts, uint64_t *offsets, uint64_t *total);

/* scalar convenience used by the shim's dubins_path_planning / _from_origin (src/dubins.rs:326, 401):
 * one path into caller buffers of capacity cap; *n_out = samples, *word = pp_word or PP_WORD_NONE. */
int pp_dubins_path(pp_ctx *ctx, double sx, double sy, double syaw, double ex, double ey, double eyaw,
                   double radius, double step, int from_origin, double *px, double *py, double *pyaw, size_t cap,
                   size_t *n_out, int *word, double *cost);

/* ------------------------------------------------------------------ RRT: tree + obstacles */
/* flat SoA mirror of the reference's Arc<Node> graph (src/rrt.rs:161-214): parent[i] < 0 for the root.
 * Replaces the RTree inserts at src/rrt.rs:345-346 (upload) and :586-589 (append). */
int pp_tree_upload(pp_ctx *ctx, size_t n, const double *x, const double *y, const double *yaw,
                   const int32_t *parent);
int pp_tree_upload_dev(pp_ctx *ctx, size_t n, const double *x, const double *y, const double *yaw,
                       const int32_t *parent);
int pp_tree_append(pp_ctx *ctx, size_t k, const double *x, const double *y, const double *yaw,
                   const int32_t *parent);
size_t pp_tree_size(pp_ctx *ctx);

/* Space{bounds, obstacles} after Space::new (src/rrt.rs:113-121): one bounds exterior ring and
 * n_rings obstacle exterior rings in CSR form (ring r = points ring_off[r] .. ring_off[r+1]).
 * Rings are closed like geo-types' Polygon::new does (first point appended when last != first). */
int pp_obstacles_upload(pp_ctx *ctx, const double *bounds_x, const double *bounds_y, size_t n_bounds,
                        const double *ring_x, const double *ring_y, const uint32_t *ring_off, size_t n_rings);

/* ------------------------------------------------------------------ RRT: nearest neighbour */
/* RRT::get_nearest_node (src/rrt.rs:378-391): idx[j] = argmin_i (dx*dx + dy*dy), non-fused f64,
 * lowest index on ties; 0xFFFFFFFF for an empty tree.  d2 may be NULL. */
int pp_nn(pp_ctx *ctx, size_t m, const double *qx, const double *qy, uint32_t *idx, double *d2, int flags);
int pp_nn_dev(pp_ctx *ctx, size_t m, const double *qx, const double *qy, uint32_t *idx, double *d2, int flags);

/* ------------------------------------------------------------------ RRT: verify */
/* Space::verify (src/rrt.rs:124-137) on m straight 2-point lines a->b: ok[i] = 1 iff both points are
 * strictly inside the bounds ring and no obstacle ring intersects / contains the line. */
int pp_collide_segments(pp_ctx *ctx, size_t m, const double *ax, const double *ay, const double *bx,
                        const double *by, uint8_t *ok, int flags);
int pp_collide_segments_dev(pp_ctx *ctx, size_t m, const double *ax, const double *ay, const double *bx,
                            const double *by, uint8_t *ok, int flags);
/* Space::verify on arbitrary polylines in CSR form (line i = points line_off[i] .. line_off[i+1]) */
int pp_verify_polylines(pp_ctx *ctx, size_t n_lines, const double *px, const double *py,
                        const uint32_t *line_off, uint8_t *ok, int flags);
/* RRT::verify_node per tree edge (src/rrt.rs:414-426 with line_to_origin, :291-321, decomposed per
 * edge): Dubins curve child->parent sampled at `step`, polyline = samples ++ [parent point]
 * (fallback [(sx,sy), parent] when no word is feasible, src/rrt.rs:313), never materialised. */
int pp_collide_dubins(pp_ctx *ctx, size_t m, const double *sx, const double *sy, const double *syaw,
                      const double *ex, const double *ey, const double *eyaw, double radius, double step,
                      uint8_t *ok, int flags);
int pp_collide_dubins_dev(pp_ctx *ctx, size_t m, const double *sx, const double *sy, const double *syaw,
                          const double *ex, const double *ey, const double *eyaw, double radius, double step,
                          uint8_t *ok, int flags);

/* one batched extend step (RRT::get_random_node + verify of the new straight edge,
 * src/rrt.rs:406-412, 169-175, 267-271): for each sample point q_j: idx = nearest node, yaw = heading
 * from q_j toward that node (compute_yaw), ok = Space::verify of the 2-point line q_j -> node. */
int pp_rrt_extend(pp_ctx *ctx, size_t m, const double *qx, const double *qy, uint32_t *idx, double *yaw,
                  uint8_t *ok, int nn_flags, int collide_flags);
int pp_rrt_extend_dev(pp_ctx *ctx, size_t m, const double *qx, const double *qy, uint32_t *idx, double *yaw,
                      uint8_t *ok, int nn_flags, int collide_flags);

/* the same step with the reference's real edge geometry (src/rrt.rs:406-426, 291-321): idx = nearest node,
 * yaw = heading from q_j toward it, ok = Space::verify of the Dubins curve (q_j, yaw) -> (node, node yaw)
 * sampled at `step` (polyline = samples ++ [node point]).  For a tree whose nodes were inserted through
 * verify_node (every chain already verified) this is verify_node of the new node. */
int pp_rrt_extend_dubins(pp_ctx *ctx, size_t m, const double *qx, const double *qy, double radius, double step,
                         uint32_t *idx, double *yaw, uint8_t *ok, int nn_flags, int collide_flags);
int pp_rrt_extend_dubins_dev(pp_ctx *ctx, size_t m, const double *qx, const double *qy, double radius, double step,
                             uint32_t *idx, double *yaw, uint8_t *ok, int nn_flags, int collide_flags);

/* ------------------------------------------------------------------ multi-GPU
 * Every pose pair, NN query and edge is independent (src/rrt.rs:607-609 treats plan_one iterations that way): each
 * device takes the contiguous slice [g*n/G, (g+1)*n/G) of a batch, tree and obstacles are REPLICATED, results are
 * gathered by per-device copies.  The only collective is the broadcast of the tree / obstacle buffers when they
 * change (ncclBroadcast over NVLink / NVSwitch) -- of just the appended tail for pp_*_tree_append, the insert site
 * src/rrt.rs:586-589.  NCCL is loaded at run time (dlopen of libnccl.so.2); without it these calls return PP_ERR_COMM.
 *
 * Two ways to drive N GPUs:
 *  (1) ONE host process (a Rust / C++ planner): pp_group owns one pp_ctx + one worker thread per device and one
 *      communicator over them (ncclCommInitAll).
 *  (2) ONE process per GPU (torchrun, MPI): each rank creates its own pp_ctx and joins a communicator with
 *      pp_ctx_comm_init (rank 0 obtains the id with pp_comm_unique_id and the launcher's own channel hands the 128
 *      bytes to the other ranks); pp_tree_*_bcast / pp_obstacles_upload_bcast are then collective calls. */
#define PP_COMM_ID_BYTES 128
/* slice of device g out of G for a batch of n: [*lo, *hi) = [g*n/G, (g+1)*n/G); pure host arithmetic */
void pp_slice_bounds(size_t n, int parts, int part, size_t *lo, size_t *hi);
int pp_comm_unique_id(void *id /* PP_COMM_ID_BYTES */);
int pp_ctx_comm_init(pp_ctx *ctx, const void *id, int n_ranks, int rank);
int pp_ctx_comm_rank(pp_ctx *ctx); /* -1 without a communicator */
int pp_ctx_comm_size(pp_ctx *ctx); /* 1 without a communicator */
/* collective over the ctx's communicator: the host arrays are read on `root` only (other ranks may pass NULL); every
 * rank ends with the same tree.  n / k must be the same on all ranks.  RTree inserts at src/rrt.rs:345-346, 586-589 */
int pp_tree_upload_bcast(pp_ctx *ctx, int root, size_t n, const double *x, const double *y, const double *yaw,
                         const int32_t *parent);
int pp_tree_append_bcast(pp_ctx *ctx, int root, size_t k, const double *x, const double *y, const double *yaw,
                         const int32_t *parent);
/* collective: `root` preprocesses and uploads the world (Space after Space::new, src/rrt.rs:113-121), its device
 * buffers are broadcast; other ranks may pass NULL / 0 for every geometry argument */
int pp_obstacles_upload_bcast(pp_ctx *ctx, int root, const double *bounds_x, const double *bounds_y, size_t n_bounds,
                              const double *ring_x, const double *ring_y, const uint32_t *ring_off, size_t n_rings);

int pp_group_create(const int *devices, int n_dev, pp_group **out);
void pp_group_destroy(pp_group *g);
int pp_group_size(pp_group *g);
pp_ctx *pp_group_ctx(pp_group *g, int i); /* device i's context (owned by the group) */
const char *pp_group_last_error(pp_group *g);
/* replicate: H2D once on device 0, ncclBroadcast to the others */
int pp_group_tree_upload(pp_group *g, size_t n, const double *x, const double *y, const double *yaw,
                         const int32_t *parent);
int pp_group_tree_append(pp_group *g, size_t k, const double *x, const double *y, const double *yaw,
                         const int32_t *parent);
int pp_group_obstacles_upload(pp_group *g, const double *bounds_x, const double *bounds_y, size_t n_bounds,
                              const double *ring_x, const double *ring_y, const uint32_t *ring_off, size_t n_rings);
/* sliced batch calls on HOST pointers: device i works on its contiguous slice concurrently with the others and
 * copies its results straight into the caller's arrays.  Same arguments and results as the pp_* calls above --
 * outputs are byte-identical for every device count. */
int pp_group_dubins_eval(pp_group *g, size_t n, const double *sx, const double *sy, const double *syaw,
                         const double *ex, const double *ey, const double *eyaw, const double *radius_arr,
                         double radius, double *cost, uint8_t *word, double *tpq);
int pp_group_nn(pp_group *g, size_t m, const double *qx, const double *qy, uint32_t *idx, double *d2, int flags);
int pp_group_collide_segments(pp_group *g, size_t m, const double *ax, const double *ay, const double *bx,
                              const double *by, uint8_t *ok, int flags);
int pp_group_collide_dubins(pp_group *g, size_t m, const double *sx, const double *sy, const double *syaw,
                            const double *ex, const double *ey, const double *eyaw, double radius, double step,
                            uint8_t *ok, int flags);
int pp_group_rrt_extend(pp_group *g, size_t m, const double *qx, const double *qy, uint32_t *idx, double *yaw,
                        uint8_t *ok, int nn_flags, int collide_flags);
int pp_group_rrt_extend_dubins(pp_group *g, size_t m, const double *qx, const double *qy, double radius, double step,
                               uint32_t *idx, double *yaw, uint8_t *ok, int nn_flags, int collide_flags);

/* ------------------------------------------------------------------ measurement helpers */
/* host<->device copy ceiling of this box for the end-to-end figures: h2d_bytes up and d2h_bytes down between pinned
 * (pinned != 0) or pageable host memory and the device, both directions at once on two streams, no kernel. */
int pp_measure_copy(pp_ctx *ctx, size_t h2d_bytes, size_t d2h_bytes, int pinned, double *ms);
/* FP64 pipe peak micro-benchmark: runs `iters` dependent DFMA chains of
 * length `chain` on every SM and returns DFMA thread-instructions per second. */
int pp_measure_fp64_peak(pp_ctx *ctx, int iters, double *dfma_per_s, double *ms);
/* per-kernel device time: when enabled, every launch group is bracketed by CUDA events on the stream it
 * is launched on; pp_timing_get sums them (kernel = "dubins_eval", "nn_scan", "nn_scan_f64", "nn_grid",
 * "nn_wide", "collide_segments", "collide_dubins", "dubins_plan", "dubins_fill", ...). */
int pp_timing_enable(pp_ctx *ctx, int on);
int pp_timing_reset(pp_ctx *ctx);
int pp_timing_get(pp_ctx *ctx, const char *kernel, double *total_ms, uint64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* PATHPLANNING_B200_H */

// collide.cu -- Space::verify (src/rrt.rs:124-137) for batches of edges (kernel 3).
// Compiled with -fmad=false; predicates in geo_predicates.cuh follow geo 0.12.2's operation order, so
// flags are bit-exact with the oracle whenever the line coordinates are.
//
// Data layout in HBM (built once per obstacle set by pp_obstacles_upload, api.cu):
//   bx/by            bounds exterior ring (closed), f64
//   bcls             bgx*bgy byte grid over the bounds' padded AABB: 0 = outside, 1 = inside, 2 = a ring
//                    segment passes within the pad of the cell -> exact get_position needed
//   ox/oy + meta[]   obstacle rings (closed) as one SoA + per-ring {exact AABB, pad, first, count}
//   aabb32[]         per-ring padded AABB rounded outward to fp32 (minx,miny,maxx,maxy), padded to the
//                    tile size with empty boxes: streamed through shared memory by the scan kernel
//   cell_start/items uniform grid over the padded ring AABBs (CSR), for polylines of short segments
//
// Exactness of the broad phases (DESIGN.md "culls"):
//   * a point outside a ring's padded AABB is never `Inside` for geo's crossing-number code: above,
//     below or right of the box no segment passes the straddle/`px <= max x` tests; left of it every
//     straddling segment counts (xints' rounding error is far below the pad) and a closed ring is
//     straddled an even number of times.
//   * a line segment whose AABB is disjoint from the ring's padded AABB cannot pass the parameter test
//     (except for rounding noise on pairs parallel to within ~2^-45 rad, where the reference's own
//     answer is noise; PP_COLLIDE_NO_CULL gives the exhaustive loop).
#include "dubins_device.cuh"
#include "geo_predicates.cuh"
#include "nn_grid.cuh"
#include "pp_common.cuh"

pp_world_view pp_make_world_view(const pp_world_dev &w);  // api.cu
int pp_build_bucket_perm(pp_ctx *ctx, size_t m, const double *kx, const double *ky, uint32_t **perm_out,
                         cudaStream_t stream);  // nn.cu

#define PP_AABB_TILE 1024  // ring boxes per shared-memory tile (16 KB)

// ---- Polygon::contains(&Point) for the bounds ring ------------------------------------------------
__device__ __forceinline__ bool pp_bounds_contains(const pp_world_view &w, double x, double y) {
    const double fx = (x - w.bminx) * w.binvx, fy = (y - w.bminy) * w.binvy;
    // floor-convert (saturating) and range-check as integers: -0.3 -> -1, +-inf -> INT_MIN / INT_MAX all fail the
    // unsigned compare.  NaN converts to cell (0, 0), a corner of the padded box and therefore never class 1; its
    // class-2 exact test is false for NaN as well.
    const int ix = __double2int_rd(fx), iy = __double2int_rd(fy);
    if ((unsigned)ix >= (unsigned)w.bgx || (unsigned)iy >= (unsigned)w.bgy) return false;
    const uint8_t cls = __ldg(w.bcls + (size_t)iy * w.bgx + ix);
    if (cls != 2) return cls == 1;
    return pp_point_position(w.bx, w.by, w.nb, x, y) == 1;
}

__device__ __forceinline__ bool pp_point_in_ring(const pp_world_view &w, const pp_ring_meta &m, double x, double y) {
    return pp_point_position(w.ox + m.first, w.oy + m.first, m.count, x, y) == 1;
}
__device__ __forceinline__ bool pp_outside_padded(const pp_ring_meta &m, double x, double y) {
    return (x < m.minx - m.pad) || (x > m.maxx + m.pad) || (y < m.miny - m.pad) || (y > m.maxy + m.pad);
}

// circle filter of ring r for the segment a-b (geo_predicates.cuh: pp_circle_class): 0 skip the ring, 1 blocked, 2 exact
__device__ __forceinline__ int pp_circle_class_at(const pp_ring_circle *cp, double ax, double ay, double bx, double by) {
    const double2 lo = __ldg(reinterpret_cast<const double2 *>(cp));
    const double2 hi = __ldg(reinterpret_cast<const double2 *>(cp) + 1);
    pp_ring_circle c;
    c.cx = lo.x;
    c.cy = lo.y;
    c.rout2 = hi.x;
    c.rin2 = hi.y;
    return pp_circle_class(c, ax, ay, bx, by);
}
__device__ __forceinline__ int pp_ring_circle_class(const pp_world_view &w, uint32_t r, double ax, double ay, double bx,
                                                    double by) {
    return pp_circle_class_at(w.circ + r, ax, ay, bx, by);
}

// ---- warp-cooperative narrow phase.  The warp is split into four groups of eight lanes; a group works on one
// pending (line segment, ring) candidate and its lane `sub` takes ring segments sub, sub + 8, ...: the rings of
// this domain have 4-20 segments, so whole-warp passes would leave three quarters of the lanes idle.  All 32
// lanes call together (the ballots are warp-wide); a group without a candidate passes n = 0.  Per (ring segment,
// line segment) pair the arithmetic is that of geo_predicates.cuh, so the answers are the same bits; what changes
// is that a ring's segments are tested side by side instead of in a serial loop on the lane that found the ring.
__device__ __forceinline__ bool pp_ring_hits_segment_g8(const double *__restrict__ rx, const double *__restrict__ ry,
                                                        uint32_t n, double b0x, double b0y, double b1x, double b1y,
                                                        int sub, unsigned gmask) {
    const double b_dx = b1x - b0x, b_dy = b1y - b0y;
    bool res = false;
    for (uint32_t base = 0;; base += 8) {
        const bool work = base + 1 < n;  // uniform inside a group
        if (__ballot_sync(0xffffffffu, work) == 0u) break;
        const uint32_t i = base + (uint32_t)sub;
        bool hit = false;
        if (work && i + 1 < n) {
            const double a0x = rx[i], a0y = ry[i];
            const double a_dx = rx[i + 1] - a0x, a_dy = ry[i + 1] - a0y;
            const double u_b = b_dy * a_dx - b_dx * a_dy;
            if (u_b != 0.0) {
                const double ua_t = b_dx * (a0y - b0y) - b_dy * (a0x - b0x);
                const double ub_t = a_dx * (a0y - b0y) - a_dy * (a0x - b0x);
                hit = pp_quot_in01(ua_t, u_b) && pp_quot_in01(ub_t, u_b);
            }
        }
        if ((__ballot_sync(0xffffffffu, hit) & gmask) != 0u) {
            res = true;
            n = 0;  // this group is done; the others may go on
        }
    }
    return res;
}

// get_position(ring, p) == Inside per group: a vertex or segment that holds the point makes it OnBoundary
// (pp_ring_has_point), otherwise the parity of the crossing count decides (pp_point_position); the `xints` carried
// between iterations in geo's loop is only read for segments with y0 != y1, which also write it, so every
// segment's contribution is independent of the others.
__device__ __forceinline__ bool pp_point_inside_ring_g8(const double *__restrict__ rx, const double *__restrict__ ry,
                                                        uint32_t n, double px, double py, int sub, unsigned gmask) {
    if (n < 2) n = 0;  // a ring of one point is never `Inside`
    uint32_t crossings = 0;
    bool boundary = false;
    for (uint32_t base = 0;; base += 8) {
        const bool work = base < n;
        if (__ballot_sync(0xffffffffu, work) == 0u) break;
        const uint32_t i = base + (uint32_t)sub;
        bool on = false, cross = false;
        if (work && i < n) {
            const double x0 = rx[i], y0 = ry[i];
            on = (x0 == px && y0 == py);
            if (i + 1 < n) {
                const double x1 = rx[i + 1], y1 = ry[i + 1];
                const double dx = x1 - x0, dy = y1 - y0;
                if (dx == 0.0 && dy == 0.0) {
                    // the vertex test above already covers it
                } else if (dy == 0.0) {
                    on = on || (py == y0 && pp_quot_in01(px - x0, dx));
                } else if (dx == 0.0) {
                    on = on || (px == x0 && pp_quot_in01(py - y0, dy));
                } else {
                    const double nx = px - x0, ny = py - y0;
                    if (pp_quot_in01(nx, dx) && pp_quot_near01(ny, dy))  // see pp_ring_has_point
                        on = on || (fabs(nx / dx - ny / dy) <= PP_F64_EPSILON);
                }
                const double ymin = (y0 < y1) ? y0 : y1, ymax = (y0 > y1) ? y0 : y1;
                const double xmax = (x0 > x1) ? x0 : x1;
                if (py > ymin && py <= ymax && px <= xmax) {
                    // y0 != y1 here (py > ymin && py <= ymax)
                    const double xints = (py - y0) * (x1 - x0) / (y1 - y0) + x0;
                    cross = (x0 == x1 || px <= xints);
                }
            }
        }
        if ((__ballot_sync(0xffffffffu, on) & gmask) != 0u) {
            boundary = true;
            n = 0;
        }
        crossings += (uint32_t)__popc(__ballot_sync(0xffffffffu, cross) & gmask);
    }
    return !boundary && (crossings & 1u) != 0u;
}

// the group's candidate: the (grp)-th lowest set bit of `pend` (or -1), and `pend` without its four lowest bits
__device__ __forceinline__ int pp_take_candidates(unsigned &pend, int grp) {
    int mine = -1;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int s = pend ? __ffs(pend) - 1 : -1;
        pend &= pend - 1;  // 0 stays 0
        if (g == grp) mine = s;
    }
    return mine;
}

// ---- exhaustive per-lane loops of PP_COLLIDE_NO_CULL: every ring, no broad phase (geo's own loop order) -------
__device__ __forceinline__ bool pp_vertex_in_any_obstacle(const pp_world_view &w, double x, double y) {
    for (uint32_t r = 0; r < w.n_rings; ++r)
        if (pp_point_in_ring(w, w.meta[r], x, y)) return true;
    return false;
}
__device__ __forceinline__ bool pp_segment_hits_any_obstacle(const pp_world_view &w, double x0, double y0, double x1,
                                                             double y1) {
    for (uint32_t r = 0; r < w.n_rings; ++r) {
        const pp_ring_meta m = w.meta[r];
        if (pp_ring_hits_segment(w.ox + m.first, w.oy + m.first, m.count, x0, y0, x1, y1)) return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// kernel 3a: straight 2-point edges, one thread per edge; the per-ring fp32 boxes of ALL obstacles
// stream through shared memory in 16 KB tiles (TMA bulk copy, 2-stage mbarrier ring); a ring whose
// box overlaps the edge's outward-rounded fp32 box takes the exact f64 test.  A warp whose edges are
// all decided (ballot) skips the box loop of the remaining tiles.
// MODE 0: tiled scan, one edge per thread in the caller's order   1: exhaustive, no cull
// ------------------------------------------------------------------------------------------------
#define PP_SEG_THREADS 128

template <int MODE>
__global__ void __launch_bounds__(PP_SEG_THREADS)
    pp_collide_segments_kernel(pp_world_view w, size_t m, const double *__restrict__ ax,
                               const double *__restrict__ ay, const double *__restrict__ bx,
                               const double *__restrict__ by, const uint32_t *__restrict__ gather_idx,
                               const double *__restrict__ node_x, const double *__restrict__ node_y,
                               double *__restrict__ yaw_out, uint8_t *__restrict__ ok) {
    __shared__ __align__(128) float4 tiles[2][PP_AABB_TILE];
    __shared__ uint64_t full[2];
    const size_t i = (size_t)blockIdx.x * PP_SEG_THREADS + threadIdx.x;
    const bool live = i < m;
    double x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (live) {
        x0 = ax[i];
        y0 = ay[i];
        if (gather_idx) {  // extend step: b = tree node nearest to the sample a (src/rrt.rs:406-411)
            // pp_nn answers 0xFFFFFFFF for a query without a nearest node (NaN / Inf coordinates, d2 overflow):
            // the reference's get_random_node returns None there, here the step reports ok = 0 and yaw = NaN
            const uint32_t g = gather_idx[i];
            const bool none = g == 0xFFFFFFFFu;
            x1 = none ? CUDART_NAN : node_x[g];
            y1 = none ? CUDART_NAN : node_y[g];
            if (yaw_out) yaw_out[i] = none ? CUDART_NAN : atan2(y1 - y0, x1 - x0);  // compute_yaw, src/rrt.rs:267-271
        } else {
            x1 = bx[i];
            y1 = by[i];
        }
    }
    // bounds.contains(line): both points strictly inside (src/rrt.rs:125)
    bool good = live && pp_bounds_contains(w, x0, y0) && pp_bounds_contains(w, x1, y1);
    bool hit = false;

    if (MODE == 1) {
        if (good)
            hit = pp_segment_hits_any_obstacle(w, x0, y0, x1, y1) || pp_vertex_in_any_obstacle(w, x0, y0) ||
                  pp_vertex_in_any_obstacle(w, x1, y1);
    } else {
        const float eminx = __double2float_rd(fmin(x0, x1)), emaxx = __double2float_ru(fmax(x0, x1));
        const float eminy = __double2float_rd(fmin(y0, y1)), emaxy = __double2float_ru(fmax(y0, y1));
        const int tid = threadIdx.x;
        if (tid == 0) {
            pp_mbar_init(&full[0], 1);
            pp_mbar_init(&full[1], 1);
            pp_fence_mbar_init();
        }
        __syncthreads();
        if (tid == 0) {
            for (uint32_t t = 0; t < 2 && t < w.n_aabb_tiles; ++t) {
                pp_mbar_expect_tx(&full[t], PP_AABB_TILE * 16);
                pp_bulk_g2s(tiles[t], w.aabb32 + (size_t)t * PP_AABB_TILE, PP_AABB_TILE * 16, &full[t]);
            }
        }
        for (uint32_t t = 0; t < w.n_aabb_tiles; ++t) {
            const int s = t & 1;
            pp_mbar_wait(&full[s], (t >> 1) & 1u);
            const bool undecided = good && !hit;
            if (__ballot_sync(0xffffffffu, undecided) != 0u) {
                if (undecided) {
                    const float4 *T = tiles[s];
#pragma unroll 4
                    for (int r = 0; r < PP_AABB_TILE; ++r) {
                        const float4 bb = T[r];
                        const bool overlap = !(emaxx < bb.x || eminx > bb.z || emaxy < bb.y || eminy > bb.w);
                        if (overlap && t * PP_AABB_TILE + r < w.n_rings) {
                            const pp_ring_meta mt = w.meta[t * PP_AABB_TILE + r];
                            const double *rx = w.ox + mt.first, *ry = w.oy + mt.first;
                            // the f64 padded-box rule decides whether the pair is tested at all (as in the grid and
                            // polyline kernels and the oracle's culled loop): the fp32 boxes are rounded outward and
                            // would otherwise let a few more near-parallel noise pairs through than those paths do
                            const bool sgx = x1 < x0, sgy = y1 < y0;
                            const bool seg_in = !((sgx ? x0 : x1) < mt.minx - mt.pad || (sgx ? x1 : x0) > mt.maxx + mt.pad ||
                                                  (sgy ? y0 : y1) < mt.miny - mt.pad || (sgy ? y1 : y0) > mt.maxy + mt.pad);
                            const int cls = pp_ring_circle_class(w, t * PP_AABB_TILE + r, x0, y0, x1, y1);  // circle filter
                            if (cls == 0) continue;
                            if (cls == 1) {
                                hit = true;
                                break;
                            }
                            if ((seg_in && pp_ring_hits_segment(rx, ry, mt.count, x0, y0, x1, y1)) ||
                                (!pp_outside_padded(mt, x0, y0) && pp_point_position(rx, ry, mt.count, x0, y0) == 1) ||
                                (!pp_outside_padded(mt, x1, y1) && pp_point_position(rx, ry, mt.count, x1, y1) == 1)) {
                                hit = true;
                                break;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            if (tid == 0 && t + 2 < w.n_aabb_tiles) {
                pp_mbar_expect_tx(&full[s], PP_AABB_TILE * 16);
                pp_bulk_g2s(tiles[s], w.aabb32 + (size_t)(t + 2) * PP_AABB_TILE, PP_AABB_TILE * 16, &full[s]);
            }
        }
    }
    if (live) ok[i] = (good && !hit) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// kernel 3a-grid (default for straight edges): one thread per edge, broad phase through the uniform obstacle
// grid, narrow phase per warp.  A lane walks the cells under its edge's box (cell range from the end points'
// cells, integer arithmetic; a row-wise difference of the CSR offsets tells at once whether any ring is registered
// there) and stops at the first ring whose outward-rounded fp32 box meets the edge's; the warp then takes the
// pending (lane, ring) candidates one by one, broadcasts the edge by shuffle and runs the exact predicates with one
// ring segment per lane.  The serial per-lane form ran the predicates at 2.6 active lanes per instruction.
// An edge whose box covers more cells than there are rings walks the ring list instead (cost bounded by O(rings)).
// ------------------------------------------------------------------------------------------------
#ifndef PP_SEGGRID_MIN_BLOCKS
#define PP_SEGGRID_MIN_BLOCKS 6  // 80 registers: 0.186 ms per 2^20 C4 edges (0.211 at 90 registers, 0.189 at 64)
#endif
#ifndef PP_SEG_FOLD_COUNT
#define PP_SEG_FOLD_COUNT 1
#endif
// Space::verify of one straight edge per lane through the obstacle grid; ALL 32 lanes of the warp call together
// (ballots and shuffles inside), lanes without an edge pass live = false
__device__ __forceinline__ bool pp_verify_segment_grid(const pp_world_view &w, bool live, double x0, double y0, double x1,
                                                       double y1, int lane) {
    // bounds.contains(line): both points strictly inside (src/rrt.rs:125)
    const bool good = live && pp_bounds_contains(w, x0, y0) && pp_bounds_contains(w, x1, y1);
    bool hit = false;
    // cell range of the edge's box: floor is monotonic, so it is spanned by the end points' cells
    const int ia = __double2int_rd((x0 - w.gminx) * w.ginv), ja = __double2int_rd((y0 - w.gminy) * w.ginv);
    const int ib = __double2int_rd((x1 - w.gminx) * w.ginv), jb = __double2int_rd((y1 - w.gminy) * w.ginv);
    int cx0 = min(ia, ib), cx1 = max(ia, ib), cy = min(ja, jb), cy1 = max(ja, jb);
    bool more = good && w.n_rings != 0u && !(cx1 < 0 || cy1 < 0 || cx0 >= w.gx || cy >= w.gy);
    cx0 = max(cx0, 0);
    cx1 = min(cx1, w.gx - 1);
    cy = max(cy, 0);
    cy1 = min(cy1, w.gy - 1);
    uint32_t kcur = 0, kend = 0;
    bool linear = false;
    if (more) {
        if ((unsigned long long)(cx1 - cx0 + 1) * (unsigned long long)(cy1 - cy + 1) > (unsigned long long)w.n_rings + 64ull) {
            linear = true;
            kend = w.n_rings;
            cy = cy1 + 1;
        } else {
#if PP_SEG_FOLD_COUNT
            // first row's run fetched here and handed to the walk: an empty single-row range ends the lane's work
            const uint32_t *row = w.cell_start + (size_t)cy * w.gx;
            kcur = __ldg(row + cx0);
            kend = __ldg(row + cx1 + 1);
            ++cy;
            if (kcur == kend && cy > cy1) more = false;
#else
            uint32_t cnt = 0;
            for (int r = cy; r <= cy1; ++r) {
                const uint32_t *row = w.cell_start + (size_t)r * w.gx;
                cnt += __ldg(row + cx1 + 1) - __ldg(row + cx0);
            }
            more = cnt != 0u;
#endif
        }
    }
    if (__ballot_sync(0xffffffffu, more) != 0u) {
        const bool swx = x1 < x0, swy = y1 < y0;
        const float q32x0 = __double2float_rd(swx ? x1 : x0), q32x1 = __double2float_ru(swx ? x0 : x1);
        const float q32y0 = __double2float_rd(swy ? y1 : y0), q32y1 = __double2float_ru(swy ? y0 : y1);
        for (;;) {
            uint32_t ring = 0xFFFFFFFFu;
            while (more) {
                if (kcur < kend) {
                    // box k of the cell-ordered copy: no dependent load through the ring id
                    const float4 bb = __ldg((linear ? w.aabb32 : w.cell_box) + kcur);
                    const uint32_t kk = kcur++;
                    if (!(q32x1 < bb.x || q32x0 > bb.z || q32y1 < bb.y || q32y0 > bb.w)) {
                        const uint32_t r = linear ? kk : __ldg(w.cell_items + kk);
                        // circle filter: most box candidates are decided here, per lane, without the warp-wide
                        // exact predicates (edge outside the ring's outer circle, or both ends inside its inner one).
                        // (A cell-ordered copy of the circles, like the boxes', was measured and dropped: the polyline
                        // kernel ran 40 % slower -- neighbouring cells share rings, and the per-ring array hits in L1.)
                        const int cls = pp_ring_circle_class(w, r, x0, y0, x1, y1);
                        if (cls == 1) {
                            hit = true;
                            more = false;
                            break;
                        }
                        if (cls == 2) {
                            ring = r;
                            break;
                        }
                    }
                } else {  // next row of cells: one contiguous run of ids
                    if (cy > cy1) {
                        more = false;
                        break;
                    }
                    const uint32_t *row = w.cell_start + (size_t)cy * w.gx;
                    kcur = __ldg(row + cx0);
                    kend = __ldg(row + cx1 + 1);
                    ++cy;
                }
            }
            uint32_t pend = __ballot_sync(0xffffffffu, ring != 0xFFFFFFFFu);
            if (pend == 0u) break;
            do {
                const int grp = lane >> 3, sub = lane & 7;
                const unsigned gmask = 0xFFu << (grp * 8);
                const int src = pp_take_candidates(pend, grp);
                const bool act = src >= 0;
                const int sl = act ? src : 0;
                const uint32_t rr = __shfl_sync(0xffffffffu, ring, sl);
                const double ex0 = __shfl_sync(0xffffffffu, x0, sl), ey0 = __shfl_sync(0xffffffffu, y0, sl);
                const double ex1 = __shfl_sync(0xffffffffu, x1, sl), ey1 = __shfl_sync(0xffffffffu, y1, sl);
                pp_ring_meta mt;
                mt.minx = mt.miny = mt.maxx = mt.maxy = mt.pad = 0.0;
                mt.first = mt.count = 0u;
                if (act) mt = w.meta[rr];
                const double *rx = w.ox + mt.first, *ry = w.oy + mt.first;
                // the f64 padded-box rule of the per-lane path decides whether the pair is tested at all
                const bool sx = ex1 < ex0, sy = ey1 < ey0;
                const bool do_seg = act && !((sx ? ex0 : ex1) < mt.minx - mt.pad || (sx ? ex1 : ex0) > mt.maxx + mt.pad ||
                                             (sy ? ey0 : ey1) < mt.miny - mt.pad || (sy ? ey1 : ey0) > mt.maxy + mt.pad);
                bool h = pp_ring_hits_segment_g8(rx, ry, do_seg ? mt.count : 0u, ex0, ey0, ex1, ey1, sub, gmask);
                const bool do_v0 = act && !h && !pp_outside_padded(mt, ex0, ey0);
                h = pp_point_inside_ring_g8(rx, ry, do_v0 ? mt.count : 0u, ex0, ey0, sub, gmask) || h;
                const bool do_v1 = act && !h && !pp_outside_padded(mt, ex1, ey1);
                h = pp_point_inside_ring_g8(rx, ry, do_v1 ? mt.count : 0u, ex1, ey1, sub, gmask) || h;
                // the verdict belongs to the lane the candidate came from
                const unsigned hb = __ballot_sync(0xffffffffu, h && sub == 0);
                bool mine = false;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int sg = __shfl_sync(0xffffffffu, src, g * 8);
                    if (((hb >> (g * 8)) & 1u) && sg == lane) mine = true;
                }
                if (mine) {
                    hit = true;
                    more = false;  // decided: stop walking
                }
            } while (pend != 0u);
        }
    }
    return good && !hit;
}

__global__ void __launch_bounds__(PP_SEG_THREADS, PP_SEGGRID_MIN_BLOCKS)
    pp_collide_segments_grid_kernel(pp_world_view w, size_t m, const double *__restrict__ ax,
                                    const double *__restrict__ ay, const double *__restrict__ bx,
                                    const double *__restrict__ by, const uint32_t *__restrict__ gather_idx,
                                    const double *__restrict__ node_x, const double *__restrict__ node_y,
                                    double *__restrict__ yaw_out, uint8_t *__restrict__ ok) {
    const size_t i = (size_t)blockIdx.x * PP_SEG_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool live = i < m;
    double x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (live) {
        x0 = ax[i];
        y0 = ay[i];
        if (gather_idx) {  // extend step: b = tree node nearest to the sample a (src/rrt.rs:406-411)
            // pp_nn answers 0xFFFFFFFF for a query without a nearest node (NaN / Inf coordinates, d2 overflow):
            // the reference's get_random_node returns None there, here the step reports ok = 0 and yaw = NaN
            const uint32_t g = gather_idx[i];
            const bool none = g == 0xFFFFFFFFu;
            x1 = none ? CUDART_NAN : node_x[g];
            y1 = none ? CUDART_NAN : node_y[g];
            if (yaw_out) yaw_out[i] = none ? CUDART_NAN : atan2(y1 - y0, x1 - x0);  // compute_yaw, src/rrt.rs:267-271
        } else {
            x1 = bx[i];
            y1 = by[i];
        }
    }
    const bool free_edge = pp_verify_segment_grid(w, live, x0, y0, x1, y1, lane);
    if (live) ok[i] = free_edge ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// The fused, cell-coherent extend step (src/rrt.rs:406-426 with straight edges): ONE launch does
//   nearest neighbour (exact grid search) -> Node::new's yaw (compute_yaw) -> Space::verify of the edge sample -> node.
// The queries are first binned by the node-grid block they fall into (pp_extend_bin / scan / scatter below: a
// counting sort whose rank comes out of the histogram's own atomicAdd), so the 32 queries of a warp lie in ONE
// 4 x 4 block of node cells: their 3 x 3 neighbourhoods overlap almost entirely, the cell runs they read are the
// same cache lines, their edges fall into the same obstacle cell, and the ring walk / narrow phase work on the same
// rings.  No idx / yaw round trip through memory between the two halves, and the edge's far end comes out of the NN
// scan's own registers instead of a second gather.  Per query the arithmetic is that of pp_nn_grid_kernel and
// pp_collide_segments_grid_kernel, so idx, yaw and ok are the same bits; only the order of the work changes.
// ------------------------------------------------------------------------------------------------
#define PP_EXT_BIN_SHIFT_MIN 2  // 4 x 4 node cells per bin (~32 queries per bin when queries ~ nodes)
#define PP_EXT_MAX_BINS 65536u

struct pp_extend_bins {
    int shift, bx, by;  // bins of (1 << shift)^2 node cells, bx x by of them
};

__device__ __forceinline__ uint32_t pp_extend_bin_of(const pp_nn_grid_view &g, const pp_extend_bins &b, double x, double y) {
    const double fx = floor((x - g.gminx) * g.ginv), fy = floor((y - g.gminy) * g.ginv);
    const int cx = (fx >= (double)g.gx) ? g.gx - 1 : ((fx > 0.0) ? (int)fx : 0);  // NaN -> 0, as the search does
    const int cy = (fy >= (double)g.gy) ? g.gy - 1 : ((fy > 0.0) ? (int)fy : 0);
    return (uint32_t)(cy >> b.shift) * (uint32_t)b.bx + (uint32_t)(cx >> b.shift);
}

// histogram of the bins; the value the atomicAdd returns is the query's rank inside its bin
__global__ void __launch_bounds__(256)
    pp_extend_bin_kernel(pp_nn_grid_view g, pp_extend_bins b, const double *__restrict__ qx, const double *__restrict__ qy,
                         uint32_t m, uint32_t *__restrict__ hist, uint32_t *__restrict__ bin_of, uint32_t *__restrict__ rank) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= m) return;
    const uint32_t bin = pp_extend_bin_of(g, b, qx[j], qy[j]);
    bin_of[j] = bin;
    rank[j] = atomicAdd(&hist[bin], 1u);
}

// exclusive scan of up to 65 536 bin counts in ONE block, in place: a thread owns 64 consecutive counts, fetched with
// sixteen independent 16-byte loads into registers (the first version walked them one dependent load at a time: ~25 us)
#define PP_EXT_SCAN_PER 64
__global__ void __launch_bounds__(1024) pp_extend_scan_kernel(uint32_t *__restrict__ hist) {
    __shared__ uint32_t warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *mine = reinterpret_cast<uint4 *>(hist + (size_t)threadIdx.x * PP_EXT_SCAN_PER);
    // two halves of 32 counts: eight independent 16-byte loads in flight at a time (all sixteen would not fit the
    // 64 registers a 1 024-thread block leaves per thread)
    uint32_t half[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = mine[8 * h + k];
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[k].x + v[k].y + v[k].z + v[k].w;
        half[h] = s;
    }
    const uint32_t sum = half[0] + half[1];
    uint32_t inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t wv = warp_sums[lane];
        uint32_t winc = wv;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_sums[lane] = winc - wv;
    }
    __syncthreads();
    uint32_t run = warp_sums[warp] + inc - sum;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = mine[8 * h + k];  // second read: this thread's own lines, L1-resident
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint4 o;
            o.x = run;
            run += v[k].x;
            o.y = run;
            run += v[k].y;
            o.z = run;
            run += v[k].z;
            o.w = run;
            run += v[k].w;
            mine[8 * h + k] = o;
        }
    }
}

__global__ void __launch_bounds__(256)
    pp_extend_scatter_kernel(uint32_t m, const uint32_t *__restrict__ start, const uint32_t *__restrict__ bin_of,
                             const uint32_t *__restrict__ rank, uint32_t *__restrict__ perm) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j < m) perm[__ldg(start + bin_of[j]) + rank[j]] = j;
}

#ifndef PP_EXTEND_MIN_BLOCKS
#define PP_EXTEND_MIN_BLOCKS 6
#endif
__global__ void __launch_bounds__(PP_SEG_THREADS, PP_EXTEND_MIN_BLOCKS)
    pp_rrt_extend_fused_kernel(pp_nn_grid_view g, pp_world_view w, uint32_t m, const uint32_t *__restrict__ perm,
                               const double *__restrict__ qx, const double *__restrict__ qy, uint32_t *__restrict__ idx_out,
                               double *__restrict__ yaw_out, uint8_t *__restrict__ ok) {
    const uint32_t t = blockIdx.x * PP_SEG_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool live = t < m;
    uint32_t j = 0, bi = 0xFFFFFFFFu;
    double x0 = 0, y0 = 0, x1 = 0, y1 = 0, best;
    if (live) {
        j = perm ? __ldg(perm + t) : t;
        x0 = qx[j];
        y0 = qy[j];
    }
    pp_nn_grid_search(g, live, x0, y0, best, bi, x1, y1);  // RRT::get_nearest_node, src/rrt.rs:378-391 (whole warp)
    if (live) {
        idx_out[j] = bi;
        const bool none = bi == 0xFFFFFFFFu;  // no nearest node: get_random_node's None -> ok = 0, yaw = NaN
        if (none) x1 = y1 = CUDART_NAN;
        if (yaw_out) yaw_out[j] = none ? CUDART_NAN : atan2(y1 - y0, x1 - x0);  // compute_yaw, src/rrt.rs:267-271
    }
    __syncwarp();
    const bool free_edge = pp_verify_segment_grid(w, live, x0, y0, x1, y1, lane);
    if (live) ok[j] = free_edge ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// kernel 3a' (PP_COLLIDE_SCAN): the same tiled scan over ALL ring boxes, but the edges are
// first binned by their start point (the counting sort of the NN scan, nn.cu), each thread owns 4 edges
// and the warp keeps the union box of its 128 undecided edges.  Per step the 32 lanes test 32 different
// ring boxes (one LDS.128 each) against the warp box; a ballot yields the rare candidate rings, whose box
// is then broadcast by shuffle and tested per edge (fp32), and only overlapping (edge, ring) pairs take
// the exact f64 predicates.  Exactness is that of the per-edge scan (the warp box only pre-filters).
// ------------------------------------------------------------------------------------------------
#define PP_SEGB_EPT 4

__global__ void __launch_bounds__(PP_SEG_THREADS)
    pp_collide_segments_bucketed_kernel(pp_world_view w, size_t m, const double *__restrict__ ax,
                                        const double *__restrict__ ay, const double *__restrict__ bx,
                                        const double *__restrict__ by, const uint32_t *__restrict__ gather_idx,
                                        const double *__restrict__ node_x, const double *__restrict__ node_y,
                                        double *__restrict__ yaw_out, const uint32_t *__restrict__ perm,
                                        uint8_t *__restrict__ ok) {
    __shared__ __align__(128) float4 tiles[2][PP_AABB_TILE];
    __shared__ uint64_t full[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t warp_base = ((size_t)blockIdx.x * (PP_SEG_THREADS / 32) + warp) * (32 * PP_SEGB_EPT);
    uint32_t ei[PP_SEGB_EPT];
    double x0[PP_SEGB_EPT], y0[PP_SEGB_EPT], x1[PP_SEGB_EPT], y1[PP_SEGB_EPT];
    float eminx[PP_SEGB_EPT], emaxx[PP_SEGB_EPT], eminy[PP_SEGB_EPT], emaxy[PP_SEGB_EPT];
    bool good[PP_SEGB_EPT], hit[PP_SEGB_EPT];
#pragma unroll
    for (int e = 0; e < PP_SEGB_EPT; ++e) {
        const size_t pos = warp_base + (size_t)e * 32 + lane;
        const bool live = pos < m;
        ei[e] = live ? perm[pos] : 0xFFFFFFFFu;
        x0[e] = y0[e] = x1[e] = y1[e] = 0.0;
        if (live) {
            const uint32_t i = ei[e];
            x0[e] = ax[i];
            y0[e] = ay[i];
            if (gather_idx) {  // extend step: b = tree node nearest to the sample a (src/rrt.rs:406-411)
                const uint32_t g = gather_idx[i];
                const bool none = g == 0xFFFFFFFFu;  // no nearest node (see pp_collide_segments_kernel)
                x1[e] = none ? CUDART_NAN : node_x[g];
                y1[e] = none ? CUDART_NAN : node_y[g];
                if (yaw_out) yaw_out[i] = none ? CUDART_NAN : atan2(y1[e] - y0[e], x1[e] - x0[e]);  // compute_yaw, src/rrt.rs:267-271
            } else {
                x1[e] = bx[i];
                y1[e] = by[i];
            }
        }
        // bounds.contains(line): both points strictly inside (src/rrt.rs:125)
        good[e] = live && pp_bounds_contains(w, x0[e], y0[e]) && pp_bounds_contains(w, x1[e], y1[e]);
        hit[e] = false;
        eminx[e] = __double2float_rd(fmin(x0[e], x1[e]));
        emaxx[e] = __double2float_ru(fmax(x0[e], x1[e]));
        eminy[e] = __double2float_rd(fmin(y0[e], y1[e]));
        emaxy[e] = __double2float_ru(fmax(y0[e], y1[e]));
    }
    if (tid == 0) {
        pp_mbar_init(&full[0], 1);
        pp_mbar_init(&full[1], 1);
        pp_fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (uint32_t t = 0; t < 2 && t < w.n_aabb_tiles; ++t) {
            pp_mbar_expect_tx(&full[t], PP_AABB_TILE * 16);
            pp_bulk_g2s(tiles[t], w.aabb32 + (size_t)t * PP_AABB_TILE, PP_AABB_TILE * 16, &full[t]);
        }
    }
    for (uint32_t t = 0; t < w.n_aabb_tiles; ++t) {
        const int s = t & 1;
        pp_mbar_wait(&full[s], (t >> 1) & 1u);
        // union box of this warp's undecided edges (an edge that left the bounds or already hit is decided)
        float wminx = CUDART_INF_F, wmaxx = -CUDART_INF_F, wminy = CUDART_INF_F, wmaxy = -CUDART_INF_F;
#pragma unroll
        for (int e = 0; e < PP_SEGB_EPT; ++e) {
            if (good[e] && !hit[e]) {
                // NaN coordinates never pass pp_bounds_contains, so the boxes here are ordered
                wminx = fminf(wminx, eminx[e]);
                wmaxx = fmaxf(wmaxx, emaxx[e]);
                wminy = fminf(wminy, eminy[e]);
                wmaxy = fmaxf(wmaxy, emaxy[e]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            wminx = fminf(wminx, __shfl_xor_sync(0xffffffffu, wminx, o));
            wmaxx = fmaxf(wmaxx, __shfl_xor_sync(0xffffffffu, wmaxx, o));
            wminy = fminf(wminy, __shfl_xor_sync(0xffffffffu, wminy, o));
            wmaxy = fmaxf(wmaxy, __shfl_xor_sync(0xffffffffu, wmaxy, o));
        }
        if (wminx <= wmaxx) {  // some edge still undecided (warp-uniform)
            const float4 *T = tiles[s];
#pragma unroll 2
            for (int c = 0; c < PP_AABB_TILE / 32; ++c) {
                const float4 bb = T[c * 32 + lane];
                unsigned mask = __ballot_sync(0xffffffffu, !(wmaxx < bb.x || wminx > bb.z || wmaxy < bb.y || wminy > bb.w));
                while (mask) {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float rminx = __shfl_sync(0xffffffffu, bb.x, j), rminy = __shfl_sync(0xffffffffu, bb.y, j);
                    const float rmaxx = __shfl_sync(0xffffffffu, bb.z, j), rmaxy = __shfl_sync(0xffffffffu, bb.w, j);
                    const uint32_t ring = t * PP_AABB_TILE + c * 32 + j;
                    if (ring >= w.n_rings) continue;  // padding boxes are empty and never get here; belt and braces
#pragma unroll
                    for (int e = 0; e < PP_SEGB_EPT; ++e) {
                        if (good[e] && !hit[e] &&
                            !(emaxx[e] < rminx || eminx[e] > rmaxx || emaxy[e] < rminy || eminy[e] > rmaxy)) {
                            const pp_ring_meta mt = w.meta[ring];
                            const double *rx = w.ox + mt.first, *ry = w.oy + mt.first;
                            // f64 padded-box rule per pair (see pp_collide_segments_kernel)
                            const bool sgx = x1[e] < x0[e], sgy = y1[e] < y0[e];
                            const bool seg_in =
                                !((sgx ? x0[e] : x1[e]) < mt.minx - mt.pad || (sgx ? x1[e] : x0[e]) > mt.maxx + mt.pad ||
                                  (sgy ? y0[e] : y1[e]) < mt.miny - mt.pad || (sgy ? y1[e] : y0[e]) > mt.maxy + mt.pad);
                            const int cls = pp_ring_circle_class(w, ring, x0[e], y0[e], x1[e], y1[e]);  // circle filter
                            if (cls == 0) continue;
                            if (cls == 1) {
                                hit[e] = true;
                                continue;
                            }
                            if ((seg_in && pp_ring_hits_segment(rx, ry, mt.count, x0[e], y0[e], x1[e], y1[e])) ||
                                (!pp_outside_padded(mt, x0[e], y0[e]) &&
                                 pp_point_position(rx, ry, mt.count, x0[e], y0[e]) == 1) ||
                                (!pp_outside_padded(mt, x1[e], y1[e]) &&
                                 pp_point_position(rx, ry, mt.count, x1[e], y1[e]) == 1))
                                hit[e] = true;
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0 && t + 2 < w.n_aabb_tiles) {
            pp_mbar_expect_tx(&full[s], PP_AABB_TILE * 16);
            pp_bulk_g2s(tiles[s], w.aabb32 + (size_t)(t + 2) * PP_AABB_TILE, PP_AABB_TILE * 16, &full[s]);
        }
    }
#pragma unroll
    for (int e = 0; e < PP_SEGB_EPT; ++e)
        if (ei[e] != 0xFFFFFFFFu) ok[ei[e]] = (good[e] && !hit[e]) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// kernel 3b: polylines, one warp per polyline.  Points come either from memory (CSR polylines,
// pp_verify_polylines) or are generated on the fly from a Dubins plan record (pp_collide_dubins):
// samples are consumed in registers and never written (config 5 would materialise ~100 GB).
// Each warp iteration covers 32 consecutive points = 31 segments (one point of overlap); lane k tests
// vertex k and segment k->k+1 (neighbour's point via shuffle); hit flags are reduced with a ballot so
// the whole warp leaves the polyline at the first hit.
// ------------------------------------------------------------------------------------------------
struct pp_points_csr {
    const double *px, *py;
    const uint32_t *off;
};

struct pp_points_dubins {
    const pp_dubins_plan *plans;
    const pp_plan_aux *aux;  // segment origins + sincos(syaw) per path, written by the plan kernel
    const double *ex, *ey;  // parent point appended after the samples
    // paths left to verify: the plan kernel's path-level test (path_box.cuh) has already answered the others.
    // todo == nullptr: every path (PP_COLLIDE_NO_CULL, or no plan-side test)
    const uint32_t *todo;
    const unsigned int *todo_count;
};

// ---- coarse pass of the Dubins verify kernel (paths of several chunks, all three segment lengths positive).  Lane L
// takes the FIRST point of chunk L of a batch of 31 chunks (a chunk = 32 consecutive points = 31 line segments) and its
// neighbour's first point = the chunk's last point.  With three positive lengths consecutive samples are exactly one
// step of path length apart, also across the junctions (src/dubins.rs:233-237: the next segment's first `pd` is the
// previous overshoot, -d - ll), and all of them lie ON the path (0 < pd <= l) -- a word with a zero-length segment is
// different (samples up to five steps off a segment's ends, even behind the start pose) and takes the fine pass only.
// So every point of the chunk lies on a curve of length <= S = 31.5 steps between the two end points: inside the
// ellipse with these foci and major axis S, hence
// inside the end points' box grown by h = sqrt(S^2 - c^2) / 2 (c = distance of the end points).  A chunk whose grown box
// (i) lies in "inside" cells of the bounds grid and (ii) meets no registered ring box cannot fail any of its own
// tests: it is skipped.  Before that (phase 1), the chunks' first points are actual polyline vertices: one outside the
// bounds or inside a ring's inner circle (pp_circle_class == 1, exact by margin) blocks the path at once -- in a dense
// world (C5: a quarter of the plane is covered) some coarse vertex nearly always is, and the path is decided without a
// fine pass.  Vertices in a ring's annulus and everything a box meets are left to the fine pass of that chunk.
#ifndef PP_POLY_COARSE
#define PP_POLY_COARSE 1  // A/B switch
#endif
#define PP_COARSE_MIN_POINTS (4u * 31u + 2u)  // below this a path goes straight to the fine pass
// phase 1: does the vertex (x, y) block the path on its own?  (outside the bounds, or inside the inner circle of a ring
// registered in its cell -- a ring is registered in every cell its padded box meets, so the vertex's cell knows every
// ring that can contain it)
__device__ __forceinline__ bool pp_coarse_vertex_blocked(const pp_world_view &w, double x, double y) {
    if (!pp_bounds_contains(w, x, y)) return true;
    if (w.n_rings == 0u) return false;
    const int cx = __double2int_rd((x - w.gminx) * w.ginv), cy = __double2int_rd((y - w.gminy) * w.ginv);
    if ((unsigned)cx >= (unsigned)w.gx || (unsigned)cy >= (unsigned)w.gy) return false;
    const uint32_t *cell = w.cell_start + (size_t)cy * w.gx + cx;
    const uint32_t k1 = __ldg(cell + 1);
    const float px0 = __double2float_rd(x), px1 = __double2float_ru(x), py0 = __double2float_rd(y), py1 = __double2float_ru(y);
    for (uint32_t k = __ldg(cell); k < k1; ++k) {
        const float4 bb = __ldg(w.cell_box + k);
        if (!(px1 < bb.x || px0 > bb.z || py1 < bb.y || py0 > bb.w) &&
            pp_ring_circle_class(w, __ldg(w.cell_items + k), x, y, x, y) == 1)
            return true;
    }
    return false;
}
// phase 2: does the chunk between (x, y) and (xn, yn) need the fine pass?  (its grown box leaves the "inside" cells of the
// bounds grid or meets a registered ring box)
__device__ __forceinline__ bool pp_coarse_chunk_flag(const pp_world_view &w, double x, double y, double xn, double yn, double S2) {
    const double dx = xn - x, dy = yn - y;
    const double c2 = dx * dx + dy * dy;
    const double t = S2 - c2;
    const double pad = 1e-9 * ((fabs(x) + fabs(y)) + 1.0);
    const double h = 0.5 * sqrt(t > 0.0 ? t : 0.0) + pad;
    const bool swx = xn < x, swy = yn < y;
    const double x0 = (swx ? xn : x) - h, x1 = (swx ? x : xn) + h, y0 = (swy ? yn : y) - h, y1 = (swy ? y : yn) + h;
    if (!((x1 - x0) < 1e300 && (y1 - y0) < 1e300)) return true;  // NaN / inf: the fine pass decides
    {   // (i) bounds cells under the box
        const int ix0 = __double2int_rd((x0 - w.bminx) * w.binvx), iy0 = __double2int_rd((y0 - w.bminy) * w.binvy);
        const int ix1 = __double2int_rd((x1 - w.bminx) * w.binvx), iy1 = __double2int_rd((y1 - w.bminy) * w.binvy);
        if (ix0 < 0 || iy0 < 0 || ix1 >= w.bgx || iy1 >= w.bgy || ix1 - ix0 > 3 || iy1 - iy0 > 3) return true;
        for (int iy = iy0; iy <= iy1; ++iy)
            for (int ix = ix0; ix <= ix1; ++ix)
                if (__ldg(w.bcls + (size_t)iy * w.bgx + ix) != 1) return true;
    }
    if (w.n_rings == 0u) return false;
    // (ii) rings registered under the box
    int cx0 = __double2int_rd((x0 - w.gminx) * w.ginv), cy0 = __double2int_rd((y0 - w.gminy) * w.ginv);
    int cx1 = __double2int_rd((x1 - w.gminx) * w.ginv), cy1 = __double2int_rd((y1 - w.gminy) * w.ginv);
    if (cx1 < 0 || cy1 < 0 || cx0 >= w.gx || cy0 >= w.gy) return false;
    cx0 = max(cx0, 0);
    cy0 = max(cy0, 0);
    cx1 = min(cx1, w.gx - 1);
    cy1 = min(cy1, w.gy - 1);
    if (cx1 - cx0 > 7 || cy1 - cy0 > 7) return true;  // a huge step: no coarse decision
    const float q0x = __double2float_rd(x0), q0y = __double2float_rd(y0), q1x = __double2float_ru(x1), q1y = __double2float_ru(y1);
    for (int r = cy0; r <= cy1; ++r) {
        const uint32_t *row = w.cell_start + (size_t)r * w.gx;
        const uint32_t k1 = __ldg(row + cx1 + 1);
        for (uint32_t k = __ldg(row + cx0); k < k1; ++k) {
            const float4 bb = __ldg(w.cell_box + k);
            if (!(q1x < bb.x || q0x > bb.z || q1y < bb.y || q0y > bb.w)) return true;
        }
    }
    return false;
}

#define PP_POLY_THREADS 128
#ifndef PP_POLY_MIN_BLOCKS
#define PP_POLY_MIN_BLOCKS 7  // 72 registers; round 2 (circle filter, cell-ordered boxes): C5 slice / Dubins extend / no-hit 1.56 / 1.98 / 3.35 ms
                              // (5 blocks, 96 registers: 1.79 / 2.19 / 3.42; 8 blocks, 64 registers: 2.41 / 3.21 / 3.37)
#endif

template <bool CULL, bool DUBINS>
__global__ void __launch_bounds__(PP_POLY_THREADS, PP_POLY_MIN_BLOCKS)
    pp_verify_polylines_kernel(pp_world_view w, size_t n_lines, pp_points_csr csr, pp_points_dubins dub,
                               uint8_t *__restrict__ ok) {
    unsigned tid;  // read once: left to itself the compiler re-reads the special register all over the chunk loop
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int lane = (int)(tid & 31u), wib = (int)(tid >> 5);
    const size_t warps_total = (size_t)gridDim.x * (PP_POLY_THREADS / 32);
    const size_t n_work = (DUBINS && dub.todo) ? (size_t)__ldg(dub.todo_count) : n_lines;
    for (size_t item = (size_t)blockIdx.x * (PP_POLY_THREADS / 32) + wib; item < n_work; item += warps_total) {
        const size_t line = (DUBINS && dub.todo) ? (size_t)__ldg(dub.todo + item) : item;
        uint32_t np;      // points of this polyline
        uint32_t base = 0;
        // The plan record and the three segment origins are warp-uniform and indexed by a run-time segment
        // number: ONE copy per warp in shared memory.  (As per-thread arrays they lived in local memory, 232 bytes
        // replicated per lane: ncu showed 513 MB of DRAM writes per launch for a kernel that outputs 1 MB.)
        __shared__ pp_dubins_plan s_plan[PP_POLY_THREADS / 32];
        __shared__ pp_plan_aux s_aux[PP_POLY_THREADS / 32];
        const pp_dubins_plan &pl = s_plan[wib];
        const pp_plan_aux &aux = s_aux[wib];
        const pp_seg_origin *o = s_aux[wib].o;
        uint32_t nsamp = 0;
        if (DUBINS) {
            __syncwarp();  // every lane is done with the previous polyline's record
            if (lane < (int)(sizeof(pp_dubins_plan) / 4))
                reinterpret_cast<uint32_t *>(&s_plan[wib])[lane] =
                    __ldg(reinterpret_cast<const uint32_t *>(dub.plans + line) + lane);
            if (lane < PP_PLAN_AUX_DOUBLES)
                reinterpret_cast<double *>(&s_aux[wib])[lane] =
                    __ldg(reinterpret_cast<const double *>(dub.aux + line) + lane);
            __syncwarp();
            if (pl.count == 0xFFFFFFFFu) {  // replay overflow: the reference would run out of memory; report blocked
                if (lane == 0) ok[line] = 0;
                continue;
            }
            if (pl.word == PP_WORD_NONE) {
                nsamp = 1;  // fallback [(sx, sy)] of src/rrt.rs:313
            } else {
                nsamp = pl.count;
            }
            np = nsamp + 1;
        } else {
            base = csr.off[line];
            np = csr.off[line + 1] - base;
        }
        bool bad = false;
        if (np == 0) {  // empty line: contains() and !intersects() are vacuously true
            if (lane == 0) ok[line] = 1;
            continue;
        }
        // point k of a Dubins polyline: 0 = the start pose, 1 .. nsamp - 1 = samples, nsamp = the parent point.  Every
        // lane evaluates a (clamped) sample slot, then the two special points are patched in: no divergence between
        // the lanes of a chunk except where a chunk straddles a segment junction
        auto dubins_point = [&](uint32_t k, double &x, double &y) {
            if (nsamp > 1u) {  // uniform; implies a feasible word
                const uint32_t kk = min(max(k, 1u), nsamp - 1u);
                double lx, ly, lyaw;
                pp_plan_sample_local(pl, o, kk, &lx, &ly, &lyaw);
                pp_local_to_world(aux.ss, aux.cs, pl.sx, pl.sy, lx, ly, &x, &y);
            }
            if (k == 0u) {  // slot 0 is exactly the start pose (0*cos + 0*sin + sx)
                x = pl.sx;
                y = pl.sy;
            }
            if (k == nsamp) {  // the parent point closes the polyline (one lane of the last chunk)
                x = __ldg(dub.ex + line);
                y = __ldg(dub.ey + line);
            }
        };
        // chunks are taken in batches of 31; bit c of `todo` = chunk c of the batch needs the fine pass
        const uint32_t n_chunks = (np <= 2u) ? 1u : (np + 29u) / 31u;  // chunk c: points 31 c .. 31 c + 31, while 31 c + 1 < np
#if PP_POLY_COARSE
        // (three positive lengths: a zero-length segment breaks the one-step spacing the chunk boxes rely on, see above)
        // (tested on the products of neighbouring lengths, as the reference's branch at src/dubins.rs:233 is)
        const bool coarse = DUBINS && CULL && nsamp > 1u && np >= PP_COARSE_MIN_POINTS && pl.len[0] * pl.len[1] > 0.0 &&
                            pl.len[1] * pl.len[2] > 0.0 && pl.len[1] > 0.0;
#else
        const bool coarse = false;
#endif
        for (uint32_t c0 = 0; c0 < n_chunks && !bad; c0 += 31u) {
        const uint32_t n_batch = min(31u, n_chunks - c0);
        uint32_t todo = (n_batch >= 32u) ? 0xFFFFFFFFu : ((1u << n_batch) - 1u);
        if (coarse) {
            double x = 0.0, y = 0.0;
            dubins_point(min(31u * (c0 + (uint32_t)lane), np - 1u), x, y);
            const double xn = __shfl_down_sync(0xffffffffu, x, 1), yn = __shfl_down_sync(0xffffffffu, y, 1);
            // path length between the first and the last point of a chunk: 31 gaps of exactly one step (half a step
            // of slack for the rounding of the `pd` values)
            const double S = 31.5 * pl.step * pl.rinv;
            // lanes 0 .. n_batch hold actual vertices (lane n_batch: the end point of the batch's last chunk)
            const bool blocked = (uint32_t)lane <= n_batch && pp_coarse_vertex_blocked(w, x, y);
            if (__ballot_sync(0xffffffffu, blocked) != 0u) {
                bad = true;
                break;
            }
            // a chunk's box is tested unless the chunk is the path's last one (its tail is the parent point, off the
            // curve): that one always takes the fine pass
            bool fine = (uint32_t)lane < n_batch;
            if (fine && c0 + (uint32_t)lane + 1u < n_chunks) fine = pp_coarse_chunk_flag(w, x, y, xn, yn, S * S);
            todo &= __ballot_sync(0xffffffffu, fine);
        }
        for (; todo != 0u; todo &= todo - 1u) {
            const uint32_t k0 = 31u * (c0 + (uint32_t)__ffs((int)todo) - 1u);
            const uint32_t k = k0 + lane;
            double x = 0.0, y = 0.0;
            const bool have = k < np;
            if (DUBINS) {
                dubins_point(k, x, y);
            } else if (have) {
                x = csr.px[base + k];
                y = csr.py[base + k];
            }
            const double xn = __shfl_down_sync(0xffffffffu, x, 1);
            const double yn = __shfl_down_sync(0xffffffffu, y, 1);
            // vertex k is owned by this iteration unless it is the overlap point (lane 31 with more to come)
            const bool own_vertex = have && (lane < 31 || k + 1 == np);
            const bool own_segment = (lane < 31) && (k + 1 < np);
            bool fail = false;
            if (!CULL) {  // exhaustive per-lane loops (PP_COLLIDE_NO_CULL)
                if (own_vertex) fail = !pp_bounds_contains(w, x, y) || pp_vertex_in_any_obstacle(w, x, y);
                if (!fail && own_segment) fail = pp_segment_hits_any_obstacle(w, x, y, xn, yn);
                if (__ballot_sync(0xffffffffu, fail) != 0u) {
                    bad = true;
                    break;
                }
                continue;
            }
            if (own_vertex) fail = !pp_bounds_contains(w, x, y);
            if (__ballot_sync(0xffffffffu, fail) != 0u) {
                bad = true;
                break;
            }
            // Broad phase per lane, narrow phase per warp.  A lane walks the grid cells under the box of its
            // segment (of its vertex, for the last point) and stops at the first ring whose outward-rounded fp32
            // box meets it; the warp then takes the pending (lane, ring) candidates one by one, broadcasts the
            // lane's segment and runs the exact predicates with one ring segment per lane.  A ring that contains
            // vertex k has a padded box that contains it, hence meets the segment's box: one walk serves both the
            // vertex test and the segment test.
            const uint32_t seg_mask = __ballot_sync(0xffffffffu, own_segment);
            // Cell range of the box: floor is monotonic, so the cells of the box's corners are the min / max of the
            // two end points' cells -- each lane converts its own point once (saturating floor, NaN -> 0), takes the
            // neighbour's cell by shuffle, and everything else is integer arithmetic (f64 min / max cost ~13
            // instructions apiece on this part: no DMNMX).
            const int ix = __double2int_rd((x - w.gminx) * w.ginv), iy = __double2int_rd((y - w.gminy) * w.ginv);
            const int ixn = __shfl_down_sync(0xffffffffu, ix, 1), iyn = __shfl_down_sync(0xffffffffu, iy, 1);
            int cx0 = own_segment ? min(ix, ixn) : ix, cx1 = own_segment ? max(ix, ixn) : ix;
            int cy = own_segment ? min(iy, iyn) : iy, cy1 = own_segment ? max(iy, iyn) : iy;
            bool more = own_vertex && w.n_rings != 0u && !(cx1 < 0 || cy1 < 0 || cx0 >= w.gx || cy >= w.gy);
            cx0 = max(cx0, 0);
            cx1 = min(cx1, w.gx - 1);
            cy = max(cy, 0);
            cy1 = min(cy1, w.gy - 1);
            uint32_t kcur = 0, kend = 0;
            bool linear = false;
            // user-supplied polylines may hold segments whose box covers more cells than there are rings: walk
            // the ring list instead (bounds the cost per segment by O(rings)); Dubins samples are a step apart
            if (!DUBINS && more &&
                (unsigned long long)(cx1 - cx0 + 1) * (unsigned long long)(cy1 - cy + 1) > (unsigned long long)w.n_rings + 64ull) {
                linear = true;
                kend = w.n_rings;
                cy = cy1 + 1;
            } else if (more) {
                // The first row's run is fetched here and handed to the walk (a sample segment nearly always lies in ONE
                // row of cells): an empty single-row run ends the lane's work after two loads, exactly as the separate
                // "rings registered under the box" count did, without loading the same offsets twice.
                const uint32_t *row = w.cell_start + (size_t)cy * w.gx;
                kcur = __ldg(row + cx0);
                kend = __ldg(row + cx1 + 1);
                ++cy;
                if (kcur == kend && cy > cy1) more = false;
            }
            if (__ballot_sync(0xffffffffu, more) == 0u) continue;
            // some lane has rings to look at: its box, rounded outward to fp32, for the per-ring overlap test
            const double xe = own_segment ? xn : x, ye = own_segment ? yn : y;
            const bool swx = xe < x, swy = ye < y;
            const float q32x0 = __double2float_rd(swx ? xe : x), q32x1 = __double2float_ru(swx ? x : xe);
            const float q32y0 = __double2float_rd(swy ? ye : y), q32y1 = __double2float_ru(swy ? y : ye);
            // the ids of a ROW of cells are one contiguous run of cell_items: the walk goes row by row
            bool inner = false;
            for (;;) {
                uint32_t ring = 0xFFFFFFFFu;
                while (more) {
                    if (kcur < kend) {
                        const float4 bb = __ldg((linear ? w.aabb32 : w.cell_box) + kcur);  // cell-ordered boxes
                        const uint32_t kk = kcur++;
                        if (!(q32x1 < bb.x || q32x0 > bb.z || q32y1 < bb.y || q32y0 > bb.w)) {
                            const uint32_t r = linear ? kk : __ldg(w.cell_items + kk);
                            // circle filter on this lane's segment (its vertex alone for the last point)
                            const int cls = pp_ring_circle_class(w, r, x, y, xe, ye);
                            if (cls == 1) {  // both ends inside the ring's inner circle: the polyline is blocked
                                inner = true;
                                more = false;
                                break;
                            }
                            if (cls == 2) {
                                ring = r;
                                break;
                            }
                        }
                    } else {
                        if (cy > cy1) {
                            more = false;
                            break;
                        }
                        const uint32_t *row = w.cell_start + (size_t)cy * w.gx;
                        kcur = __ldg(row + cx0);
                        kend = __ldg(row + cx1 + 1);
                        ++cy;
                    }
                }
                if (__ballot_sync(0xffffffffu, inner) != 0u) {
                    bad = true;
                    break;
                }
                uint32_t pend = __ballot_sync(0xffffffffu, ring != 0xFFFFFFFFu);
                if (pend == 0u) break;
                do {
                    const int grp = lane >> 3, sub = lane & 7;
                    const unsigned gmask = 0xFFu << (grp * 8);
                    const int src = pp_take_candidates(pend, grp);
                    const bool act = src >= 0;
                    const int sl = act ? src : 0;
                    const uint32_t rr = __shfl_sync(0xffffffffu, ring, sl);
                    const double ax0 = __shfl_sync(0xffffffffu, x, sl), ay0 = __shfl_sync(0xffffffffu, y, sl);
                    const double ax1 = __shfl_sync(0xffffffffu, xn, sl), ay1 = __shfl_sync(0xffffffffu, yn, sl);
                    pp_ring_meta mt;
                    mt.minx = mt.miny = mt.maxx = mt.maxy = mt.pad = 0.0;
                    mt.first = mt.count = 0u;
                    if (act) mt = w.meta[rr];
                    const double *rx = w.ox + mt.first, *ry = w.oy + mt.first;
                    // the f64 padded-box rule of the per-lane path decides whether the pair is tested at all
                    const bool sx = ax1 < ax0, sy = ay1 < ay0;
                    const bool do_seg = act && ((seg_mask >> sl) & 1u) &&
                                        !((sx ? ax0 : ax1) < mt.minx - mt.pad || (sx ? ax1 : ax0) > mt.maxx + mt.pad ||
                                          (sy ? ay0 : ay1) < mt.miny - mt.pad || (sy ? ay1 : ay0) > mt.maxy + mt.pad);
                    bool hit = pp_ring_hits_segment_g8(rx, ry, do_seg ? mt.count : 0u, ax0, ay0, ax1, ay1, sub, gmask);
                    const bool do_vtx = act && !hit && !pp_outside_padded(mt, ax0, ay0);
                    hit = pp_point_inside_ring_g8(rx, ry, do_vtx ? mt.count : 0u, ax0, ay0, sub, gmask) || hit;
                    if (__ballot_sync(0xffffffffu, hit) != 0u) bad = true;
                } while (pend != 0u && !bad);
                if (bad) break;
            }
            if (bad) break;
        }
        }
        if (lane == 0) ok[line] = bad ? 0 : 1;
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
int pp_launch_collide_segments(pp_ctx *ctx, size_t m, const double *ax, const double *ay, const double *bx,
                               const double *by, const uint32_t *gather_idx, double *yaw_out, uint8_t *ok, int flags,
                               cudaStream_t stream) {
    if (m == 0) return PP_OK;
    pp_world_view w = pp_make_world_view(ctx->world);
    const unsigned grid = (unsigned)((m + PP_SEG_THREADS - 1) / PP_SEG_THREADS);
    const double *nx = ctx->tree.x, *ny = ctx->tree.y;
    if (flags & PP_COLLIDE_NO_CULL) {
        pp_launch_scope scope(ctx, "collide_segments_nocull");
        pp_collide_segments_kernel<1><<<grid, PP_SEG_THREADS, 0, stream>>>(w, m, ax, ay, bx, by, gather_idx, nx, ny,
                                                                           yaw_out, ok);
    } else if ((flags & PP_COLLIDE_USE_GRID) || !(flags & (PP_COLLIDE_UNSORTED | PP_COLLIDE_SCAN))) {
        // default: the obstacle grid built by pp_obstacles_upload (0.19 / 0.05 ms against 1.32 / 0.21 ms for the
        // binned tiled scan on the C4 hit / no-hit sets; identical flags)
        pp_launch_scope scope(ctx, "collide_segments_grid");
        pp_collide_segments_grid_kernel<<<grid, PP_SEG_THREADS, 0, stream>>>(w, m, ax, ay, bx, by, gather_idx, nx, ny,
                                                                             yaw_out, ok);
    } else if (flags & PP_COLLIDE_UNSORTED) {
        pp_launch_scope scope(ctx, "collide_segments_unsorted");
        pp_collide_segments_kernel<0><<<grid, PP_SEG_THREADS, 0, stream>>>(w, m, ax, ay, bx, by, gather_idx, nx, ny,
                                                                           yaw_out, ok);
    } else {
        uint32_t *perm = nullptr;
        int rc = pp_build_bucket_perm(ctx, m, ax, ay, &perm, stream);
        if (rc) return rc;
        const unsigned grid_b = (unsigned)((m + PP_SEG_THREADS * PP_SEGB_EPT - 1) / (PP_SEG_THREADS * PP_SEGB_EPT));
        pp_launch_scope scope(ctx, "collide_segments");
        pp_collide_segments_bucketed_kernel<<<grid_b, PP_SEG_THREADS, 0, stream>>>(w, m, ax, ay, bx, by, gather_idx, nx,
                                                                                   ny, yaw_out, perm, ok);
    }
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

// fused extend step (NN + yaw + verify in one launch, queries binned by node-grid block).  The caller has made the
// node grid usable (pp_nn_prepare).  Queries below PP_EXT_SORT_MIN keep their order (the three binning launches would
// cost more than they save).
#define PP_EXT_SORT_MIN 8192
int pp_launch_rrt_extend_fused(pp_ctx *ctx, size_t m, const double *qx, const double *qy, uint32_t *idx, double *yaw,
                               uint8_t *ok, cudaStream_t stream) {
    if (m == 0) return PP_OK;
    if (m >= 0xFFFFFFF0ull) return pp_fail(ctx, PP_ERR_INVALID, "too many queries for one extend call");
    const pp_nn_grid_view g = pp_nn_make_grid_view(ctx->tree);
    const pp_world_view w = pp_make_world_view(ctx->world);
    uint32_t *perm = nullptr;
    if (m >= PP_EXT_SORT_MIN && g.grid_n > 0) {
        pp_extend_bins b;
        b.shift = PP_EXT_BIN_SHIFT_MIN;
        for (;;) {
            b.bx = ((g.gx - 1) >> b.shift) + 1;
            b.by = ((g.gy - 1) >> b.shift) + 1;
            if ((uint64_t)b.bx * (uint64_t)b.by <= PP_EXT_MAX_BINS) break;
            ++b.shift;
        }
        const uint32_t nb = (uint32_t)b.bx * (uint32_t)b.by;
        const size_t hist_words = 1024 * PP_EXT_SCAN_PER;  // the scan block always covers 65 536 slots (zero beyond nb)
        int rc = pp_scratch_reserve(ctx, (hist_words + 3 * m) * 4);
        if (rc) return rc;
        uint32_t *hist = (uint32_t *)ctx->scratch, *bin_of = hist + hist_words, *rank = bin_of + m;
        perm = rank + m;
        pp_launch_scope scope(ctx, "extend_sort", 3);
        PP_CUDA(ctx, cudaMemsetAsync(hist, 0, hist_words * 4, stream));
        const unsigned g1 = (unsigned)((m + 255) / 256);
        pp_extend_bin_kernel<<<g1, 256, 0, stream>>>(g, b, qx, qy, (uint32_t)m, hist, bin_of, rank);
        pp_extend_scan_kernel<<<1, 1024, 0, stream>>>(hist);
        pp_extend_scatter_kernel<<<g1, 256, 0, stream>>>((uint32_t)m, hist, bin_of, rank, perm);
        PP_CUDA(ctx, cudaGetLastError());
    }
    pp_launch_scope scope(ctx, "extend_fused");
    const unsigned grid = (unsigned)((m + PP_SEG_THREADS - 1) / PP_SEG_THREADS);
    pp_rrt_extend_fused_kernel<<<grid, PP_SEG_THREADS, 0, stream>>>(g, w, (uint32_t)m, perm, qx, qy, idx, yaw, ok);
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

static unsigned pp_poly_grid(pp_ctx *ctx, size_t n_lines) {
    size_t blocks = (n_lines + (PP_POLY_THREADS / 32) - 1) / (PP_POLY_THREADS / 32);
    size_t max_blocks = (size_t)ctx->sm_count * 32;
    return (unsigned)(blocks < max_blocks ? blocks : max_blocks);
}

int pp_launch_verify_polylines(pp_ctx *ctx, size_t n_lines, const double *px, const double *py, const uint32_t *off,
                               uint8_t *ok, int flags, cudaStream_t stream) {
    if (n_lines == 0) return PP_OK;
    pp_world_view w = pp_make_world_view(ctx->world);
    pp_points_csr csr{px, py, off};
    pp_points_dubins dub{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    pp_launch_scope scope(ctx, "verify_polylines");
    if (flags & PP_COLLIDE_NO_CULL)
        pp_verify_polylines_kernel<false, false><<<pp_poly_grid(ctx, n_lines), PP_POLY_THREADS, 0, stream>>>(
            w, n_lines, csr, dub, ok);
    else
        pp_verify_polylines_kernel<true, false><<<pp_poly_grid(ctx, n_lines), PP_POLY_THREADS, 0, stream>>>(
            w, n_lines, csr, dub, ok);
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

int pp_launch_collide_dubins(pp_ctx *ctx, size_t m, const void *plans, const void *aux, const double *ex,
                             const double *ey, uint8_t *ok, int flags, const uint32_t *todo, const unsigned int *todo_count,
                             cudaStream_t stream) {
    if (m == 0) return PP_OK;
    pp_world_view w = pp_make_world_view(ctx->world);
    pp_points_csr csr{nullptr, nullptr, nullptr};
    pp_launch_scope scope(ctx, "collide_dubins");
    if (flags & PP_COLLIDE_NO_CULL) {
        pp_points_dubins dub{(const pp_dubins_plan *)plans, (const pp_plan_aux *)aux, ex, ey, nullptr, nullptr};
        pp_verify_polylines_kernel<false, true><<<pp_poly_grid(ctx, m), PP_POLY_THREADS, 0, stream>>>(w, m, csr, dub,
                                                                                                      ok);
    } else {
        // with a todo list the number of paths is only known on the device: the grid is sized for all m, warps beyond
        // the list's length leave at once
        pp_points_dubins dub{(const pp_dubins_plan *)plans, (const pp_plan_aux *)aux, ex, ey, todo, todo_count};
        pp_verify_polylines_kernel<true, true><<<pp_poly_grid(ctx, m), PP_POLY_THREADS, 0, stream>>>(w, m, csr, dub,
                                                                                                     ok);
    }
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

// uploads this translation unit's copy of the math coefficient tables (pp_math.cuh) to the current device
int pp_collide_tu_init(pp_ctx *ctx) {
    PP_CUDA(ctx, pp_math_upload_tables());
    return PP_OK;
}

// ------------------------------------------------------------------------------------------------
// extend step with Dubins edges: Node::new for every sample (src/rrt.rs:169-175, 267-271) -- the new
// node's yaw aims at its nearest node, which becomes the goal pose of the edge
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    pp_extend_gather_kernel(size_t m, const double *__restrict__ qx, const double *__restrict__ qy,
                            const uint32_t *__restrict__ idx, const double *__restrict__ node_x,
                            const double *__restrict__ node_y, const double *__restrict__ node_yaw,
                            double *__restrict__ syaw, double *__restrict__ ex, double *__restrict__ ey,
                            double *__restrict__ eyaw) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    const uint32_t p = idx[i];
    if (p == 0xFFFFFFFFu) {  // no nearest node: a NaN goal pose has no feasible word and is never `contained`
        ex[i] = ey[i] = eyaw[i] = syaw[i] = CUDART_NAN;
        return;
    }
    const double px = node_x[p], py = node_y[p];
    ex[i] = px;
    ey[i] = py;
    eyaw[i] = node_yaw[p];
    syaw[i] = atan2(py - qy[i], px - qx[i]);  // compute_yaw(from = the new point, to = its parent)
}

int pp_launch_extend_gather(pp_ctx *ctx, size_t m, const double *qx, const double *qy, const uint32_t *idx,
                            double *syaw, double *ex, double *ey, double *eyaw, cudaStream_t stream) {
    if (m == 0) return PP_OK;
    pp_launch_scope scope(ctx, "extend_gather");
    pp_extend_gather_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(m, qx, qy, idx, ctx->tree.x, ctx->tree.y,
                                                                            ctx->tree.yaw, syaw, ex, ey, eyaw);
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

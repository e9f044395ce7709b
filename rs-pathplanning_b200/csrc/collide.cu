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
#include "pp_common.cuh"

pp_world_view pp_make_world_view(const pp_world_dev &w);  // api.cu
int pp_build_bucket_perm(pp_ctx *ctx, size_t m, const double *kx, const double *ky, uint32_t **perm_out,
                         cudaStream_t stream);  // nn.cu

#define PP_AABB_TILE 1024  // ring boxes per shared-memory tile (16 KB)

// ---- Polygon::contains(&Point) for the bounds ring ------------------------------------------------
__device__ __forceinline__ bool pp_bounds_contains(const pp_world_view &w, double x, double y) {
    const double fx = (x - w.bminx) * w.binvx, fy = (y - w.bminy) * w.binvy;
    if (!(fx >= 0.0 && fx < (double)w.bgx && fy >= 0.0 && fy < (double)w.bgy)) return false;  // also NaN
    const uint8_t cls = __ldg(w.bcls + (size_t)(int)fy * w.bgx + (int)fx);
    if (cls != 2) return cls == 1;
    return pp_point_position(w.bx, w.by, w.nb, x, y) == 1;
}

__device__ __forceinline__ bool pp_point_in_ring(const pp_world_view &w, const pp_ring_meta &m, double x, double y) {
    return pp_point_position(w.ox + m.first, w.oy + m.first, m.count, x, y) == 1;
}
__device__ __forceinline__ bool pp_outside_padded(const pp_ring_meta &m, double x, double y) {
    return (x < m.minx - m.pad) || (x > m.maxx + m.pad) || (y < m.miny - m.pad) || (y > m.maxy + m.pad);
}

__device__ __forceinline__ int pp_cell_clamp(double f, int g) {
    // f may be NaN/inf: comparisons first, conversion only for in-range values
    if (!(f > 0.0)) return 0;
    if (f >= (double)g) return g - 1;
    return (int)f;
}

// ---- any obstacle polygon contains the point? ------------------------------------------------------
template <bool CULL>
__device__ __forceinline__ bool pp_vertex_in_obstacle(const pp_world_view &w, double x, double y) {
    if (!CULL) {
        for (uint32_t r = 0; r < w.n_rings; ++r)
            if (pp_point_in_ring(w, w.meta[r], x, y)) return true;
        return false;
    }
    if (w.n_rings == 0) return false;
    const double fx = (x - w.gminx) * w.ginv, fy = (y - w.gminy) * w.ginv;
    if (!(fx >= 0.0 && fx < (double)w.gx && fy >= 0.0 && fy < (double)w.gy)) return false;
    const size_t c = (size_t)(int)fy * w.gx + (int)fx;
    const uint32_t c0 = __ldg(w.cell_start + c), c1 = __ldg(w.cell_start + c + 1);
    for (uint32_t k = c0; k < c1; ++k) {
        const pp_ring_meta m = w.meta[__ldg(w.cell_items + k)];
        if (pp_outside_padded(m, x, y)) continue;
        if (pp_point_in_ring(w, m, x, y)) return true;
    }
    return false;
}

// ---- any obstacle ring meets the segment? ----------------------------------------------------------
// LONG_SEGMENTS: the caller's segments may be long compared with the grid (user-supplied straight edges); the
// polyline kernels pass false (sample spacing <= a cell or two) and skip the extra branch.
template <bool CULL, bool LONG_SEGMENTS = false>
__device__ __forceinline__ bool pp_segment_hits_obstacle(const pp_world_view &w, double x0, double y0, double x1,
                                                         double y1) {
    if (!CULL) {
        for (uint32_t r = 0; r < w.n_rings; ++r) {
            const pp_ring_meta m = w.meta[r];
            if (pp_ring_hits_segment(w.ox + m.first, w.oy + m.first, m.count, x0, y0, x1, y1)) return true;
        }
        return false;
    }
    if (w.n_rings == 0) return false;
    const double sminx = fmin(x0, x1), smaxx = fmax(x0, x1), sminy = fmin(y0, y1), smaxy = fmax(y0, y1);
    const double fx0 = (sminx - w.gminx) * w.ginv, fx1 = (smaxx - w.gminx) * w.ginv;
    const double fy0 = (sminy - w.gminy) * w.ginv, fy1 = (smaxy - w.gminy) * w.ginv;
    if (fx1 < 0.0 || fy1 < 0.0 || fx0 >= (double)w.gx || fy0 >= (double)w.gy) return false;
    const int cx0 = pp_cell_clamp(fx0, w.gx), cx1 = pp_cell_clamp(fx1, w.gx);
    const int cy0 = pp_cell_clamp(fy0, w.gy), cy1 = pp_cell_clamp(fy1, w.gy);
    // a long segment whose box covers more cells than there are rings (not the RRT's short edges): walking the
    // ring list once is cheaper than walking the cells, and bounds the cost per segment by O(rings)
    if (LONG_SEGMENTS &&
        (unsigned long long)(cx1 - cx0 + 1) * (unsigned long long)(cy1 - cy0 + 1) > (unsigned long long)w.n_rings + 64ull) {
        for (uint32_t r = 0; r < w.n_rings; ++r) {
            const pp_ring_meta m = w.meta[r];
            if (smaxx < m.minx - m.pad || sminx > m.maxx + m.pad || smaxy < m.miny - m.pad || sminy > m.maxy + m.pad)
                continue;
            if (pp_ring_hits_segment(w.ox + m.first, w.oy + m.first, m.count, x0, y0, x1, y1)) return true;
        }
        return false;
    }
    for (int cy = cy0; cy <= cy1; ++cy)
        for (int cx = cx0; cx <= cx1; ++cx) {
            const size_t c = (size_t)cy * w.gx + cx;
            const uint32_t c0 = __ldg(w.cell_start + c), c1 = __ldg(w.cell_start + c + 1);
            for (uint32_t k = c0; k < c1; ++k) {
                const pp_ring_meta m = w.meta[__ldg(w.cell_items + k)];
                if (smaxx < m.minx - m.pad || sminx > m.maxx + m.pad || smaxy < m.miny - m.pad ||
                    sminy > m.maxy + m.pad)
                    continue;
                if (pp_ring_hits_segment(w.ox + m.first, w.oy + m.first, m.count, x0, y0, x1, y1)) return true;
            }
        }
    return false;
}

// ------------------------------------------------------------------------------------------------
// kernel 3a: straight 2-point edges, one thread per edge; the per-ring fp32 boxes of ALL obstacles
// stream through shared memory in 16 KB tiles (TMA bulk copy, 2-stage mbarrier ring); a ring whose
// box overlaps the edge's outward-rounded fp32 box takes the exact f64 test.  A warp whose edges are
// all decided (ballot) skips the box loop of the remaining tiles.
// MODE 0: tiled scan, one edge per thread in the caller's order   1: exhaustive, no cull   2: grid broad phase (default)
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
            const uint32_t g = gather_idx[i];
            x1 = node_x[g];
            y1 = node_y[g];
            if (yaw_out) yaw_out[i] = atan2(y1 - y0, x1 - x0);  // compute_yaw, src/rrt.rs:267-271
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
            hit = pp_segment_hits_obstacle<false>(w, x0, y0, x1, y1) || pp_vertex_in_obstacle<false>(w, x0, y0) ||
                  pp_vertex_in_obstacle<false>(w, x1, y1);
    } else if (MODE == 2) {
        if (good)
            hit = pp_segment_hits_obstacle<true, true>(w, x0, y0, x1, y1) || pp_vertex_in_obstacle<true>(w, x0, y0) ||
                  pp_vertex_in_obstacle<true>(w, x1, y1);
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
                            if (pp_ring_hits_segment(rx, ry, mt.count, x0, y0, x1, y1) ||
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
                x1[e] = node_x[g];
                y1[e] = node_y[g];
                if (yaw_out) yaw_out[i] = atan2(y1[e] - y0[e], x1[e] - x0[e]);  // compute_yaw, src/rrt.rs:267-271
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
                            if (pp_ring_hits_segment(rx, ry, mt.count, x0[e], y0[e], x1[e], y1[e]) ||
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
    const double *ex, *ey;  // parent point appended after the samples
};

#define PP_POLY_THREADS 128
#ifndef PP_POLY_MIN_BLOCKS
#define PP_POLY_MIN_BLOCKS 8  // 64 registers (some spills): measured faster than 80 / 96 / 109 registers on the C5 slice
#endif

template <bool CULL, bool DUBINS>
__global__ void __launch_bounds__(PP_POLY_THREADS, PP_POLY_MIN_BLOCKS)
    pp_verify_polylines_kernel(pp_world_view w, size_t n_lines, pp_points_csr csr, pp_points_dubins dub,
                               uint8_t *__restrict__ ok) {
    const int lane = threadIdx.x & 31;
    const size_t warps_total = (size_t)gridDim.x * (PP_POLY_THREADS / 32);
    for (size_t line = (size_t)blockIdx.x * (PP_POLY_THREADS / 32) + (threadIdx.x >> 5); line < n_lines;
         line += warps_total) {
        uint32_t np;      // points of this polyline
        uint32_t base = 0;
        pp_dubins_plan pl;
        pp_seg_origin o[3];
        double ss = 0.0, cs = 1.0, gx_unused, pex = 0.0, pey = 0.0;
        uint32_t nsamp = 0;
        if (DUBINS) {
            pl = dub.plans[line];
            pex = dub.ex[line];
            pey = dub.ey[line];
            if (pl.count == 0xFFFFFFFFu) {  // replay overflow: the reference would run out of memory; report blocked
                if (lane == 0) ok[line] = 0;
                continue;
            }
            if (pl.word == PP_WORD_NONE) {
                nsamp = 1;  // fallback [(sx, sy)] of src/rrt.rs:313
            } else {
                nsamp = pl.count;
                pp_segment_origins(pl, o, &gx_unused);
                pp_sincos1(pl.syaw, &ss, &cs);
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
        for (uint32_t k0 = 0; k0 == 0 || k0 + 1 < np; k0 += 31) {
            const uint32_t k = k0 + lane;
            double x = 0.0, y = 0.0;
            const bool have = k < np;
            if (have) {
                if (DUBINS) {
                    if (k >= nsamp) {
                        x = pex;
                        y = pey;
                    } else if (pl.word == PP_WORD_NONE || k == 0) {
                        x = pl.sx;  // sample 0 is exactly the start pose (0*cos + 0*sin + sx)
                        y = pl.sy;
                    } else {
                        double lx, ly, lyaw;
                        pp_plan_sample_local(pl, o, k, &lx, &ly, &lyaw);
                        pp_local_to_world(ss, cs, pl.sx, pl.sy, lx, ly, &x, &y);
                    }
                } else {
                    x = csr.px[base + k];
                    y = csr.py[base + k];
                }
            }
            const double xn = __shfl_down_sync(0xffffffffu, x, 1);
            const double yn = __shfl_down_sync(0xffffffffu, y, 1);
            // vertex k is owned by this iteration unless it is the overlap point (lane 31 with more to come)
            const bool own_vertex = have && (lane < 31 || k + 1 == np);
            const bool own_segment = (lane < 31) && (k + 1 < np);
            bool fail = false;
            if (own_vertex) fail = !pp_bounds_contains(w, x, y) || pp_vertex_in_obstacle<CULL>(w, x, y);
            if (!fail && own_segment) fail = pp_segment_hits_obstacle<CULL, !DUBINS>(w, x, y, xn, yn);
            if (__ballot_sync(0xffffffffu, fail) != 0u) {
                bad = true;
                break;
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
        // default: the obstacle grid built by pp_obstacles_upload (0.65 / 0.05 ms against 1.32 / 0.21 ms for the
        // binned tiled scan on the C4 hit / no-hit sets; identical flags)
        pp_launch_scope scope(ctx, "collide_segments_grid");
        pp_collide_segments_kernel<2><<<grid, PP_SEG_THREADS, 0, stream>>>(w, m, ax, ay, bx, by, gather_idx, nx, ny,
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
    pp_points_dubins dub{nullptr, nullptr, nullptr};
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

int pp_launch_collide_dubins(pp_ctx *ctx, size_t m, const void *plans, const double *ex, const double *ey,
                             uint8_t *ok, int flags, cudaStream_t stream) {
    if (m == 0) return PP_OK;
    pp_world_view w = pp_make_world_view(ctx->world);
    pp_points_csr csr{nullptr, nullptr, nullptr};
    pp_points_dubins dub{(const pp_dubins_plan *)plans, ex, ey};
    pp_launch_scope scope(ctx, "collide_dubins");
    if (flags & PP_COLLIDE_NO_CULL)
        pp_verify_polylines_kernel<false, true><<<pp_poly_grid(ctx, m), PP_POLY_THREADS, 0, stream>>>(w, m, csr, dub,
                                                                                                      ok);
    else
        pp_verify_polylines_kernel<true, true><<<pp_poly_grid(ctx, m), PP_POLY_THREADS, 0, stream>>>(w, m, csr, dub,
                                                                                                     ok);
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

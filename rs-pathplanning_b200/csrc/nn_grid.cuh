// nn_grid.cuh -- the exact uniform-grid nearest-neighbour search as a device function, shared by the stand-alone
// NN kernel (nn.cu) and the fused extend kernel (collide.cu).
#pragma once
#include "pp_common.cuh"

// The grid describes the nodes [0, grid_n); nodes appended since (the TAIL [grid_n, n_nodes), src/rrt.rs:586-589
// inserts one node per iteration) are not filed into cells but compared one by one BEFORE the cell search -- every
// lane of a warp reads the same tail node (one broadcast load), and the tighter `best` shortens the ring walk.
// Tail ids are larger than every grid id and the cell scan takes ties to the lower id, so the lexicographic
// (d2, id) minimum -- hence the lowest-index tie-break -- is unchanged.  The host rebuilds the grid when the tail
// outgrows its budget (pp_nn_grid_policy).
struct pp_nn_grid_view {
    uint32_t n_nodes, grid_n;
    const uint32_t *cell_start, *cell_items;
    const double2 *cell_xy;
    const double *node_x, *node_y;
    int gx, gy;
    double gminx, gminy, gcell, ginv;
};

// (bx_out, by_out): coordinates of the winner (undefined when bi_out == 0xFFFFFFFF) -- the extend step needs them
// for the edge and gets them from the scan instead of a second gather.
// Latency structure (ncu r02: 63 % of this function's samples were load-scoreboard stalls): the six row offsets of
// the 3 x 3 block are loaded before any of them is used, the coordinate loop runs one load ahead, and a candidate's
// node id is NOT fetched inside the loop -- the winner is tracked by its position in the cell arrays and its id is
// read once at the end; only an exact tie of two distances (rare) needs ids early.
// ALL 32 lanes of the warp call together (the tail is fetched cooperatively); a lane without a query passes
// live = false and gets bi_out = 0xFFFFFFFF.
__device__ __forceinline__ void pp_nn_grid_search(const pp_nn_grid_view &g, bool live, double x, double y, double &best_out,
                                                  uint32_t &bi_out, double &bx_out, double &by_out) {
    double best = CUDART_INF, bx = 0.0, by = 0.0;
    uint32_t bi = 0xFFFFFFFFu;  // id of the winner when known (src == 1)
    uint32_t bk = 0;            // position of the winner in the cell arrays (src == 2)
    int src = 0;                // 0: no node yet, 1: id known (tail node or resolved tie), 2: position known
    // the tail, in index order, strict compare (as the scans).  Every lane needs every tail node: the warp loads 32
    // nodes at a time, one per lane, and hands them round by shuffle -- one load latency per 32 nodes instead of one
    // per node (a per-thread loop over a 4 096-node tail spent ~1 ms waiting on L2, profiles/r02_append_latency.json)
    {
        const int lane = (int)(threadIdx.x & 31u);
        for (uint32_t base = g.grid_n; base < g.n_nodes; base += 32u) {
            const uint32_t mine = base + (uint32_t)lane;
            const double tx = (mine < g.n_nodes) ? __ldg(g.node_x + mine) : CUDART_INF;
            const double ty = (mine < g.n_nodes) ? __ldg(g.node_y + mine) : CUDART_INF;
            const int cnt = (int)min(32u, g.n_nodes - base);
            for (int k = 0; k < cnt; ++k) {
                const double nx = __shfl_sync(0xffffffffu, tx, k), ny = __shfl_sync(0xffffffffu, ty, k);
                const double dx = nx - x, dy = ny - y;
                const double v = dx * dx + dy * dy;
                if (v < best) {
                    best = v;
                    bi = base + (uint32_t)k;
                    bx = nx;
                    by = ny;
                    src = 1;
                }
            }
        }
    }
    // candidates k0 <= k < k1 of the cell-sorted arrays: same d2 arithmetic as the scans, (d2, id) lexicographic minimum
    auto scan = [&](uint32_t k0, uint32_t k1) {
        if (k0 >= k1) return;
        double2 p = __ldg(g.cell_xy + k0);
        for (uint32_t k = k0; k < k1; ++k) {
            const double2 q = p;
            if (k + 1 < k1) p = __ldg(g.cell_xy + k + 1);  // one load ahead of the arithmetic
            const double dx = q.x - x, dy = q.y - y;
            const double v = dx * dx + dy * dy;
            if (v < best) {
                best = v;
                bk = k;
                bx = q.x;
                by = q.y;
                src = 2;
            } else if (v == best && src != 0) {
                // exact tie: the lower id wins.  (d2 = +inf -- a query or node at 1e300, +-inf -- ties with the initial
                // best and must leave "no node", as the strict compare of the scans does: src == 0 skips it.)
                const uint32_t i = __ldg(g.cell_items + k);
                const uint32_t cur = (src == 1) ? bi : __ldg(g.cell_items + bk);
                if (i < cur) {
                    bx = q.x;
                    by = q.y;
                }
                bi = (i < cur) ? i : cur;
                src = 1;
            }
        }
    };
    if (live && g.grid_n > 0) {
        const int gx = g.gx, gy = g.gy;
        double fx = floor((x - g.gminx) * g.ginv), fy = floor((y - g.gminy) * g.ginv);
        int cx = (fx >= (double)gx) ? gx - 1 : ((fx > 0.0) ? (int)fx : 0);  // NaN -> 0
        int cy = (fy >= (double)gy) ? gy - 1 : ((fy > 0.0) ? (int)fy : 0);
        // rings 0 and 1 together: the 3 x 3 block is three runs of the cell-sorted arrays (a row of cells is
        // contiguous); all six offsets are requested before the first run is walked
        {
            const int xa = max(cx - 1, 0), xb = min(cx + 1, gx - 1);
            const int ya = max(cy - 1, 0), yb = min(cy + 1, gy - 1);
            uint32_t r0[3], r1[3];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int yy = min(ya + t, yb);
                const uint32_t *row = g.cell_start + (size_t)yy * gx;
                r0[t] = __ldg(row + xa);
                r1[t] = __ldg(row + xb + 1);
            }
#pragma unroll
            for (int t = 0; t < 3; ++t)
                if (ya + t <= yb) scan(r0[t], r1[t]);
        }
        const int maxr = max(gx, gy);
        for (int r = 2; r <= maxr; ++r) {
            if (src != 0) {
                double lim = (double)(r - 1) * g.gcell * (1.0 - 0x1p-30);
                if (best < lim * lim) break;
            }
            const int y0 = cy - r, y1 = cy + r, x0 = cx - r, x1 = cx + r;
            for (int yy = max(y0, 0); yy <= min(y1, gy - 1); ++yy) {
                const uint32_t *row = g.cell_start + (size_t)yy * gx;
                if (yy == y0 || yy == y1) {  // full row of the ring: one run
                    scan(__ldg(row + max(x0, 0)), __ldg(row + min(x1, gx - 1) + 1));
                } else {  // the two end cells
                    if (x0 >= 0) scan(__ldg(row + x0), __ldg(row + x0 + 1));
                    if (x1 < gx) scan(__ldg(row + x1), __ldg(row + x1 + 1));
                }
            }
        }
    }
    if (src == 2) bi = __ldg(g.cell_items + bk);
    best_out = best;
    bi_out = (src == 0 || !live) ? 0xFFFFFFFFu : bi;
    bx_out = bx;
    by_out = by;
}

pp_nn_grid_view pp_nn_make_grid_view(const pp_tree_dev &t);  // nn.cu (host)


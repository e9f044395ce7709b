// nn.cu -- exact nearest neighbour over the RRT tree's nodes (kernel 2).
// Compiled with -fmad=false: d2 = dx*dx + dy*dy is evaluated non-fused so that the argmin (lowest
// index on ties) is bit-exact with the oracle's definition of RRT::get_nearest_node
// (src/rrt.rs:378-391; metric of src/rrt.rs:239-246 squared -- see SURVEY B.2).
//
// Three kernels, one contract:
//   pp_nn_scan_kernel   many queries: each thread owns QPT queries, the node arrays stream through
//                       shared memory in tiles moved by the TMA engine (cp.async.bulk + mbarrier,
//                       3-stage ring).  With PREFILTER a node is first rejected in fp32: the query
//                       keeps outward-rounded float bounds [lo, hi] of the x-interval in which a
//                       node could still beat the current best (|x - qx| < sqrt_ru(best)); a node
//                       whose fl32(x) lies outside cannot have d2 < best, so skipping it is exact.
//                       Survivors take the f64 path.  Without PREFILTER every pair is evaluated in
//                       f64 (6 DP instructions per pair: the yard-stick kernel of SURVEY 8d).
//   pp_nn_wide_kernel   few queries (the scalar get_nearest_node call): all threads of the grid
//                       split the nodes of one query, per-thread running minimum, warp-shuffle
//                       argmin with lowest-index tie-break, last CTA reduces the partials.
//   pp_nn_grid_kernel   exact ring-expanding search over a uniform grid of the nodes.
#include "pp_common.cuh"

#define PP_NN_TILE 1024   // nodes per tile (tree arrays are padded to this with +inf sentinels)
#define PP_NN_STAGES 3
#define PP_NN_THREADS 128
#define PP_NN_QPT 4

struct __align__(128) pp_nn_stage {
    double x[PP_NN_TILE];
    double y[PP_NN_TILE];
    float x32[PP_NN_TILE];
};
#define PP_NN_STAGE_BYTES ((uint32_t)(PP_NN_TILE * (8 + 8 + 4)))
#define PP_NN_SMEM_BYTES (PP_NN_STAGES * sizeof(pp_nn_stage) + PP_NN_STAGES * sizeof(uint64_t))

__device__ __forceinline__ void pp_nn_issue_tile(pp_nn_stage *st, uint64_t *bar, const double *nx, const double *ny,
                                                 const float *nx32, uint32_t tile) {
    size_t base = (size_t)tile * PP_NN_TILE;
    pp_mbar_expect_tx(bar, PP_NN_STAGE_BYTES);
    pp_bulk_g2s(st->x, nx + base, PP_NN_TILE * 8, bar);
    pp_bulk_g2s(st->y, ny + base, PP_NN_TILE * 8, bar);
    pp_bulk_g2s(st->x32, nx32 + base, PP_NN_TILE * 4, bar);
}

template <bool PREFILTER>
__global__ void __launch_bounds__(PP_NN_THREADS)
    pp_nn_scan_kernel(const double *__restrict__ nx, const double *__restrict__ ny, const float *__restrict__ nx32,
                      uint32_t n_tiles, const double *__restrict__ qx, const double *__restrict__ qy, size_t m,
                      uint32_t *__restrict__ idx_out, double *__restrict__ d2_out) {
    extern __shared__ __align__(128) unsigned char pp_nn_smem[];
    pp_nn_stage *stages = reinterpret_cast<pp_nn_stage *>(pp_nn_smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(pp_nn_smem + PP_NN_STAGES * sizeof(pp_nn_stage));
    const int tid = threadIdx.x;

    // queries: q-th query of this thread = block_base + q*THREADS + tid (coalesced)
    const size_t block_base = (size_t)blockIdx.x * (PP_NN_THREADS * PP_NN_QPT);
    double x[PP_NN_QPT], y[PP_NN_QPT], best[PP_NN_QPT];
    float lo[PP_NN_QPT], hi[PP_NN_QPT];
    uint32_t bi[PP_NN_QPT];
#pragma unroll
    for (int q = 0; q < PP_NN_QPT; ++q) {
        size_t j = block_base + (size_t)q * PP_NN_THREADS + tid;
        bool live = j < m;
        x[q] = live ? qx[j] : 0.0;
        y[q] = live ? qy[j] : 0.0;
        best[q] = CUDART_INF;
        bi[q] = 0xFFFFFFFFu;
        lo[q] = -CUDART_INF_F;
        hi[q] = CUDART_INF_F;
    }

    if (tid == 0) {
        for (int s = 0; s < PP_NN_STAGES; ++s) pp_mbar_init(&full[s], 1);
        pp_fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (uint32_t t = 0; t < PP_NN_STAGES && t < n_tiles; ++t)
            pp_nn_issue_tile(&stages[t], &full[t], nx, ny, nx32, t);
    }

    for (uint32_t t = 0; t < n_tiles; ++t) {
        const int s = t % PP_NN_STAGES;
        pp_mbar_wait(&full[s], (t / PP_NN_STAGES) & 1u);
        const pp_nn_stage &T = stages[s];
        const uint32_t base = t * PP_NN_TILE;
        if (PREFILTER) {
#pragma unroll 2
            for (int j = 0; j < PP_NN_TILE; j += 4) {
                const float4 xs = *reinterpret_cast<const float4 *>(&T.x32[j]);
                const float xv[4] = {xs.x, xs.y, xs.z, xs.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    bool surv[PP_NN_QPT];
                    bool any = false;
#pragma unroll
                    for (int q = 0; q < PP_NN_QPT; ++q) {
                        surv[q] = !((xv[u] > hi[q]) || (xv[u] < lo[q]));
                        any |= surv[q];
                    }
                    if (any) {
                        const double nxv = T.x[j + u], nyv = T.y[j + u];
#pragma unroll
                        for (int q = 0; q < PP_NN_QPT; ++q) {
                            if (surv[q]) {
                                double dx = nxv - x[q], dy = nyv - y[q];
                                double v = dx * dx + dy * dy;
                                if (v < best[q]) {
                                    best[q] = v;
                                    bi[q] = base + j + u;
                                    // any node with |nx - qx| >= r has dx*dx >= best, hence d2 >= best
                                    double r = __dsqrt_ru(v);
                                    hi[q] = __double2float_ru(__dadd_ru(x[q], r));
                                    lo[q] = __double2float_rd(__dadd_rd(x[q], -r));
                                }
                            }
                        }
                    }
                }
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < PP_NN_TILE; ++j) {
                const double nxv = T.x[j], nyv = T.y[j];
#pragma unroll
                for (int q = 0; q < PP_NN_QPT; ++q) {
                    double dx = nxv - x[q], dy = nyv - y[q];
                    double v = dx * dx + dy * dy;
                    if (v < best[q]) {
                        best[q] = v;
                        bi[q] = base + j;
                    }
                }
            }
        }
        __syncthreads();  // every thread is done with stage s
        if (tid == 0 && t + PP_NN_STAGES < n_tiles)
            pp_nn_issue_tile(&stages[s], &full[s], nx, ny, nx32, t + PP_NN_STAGES);
    }

#pragma unroll
    for (int q = 0; q < PP_NN_QPT; ++q) {
        size_t j = block_base + (size_t)q * PP_NN_THREADS + tid;
        if (j < m) {
            idx_out[j] = bi[q];
            if (d2_out) d2_out[j] = best[q];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// few queries: grid = (G, m); CTA (g, j) scans nodes g*256 + tid, stride G*256.
// ---------------------------------------------------------------------------------------------
#define PP_NN_WIDE_THREADS 256

__device__ __forceinline__ void pp_argmin_combine(double &v, uint32_t &i, double ov, uint32_t oi) {
    if (ov < v || (ov == v && oi < i)) {
        v = ov;
        i = oi;
    }
}

__global__ void __launch_bounds__(PP_NN_WIDE_THREADS)
    pp_nn_wide_kernel(const double *__restrict__ nx, const double *__restrict__ ny, uint32_t n_nodes,
                      const double *__restrict__ qx, const double *__restrict__ qy, double *__restrict__ part_v,
                      uint32_t *__restrict__ part_i, unsigned int *__restrict__ tickets,
                      uint32_t *__restrict__ idx_out, double *__restrict__ d2_out) {
    __shared__ double sv[PP_NN_WIDE_THREADS / 32];
    __shared__ uint32_t si[PP_NN_WIDE_THREADS / 32];
    __shared__ bool is_last;
    const uint32_t j = blockIdx.y;
    const double x = qx[j], y = qy[j];
    double best = CUDART_INF;
    uint32_t bi = 0xFFFFFFFFu;
    for (uint32_t i = blockIdx.x * PP_NN_WIDE_THREADS + threadIdx.x; i < n_nodes;
         i += gridDim.x * PP_NN_WIDE_THREADS) {
        double dx = __ldg(nx + i) - x, dy = __ldg(ny + i) - y;
        double v = dx * dx + dy * dy;
        if (v < best) {  // increasing i per thread: first minimum = lowest index
            best = v;
            bi = i;
        }
    }
    // warp-shuffle argmin, (value, index) lexicographic so the lowest index wins ties
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_down_sync(0xffffffffu, best, o);
        uint32_t oi = __shfl_down_sync(0xffffffffu, bi, o);
        pp_argmin_combine(best, bi, ov, oi);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sv[warp] = best;
        si[warp] = bi;
    }
    __syncthreads();
    if (warp == 0) {
        best = (lane < PP_NN_WIDE_THREADS / 32) ? sv[lane] : CUDART_INF;
        bi = (lane < PP_NN_WIDE_THREADS / 32) ? si[lane] : 0xFFFFFFFFu;
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_down_sync(0xffffffffu, best, o);
            uint32_t oi = __shfl_down_sync(0xffffffffu, bi, o);
            pp_argmin_combine(best, bi, ov, oi);
        }
        if (lane == 0) {
            part_v[(size_t)j * gridDim.x + blockIdx.x] = best;
            part_i[(size_t)j * gridDim.x + blockIdx.x] = bi;
            __threadfence();
            unsigned int t = atomicAdd(&tickets[j], 1u);
            is_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (is_last && warp == 0) {
        __threadfence();
        best = CUDART_INF;
        bi = 0xFFFFFFFFu;
        for (uint32_t g = lane; g < gridDim.x; g += 32)
            pp_argmin_combine(best, bi, part_v[(size_t)j * gridDim.x + g], part_i[(size_t)j * gridDim.x + g]);
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_down_sync(0xffffffffu, best, o);
            uint32_t oi = __shfl_down_sync(0xffffffffu, bi, o);
            pp_argmin_combine(best, bi, ov, oi);
        }
        if (lane == 0) {
            idx_out[j] = bi;
            if (d2_out) d2_out[j] = best;
            tickets[j] = 0;  // re-arm for the next call
        }
    }
}

// ---------------------------------------------------------------------------------------------
// uniform grid (built by pp_tree_build_grid in api.cu): cell c holds node ids
// cell_items[cell_start[c] .. cell_start[c+1]) in ascending order.
// Search: rings of cells of Chebyshev radius r = 0, 1, ... around the query's (clamped) cell; every
// node in a cell at Chebyshev distance >= r is farther than (r-1)*cell from the query, so the search
// stops once best < ((r-1)*cell*(1-2^-30))^2.  Same d2 arithmetic, ties to the lowest index.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    pp_nn_grid_kernel(const double *__restrict__ nx, const double *__restrict__ ny, uint32_t n_nodes,
                      const uint32_t *__restrict__ cell_start, const uint32_t *__restrict__ cell_items, int gx, int gy,
                      double gminx, double gminy, double gcell, double ginv, const double *__restrict__ qx,
                      const double *__restrict__ qy, size_t m, uint32_t *__restrict__ idx_out,
                      double *__restrict__ d2_out) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double x = qx[j], y = qy[j];
    double best = CUDART_INF;
    uint32_t bi = 0xFFFFFFFFu;
    if (n_nodes > 0) {
        double fx = floor((x - gminx) * ginv), fy = floor((y - gminy) * ginv);
        int cx = (fx >= (double)gx) ? gx - 1 : ((fx > 0.0) ? (int)fx : 0);  // NaN -> 0
        int cy = (fy >= (double)gy) ? gy - 1 : ((fy > 0.0) ? (int)fy : 0);
        const int maxr = max(gx, gy);
        for (int r = 0; r <= maxr; ++r) {
            if (r >= 2 && bi != 0xFFFFFFFFu) {
                double lim = (double)(r - 1) * gcell * (1.0 - 0x1p-30);
                if (best < lim * lim) break;
            }
            const int y0 = cy - r, y1 = cy + r, x0 = cx - r, x1 = cx + r;
            for (int yy = max(y0, 0); yy <= min(y1, gy - 1); ++yy) {
                const bool edge_row = (yy == y0) || (yy == y1);
                const int xstep = edge_row ? 1 : max(x1 - x0, 1);
                for (int xx = x0; xx <= x1; xx += xstep) {
                    if (xx < 0 || xx >= gx) continue;
                    const uint32_t c0 = cell_start[(size_t)yy * gx + xx], c1 = cell_start[(size_t)yy * gx + xx + 1];
                    for (uint32_t k = c0; k < c1; ++k) {
                        const uint32_t i = cell_items[k];
                        double dx = __ldg(nx + i) - x, dy = __ldg(ny + i) - y;
                        double v = dx * dx + dy * dy;
                        if (v < best || (v == best && i < bi)) {
                            best = v;
                            bi = i;
                        }
                    }
                }
            }
        }
    }
    idx_out[j] = bi;
    if (d2_out) d2_out[j] = best;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int pp_nn_configure(pp_ctx *ctx) {
    PP_CUDA(ctx, cudaFuncSetAttribute(pp_nn_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)PP_NN_SMEM_BYTES));
    PP_CUDA(ctx, cudaFuncSetAttribute(pp_nn_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)PP_NN_SMEM_BYTES));
    return PP_OK;
}

size_t pp_nn_tile_nodes() { return PP_NN_TILE; }

int pp_launch_nn(pp_ctx *ctx, size_t m, const double *qx, const double *qy, uint32_t *idx, double *d2, int flags,
                 cudaStream_t stream) {
    if (m == 0) return PP_OK;
    const pp_tree_dev &t = ctx->tree;
    const uint32_t n_nodes = (uint32_t)t.n;
    if (flags & PP_NN_GRID) {
        pp_launch_scope scope(ctx, "nn_grid");
        pp_nn_grid_kernel<<<(unsigned)((m + 127) / 128), 128, 0, stream>>>(t.x, t.y, n_nodes, t.cell_start, t.cell_items,
                                                                          t.gx, t.gy, t.gminx, t.gminy, t.gcell, t.ginv,
                                                                          qx, qy, m, idx, d2);
        PP_CUDA(ctx, cudaGetLastError());
        return PP_OK;
    }
    const uint32_t n_tiles = (uint32_t)((t.n + PP_NN_TILE - 1) / PP_NN_TILE);
    // few queries: split the nodes instead of the queries
    if (m <= 64 && !(flags & PP_NN_PLAIN_F64)) {
        unsigned G = (unsigned)((t.n + PP_NN_WIDE_THREADS * 4 - 1) / (PP_NN_WIDE_THREADS * 4));
        unsigned maxG = (unsigned)ctx->sm_count * 4;
        if (G > maxG) G = maxG;
        if (G < 1) G = 1;
        size_t need = (size_t)m * G * (sizeof(double) + sizeof(uint32_t));
        int rc = pp_scratch_reserve(ctx, need);
        if (rc) return rc;
        // tickets live in their own zero-initialised allocation and are re-armed by the kernel
        unsigned int *tickets = ctx->tickets;
        double *part_v = (double *)ctx->scratch;
        uint32_t *part_i = (uint32_t *)(part_v + (size_t)m * G);
        pp_launch_scope scope(ctx, "nn_wide");
        pp_nn_wide_kernel<<<dim3(G, (unsigned)m), PP_NN_WIDE_THREADS, 0, stream>>>(t.x, t.y, n_nodes, qx, qy, part_v,
                                                                                   part_i, tickets, idx, d2);
        PP_CUDA(ctx, cudaGetLastError());
        return PP_OK;
    }
    const unsigned grid = (unsigned)((m + PP_NN_THREADS * PP_NN_QPT - 1) / (PP_NN_THREADS * PP_NN_QPT));
    if (flags & PP_NN_PLAIN_F64) {
        pp_launch_scope scope(ctx, "nn_scan_f64");
        pp_nn_scan_kernel<false><<<grid, PP_NN_THREADS, PP_NN_SMEM_BYTES, stream>>>(t.x, t.y, t.x32, n_tiles, qx, qy, m,
                                                                                     idx, d2);
    } else {
        pp_launch_scope scope(ctx, "nn_scan");
        pp_nn_scan_kernel<true><<<grid, PP_NN_THREADS, PP_NN_SMEM_BYTES, stream>>>(t.x, t.y, t.x32, n_tiles, qx, qy, m,
                                                                                    idx, d2);
    }
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

// fl32 copy of x (+inf sentinel padding is written by the uploader)
__global__ void pp_tree_x32_kernel(const double *__restrict__ x, float *__restrict__ x32, size_t first, size_t n) {
    size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x32[i] = __double2float_rn(x[i]);
}
__global__ void pp_tree_pad_kernel(double *x, double *y, float *x32, size_t first, size_t end) {
    size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < end) {
        x[i] = CUDART_INF;
        y[i] = CUDART_INF;
        x32[i] = CUDART_INF_F;
    }
}

int pp_launch_tree_finish(pp_ctx *ctx, size_t first, size_t n, size_t padded_end, cudaStream_t stream) {
    pp_tree_dev &t = ctx->tree;
    pp_launch_scope scope(ctx, "tree_finish", 2);
    if (n > first)
        pp_tree_x32_kernel<<<(unsigned)((n - first + 255) / 256), 256, 0, stream>>>(t.x, t.x32, first, n);
    if (padded_end > n)
        pp_tree_pad_kernel<<<(unsigned)((padded_end - n + 255) / 256), 256, 0, stream>>>(t.x, t.y, t.x32, n, padded_end);
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

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
#include <cstdlib>

#include "nn_grid.cuh"
#include "pp_common.cuh"

#define PP_NN_TILE 1024   // nodes per tile (tree arrays are padded to this with +inf sentinels)
#define PP_NN_STAGES 3
#define PP_NN_THREADS 128
#define PP_NN_QPT 4

// fp32 copies of the node coordinates live in ONE array, blocked so that a bucketed-scan tile is a single
// contiguous 16 KB TMA copy: block b = nodes [2048 b, 2048 b + 2048) stored as 2048 x-values then 2048 y-values
#define PP_XY32_BLOCK 2048
__host__ __device__ __forceinline__ size_t pp_xy32_index(size_t node) {  // index of fl32(x); fl32(y) is +2048
    return (node / PP_XY32_BLOCK) * (2 * PP_XY32_BLOCK) + (node % PP_XY32_BLOCK);
}

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
    pp_bulk_g2s(st->x32, nx32 + pp_xy32_index(base), PP_NN_TILE * 4, bar);
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
// cell_items[cell_start[c] .. cell_start[c+1]) (any order) with their coordinates in cell_xy[] at the same positions.
// Search: rings of cells of Chebyshev radius r = 0, 1, ... around the query's (clamped) cell; every
// node in a cell at Chebyshev distance >= r is farther than (r-1)*cell from the query, so the search
// stops once best < ((r-1)*cell*(1-2^-30))^2.  Same d2 arithmetic, ties to the lowest index.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    pp_nn_grid_kernel(pp_nn_grid_view g, const double *__restrict__ qx, const double *__restrict__ qy, size_t m,
                      uint32_t *__restrict__ idx_out, double *__restrict__ d2_out) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = j < m;  // no early return: the whole warp walks the tail together
    double best, bx, by;
    uint32_t bi;
    pp_nn_grid_search(g, live, live ? qx[j] : 0.0, live ? qy[j] : 0.0, best, bi, bx, by);
    if (live) {
        idx_out[j] = bi;
        if (d2_out) d2_out[j] = best;
    }
}


// ---------------------------------------------------------------------------------------------
// Bucketed scan (PP_NN_SCAN; the tiled brute-force design).  The queries are binned into the cells of a G x G grid over
// their bounding box (counting sort: histogram, single-block scan, scatter -> a permutation), so the 128
// queries of a warp lie in a small box.  The warp keeps ONE outward-rounded float box = union of its
// threads' rejection boxes [q - r, q + r] (r = sqrt_ru(best)) and tests 128 nodes per step: lane l compares
// fl32(x), fl32(y) of nodes 4l..4l+3 (two LDS.128) with the warp box, a ballot yields the (rare) candidate
// nodes, and only those go through the per-thread fp32 test and the exact f64 evaluation.  A node outside a
// thread's box has |dx| >= r or |dy| >= r, hence d2 >= best: skipping it is exact.  Every node is still
// visited for every query (brute force, no index structure on the tree), in increasing index order with a
// strict comparison, so the result is the same bit-exact argmin with lowest-index tie-break.  Only the fp32
// copies stream through shared memory (16 KB TMA tiles of 2048 nodes); x and y of a candidate are fetched
// from L2 (one broadcast load per warp).
// ---------------------------------------------------------------------------------------------
#define PP_NNS_TILE PP_XY32_BLOCK
#define PP_NNS_STAGES 3
#define PP_NNS_THREADS 128
#define PP_NNS_QPT 4
#define PP_NNS_TILE_BYTES (PP_NNS_TILE * 8)
#define PP_NNS_SMEM_BYTES (PP_NNS_STAGES * PP_NNS_TILE_BYTES + PP_NNS_STAGES * 8)

__device__ __forceinline__ unsigned long long pp_f64_key(double v) {  // order-preserving u64 key
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double pp_key_f64(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

// mm[0..1] = min / max key of the non-NaN query x, mm[2..3] of y (preset to {~0, 0, ~0, 0})
__global__ void __launch_bounds__(256)
    pp_nn_qrange_kernel(const double *__restrict__ qx, const double *__restrict__ qy, size_t m,
                        unsigned long long *__restrict__ mm) {
    unsigned long long lo[2] = {~0ull, ~0ull}, hi[2] = {0ull, 0ull};
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < m; i += (size_t)gridDim.x * 256) {
        const double v[2] = {qx[i], qy[i]};
#pragma unroll
        for (int a = 0; a < 2; ++a)
            if (v[a] == v[a]) {
                const unsigned long long k = pp_f64_key(v[a]);
                lo[a] = min(lo[a], k);
                hi[a] = max(hi[a], k);
            }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = min(lo[a], __shfl_down_sync(0xffffffffu, lo[a], o));
            hi[a] = max(hi[a], __shfl_down_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&mm[2 * a], lo[a]);
            atomicMax(&mm[2 * a + 1], hi[a]);
        }
    }
}

__device__ __forceinline__ uint32_t pp_nn_axis_cell(double v, unsigned long long klo, unsigned long long khi,
                                                    uint32_t g) {
    const double lo = pp_key_f64(klo), hi = pp_key_f64(khi);
    const double span = hi - lo;
    if (!(v == v) || !(span > 0.0)) return 0;
    const double f = (v - lo) * ((double)g / span);
    if (!(f > 0.0)) return 0;
    return (f >= (double)g) ? g - 1 : (uint32_t)f;
}
// row-major cell of the g x g grid; consecutive buckets are x-neighbours, so a warp that spans two
// buckets still covers a compact box
__device__ __forceinline__ uint32_t pp_nn_bucket_of(double x, double y, const unsigned long long *mm, uint32_t g) {
    return pp_nn_axis_cell(y, mm[2], mm[3], g) * g + pp_nn_axis_cell(x, mm[0], mm[1], g);
}

__global__ void __launch_bounds__(256)
    pp_nn_bucket_count_kernel(const double *__restrict__ qx, const double *__restrict__ qy, size_t m,
                              const unsigned long long *__restrict__ mm, uint32_t g, uint32_t *__restrict__ hist) {
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < m) atomicAdd(&hist[pp_nn_bucket_of(qx[i], qy[i], mm, g)], 1u);
}

// single block: exclusive scan of hist[nb] into cursor[nb] (nb <= 16384)
__global__ void __launch_bounds__(1024) pp_nn_bucket_scan_kernel(const uint32_t *__restrict__ hist, uint32_t nb,
                                                                 uint32_t *__restrict__ cursor) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = (i < nb) ? hist[i] : 0;
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane], winc = w;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_sums[lane] = winc - w;
        }
        __syncthreads();
        const uint32_t excl = carry + warp_sums[warp] + inc - v;
        if (i < nb) cursor[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
    pp_nn_bucket_scatter_kernel(const double *__restrict__ qx, const double *__restrict__ qy, size_t m,
                                const unsigned long long *__restrict__ mm, uint32_t g, uint32_t *__restrict__ cursor,
                                uint32_t *__restrict__ perm) {
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < m) perm[atomicAdd(&cursor[pp_nn_bucket_of(qx[i], qy[i], mm, g)], 1u)] = (uint32_t)i;
}

__global__ void __launch_bounds__(PP_NNS_THREADS)
    pp_nn_bucketed_kernel(const double *__restrict__ nx, const double *__restrict__ ny, const float *__restrict__ nxy32,
                          uint32_t n_tiles, const double *__restrict__ qx, const double *__restrict__ qy,
                          const uint32_t *__restrict__ perm, size_t m, uint32_t *__restrict__ idx_out,
                          double *__restrict__ d2_out) {
    extern __shared__ __align__(128) unsigned char pp_nns_smem[];
    float *tiles = reinterpret_cast<float *>(pp_nns_smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(pp_nns_smem + PP_NNS_STAGES * PP_NNS_TILE_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // the q-th query of this thread: consecutive positions of the bucket order within the warp
    const size_t warp_base = ((size_t)blockIdx.x * (PP_NNS_THREADS / 32) + warp) * (32 * PP_NNS_QPT);
    uint32_t qi[PP_NNS_QPT];
    double x[PP_NNS_QPT], y[PP_NNS_QPT], best[PP_NNS_QPT];
    float xlo[PP_NNS_QPT], xhi[PP_NNS_QPT], ylo[PP_NNS_QPT], yhi[PP_NNS_QPT];
    uint32_t bi[PP_NNS_QPT];
#pragma unroll
    for (int q = 0; q < PP_NNS_QPT; ++q) {
        const size_t pos = warp_base + (size_t)q * 32 + lane;
        const bool live = pos < m;
        qi[q] = live ? perm[pos] : 0xFFFFFFFFu;
        x[q] = live ? qx[qi[q]] : 0.0;
        y[q] = live ? qy[qi[q]] : 0.0;
        best[q] = CUDART_INF;
        bi[q] = 0xFFFFFFFFu;
        // dead slots get an empty box so that they never widen the warp's
        xlo[q] = ylo[q] = live ? -CUDART_INF_F : CUDART_INF_F;
        xhi[q] = yhi[q] = live ? CUDART_INF_F : -CUDART_INF_F;
    }

    if (tid == 0) {
        for (int s = 0; s < PP_NNS_STAGES; ++s) pp_mbar_init(&full[s], 1);
        pp_fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (uint32_t t = 0; t < PP_NNS_STAGES && t < n_tiles; ++t) {
            pp_mbar_expect_tx(&full[t], PP_NNS_TILE_BYTES);
            pp_bulk_g2s(tiles + (size_t)t * (2 * PP_NNS_TILE), nxy32 + (size_t)t * (2 * PP_NNS_TILE), PP_NNS_TILE_BYTES,
                        &full[t]);
        }
    }

    for (uint32_t t = 0; t < n_tiles; ++t) {
        const int s = t % PP_NNS_STAGES;
        pp_mbar_wait(&full[s], (t / PP_NNS_STAGES) & 1u);
        const uint32_t tile_addr = pp_smem_u32(tiles + (size_t)s * (2 * PP_NNS_TILE)) + (uint32_t)lane * 16u;
        const uint32_t base = t * PP_NNS_TILE;
        float wxlo = CUDART_INF_F, wxhi = -CUDART_INF_F, wylo = CUDART_INF_F, wyhi = -CUDART_INF_F;
        // 128 nodes per step: lane l holds nodes 4l .. 4l+3 of the chunk, so the candidate order (lane, then k)
        // is the node index order
#pragma unroll 2
        for (int c = 0; c < PP_NNS_TILE / 128; ++c) {
            if ((c & 3) == 0) {  // refresh the warp box every 512 nodes (a stale one is only larger)
                wxlo = fminf(fminf(xlo[0], xlo[1]), fminf(xlo[2], xlo[3]));
                wxhi = fmaxf(fmaxf(xhi[0], xhi[1]), fmaxf(xhi[2], xhi[3]));
                wylo = fminf(fminf(ylo[0], ylo[1]), fminf(ylo[2], ylo[3]));
                wyhi = fmaxf(fmaxf(yhi[0], yhi[1]), fmaxf(yhi[2], yhi[3]));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    wxlo = fminf(wxlo, __shfl_xor_sync(0xffffffffu, wxlo, o));
                    wxhi = fmaxf(wxhi, __shfl_xor_sync(0xffffffffu, wxhi, o));
                    wylo = fminf(wylo, __shfl_xor_sync(0xffffffffu, wylo, o));
                    wyhi = fmaxf(wyhi, __shfl_xor_sync(0xffffffffu, wyhi, o));
                }
            }
            float4 xq, yq;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(xq.x), "=f"(xq.y), "=f"(xq.z), "=f"(xq.w)
                         : "r"(tile_addr + (uint32_t)c * 512u));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(yq.x), "=f"(yq.y), "=f"(yq.z), "=f"(yq.w)
                         : "r"(tile_addr + (uint32_t)(PP_NNS_TILE * 4) + (uint32_t)c * 512u));
            const float xv[4] = {xq.x, xq.y, xq.z, xq.w}, yv[4] = {yq.x, yq.y, yq.z, yq.w};
            // common case: no lane holds a candidate -> sixteen compares, one vote, one branch
            bool in[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                in[k] = !((xv[k] > wxhi) || (xv[k] < wxlo)) && !((yv[k] > wyhi) || (yv[k] < wylo));
            unsigned mask = __ballot_sync(0xffffffffu, in[0] || in[1] || in[2] || in[3]);
            if (mask == 0u) continue;
            const unsigned nib = (in[0] ? 1u : 0u) | (in[1] ? 2u : 0u) | (in[2] ? 4u : 0u) | (in[3] ? 8u : 0u);
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                const unsigned nj = __shfl_sync(0xffffffffu, nib, j);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!(nj & (1u << k))) continue;  // warp-uniform
                    const float xs = __shfl_sync(0xffffffffu, xv[k], j), ys = __shfl_sync(0xffffffffu, yv[k], j);
                    bool surv[PP_NNS_QPT];
                    bool any = false;
#pragma unroll
                    for (int q = 0; q < PP_NNS_QPT; ++q) {
                        surv[q] = !((xs > xhi[q]) || (xs < xlo[q])) && !((ys > yhi[q]) || (ys < ylo[q]));
                        any |= surv[q];
                    }
                    if (any) {
                        const uint32_t node = base + c * 128 + j * 4 + k;
                        const double nxv = __ldg(nx + node), nyv = __ldg(ny + node);
#pragma unroll
                        for (int q = 0; q < PP_NNS_QPT; ++q) {
                            if (surv[q]) {
                                const double dx = nxv - x[q], dy = nyv - y[q];
                                const double v = dx * dx + dy * dy;
                                if (v < best[q]) {
                                    best[q] = v;
                                    bi[q] = node;
                                    // |nx - qx| >= r or |ny - qy| >= r  =>  d2 >= best
                                    const double r = __dsqrt_ru(v);
                                    xhi[q] = __double2float_ru(__dadd_ru(x[q], r));
                                    xlo[q] = __double2float_rd(__dadd_rd(x[q], -r));
                                    yhi[q] = __double2float_ru(__dadd_ru(y[q], r));
                                    ylo[q] = __double2float_rd(__dadd_rd(y[q], -r));
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0 && t + PP_NNS_STAGES < n_tiles) {
            pp_mbar_expect_tx(&full[s], PP_NNS_TILE_BYTES);
            pp_bulk_g2s(tiles + (size_t)s * (2 * PP_NNS_TILE), nxy32 + (size_t)(t + PP_NNS_STAGES) * (2 * PP_NNS_TILE),
                        PP_NNS_TILE_BYTES, &full[s]);
        }
    }
#pragma unroll
    for (int q = 0; q < PP_NNS_QPT; ++q) {
        if (qi[q] != 0xFFFFFFFFu) {
            idx_out[qi[q]] = bi[q];
            if (d2_out) d2_out[qi[q]] = best[q];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int pp_nn_configure(pp_ctx *ctx) {
    PP_CUDA(ctx, cudaFuncSetAttribute(pp_nn_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)PP_NN_SMEM_BYTES));
    PP_CUDA(ctx, cudaFuncSetAttribute(pp_nn_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)PP_NN_SMEM_BYTES));
    PP_CUDA(ctx, cudaFuncSetAttribute(pp_nn_bucketed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)PP_NNS_SMEM_BYTES));
    return PP_OK;
}

size_t pp_nn_tile_nodes() { return 4096; }  // padding granule: a multiple of every tile size used here
size_t pp_nn_xy32_floats(size_t cap) { return 2 * cap; }  // cap is a multiple of the granule

// Counting sort of m points into the cells of a g x g grid over their bounding box (~64 points per cell):
// *perm_out (in the ctx scratch, valid until the next call that uses the scratch) lists the point indices
// cell by cell.  Shared by the bucketed NN scan and the bucketed straight-edge verify (collide.cu).
int pp_build_bucket_perm(pp_ctx *ctx, size_t m, const double *kx, const double *ky, uint32_t **perm_out,
                         cudaStream_t stream) {
    if (m >= 0xFFFFFFF0ull) return pp_fail(ctx, PP_ERR_INVALID, "too many items for one call");
    uint32_t g = (uint32_t)sqrt((double)m / 64.0);
    g = g < 1 ? 1 : (g > 128 ? 128 : g);
    const uint32_t nb = g * g;
    const size_t need = 64 + (size_t)nb * 8 + m * 4;
    int rc = pp_scratch_reserve(ctx, need);
    if (rc) return rc;
    unsigned long long *mm = (unsigned long long *)ctx->scratch;
    uint32_t *hist = (uint32_t *)((char *)ctx->scratch + 64);
    uint32_t *cursor = hist + nb;
    uint32_t *perm = cursor + nb;
    pp_launch_scope scope(ctx, "bucket_sort", 4);
    const unsigned long long init[4] = {~0ull, 0ull, ~0ull, 0ull};
    PP_CUDA(ctx, cudaMemcpyAsync(mm, init, sizeof init, cudaMemcpyHostToDevice, stream));
    PP_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)nb * 4, stream));
    const unsigned g1 = (unsigned)((m + 255) / 256);
    const unsigned gr = g1 < (unsigned)ctx->sm_count * 8 ? g1 : (unsigned)ctx->sm_count * 8;
    pp_nn_qrange_kernel<<<gr, 256, 0, stream>>>(kx, ky, m, mm);
    pp_nn_bucket_count_kernel<<<g1, 256, 0, stream>>>(kx, ky, m, mm, g, hist);
    pp_nn_bucket_scan_kernel<<<1, 1024, 0, stream>>>(hist, nb, cursor);
    pp_nn_bucket_scatter_kernel<<<g1, 256, 0, stream>>>(kx, ky, m, mm, g, cursor, perm);
    PP_CUDA(ctx, cudaGetLastError());
    *perm_out = perm;
    return PP_OK;
}

pp_nn_grid_view pp_nn_make_grid_view(const pp_tree_dev &t) {
    pp_nn_grid_view g;
    g.n_nodes = (uint32_t)t.n;
    g.grid_n = (t.grid_n == (size_t)-1 || t.grid_n > t.n) ? 0u : (uint32_t)t.grid_n;
    g.cell_start = t.cell_start;
    g.cell_items = t.cell_items;
    g.cell_xy = t.cell_xy;
    g.node_x = t.x;
    g.node_y = t.y;
    g.gx = t.gx;
    g.gy = t.gy;
    g.gminx = t.gminx;
    g.gminy = t.gminy;
    g.gcell = t.gcell;
    g.ginv = t.ginv;
    return g;
}

// Tail budget of the incremental node grid: nodes appended since the last build are scanned linearly by every
// query (pp_nn_grid_search), ~6 ns per tail node and thread.  A rebuild costs O(n) on the device (~0.25 ms at 2^20
// nodes).  Rebuild when the tail exceeds PP_NN_TAIL_MAX nodes (bounds a scalar query's latency; amortised cost of
// the rebuild: n / 4096 node visits per appended node) or when this call's m * tail pair evaluations outweigh a
// rebuild.  Returns true when the grid must be rebuilt before it is searched.
#define PP_NN_TAIL_MAX 4096  // ~12 warp instructions per tail node and warp: <= 25 us per call
bool pp_nn_grid_policy(const pp_tree_dev &t, size_t m) {
    if (t.grid_n == (size_t)-1 || t.grid_n > t.n || (t.grid_n == 0 && t.n != 0)) return true;
    const size_t tail = t.n - t.grid_n;
    if (tail == 0) return false;
    if (tail > PP_NN_TAIL_MAX) return true;
    return (double)m * (double)tail > 32.0 * (double)t.n + 65536.0;
}

int pp_launch_nn(pp_ctx *ctx, size_t m, const double *qx, const double *qy, uint32_t *idx, double *d2, int flags,
                 cudaStream_t stream) {
    if (m == 0) return PP_OK;
    const pp_tree_dev &t = ctx->tree;
    const uint32_t n_nodes = (uint32_t)t.n;
    if (flags & PP_NN_GRID) {
        pp_launch_scope scope(ctx, "nn_grid");
        pp_nn_grid_kernel<<<(unsigned)((m + 127) / 128), 128, 0, stream>>>(pp_nn_make_grid_view(t), qx, qy, m, idx, d2);
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
    } else if (flags & PP_NN_UNSORTED) {
        pp_launch_scope scope(ctx, "nn_scan_unsorted");
        pp_nn_scan_kernel<true><<<grid, PP_NN_THREADS, PP_NN_SMEM_BYTES, stream>>>(t.x, t.y, t.x32, n_tiles, qx, qy, m,
                                                                                    idx, d2);
    } else {
        // bin the queries into grid cells, then the bucketed scan
        uint32_t *perm = nullptr;
        int rc = pp_build_bucket_perm(ctx, m, qx, qy, &perm, stream);
        if (rc) return rc;
        const uint32_t n_tiles_s = (uint32_t)((t.n + PP_NNS_TILE - 1) / PP_NNS_TILE);
        const unsigned grid_s = (unsigned)((m + PP_NNS_THREADS * PP_NNS_QPT - 1) / (PP_NNS_THREADS * PP_NNS_QPT));
        pp_launch_scope scope(ctx, "nn_scan");
        pp_nn_bucketed_kernel<<<grid_s, PP_NNS_THREADS, PP_NNS_SMEM_BYTES, stream>>>(t.x, t.y, t.x32, n_tiles_s, qx, qy,
                                                                                       perm, m, idx, d2);
    }
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

// ---------------------------------------------------------------------------------------------
// Device-side build of the uniform node grid of pp_nn_grid_kernel (counting sort of the nodes by cell):
// range of the finite nodes -> 32 bytes to the host, which fixes the grid geometry -> histogram ->
// single-block scan -> scatter.  The order of the ids inside a cell is whatever the atomics produce; the
// search compares (d2, index) explicitly, so the answer does not depend on it.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    pp_nn_noderange_kernel(const double *__restrict__ x, const double *__restrict__ y, size_t n,
                           unsigned long long *__restrict__ mm) {
    unsigned long long lo[2] = {~0ull, ~0ull}, hi[2] = {0ull, 0ull};
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const double v[2] = {x[i], y[i]};
        if (isfinite(v[0]) && isfinite(v[1])) {  // a node with a non-finite coordinate can never be nearest
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const unsigned long long k = pp_f64_key(v[a]);
                lo[a] = min(lo[a], k);
                hi[a] = max(hi[a], k);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = min(lo[a], __shfl_down_sync(0xffffffffu, lo[a], o));
            hi[a] = max(hi[a], __shfl_down_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&mm[2 * a], lo[a]);
            atomicMax(&mm[2 * a + 1], hi[a]);
        }
    }
}

// the cell a node is filed under: the same expression the search uses for a query
__device__ __forceinline__ uint32_t pp_nn_grid_cell(double x, double y, double gminx, double gminy, double ginv, int gx,
                                                    int gy) {
    const double fx = floor((x - gminx) * ginv), fy = floor((y - gminy) * ginv);
    const int cx = (fx >= (double)gx) ? gx - 1 : ((fx > 0.0) ? (int)fx : 0);  // NaN -> 0
    const int cy = (fy >= (double)gy) ? gy - 1 : ((fy > 0.0) ? (int)fy : 0);
    return (uint32_t)cy * (uint32_t)gx + (uint32_t)cx;
}

// hist1 = cell_start + 1: after the inclusive scan cell_start[c] is the first slot of cell c
__global__ void __launch_bounds__(256)
    pp_nn_grid_count_kernel(const double *__restrict__ x, const double *__restrict__ y, uint32_t n, double gminx,
                            double gminy, double ginv, int gx, int gy, uint32_t *__restrict__ hist1) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) atomicAdd(&hist1[pp_nn_grid_cell(x[i], y[i], gminx, gminy, ginv, gx, gy)], 1u);
}

// In-place inclusive scan v[i] <- v[0] + ... + v[i] plus a copy of the exclusive prefix in cursor[] (the scatter's
// running slot per cell), in three small launches: every block scans its own 16 384 values (16 per thread) and
// records its total, one block scans the <= 1 024 totals, and the offsets are added back.  (A single block looping
// over all cells took 0.70 ms for the 5.2e5 cells of a 2^20-node tree: 32 rounds of strided accesses and barriers.)
#define PP_GRID_SCAN_VPT 16
#define PP_GRID_SCAN_BLOCK (1024 * PP_GRID_SCAN_VPT)
__global__ void __launch_bounds__(1024)
    pp_nn_grid_scan_local_kernel(uint32_t *__restrict__ v, uint32_t nb, uint32_t *__restrict__ cursor,
                                 uint32_t *__restrict__ block_sums) {
    __shared__ uint32_t warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t first = blockIdx.x * PP_GRID_SCAN_BLOCK + threadIdx.x * PP_GRID_SCAN_VPT;
    uint32_t a[PP_GRID_SCAN_VPT];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < PP_GRID_SCAN_VPT; ++k) {
        a[k] = (first + k < nb) ? v[first + k] : 0u;
        sum += a[k];
    }
    uint32_t inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = warp_sums[lane];
        uint32_t winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_sums[lane] = winc - w;
        if (lane == 31) block_sums[blockIdx.x] = winc;
    }
    __syncthreads();
    uint32_t run = warp_sums[warp] + inc - sum;  // exclusive prefix of this thread's first value inside the block
#pragma unroll
    for (int k = 0; k < PP_GRID_SCAN_VPT; ++k) {
        if (first + k < nb) {
            cursor[first + k] = run;
            run += a[k];
            v[first + k] = run;
        }
    }
}

// exclusive scan of the <= 1 024 block totals, in place
__global__ void __launch_bounds__(1024) pp_nn_grid_scan_sums_kernel(uint32_t *__restrict__ block_sums, uint32_t nblk) {
    __shared__ uint32_t warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t val = (threadIdx.x < nblk) ? block_sums[threadIdx.x] : 0u;
    uint32_t inc = val;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = warp_sums[lane];
        uint32_t winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_sums[lane] = winc - w;
    }
    __syncthreads();
    if (threadIdx.x < nblk) block_sums[threadIdx.x] = warp_sums[warp] + inc - val;
}

__global__ void __launch_bounds__(1024)
    pp_nn_grid_scan_add_kernel(uint32_t *__restrict__ v, uint32_t nb, uint32_t *__restrict__ cursor,
                               const uint32_t *__restrict__ block_offsets) {
    const uint32_t off = block_offsets[blockIdx.x];
    if (off == 0u) return;
    const uint32_t first = blockIdx.x * PP_GRID_SCAN_BLOCK + threadIdx.x * PP_GRID_SCAN_VPT;
#pragma unroll
    for (int k = 0; k < PP_GRID_SCAN_VPT; ++k) {
        if (first + k < nb) {
            v[first + k] += off;
            cursor[first + k] += off;
        }
    }
}

__global__ void __launch_bounds__(256)
    pp_nn_grid_scatter_kernel(const double *__restrict__ x, const double *__restrict__ y, uint32_t n, double gminx,
                              double gminy, double ginv, int gx, int gy, uint32_t *__restrict__ cursor,
                              uint32_t *__restrict__ items, double2 *__restrict__ items_xy) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        const double xi = x[i], yi = y[i];
        const uint32_t pos = atomicAdd(&cursor[pp_nn_grid_cell(xi, yi, gminx, gminy, ginv, gx, gy)], 1u);
        items[pos] = i;
        items_xy[pos] = make_double2(xi, yi);  // the search reads coordinates cell by cell, ids only on improvement
    }
}

#ifndef PP_NN_GRID_OCC_DEFAULT
#define PP_NN_GRID_OCC_DEFAULT 2.0
#endif
int pp_tree_build_grid(pp_ctx *ctx, cudaStream_t stream) {
    pp_tree_dev &t = ctx->tree;
    if (t.grid_n == t.n) return PP_OK;
    ctx->grid_builds += 1;
    const size_t n = t.n;
    // geometry: ~2 nodes per cell, square cells over the bounding box of the finite nodes
    // (PP_NN_GRID_OCC: developer knob for the nodes-per-cell target, read at build time)
    double occ = PP_NN_GRID_OCC_DEFAULT;
    if (const char *e = getenv("PP_NN_GRID_OCC")) occ = atof(e) > 0.0 ? atof(e) : occ;
    long g = (long)floor(sqrt((double)(n > 1 ? n : 1) / occ));
    g = g < 1 ? 1 : (g > 4095 ? 4095 : g);  // (g + 1)^2 cells <= 1 024 scan blocks
    const size_t max_cells = (size_t)(g + 1) * (size_t)(g + 1);
    int rc = pp_scratch_reserve(ctx, 64 + 4096 + max_cells * 4);
    if (rc) return rc;
    unsigned long long *mm = (unsigned long long *)ctx->scratch;
    uint32_t *block_sums = (uint32_t *)((char *)ctx->scratch + 64);  // <= 1 024 scan blocks (g <= 4096)
    uint32_t *cursor = block_sums + 1024;
    if (max_cells + 1 > t.cell_cap) {
        PP_CUDA(ctx, cudaStreamSynchronize(stream));
        cudaFree(t.cell_start);
        t.cell_start = nullptr;
        t.cell_cap = 0;
        const size_t cap = 2 * max_cells + 1;
        if (cudaMalloc(&t.cell_start, cap * 4) != cudaSuccess) {
            cudaGetLastError();
            return pp_fail(ctx, PP_ERR_NOMEM, "nn grid allocation failed");
        }
        t.cell_cap = cap;
    }
    if ((n > 1 ? n : 1) > t.item_cap) {
        PP_CUDA(ctx, cudaStreamSynchronize(stream));
        cudaFree(t.cell_items);
        cudaFree(t.cell_xy);
        t.cell_items = nullptr;
        t.cell_xy = nullptr;
        t.item_cap = 0;
        const size_t cap = t.cap > n ? t.cap : (n > 1 ? n : 1);  // grows with the node arrays
        if (cudaMalloc(&t.cell_items, cap * 4) != cudaSuccess || cudaMalloc(&t.cell_xy, cap * sizeof(double2)) != cudaSuccess) {
            cudaGetLastError();
            return pp_fail(ctx, PP_ERR_NOMEM, "nn grid allocation failed");
        }
        t.item_cap = cap;
    }
    pp_launch_scope scope(ctx, "nn_grid_build", 6);
    const unsigned long long init[4] = {~0ull, 0ull, ~0ull, 0ull};
    unsigned long long got[4] = {~0ull, 0ull, ~0ull, 0ull};
    if (n) {
        PP_CUDA(ctx, cudaMemcpyAsync(mm, init, sizeof init, cudaMemcpyHostToDevice, stream));
        const unsigned g1 = (unsigned)((n + 255) / 256);
        const unsigned gr = g1 < (unsigned)ctx->sm_count * 8 ? g1 : (unsigned)ctx->sm_count * 8;
        pp_nn_noderange_kernel<<<gr, 256, 0, stream>>>(t.x, t.y, n, mm);
        PP_CUDA(ctx, cudaMemcpyAsync(got, mm, sizeof got, cudaMemcpyDeviceToHost, stream));
        PP_CUDA(ctx, cudaStreamSynchronize(stream));
    }
    auto unkey = [](unsigned long long k) {
        const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
        double d;
        memcpy(&d, &b, 8);
        return d;
    };
    double minx = 0.0, maxx = 0.0, miny = 0.0, maxy = 0.0;
    if (got[0] <= got[1]) {  // at least one finite node
        minx = unkey(got[0]);
        maxx = unkey(got[1]);
        miny = unkey(got[2]);
        maxy = unkey(got[3]);
    }
    const double w = maxx - minx, h = maxy - miny;
    double cell = (w > h ? w : h) / (double)g;
    if (!(cell > 0.0) || !isfinite(cell)) cell = 1.0;
    const double inv = 1.0 / cell;
    const double fgx = floor(w * inv) + 1.0, fgy = floor(h * inv) + 1.0;
    const int gx = (int)(fgx < (double)(g + 1) ? fgx : (double)(g + 1));
    const int gy = (int)(fgy < (double)(g + 1) ? fgy : (double)(g + 1));
    const uint32_t ncell = (uint32_t)gx * (uint32_t)gy;
    PP_CUDA(ctx, cudaMemsetAsync(t.cell_start, 0, ((size_t)ncell + 1) * 4, stream));
    if (n) {
        const unsigned g1 = (unsigned)((n + 255) / 256);
        pp_nn_grid_count_kernel<<<g1, 256, 0, stream>>>(t.x, t.y, (uint32_t)n, minx, miny, inv, gx, gy, t.cell_start + 1);
        const unsigned nblk = (ncell + PP_GRID_SCAN_BLOCK - 1) / PP_GRID_SCAN_BLOCK;  // <= 1 024 + 1
        pp_nn_grid_scan_local_kernel<<<nblk, 1024, 0, stream>>>(t.cell_start + 1, ncell, cursor, block_sums);
        pp_nn_grid_scan_sums_kernel<<<1, 1024, 0, stream>>>(block_sums, nblk);
        pp_nn_grid_scan_add_kernel<<<nblk, 1024, 0, stream>>>(t.cell_start + 1, ncell, cursor, block_sums);
        pp_nn_grid_scatter_kernel<<<g1, 256, 0, stream>>>(t.x, t.y, (uint32_t)n, minx, miny, inv, gx, gy, cursor,
                                                          t.cell_items, t.cell_xy);
    }
    PP_CUDA(ctx, cudaGetLastError());
    t.gx = gx;
    t.gy = gy;
    t.gminx = minx;
    t.gminy = miny;
    t.gcell = cell;
    t.ginv = inv;
    t.grid_n = n;
    return PP_OK;
}

// fl32 copies of x and y in the blocked layout (+inf sentinel padding up to the granule)
__global__ void pp_tree_x32_kernel(const double *__restrict__ x, const double *__restrict__ y, float *__restrict__ xy32,
                                   size_t first, size_t n) {
    size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const size_t k = pp_xy32_index(i);
        xy32[k] = __double2float_rn(x[i]);
        xy32[k + PP_XY32_BLOCK] = __double2float_rn(y[i]);
    }
}
__global__ void pp_tree_pad_kernel(double *x, double *y, float *xy32, size_t first, size_t end) {
    size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < end) {
        x[i] = CUDART_INF;
        y[i] = CUDART_INF;
        const size_t k = pp_xy32_index(i);
        xy32[k] = CUDART_INF_F;
        xy32[k + PP_XY32_BLOCK] = CUDART_INF_F;
    }
}

int pp_launch_tree_finish(pp_ctx *ctx, size_t first, size_t n, size_t padded_end, cudaStream_t stream) {
    pp_tree_dev &t = ctx->tree;
    pp_launch_scope scope(ctx, "tree_finish", 2);
    if (n > first)
        pp_tree_x32_kernel<<<(unsigned)((n - first + 255) / 256), 256, 0, stream>>>(t.x, t.y, t.x32, first, n);
    if (padded_end > n)
        pp_tree_pad_kernel<<<(unsigned)((padded_end - n + 255) / 256), 256, 0, stream>>>(t.x, t.y, t.x32, n, padded_end);
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

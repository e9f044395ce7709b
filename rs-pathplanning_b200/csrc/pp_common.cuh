// pp_common.cuh -- context, error plumbing and sm_100a async-copy primitives shared by all kernels.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pathplanning_b200.h"

#define PP_SM_COUNT_B200 148

// ---------------------------------------------------------------------------------------------
// device mirrors
// ---------------------------------------------------------------------------------------------
struct pp_tree_dev {
    size_t n = 0, cap = 0;
    double *x = nullptr, *y = nullptr, *yaw = nullptr;
    int32_t *parent = nullptr;
    float *x32 = nullptr;  // fl32(x), fl32(y) in 2048-node blocks (nn.cu: pp_xy32_index) for the exact fp32 pre-rejection
    // uniform grid over the nodes [0, grid_n) (PP_NN_GRID); nodes [grid_n, n) appended since are the linearly
    // scanned tail; rebuilt when the tail outgrows its budget (nn.cu: pp_nn_grid_policy).  (size_t)-1: no grid
    size_t grid_n = (size_t)-1;
    int gx = 0, gy = 0;
    double gminx = 0, gminy = 0, gcell = 1, ginv = 1;
    uint32_t *cell_start = nullptr;  // gx*gy+1
    uint32_t *cell_items = nullptr;  // n node ids, cell by cell (any order inside a cell)
    double2 *cell_xy = nullptr;      // (x, y) of cell_items[k]: a row of cells is one contiguous run of coordinates
    size_t cell_cap = 0, item_cap = 0;
};

struct pp_ring_circle;  // geo_predicates.cuh

struct pp_ring_meta {  // per obstacle ring
    double minx, miny, maxx, maxy;  // exact AABB of the ring's points
    double pad;                     // conservative rounding pad (see collide.cu)
    uint32_t first, count;          // points [first, first+count) in ox/oy (closed ring)
    uint32_t _r0, _r1;
};

struct pp_world_dev {
    bool valid = false;
    // bounds ring
    double *bx = nullptr, *by = nullptr;
    uint32_t nb = 0;
    // obstacle rings
    double *ox = nullptr, *oy = nullptr;
    uint32_t n_pts = 0, n_rings = 0;
    pp_ring_meta *meta = nullptr;
    float4 *aabb32 = nullptr;  // outward-rounded padded AABBs (minx, miny, maxx, maxy) for the fp32 broad phase
    struct pp_ring_circle *circ = nullptr;  // per ring: centre, inflated outer / deflated inner radius^2 (geo_predicates.cuh)
    uint32_t n_aabb_tiles = 0;
    // byte grid classifying cells of the bounds' AABB: 0 outside, 1 inside, 2 needs the exact test
    uint8_t *bcls = nullptr;
    int bgx = 1, bgy = 1;
    double bminx = 0, bminy = 0, binvx = 1, binvy = 1;
    // uniform grid over padded ring AABBs
    int gx = 0, gy = 0;
    double gminx = 0, gminy = 0, gcell = 1, ginv = 1;
    uint32_t *cell_start = nullptr, *cell_items = nullptr;
    float4 *cell_box = nullptr;  // aabb32[cell_items[k]] at position k: the walk tests a box without knowing the ring id
    uint32_t n_cell_items = 0;
};

struct pp_timing_slot {
    double total_ms = 0;
    uint64_t launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

#define PP_PATH_BOX_CELLS 4  // obstacle-grid cells per axis a path box may span for the path-level test (path_box.cuh): 16 costs the C5 slice 9 % (a third of its edges then compute a box that never passes)
#define PP_TICKETS 128
#define PP_TICKETS_FILL 64

struct pp_ctx {
    int device = 0;
    int sm_count = PP_SM_COUNT_B200;
    cudaStream_t stream = nullptr;             // stream of the _dev entry points (may be caller-owned)
    cudaStream_t own_stream_handle = nullptr;  // the stream created with the ctx
    cudaStream_t active_stream = nullptr;      // stream the current API call launches on (set under `mu`)
    cudaStream_t copy_streams[3] = {nullptr, nullptr, nullptr};
    std::mutex mu;
    std::string last_error;
    uint64_t launches = 0;
    uint64_t grid_builds = 0;  // O(n) rebuilds of the node grid so far (pp_nn_grid_builds)
    int fill_resident = 0;     // resident CTAs per SM of the sample fill kernel (0 = not asked yet)
    bool timing = false;
    std::map<std::string, pp_timing_slot> timings;
    std::vector<cudaEvent_t> event_pool;
    pp_tree_dev tree;
    pp_world_dev world;
    // scratch (grown on demand)
    unsigned int *tickets = nullptr;  // PP_TICKETS zeroed counters: [0, 64) last-block-done reductions (nn.cu),
                                      // [PP_TICKETS_FILL, +2) work / done counters of the sample fill kernel
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
    // pinned ring + copy threads of the batch entry points when the caller's arrays are pageable (pp_stage.hpp)
    void *stage = nullptr;
    size_t stage_bytes = 0;
    class pp_stage_pool *stage_pool = nullptr;
    // NCCL communicator this ctx is a rank of (group.cu); null = single device
    void *comm = nullptr;
    int comm_rank = -1, comm_size = 1;
};

// every API entry: serialise on the ctx, select its device, pick the stream
struct pp_guard {
    std::lock_guard<std::mutex> lk;
    explicit pp_guard(pp_ctx *ctx) : lk(ctx->mu) {
        cudaSetDevice(ctx->device);
        ctx->active_stream = ctx->stream;
    }
};
void pp_comm_release(pp_ctx *ctx);                // group.cu: destroys ctx->comm if any
void pp_world_free(pp_world_dev &w);              // api.cu

int pp_fail(pp_ctx *ctx, int status, const char *what, cudaError_t e = cudaSuccess);

#define PP_CUDA(ctx, call)                                                    \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess) return pp_fail((ctx), PP_ERR_CUDA, #call, _e); \
    } while (0)

// RAII kernel-launch accounting (+ optional CUDA-event timing on the ctx stream)
struct pp_launch_scope {
    pp_ctx *ctx;
    const char *name;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    pp_launch_scope(pp_ctx *c, const char *n, int n_launches = 1);
    ~pp_launch_scope();
};

int pp_scratch_reserve(pp_ctx *ctx, size_t bytes);
int pp_tree_reserve(pp_ctx *ctx, size_t n_total, bool keep);
int pp_tree_copy_in(pp_ctx *ctx, size_t first, size_t k, const double *x, const double *y, const double *yaw,
                    const int32_t *parent, cudaMemcpyKind kind);
int pp_tree_commit(pp_ctx *ctx, size_t first, size_t k, bool sync);

// kernel-side view of pp_world_dev (passed by value)
struct pp_world_view {
    const double *bx, *by;
    uint32_t nb;
    const uint8_t *bcls;
    int bgx, bgy;
    double bminx, bminy, binvx, binvy;
    const double *ox, *oy;
    const pp_ring_meta *meta;
    uint32_t n_rings;
    const float4 *aabb32;
    const struct pp_ring_circle *circ;
    uint32_t n_aabb_tiles;
    const uint32_t *cell_start, *cell_items;
    const float4 *cell_box;
    int gx, gy;
    double gminx, gminy, ginv;
};


// ---------------------------------------------------------------------------------------------
// sm_100a primitives: mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP)
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t pp_smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void pp_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pp_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void pp_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void pp_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void pp_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pp_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void pp_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(pp_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned; completion on `bar`
__device__ __forceinline__ void pp_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     pp_smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(pp_smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (TMA engine; completion through bulk async-groups); bytes % 16 == 0, both addresses
// 16-byte aligned.  The data must have been made visible to the async proxy (pp_fence_proxy_async) by its writers.
__device__ __forceinline__ void pp_bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(pp_smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void pp_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source (the staging row is free again)
template <int N>
__device__ __forceinline__ void pp_bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void pp_bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
#endif

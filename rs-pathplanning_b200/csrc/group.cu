// group.cu -- multi-GPU: replication of tree / obstacles by ncclBroadcast over NVLink / NVSwitch and contiguous
// slicing of batches across devices (include/pathplanning_b200.h "multi-GPU").
//
// The path has no exchange step -- every pose pair, NN query and edge is independent (src/rrt.rs:607-609) -- so the
// only collective is the broadcast of the replicated buffers when they change: the whole tree at RRT::new
// (src/rrt.rs:345-346), the appended tail at the insert site (src/rrt.rs:586-589), the obstacle buffers after Space::new.
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy a host such as PyTorch already loaded, else the
// system's), so the library has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>  // types and enumerators only; every function is looked up with dlsym

#include <condition_variable>
#include <cstring>
#include <functional>
#include <thread>

#include "geo_predicates.cuh"
#include "pp_common.cuh"

// ------------------------------------------------------------------------------------------------ NCCL binding
namespace {
struct nccl_api {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

nccl_api *nccl() {
    static nccl_api api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        auto sym = [&](const char *n) { return dlsym(api.handle, n); };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.Broadcast &&
                 api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    return api.ok ? &api : nullptr;
}

int comm_fail(pp_ctx *ctx, const char *what, ncclResult_t r) {
    nccl_api *a = nccl();
    std::string msg = std::string(what) + ": " + (a ? a->GetErrorString(r) : "NCCL not loaded");
    return pp_fail(ctx, PP_ERR_COMM, msg.c_str());
}
#define PP_NCCL(ctx, call)                                   \
    do {                                                     \
        ncclResult_t _r = (call);                            \
        if (_r != ncclSuccess) return comm_fail(ctx, #call, _r); \
    } while (0)
}  // namespace

void pp_slice_bounds(size_t n, int parts, int part, size_t *lo, size_t *hi) {
    if (parts < 1) parts = 1;
    if (part < 0) part = 0;
    if (part >= parts) part = parts - 1;
    // g*n/G without overflow for n up to 2^57 and G <= 64
    const unsigned __int128 N = n;
    if (lo) *lo = (size_t)(N * (unsigned)part / (unsigned)parts);
    if (hi) *hi = (size_t)(N * (unsigned)(part + 1) / (unsigned)parts);
}

void pp_comm_release(pp_ctx *ctx) {
    if (ctx && ctx->comm) {
        if (nccl_api *a = nccl()) a->CommDestroy((ncclComm_t)ctx->comm);
        ctx->comm = nullptr;
        ctx->comm_rank = -1;
        ctx->comm_size = 1;
    }
}

int pp_comm_unique_id(void *id) {
    if (!id) return PP_ERR_INVALID;
    nccl_api *a = nccl();
    if (!a) return PP_ERR_COMM;
    static_assert(sizeof(ncclUniqueId) == PP_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    if (a->GetUniqueId(&u) != ncclSuccess) return PP_ERR_COMM;
    memcpy(id, &u, sizeof u);
    return PP_OK;
}

int pp_ctx_comm_init(pp_ctx *ctx, const void *id, int n_ranks, int rank) {
    if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return PP_ERR_INVALID;
    nccl_api *a = nccl();
    if (!a) return pp_fail(ctx, PP_ERR_COMM, "libnccl.so.2 could not be loaded");
    pp_guard g(ctx);
    pp_comm_release(ctx);
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t c = nullptr;
    PP_NCCL(ctx, a->CommInitRank(&c, n_ranks, u, rank));
    ctx->comm = c;
    ctx->comm_rank = rank;
    ctx->comm_size = n_ranks;
    return PP_OK;
}
int pp_ctx_comm_rank(pp_ctx *ctx) { return ctx ? ctx->comm_rank : -1; }
int pp_ctx_comm_size(pp_ctx *ctx) { return ctx ? ctx->comm_size : 1; }

// ------------------------------------------------------------------------------------------------ tree replication
// slots [first, first + k): host arrays -> root's device tree, one fused NCCL launch broadcasts the four SoA arrays in
// place over NVLink, every rank then runs the same finishing kernels (fp32 copies, sentinel padding)
static int tree_bcast(pp_ctx *ctx, int root, size_t first, size_t k, const double *x, const double *y, const double *yaw,
                      const int32_t *parent) {
    const bool multi = ctx->comm && ctx->comm_size > 1;
    const bool is_root = !multi || ctx->comm_rank == root;
    if (multi && (root < 0 || root >= ctx->comm_size)) return pp_fail(ctx, PP_ERR_INVALID, "broadcast root out of range");
    if (is_root && k && (!x || !y)) return pp_fail(ctx, PP_ERR_INVALID, "root rank needs the node arrays");
    if (first + k >= 0xFFFFFFF0ull) return PP_ERR_INVALID;
    int rc = pp_tree_reserve(ctx, first + k, first != 0);
    if (rc) return rc;
    if (is_root) {
        rc = pp_tree_copy_in(ctx, first, k, x, y, yaw, parent, cudaMemcpyHostToDevice);
        if (rc) return rc;
    }
    if (multi && k) {
        nccl_api *a = nccl();
        if (!a) return pp_fail(ctx, PP_ERR_COMM, "libnccl.so.2 could not be loaded");
        pp_tree_dev &t = ctx->tree;
        ncclComm_t c = (ncclComm_t)ctx->comm;
        cudaStream_t s = ctx->stream;
        PP_NCCL(ctx, a->GroupStart());
        ncclResult_t r0 = a->Broadcast(t.x + first, t.x + first, k, ncclDouble, root, c, s);
        ncclResult_t r1 = a->Broadcast(t.y + first, t.y + first, k, ncclDouble, root, c, s);
        ncclResult_t r2 = a->Broadcast(t.yaw + first, t.yaw + first, k, ncclDouble, root, c, s);
        ncclResult_t r3 = a->Broadcast(t.parent + first, t.parent + first, k, ncclInt32, root, c, s);
        ncclResult_t re = a->GroupEnd();
        for (ncclResult_t r : {r0, r1, r2, r3, re})
            if (r != ncclSuccess) return comm_fail(ctx, "ncclBroadcast(tree)", r);
        ctx->launches += 1;  // NCCL fuses the grouped broadcasts into one kernel
    }
    return pp_tree_commit(ctx, first, k, true);
}

int pp_tree_upload_bcast(pp_ctx *ctx, int root, size_t n, const double *x, const double *y, const double *yaw,
                         const int32_t *parent) {
    if (!ctx) return PP_ERR_INVALID;
    pp_guard g(ctx);
    return tree_bcast(ctx, root, 0, n, x, y, yaw, parent);
}
int pp_tree_append_bcast(pp_ctx *ctx, int root, size_t k, const double *x, const double *y, const double *yaw,
                         const int32_t *parent) {
    if (!ctx) return PP_ERR_INVALID;
    pp_guard g(ctx);
    return tree_bcast(ctx, root, ctx->tree.n, k, x, y, yaw, parent);
}

// ------------------------------------------------------------------------------------------------ world replication
namespace {
struct world_header {  // everything of pp_world_dev that is not a device array
    int32_t status;
    uint32_t nb, n_pts, n_rings, n_aabb_tiles, n_cell_items;
    int32_t bgx, bgy, gx, gy;
    double d[8];  // bminx, bminy, binvx, binvy, gminx, gminy, gcell, ginv
};
template <class T>
int alloc_dev(pp_ctx *ctx, T **p, size_t count) {
    *p = nullptr;
    if (cudaMalloc((void **)p, (count ? count : 1) * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        return pp_fail(ctx, PP_ERR_NOMEM, "obstacle allocation failed");
    }
    return PP_OK;
}
}  // namespace

int pp_obstacles_upload_bcast(pp_ctx *ctx, int root, const double *bounds_x, const double *bounds_y, size_t n_bounds,
                              const double *ring_x, const double *ring_y, const uint32_t *ring_off, size_t n_rings) {
    if (!ctx) return PP_ERR_INVALID;
    const bool multi = ctx->comm && ctx->comm_size > 1;
    if (!multi) return pp_obstacles_upload(ctx, bounds_x, bounds_y, n_bounds, ring_x, ring_y, ring_off, n_rings);
    if (root < 0 || root >= ctx->comm_size) return pp_fail(ctx, PP_ERR_INVALID, "broadcast root out of range");
    nccl_api *a = nccl();
    if (!a) return pp_fail(ctx, PP_ERR_COMM, "libnccl.so.2 could not be loaded");
    const bool is_root = ctx->comm_rank == root;
    int up_rc = PP_OK;
    if (is_root)  // host-side preprocessing (grids, padded boxes) happens once, on the root
        up_rc = pp_obstacles_upload(ctx, bounds_x, bounds_y, n_bounds, ring_x, ring_y, ring_off, n_rings);
    pp_guard g(ctx);
    cudaStream_t s = ctx->stream;
    ncclComm_t c = (ncclComm_t)ctx->comm;
    pp_world_dev &w = ctx->world;
    world_header h;
    memset(&h, 0, sizeof h);
    if (is_root) {
        h.status = up_rc;
        if (up_rc == PP_OK) {
            h.nb = w.nb; h.n_pts = w.n_pts; h.n_rings = w.n_rings; h.n_aabb_tiles = w.n_aabb_tiles;
            h.n_cell_items = w.n_cell_items;
            h.bgx = w.bgx; h.bgy = w.bgy; h.gx = w.gx; h.gy = w.gy;
            const double d[8] = {w.bminx, w.bminy, w.binvx, w.binvy, w.gminx, w.gminy, w.gcell, w.ginv};
            memcpy(h.d, d, sizeof d);
        }
    }
    int rc = pp_scratch_reserve(ctx, sizeof h);
    if (rc) return rc;
    if (is_root) PP_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, &h, sizeof h, cudaMemcpyHostToDevice, s));
    PP_NCCL(ctx, a->Broadcast(ctx->scratch, ctx->scratch, sizeof h, ncclUint8, root, c, s));
    ctx->launches += 1;
    PP_CUDA(ctx, cudaMemcpyAsync(&h, ctx->scratch, sizeof h, cudaMemcpyDeviceToHost, s));
    PP_CUDA(ctx, cudaStreamSynchronize(s));
    if (h.status != PP_OK) return pp_fail(ctx, h.status, "obstacle upload failed on the root rank");
    const size_t n_bcls = (size_t)h.bgx * (size_t)h.bgy, n_cells = (size_t)h.gx * (size_t)h.gy + 1;
    const size_t n_aabb = (size_t)h.n_aabb_tiles * 1024;
    if (!is_root) {
        pp_world_free(w);
        if ((rc = alloc_dev(ctx, &w.bx, h.nb)) || (rc = alloc_dev(ctx, &w.by, h.nb)) ||
            (rc = alloc_dev(ctx, &w.bcls, n_bcls)) || (rc = alloc_dev(ctx, &w.ox, h.n_pts)) ||
            (rc = alloc_dev(ctx, &w.oy, h.n_pts)) || (rc = alloc_dev(ctx, &w.meta, h.n_rings)) ||
            (rc = alloc_dev(ctx, &w.aabb32, n_aabb)) || (rc = alloc_dev(ctx, &w.circ, h.n_rings)) ||
            (rc = alloc_dev(ctx, &w.cell_start, n_cells)) ||
            (rc = alloc_dev(ctx, &w.cell_items, h.n_cell_items)) || (rc = alloc_dev(ctx, &w.cell_box, h.n_cell_items)))
            return rc;
        w.nb = h.nb; w.n_pts = h.n_pts; w.n_rings = h.n_rings; w.n_aabb_tiles = h.n_aabb_tiles;
        w.n_cell_items = h.n_cell_items;
        w.bgx = h.bgx; w.bgy = h.bgy; w.gx = h.gx; w.gy = h.gy;
        w.bminx = h.d[0]; w.bminy = h.d[1]; w.binvx = h.d[2]; w.binvy = h.d[3];
        w.gminx = h.d[4]; w.gminy = h.d[5]; w.gcell = h.d[6]; w.ginv = h.d[7];
    }
    struct item { void *p; size_t bytes; };
    const item items[11] = {{w.circ, (size_t)h.n_rings * sizeof(pp_ring_circle)},
                           {w.cell_box, (size_t)h.n_cell_items * sizeof(float4)},
                           {w.bx, (size_t)h.nb * 8}, {w.by, (size_t)h.nb * 8}, {w.bcls, n_bcls},
                           {w.ox, (size_t)h.n_pts * 8}, {w.oy, (size_t)h.n_pts * 8},
                           {w.meta, (size_t)h.n_rings * sizeof(pp_ring_meta)}, {w.aabb32, n_aabb * sizeof(float4)},
                           {w.cell_start, n_cells * 4}, {w.cell_items, (size_t)h.n_cell_items * 4}};
    PP_NCCL(ctx, a->GroupStart());
    ncclResult_t bad = ncclSuccess;
    for (const item &it : items)
        if (it.bytes) {
            ncclResult_t r = a->Broadcast(it.p, it.p, it.bytes, ncclUint8, root, c, s);
            if (r != ncclSuccess) bad = r;
        }
    ncclResult_t re = a->GroupEnd();
    if (bad != ncclSuccess || re != ncclSuccess) return comm_fail(ctx, "ncclBroadcast(world)", bad != ncclSuccess ? bad : re);
    ctx->launches += 1;
    PP_CUDA(ctx, cudaStreamSynchronize(s));
    w.valid = true;
    return PP_OK;
}

// ------------------------------------------------------------------------------------------------ one process, N devices
struct pp_group {
    std::vector<pp_ctx *> ctx;
    std::vector<std::thread> workers;  // worker i owns device i: the sliced calls and the collectives run on it
    std::mutex mu, call_mu;
    std::condition_variable cv_job, cv_done;
    std::function<int(int)> job;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;
    std::vector<int> rc;
    std::string last_error;

    void worker(int i) {
        uint64_t seen = 0;
        for (;;) {
            std::function<int(int)> f;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_job.wait(lk, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
                f = job;
            }
            const int r = f(i);
            {
                std::lock_guard<std::mutex> lk(mu);
                rc[i] = r;
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    // runs f(i) on every device's worker at once; first failing status wins
    int run(std::function<int(int)> f) {
        std::lock_guard<std::mutex> call(call_mu);
        {
            std::lock_guard<std::mutex> lk(mu);
            job = std::move(f);
            pending = (int)ctx.size();
            ++generation;
        }
        cv_job.notify_all();
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return pending == 0; });
        }
        for (size_t i = 0; i < ctx.size(); ++i)
            if (rc[i] != PP_OK) {
                last_error = "device " + std::to_string(ctx[i]->device) + ": " + ctx[i]->last_error;
                return rc[i];
            }
        return PP_OK;
    }
};

int pp_group_create(const int *devices, int n_dev, pp_group **out) {
    if (!out || !devices || n_dev < 1 || n_dev > 64) return PP_ERR_INVALID;
    *out = nullptr;
    for (int i = 0; i < n_dev; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return PP_ERR_INVALID;
    pp_group *g = new pp_group();
    for (int i = 0; i < n_dev; ++i) {
        pp_ctx *c = nullptr;
        int rc = pp_ctx_create(devices[i], &c);
        if (rc != PP_OK) {
            pp_group_destroy(g);
            return rc;
        }
        g->ctx.push_back(c);
    }
    if (n_dev > 1) {
        nccl_api *a = nccl();
        std::vector<ncclComm_t> comms(n_dev, nullptr);
        if (!a || a->CommInitAll(comms.data(), n_dev, devices) != ncclSuccess) {
            pp_group_destroy(g);
            return PP_ERR_COMM;
        }
        for (int i = 0; i < n_dev; ++i) {
            g->ctx[i]->comm = comms[i];
            g->ctx[i]->comm_rank = i;
            g->ctx[i]->comm_size = n_dev;
        }
    }
    g->rc.assign(n_dev, PP_OK);
    for (int i = 0; i < n_dev; ++i) g->workers.emplace_back([g, i] { g->worker(i); });
    *out = g;
    return PP_OK;
}

void pp_group_destroy(pp_group *g) {
    if (!g) return;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->stop = true;
    }
    g->cv_job.notify_all();
    for (std::thread &t : g->workers)
        if (t.joinable()) t.join();
    for (pp_ctx *c : g->ctx) pp_ctx_destroy(c);
    delete g;
}

int pp_group_size(pp_group *g) { return g ? (int)g->ctx.size() : 0; }
pp_ctx *pp_group_ctx(pp_group *g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }
const char *pp_group_last_error(pp_group *g) { return g ? g->last_error.c_str() : "null group"; }

int pp_group_tree_upload(pp_group *g, size_t n, const double *x, const double *y, const double *yaw,
                         const int32_t *parent) {
    if (!g || (n && (!x || !y))) return PP_ERR_INVALID;
    return g->run([=](int i) { return pp_tree_upload_bcast(g->ctx[i], 0, n, x, y, yaw, parent); });
}
int pp_group_tree_append(pp_group *g, size_t k, const double *x, const double *y, const double *yaw,
                         const int32_t *parent) {
    if (!g || (k && (!x || !y))) return PP_ERR_INVALID;
    return g->run([=](int i) { return pp_tree_append_bcast(g->ctx[i], 0, k, x, y, yaw, parent); });
}
int pp_group_obstacles_upload(pp_group *g, const double *bounds_x, const double *bounds_y, size_t n_bounds,
                              const double *ring_x, const double *ring_y, const uint32_t *ring_off, size_t n_rings) {
    if (!g) return PP_ERR_INVALID;
    return g->run([=](int i) {
        return pp_obstacles_upload_bcast(g->ctx[i], 0, bounds_x, bounds_y, n_bounds, ring_x, ring_y, ring_off, n_rings);
    });
}

#define PP_SLICE(n)                                  \
    size_t lo, hi;                                   \
    pp_slice_bounds((n), (int)g->ctx.size(), i, &lo, &hi); \
    const size_t cnt = hi - lo;                      \
    pp_ctx *c = g->ctx[i];                           \
    (void)c;                                         \
    if (cnt == 0) return (int)PP_OK

int pp_group_dubins_eval(pp_group *g, size_t n, const double *sx, const double *sy, const double *syaw,
                         const double *ex, const double *ey, const double *eyaw, const double *radius_arr,
                         double radius, double *cost, uint8_t *word, double *tpq) {
    if (!g || (n && (!sx || !sy || !syaw || !ex || !ey || !eyaw || !cost || !word))) return PP_ERR_INVALID;
    return g->run([=](int i) {
        PP_SLICE(n);
        return pp_dubins_eval(c, cnt, sx + lo, sy + lo, syaw + lo, ex + lo, ey + lo, eyaw + lo,
                              radius_arr ? radius_arr + lo : nullptr, radius, cost + lo, word + lo,
                              tpq ? tpq + 3 * lo : nullptr);
    });
}
int pp_group_nn(pp_group *g, size_t m, const double *qx, const double *qy, uint32_t *idx, double *d2, int flags) {
    if (!g || (m && (!qx || !qy || !idx))) return PP_ERR_INVALID;
    return g->run([=](int i) {
        PP_SLICE(m);
        return pp_nn(c, cnt, qx + lo, qy + lo, idx + lo, d2 ? d2 + lo : nullptr, flags);
    });
}
int pp_group_collide_segments(pp_group *g, size_t m, const double *ax, const double *ay, const double *bx,
                              const double *by, uint8_t *ok, int flags) {
    if (!g || (m && (!ax || !ay || !bx || !by || !ok))) return PP_ERR_INVALID;
    return g->run([=](int i) {
        PP_SLICE(m);
        return pp_collide_segments(c, cnt, ax + lo, ay + lo, bx + lo, by + lo, ok + lo, flags);
    });
}
int pp_group_collide_dubins(pp_group *g, size_t m, const double *sx, const double *sy, const double *syaw,
                            const double *ex, const double *ey, const double *eyaw, double radius, double step,
                            uint8_t *ok, int flags) {
    if (!g || (m && (!sx || !sy || !syaw || !ex || !ey || !eyaw || !ok))) return PP_ERR_INVALID;
    return g->run([=](int i) {
        PP_SLICE(m);
        return pp_collide_dubins(c, cnt, sx + lo, sy + lo, syaw + lo, ex + lo, ey + lo, eyaw + lo, radius, step, ok + lo,
                                 flags);
    });
}
int pp_group_rrt_extend(pp_group *g, size_t m, const double *qx, const double *qy, uint32_t *idx, double *yaw,
                        uint8_t *ok, int nn_flags, int collide_flags) {
    if (!g || (m && (!qx || !qy || !idx || !ok))) return PP_ERR_INVALID;
    return g->run([=](int i) {
        PP_SLICE(m);
        return pp_rrt_extend(c, cnt, qx + lo, qy + lo, idx + lo, yaw ? yaw + lo : nullptr, ok + lo, nn_flags,
                             collide_flags);
    });
}
int pp_group_rrt_extend_dubins(pp_group *g, size_t m, const double *qx, const double *qy, double radius, double step,
                               uint32_t *idx, double *yaw, uint8_t *ok, int nn_flags, int collide_flags) {
    if (!g || (m && (!qx || !qy || !idx || !ok))) return PP_ERR_INVALID;
    return g->run([=](int i) {
        PP_SLICE(m);
        return pp_rrt_extend_dubins(c, cnt, qx + lo, qy + lo, radius, step, idx + lo, yaw ? yaw + lo : nullptr, ok + lo,
                                    nn_flags, collide_flags);
    });
}

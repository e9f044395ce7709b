// api.cu -- the C-ABI of libpathplanning_b200.so: context, device mirrors of tree / obstacles,
// host-pointer wrappers (pinned staging, chunked copy/compute overlap) and measurement helpers.
// Host-side geometry set-up uses the same predicates as the kernels (geo_predicates.cuh).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "geo_predicates.cuh"
#include "pp_common.cuh"
#include "pp_stage.hpp"

// ---- launchers defined in the kernel translation units ---------------------------------------------
int pp_launch_dubins_eval(pp_ctx *, size_t, const double *, const double *, const double *, const double *,
                          const double *, const double *, const double *, double, double *, uint8_t *, double *,
                          cudaStream_t);
int pp_launch_dubins_words(pp_ctx *, size_t, const double *, const double *, const double *, double *, uint8_t *,
                           cudaStream_t);
int pp_launch_mod2pi(pp_ctx *, size_t, const double *, double *, int, cudaStream_t);
int pp_launch_dubins_plan(pp_ctx *, size_t, const double *, const double *, const double *, const double *,
                          const double *, const double *, double, double, int, uint32_t *, void *, void *, cudaStream_t,
                          uint8_t *ok = nullptr, uint32_t *todo = nullptr, unsigned int *todo_count = nullptr);
int pp_launch_dubins_fill(pp_ctx *, size_t, const void *, const uint64_t *, double *, cudaStream_t);
int pp_launch_dubins_path(pp_ctx *, double, double, double, double, double, double, double, double, int, uint32_t, void *,
                          double *, cudaStream_t);
int pp_launch_exclusive_scan(pp_ctx *, size_t, const uint32_t *, uint64_t *, uint64_t *, uint64_t *, cudaStream_t);
int pp_nn_configure(pp_ctx *);
int pp_dubins_tu_init(pp_ctx *);
int pp_collide_tu_init(pp_ctx *);
size_t pp_nn_tile_nodes();
size_t pp_nn_xy32_floats(size_t cap);
int pp_launch_nn(pp_ctx *, size_t, const double *, const double *, uint32_t *, double *, int, cudaStream_t);
int pp_launch_tree_finish(pp_ctx *, size_t, size_t, size_t, cudaStream_t);
int pp_tree_build_grid(pp_ctx *, cudaStream_t);  // nn.cu: device-side counting sort of the nodes by cell
bool pp_nn_grid_policy(const pp_tree_dev &, size_t m);  // nn.cu: true when the appended tail is over budget
int pp_launch_collide_segments(pp_ctx *, size_t, const double *, const double *, const double *, const double *,
                               const uint32_t *, double *, uint8_t *, int, cudaStream_t);
int pp_launch_verify_polylines(pp_ctx *, size_t, const double *, const double *, const uint32_t *, uint8_t *, int,
                               cudaStream_t);
int pp_launch_collide_dubins(pp_ctx *, size_t, const void *, const void *, const double *, const double *, uint8_t *,
                             int, const uint32_t *, const unsigned int *, cudaStream_t);
int pp_launch_fp64_peak(pp_ctx *, int, double *, cudaStream_t, unsigned *, unsigned *);
int pp_launch_rrt_extend_fused(pp_ctx *, size_t, const double *, const double *, uint32_t *, double *, uint8_t *,
                               cudaStream_t);
int pp_launch_extend_gather(pp_ctx *, size_t, const double *, const double *, const uint32_t *, double *, double *,
                            double *, double *, cudaStream_t);

#define PP_AABB_TILE 1024
#define PP_BOUNDS_GRID 256
#ifndef PP_GRID_CELL_SCALE_DEFAULT
#define PP_GRID_CELL_SCALE_DEFAULT 0.5  // obstacle-grid cell edge in units of sqrt(area / rings); swept on the GPU
                                        // (profiles/r03_grid_sweep.txt): 1.0 -> 0.5 takes the C5 slice from 1.36 to 1.29 ms
                                        // and the C4 straight-edge verify from 0.114 to 0.105 ms; 0.25 loses again
#endif
#define PP_SCAN_TILE_ITEMS 2048

// ---------------------------------------------------------------------------------------------------
// errors, launch accounting, timing
// ---------------------------------------------------------------------------------------------------
int pp_fail(pp_ctx *ctx, int status, const char *what, cudaError_t e) {
    if (ctx) {
        char buf[512];
        if (e != cudaSuccess)
            snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
        else
            snprintf(buf, sizeof buf, "%s", what);
        ctx->last_error = buf;
    }
    if (e != cudaSuccess) cudaGetLastError();  // clear the sticky-less error state
    return status;
}

pp_launch_scope::pp_launch_scope(pp_ctx *c, const char *n, int n_launches) : ctx(c), name(n) {
    ctx->launches += (uint64_t)n_launches;
    if (ctx->timing) {
        for (cudaEvent_t *e : {&e0, &e1}) {
            if (!ctx->event_pool.empty()) {
                *e = ctx->event_pool.back();
                ctx->event_pool.pop_back();
            } else if (cudaEventCreate(e) != cudaSuccess) {
                *e = nullptr;
            }
        }
        if (e0) cudaEventRecord(e0, ctx->active_stream);
    }
}
pp_launch_scope::~pp_launch_scope() {
    if (e0 && e1) {
        cudaEventRecord(e1, ctx->active_stream);
        ctx->timings[name].pending.emplace_back(e0, e1);
    }
}

const char *pp_status_string(int s) {
    switch (s) {
    case PP_OK: return "ok";
    case PP_ERR_INVALID: return "invalid argument";
    case PP_ERR_NO_DEVICE: return "no sm_100 CUDA device (there is no CPU fallback)";
    case PP_ERR_CUDA: return "CUDA error";
    case PP_ERR_NOMEM: return "out of memory";
    case PP_ERR_STATE: return "tree or obstacles not uploaded";
    case PP_ERR_OVERFLOW: return "output capacity too small";
    case PP_ERR_COMM: return "NCCL unavailable or collective failed";
    }
    return "unknown status";
}

int pp_abi_version(void) { return PP_ABI_VERSION; }

static bool pp_device_ok(int dev) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return false;
    return p.major == 10;
}

int pp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int d = 0; d < n; ++d)
        if (pp_device_ok(d)) ++ok;
    return ok;
}

int pp_ctx_create(int device, pp_ctx **out) {
    if (!out) return PP_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return PP_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n || !pp_device_ok(device)) return PP_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return PP_ERR_NO_DEVICE;
    pp_ctx *ctx = new pp_ctx();
    ctx->device = device;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, device);
    ctx->sm_count = p.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    ctx->own_stream_handle = ctx->stream;
    for (int i = 0; i < 3 && ok; ++i)
        ok = cudaStreamCreateWithFlags(&ctx->copy_streams[i], cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->tickets, PP_TICKETS * sizeof(unsigned int)) == cudaSuccess;
    ok = ok && cudaMemset(ctx->tickets, 0, PP_TICKETS * sizeof(unsigned int)) == cudaSuccess;
    if (ok) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;  // keep freed blocks cached: the host wrappers allocate per call
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        ctx->active_stream = ctx->stream;
        ok = pp_nn_configure(ctx) == PP_OK && pp_dubins_tu_init(ctx) == PP_OK && pp_collide_tu_init(ctx) == PP_OK;
    }
    if (!ok) {
        pp_ctx_destroy(ctx);
        cudaGetLastError();
        return PP_ERR_CUDA;
    }
    *out = ctx;
    return PP_OK;
}

static void pp_tree_free(pp_tree_dev &t) {
    cudaFree(t.x);
    cudaFree(t.y);
    cudaFree(t.yaw);
    cudaFree(t.parent);
    cudaFree(t.x32);
    cudaFree(t.cell_start);
    cudaFree(t.cell_items);
    cudaFree(t.cell_xy);
    t = pp_tree_dev();
}
void pp_world_free(pp_world_dev &w) {
    cudaFree(w.bx);
    cudaFree(w.by);
    cudaFree(w.bcls);
    cudaFree(w.ox);
    cudaFree(w.oy);
    cudaFree(w.meta);
    cudaFree(w.aabb32);
    cudaFree(w.circ);
    cudaFree(w.cell_start);
    cudaFree(w.cell_items);
    cudaFree(w.cell_box);
    w = pp_world_dev();
}

void pp_ctx_destroy(pp_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    pp_comm_release(ctx);
    for (auto &kv : ctx->timings)
        for (auto &pr : kv.second.pending) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    pp_tree_free(ctx->tree);
    pp_world_free(ctx->world);
    cudaFree(ctx->tickets);
    cudaFree(ctx->scratch);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    delete ctx->stage_pool;
    if (ctx->own_stream_handle) cudaStreamDestroy(ctx->own_stream_handle);
    for (int i = 0; i < 3; ++i)
        if (ctx->copy_streams[i]) cudaStreamDestroy(ctx->copy_streams[i]);
    cudaGetLastError();
    delete ctx;
}

const char *pp_last_error(pp_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
int pp_ctx_device(pp_ctx *ctx) { return ctx ? ctx->device : -1; }
int pp_ctx_sm_count(pp_ctx *ctx) { return ctx ? ctx->sm_count : 0; }
void *pp_ctx_stream(pp_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int pp_ctx_set_stream(pp_ctx *ctx, void *s) {
    if (!ctx) return PP_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream_handle;
    return PP_OK;
}
int pp_sync(pp_ctx *ctx) {
    if (!ctx) return PP_ERR_INVALID;
    pp_guard g(ctx);
    PP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PP_OK;
}
uint64_t pp_launch_count(pp_ctx *ctx) { return ctx ? ctx->launches : 0; }
uint64_t pp_nn_grid_builds(pp_ctx *ctx) { return ctx ? ctx->grid_builds : 0; }

int pp_host_alloc(size_t bytes, void **out) {
    if (!out) return PP_ERR_INVALID;
    *out = nullptr;
    if (bytes == 0) bytes = 1;
    if (cudaHostAlloc(out, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return PP_ERR_NOMEM;
    }
    return PP_OK;
}
int pp_host_free(void *p) {
    if (p && cudaFreeHost(p) != cudaSuccess) {
        cudaGetLastError();
        return PP_ERR_CUDA;
    }
    return PP_OK;
}

int pp_timing_enable(pp_ctx *ctx, int on) {
    if (!ctx) return PP_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->timing = on != 0;
    return PP_OK;
}
static void pp_timing_drain(pp_ctx *ctx, pp_timing_slot &s) {
    for (auto &pr : s.pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
            s.total_ms += ms;
            s.launches += 1;
        }
        ctx->event_pool.push_back(pr.first);
        ctx->event_pool.push_back(pr.second);
    }
    s.pending.clear();
}
int pp_timing_reset(pp_ctx *ctx) {
    if (!ctx) return PP_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->device);
    for (auto &kv : ctx->timings) {
        pp_timing_drain(ctx, kv.second);
        kv.second.total_ms = 0;
        kv.second.launches = 0;
    }
    return PP_OK;
}
int pp_timing_get(pp_ctx *ctx, const char *kernel, double *total_ms, uint64_t *launches) {
    if (!ctx || !kernel) return PP_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->device);
    auto it = ctx->timings.find(kernel);
    double t = 0;
    uint64_t l = 0;
    if (it != ctx->timings.end()) {
        pp_timing_drain(ctx, it->second);
        t = it->second.total_ms;
        l = it->second.launches;
    }
    if (total_ms) *total_ms = t;
    if (launches) *launches = l;
    return PP_OK;
}

int pp_scratch_reserve(pp_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return PP_OK;
    PP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // nothing in flight may still use the old block
    cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = std::max(bytes, (size_t)1 << 20);
    if (cudaMalloc(&ctx->scratch, want) != cudaSuccess) {
        cudaGetLastError();
        return pp_fail(ctx, PP_ERR_NOMEM, "scratch allocation failed");
    }
    ctx->scratch_bytes = want;
    return PP_OK;
}

static int pp_pinned_reserve(pp_ctx *ctx, size_t bytes);

// stream-ordered temporary device buffer
struct pp_tmp {
    void *p = nullptr;
    cudaStream_t s;
    explicit pp_tmp(cudaStream_t st) : s(st) {}
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 1, s); }
    template <class T>
    T *as() { return static_cast<T *>(p); }
    ~pp_tmp() {
        if (p) cudaFreeAsync(p, s);
    }
};

#define PP_TMP(ctx, var, stream, bytes)                                                            \
    pp_tmp var(stream);                                                                            \
    do {                                                                                           \
        if (var.alloc(bytes) != cudaSuccess) {                                                     \
            cudaGetLastError();                                                                    \
            return pp_fail(ctx, PP_ERR_NOMEM, "device allocation failed (" #var ")");              \
        }                                                                                          \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// Dubins
// ---------------------------------------------------------------------------------------------------
static bool pp_pos_finite(double v) { return v > 0.0 && std::isfinite(v); }

int pp_mod2pi(pp_ctx *ctx, size_t n, const double *x, double *out, int pi_2_pi) {
    if (!ctx || (n && (!x || !out))) return PP_ERR_INVALID;
    pp_guard g(ctx);
    cudaStream_t s = ctx->stream;
    PP_TMP(ctx, dx, s, n * 8);
    PP_TMP(ctx, dout, s, n * 8);
    PP_CUDA(ctx, cudaMemcpyAsync(dx.p, x, n * 8, cudaMemcpyHostToDevice, s));
    int rc = pp_launch_mod2pi(ctx, n, dx.as<double>(), dout.as<double>(), pi_2_pi, s);
    if (rc) return rc;
    PP_CUDA(ctx, cudaMemcpyAsync(out, dout.p, n * 8, cudaMemcpyDeviceToHost, s));
    PP_CUDA(ctx, cudaStreamSynchronize(s));
    return PP_OK;
}

int pp_dubins_words(pp_ctx *ctx, size_t n, const double *alpha, const double *beta, const double *d, double *tpq,
                    uint8_t *feasible) {
    if (!ctx || (n && (!alpha || !beta || !d || !tpq || !feasible))) return PP_ERR_INVALID;
    // domain of the word formulas as the reference calls them (src/dubins.rs:336-338: alpha and beta come out of mod2pi,
    // d = hypot * c): the kernels' range reductions are specialised to it, so anything else is refused, not guessed
    for (size_t i = 0; i < n; ++i) {
        const bool ok = alpha[i] >= 0.0 && alpha[i] <= 6.283185307179586 && beta[i] >= 0.0 && beta[i] <= 6.283185307179586 &&
                        d[i] >= 0.0 && d[i] < 1.0e5;
        if (!ok && !(std::isnan(alpha[i]) || std::isnan(beta[i]) || std::isnan(d[i])))  // NaN in -> infeasible out (Q5)
            return pp_fail(ctx, PP_ERR_INVALID, "pp_dubins_words: alpha, beta must lie in [0, 2*pi] and d in [0, 1e5)");
    }
    pp_guard g(ctx);
    cudaStream_t s = ctx->stream;
    PP_TMP(ctx, da, s, n * 8);
    PP_TMP(ctx, db, s, n * 8);
    PP_TMP(ctx, dd, s, n * 8);
    PP_TMP(ctx, dt, s, n * 18 * 8);
    PP_TMP(ctx, df, s, n * 6);
    PP_CUDA(ctx, cudaMemcpyAsync(da.p, alpha, n * 8, cudaMemcpyHostToDevice, s));
    PP_CUDA(ctx, cudaMemcpyAsync(db.p, beta, n * 8, cudaMemcpyHostToDevice, s));
    PP_CUDA(ctx, cudaMemcpyAsync(dd.p, d, n * 8, cudaMemcpyHostToDevice, s));
    int rc = pp_launch_dubins_words(ctx, n, da.as<double>(), db.as<double>(), dd.as<double>(), dt.as<double>(),
                                    df.as<uint8_t>(), s);
    if (rc) return rc;
    PP_CUDA(ctx, cudaMemcpyAsync(tpq, dt.p, n * 18 * 8, cudaMemcpyDeviceToHost, s));
    PP_CUDA(ctx, cudaMemcpyAsync(feasible, df.p, n * 6, cudaMemcpyDeviceToHost, s));
    PP_CUDA(ctx, cudaStreamSynchronize(s));
    return PP_OK;
}

int pp_dubins_eval_dev(pp_ctx *ctx, size_t n, const double *sx, const double *sy, const double *syaw,
                       const double *ex, const double *ey, const double *eyaw, const double *radius_arr,
                       double radius, double *cost, uint8_t *word, double *tpq) {
    if (!ctx || (n && (!sx || !sy || !syaw || !ex || !ey || !eyaw || !cost || !word))) return PP_ERR_INVALID;
    if (!radius_arr && !pp_pos_finite(radius)) return PP_ERR_INVALID;
    pp_guard g(ctx);
    return pp_launch_dubins_eval(ctx, n, sx, sy, syaw, ex, ey, eyaw, radius_arr, radius, cost, word, tpq, ctx->stream);
}

// host pointers: the batch is cut into chunks that rotate over three streams so that the H2D copy of
// chunk c+1, the kernel of chunk c and the D2H copy of chunk c-1 overlap (both PCIe directions busy).
// With pinned buffers (pp_host_alloc) the copies are truly asynchronous.
#ifndef PP_EXTEND_CHUNK_LOG2
#define PP_EXTEND_CHUNK_LOG2 18  // queries per pipelined chunk of the host-pointer extend step (4 MB up, 3.4 MB down); swept on
                                 // the GPU (profiles/r05_extend_e2e.json): 2^15 1.11, 2^16 0.77, 2^17 0.63, 2^18 0.60, 2^19 0.62 ms,
                                 // one chunk (the single-stream route) 0.79 ms per 2^20-query call from pinned memory
#endif
#ifndef PP_EVAL_CHUNK_LOG2
#define PP_EVAL_CHUNK_LOG2 20  // pairs per pipelined chunk of the host-pointer entry (8 MB per input array)
#endif
// true for ordinary host memory (malloc, Vec<f64>, numpy): neither pinned nor registered nor device / managed
static bool pp_is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// pinned ring of the staged path (grown on demand) and its copy threads (PP_STAGE_THREADS overrides; counts the caller)
static int pp_stage_reserve(pp_ctx *ctx, size_t bytes) {
    if (!ctx->stage_pool) {
        // half of the host's hardware threads, between 2 and 8 (swept on a 16-thread B200 host,
        // profiles/r05_stage_threads.json: 2 -> 3.0e8, 4 -> 5.1e8, 6 and 8 -> 6.5e8, 16 -> 7.8e8 pairs/s; the driver's
        // own pageable path: 2.1e8); the cap leaves room for the other ranks of an 8-GPU host
        int t = std::max(2, std::min(8, (int)std::thread::hardware_concurrency() / 2));
        if (const char *e = getenv("PP_STAGE_THREADS")) t = atoi(e);
        ctx->stage_pool = new pp_stage_pool(std::max(1, std::min(t, 32)));
    }
    if (bytes <= ctx->stage_bytes) return PP_OK;
    if (ctx->stage) cudaFreeHost(ctx->stage);
    ctx->stage = nullptr;
    ctx->stage_bytes = 0;
    if (cudaHostAlloc(&ctx->stage, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return pp_fail(ctx, PP_ERR_NOMEM, "pinned staging ring allocation failed");
    }
    ctx->stage_bytes = bytes;
    return PP_OK;
}

// The host-pointer evaluation for PAGEABLE caller memory: the same three-slot chunk pipeline as below, but every
// chunk goes through the context's pinned ring.  The pool copies chunk c's inputs into its slot (and chunk c-3's
// results out of it) with several threads while chunks c-1 and c-2 are on the wire, instead of leaving the staging
// to the driver's single-threaded bounce-buffer path.  Device layout per slot as in pp_dubins_eval.
static int pp_dubins_eval_staged(pp_ctx *ctx, size_t n, const double *const in[7], int n_in, double radius, double *cost,
                                 uint8_t *word, double *tpq) {
    const size_t chunk = (size_t)1 << PP_EVAL_CHUNK_LOG2;
    const size_t per_slot = chunk * (8 * (size_t)n_in + 8 + 8 /*word, padded*/ + (tpq ? 24 : 0));
    int rc = pp_scratch_reserve(ctx, per_slot * 3 + 4096);
    if (rc) return rc;
    rc = pp_stage_reserve(ctx, per_slot * 3);
    if (rc) return rc;
    PP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // see pp_dubins_eval
    struct slot_events {
        cudaEvent_t e[3] = {nullptr, nullptr, nullptr};
        ~slot_events() {
            for (cudaEvent_t x : e)
                if (x) cudaEventDestroy(x);
        }
    } done;
    for (int k = 0; k < 3; ++k) PP_CUDA(ctx, cudaEventCreateWithFlags(&done.e[k], cudaEventDisableTiming));
    const size_t n_chunks = (n + chunk - 1) / chunk;
    auto slot_host = [&](size_t c) { return (char *)ctx->stage + (c % 3) * per_slot; };
    // results of chunk c: pinned slot -> caller's arrays
    auto drain_pieces = [&](size_t c, std::vector<pp_copy_piece> &v) {
        const size_t off = c * chunk, cnt = std::min(chunk, n - off);
        char *h = slot_host(c) + (size_t)n_in * chunk * 8;
        v.push_back({cost + off, h, cnt * 8});
        v.push_back({word + off, h + chunk * 8, cnt});
        if (tpq) v.push_back({tpq + 3 * off, h + chunk * 16, cnt * 24});
    };
    std::vector<pp_copy_piece> pieces;
    for (size_t c = 0; c < n_chunks; ++c) {
        const size_t off = c * chunk, cnt = std::min(chunk, n - off);
        cudaStream_t s = ctx->copy_streams[c % 3];
        ctx->active_stream = s;
        char *h = slot_host(c), *base = (char *)ctx->scratch + (c % 3) * per_slot;
        pieces.clear();
        if (c >= 3) {  // the slot's previous occupant must have landed before it is overwritten
            PP_CUDA(ctx, cudaEventSynchronize(done.e[c % 3]));
            drain_pieces(c - 3, pieces);
        }
        std::vector<pp_copy_piece> fill;
        for (int k = 0; k < n_in; ++k) fill.push_back({h + (size_t)k * chunk * 8, in[k] + off, cnt * 8});
        // drain first: its source is the slot region the fill does not touch (inputs and results do not overlap)
        pieces.insert(pieces.end(), fill.begin(), fill.end());
        ctx->stage_pool->run(pieces.data(), pieces.size());
        double *din[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        for (int k = 0; k < n_in; ++k) din[k] = (double *)(base + (size_t)k * chunk * 8);
        if (cnt == chunk) {  // the slot's input block is contiguous: one copy
            PP_CUDA(ctx, cudaMemcpyAsync(base, h, (size_t)n_in * chunk * 8, cudaMemcpyHostToDevice, s));
        } else {
            for (int k = 0; k < n_in; ++k)
                PP_CUDA(ctx, cudaMemcpyAsync(din[k], h + (size_t)k * chunk * 8, cnt * 8, cudaMemcpyHostToDevice, s));
        }
        double *dcost = (double *)(base + (size_t)n_in * chunk * 8);
        uint8_t *dword = (uint8_t *)(dcost + chunk);
        double *dtpq = tpq ? (double *)(dword + chunk * 8) : nullptr;
        rc = pp_launch_dubins_eval(ctx, cnt, din[0], din[1], din[2], din[3], din[4], din[5], din[6], radius, dcost, dword,
                                   dtpq, s);
        if (rc) return rc;
        char *hout = h + (size_t)n_in * chunk * 8;
        PP_CUDA(ctx, cudaMemcpyAsync(hout, dcost, cnt * 8, cudaMemcpyDeviceToHost, s));
        PP_CUDA(ctx, cudaMemcpyAsync(hout + chunk * 8, dword, cnt, cudaMemcpyDeviceToHost, s));
        if (tpq) PP_CUDA(ctx, cudaMemcpyAsync(hout + chunk * 16, dtpq, cnt * 24, cudaMemcpyDeviceToHost, s));
        PP_CUDA(ctx, cudaEventRecord(done.e[c % 3], s));
    }
    for (size_t c = n_chunks > 3 ? n_chunks - 3 : 0; c < n_chunks; ++c) {
        PP_CUDA(ctx, cudaEventSynchronize(done.e[c % 3]));
        pieces.clear();
        drain_pieces(c, pieces);
        ctx->stage_pool->run(pieces.data(), pieces.size());
    }
    return PP_OK;
}

int pp_dubins_eval(pp_ctx *ctx, size_t n, const double *sx, const double *sy, const double *syaw, const double *ex,
                   const double *ey, const double *eyaw, const double *radius_arr, double radius, double *cost,
                   uint8_t *word, double *tpq) {
    if (!ctx || (n && (!sx || !sy || !syaw || !ex || !ey || !eyaw || !cost || !word))) return PP_ERR_INVALID;
    if (!radius_arr && !pp_pos_finite(radius)) return PP_ERR_INVALID;
    if (n == 0) return PP_OK;
    pp_guard g(ctx);
    const int n_in = radius_arr ? 7 : 6;
    if (n >= ((size_t)2 << PP_EVAL_CHUNK_LOG2) && (pp_is_pageable(sx) || pp_is_pageable(cost))) {
        // large batch in pageable memory: stage it through the pinned ring with the copy threads
        const double *in_all[7] = {sx, sy, syaw, ex, ey, eyaw, radius_arr};
        int rc = pp_dubins_eval_staged(ctx, n, in_all, n_in, radius, cost, word, tpq);
        if (rc != PP_OK) {  // leave nothing in flight that still targets the ring or the scratch block
            for (int k = 0; k < 3; ++k) cudaStreamSynchronize(ctx->copy_streams[k]);
            cudaGetLastError();
        }
        return rc;
    }
    const size_t chunk = std::min(n, (size_t)1 << PP_EVAL_CHUNK_LOG2);
    const size_t per_slot = chunk * (8 * (size_t)n_in + 8 + 8 /*word, padded*/ + (tpq ? 24 : 0));
    int rc = pp_scratch_reserve(ctx, per_slot * 3 + 4096);
    if (rc) return rc;
    // the chunks run on the copy streams, which are not ordered after the context stream: work a `_dev` call left
    // in flight there (scan, grid build) may still be using the scratch block
    PP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double *in[7] = {sx, sy, syaw, ex, ey, eyaw, radius_arr};
    size_t c = 0;
    for (size_t off = 0; off < n; off += chunk, ++c) {
        const size_t cnt = std::min(chunk, n - off);
        cudaStream_t s = ctx->copy_streams[c % 3];
        ctx->active_stream = s;
        char *base = (char *)ctx->scratch + (c % 3) * per_slot;
        double *din[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        for (int k = 0; k < n_in; ++k) {
            din[k] = (double *)(base + (size_t)k * chunk * 8);
            PP_CUDA(ctx, cudaMemcpyAsync(din[k], in[k] + off, cnt * 8, cudaMemcpyHostToDevice, s));
        }
        double *dcost = (double *)(base + (size_t)n_in * chunk * 8);
        uint8_t *dword = (uint8_t *)(dcost + chunk);
        double *dtpq = tpq ? (double *)(dword + chunk * 8) : nullptr;
        rc = pp_launch_dubins_eval(ctx, cnt, din[0], din[1], din[2], din[3], din[4], din[5], din[6], radius, dcost,
                                   dword, dtpq, s);
        if (rc) return rc;
        PP_CUDA(ctx, cudaMemcpyAsync(cost + off, dcost, cnt * 8, cudaMemcpyDeviceToHost, s));
        PP_CUDA(ctx, cudaMemcpyAsync(word + off, dword, cnt, cudaMemcpyDeviceToHost, s));
        if (tpq) PP_CUDA(ctx, cudaMemcpyAsync(tpq + 3 * off, dtpq, cnt * 24, cudaMemcpyDeviceToHost, s));
    }
    for (int k = 0; k < 3; ++k) PP_CUDA(ctx, cudaStreamSynchronize(ctx->copy_streams[k]));
    return PP_OK;
}

int pp_dubins_sample_count_dev(pp_ctx *ctx, size_t n, const double *sx, const double *sy, const double *syaw,
                               const double *ex, const double *ey, const double *eyaw, double radius, double step,
                               int from_origin, uint32_t *counts, void *plan) {
    if (!ctx || (n && (!ex || !ey || !eyaw || !counts))) return PP_ERR_INVALID;
    if (n && !from_origin && (!sx || !sy || !syaw)) return PP_ERR_INVALID;
    if (!pp_pos_finite(radius) || !pp_pos_finite(step)) return PP_ERR_INVALID;
    pp_guard g(ctx);
    return pp_launch_dubins_plan(ctx, n, sx, sy, syaw, ex, ey, eyaw, radius, step, from_origin, counts, plan, nullptr,
                                 ctx->stream);
}

int pp_dubins_sample_count(pp_ctx *ctx, size_t n, const double *sx, const double *sy, const double *syaw,
                           const double *ex, const double *ey, const double *eyaw, double radius, double step,
                           int from_origin, uint32_t *counts, void *plan) {
    if (!ctx || (n && (!ex || !ey || !eyaw || !counts))) return PP_ERR_INVALID;
    if (n && !from_origin && (!sx || !sy || !syaw)) return PP_ERR_INVALID;
    if (!pp_pos_finite(radius) || !pp_pos_finite(step)) return PP_ERR_INVALID;
    if (n == 0) return PP_OK;
    pp_guard g(ctx);
    cudaStream_t s = ctx->stream;
    PP_TMP(ctx, din, s, n * 48);
    PP_TMP(ctx, dcnt, s, n * 4);
    PP_TMP(ctx, dplan, s, n * PP_DUBINS_PLAN_BYTES);
    double *d = din.as<double>();
    const double *in[6] = {sx, sy, syaw, ex, ey, eyaw};
    for (int k = 0; k < 6; ++k) {
        if (in[k])
            PP_CUDA(ctx, cudaMemcpyAsync(d + k * n, in[k], n * 8, cudaMemcpyHostToDevice, s));
        else
            PP_CUDA(ctx, cudaMemsetAsync(d + k * n, 0, n * 8, s));
    }
    int rc = pp_launch_dubins_plan(ctx, n, d, d + n, d + 2 * n, d + 3 * n, d + 4 * n, d + 5 * n, radius, step,
                                   from_origin, dcnt.as<uint32_t>(), plan ? dplan.p : nullptr, nullptr, s);
    if (rc) return rc;
    PP_CUDA(ctx, cudaMemcpyAsync(counts, dcnt.p, n * 4, cudaMemcpyDeviceToHost, s));
    if (plan) PP_CUDA(ctx, cudaMemcpyAsync(plan, dplan.p, n * PP_DUBINS_PLAN_BYTES, cudaMemcpyDeviceToHost, s));
    PP_CUDA(ctx, cudaStreamSynchronize(s));
    return PP_OK;
}

int pp_dubins_sample_fill_dev(pp_ctx *ctx, size_t n, const void *plan, const uint64_t *offsets, uint64_t total,
                              double *out) {
    (void)total;
    if (!ctx || (n && (!plan || !offsets || !out))) return PP_ERR_INVALID;
    pp_guard g(ctx);
    return pp_launch_dubins_fill(ctx, n, plan, offsets, out, ctx->stream);
}

int pp_dubins_sample_fill(pp_ctx *ctx, size_t n, const void *plan, const uint64_t *offsets, uint64_t total,
                          double *out) {
    if (!ctx || (n && (!plan || !offsets)) || (total && !out)) return PP_ERR_INVALID;
    if (n == 0 || total == 0) return PP_OK;
    pp_guard g(ctx);
    cudaStream_t s = ctx->stream;
    PP_TMP(ctx, dplan, s, n * PP_DUBINS_PLAN_BYTES);
    PP_TMP(ctx, doff, s, n * 8);
    PP_TMP(ctx, dout, s, total * 24);
    PP_CUDA(ctx, cudaMemcpyAsync(dplan.p, plan, n * PP_DUBINS_PLAN_BYTES, cudaMemcpyHostToDevice, s));
    PP_CUDA(ctx, cudaMemcpyAsync(doff.p, offsets, n * 8, cudaMemcpyHostToDevice, s));
    int rc = pp_launch_dubins_fill(ctx, n, dplan.p, doff.as<uint64_t>(), dout.as<double>(), s);
    if (rc) return rc;
    PP_CUDA(ctx, cudaMemcpyAsync(out, dout.p, total * 24, cudaMemcpyDeviceToHost, s));
    PP_CUDA(ctx, cudaStreamSynchronize(s));
    return PP_OK;
}

int pp_exclusive_scan_u32_dev(pp_ctx *ctx, size_t n, const uint32_t *counts, uint64_t *offsets, uint64_t *total) {
    if (!ctx || !total || (n && (!counts || !offsets))) return PP_ERR_INVALID;
    pp_guard g(ctx);
    size_t tiles = (n + PP_SCAN_TILE_ITEMS - 1) / PP_SCAN_TILE_ITEMS;
    int rc = pp_scratch_reserve(ctx, (tiles + 1) * 8);
    if (rc) return rc;
    return pp_launch_exclusive_scan(ctx, n, counts, offsets, total, (uint64_t *)ctx->scratch, ctx->stream);
}

// pinned, device-mapped staging buffer of the scalar calls (grown on demand)
static int pp_pinned_reserve(pp_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->pinned_bytes) return PP_OK;
    PP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
    size_t want = std::max(bytes, (size_t)1 << 20);
    if (cudaHostAlloc(&ctx->pinned, want, cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        return pp_fail(ctx, PP_ERR_NOMEM, "pinned staging allocation failed");
    }
    ctx->pinned_bytes = want;
    return PP_OK;
}

struct pp_path_header_host {  // mirrors pp_path_header in dubins.cu
    uint32_t count, word;
    double cost, len[3];
};

int pp_dubins_path(pp_ctx *ctx, double sx, double sy, double syaw, double ex, double ey, double eyaw, double radius,
                   double step, int from_origin, double *px, double *py, double *pyaw, size_t cap, size_t *n_out,
                   int *word, double *cost) {
    if (!ctx || !n_out) return PP_ERR_INVALID;
    if (!pp_pos_finite(radius) || !pp_pos_finite(step)) return PP_ERR_INVALID;
    pp_guard g(ctx);
    cudaStream_t s = ctx->stream;
    *n_out = 0;
    // one launch writes the header and the samples straight into mapped pinned memory
    const size_t cap_eff = std::min(cap, (size_t)1 << 26);
    int rc = pp_pinned_reserve(ctx, 64 + cap_eff * 24);
    if (rc) return rc;
    pp_path_header_host *hdr = (pp_path_header_host *)ctx->pinned;
    double *samples = (double *)((char *)ctx->pinned + 64);
    rc = pp_launch_dubins_path(ctx, sx, sy, syaw, ex, ey, eyaw, radius, step, from_origin, (uint32_t)cap_eff, hdr, samples,
                               s);
    if (rc) return rc;
    PP_CUDA(ctx, cudaStreamSynchronize(s));
    const int w = (int)hdr->word;
    if (word) *word = w;
    if (cost) *cost = hdr->cost;
    if (w == PP_WORD_NONE) return PP_OK;  // reference: None
    const uint32_t h_cnt = hdr->count;
    if (h_cnt == 0xFFFFFFFFu) return pp_fail(ctx, PP_ERR_OVERFLOW, "path too long for the sample replay");
    if (h_cnt > cap_eff) {
        *n_out = h_cnt;
        return pp_fail(ctx, PP_ERR_OVERFLOW, "sample capacity too small");
    }
    if (h_cnt == 0) return PP_OK;
    if (!px || !py) return PP_ERR_INVALID;
    for (uint32_t k = 0; k < h_cnt; ++k) {
        px[k] = samples[3 * k];
        py[k] = samples[3 * k + 1];
        if (pyaw) pyaw[k] = samples[3 * k + 2];
    }
    *n_out = h_cnt;
    return PP_OK;
}

// ---------------------------------------------------------------------------------------------------
// tree
// ---------------------------------------------------------------------------------------------------
int pp_tree_reserve(pp_ctx *ctx, size_t n_total, bool keep) {
    pp_tree_dev &t = ctx->tree;
    const size_t tile = pp_nn_tile_nodes();
    size_t need = ((n_total + tile - 1) / tile) * tile;
    if (need == 0) need = tile;
    if (need <= t.cap) return PP_OK;
    size_t cap = std::max(need, t.cap * 2);
    cap = ((cap + tile - 1) / tile) * tile;
    PP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double *x = nullptr, *y = nullptr, *yaw = nullptr;
    int32_t *parent = nullptr;
    float *x32 = nullptr;
    bool ok = cudaMalloc(&x, cap * 8) == cudaSuccess && cudaMalloc(&y, cap * 8) == cudaSuccess &&
              cudaMalloc(&yaw, cap * 8) == cudaSuccess && cudaMalloc(&parent, cap * 4) == cudaSuccess &&
              cudaMalloc(&x32, pp_nn_xy32_floats(cap) * 4) == cudaSuccess;
    if (!ok) {
        cudaFree(x);
        cudaFree(y);
        cudaFree(yaw);
        cudaFree(parent);
        cudaFree(x32);
        cudaGetLastError();
        return pp_fail(ctx, PP_ERR_NOMEM, "tree allocation failed");
    }
    if (keep && t.n) {
        cudaStream_t s = ctx->stream;
        PP_CUDA(ctx, cudaMemcpyAsync(x, t.x, t.n * 8, cudaMemcpyDeviceToDevice, s));
        PP_CUDA(ctx, cudaMemcpyAsync(y, t.y, t.n * 8, cudaMemcpyDeviceToDevice, s));
        PP_CUDA(ctx, cudaMemcpyAsync(yaw, t.yaw, t.n * 8, cudaMemcpyDeviceToDevice, s));
        PP_CUDA(ctx, cudaMemcpyAsync(parent, t.parent, t.n * 4, cudaMemcpyDeviceToDevice, s));
        PP_CUDA(ctx, cudaMemcpyAsync(x32, t.x32, pp_nn_xy32_floats(t.cap) * 4, cudaMemcpyDeviceToDevice, s));  // blocked layout
        PP_CUDA(ctx, cudaStreamSynchronize(s));
    }
    cudaFree(t.x);
    cudaFree(t.y);
    cudaFree(t.yaw);
    cudaFree(t.parent);
    cudaFree(t.x32);
    t.x = x;
    t.y = y;
    t.yaw = yaw;
    t.parent = parent;
    t.x32 = x32;
    t.cap = cap;
    return PP_OK;
}

// copies k nodes into slots [first, first + k) of the device tree (host or device source)
int pp_tree_copy_in(pp_ctx *ctx, size_t first, size_t k, const double *x, const double *y, const double *yaw,
                    const int32_t *parent, cudaMemcpyKind kind) {
    pp_tree_dev &t = ctx->tree;
    cudaStream_t s = ctx->stream;
    if (k) {
        PP_CUDA(ctx, cudaMemcpyAsync(t.x + first, x, k * 8, kind, s));
        PP_CUDA(ctx, cudaMemcpyAsync(t.y + first, y, k * 8, kind, s));
        if (yaw)
            PP_CUDA(ctx, cudaMemcpyAsync(t.yaw + first, yaw, k * 8, kind, s));
        else
            PP_CUDA(ctx, cudaMemsetAsync(t.yaw + first, 0, k * 8, s));
        if (parent)
            PP_CUDA(ctx, cudaMemcpyAsync(t.parent + first, parent, k * 4, kind, s));
        else
            PP_CUDA(ctx, cudaMemsetAsync(t.parent + first, 0xFF, k * 4, s));
    }
    return PP_OK;
}

// makes slots [first, first + k) part of the tree: fp32 copies, sentinel padding, NN index bookkeeping
int pp_tree_commit(pp_ctx *ctx, size_t first, size_t k, bool sync) {
    pp_tree_dev &t = ctx->tree;
    cudaStream_t s = ctx->stream;
    t.n = first + k;
    if (first == 0) t.grid_n = (size_t)-1;  // a new tree: the node grid (if any) describes another one
    const size_t tile = pp_nn_tile_nodes();
    size_t padded_end = ((t.n + tile - 1) / tile) * tile;
    int rc = pp_launch_tree_finish(ctx, first, t.n, padded_end, s);
    if (rc) return rc;
    if (sync) PP_CUDA(ctx, cudaStreamSynchronize(s));
    return PP_OK;
}

static int pp_tree_put(pp_ctx *ctx, size_t first, size_t k, const double *x, const double *y, const double *yaw,
                       const int32_t *parent, cudaMemcpyKind kind) {
    int rc = pp_tree_copy_in(ctx, first, k, x, y, yaw, parent, kind);
    if (rc) return rc;
    return pp_tree_commit(ctx, first, k, kind == cudaMemcpyHostToDevice);
}

int pp_tree_upload(pp_ctx *ctx, size_t n, const double *x, const double *y, const double *yaw,
                   const int32_t *parent) {
    if (!ctx || (n && (!x || !y))) return PP_ERR_INVALID;
    if (n >= 0xFFFFFFF0ull) return PP_ERR_INVALID;
    pp_guard g(ctx);
    int rc = pp_tree_reserve(ctx, n, false);
    if (rc) return rc;
    return pp_tree_put(ctx, 0, n, x, y, yaw, parent, cudaMemcpyHostToDevice);
}
int pp_tree_upload_dev(pp_ctx *ctx, size_t n, const double *x, const double *y, const double *yaw,
                       const int32_t *parent) {
    if (!ctx || (n && (!x || !y))) return PP_ERR_INVALID;
    if (n >= 0xFFFFFFF0ull) return PP_ERR_INVALID;
    pp_guard g(ctx);
    int rc = pp_tree_reserve(ctx, n, false);
    if (rc) return rc;
    return pp_tree_put(ctx, 0, n, x, y, yaw, parent, cudaMemcpyDeviceToDevice);
}
int pp_tree_append(pp_ctx *ctx, size_t k, const double *x, const double *y, const double *yaw,
                   const int32_t *parent) {
    if (!ctx || (k && (!x || !y))) return PP_ERR_INVALID;
    pp_guard g(ctx);
    if (ctx->tree.n + k >= 0xFFFFFFF0ull) return PP_ERR_INVALID;
    size_t first = ctx->tree.n;
    int rc = pp_tree_reserve(ctx, first + k, true);
    if (rc) return rc;
    return pp_tree_put(ctx, first, k, x, y, yaw, parent, cudaMemcpyHostToDevice);
}
size_t pp_tree_size(pp_ctx *ctx) { return ctx ? ctx->tree.n : 0; }

// ---------------------------------------------------------------------------------------------------
// obstacles
// ---------------------------------------------------------------------------------------------------

static void pp_close_ring(std::vector<double> &x, std::vector<double> &y) {
    // geo-types 0.4 Polygon::new -> LineString::close: push the first point when last != first
    if (!x.empty() && (x.front() != x.back() || y.front() != y.back())) {
        x.push_back(x.front());
        y.push_back(y.front());
    }
}

template <class T>
static int pp_upload_vec(pp_ctx *ctx, T **dst, const std::vector<T> &v) {
    *dst = nullptr;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    if (cudaMalloc((void **)dst, bytes) != cudaSuccess) {
        cudaGetLastError();
        return pp_fail(ctx, PP_ERR_NOMEM, "obstacle allocation failed");
    }
    if (!v.empty()) PP_CUDA(ctx, cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return PP_OK;
}

int pp_obstacles_upload(pp_ctx *ctx, const double *bounds_x, const double *bounds_y, size_t n_bounds,
                        const double *ring_x, const double *ring_y, const uint32_t *ring_off, size_t n_rings) {
    if (!ctx || (n_bounds && (!bounds_x || !bounds_y)) || (n_rings && (!ring_x || !ring_y || !ring_off)))
        return PP_ERR_INVALID;
    if (n_rings >= 0x7FFFFFFFull) return PP_ERR_INVALID;
    pp_guard g(ctx);
    PP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    pp_world_free(ctx->world);
    pp_world_dev &w = ctx->world;

    // ---- bounds ring + classification grid
    std::vector<double> bx(bounds_x, bounds_x + n_bounds), by(bounds_y, bounds_y + n_bounds);
    pp_close_ring(bx, by);
    const uint32_t nb = (uint32_t)bx.size();
    double bminx = INFINITY, bminy = INFINITY, bmaxx = -INFINITY, bmaxy = -INFINITY;
    for (uint32_t i = 0; i < nb; ++i) {
        bminx = std::min(bminx, bx[i]);
        bmaxx = std::max(bmaxx, bx[i]);
        bminy = std::min(bminy, by[i]);
        bmaxy = std::max(bmaxy, by[i]);
    }
    int bgx = PP_BOUNDS_GRID, bgy = PP_BOUNDS_GRID;
    std::vector<uint8_t> bcls;
    double gminx = 0, gminy = 0, binvx = 1, binvy = 1;
    if (nb < 2 || !(bminx <= bmaxx) || !std::isfinite(bminx + bmaxx + bminy + bmaxy)) {
        // no usable grid: fewer than two points -> nothing is ever inside; a non-finite ring -> one cell
        // that swallows every finite point and always takes the exact path
        bgx = bgy = 1;
        bcls.assign(1, nb >= 2 ? 2 : 0);
        gminx = gminy = -1.0e308;
        binvx = binvy = 1e-309;
    } else {
        const double pad = 4.0 * pp_aabb_pad(bminx, bminy, bmaxx, bmaxy);
        double width = (bmaxx - bminx) + 2.0 * pad, height = (bmaxy - bminy) + 2.0 * pad;
        if (!(width > 0.0)) width = 1.0;
        if (!(height > 0.0)) height = 1.0;
        gminx = bminx - pad;
        gminy = bminy - pad;
        const double cw = width / bgx, ch = height / bgy;
        binvx = 1.0 / cw;
        binvy = 1.0 / ch;
        bcls.assign((size_t)bgx * bgy, 3);
        const double mx = 1e-6 * cw + pad, my = 1e-6 * ch + pad;
        auto cellf = [](double v, double lo, double inv, int gmax) {
            double f = std::floor((v - lo) * inv);
            if (!(f > 0.0)) return 0;
            if (f >= (double)gmax) return gmax - 1;
            return (int)f;
        };
        for (uint32_t i = 0; i + 1 < nb; ++i) {
            const double x0 = bx[i], y0 = by[i], x1 = bx[i + 1], y1 = by[i + 1];
            const int cx0 = cellf(std::min(x0, x1) - mx, gminx, binvx, bgx), cx1 = cellf(std::max(x0, x1) + mx, gminx, binvx, bgx);
            const int cy0 = cellf(std::min(y0, y1) - my, gminy, binvy, bgy), cy1 = cellf(std::max(y0, y1) + my, gminy, binvy, bgy);
            const double dx = x1 - x0, dy = y1 - y0;
            const double tol = 1e-9 * (std::fabs(dx) + std::fabs(dy) + 1e-300) * (cw + ch);
            for (int cy = cy0; cy <= cy1; ++cy)
                for (int cx = cx0; cx <= cx1; ++cx) {
                    // cell box grown by the margin; separated from the segment's line iff all four corners
                    // lie strictly on one side
                    const double lx = gminx + cx * cw - mx, hx = gminx + (cx + 1) * cw + mx;
                    const double ly = gminy + cy * ch - my, hy = gminy + (cy + 1) * ch + my;
                    const double c0 = dx * (ly - y0) - dy * (lx - x0), c1 = dx * (ly - y0) - dy * (hx - x0);
                    const double c2 = dx * (hy - y0) - dy * (lx - x0), c3 = dx * (hy - y0) - dy * (hx - x0);
                    const bool pos = c0 > tol && c1 > tol && c2 > tol && c3 > tol;
                    const bool neg = c0 < -tol && c1 < -tol && c2 < -tol && c3 < -tol;
                    if (!(pos || neg)) bcls[(size_t)cy * bgx + cx] = 2;
                }
        }
        for (int cy = 0; cy < bgy; ++cy)
            for (int cx = 0; cx < bgx; ++cx) {
                uint8_t &c = bcls[(size_t)cy * bgx + cx];
                if (c != 3) continue;
                const int pos = pp_point_position(bx.data(), by.data(), nb, gminx + (cx + 0.5) * cw, gminy + (cy + 0.5) * ch);
                c = (pos == 1) ? 1 : ((pos == 0) ? 0 : 2);
            }
    }

    // ---- obstacle rings
    std::vector<double> ox, oy;
    std::vector<pp_ring_meta> meta(n_rings);
    std::vector<pp_ring_circle> circ(n_rings);
    for (size_t r = 0; r < n_rings; ++r) {
        if (ring_off[r + 1] < ring_off[r]) return pp_fail(ctx, PP_ERR_INVALID, "ring_off must be non-decreasing");
        std::vector<double> rx(ring_x + ring_off[r], ring_x + ring_off[r + 1]),
            ry(ring_y + ring_off[r], ring_y + ring_off[r + 1]);
        pp_close_ring(rx, ry);
        pp_ring_meta &m = meta[r];
        m.first = (uint32_t)ox.size();
        m.count = (uint32_t)rx.size();
        m._r0 = m._r1 = 0;
        m.minx = m.miny = INFINITY;
        m.maxx = m.maxy = -INFINITY;
        bool finite = true;
        for (size_t i = 0; i < rx.size(); ++i) {
            finite = finite && std::isfinite(rx[i]) && std::isfinite(ry[i]);
            m.minx = std::min(m.minx, rx[i]);
            m.maxx = std::max(m.maxx, rx[i]);
            m.miny = std::min(m.miny, ry[i]);
            m.maxy = std::max(m.maxy, ry[i]);
        }
        if (!finite) {  // never cull a ring with non-finite coordinates
            m.minx = m.miny = -INFINITY;
            m.maxx = m.maxy = INFINITY;
            m.pad = 0.0;
        } else if (rx.empty()) {
            m.pad = 0.0;
        } else {
            m.pad = pp_aabb_pad(m.minx, m.miny, m.maxx, m.maxy);
        }
        circ[r] = pp_make_ring_circle(rx.data(), ry.data(), (uint32_t)rx.size(), m.minx, m.miny, m.maxx, m.maxy, m.pad,
                                      finite && !rx.empty());
        ox.insert(ox.end(), rx.begin(), rx.end());
        oy.insert(oy.end(), ry.begin(), ry.end());
    }
    if (ox.size() >= 0xFFFFFFF0ull) return pp_fail(ctx, PP_ERR_INVALID, "too many obstacle points");

    // fp32 boxes, rounded outward, padded to the tile size with empty boxes
    const size_t n_tiles = (n_rings + PP_AABB_TILE - 1) / PP_AABB_TILE;
    std::vector<float4> aabb32(n_tiles * PP_AABB_TILE);
    for (size_t r = 0; r < aabb32.size(); ++r) {
        float4 b;
        if (r < n_rings && meta[r].count > 0) {
            const pp_ring_meta &m = meta[r];
            b.x = std::nextafterf((float)(m.minx - m.pad), -INFINITY);
            b.y = std::nextafterf((float)(m.miny - m.pad), -INFINITY);
            b.z = std::nextafterf((float)(m.maxx + m.pad), INFINITY);
            b.w = std::nextafterf((float)(m.maxy + m.pad), INFINITY);
        } else {
            b.x = b.y = INFINITY;
            b.z = b.w = -INFINITY;
        }
        aabb32[r] = b;
    }

    // uniform grid over the padded boxes
    double wminx = INFINITY, wminy = INFINITY, wmaxx = -INFINITY, wmaxy = -INFINITY, ext_sum = 0.0;
    size_t n_fin = 0;
    for (size_t r = 0; r < n_rings; ++r) {
        const pp_ring_meta &m = meta[r];
        if (m.count == 0 || !std::isfinite(m.minx)) continue;
        wminx = std::min(wminx, m.minx - m.pad);
        wmaxx = std::max(wmaxx, m.maxx + m.pad);
        wminy = std::min(wminy, m.miny - m.pad);
        wmaxy = std::max(wmaxy, m.maxy + m.pad);
        ext_sum += std::max(m.maxx - m.minx, m.maxy - m.miny);
        ++n_fin;
    }
    int gx = 1, gy = 1;
    double cell = 1.0, ginv = 1.0, g0x = 0.0, g0y = 0.0;
    if (n_fin) {
        const double gw = std::max(wmaxx - wminx, 1e-300), gh = std::max(wmaxy - wminy, 1e-300);
        // tuning knobs (developer A/B, read at upload): PP_GRID_CELL_SCALE scales the "one ring per cell" edge,
        // PP_GRID_EXT_SCALE the floor of 0.75 mean ring extents
        double cs = PP_GRID_CELL_SCALE_DEFAULT, es = 0.75;
        if (const char *e = getenv("PP_GRID_CELL_SCALE")) cs = atof(e) > 0.0 ? atof(e) : cs;
        if (const char *e = getenv("PP_GRID_EXT_SCALE")) es = atof(e) > 0.0 ? atof(e) : es;
        cell = cs * std::sqrt(gw * gh / (double)n_fin);
        cell = std::max(cell, es * ext_sum / (double)n_fin);
        cell = std::max(cell, std::max(gw, gh) / 2048.0);
        if (!(cell > 0.0) || !std::isfinite(cell)) cell = std::max(gw, gh);
        ginv = 1.0 / cell;
        g0x = wminx - 1e-6 * cell;
        g0y = wminy - 1e-6 * cell;
        gx = (int)std::min(4096.0, std::floor((wmaxx - g0x) * ginv) + 2.0);
        gy = (int)std::min(4096.0, std::floor((wmaxy - g0y) * ginv) + 2.0);
    }
    std::vector<uint32_t> cstart((size_t)gx * gy + 1, 0), citems;
    auto cellr = [&](double v, double lo, int gmax) {
        double f = std::floor((v - lo) * ginv);
        if (!(f > 0.0)) return 0;
        if (f >= (double)gmax) return gmax - 1;
        return (int)f;
    };
    for (int pass = 0; pass < 2; ++pass) {
        std::vector<uint32_t> cur;
        if (pass == 1) {
            for (size_t c = 0; c < (size_t)gx * gy; ++c) cstart[c + 1] += cstart[c];
            citems.resize(std::max<size_t>(cstart.back(), 1));
            cur.assign(cstart.begin(), cstart.end() - 1);
        }
        for (size_t r = 0; r < n_rings; ++r) {
            const pp_ring_meta &m = meta[r];
            if (m.count == 0) continue;
            int cx0 = 0, cx1 = gx - 1, cy0 = 0, cy1 = gy - 1;
            if (std::isfinite(m.minx)) {
                const double e = m.pad + 1e-6 * cell;  // also covers rounding of the cell index itself
                cx0 = cellr(m.minx - e, g0x, gx);
                cx1 = cellr(m.maxx + e, g0x, gx);
                cy0 = cellr(m.miny - e, g0y, gy);
                cy1 = cellr(m.maxy + e, g0y, gy);
            }
            for (int cy = cy0; cy <= cy1; ++cy)
                for (int cx = cx0; cx <= cx1; ++cx) {
                    size_t c = (size_t)cy * gx + cx;
                    if (pass == 0)
                        cstart[c + 1]++;
                    else
                        citems[cur[c]++] = (uint32_t)r;
                }
        }
    }

    int rc;
    if ((rc = pp_upload_vec(ctx, &w.bx, bx))) return rc;
    if ((rc = pp_upload_vec(ctx, &w.by, by))) return rc;
    if ((rc = pp_upload_vec(ctx, &w.bcls, bcls))) return rc;
    if ((rc = pp_upload_vec(ctx, &w.ox, ox))) return rc;
    if ((rc = pp_upload_vec(ctx, &w.oy, oy))) return rc;
    if ((rc = pp_upload_vec(ctx, &w.meta, meta))) return rc;
    if ((rc = pp_upload_vec(ctx, &w.aabb32, aabb32))) return rc;
    if ((rc = pp_upload_vec(ctx, &w.circ, circ))) return rc;
    if ((rc = pp_upload_vec(ctx, &w.cell_start, cstart))) return rc;
    if ((rc = pp_upload_vec(ctx, &w.cell_items, citems))) return rc;
    {
        // the ring boxes once more, in cell order: the walk reads box k and item k side by side instead of item -> box
        std::vector<float4> cbox(citems.size());
        for (size_t k = 0; k < cstart.back(); ++k) cbox[k] = aabb32[citems[k]];
        if ((rc = pp_upload_vec(ctx, &w.cell_box, cbox))) return rc;
    }
    w.nb = nb;
    w.bgx = bgx;
    w.bgy = bgy;
    w.bminx = gminx;
    w.bminy = gminy;
    w.binvx = binvx;
    w.binvy = binvy;
    w.n_pts = (uint32_t)ox.size();
    w.n_rings = (uint32_t)n_rings;
    w.n_aabb_tiles = (uint32_t)n_tiles;
    w.n_cell_items = (uint32_t)citems.size();
    w.gx = gx;
    w.gy = gy;
    w.gminx = g0x;
    w.gminy = g0y;
    w.gcell = cell;
    w.ginv = ginv;
    w.valid = true;
    return PP_OK;
}

pp_world_view pp_make_world_view(const pp_world_dev &w) {
    pp_world_view v;
    v.bx = w.bx;
    v.by = w.by;
    v.nb = w.nb;
    v.bcls = w.bcls;
    v.bgx = w.bgx;
    v.bgy = w.bgy;
    v.bminx = w.bminx;
    v.bminy = w.bminy;
    v.binvx = w.binvx;
    v.binvy = w.binvy;
    v.ox = w.ox;
    v.oy = w.oy;
    v.meta = w.meta;
    v.n_rings = w.n_rings;
    v.aabb32 = w.aabb32;
    v.circ = w.circ;
    v.n_aabb_tiles = w.n_aabb_tiles;
    v.cell_start = w.cell_start;
    v.cell_items = w.cell_items;
    v.cell_box = w.cell_box;
    v.gx = w.gx;
    v.gy = w.gy;
    v.gminx = w.gminx;
    v.gminy = w.gminy;
    v.ginv = w.ginv;
    return v;
}

// ---------------------------------------------------------------------------------------------------
// NN / verify / extend
// ---------------------------------------------------------------------------------------------------
// Resolves PP_NN_DEFAULT to a concrete method and keeps the node grid usable when the grid search is chosen.
// Every method returns the same bit-exact argmin (lowest index on ties), so the choice is about time only.
static int pp_nn_prepare(pp_ctx *ctx, size_t m, int *flags) {
    const pp_tree_dev &t = ctx->tree;
    const bool forced = (*flags & (PP_NN_GRID | PP_NN_SCAN | PP_NN_PLAIN_F64 | PP_NN_UNSORTED)) != 0;
    if (!forced && t.n >= PP_NN_GRID_MIN_NODES) *flags |= PP_NN_GRID;
    // the grid is incremental: appended nodes form a linearly scanned tail, and the O(n) rebuild happens once per
    // PP_NN_TAIL_MAX appended nodes (or when this call's m * tail outweighs it), never once per append -- the scalar
    // plan_one loop (one append + one query per iteration, src/rrt.rs:583-597) pays it every 4 096 iterations
    if ((*flags & PP_NN_GRID) && pp_nn_grid_policy(t, m)) return pp_tree_build_grid(ctx, ctx->stream);
    return PP_OK;
}

int pp_nn_dev(pp_ctx *ctx, size_t m, const double *qx, const double *qy, uint32_t *idx, double *d2, int flags) {
    if (!ctx || (m && (!qx || !qy || !idx))) return PP_ERR_INVALID;
    pp_guard g(ctx);
    int rc = pp_nn_prepare(ctx, m, &flags);
    if (rc) return rc;
    return pp_launch_nn(ctx, m, qx, qy, idx, d2, flags, ctx->stream);
}

int pp_nn(pp_ctx *ctx, size_t m, const double *qx, const double *qy, uint32_t *idx, double *d2, int flags) {
    if (!ctx || (m && (!qx || !qy || !idx))) return PP_ERR_INVALID;
    if (m == 0) return PP_OK;
    pp_guard g(ctx);
    int rc = pp_nn_prepare(ctx, m, &flags);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    if (m <= 64) {
        // the scalar get_nearest_node call: queries and answers live in mapped pinned memory, so the whole
        // call is one kernel launch and one stream synchronisation (no allocation, no explicit copy)
        rc = pp_pinned_reserve(ctx, 4096);
        if (rc) return rc;
        double *q = (double *)ctx->pinned;          // [0, 1024): qx, qy
        double *od2 = q + 128;                       // [1024, 1536)
        uint32_t *oidx = (uint32_t *)(q + 192);      // [1536, 1792)
        memcpy(q, qx, m * 8);
        memcpy(q + 64, qy, m * 8);
        rc = pp_launch_nn(ctx, m, q, q + 64, oidx, d2 ? od2 : nullptr, flags, s);
        if (rc) return rc;
        PP_CUDA(ctx, cudaStreamSynchronize(s));
        memcpy(idx, oidx, m * 4);
        if (d2) memcpy(d2, od2, m * 8);
        return PP_OK;
    }
    PP_TMP(ctx, dq, s, m * 16);
    PP_TMP(ctx, di, s, m * 4);
    PP_TMP(ctx, dd, s, m * 8);
    double *q = dq.as<double>();
    PP_CUDA(ctx, cudaMemcpyAsync(q, qx, m * 8, cudaMemcpyHostToDevice, s));
    PP_CUDA(ctx, cudaMemcpyAsync(q + m, qy, m * 8, cudaMemcpyHostToDevice, s));
    rc = pp_launch_nn(ctx, m, q, q + m, di.as<uint32_t>(), d2 ? dd.as<double>() : nullptr, flags, s);
    if (rc) return rc;
    PP_CUDA(ctx, cudaMemcpyAsync(idx, di.p, m * 4, cudaMemcpyDeviceToHost, s));
    if (d2) PP_CUDA(ctx, cudaMemcpyAsync(d2, dd.p, m * 8, cudaMemcpyDeviceToHost, s));
    PP_CUDA(ctx, cudaStreamSynchronize(s));
    return PP_OK;
}

int pp_collide_segments_dev(pp_ctx *ctx, size_t m, const double *ax, const double *ay, const double *bx,
                            const double *by, uint8_t *ok, int flags) {
    if (!ctx || (m && (!ax || !ay || !bx || !by || !ok))) return PP_ERR_INVALID;
    pp_guard g(ctx);
    if (!ctx->world.valid) return pp_fail(ctx, PP_ERR_STATE, "obstacles not uploaded");
    return pp_launch_collide_segments(ctx, m, ax, ay, bx, by, nullptr, nullptr, ok, flags, ctx->stream);
}

int pp_collide_segments(pp_ctx *ctx, size_t m, const double *ax, const double *ay, const double *bx,
                        const double *by, uint8_t *ok, int flags) {
    if (!ctx || (m && (!ax || !ay || !bx || !by || !ok))) return PP_ERR_INVALID;
    if (m == 0) return PP_OK;
    pp_guard g(ctx);
    if (!ctx->world.valid) return pp_fail(ctx, PP_ERR_STATE, "obstacles not uploaded");
    cudaStream_t s = ctx->stream;
    PP_TMP(ctx, de, s, m * 32);
    PP_TMP(ctx, dok, s, m);
    double *e = de.as<double>();
    const double *in[4] = {ax, ay, bx, by};
    for (int k = 0; k < 4; ++k) PP_CUDA(ctx, cudaMemcpyAsync(e + k * m, in[k], m * 8, cudaMemcpyHostToDevice, s));
    int rc = pp_launch_collide_segments(ctx, m, e, e + m, e + 2 * m, e + 3 * m, nullptr, nullptr, dok.as<uint8_t>(),
                                        flags, s);
    if (rc) return rc;
    PP_CUDA(ctx, cudaMemcpyAsync(ok, dok.p, m, cudaMemcpyDeviceToHost, s));
    PP_CUDA(ctx, cudaStreamSynchronize(s));
    return PP_OK;
}

int pp_verify_polylines(pp_ctx *ctx, size_t n_lines, const double *px, const double *py, const uint32_t *line_off,
                        uint8_t *ok, int flags) {
    if (!ctx || (n_lines && (!line_off || !ok))) return PP_ERR_INVALID;
    if (n_lines == 0) return PP_OK;
    const size_t n_pts = line_off[n_lines];
    if (n_pts && (!px || !py)) return PP_ERR_INVALID;
    pp_guard g(ctx);
    if (!ctx->world.valid) return pp_fail(ctx, PP_ERR_STATE, "obstacles not uploaded");
    cudaStream_t s = ctx->stream;
    PP_TMP(ctx, dp, s, n_pts * 16);
    PP_TMP(ctx, doff, s, (n_lines + 1) * 4);
    PP_TMP(ctx, dok, s, n_lines);
    double *p = dp.as<double>();
    if (n_pts) {
        PP_CUDA(ctx, cudaMemcpyAsync(p, px, n_pts * 8, cudaMemcpyHostToDevice, s));
        PP_CUDA(ctx, cudaMemcpyAsync(p + n_pts, py, n_pts * 8, cudaMemcpyHostToDevice, s));
    }
    PP_CUDA(ctx, cudaMemcpyAsync(doff.p, line_off, (n_lines + 1) * 4, cudaMemcpyHostToDevice, s));
    int rc = pp_launch_verify_polylines(ctx, n_lines, p, p + n_pts, doff.as<uint32_t>(), dok.as<uint8_t>(), flags, s);
    if (rc) return rc;
    PP_CUDA(ctx, cudaMemcpyAsync(ok, dok.p, n_lines, cudaMemcpyDeviceToHost, s));
    PP_CUDA(ctx, cudaStreamSynchronize(s));
    return PP_OK;
}

static int pp_collide_dubins_impl(pp_ctx *ctx, size_t m, const double *sx, const double *sy, const double *syaw,
                                  const double *ex, const double *ey, const double *eyaw, double radius, double step,
                                  uint8_t *ok, int flags, cudaStream_t s) {
    // pass 1: plan records (evaluate + the sampling loop's counts) and, per path, the path-level box test -- a path
    // with nothing registered under its bounding box is answered there; pass 2: generate-and-test per warp for the
    // paths on pass 1's list (all of them under PP_COLLIDE_NO_CULL)
    PP_TMP(ctx, dcnt, s, m * 4);
    PP_TMP(ctx, dplan, s, m * PP_DUBINS_PLAN_BYTES);
    PP_TMP(ctx, daux, s, m * 136);  // pp_plan_aux: segment origins + sincos(syaw), computed once per path by pass 1
    uint32_t *todo = nullptr;
    unsigned int *todo_count = nullptr;
    PP_TMP(ctx, dtodo, s, (m + 1) * 4);
    if (!(flags & PP_COLLIDE_NO_CULL) && m < 0xFFFFFFFFull) {
        todo_count = dtodo.as<unsigned int>();
        todo = dtodo.as<uint32_t>() + 1;
        PP_CUDA(ctx, cudaMemsetAsync(todo_count, 0, 4, s));
    }
    int rc = pp_launch_dubins_plan(ctx, m, sx, sy, syaw, ex, ey, eyaw, radius, step, 0, dcnt.as<uint32_t>(), dplan.p,
                                   daux.p, s, ok, todo, todo_count);
    if (rc) return rc;
    return pp_launch_collide_dubins(ctx, m, dplan.p, daux.p, ex, ey, ok, flags, todo, todo_count, s);
}

int pp_collide_dubins_dev(pp_ctx *ctx, size_t m, const double *sx, const double *sy, const double *syaw,
                          const double *ex, const double *ey, const double *eyaw, double radius, double step,
                          uint8_t *ok, int flags) {
    if (!ctx || (m && (!sx || !sy || !syaw || !ex || !ey || !eyaw || !ok))) return PP_ERR_INVALID;
    if (!pp_pos_finite(radius) || !pp_pos_finite(step)) return PP_ERR_INVALID;
    if (m == 0) return PP_OK;
    pp_guard g(ctx);
    if (!ctx->world.valid) return pp_fail(ctx, PP_ERR_STATE, "obstacles not uploaded");
    return pp_collide_dubins_impl(ctx, m, sx, sy, syaw, ex, ey, eyaw, radius, step, ok, flags, ctx->stream);
}

int pp_collide_dubins(pp_ctx *ctx, size_t m, const double *sx, const double *sy, const double *syaw,
                      const double *ex, const double *ey, const double *eyaw, double radius, double step, uint8_t *ok,
                      int flags) {
    if (!ctx || (m && (!sx || !sy || !syaw || !ex || !ey || !eyaw || !ok))) return PP_ERR_INVALID;
    if (!pp_pos_finite(radius) || !pp_pos_finite(step)) return PP_ERR_INVALID;
    if (m == 0) return PP_OK;
    pp_guard g(ctx);
    if (!ctx->world.valid) return pp_fail(ctx, PP_ERR_STATE, "obstacles not uploaded");
    cudaStream_t s = ctx->stream;
    if (m <= 1024) {
        // verify_node on a chain of a few edges: poses in, flags out through mapped pinned memory
        int rc0 = pp_pinned_reserve(ctx, 64 + 1024 * 56);
        if (rc0) return rc0;
        double *d = (double *)ctx->pinned;
        uint8_t *pok = (uint8_t *)(d + 6 * m);
        const double *in[6] = {sx, sy, syaw, ex, ey, eyaw};
        for (int k = 0; k < 6; ++k) memcpy(d + k * m, in[k], m * 8);
        rc0 = pp_collide_dubins_impl(ctx, m, d, d + m, d + 2 * m, d + 3 * m, d + 4 * m, d + 5 * m, radius, step, pok,
                                     flags, s);
        if (rc0) return rc0;
        PP_CUDA(ctx, cudaStreamSynchronize(s));
        memcpy(ok, pok, m);
        return PP_OK;
    }
    PP_TMP(ctx, din, s, m * 48);
    PP_TMP(ctx, dok, s, m);
    double *d = din.as<double>();
    const double *in[6] = {sx, sy, syaw, ex, ey, eyaw};
    for (int k = 0; k < 6; ++k) PP_CUDA(ctx, cudaMemcpyAsync(d + k * m, in[k], m * 8, cudaMemcpyHostToDevice, s));
    int rc = pp_collide_dubins_impl(ctx, m, d, d + m, d + 2 * m, d + 3 * m, d + 4 * m, d + 5 * m, radius, step,
                                    dok.as<uint8_t>(), flags, s);
    if (rc) return rc;
    PP_CUDA(ctx, cudaMemcpyAsync(ok, dok.p, m, cudaMemcpyDeviceToHost, s));
    PP_CUDA(ctx, cudaStreamSynchronize(s));
    return PP_OK;
}

// one extend step on device pointers: the NN kernel followed by the verify kernel the flags name; with
// PP_COLLIDE_FUSED (and both halves on their grid route) the binned, fused kernel instead
static int pp_extend_step(pp_ctx *ctx, size_t m, const double *qx, const double *qy, uint32_t *idx, double *yaw,
                          uint8_t *ok, int nn_flags, int collide_flags, cudaStream_t s) {
    const bool fused = (collide_flags & PP_COLLIDE_FUSED) && (nn_flags & PP_NN_GRID) &&
                       !(nn_flags & (PP_NN_SCAN | PP_NN_PLAIN_F64 | PP_NN_UNSORTED)) &&
                       !(collide_flags & (PP_COLLIDE_NO_CULL | PP_COLLIDE_UNSORTED | PP_COLLIDE_SCAN));
    if (fused) return pp_launch_rrt_extend_fused(ctx, m, qx, qy, idx, yaw, ok, s);
    int rc = pp_launch_nn(ctx, m, qx, qy, idx, nullptr, nn_flags, s);
    if (rc) return rc;
    return pp_launch_collide_segments(ctx, m, qx, qy, nullptr, nullptr, idx, yaw, ok, collide_flags, s);
}

int pp_rrt_extend_dev(pp_ctx *ctx, size_t m, const double *qx, const double *qy, uint32_t *idx, double *yaw,
                      uint8_t *ok, int nn_flags, int collide_flags) {
    if (!ctx || (m && (!qx || !qy || !idx || !ok))) return PP_ERR_INVALID;
    if (m == 0) return PP_OK;
    pp_guard g(ctx);
    if (!ctx->world.valid) return pp_fail(ctx, PP_ERR_STATE, "obstacles not uploaded");
    if (ctx->tree.n == 0) return pp_fail(ctx, PP_ERR_STATE, "tree is empty");
    int rc = pp_nn_prepare(ctx, m, &nn_flags);
    if (rc) return rc;
    return pp_extend_step(ctx, m, qx, qy, idx, yaw, ok, nn_flags, collide_flags, ctx->stream);
}

int pp_rrt_extend(pp_ctx *ctx, size_t m, const double *qx, const double *qy, uint32_t *idx, double *yaw, uint8_t *ok,
                  int nn_flags, int collide_flags) {
    if (!ctx || (m && (!qx || !qy || !idx || !ok))) return PP_ERR_INVALID;
    if (m == 0) return PP_OK;
    pp_guard g(ctx);
    if (!ctx->world.valid) return pp_fail(ctx, PP_ERR_STATE, "obstacles not uploaded");
    if (ctx->tree.n == 0) return pp_fail(ctx, PP_ERR_STATE, "tree is empty");
    int rc = pp_nn_prepare(ctx, m, &nn_flags);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    // Large batch on the default route (grid NN + grid verify: neither kernel touches the context's scratch block or
    // tickets, so chunks may run side by side): cut it into chunks that rotate over the three copy streams, as
    // pp_dubins_eval does -- the upload of chunk c+1, the two kernels of chunk c and the download of chunk c-1 overlap,
    // instead of 16 B in, two kernels and 13 B out per query strictly one after the other.
    const bool grid_route = (nn_flags & PP_NN_GRID) && !(nn_flags & (PP_NN_SCAN | PP_NN_PLAIN_F64 | PP_NN_UNSORTED)) &&
                            !(collide_flags & (PP_COLLIDE_NO_CULL | PP_COLLIDE_UNSORTED | PP_COLLIDE_SCAN | PP_COLLIDE_FUSED));
    int chunk_log2 = PP_EXTEND_CHUNK_LOG2;
    if (const char *e = getenv("PP_EXTEND_CHUNK_LOG2")) chunk_log2 = std::max(12, std::min(26, atoi(e)));  // developer sweep
    const size_t chunk = (size_t)1 << chunk_log2;
    if (grid_route && m >= 2 * chunk) {
        const size_t per_slot = chunk * (16 + 8 + 4 + 4 /*ok, padded*/);
        rc = pp_scratch_reserve(ctx, per_slot * 3);
        if (rc) return rc;
        PP_CUDA(ctx, cudaStreamSynchronize(s));  // the node grid (built on the context stream) is complete; scratch is free
        size_t c = 0;
        for (size_t off = 0; off < m; off += chunk, ++c) {
            const size_t cnt = std::min(chunk, m - off);
            cudaStream_t cs = ctx->copy_streams[c % 3];
            ctx->active_stream = cs;
            char *base = (char *)ctx->scratch + (c % 3) * per_slot;
            double *dqx = (double *)base, *dqy = dqx + chunk, *dyaw = dqy + chunk;
            uint32_t *didx = (uint32_t *)(dyaw + chunk);
            uint8_t *dok = (uint8_t *)(didx + chunk);
            PP_CUDA(ctx, cudaMemcpyAsync(dqx, qx + off, cnt * 8, cudaMemcpyHostToDevice, cs));
            PP_CUDA(ctx, cudaMemcpyAsync(dqy, qy + off, cnt * 8, cudaMemcpyHostToDevice, cs));
            rc = pp_extend_step(ctx, cnt, dqx, dqy, didx, dyaw, dok, nn_flags, collide_flags, cs);
            if (rc) break;
            PP_CUDA(ctx, cudaMemcpyAsync(idx + off, didx, cnt * 4, cudaMemcpyDeviceToHost, cs));
            if (yaw) PP_CUDA(ctx, cudaMemcpyAsync(yaw + off, dyaw, cnt * 8, cudaMemcpyDeviceToHost, cs));
            PP_CUDA(ctx, cudaMemcpyAsync(ok + off, dok, cnt, cudaMemcpyDeviceToHost, cs));
        }
        for (int k = 0; k < 3; ++k) PP_CUDA(ctx, cudaStreamSynchronize(ctx->copy_streams[k]));
        return rc;
    }
    PP_TMP(ctx, dq, s, m * 16);
    PP_TMP(ctx, di, s, m * 4);
    PP_TMP(ctx, dy, s, m * 8);
    PP_TMP(ctx, dok, s, m);
    double *q = dq.as<double>();
    PP_CUDA(ctx, cudaMemcpyAsync(q, qx, m * 8, cudaMemcpyHostToDevice, s));
    PP_CUDA(ctx, cudaMemcpyAsync(q + m, qy, m * 8, cudaMemcpyHostToDevice, s));
    rc = pp_extend_step(ctx, m, q, q + m, di.as<uint32_t>(), dy.as<double>(), dok.as<uint8_t>(), nn_flags, collide_flags, s);
    if (rc) return rc;
    PP_CUDA(ctx, cudaMemcpyAsync(idx, di.p, m * 4, cudaMemcpyDeviceToHost, s));
    if (yaw) PP_CUDA(ctx, cudaMemcpyAsync(yaw, dy.p, m * 8, cudaMemcpyDeviceToHost, s));
    PP_CUDA(ctx, cudaMemcpyAsync(ok, dok.p, m, cudaMemcpyDeviceToHost, s));
    PP_CUDA(ctx, cudaStreamSynchronize(s));
    return PP_OK;
}

static int pp_extend_dubins_impl(pp_ctx *ctx, size_t m, const double *qx, const double *qy, double radius, double step,
                                 uint32_t *idx, double *yaw, uint8_t *ok, int nn_flags, int collide_flags,
                                 cudaStream_t s) {
    int rc = pp_launch_nn(ctx, m, qx, qy, idx, nullptr, nn_flags, s);
    if (rc) return rc;
    PP_TMP(ctx, de, s, m * 24);  // ex, ey, eyaw of the nearest nodes
    double *e = de.as<double>();
    rc = pp_launch_extend_gather(ctx, m, qx, qy, idx, yaw, e, e + m, e + 2 * m, s);
    if (rc) return rc;
    return pp_collide_dubins_impl(ctx, m, qx, qy, yaw, e, e + m, e + 2 * m, radius, step, ok, collide_flags, s);
}

int pp_rrt_extend_dubins_dev(pp_ctx *ctx, size_t m, const double *qx, const double *qy, double radius, double step,
                             uint32_t *idx, double *yaw, uint8_t *ok, int nn_flags, int collide_flags) {
    if (!ctx || (m && (!qx || !qy || !idx || !yaw || !ok))) return PP_ERR_INVALID;
    if (!pp_pos_finite(radius) || !pp_pos_finite(step)) return PP_ERR_INVALID;
    if (m == 0) return PP_OK;
    pp_guard g(ctx);
    if (!ctx->world.valid) return pp_fail(ctx, PP_ERR_STATE, "obstacles not uploaded");
    if (ctx->tree.n == 0) return pp_fail(ctx, PP_ERR_STATE, "tree is empty");
    int rc = pp_nn_prepare(ctx, m, &nn_flags);
    if (rc) return rc;
    return pp_extend_dubins_impl(ctx, m, qx, qy, radius, step, idx, yaw, ok, nn_flags, collide_flags, ctx->stream);
}

int pp_rrt_extend_dubins(pp_ctx *ctx, size_t m, const double *qx, const double *qy, double radius, double step,
                         uint32_t *idx, double *yaw, uint8_t *ok, int nn_flags, int collide_flags) {
    if (!ctx || (m && (!qx || !qy || !idx || !ok))) return PP_ERR_INVALID;
    if (!pp_pos_finite(radius) || !pp_pos_finite(step)) return PP_ERR_INVALID;
    if (m == 0) return PP_OK;
    pp_guard g(ctx);
    if (!ctx->world.valid) return pp_fail(ctx, PP_ERR_STATE, "obstacles not uploaded");
    if (ctx->tree.n == 0) return pp_fail(ctx, PP_ERR_STATE, "tree is empty");
    int rc = pp_nn_prepare(ctx, m, &nn_flags);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    PP_TMP(ctx, dq, s, m * 16);
    PP_TMP(ctx, di, s, m * 4);
    PP_TMP(ctx, dy, s, m * 8);
    PP_TMP(ctx, dok, s, m);
    double *q = dq.as<double>();
    PP_CUDA(ctx, cudaMemcpyAsync(q, qx, m * 8, cudaMemcpyHostToDevice, s));
    PP_CUDA(ctx, cudaMemcpyAsync(q + m, qy, m * 8, cudaMemcpyHostToDevice, s));
    rc = pp_extend_dubins_impl(ctx, m, q, q + m, radius, step, di.as<uint32_t>(), dy.as<double>(), dok.as<uint8_t>(),
                               nn_flags, collide_flags, s);
    if (rc) return rc;
    PP_CUDA(ctx, cudaMemcpyAsync(idx, di.p, m * 4, cudaMemcpyDeviceToHost, s));
    if (yaw) PP_CUDA(ctx, cudaMemcpyAsync(yaw, dy.p, m * 8, cudaMemcpyDeviceToHost, s));
    PP_CUDA(ctx, cudaMemcpyAsync(ok, dok.p, m, cudaMemcpyDeviceToHost, s));
    PP_CUDA(ctx, cudaStreamSynchronize(s));
    return PP_OK;
}

// ---------------------------------------------------------------------------------------------------
// measurement helpers
// ---------------------------------------------------------------------------------------------------
int pp_measure_fp64_peak(pp_ctx *ctx, int iters, double *dfma_per_s, double *ms_out) {
    if (!ctx || !dfma_per_s || iters <= 0) return PP_ERR_INVALID;
    pp_guard g(ctx);
    cudaStream_t s = ctx->stream;
    PP_TMP(ctx, dsink, s, 1 << 20);
    cudaEvent_t e0, e1;
    PP_CUDA(ctx, cudaEventCreate(&e0));
    PP_CUDA(ctx, cudaEventCreate(&e1));
    unsigned threads = 0, per_thread = 0;
    int rc = pp_launch_fp64_peak(ctx, 8, dsink.as<double>(), s, &threads, &per_thread);  // warm-up
    if (rc == PP_OK) {
        cudaEventRecord(e0, s);
        rc = pp_launch_fp64_peak(ctx, iters, dsink.as<double>(), s, &threads, &per_thread);
        cudaEventRecord(e1, s);
    }
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return rc;
    PP_CUDA(ctx, e);
    *dfma_per_s = (double)threads * (double)per_thread * (double)iters / ((double)ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return PP_OK;
}

int pp_measure_copy(pp_ctx *ctx, size_t h2d_bytes, size_t d2h_bytes, int pinned, double *ms_out) {
    if (!ctx || !ms_out) return PP_ERR_INVALID;
    pp_guard g(ctx);
    void *h_up = nullptr, *h_dn = nullptr, *d_up = nullptr, *d_dn = nullptr;
    auto cleanup = [&]() {
        if (pinned) {
            if (h_up) cudaFreeHost(h_up);
            if (h_dn) cudaFreeHost(h_dn);
        } else {
            free(h_up);
            free(h_dn);
        }
        cudaFree(d_up);
        cudaFree(d_dn);
        cudaGetLastError();
    };
    bool ok = true;
    if (pinned) {
        ok = cudaHostAlloc(&h_up, h2d_bytes ? h2d_bytes : 1, cudaHostAllocDefault) == cudaSuccess &&
             cudaHostAlloc(&h_dn, d2h_bytes ? d2h_bytes : 1, cudaHostAllocDefault) == cudaSuccess;
    } else {
        h_up = malloc(h2d_bytes ? h2d_bytes : 1);
        h_dn = malloc(d2h_bytes ? d2h_bytes : 1);
        ok = h_up && h_dn;
    }
    ok = ok && cudaMalloc(&d_up, h2d_bytes ? h2d_bytes : 1) == cudaSuccess &&
         cudaMalloc(&d_dn, d2h_bytes ? d2h_bytes : 1) == cudaSuccess;
    if (!ok) {
        cleanup();
        return pp_fail(ctx, PP_ERR_NOMEM, "copy benchmark allocation failed");
    }
    memset(h_up, 1, h2d_bytes ? h2d_bytes : 1);  // touch the pages: first-touch faults are not copy time
    memset(h_dn, 1, d2h_bytes ? d2h_bytes : 1);
    cudaStream_t s0 = ctx->copy_streams[0], s1 = ctx->copy_streams[1];
    const size_t chunk = (size_t)32 << 20;
    double best = 1e300;
    for (int rep = 0; rep < 3; ++rep) {  // best of 3 (the first pass also warms the driver's staging path)
        cudaDeviceSynchronize();
        auto t0 = std::chrono::steady_clock::now();
        for (size_t off = 0; off < h2d_bytes || off < d2h_bytes; off += chunk) {
            if (off < h2d_bytes)
                cudaMemcpyAsync((char *)d_up + off, (char *)h_up + off, std::min(chunk, h2d_bytes - off),
                                cudaMemcpyHostToDevice, s0);
            if (off < d2h_bytes)
                cudaMemcpyAsync((char *)h_dn + off, (char *)d_dn + off, std::min(chunk, d2h_bytes - off),
                                cudaMemcpyDeviceToHost, s1);
        }
        cudaStreamSynchronize(s0);
        cudaStreamSynchronize(s1);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        best = std::min(best, ms);
    }
    cudaError_t e = cudaGetLastError();
    cleanup();
    PP_CUDA(ctx, e);
    *ms_out = best;
    return PP_OK;
}

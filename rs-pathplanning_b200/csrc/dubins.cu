// dubins.cu -- batched Dubins kernels: evaluate (six words + minimum), plan/count and sample fill.
// Compiled with -fmad=false (see dubins_device.cuh).
//
// Reference arithmetic: src/dubins.rs:14-428.  Layout: SoA f64 arrays sx[], sy[], syaw[], ex[], ey[],
// eyaw[] (one coalesced 8-byte load per lane and array); outputs cost[] f64, word[] u8, optional tpq[3n].
#include <algorithm>

#include "dubins_device.cuh"
#include "pp_common.cuh"
#include "path_box.cuh"
#include "pp_replay.cuh"

pp_world_view pp_make_world_view(const pp_world_dev &w);  // api.cu

// ------------------------------------------------------------------------------------------------
// kernel 1: evaluate.  One thread per pose pair; FP64-pipe bound (about 490 DP instructions per
// 57 algorithmic bytes), so the grid is simply n/128 CTAs of 4 warps: small CTAs keep the tail short.
// ------------------------------------------------------------------------------------------------
#define PP_EVAL_THREADS 128

#ifndef PP_EVAL_MIN_BLOCKS
#define PP_EVAL_MIN_BLOCKS 6  // 80 registers; measured per 2^24 pairs: 0.658 ms (6), 0.661 (5), 0.665 (7), 0.692 (8)
#endif
// one pose pair: to_local + frame + six words + selection (src/dubins.rs:401-408, 333-363); c = 1 / turn_radius
template <bool WANT_TPQ>
__device__ __forceinline__ void pp_eval_one(double sx, double sy, double syaw, double ex, double ey, double eyaw,
                                            double c, size_t i, double *__restrict__ cost,
                                            uint8_t *__restrict__ word, double *__restrict__ tpq) {
    const pp_dubins_frame f = pp_dubins_frame_world(ex - sx, ey - sy, syaw, eyaw - syaw, c);
    pp_dubins_sol s = pp_dubins_solve<false>(f.alpha, f.beta, f.d, nullptr, nullptr);
    cost[i] = s.cost;
    word[i] = (uint8_t)s.word;
    if (WANT_TPQ) {
        tpq[3 * i + 0] = s.t;
        tpq[3 * i + 1] = s.p;
        tpq[3 * i + 2] = s.q;
    }
}

// One thread per pair.  A persistent variant (per-warp two-stage ring in shared memory fed by 1-D bulk copies,
// 148 x 6 CTAs) was measured and dropped: 0.785 ms against 0.651 ms per 2^24 pairs -- inside a loop ptxas hoists
// the coefficient loads, runs out of uniform registers and executes 1 160 instead of ~950 instructions per pair,
// and even behind a call the ring bookkeeping costs more than the 10 % of load-scoreboard stalls it removes
// (profiles/r01_summary.md).
template <bool HAS_RADIUS_ARR, bool WANT_TPQ>
__global__ void __launch_bounds__(PP_EVAL_THREADS, PP_EVAL_MIN_BLOCKS)
    pp_dubins_eval_kernel(size_t n, const double *__restrict__ sx, const double *__restrict__ sy,
                          const double *__restrict__ syaw, const double *__restrict__ ex,
                          const double *__restrict__ ey, const double *__restrict__ eyaw,
                          const double *__restrict__ radius_arr, double inv_radius, double *__restrict__ cost,
                          uint8_t *__restrict__ word, double *__restrict__ tpq) {
    size_t i = (size_t)blockIdx.x * PP_EVAL_THREADS + threadIdx.x;
    if (i >= n) return;
    // src/dubins.rs:404 `c = 1 / turn_radius`: for a scalar radius the host passes the quotient (same IEEE division)
    const double c = HAS_RADIUS_ARR ? 1.0 / __ldg(radius_arr + i) : inv_radius;
    pp_eval_one<WANT_TPQ>(__ldg(sx + i), __ldg(sy + i), __ldg(syaw + i), __ldg(ex + i), __ldg(ey + i), __ldg(eyaw + i),
                          c, i, cost, word, tpq);
}

// Two pairs per thread with both pairs' loads issued before the first evaluation (behind a call, or inlined twice) were
// measured in round 2 and removed: 0.636 / 0.623 / 0.619 ms against 0.616 ms (profiles/r02_summary.md section 4) -- while
// some warps wait on their loads, the others already saturate the issue port and the FP64 pipe together.
int pp_launch_dubins_eval(pp_ctx *ctx, size_t n, const double *sx, const double *sy, const double *syaw,
                          const double *ex, const double *ey, const double *eyaw, const double *radius_arr,
                          double radius, double *cost, uint8_t *word, double *tpq, cudaStream_t stream) {
    if (n == 0) return PP_OK;
    pp_launch_scope scope(ctx, "dubins_eval");
    const double inv_radius = 1.0 / radius;
    const unsigned grid = (unsigned)((n + PP_EVAL_THREADS - 1) / PP_EVAL_THREADS);
#define PP_GO(RA, TPQ)                                                                                          \
    pp_dubins_eval_kernel<RA, TPQ><<<grid, PP_EVAL_THREADS, 0, stream>>>(n, sx, sy, syaw, ex, ey, eyaw, radius_arr, \
                                                                         inv_radius, cost, word, tpq)
    if (radius_arr) {
        if (tpq) PP_GO(true, true); else PP_GO(true, false);
    } else {
        if (tpq) PP_GO(false, true); else PP_GO(false, false);
    }
#undef PP_GO
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

// diagnostic: all six words for explicit (alpha, beta, d)
__global__ void pp_dubins_words_kernel(size_t n, const double *__restrict__ alpha, const double *__restrict__ beta,
                                       const double *__restrict__ d, double *__restrict__ tpq,
                                       uint8_t *__restrict__ feas) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double all[18];
    uint8_t fz[6];
    pp_dubins_solve<true>(alpha[i], beta[i], d[i], all, fz);
    for (int k = 0; k < 18; ++k) tpq[18 * i + k] = all[k];
    for (int k = 0; k < 6; ++k) feas[6 * i + k] = fz[k];
}

int pp_launch_dubins_words(pp_ctx *ctx, size_t n, const double *alpha, const double *beta, const double *d,
                           double *tpq, uint8_t *feas, cudaStream_t stream) {
    if (n == 0) return PP_OK;
    pp_launch_scope scope(ctx, "dubins_words");
    pp_dubins_words_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(n, alpha, beta, d, tpq, feas);
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

__global__ void pp_mod2pi_kernel(size_t n, const double *__restrict__ x, double *__restrict__ out, int pi2pi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = pi2pi ? pp_pi_2_pi_fast(x[i]) : pp_mod2pi(x[i]);
}

int pp_launch_mod2pi(pp_ctx *ctx, size_t n, const double *x, double *out, int pi2pi, cudaStream_t stream) {
    if (n == 0) return PP_OK;
    pp_launch_scope scope(ctx, "mod2pi");
    pp_mod2pi_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, x, out, pi2pi);
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

// ------------------------------------------------------------------------------------------------
// plan / count: evaluate, then replay generate_local_course's index arithmetic (src/dubins.rs:200-289)
// without producing samples: per segment the first `pd`, the number of loop iterations obtained by
// the reference's own repeated addition (Q10), the carried remainder `ll`, and finally the trim rule
// (Q6/Q7) which needs the local x of the trailing slots.  One thread per path.  The literal replay is a
// dependent DADD/DSETP/BRA chain per sample (80 % of this kernel's instructions on C5 paths); pp_replay.cuh
// produces the same count and the same final `pd` by jumping binade by binade.
// ------------------------------------------------------------------------------------------------
#define PP_PLAN_MAX_ITERS (1u << 26)
#ifndef PP_PLAN_LITERAL_REPLAY
#define PP_PLAN_LITERAL_REPLAY 0  // A/B switch.  1: the serial `pd += d` loop (round 1 / early round 2)
#endif

// one path: evaluate + replay; (sx, sy, syaw) are ignored when from_origin
__device__ __forceinline__ pp_dubins_plan pp_make_plan(double sx, double sy, double syaw, double ex, double ey,
                                                       double eyaw, double radius, double step, int from_origin,
                                                       pp_plan_aux *aux = nullptr, double box_limit = 0.0,
                                                       double *box = nullptr) {
    pp_dubins_plan pl;
    if (from_origin) {  // the goal is given in the start frame: start pose = origin
        pl.sx = pl.sy = pl.syaw = 0.0;
    } else {
        pl.sx = sx;
        pl.sy = sy;
        pl.syaw = syaw;
    }
    double c = 1.0 / radius;
    pp_dubins_frame f = from_origin ? pp_dubins_frame_world(ex, ey, 0.0, eyaw, c, true)
                                    : pp_dubins_frame_world(ex - sx, ey - sy, syaw, eyaw - syaw, c);
    pp_dubins_sol s = pp_dubins_solve<false>(f.alpha, f.beta, f.d, nullptr, nullptr);
    pl.word = (uint8_t)s.word;
    pl.from_origin = (uint8_t)(from_origin != 0);
    pl.rinv = 1.0 / c;
    pl.step = step;
    pl.len[0] = s.t;
    pl.len[1] = s.p;
    pl.len[2] = s.q;
    pl.n[0] = pl.n[1] = pl.n[2] = 0;
    pl.pd0[0] = pl.pd0[1] = pl.pd0[2] = 0.0;
    pl.count = 0;
    for (int k = 0; k < 6; ++k) pl._pad[k] = 0;
    if (s.word != PP_WORD_NONE) {
        double ll = 0.0;
        bool overflow = false;
        for (int sgm = 0; sgm < 3; ++sgm) {
            double l = pl.len[sgm];
            double d = (l > 0.0) ? step : -step;  // :228
            double pd = (sgm >= 1 && (pl.len[sgm - 1] * l) > 0.0) ? (-d - ll) : (d - ll);  // :233-237
            pl.pd0[sgm] = pd;
            uint32_t cnt = 0;
            double al = fabs(l);
#if PP_PLAN_LITERAL_REPLAY
            while (fabs(pd) <= al) {  // :239
                pd += d;
                if (++cnt >= PP_PLAN_MAX_ITERS) {
                    overflow = true;
                    break;
                }
            }
#else
            // the same loop, bit for bit, in O(binades) steps (pp_replay.cuh)
            if (!pp_replay_segment(pd, d, al, PP_PLAN_MAX_ITERS, &cnt, &pd)) overflow = true;
#endif
            pl.n[sgm] = cnt;
            ll = (l - pd) - d;  // :256
        }
        if (overflow) {
            pl.count = 0xFFFFFFFFu;
        } else {
            // trim (:281-288): count = index of the last slot whose local x is non-zero
            uint32_t N = pl.n[0] + pl.n[1] + pl.n[2];
            pp_seg_origin o[3];
            double gx, gy;
            pp_segment_origins(pl, o, &gx, &gy);
            if (aux) {
                pp_sincos1(pl.syaw, &aux->ss, &aux->cs);
                // The path's bounding box for the path-level test (path_box.cuh).  That test can only use boxes below
                // `box_limit` in either extent (its cell-count caps, set by the launcher from the world's grids); the
                // start-to-goal offset is a lower bound of the extent, so long edges (C5) skip the computation.
                // A word with a zero-length segment gets no box either: there the reference's index arithmetic puts
                // samples up to five steps off a segment's ends, even BEHIND the start pose (l = 0 makes d negative
                // and the carried `ll` positive, src/dubins.rs:228-237), so "every point lies on the three segments"
                // does not hold.  With three positive lengths every first `pd` is the previous overshoot in (0, d].
                // (the reference branches on the PRODUCTS of neighbouring lengths, :233: tested as such, so that an
                // underflowing product of two tiny lengths counts as the zero-length case it would be treated as)
                if (box && fabs(ex - sx) < box_limit && fabs(ey - sy) < box_limit && pl.len[0] * pl.len[1] > 0.0 &&
                    pl.len[1] * pl.len[2] > 0.0 && pl.len[1] > 0.0)
                    pp_path_box(pl, o, aux->ss, aux->cs, gx, gy, from_origin ? pl.sx : ex, from_origin ? pl.sy : ey, box);
                aux->o[0] = o[0];
                aux->o[1] = o[1];
                aux->o[2] = o[2];
            }
            uint32_t cntout;
            if (gx != 0.0) {
                cntout = N + 1;
            } else {
                cntout = 0;
                for (uint32_t k = N; k >= 1; --k) {
                    double x, y, yaw;
                    pp_plan_sample_local(pl, o, k, &x, &y, &yaw);
                    if (x != 0.0) {
                        cntout = k;
                        break;
                    }
                }
            }
            pl.count = cntout;
        }
    }
    if (aux && (s.word == PP_WORD_NONE || pl.count == 0xFFFFFFFFu)) pp_sincos1(pl.syaw, &aux->ss, &aux->cs);
    return pl;
}

#ifndef PP_PLAN_MIN_BLOCKS
#define PP_PLAN_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(128, PP_PLAN_MIN_BLOCKS)
    pp_dubins_plan_kernel(size_t n, const double *__restrict__ sx, const double *__restrict__ sy,
                          const double *__restrict__ syaw, const double *__restrict__ ex,
                          const double *__restrict__ ey, const double *__restrict__ eyaw, double radius, double step,
                          int from_origin, uint32_t *__restrict__ counts, pp_dubins_plan *__restrict__ plans,
                          pp_plan_aux *__restrict__ aux_out, double box_limit, pp_world_view w,
                          uint8_t *__restrict__ ok, uint32_t *__restrict__ todo, unsigned int *__restrict__ todo_count) {
    // Records leave through shared memory: a thread's own 112 / 136-byte record written straight to global memory is
    // seven / seventeen warp stores that each touch 32 half-used sectors (ncu r03: lg_throttle on these stores was 31 %
    // of the kernel's stall samples for 5 % of its instructions).  A warp's 32 records are contiguous in global
    // memory, so they are parked per warp and copied out as coalesced 16-byte vectors.
    __shared__ __align__(16) unsigned char s_plan[128 / 32][32 * sizeof(pp_dubins_plan)];
    __shared__ __align__(16) unsigned char s_aux[128 / 32][32 * sizeof(pp_plan_aux)];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t first = i - (size_t)lane;  // first path of this warp
    if (first >= n) return;                 // warp-uniform
    const uint32_t live = (uint32_t)((n - first < 32) ? (n - first) : 32);
    bool verify = false;  // this path still needs the verify kernel (only meaningful with a todo list)
    if (i < n) {
        pp_plan_aux aux;
#pragma unroll
        for (int k = 0; k < 3; ++k) aux.o[k].ox = aux.o[k].oy = aux.o[k].oyaw = aux.o[k].so = aux.o[k].co = 0.0;
        pp_plan_aux *ap = aux_out ? &aux : nullptr;
        double box[4] = {CUDART_NAN, CUDART_NAN, CUDART_NAN, CUDART_NAN};
        const pp_dubins_plan pl =
            from_origin ? pp_make_plan(0.0, 0.0, 0.0, ex[i], ey[i], eyaw[i], radius, step, 1, ap, box_limit, todo ? box : nullptr)
                        : pp_make_plan(sx[i], sy[i], syaw[i], ex[i], ey[i], eyaw[i], radius, step, 0, ap, box_limit,
                                       todo ? box : nullptr);
        counts[i] = pl.count;
        if (todo) {
            // path-level test (path_box.cuh): a free path is answered here, the others go on the verify kernel's list
            verify = !pp_path_box_free_thread(w, box[0], box[1], box[2], box[3]);
            if (!verify) ok[i] = 1;
        }
        if (plans) *reinterpret_cast<pp_dubins_plan *>(&s_plan[wib][lane * sizeof(pp_dubins_plan)]) = pl;
        if (aux_out) {
            double *sa = reinterpret_cast<double *>(&s_aux[wib][lane * sizeof(pp_plan_aux)]);
            const double *a = reinterpret_cast<const double *>(&aux);
#pragma unroll
            for (int k = 0; k < PP_PLAN_AUX_DOUBLES; ++k) sa[k] = a[k];
        }
    }
    __syncwarp();
    if (todo) {  // append this warp's failing paths: one atomic per warp
        const unsigned need = __ballot_sync(0xffffffffu, verify);
        unsigned int base = 0;
        if (lane == 0 && need) base = atomicAdd(todo_count, (unsigned int)__popc(need));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (verify) todo[base + __popc(need & ((1u << lane) - 1u))] = (uint32_t)i;
    }
    if (plans) {  // sizeof(pp_dubins_plan) = 7 x 16 bytes
        const uint4 *src = reinterpret_cast<const uint4 *>(s_plan[wib]);
        uint4 *dst = reinterpret_cast<uint4 *>(plans + first);
        for (uint32_t v = lane; v < live * (uint32_t)(sizeof(pp_dubins_plan) / 16); v += 32) dst[v] = src[v];
    }
    if (aux_out) {  // 17 doubles per record; 32 records start 16-byte aligned, an odd number of them ends on a half vector
        const double2 *src = reinterpret_cast<const double2 *>(s_aux[wib]);
        double2 *dst = reinterpret_cast<double2 *>(aux_out + first);
        const uint32_t nd = live * PP_PLAN_AUX_DOUBLES;
        for (uint32_t v = lane; v < nd / 2; v += 32) dst[v] = src[v];
        if ((nd & 1u) && lane == 0)
            reinterpret_cast<double *>(aux_out + first)[nd - 1] = reinterpret_cast<const double *>(s_aux[wib])[nd - 1];
    }
}

// ------------------------------------------------------------------------------------------------
// scalar call (dubins_path_planning(&conf) of the reference, a batch of one): ONE launch, arguments by
// value, plan by thread 0, samples by the whole CTA straight into mapped pinned host memory -- no
// allocation, no explicit copy, one stream synchronisation on the host side.
// ------------------------------------------------------------------------------------------------
struct pp_path_header {
    uint32_t count, word;
    double cost, len[3];
};

__global__ void __launch_bounds__(128)
    pp_dubins_path_kernel(double sx, double sy, double syaw, double ex, double ey, double eyaw, double radius,
                          double step, int from_origin, uint32_t cap, pp_path_header *__restrict__ hdr,
                          double *__restrict__ out) {
    __shared__ pp_dubins_plan spl;
    if (threadIdx.x == 0) {
        spl = pp_make_plan(sx, sy, syaw, ex, ey, eyaw, radius, step, from_origin);
        hdr->count = spl.count;
        hdr->word = spl.word;
        hdr->len[0] = spl.len[0];
        hdr->len[1] = spl.len[1];
        hdr->len[2] = spl.len[2];
        hdr->cost = (spl.word == PP_WORD_NONE) ? CUDART_INF : (fabs(spl.len[0]) + fabs(spl.len[1])) + fabs(spl.len[2]);
    }
    __syncthreads();
    const pp_dubins_plan pl = spl;
    if (pl.word == PP_WORD_NONE || pl.count == 0 || pl.count == 0xFFFFFFFFu || pl.count > cap) return;
    double ss = 0.0, cs = 1.0;
    if (!pl.from_origin) pp_sincos1(pl.syaw, &ss, &cs);
    if (threadIdx.x == 0) {
        out[0] = pl.from_origin ? 0.0 : (cs * 0.0 + (-ss) * 0.0) + pl.sx;
        out[1] = pl.from_origin ? 0.0 : (ss * 0.0 + cs * 0.0) + pl.sy;
        out[2] = pl.from_origin ? 0.0 : pp_pi_2_pi_fast(0.0 + pl.syaw);
    }
    // samples exactly as pp_dubins_fill_kernel makes them (same chunk / lane split of the arc parameter, same device
    // functions): the scalar call and the batch return the same bits
    double ox = 0.0, oy = 0.0, oyaw = 0.0, so = 0.0, co = 1.0;
    uint32_t base = 1;
    double sB, cB;
    pp_sincos1((double)(threadIdx.x & 31u) * pl.step, &sB, &cB);
#pragma unroll
    for (int seg = 0; seg < 3; ++seg) {
        const int mode = pp_word_mode(pl.word, seg);
        const double len = pl.len[seg], pd0 = pl.pd0[seg];
        const double d = (len > 0.0) ? pl.step : -pl.step;
        const uint32_t ns = pl.n[seg];
        const pp_seg_world sw = pp_seg_world_make(ss, cs, pl.sx, pl.sy, ox, oy, so, co, pl.rinv, mode);
        const double sBd = (d > 0.0) ? sB : -sB;
        const bool right = mode == PP_MODE_R;
        for (uint32_t j = threadIdx.x; j < ns; j += blockDim.x) {
            const uint32_t k = base + j;
            if (k >= pl.count) break;
            const double jf = (double)j;
            double x, y, yaw;
            if (mode == PP_MODE_S) {
                pp_line_sample(sw, fma(jf, d, pd0), &x, &y);
                yaw = pl.from_origin ? oyaw : pp_pi_2_pi_fast(oyaw + pl.syaw);
            } else {
                const pp_arc_coef kc = pp_arc_coef_make(sw, pd0 + (double)(j & ~31u) * d);
                pp_arc_sample(sw, kc, sBd, cB, &x, &y);
                yaw = oyaw + fma(jf, right ? -d : d, right ? -pd0 : pd0);
                if (!pl.from_origin) yaw = pp_pi_2_pi_fast(yaw + pl.syaw);
            }
            out[3 * (size_t)k + 0] = x;
            out[3 * (size_t)k + 1] = y;
            out[3 * (size_t)k + 2] = yaw;
        }
        base += ns;
        if (seg < 2) {
            double ex_, ey_, eyaw_;
            pp_interpolate(mode, len, ox, oy, oyaw, so, co, pl.rinv, &ex_, &ey_, &eyaw_);
            ox = ex_;
            oy = ey_;
            oyaw = eyaw_;
            pp_sincos1(oyaw, &so, &co);
        }
    }
}

int pp_launch_dubins_path(pp_ctx *ctx, double sx, double sy, double syaw, double ex, double ey, double eyaw,
                          double radius, double step, int from_origin, uint32_t cap, void *hdr, double *out,
                          cudaStream_t stream) {
    pp_launch_scope scope(ctx, "dubins_path");
    pp_dubins_path_kernel<<<1, 128, 0, stream>>>(sx, sy, syaw, ex, ey, eyaw, radius, step, from_origin, cap,
                                                 (pp_path_header *)hdr, out);
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

#ifndef PP_PLAN_PATH_BOX
#define PP_PLAN_PATH_BOX 1  // A/B switch.  0: no path-level test, every path goes to the verify kernel
#endif
// ok / todo / todo_count != nullptr ("verify mode", with aux): the kernel answers the paths that pass the path-level test
// itself (ok[i] = 1) and appends the indices of the others to todo[0 .. *todo_count); *todo_count must be zero.
int pp_launch_dubins_plan(pp_ctx *ctx, size_t n, const double *sx, const double *sy, const double *syaw,
                          const double *ex, const double *ey, const double *eyaw, double radius, double step,
                          int from_origin, uint32_t *counts, void *plans, void *aux, cudaStream_t stream,
                          uint8_t *ok, uint32_t *todo, unsigned int *todo_count) {
    if (n == 0) return PP_OK;
    pp_launch_scope scope(ctx, "dubins_plan");
    // largest box extent the path-level test accepts (path_box.cuh -- at most 32 cells of the bounds grid,
    // PP_PATH_BOX_CELLS^2 cells of the obstacle grid); 0 = no box is computed
    double box_limit = 0.0;
    pp_world_view w{};
    if (todo && aux && ok && todo_count && ctx->world.valid && n < 0xFFFFFFFFull) {
        const pp_world_dev &wd = ctx->world;
        w = pp_make_world_view(wd);
#if PP_PLAN_PATH_BOX
        box_limit = std::min(32.0 / wd.binvx, 32.0 / wd.binvy);
        if (wd.n_rings) box_limit = std::min(box_limit, (PP_PATH_BOX_CELLS + 1.0) * wd.gcell);
        if (!(box_limit > 0.0)) box_limit = 0.0;
#endif
    } else {
        todo = nullptr;
    }
    pp_dubins_plan_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
        n, sx, sy, syaw, ex, ey, eyaw, radius, step, from_origin, counts, (pp_dubins_plan *)plans, (pp_plan_aux *)aux,
        box_limit, w, ok, todo, todo_count);
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

// ------------------------------------------------------------------------------------------------
// fill: one warp per path, lane k takes samples k, k+32, ...; each warp iteration writes 32
// consecutive (x,y,yaw) triples = 768 contiguous bytes.  HBM-store bound for straight segments,
// near the FP64 ridge for arcs (sincos per sample).
// ------------------------------------------------------------------------------------------------
#define PP_FILL_THREADS 128

#ifndef PP_FILL_LEGACY
#define PP_FILL_LEGACY 0  // A/B switch.  1: the round-1 kernel (one sincos per SAMPLE, local frame + per-sample rotation)
#endif
#if PP_FILL_LEGACY
// Variants measured on 5.1e7 samples (2^16 C5 paths) and NOT adopted (round 2, profiles/r02_summary.md): rows leaving
// through the TMA engine (cp.async.bulk shared -> global, two rows per warp; removed since) 0.378 ms; rows of
// 64 samples with two interpolations in flight per lane 0.366 ms; this kernel under __launch_bounds__(128, 6) 0.418 ms;
// this kernel as it stands 0.280 ms = 4.4 TB/s (a pure store stream, torch fill_, reaches 7.26 TB/s on the same box).
__global__ void __launch_bounds__(PP_FILL_THREADS)
    pp_dubins_fill_kernel(size_t n, const pp_dubins_plan *__restrict__ plans, const uint64_t *__restrict__ offsets,
                          double *__restrict__ out) {
    __shared__ __align__(16) double fill_stage[PP_FILL_THREADS / 32][96];
    const int lane = threadIdx.x & 31;
    double *stage = fill_stage[threadIdx.x >> 5];
    const size_t warps_total = (size_t)gridDim.x * (PP_FILL_THREADS / 32);
    for (size_t path = (size_t)blockIdx.x * (PP_FILL_THREADS / 32) + (threadIdx.x >> 5); path < n;
         path += warps_total) {
        const pp_dubins_plan pl = plans[path];
        if (pl.count == 0 || pl.count == 0xFFFFFFFFu) continue;
        double ss = 0.0, cs = 1.0;
        if (!pl.from_origin) pp_sincos1(pl.syaw, &ss, &cs);
        double *dst = out + 3 * offsets[path];
        if (lane == 0) {  // slot 0: the untouched zero of the reference's buffer = the start pose
            dst[0] = pl.from_origin ? 0.0 : (cs * 0.0 + (-ss) * 0.0) + pl.sx;
            dst[1] = pl.from_origin ? 0.0 : (ss * 0.0 + cs * 0.0) + pl.sy;
            dst[2] = pl.from_origin ? 0.0 : pp_pi_2_pi_fast(0.0 + pl.syaw);
        }
        // one loop per segment: origin, mode and first pd are loop invariants (no per-sample segment search)
        double ox = 0.0, oy = 0.0, oyaw = 0.0, so = 0.0, co = 1.0;
        uint32_t base = 1;
#pragma unroll
        for (int seg = 0; seg < 3; ++seg) {
            const int mode = pp_word_mode(pl.word, seg);
            const double len = pl.len[seg], pd0 = pl.pd0[seg];
            const double d = (len > 0.0) ? pl.step : -pl.step;
            const uint32_t ns = pl.n[seg];
            for (uint32_t j0 = 0; j0 < ns; j0 += 32) {
                const uint32_t k0 = base + j0;
                if (k0 >= pl.count) break;  // the trim rule may cut the tail (Q6/Q7); warp-uniform
                const uint32_t cnt = min(min(32u, ns - j0), pl.count - k0);
                if ((uint32_t)lane < cnt) {
                    double x, y, yaw;
                    pp_interpolate(mode, pd0 + (double)(j0 + lane) * d, ox, oy, oyaw, so, co, pl.rinv, &x, &y, &yaw);
                    if (!pl.from_origin) {  // src/dubins.rs:412-422
                        const double xw = (cs * x + (-ss) * y) + pl.sx;
                        const double yw = (ss * x + cs * y) + pl.sy;
                        x = xw;
                        y = yw;
                        yaw = pp_pi_2_pi_fast(yaw + pl.syaw);
                    }
                    stage[3 * lane + 0] = x;  // stride of 3 doubles: conflict-free
                    stage[3 * lane + 1] = y;
                    stage[3 * lane + 2] = yaw;
                }
                __syncwarp();
                // Per-lane (x,y,yaw) stores would touch 24 sectors per warp instruction (ncu: 3x the ideal
                // sector count, L1 79 % busy, lg_throttle the top stall).  The 3*cnt doubles of this step are
                // contiguous in `out`: one scalar store to reach 16-byte alignment if needed, then 16-byte
                // vector stores straight from the staging row (a full step = 768 B = 48 vectors).
                double *d0 = dst + 3 * (size_t)k0;
                const uint32_t D = 3 * cnt;
                const uint32_t head = (uint32_t)((reinterpret_cast<uintptr_t>(d0) >> 3) & 1u);
                const uint32_t nvec = (D - head) >> 1;
                if (lane == 0 && head) d0[0] = stage[0];
#pragma unroll
                for (uint32_t v = lane; v < 64; v += 32) {
                    if (v < nvec) {
                        const uint32_t e = head + 2 * v;
                        *reinterpret_cast<double2 *>(d0 + e) = make_double2(stage[e], stage[e + 1]);
                    }
                }
                if (lane == 0 && ((D - head) & 1u)) d0[D - 1] = stage[D - 1];
                __syncwarp();
            }
            base += ns;
            if (seg < 2) {  // next origin = this segment's end point (src/dubins.rs:258-271, read back at :230)
                double ex_, ey_, eyaw_;
                pp_interpolate(mode, len, ox, oy, oyaw, so, co, pl.rinv, &ex_, &ey_, &eyaw_);
                ox = ex_;
                oy = ey_;
                oyaw = eyaw_;
                pp_sincos1(oyaw, &so, &co);
            }
        }
    }
}
#else
// Default since round 2 (late): the same warp-per-path, 32-samples-per-row staging, but the samples are produced in
// the WORLD frame from per-segment constants (pp_seg_world) and, for arcs, per-chunk coefficients (pp_arc_coef):
// one sincos per 32-sample chunk -- computed 32 chunks at a time, one chunk per lane, parked in shared memory -- and
// one sincos per lane and path (lane * step) instead of one sincos plus two rotations per sample.  The round-1 kernel
// (PP_FILL_LEGACY=1) executed ~129 warp instructions per row and ran at 0.61 of a pure store stream because of them.
// The staging row is shifted by the parity of the row's first global element (constant per segment: rows are 96
// doubles), so both the shared-memory reads and the global stores of a row are 16-byte vectors; full rows -- all but
// the last of a segment -- take a path with no per-row index arithmetic.
#define PP_FILL_STAGE 100  // 96 doubles + the parity shift, padded to a 16-byte multiple
#ifndef PP_FILL_MIN_BLOCKS
#define PP_FILL_MIN_BLOCKS 4  // 120 registers
#endif

// one row of `cnt` samples (3 cnt doubles, contiguous in `out` from d0) leaves the staging row: the generic form for
// the last, partial row of a segment
__device__ __forceinline__ void pp_fill_store_partial(double *d0, const double *stage, const double2 *svec, uint32_t head,
                                                      uint32_t cnt, int lane) {
    const uint32_t D = 3 * cnt;
    const uint32_t nvec = (D - head) >> 1;
    double2 *gvec = reinterpret_cast<double2 *>(d0 + head);
    if (lane == 0 && head) d0[0] = stage[0];
#pragma unroll
    for (uint32_t v = lane; v < 64; v += 32)
        if (v < nvec) gvec[v] = svec[v];
    if (lane == 0 && ((D - head) & 1u)) d0[D - 1] = stage[D - 1];
}

//
// Work distribution: path lengths differ by orders of magnitude, so a static path -> warp map leaves the SMs idle at
// the end (and the short warps of a CTA idle all along).  The grid is exactly the resident CTAs and every warp draws
// `batch` consecutive paths at a time from a counter in the context (work[0]; work[1] counts finished CTAs, the last
// one re-arms both for the next launch on the stream).
__global__ void __launch_bounds__(PP_FILL_THREADS, PP_FILL_MIN_BLOCKS)
    pp_dubins_fill_kernel(uint32_t n, uint32_t batch, unsigned int *__restrict__ work,
                          const pp_dubins_plan *__restrict__ plans, const uint64_t *__restrict__ offsets,
                          double *__restrict__ out) {
    __shared__ __align__(16) double fill_stage[PP_FILL_THREADS / 32][PP_FILL_STAGE];
    __shared__ pp_arc_coef fill_coef[PP_FILL_THREADS / 32][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    pp_arc_coef *coef = fill_coef[wib];
    uint32_t path = 0, path_end = 0;
    for (;; ++path) {
        if (path >= path_end) {  // warp-uniform
            unsigned int first = 0;
            if (lane == 0) first = atomicAdd(&work[0], batch);
            first = __shfl_sync(0xffffffffu, first, 0);
            if (first >= n) break;
            path = first;
            path_end = min(n, first + batch);
        }
        const pp_dubins_plan pl = plans[path];
        if (pl.count == 0 || pl.count == 0xFFFFFFFFu) continue;
        double ss = 0.0, cs = 1.0;
        if (!pl.from_origin) pp_sincos1(pl.syaw, &ss, &cs);
        double *dst = out + 3 * offsets[path];
        if (lane == 0) {  // slot 0: the untouched zero of the reference's buffer = the start pose
            dst[0] = pl.from_origin ? 0.0 : (cs * 0.0 + (-ss) * 0.0) + pl.sx;
            dst[1] = pl.from_origin ? 0.0 : (ss * 0.0 + cs * 0.0) + pl.sy;
            dst[2] = pl.from_origin ? 0.0 : pp_pi_2_pi_fast(0.0 + pl.syaw);
        }
        double sB, cB;  // this lane's offset inside a chunk: lane * step
        pp_sincos1((double)lane * pl.step, &sB, &cB);
        double ox = 0.0, oy = 0.0, oyaw = 0.0, so = 0.0, co = 1.0;
        uint32_t base = 1;
#pragma unroll 1
        for (int seg = 0; seg < 3; ++seg) {
            const int mode = pp_word_mode(pl.word, seg);
            const double len = pl.len[seg], pd0 = pl.pd0[seg];
            const double d = (len > 0.0) ? pl.step : -pl.step;
            const uint32_t ns = pl.n[seg];
            // samples of this segment that survive the trim rule (Q6/Q7), as full rows of 32 + one partial row
            const uint32_t nseg = min(ns, pl.count > base ? pl.count - base : 0u);
            const uint32_t nfull = nseg >> 5, rem = nseg & 31u;
            if (nseg != 0u) {
                const pp_seg_world sw = pp_seg_world_make(ss, cs, pl.sx, pl.sy, ox, oy, so, co, pl.rinv, mode);
                double *d0 = dst + 3 * (size_t)base;
                // parity of the first global element of every row of this segment (a full row is 96 doubles)
                const uint32_t head = (uint32_t)((reinterpret_cast<uintptr_t>(d0) >> 3) & 1u);
                double *stage = fill_stage[wib] + head;  // element e of a row sits at stage[e]: e = head (mod 2) is 16-byte aligned
                const double2 *svec = reinterpret_cast<const double2 *>(fill_stage[wib]) + head;  // vector v = elements head + 2v, + 1
                double *mine = stage + 3 * lane;  // stride of 3 doubles: conflict-free
                double2 *gvec = reinterpret_cast<double2 *>(d0 + head) + lane;
                // a full row is 48 vectors (head = 0) or 1 + 47 vectors + 1: lane 0's vector pointer sits one element
                // after the row's first, lane 15's second-round pointer exactly on its last
                const bool second = lane < 16 - (int)head, first_el = head && lane == 0, last_el = head && lane == 15;
                double jf = (double)lane;  // sample index inside the segment, as a double (exact)
                if (mode == PP_MODE_S) {
                    // a straight segment keeps its origin's yaw (src/dubins.rs:166): one normalisation per segment
                    const double yaw_s = pl.from_origin ? oyaw : pp_pi_2_pi_fast(oyaw + pl.syaw);
                    for (uint32_t c = 0; c < nfull; ++c) {
                        double x, y;
                        pp_line_sample(sw, fma(jf, d, pd0), &x, &y);
                        mine[0] = x;
                        mine[1] = y;
                        mine[2] = yaw_s;
                        __syncwarp();
                        gvec[0] = svec[lane];
                        if (second) gvec[32] = svec[lane + 32];
                        if (first_el) reinterpret_cast<double *>(gvec)[-1] = reinterpret_cast<const double *>(svec + lane)[-1];
                        if (last_el) *reinterpret_cast<double *>(gvec + 32) = *reinterpret_cast<const double *>(svec + lane + 32);
                        __syncwarp();
                        gvec += 48;
                        jf += 32.0;
                    }
                    if (rem != 0u) {
                        if ((uint32_t)lane < rem) {
                            double x, y;
                            pp_line_sample(sw, fma(jf, d, pd0), &x, &y);
                            mine[0] = x;
                            mine[1] = y;
                            mine[2] = yaw_s;
                        }
                        __syncwarp();
                        pp_fill_store_partial(d0 + 96 * (size_t)nfull, stage, svec, head, rem, lane);
                        __syncwarp();
                    }
                } else {
                    const double sBd = (d > 0.0) ? sB : -sB;
                    // yaw of an arc sample: oyaw + pd (left) / oyaw - pd (right); the negation is exact, so the right
                    // turn walks (-pd0) + j (-d)
                    const bool right = mode == PP_MODE_R;
                    const double pdy0 = right ? -pd0 : pd0, dy = right ? -d : d;
                    for (uint32_t g0 = 0; g0 < nfull; g0 += 32) {
                        // coefficients of the next 32 chunks, one chunk per lane
                        __syncwarp();
                        coef[lane] = pp_arc_coef_make(sw, pd0 + (double)(32u * min(g0 + (uint32_t)lane, nfull)) * d);
                        __syncwarp();
                        const pp_arc_coef *cp = coef, *cend = coef + min(32u, nfull - g0);
                        do {
                            double x, y;
                            pp_arc_sample(sw, *cp, sBd, cB, &x, &y);
                            double yaw = oyaw + fma(jf, dy, pdy0);
                            if (!pl.from_origin) yaw = pp_pi_2_pi_fast(yaw + pl.syaw);
                            mine[0] = x;
                            mine[1] = y;
                            mine[2] = yaw;
                            __syncwarp();
                            gvec[0] = svec[lane];
                            if (second) gvec[32] = svec[lane + 32];
                            if (first_el) reinterpret_cast<double *>(gvec)[-1] = reinterpret_cast<const double *>(svec + lane)[-1];
                            if (last_el) *reinterpret_cast<double *>(gvec + 32) = *reinterpret_cast<const double *>(svec + lane + 32);
                            __syncwarp();
                            gvec += 48;
                            jf += 32.0;
                        } while (++cp != cend);
                    }
                    if (rem != 0u) {
                        if ((uint32_t)lane < rem) {
                            const pp_arc_coef kc = pp_arc_coef_make(sw, pd0 + (double)(32u * nfull) * d);
                            double x, y;
                            pp_arc_sample(sw, kc, sBd, cB, &x, &y);
                            double yaw = oyaw + fma(jf, dy, pdy0);
                            if (!pl.from_origin) yaw = pp_pi_2_pi_fast(yaw + pl.syaw);
                            mine[0] = x;
                            mine[1] = y;
                            mine[2] = yaw;
                        }
                        __syncwarp();
                        pp_fill_store_partial(d0 + 96 * (size_t)nfull, stage, svec, head, rem, lane);
                        __syncwarp();
                    }
                }
            }
            base += ns;
            if (seg < 2) {  // next origin = this segment's end point (src/dubins.rs:258-271, read back at :230)
                double ex_, ey_, eyaw_;
                pp_interpolate(mode, len, ox, oy, oyaw, so, co, pl.rinv, &ex_, &ey_, &eyaw_);
                ox = ex_;
                oy = ey_;
                oyaw = eyaw_;
                pp_sincos1(oyaw, &so, &co);
            }
        }
    }
    __syncthreads();  // every warp of this CTA has drawn its last batch
    if (threadIdx.x == 0 && atomicAdd(&work[1], 1u) == gridDim.x - 1u) {
        work[0] = 0u;
        work[1] = 0u;
    }
}
#endif

int pp_launch_dubins_fill(pp_ctx *ctx, size_t n, const void *plans, const uint64_t *offsets, double *out,
                          cudaStream_t stream) {
    if (n == 0) return PP_OK;
#if PP_FILL_LEGACY
    pp_launch_scope scope(ctx, "dubins_fill");
    size_t warps = n;
    size_t blocks = (warps + (PP_FILL_THREADS / 32) - 1) / (PP_FILL_THREADS / 32);
    size_t max_blocks = (size_t)ctx->sm_count * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    pp_dubins_fill_kernel<<<(unsigned)blocks, PP_FILL_THREADS, 0, stream>>>(n, (const pp_dubins_plan *)plans, offsets,
                                                                            out);
    PP_CUDA(ctx, cudaGetLastError());
#else
    if (ctx->fill_resident == 0) {  // CTAs of this kernel per SM, asked once per context (contexts are per device and
                                    // serialised by their mutex: no shared state between the workers of a pp_group)
        int r = 0;
        PP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, pp_dubins_fill_kernel, PP_FILL_THREADS, 0));
        ctx->fill_resident = r > 0 ? r : 1;
    }
    const size_t slots = (size_t)ctx->sm_count * ctx->fill_resident;  // CTAs that run side by side
    const size_t piece = (size_t)1 << 30;                    // the work counter is 32 bits wide
    for (size_t first = 0; first < n; first += piece) {
        const size_t m = std::min(piece, n - first);
        const size_t warps = slots * (PP_FILL_THREADS / 32);
        // ~8 draws per warp: fine enough to level the tail, coarse enough that 10^6 short paths do not queue on one
        // address
        const uint32_t batch = (uint32_t)std::min<size_t>(64, std::max<size_t>(1, m / (warps * 8)));
        const size_t blocks = std::min(slots, (m + (PP_FILL_THREADS / 32) - 1) / (PP_FILL_THREADS / 32));
        pp_launch_scope scope(ctx, "dubins_fill");
        pp_dubins_fill_kernel<<<(unsigned)blocks, PP_FILL_THREADS, 0, stream>>>(
            (uint32_t)m, batch, ctx->tickets + PP_TICKETS_FILL, (const pp_dubins_plan *)plans + first, offsets + first, out);
        PP_CUDA(ctx, cudaGetLastError());
    }
#endif
    return PP_OK;
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of u32 counts into u64 offsets: single-pass chained scan is overkill for this
// helper (n <= a few million, 12 B/elem): three small kernels (block sums, scan of sums, apply).
// ------------------------------------------------------------------------------------------------
#define PP_SCAN_THREADS 256
#define PP_SCAN_ITEMS 8
#define PP_SCAN_TILE (PP_SCAN_THREADS * PP_SCAN_ITEMS)

__device__ __forceinline__ uint64_t pp_block_exclusive_scan(uint64_t v, uint64_t *total, uint64_t *smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = (lane < PP_SCAN_THREADS / 32) ? smem[lane] : 0;
        uint64_t winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < PP_SCAN_THREADS / 32) smem[lane] = winc - w;
        if (lane == PP_SCAN_THREADS / 32 - 1) smem[32] = winc;
    }
    __syncthreads();
    uint64_t res = smem[warp] + inc - v;
    *total = smem[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(PP_SCAN_THREADS)
    pp_scan_tile_sums(size_t n, const uint32_t *__restrict__ counts, uint64_t *__restrict__ tile_sums) {
    __shared__ uint64_t sm[33];
    size_t base = (size_t)blockIdx.x * PP_SCAN_TILE + (size_t)threadIdx.x * PP_SCAN_ITEMS;
    uint64_t s = 0;
    for (int k = 0; k < PP_SCAN_ITEMS; ++k)
        if (base + k < n) {
            uint32_t c = counts[base + k];
            s += (c == 0xFFFFFFFFu) ? 0 : c;
        }
    uint64_t total;
    pp_block_exclusive_scan(s, &total, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(PP_SCAN_THREADS)
    pp_scan_of_sums(size_t n_tiles, uint64_t *__restrict__ tile_sums, uint64_t *__restrict__ total_out) {
    __shared__ uint64_t sm[33];
    uint64_t carry = 0;
    for (size_t base = 0; base < n_tiles; base += PP_SCAN_THREADS) {
        size_t i = base + threadIdx.x;
        uint64_t v = (i < n_tiles) ? tile_sums[i] : 0;
        uint64_t total;
        uint64_t e = pp_block_exclusive_scan(v, &total, sm);
        if (i < n_tiles) tile_sums[i] = carry + e;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(PP_SCAN_THREADS)
    pp_scan_apply(size_t n, const uint32_t *__restrict__ counts, const uint64_t *__restrict__ tile_sums,
                  uint64_t *__restrict__ offsets) {
    __shared__ uint64_t sm[33];
    size_t base = (size_t)blockIdx.x * PP_SCAN_TILE + (size_t)threadIdx.x * PP_SCAN_ITEMS;
    uint32_t c[PP_SCAN_ITEMS];
    uint64_t s = 0;
    for (int k = 0; k < PP_SCAN_ITEMS; ++k) {
        uint32_t v = (base + k < n) ? counts[base + k] : 0;
        c[k] = (v == 0xFFFFFFFFu) ? 0 : v;
        s += c[k];
    }
    uint64_t total;
    uint64_t e = pp_block_exclusive_scan(s, &total, sm) + tile_sums[blockIdx.x];
    for (int k = 0; k < PP_SCAN_ITEMS; ++k) {
        if (base + k < n) offsets[base + k] = e;
        e += c[k];
    }
}

int pp_launch_exclusive_scan(pp_ctx *ctx, size_t n, const uint32_t *counts, uint64_t *offsets, uint64_t *total_dev,
                             uint64_t *tile_sums /* >= ceil(n/tile) */, cudaStream_t stream) {
    if (n == 0) {
        PP_CUDA(ctx, cudaMemsetAsync(total_dev, 0, sizeof(uint64_t), stream));
        return PP_OK;
    }
    size_t tiles = (n + PP_SCAN_TILE - 1) / PP_SCAN_TILE;
    pp_launch_scope scope(ctx, "exclusive_scan", 3);
    pp_scan_tile_sums<<<(unsigned)tiles, PP_SCAN_THREADS, 0, stream>>>(n, counts, tile_sums);
    pp_scan_of_sums<<<1, PP_SCAN_THREADS, 0, stream>>>(tiles, tile_sums, total_dev);
    pp_scan_apply<<<(unsigned)tiles, PP_SCAN_THREADS, 0, stream>>>(n, counts, tile_sums, offsets);
    PP_CUDA(ctx, cudaGetLastError());
    return PP_OK;
}

// uploads this translation unit's copy of the math coefficient tables (pp_math.cuh) to the current device
int pp_dubins_tu_init(pp_ctx *ctx) {
    PP_CUDA(ctx, pp_math_upload_tables());
    return PP_OK;
}

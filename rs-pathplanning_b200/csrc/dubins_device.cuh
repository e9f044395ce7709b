// dubins_device.cuh -- device-side Dubins mathematics (f64), shared by the evaluate, sample and
// fused sample-and-verify kernels.
//
// Restates the arithmetic of src/dubins.rs (reference crate root) for one pose pair per thread.
// Design notes (not a translation of the Rust control flow):
//   * the six words are evaluated branch-free from ONE set of trig values (the reference
//     recomputes sin/cos of alpha, beta five times per word, src/dubins.rs:28-32 etc.);
//     identical sub-expressions are shared only where they are bit-identical
//     (atan2 of RSR == RLR's, LRL's == -LSL's since atan2 is odd in y).
//   * mod2pi (src/dubins.rs:14-20) is x - 2pi*floor(x/2pi).  The division is replaced by a
//     reciprocal multiply plus an exact guard: when the quotient is within 1e-9 of an integer
//     the IEEE division is redone, so the result is bit-identical to the reference form for
//     every input (Q1/Q3 wrap behaviour included).  Where the argument range is known
//     ([0,2pi], [pi,2pi] ...) the floor is resolved by a compare.
//   * this translation unit is compiled with -fmad=false: outside the CUDA libm calls no
//     multiply-add is fused, as in rustc's output.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/pathplanning_b200.h"
#include "pp_math.cuh"

#define PP_PI 3.14159265358979323846
#define PP_TWO_PI 6.28318530717958647692   // == 2.0 * PI exactly in binary64
#define PP_INV_TWO_PI 0.15915494309189533577

// x - 2pi*floor(x/2pi), bit-exact with the division form (src/dubins.rs:14-20)
__device__ __forceinline__ double pp_mod2pi(double x) {
    double q = x * PP_INV_TWO_PI;
    double k = floor(q);
    double f = q - k;
    // q is within 3 ulp of x/2pi; floor can only differ when q sits next to an integer
    if (!(fabs(f - 0.5) < 0.5 - 1e-9) || !(fabs(q) < 1e5)) k = floor(x / PP_TWO_PI);
    return x - PP_TWO_PI * k;
}
// mod2pi for x already known to lie in [0, 2pi]: only x == 2pi wraps (to 0)
// (bit-pattern compare, integer pipe: 2 pi = 0x401921FB54442D18)
__device__ __forceinline__ double pp_mod2pi_unit(double x) { return (pp_bits(x) == 0x401921FB54442D18LL) ? 0.0 : x; }

// src/dubins.rs:22-24 ; Rust % == fmod
__device__ __forceinline__ double pp_pi_2_pi(double a) { return fmod(a + PP_PI, PP_TWO_PI) - PP_PI; }

// Same value, bit for bit, without the library fmod: fmod's result is exactly representable, so ONE fma
// reproduces it when the integer quotient is right, and a quotient that is off by one (product rounding next
// to a multiple of 2pi) is repaired by redoing the fma with the neighbouring integer.
// b = a + pi with |b| >= 2 pi (or NaN): the quotient route.  Out of line: the callers' common case is the one-line
// test in pp_pi_2_pi_fast.
static __device__ __noinline__ double pp_pi_2_pi_wrap(double b) {
    const double ab = fabs(b);
    if (!(ab < 1e15)) return fmod(b, PP_TWO_PI) - PP_PI;  // huge, infinite or NaN
    double k = floor(ab * PP_INV_TWO_PI);
    double r = fma(-k, PP_TWO_PI, ab);
    if (r < 0.0) {
        k -= 1.0;
        r = fma(-k, PP_TWO_PI, ab);
    } else if (r >= PP_TWO_PI) {
        k += 1.0;
        r = fma(-k, PP_TWO_PI, ab);
    }
    return copysign(r, b) - PP_PI;
}
__device__ __forceinline__ double pp_pi_2_pi_fast(double a) {
    const double b = a + PP_PI;
    if (fabs(b) < PP_TWO_PI) return b - PP_PI;  // fmod(b, 2 pi) == b
    return pp_pi_2_pi_wrap(b);
}

struct pp_dubins_sol {
    double t, p, q, cost;
    int word;  // pp_word or PP_WORD_NONE
};

struct pp_dubins_frame {
    double alpha, beta, d;
};

// src/dubins.rs:402-408 + 333-338 : d, theta, alpha, beta of the goal as seen from the start pose.
// The reference rotates the goal offset into the start frame (sincos(syaw)) and takes atan2 / hypot of the rotated
// vector.  A rotation changes neither the norm nor -- up to the subtraction -- the polar angle:
// theta = atan2(dy, dx) - syaw, which saves the sincos and the rotation (~30 FP64 instructions per pair; a few ulp
// on theta, the class of difference the 1e-9 contract already covers).  Start yaws beyond +-64 rad (the subtraction
// would lose digits that the reference's exact range reduction keeps) and coincident positions (atan2 of signed
// zeros) take the reference's route.  Every kernel (evaluate, plan, scalar path) goes through this one function, so
// they agree on the chosen word.
// local_input: (dx, dy, leyaw) are already in the start frame (dubins_path_planning_from_origin, syaw = 0): the
// reference does not rotate at all, so neither does this.
__device__ __forceinline__ pp_dubins_frame pp_dubins_frame_world(double dx, double dy, double syaw, double leyaw,
                                                                 double c, bool local_input = false) {
    double ay_ = dy, ax_ = dx, sub = syaw;
    if (!local_input && !(pp_abs_below_pow2(syaw, 6) && (dx != 0.0 || dy != 0.0))) {
        double ss, cs;
#ifdef PP_DUBINS_LIBM
        sincos(syaw, &ss, &cs);
#else
        pp_sincos1(syaw, &ss, &cs);
#endif
        ax_ = cs * dx + ss * dy;
        ay_ = -ss * dx + cs * dy;
        sub = 0.0;
    }
    pp_dubins_frame f;
#ifdef PP_DUBINS_LIBM
    f.d = hypot(ax_, ay_) * c;
    const double a = atan2(ay_, ax_) - sub;
#else
    f.d = pp_sqrt_pos(fma(ax_, ax_, ay_ * ay_)) * c;  // world-scale coordinates: no overflow guard needed (<= 1 ulp)
    const double a = pp_atan2(ay_, ax_) - sub;
#endif
    const double theta = pp_mod2pi(a);  // [0, 2pi]; for the rotated route a is in [-pi, pi] and -0 becomes +0
    // alpha = mod2pi(-theta), theta in [0, 2pi]: floor(-theta/2pi) is -1 unless theta == 0
    f.alpha = (theta > 0.0) ? (PP_TWO_PI - theta) : ((theta == 0.0) ? 0.0 : theta /*NaN*/);
    f.beta = pp_mod2pi(leyaw - theta);
    return f;
}

// the six words (src/dubins.rs:27-153) + the selection fold (src/dubins.rs:347-363).
// WANT_ALL: also store every word's (t,p,q) / feasibility (diagnostic entry pp_dubins_words).
template <bool WANT_ALL>
__device__ __forceinline__ pp_dubins_sol pp_dubins_solve_libm(double alpha, double beta, double d, double *all_tpq,
                                                              uint8_t *all_feas) {
    double sa, ca, sb, cb;
    sincos(alpha, &sa, &ca);
    sincos(beta, &sb, &cb);
    const double c_ab = cos(alpha - beta);
    const double dd = d * d;
    const double two_cab = 2.0 * c_ab;
    const double two_d = 2.0 * d;
    const double mbeta = pp_mod2pi_unit(beta);  // mod2pi(beta), beta already in [0, 2pi]

    pp_dubins_sol best;
    best.cost = CUDART_INF;
    best.word = PP_WORD_NONE;
    best.t = best.p = best.q = CUDART_NAN;

#define PP_CONSIDER(W, FEAS, T, P, Q)                              \
    do {                                                           \
        double _c = (fabs(T) + fabs(P)) + fabs(Q);                 \
        bool _f = (FEAS);                                          \
        if (WANT_ALL) {                                            \
            all_feas[W] = _f ? 1 : 0;                              \
            all_tpq[3 * W + 0] = _f ? (T) : CUDART_NAN;            \
            all_tpq[3 * W + 1] = _f ? (P) : CUDART_NAN;            \
            all_tpq[3 * W + 2] = _f ? (Q) : CUDART_NAN;            \
        }                                                          \
        if (_f && _c < best.cost) { /* strict: first word wins */ \
            best.cost = _c;                                        \
            best.word = W;                                         \
            best.t = (T);                                          \
            best.p = (P);                                          \
            best.q = (Q);                                          \
        }                                                          \
    } while (0)

    // ---- LSL (src/dubins.rs:27-48) and the shared atan2 of LRL
    const double sa_m_sb = sa - sb;
    const double base_csc = (2.0 + dd) - two_cab;
    const double at_lsl = atan2(cb - ca, (d + sa) - sb);
    {
        double psq = base_csc + two_d * sa_m_sb;
        double t = pp_mod2pi(-alpha + at_lsl);
        double p = sqrt(fmax(psq, 0.0));
        double q = pp_mod2pi(beta - at_lsl);
        PP_CONSIDER(PP_LSL, !(psq < 0.0), t, p, q);
    }
    // ---- RSR (src/dubins.rs:51-71) and the shared atan2 of RLR
    const double at_rsr = atan2(ca - cb, (d - sa) + sb);
    {
        double psq = base_csc + two_d * (sb - sa);
        double t = pp_mod2pi(alpha - at_rsr);
        double p = sqrt(fmax(psq, 0.0));
        double q = pp_mod2pi(-beta + at_rsr);
        PP_CONSIDER(PP_RSR, !(psq < 0.0), t, p, q);
    }
    // ---- LSR (src/dubins.rs:74-92)
    const double sa_p_sb = sa + sb;
    const double base_cross = (-2.0 + dd) + two_cab;
    {
        double psq = base_cross + two_d * sa_p_sb;
        double p = sqrt(fmax(psq, 0.0));
        double tmp = atan2(-ca - cb, (d + sa) + sb) - atan2(-2.0, p);
        double t = pp_mod2pi(-alpha + tmp);
        double q = pp_mod2pi(-mbeta + tmp);
        PP_CONSIDER(PP_LSR, !(psq < 0.0), t, p, q);
    }
    // ---- RSL (src/dubins.rs:95-113)
    {
        double psq = base_cross - two_d * sa_p_sb;
        double p = sqrt(fmax(psq, 0.0));
        double tmp = atan2(ca + cb, (d - sa) - sb) - atan2(2.0, p);
        double t = pp_mod2pi(alpha - tmp);
        double q = pp_mod2pi(beta - tmp);
        PP_CONSIDER(PP_RSL, !(psq < 0.0), t, p, q);
    }
    // ---- RLR (src/dubins.rs:116-133): p in [pi,2pi] -> mod2pi(p/2) = p/2 and mod2pi(p) = p exactly
    const double base_ccc = (6.0 - dd) + two_cab;
    {
        double tmp = (base_ccc + two_d * sa_m_sb) * 0.125;
        bool feas = !(fabs(tmp) > 1.0);
        double ac = acos(fmin(fmax(tmp, -1.0), 1.0));
        double p = pp_mod2pi_unit(PP_TWO_PI - ac);
        double t = pp_mod2pi((alpha - at_rsr) + p * 0.5);
        double q = pp_mod2pi(((alpha - beta) - t) + p);
        PP_CONSIDER(PP_RLR, feas, t, p, q);
    }
    // ---- LRL (src/dubins.rs:136-153): atan2(ca-cb, d+sa-sb) == -atan2(cb-ca, d+sa-sb)
    {
        double tmp = (base_ccc + two_d * (sb - sa)) * 0.125;
        bool feas = !(fabs(tmp) > 1.0);
        double ac = acos(fmin(fmax(tmp, -1.0), 1.0));
        double p = pp_mod2pi_unit(PP_TWO_PI - ac);
        double t = pp_mod2pi((-alpha + at_lsl) + p * 0.5);
        double q = pp_mod2pi(((mbeta - alpha) - t) + p);
        PP_CONSIDER(PP_LRL, feas, t, p, q);
    }
#undef PP_CONSIDER
    return best;
}


// N-way mod2pi (bit-exact like pp_mod2pi): the N reductions are interleaved and share ONE guard branch.
// BOUNDED: the caller guarantees |x| < 1e5 (or NaN), so the large-quotient guard is not needed.
template <int N, bool BOUNDED = false>
__device__ __forceinline__ void pp_mod2pi_n(double (&x)[N]) {
    // The guard looks at the RESULT: with the right floor the value is bit-identical to the reference form and
    // lies in [0, 2pi]; a floor that is off by one (quotient rounding next to an integer) puts it outside, and a
    // result within 1e-9 of either end is re-derived with the IEEE division as well.
    double r[N];
    bool slow = false;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double q = x[i] * PP_INV_TWO_PI;
        r[i] = x[i] - PP_TWO_PI * floor(q);
        // 1e-9 < r < 2pi - 1.3e-6, tested on the high word with integer instructions (pp_math.cuh: the FP64
        // pipe is the bound); negative values and NaN fall outside the unsigned window as well
        slow |= !((pp_hi32(r[i]) - 0x3E112E0Cu) < (0x401921FBu - 0x3E112E0Cu));
        if (!BOUNDED) slow |= !pp_abs_below_pow2(q, 16);
    }
    if (slow) {
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = x[i] - PP_TWO_PI * floor(x[i] / PP_TWO_PI);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = r[i];
}

// The six words + selection, written for the FP64 pipe (default).  Differences from the libm version:
//   * sincos(alpha), sincos(beta) as one 2-way batch of the hand-written kernel; cos(alpha-beta) from the
//     angle-difference identity (3 DP instructions instead of a third trig call)
//   * LSR / RSL: atan2(y1,x1) - atan2(y2,x2) is folded into ONE atan2(y1 x2 - x1 y2, x1 x2 + y1 y2); the
//     result differs by a multiple of 2 pi, which the following mod2pi removes
//   * the two acos of the CCC words as one 2-way batch of the direct asin-polynomial kernel (no division)
//   * the four atan2 and the twelve mod2pi run as interleaved batches (coefficients fetched once, 4-10
//     independent dependency chains in flight)
// All of it stays inside the 1e-9 contract; t/q agree with the reference to a few ulp away from the wrap.
template <bool WANT_ALL>
__device__ __forceinline__ pp_dubins_sol pp_dubins_solve(double alpha, double beta, double d, double *all_tpq,
                                                         uint8_t *all_feas) {
    double sn[2], cs[2];
    {
        const double ang[2] = {alpha, beta};
        pp_sincos_n<2>(ang, sn, cs);
    }
    const double sa = sn[0], sb = sn[1], ca = cs[0], cb = cs[1];
    const double c_ab = fma(ca, cb, sa * sb);
    const double dd = d * d;
    const double two_cab = 2.0 * c_ab;
    const double two_d = 2.0 * d;
    const double mbeta = pp_mod2pi_unit(beta);

    const double sa_m_sb = sa - sb, sa_p_sb = sa + sb;
    const double base_csc = (2.0 + dd) - two_cab;
    const double base_cross = (-2.0 + dd) + two_cab;
    const double base_ccc = (6.0 - dd) + two_cab;
    const double psq_lsl = base_csc + two_d * sa_m_sb;
    const double psq_rsr = base_csc - two_d * sa_m_sb;
    const double psq_lsr = base_cross + two_d * sa_p_sb;
    const double psq_rsl = base_cross - two_d * sa_p_sb;
    const double tmp_rlr = (base_ccc + two_d * sa_m_sb) * 0.125;
    const double tmp_lrl = (base_ccc - two_d * sa_m_sb) * 0.125;
    const bool f_lsl = !(psq_lsl < 0.0), f_rsr = !(psq_rsr < 0.0), f_lsr = !(psq_lsr < 0.0), f_rsl = !(psq_rsl < 0.0);
    const bool f_rlr = !(fabs(tmp_rlr) > 1.0), f_lrl = !(fabs(tmp_lrl) > 1.0);
    // infeasible words get harmless stand-in arguments (masked at the selection): a clamp to 0 / +-1 would
    // send sqrt(0) and 0/x down the IEEE slow paths of sqrt and division in nearly every warp
    const double v_rlr = f_rlr ? tmp_rlr : 0.5, v_lrl = f_lrl ? tmp_lrl : 0.5;

    const double p_lsl = pp_sqrt_pos(f_lsl ? psq_lsl : 1.0), p_rsr = pp_sqrt_pos(f_rsr ? psq_rsr : 1.0);
    const double p_lsr = pp_sqrt_pos(f_lsr ? psq_lsr : 1.0), p_rsl = pp_sqrt_pos(f_rsl ? psq_rsl : 1.0);

    // ---- four atan2 in one batch, the two acos of the CCC words in another
    double ay[4], ax[4], at[4];
    ay[0] = cb - ca;  // LSL (and LRL, negated)
    ax[0] = (d + sa) - sb;
    ay[1] = ca - cb;  // RSR (and RLR)
    ax[1] = (d - sa) + sb;
    {   // LSR: atan2(-ca-cb, d+sa+sb) - atan2(-2, p)
        const double y1 = -ca - cb, x1 = (d + sa) + sb;
        ay[2] = fma(y1, p_lsr, 2.0 * x1);
        ax[2] = fma(x1, p_lsr, -2.0 * y1);
    }
    {   // RSL: atan2(ca+cb, d-sa-sb) - atan2(2, p)
        const double y1 = ca + cb, x1 = (d - sa) - sb;
        ay[3] = fma(y1, p_rsl, -2.0 * x1);
        ax[3] = fma(x1, p_rsl, 2.0 * y1);
    }
    pp_atan2_n<4>(ay, ax, at);
    double ac[2];
    {
        const double vv[2] = {v_rlr, v_lrl};
        pp_acos_n<2>(vv, ac);
    }

    const double pc_rlr = pp_mod2pi_unit(PP_TWO_PI - ac[0]);  // p of RLR, in [pi, 2pi] -> only 2pi wraps
    const double pc_lrl = pp_mod2pi_unit(PP_TWO_PI - ac[1]);

    // ---- twelve mod2pi in two batches (q of the CCC words needs their t)
    double m[10];
    m[0] = -alpha + at[0];                     // LSL t
    m[1] = beta - at[0];                       // LSL q
    m[2] = alpha - at[1];                      // RSR t
    m[3] = -beta + at[1];                      // RSR q
    m[4] = -alpha + at[2];                     // LSR t
    m[5] = -mbeta + at[2];                     // LSR q
    m[6] = alpha - at[3];                      // RSL t
    m[7] = beta - at[3];                       // RSL q
    m[8] = (alpha - at[1]) + pc_rlr * 0.5;     // RLR t
    m[9] = (-alpha + at[0]) + pc_lrl * 0.5;    // LRL t   (atan2(ca-cb, d+sa-sb) = -at[0])
    pp_mod2pi_n<10, true>(m);  // sums of a few angles in [-2pi, 2pi]
    double m2[2];
    m2[0] = ((alpha - beta) - m[8]) + pc_rlr;  // RLR q
    m2[1] = ((mbeta - alpha) - m[9]) + pc_lrl; // LRL q
    pp_mod2pi_n<2, true>(m2);

    pp_dubins_sol best;
    best.cost = CUDART_INF;
    best.word = PP_WORD_NONE;
    best.t = best.p = best.q = CUDART_NAN;
#define PP_CONSIDER2(W, FEAS, T, P, Q)                    \
    do {                                                  \
        const double _c = (fabs(T) + fabs(P)) + fabs(Q);  \
        const bool _f = (FEAS);                           \
        if (WANT_ALL) {                                   \
            all_feas[W] = _f ? 1 : 0;                     \
            all_tpq[3 * W + 0] = _f ? (T) : CUDART_NAN;   \
            all_tpq[3 * W + 1] = _f ? (P) : CUDART_NAN;   \
            all_tpq[3 * W + 2] = _f ? (Q) : CUDART_NAN;   \
        }                                                 \
        /* costs are sums of |.|: non-negative or NaN, so the unsigned bit patterns order like the */ \
        /* values (NaN above +inf, never selected) and the compare stays off the FP64 pipe */          \
        if (_f && (unsigned long long)pp_bits(_c) < (unsigned long long)pp_bits(best.cost)) { \
            best.cost = _c;                               \
            best.word = W;                                \
            best.t = (T);                                 \
            best.p = (P);                                 \
            best.q = (Q);                                 \
        }                                                 \
    } while (0)
    PP_CONSIDER2(PP_LSL, f_lsl, m[0], p_lsl, m[1]);
    PP_CONSIDER2(PP_RSR, f_rsr, m[2], p_rsr, m[3]);
    PP_CONSIDER2(PP_LSR, f_lsr, m[4], p_lsr, m[5]);
    PP_CONSIDER2(PP_RSL, f_rsl, m[6], p_rsl, m[7]);
    PP_CONSIDER2(PP_RLR, f_rlr, m[8], pc_rlr, m2[0]);
    PP_CONSIDER2(PP_LRL, f_lrl, m[9], pc_lrl, m2[1]);
#undef PP_CONSIDER2
    return best;
}

// segment modes per word (src/dubins.rs:26,50,73,94,115,135): 2 bits per segment, 0 L / 1 S / 2 R
__device__ __forceinline__ int pp_word_mode(int word, int seg) {
    // LSL 0,1,0  RSR 2,1,2  LSR 0,1,2  RSL 2,1,0  RLR 2,0,2  LRL 0,2,0
    const uint32_t table = (0u | 1u << 2 | 0u << 4) | ((2u | 1u << 2 | 2u << 4) << 6) | ((0u | 1u << 2 | 2u << 4) << 12) |
                           ((2u | 1u << 2 | 0u << 4) << 18) | ((2u | 0u << 2 | 2u << 4) << 24);
    if (word == PP_LRL) return (seg == 1) ? PP_MODE_R : PP_MODE_L;
    return (int)((table >> (6 * word + 2 * seg)) & 3u);
}

// src/dubins.rs:155-198 with the per-segment constants hoisted: (so, co) = sincos(origin_yaw), rinv = 1/c.
// (the reference divides by c per sample; multiplying by 1/c differs by <= 1 ulp, inside the 1e-9 contract)
__device__ __forceinline__ void pp_interpolate(int mode, double len, double ox, double oy, double oyaw, double so,
                                               double co, double rinv, double *x, double *y, double *yaw) {
    if (mode == PP_MODE_S) {
        double l = len * rinv;
        *x = ox + l * co;
        *y = oy + l * so;
        *yaw = oyaw;
    } else {
        double sl, cl;
#ifdef PP_INTERP_LIBM
        sincos(len, &sl, &cl);
#else
        pp_sincos1(len, &sl, &cl);
#endif
        double ldx = sl * rinv;
        double ldy = (1.0 - cl) * rinv;
        if (mode == PP_MODE_R) ldy = -ldy;
        // cos(-o) = co, sin(-o) = -so
        double gdx = co * ldx + (-so) * ldy;
        double gdy = so * ldx + co * ldy;
        *x = ox + gdx;
        *y = oy + gdy;
        *yaw = (mode == PP_MODE_L) ? (oyaw + len) : (oyaw - len);
    }
}

// per-path plan record shared by the count / fill / collide kernels (PP_DUBINS_PLAN_BYTES)
struct __align__(16) pp_dubins_plan {
    double len[3];   // t, p, q
    double pd0[3];   // first `pd` of each segment's loop (src/dubins.rs:233-237)
    double sx, sy, syaw;
    uint32_t n[3];   // loop iterations per segment (src/dubins.rs:239-255)
    uint32_t count;  // samples the reference returns after its trim loop (src/dubins.rs:281-288)
    double rinv;     // turn radius (1/c)
    double step;
    uint8_t word;    // pp_word or PP_WORD_NONE
    uint8_t from_origin;
    uint8_t _pad[6];
};
static_assert(sizeof(pp_dubins_plan) == PP_DUBINS_PLAN_BYTES, "plan record size is part of the ABI");

struct pp_seg_origin {
    double ox, oy, oyaw, so, co;
};

// What the fused sample-and-verify kernel needs besides the plan record.  The plan kernel has it anyway (one
// thread per path); recomputing it in the verify kernel costs every WARP four sincos and three interpolations per
// path, ~340 warp-instructions of the ~1 900 a short extend-step edge takes.
struct pp_plan_aux {
    pp_seg_origin o[3];
    double ss, cs;  // sincos(syaw): local -> world
};
#define PP_PLAN_AUX_DOUBLES 17
static_assert(sizeof(pp_plan_aux) == PP_PLAN_AUX_DOUBLES * 8, "aux record is copied as doubles");

// origins of the three segments: segment i starts at segment i-1's end point, written by
// interpolate(ind, l) at src/dubins.rs:258-271 and read back at :230.
__device__ __forceinline__ void pp_segment_origins(const pp_dubins_plan &pl, pp_seg_origin o[3], double *gx,
                                                   double *gy = nullptr) {
    o[0].ox = 0.0;
    o[0].oy = 0.0;
    o[0].oyaw = 0.0;
    o[0].so = 0.0;
    o[0].co = 1.0;
    for (int i = 0; i < 3; ++i) {
        double x, y, yaw;
        pp_interpolate(pp_word_mode(pl.word, i), pl.len[i], o[i].ox, o[i].oy, o[i].oyaw, o[i].so, o[i].co, pl.rinv,
                       &x, &y, &yaw);
        if (i < 2) {
            o[i + 1].ox = x;
            o[i + 1].oy = y;
            o[i + 1].oyaw = yaw;
            pp_sincos1(yaw, &o[i + 1].so, &o[i + 1].co);
        } else {
            *gx = x;
            if (gy) *gy = y;
        }
    }
}

// Axis-aligned box (world frame) around EVERY point the sampled path can produce: the start pose, the three segment
// origins, the end point, the parent point, and for each arc the axis extremes of its circle that fall inside the swept
// angle.  Valid for words with three positive lengths only (the caller checks).  The plan kernel hands it to the
// path-level test (path_box.cuh), which dismisses a whole path with one look at the rings registered under the box.
// Conservative by construction: extremes are included with a 1e-6 rad margin and the box is padded by 1e-9 of the
// coordinates' scale, orders of magnitude above the ulp-level differences between this arithmetic and the samples'.
// (ss, cs) = sincos(pl.syaw); (gx, gy) = local end point of the third segment; (ex, ey) = the goal as given (the
// verify kernel appends it to the samples as the parent point).
__device__ __forceinline__ void pp_path_box(const pp_dubins_plan &pl, const pp_seg_origin o[3], double ss, double cs,
                                            double gx, double gy, double ex, double ey, double box[4]) {
    double minx = (ex < pl.sx) ? ex : pl.sx, maxx = (ex > pl.sx) ? ex : pl.sx;
    double miny = (ey < pl.sy) ? ey : pl.sy, maxy = (ey > pl.sy) ? ey : pl.sy;
    double nan_guard = ex + ey;
    const double r = pl.rinv;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // world position of origin i (i = 3: the end point)
        const double lx = (i < 3) ? o[i].ox : gx, ly = (i < 3) ? o[i].oy : gy;
        const double wx = (cs * lx + (-ss) * ly) + pl.sx, wy = (ss * lx + cs * ly) + pl.sy;
        // plain compare-and-select (fmin / fmax cost ~13 instructions apiece here); a NaN coordinate poisons the pad
        // below and with it the whole box, which the verify kernel then refuses
        minx = (wx < minx) ? wx : minx;
        maxx = (wx > maxx) ? wx : maxx;
        miny = (wy < miny) ? wy : miny;
        maxy = (wy > maxy) ? wy : maxy;
        nan_guard += wx + wy;
        if (i == 3) break;
        const int mode = pp_word_mode(pl.word, i);
        const double len = pl.len[i];
        if (mode == PP_MODE_S || !(len > 0.0)) continue;
        const double C = fma(cs, o[i].co, -(ss * o[i].so)), S = fma(ss, o[i].co, cs * o[i].so);  // heading at the origin
        const double theta0 = pl.syaw + o[i].oyaw;
        const bool left = mode == PP_MODE_L;
        // centre of the turn circle; phi0 = direction centre -> origin; the sweep is counter-clockwise for a left turn
        const double ccx = left ? wx - r * S : wx + r * S, ccy = left ? wy + r * C : wy - r * C;
        const double phi0 = left ? theta0 - 0.5 * PP_PI : theta0 + 0.5 * PP_PI;
        // angle from phi0, in sweep direction, to the +x axis (k = 0); the other axes follow a quarter turn apart
        const double v = left ? -phi0 : phi0;
        const double a0 = v - PP_TWO_PI * floor(v * PP_INV_TWO_PI);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // left: axis k sits at k pi/2 counter-clockwise of +x; right (clockwise sweep): at k pi/2 clockwise, i.e. axis (4 - k) & 3
            double a = a0 + (double)k * (0.5 * PP_PI);
            if (a >= PP_TWO_PI) a -= PP_TWO_PI;
            const bool in = !(a > len + 1e-6) || (a > PP_TWO_PI - 1e-6);  // NaN: included
            const int axis = left ? k : ((4 - k) & 3);
            if (in) {
                if (axis == 0) maxx = (ccx + r > maxx) ? ccx + r : maxx;
                if (axis == 1) maxy = (ccy + r > maxy) ? ccy + r : maxy;
                if (axis == 2) minx = (ccx - r < minx) ? ccx - r : minx;
                if (axis == 3) miny = (ccy - r < miny) ? ccy - r : miny;
            }
            nan_guard += ccx + ccy;
        }
    }
    const double pad = 1e-9 * (((fabs(minx) + fabs(maxx)) + (fabs(miny) + fabs(maxy))) + (fabs(r) + 1.0)) + 0.0 * nan_guard;
    box[0] = minx - pad;
    box[1] = miny - pad;
    box[2] = maxx + pad;
    box[3] = maxy + pad;
}

// local-frame sample of output slot k (1 <= k <= n0+n1+n2)
__device__ __forceinline__ void pp_plan_sample_local(const pp_dubins_plan &pl, const pp_seg_origin o[3], uint32_t k,
                                                     double *x, double *y, double *yaw) {
    // the run-time `seg` index puts pl / o in local memory (L1-resident); measured faster than a select
    // tree, which costs ~30 instructions per sample and spills in the verify kernel
    uint32_t j = k - 1;
    int seg = 0;
    if (j >= pl.n[0]) {
        j -= pl.n[0];
        seg = 1;
        if (j >= pl.n[1]) {
            j -= pl.n[1];
            seg = 2;
        }
    }
    const double d = (pl.len[seg] > 0.0) ? pl.step : -pl.step;
    const double pd = pl.pd0[seg] + (double)j * d;  // reference accumulates; differs by <= j ulp (Q10)
    pp_interpolate(pp_word_mode(pl.word, seg), pd, o[seg].ox, o[seg].oy, o[seg].oyaw, o[seg].so, o[seg].co, pl.rinv, x,
                   y, yaw);
}


// ---- samples of one segment straight in the world frame (sample fill / scalar path kernels) ----------------------
// The reference interpolates in the local frame (src/dubins.rs:155-198) and rotates / translates every sample
// afterwards (:412-422).  Both maps are rigid, so they compose per SEGMENT: world origin (wox, woy) and the sine /
// cosine of (syaw + origin yaw) by the angle-addition formulas.  An arc sample at arc parameter pd = A + B, with A the
// parameter of the first sample of a 32-sample chunk and B = lane * step, is then LINEAR in (sin B, cos B):
//     x = X0 + cos B * Px + sin B * Qx,   y = Y0 + cos B * Py + sin B * Qy
// with four per-chunk coefficients (pp_arc_coef, one sincos per CHUNK instead of one per sample) and one sincos per
// lane and path for B.  Rounding differs from the reference's order by a few ulp of the path's extent -- the same
// class as `pd0 + j d` against the accumulated `pd` (Q10), far inside the 1e-9 contract.
struct pp_seg_world {
    double wox, woy;  // segment origin, world frame
    double Cr, Sr;    // radius * cos / sin (syaw + origin yaw)
    double X0, Y0;    // arc centre-relative constants (arc modes only)
    double Sg, Cg;    // Sr, Cr signed by the turn direction (+ left, - right)
};
__device__ __forceinline__ pp_seg_world pp_seg_world_make(double ss, double cs, double sx, double sy, double ox, double oy,
                                                          double so, double co, double rinv, int mode) {
    pp_seg_world s;
    s.wox = (cs * ox + (-ss) * oy) + sx;  // pp_local_to_world of the origin (exact identity for from_origin: ss = 0, cs = 1)
    s.woy = (ss * ox + cs * oy) + sy;
    const double C = fma(cs, co, -(ss * so)), S = fma(ss, co, cs * so);
    s.Cr = C * rinv;
    s.Sr = S * rinv;
    const bool right = mode == PP_MODE_R;
    s.Sg = right ? -s.Sr : s.Sr;
    s.Cg = right ? -s.Cr : s.Cr;
    s.X0 = s.wox - s.Sg;
    s.Y0 = s.woy + s.Cg;
    return s;
}
struct __align__(16) pp_arc_coef {
    double Px, Qx, Py, Qy;
};
// coefficients of the chunk whose first sample has arc parameter A
__device__ __forceinline__ pp_arc_coef pp_arc_coef_make(const pp_seg_world &s, double A) {
    double sA, cA;
    pp_sincos1(A, &sA, &cA);
    pp_arc_coef k;
    k.Px = fma(s.Cr, sA, s.Sg * cA);
    k.Qx = fma(s.Cr, cA, -(s.Sg * sA));
    k.Py = fma(s.Sr, sA, -(s.Cg * cA));
    k.Qy = fma(s.Sr, cA, s.Cg * sA);
    return k;
}
__device__ __forceinline__ void pp_arc_sample(const pp_seg_world &s, const pp_arc_coef &k, double sB, double cB, double *x,
                                              double *y) {
    *x = fma(cB, k.Px, fma(sB, k.Qx, s.X0));
    *y = fma(cB, k.Py, fma(sB, k.Qy, s.Y0));
}
__device__ __forceinline__ void pp_line_sample(const pp_seg_world &s, double pd, double *x, double *y) {
    *x = fma(pd, s.Cr, s.wox);
    *y = fma(pd, s.Sr, s.woy);
}

// local -> world (src/dubins.rs:412-422); (ss, cs) = sincos(syaw)
__device__ __forceinline__ void pp_local_to_world(double ss, double cs, double sx, double sy, double x, double y,
                                                  double *xw, double *yw) {
    *xw = (cs * x + (-ss) * y) + sx;
    *yw = (ss * x + cs * y) + sy;
}

// pp_replay.cuh -- exact fast-forward of the sample loop of generate_local_course (src/dubins.rs:239-255):
//
//     while pd.abs() <= l.abs() { ...; pd += d; }          // one sample per iteration
//
// The number of iterations and the final `pd` (it seeds the next segment through `ll`, :256) are defined by the
// REPEATED ROUNDED ADDITION, not by (l - pd0) / d: after the first few steps every `pd += d` rounds, and which side of
// `l` the last iterate lands on decides whether a path has n or n + 1 samples.  The plan kernel used to replay the
// loop literally: a serial DADD / DSETP / BRA chain of ~800 iterations per C5 path, 80 % of that kernel's
// instructions.  This header reproduces the loop's result bit for bit in O(binades) instead of O(samples).
//
// Why that is possible.  Take d > 0 and pd > 0 (the loop is symmetric under negation of both).  Inside one binade
// [2^e, 2^(e+1)) every iterate is a multiple of u = 2^(e-52), so round-to-nearest-even of pd + d moves pd by a
// CONSTANT D: the multiple of u nearest to d -- unless d lies exactly half-way between two multiples (common: the
// binade just above d's own drops exactly one bit of d), where the tie goes to the even neighbour and therefore
// depends on the parity of pd.  But a tie result is always even, so from the SECOND in-binade iterate onward the parity
// is fixed and the increment is constant again.  Hence: take three literal steps p1, p2, p3; if all three lie in one
// binade, D = p3 - p2 (exact) is the increment of every further step that stays in the binade, and
//     m = floor((min(|l|, top of the binade) - p3) / D)
// steps can be taken at once (all quantities are multiples of u below 2^53 u: the difference, the remainder check
// that repairs the floating quotient, and p3 + m D are exact).  Then literal steps again across the binade boundary.
// A path of 800 samples at step 0.05 crosses ~10 binades.
//
// A step that does not move pd (d below half an ulp of pd, or d = 0) means the reference never leaves the loop:
// reported as overflow at once.
// Mixed signs, zeros, subnormals, NaN, infinities take literal steps only (the reference's call sites produce
// pd0 in (0, 3d], src/dubins.rs:233-237).
//
// The header compiles as plain C++ too: tools/replay_check.cpp compares it with the literal loop on 10^7 cases.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define PP_REPLAY_FN __host__ __device__ __forceinline__
#else
#define PP_REPLAY_FN static inline
#endif

PP_REPLAY_FN uint64_t pp_replay_bits(double x) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
#endif
}
PP_REPLAY_FN double pp_replay_from_bits(uint64_t u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, 8);
    return x;
#endif
}

// Runs `while (fabs(pd) <= al) { pd += d; ++cnt; }` from (pd, cnt = 0).  Returns false when cnt reaches max_iters
// (the literal loop's overflow guard) -- then *cnt_out = max_iters and *pd_out is unspecified.
PP_REPLAY_FN bool pp_replay_segment(double pd, double d, double al, uint32_t max_iters, uint32_t *cnt_out, double *pd_out) {
    uint32_t cnt = 0;
    // work on magnitudes when pd and d point the same way; `flip` restores the sign at the end
    const bool flip = d < 0.0;
    if (flip) {
        pd = -pd;
        d = -d;
    }
    for (;;) {
        if (!(fabs(pd) <= al)) break;
        // three literal steps
        double p = pd;
        uint64_t e0 = 0;
        bool same = true, out = false;
#ifdef __CUDACC__
#pragma unroll
#endif
        for (int k = 0; k < 3; ++k) {
            const double prev = p;
            p = p + d;
            if (++cnt >= max_iters) {
                *cnt_out = max_iters;
                *pd_out = flip ? -p : p;
                return false;
            }
            const uint64_t e = pp_replay_bits(p) >> 52;  // sign + exponent field
            if (k == 0) e0 = e;
            same = same && (e == e0);
            if (!(fabs(p) <= al)) {
                out = true;
                break;
            }
            if (p == prev) {  // no progress and still inside: the reference loops for ever
                *cnt_out = max_iters;
                *pd_out = flip ? -p : p;
                return false;
            }
            if (k == 2) {
                // p1, p2, p3 in one positive normal binade: constant increment from here on
                if (same && e0 > 0 && e0 < 0x7ffull) {
                    const double D = p - prev;  // exact, > 0 (p != prev, and d > 0 cannot move a positive p down)
                    const double top = pp_replay_from_bits((e0 << 52) | 0xFFFFFFFFFFFFFull);  // largest value of the binade
                    const double lim = (al < top) ? al : top;
                    const double x = lim - p;  // exact: both multiples of the binade's ulp (al >= p lies in it or above)
                    double q = floor(x / D);
                    const double r = fma(-q, D, x);  // exact remainder
                    if (r < 0.0)
                        q -= 1.0;
                    else if (r >= D)
                        q += 1.0;
                    if (q >= (double)(max_iters - cnt)) {
                        *cnt_out = max_iters;
                        *pd_out = flip ? -p : p;
                        return false;
                    }
                    p = fma(q, D, p);  // exact
                    cnt += (uint32_t)q;
                }
            }
        }
        pd = p;
        if (out) break;
    }
    *cnt_out = cnt;
    *pd_out = flip ? -pd : pd;
    return true;
}

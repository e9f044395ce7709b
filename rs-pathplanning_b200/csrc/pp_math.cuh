// pp_math.cuh -- hand-written f64 sincos / atan2 / acos for the Dubins kernels.
//
// Why not CUDA's libm: on sm_100 a double constant is materialised with two 32-bit uniform moves, so the
// inlined libm polynomials spend more issue slots on constants than on DFMAs (ncu, profiles/r01_summary.md:
// 2 146 instructions per pose pair, only ~730 of them on the FP64 pipe).  Here the coefficient tables live in
// __constant__ memory (one LDCU.128 fetches two doubles) and every routine is written as an N-way batch:
// the N evaluations are interleaved instruction by instruction, so a coefficient is fetched once per batch
// and the N independent dependency chains give the FP64 pipe its ILP.
//
// Accuracy (tests/test_math_host.py, 10^6 samples per function against glibc): sincos <= 1.5 ulp for
// |x| < 2^16, atan2 <= 1.5 ulp, acos <= 1.2 ulp.  Coefficients: tools/gen_math_tables.py.
// The header also compiles as plain C++ (g++) for that host-side accuracy test.
#pragma once
#include <math.h>
#include <string.h>

#include "pp_math_tables.inc"

#ifdef __CUDACC__
#define PP_MATH_FN __device__ __forceinline__
#define PP_MATH_TABLE static __constant__ double
#define PP_UNROLL _Pragma("unroll")
#else
#define PP_MATH_FN static inline
#define PP_MATH_TABLE static const double
#define PP_UNROLL
#endif

// Coefficient tables.  On the device they are __constant__ arrays WITHOUT a static initialiser, filled by
// pp_math_upload_tables() when a context is created: with an initialiser nvcc folds the values back into
// the instruction stream as pairs of 32-bit uniform moves, which is exactly the overhead to avoid.  Each
// translation unit that includes this header owns (and uploads) its private copy; 16-byte alignment lets a
// pair of coefficients arrive with one LDCU.128.
#ifdef __CUDACC__
static __constant__ __attribute__((aligned(16))) double pp_sin_c[6];
static __constant__ __attribute__((aligned(16))) double pp_cos_c[6];
static __constant__ __attribute__((aligned(16))) double pp_atan_c[22];
static __constant__ __attribute__((aligned(16))) double pp_asin_c[14];  // 13 used
static inline cudaError_t pp_math_upload_tables() {
    static const double h_sin[6] = PP_SIN_COEF, h_cos[6] = PP_COS_COEF, h_atan[22] = PP_ATAN_COEF;
    static const double h_asin[14] = PP_ASIN_COEF;
    cudaError_t e = cudaMemcpyToSymbol(pp_sin_c, h_sin, sizeof h_sin);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(pp_cos_c, h_cos, sizeof h_cos);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(pp_atan_c, h_atan, sizeof h_atan);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(pp_asin_c, h_asin, sizeof h_asin);
    return e;
}
#else
PP_MATH_TABLE pp_sin_c[6] = PP_SIN_COEF;
PP_MATH_TABLE pp_cos_c[6] = PP_COS_COEF;
PP_MATH_TABLE pp_atan_c[22] = PP_ATAN_COEF;
PP_MATH_TABLE pp_asin_c[14] = PP_ASIN_COEF;
#endif

// Range checks that do not need the value's low word are done on the HIGH 32 bits with integer instructions:
// the FP64 pipe is the kernels' bound (64 lanes per SM), the integer ALU is nearly idle, and a DSETP costs the
// same FP64 issue slot as a DFMA.  For non-negative doubles the bit patterns order like the values.
PP_MATH_FN unsigned pp_hi32(double x) {
#ifdef __CUDA_ARCH__
    return (unsigned)__double2hiint(x);
#else
    unsigned long long u;
    memcpy(&u, &x, 8);
    return (unsigned)(u >> 32);
#endif
}
PP_MATH_FN long long pp_bits(double x) {
#ifdef __CUDA_ARCH__
    return __double_as_longlong(x);
#else
    long long u;
    memcpy(&u, &x, 8);
    return u;
#endif
}
// |x| < 2^k for a finite x (false for NaN / inf): exponent-field compare
PP_MATH_FN bool pp_abs_below_pow2(double x, int k) { return (pp_hi32(x) & 0x7fffffffu) < (unsigned)(1023 + k) << 20; }
// x with its sign flipped when neg (integer XOR on the high word)
PP_MATH_FN double pp_negate_if(double x, bool neg) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(x) ^ (neg ? (int)0x80000000u : 0), __double2loint(x));
#else
    return neg ? -x : x;
#endif
}

// ---- sincos: Cody-Waite reduction by pi/2 with three FMAs (pi/2 = P1 + P2 + P3 to ~160 bits), then the
// two kernels sin(r) = r + r z S(z), cos(r) = 1 - z/2 + z^2 C(z) on |r| <= pi/4, swapped / negated by quadrant.
// Used for |x| < 2^16; larger or non-finite arguments take the library routine (rare branch).
template <int N>
PP_MATH_FN void pp_sincos_n(const double (&x)[N], double (&s)[N], double (&c)[N]) {
    double r[N], z[N], ps[N], pc[N];
    int q[N];
    bool slow = false;
    PP_UNROLL
    for (int i = 0; i < N; ++i) {
        // sin is odd, cos even: reduce |x| and put the sign back at the end (keeps sin(-0.0) = -0.0)
        const double xa = fabs(x[i]);
        slow |= !pp_abs_below_pow2(x[i], 16);
        const double kf = rint(xa * PP_TWO_OVER_PI);
        q[i] = (int)kf;
        double t = fma(-kf, PP_PIO2_1, xa);
        t = fma(-kf, PP_PIO2_2, t);
        r[i] = fma(-kf, PP_PIO2_3, t);
        z[i] = r[i] * r[i];
        ps[i] = pp_sin_c[5];
        pc[i] = pp_cos_c[5];
    }
    PP_UNROLL
    for (int j = 4; j >= 0; --j) {
        PP_UNROLL
        for (int i = 0; i < N; ++i) {
            ps[i] = fma(ps[i], z[i], pp_sin_c[j]);
            pc[i] = fma(pc[i], z[i], pp_cos_c[j]);
        }
    }
    PP_UNROLL
    for (int i = 0; i < N; ++i) {
        const double sr = fma(r[i] * z[i], ps[i], r[i]);
        const double hz = 0.5 * z[i];
        const double w = 1.0 - hz;
        const double cr = w + (((1.0 - w) - hz) + (z[i] * z[i]) * pc[i]);
        const bool swap = (q[i] & 1) != 0;
        const double sv = swap ? cr : sr;
        const double cv = swap ? sr : cr;
        s[i] = (((q[i] & 2) != 0) != signbit(x[i])) ? -sv : sv;
        c[i] = ((q[i] + 1) & 2) ? -cv : cv;
    }
    if (slow) {
        PP_UNROLL
        for (int i = 0; i < N; ++i)
            if (!pp_abs_below_pow2(x[i], 16)) {
#ifdef __CUDACC__
                sincos(x[i], &s[i], &c[i]);
#else
                s[i] = sin(x[i]);
                c[i] = cos(x[i]);
#endif
            }
    }
}

PP_MATH_FN void pp_sincos1(double x, double *s, double *c) {
    const double xi[1] = {x};
    double so[1], co[1];
    pp_sincos_n<1>(xi, so, co);
    *s = so[0];
    *c = co[0];
}

// quotient n / d for 0 <= n <= d, d a positive normal number: hardware reciprocal seed (MUFU.RCP64H), two Newton
// steps and one residual correction (<= 1 ulp); the IEEE division's range checks and slow-path call are not
// needed for these operands.  The host build (accuracy harness) uses the plain division.
PP_MATH_FN double pp_div_01(double n, double d) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    const double q = n * r;
    return fma(fma(-d, q, n), r, q);
#else
    return n / d;
#endif
}

// sqrt(x) for finite x >= 0: hardware reciprocal-square-root seed (MUFU.RSQ64H, ~2^-22), one coupled Newton
// step on (g ~ sqrt x, h ~ 1/(2 sqrt x)) and one residual correction (<= 1 ulp, exact for 0); no range checks.
PP_MATH_FN double pp_sqrt_pos(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x + 0x1p-1000));  // keeps the seed finite for x = 0
    double g = x * y, h = 0.5 * y;
    const double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    return fma(fma(-g, g, x), h, g);
#else
    return sqrt(x);
#endif
}

// ---- atan2: a = min(|x|,|y|) / max(|x|,|y|) in [0,1], atan(a) = a + a s A(s) (degree-21 polynomial in
// s = a^2), then the octant fix-ups pi/2 - r, pi - r and the sign of y.  Signed zeros follow C99
// (atan2(+-0, -x) = +-pi, atan2(+-0, +x) = +-0); NaN propagates; (0, 0) gives +-0 or +-pi like glibc.
template <int N>
PP_MATH_FN void pp_atan2_n(const double (&y)[N], const double (&x)[N], double (&out)[N]) {
    double a[N], s[N], p[N];
    bool swp[N];
    PP_UNROLL
    for (int i = 0; i < N; ++i) {
        const double ax = fabs(x[i]), ay = fabs(y[i]);
        // one compare orders the pair (and is reused for the octant fix-up); a NaN operand makes the
        // compare false and then reaches the quotient, so NaN propagates; inf/inf gives NaN (libm: pi/4 ...),
        // which only happens for non-finite poses whose cost is not finite either
        // (integer compare of the two non-negative bit patterns: same order, NaN sorts above infinity)
        swp[i] = pp_bits(ay) > pp_bits(ax);
        const double mx = swp[i] ? ay : ax, mn = swp[i] ? ax : ay;
        a[i] = pp_div_01(mn, mx + 0x1p-1000);  // (0, 0) -> 0; leaves every mx > 1e-285 unchanged; NaN propagates
        s[i] = a[i] * a[i];
        p[i] = pp_atan_c[21];
    }
    PP_UNROLL
    for (int j = 20; j >= 0; --j) {
        PP_UNROLL
        for (int i = 0; i < N; ++i) p[i] = fma(p[i], s[i], pp_atan_c[j]);
    }
    PP_UNROLL
    for (int i = 0; i < N; ++i) {
        // octants in one step: (B + sgn r) with B = 0, pi/2, pi, pi/2 and sgn = +,-,-,+ for
        // (swap, x<0) = (0,0), (1,0), (0,1), (1,1); base and sign are picked with integer selects
        const double r = fma(a[i] * s[i], p[i], a[i]);
        const bool xneg = signbit(x[i]);
        const double bh = swp[i] ? PP_PIO2 : (xneg ? PP_PI_HI : 0.0);
        const double bl = swp[i] ? PP_PIO2_LO : (xneg ? PP_PI_LO : 0.0);
        out[i] = copysign((bh + pp_negate_if(r, swp[i] != xneg)) + bl, y[i]);
    }
}

PP_MATH_FN double pp_atan2(double y, double x) {
    const double yi[1] = {y}, xi[1] = {x};
    double o[1];
    pp_atan2_n<1>(yi, xi, o);
    return o[0];
}

// ---- acos, direct: asin(r) = r + r z Q(z) on z = r^2 <= 1/4 (13 coefficients), with the usual two ranges
//   |v| <= 1/2 : acos(v) = pi/2 - asin(v)
//   |v| >  1/2 : r = sqrt((1 - |v|) / 2) (exact subtraction), acos(|v|) = 2 asin(r), acos(-|v|) = pi - 2 asin(r)
// |v| > 1 gives NaN (the square root of a negative number), NaN propagates.  About 30 FP64 instructions per
// value against ~47 for the atan2 form, and no division.
template <int N>
PP_MATH_FN void pp_acos_n(const double (&v)[N], double (&out)[N]) {
    double s[N], z[N], p[N];
    bool big[N];
    PP_UNROLL
    for (int i = 0; i < N; ++i) {
        const double av = fabs(v[i]);
        big[i] = (pp_hi32(av) >= 0x3fe00000u);  // av >= 0.5 (both branches are valid at 0.5)
        z[i] = big[i] ? (1.0 - av) * 0.5 : av * av;
        p[i] = pp_asin_c[12];
    }
    PP_UNROLL
    for (int i = 0; i < N; ++i) {
#ifdef __CUDA_ARCH__
        const double rt = pp_sqrt_pos(z[i]);
#else
        const double rt = sqrt(z[i]);
#endif
        s[i] = big[i] ? rt : v[i];
    }
    PP_UNROLL
    for (int j = 11; j >= 0; --j) {
        PP_UNROLL
        for (int i = 0; i < N; ++i) p[i] = fma(p[i], z[i], pp_asin_c[j]);
    }
    PP_UNROLL
    for (int i = 0; i < N; ++i) {
        const double corr = (s[i] * z[i]) * p[i];  // asin(s) - s
        if (big[i]) {
            const double r2 = 2.0 * (s[i] + corr);
            out[i] = signbit(v[i]) ? (PP_PI_HI - r2) + PP_PI_LO : r2;
        } else {
            out[i] = PP_PIO2 - (s[i] - (PP_PIO2_LO - corr));
        }
    }
}
PP_MATH_FN double pp_acos(double v) {
    const double vi[1] = {v};
    double o[1];
    pp_acos_n<1>(vi, o);
    return o[0];
}

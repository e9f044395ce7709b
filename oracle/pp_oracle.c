/*
 * pp_oracle.c -- CPU ORACLE (test infrastructure, NOT product code; see pp_oracle.h).
 *
 * Each function cites the reference lines (under /root/reference) it restates.
 * Operation order follows the Rust source exactly; compile with
 * -ffp-contract=off so no multiply-add is fused (rustc never contracts).
 * PARITY UNPINNED by the reference (it has no tests); see pp_oracle.h.
 */
#define _GNU_SOURCE
#include "pp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* std::f64::consts::PI is the double nearest pi, same as M_PI */
#define PI 3.14159265358979323846
#define TWO_PI (2.0 * PI)

/* host cores available to this process; deliberately NOT omp_get_max_threads(): torchrun exports
 * OMP_NUM_THREADS=1, and the CPU baseline must use all the host threads it can */
int ppo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
static int resolve_threads(int nthreads) {
    int mx = ppo_max_threads();
    if (nthreads <= 0) return mx;
    return nthreads;
}

/* ------------------------------------------------------------------ dubins.rs */

/* src/dubins.rs:14-16  fn fmodr */
static double fmodr(double x, double y) { return x - y * floor(x / y); }
/* src/dubins.rs:18-20 */
double ppo_mod2pi(double theta) { return fmodr(theta, TWO_PI); }
/* src/dubins.rs:22-24 ; Rust `%` on f64 is C fmod */
double ppo_pi_2_pi(double angle) { return fmod(angle + PI, TWO_PI) - PI; }

/* src/dubins.rs:27-48 */
static int w_lsl(double alpha, double beta, double d, double o[3]) {
    double sa = sin(alpha), sb = sin(beta), ca = cos(alpha), cb = cos(beta), c_ab = cos(alpha - beta);
    double tmp0 = d + sa - sb;
    double p_squared = 2.0 + (d * d) - (2.0 * c_ab) + (2.0 * d * (sa - sb));
    if (p_squared < 0.0) return 0;
    double tmp1 = atan2(cb - ca, tmp0);
    o[0] = ppo_mod2pi(-alpha + tmp1);
    o[1] = sqrt(p_squared);
    o[2] = ppo_mod2pi(beta - tmp1);
    return 1;
}
/* src/dubins.rs:51-71 */
static int w_rsr(double alpha, double beta, double d, double o[3]) {
    double sa = sin(alpha), sb = sin(beta), ca = cos(alpha), cb = cos(beta), c_ab = cos(alpha - beta);
    double tmp0 = d - sa + sb;
    double p_squared = 2.0 + (d * d) - (2.0 * c_ab) + (2.0 * d * (sb - sa));
    if (p_squared < 0.0) return 0;
    double tmp1 = atan2(ca - cb, tmp0);
    o[0] = ppo_mod2pi(alpha - tmp1);
    o[1] = sqrt(p_squared);
    o[2] = ppo_mod2pi(-beta + tmp1);
    return 1;
}
/* src/dubins.rs:74-92 */
static int w_lsr(double alpha, double beta, double d, double o[3]) {
    double sa = sin(alpha), sb = sin(beta), ca = cos(alpha), cb = cos(beta), c_ab = cos(alpha - beta);
    double p_squared = -2.0 + (d * d) + (2.0 * c_ab) + (2.0 * d * (sa + sb));
    if (p_squared < 0.0) return 0;
    double p = sqrt(p_squared);
    double tmp = atan2(-ca - cb, d + sa + sb) - atan2(-2.0, p);
    o[0] = ppo_mod2pi(-alpha + tmp);
    o[1] = p;
    o[2] = ppo_mod2pi(-ppo_mod2pi(beta) + tmp);
    return 1;
}
/* src/dubins.rs:95-113 */
static int w_rsl(double alpha, double beta, double d, double o[3]) {
    double sa = sin(alpha), sb = sin(beta), ca = cos(alpha), cb = cos(beta), c_ab = cos(alpha - beta);
    double p_squared = -2.0 + (d * d) + (2.0 * c_ab) - (2.0 * d * (sa + sb));
    if (p_squared < 0.0) return 0;
    double p = sqrt(p_squared);
    double tmp = atan2(ca + cb, d - sa - sb) - atan2(2.0, p);
    o[0] = ppo_mod2pi(alpha - tmp);
    o[1] = p;
    o[2] = ppo_mod2pi(beta - tmp);
    return 1;
}
/* src/dubins.rs:116-133 */
static int w_rlr(double alpha, double beta, double d, double o[3]) {
    double sa = sin(alpha), sb = sin(beta), ca = cos(alpha), cb = cos(beta), c_ab = cos(alpha - beta);
    double tmp_rlr = (6.0 - d * d + 2.0 * c_ab + 2.0 * d * (sa - sb)) / 8.0;
    if (fabs(tmp_rlr) > 1.0) return 0;
    double p = ppo_mod2pi(2.0 * PI - acos(tmp_rlr));
    double t = ppo_mod2pi(alpha - atan2(ca - cb, d - sa + sb) + ppo_mod2pi(p / 2.0));
    double q = ppo_mod2pi(alpha - beta - t + ppo_mod2pi(p));
    o[0] = t;
    o[1] = p;
    o[2] = q;
    return 1;
}
/* src/dubins.rs:136-153 */
static int w_lrl(double alpha, double beta, double d, double o[3]) {
    double sa = sin(alpha), sb = sin(beta), ca = cos(alpha), cb = cos(beta), c_ab = cos(alpha - beta);
    double tmp_lrl = (6.0 - d * d + 2.0 * c_ab + 2.0 * d * (-sa + sb)) / 8.0;
    if (fabs(tmp_lrl) > 1.0) return 0;
    double p = ppo_mod2pi(2.0 * PI - acos(tmp_lrl));
    double t = ppo_mod2pi(-alpha - atan2(ca - cb, d + sa - sb) + p / 2.0);
    double q = ppo_mod2pi(ppo_mod2pi(beta) - alpha - t + ppo_mod2pi(p));
    o[0] = t;
    o[1] = p;
    o[2] = q;
    return 1;
}

int ppo_dubins_word(int word, double alpha, double beta, double d, double tpq[3]) {
    switch (word) { /* ALL_PLANNERS order, src/dubins.rs:291 */
    case PPO_LSL: return w_lsl(alpha, beta, d, tpq);
    case PPO_RSR: return w_rsr(alpha, beta, d, tpq);
    case PPO_LSR: return w_lsr(alpha, beta, d, tpq);
    case PPO_RSL: return w_rsl(alpha, beta, d, tpq);
    case PPO_RLR: return w_rlr(alpha, beta, d, tpq);
    case PPO_LRL: return w_lrl(alpha, beta, d, tpq);
    }
    return 0;
}

/* HARNESS: signed distance of a word's feasibility test from flipping, relative to 1 + d^2
 * (CSC: p^2 >= 0, src/dubins.rs:38,61,82,103; CCC: |tmp| <= 1, :124,144) */
double ppo_dubins_word_margin(int word, double alpha, double beta, double d) {
    double sa = sin(alpha), sb = sin(beta), c_ab = cos(alpha - beta), m = 0.0;
    switch (word) {
    case PPO_LSL: m = 2.0 + d * d - 2.0 * c_ab + 2.0 * d * (sa - sb); break;
    case PPO_RSR: m = 2.0 + d * d - 2.0 * c_ab + 2.0 * d * (sb - sa); break;
    case PPO_LSR: m = -2.0 + d * d + 2.0 * c_ab + 2.0 * d * (sa + sb); break;
    case PPO_RSL: m = -2.0 + d * d + 2.0 * c_ab - 2.0 * d * (sa + sb); break;
    case PPO_RLR: m = 1.0 - fabs((6.0 - d * d + 2.0 * c_ab + 2.0 * d * (sa - sb)) / 8.0); break;
    case PPO_LRL: m = 1.0 - fabs((6.0 - d * d + 2.0 * c_ab + 2.0 * d * (-sa + sb)) / 8.0); break;
    }
    return m / (1.0 + d * d);
}

/* segment modes per word: 0 = L, 1 = S, 2 = R  (src/dubins.rs:26,50,73,94,115,135) */
static const int WORD_MODES[6][3] = {{0, 1, 0}, {2, 1, 2}, {0, 1, 2}, {2, 1, 0}, {2, 0, 2}, {0, 2, 0}};

static int near_wrap(double v) { return (v < PPO_WRAP_EPS) || (TWO_PI - v < PPO_WRAP_EPS); }

/* src/dubins.rs:333-363: set-up in the start frame + selection fold.
 * (lex, ley, leyaw) already local; c = 1/turn_radius. */
static int eval_from_origin(double dx, double dy, double eyaw, double c, double *cost, double tpq[3],
                            uint32_t *flags) {
    double hyp = hypot(dx, dy);
    double d = hyp * c;
    double theta = ppo_mod2pi(atan2(dy, dx));
    double alpha = ppo_mod2pi(-theta);
    double beta = ppo_mod2pi(eyaw - theta);

    double bcost = INFINITY, second = INFINITY;
    int bword = PPO_NONE;
    double b[3] = {0, 0, 0};
    double costs[6];
    int wraps[6];
    for (int w = 0; w < 6; ++w) {
        double o[3];
        costs[w] = INFINITY;
        wraps[w] = 0;
        if (ppo_dubins_word(w, alpha, beta, d, o)) {
            double wcost = fabs(o[0]) + fabs(o[1]) + fabs(o[2]); /* :352 */
            costs[w] = wcost;
            /* t and q of every word, and p of the CCC words, come out of mod2pi */
            wraps[w] = near_wrap(o[0]) || near_wrap(o[2]) || (w >= 4 && near_wrap(o[1]));
            if (bcost > wcost) { /* :354 strict, first wins ties; NaN never wins */
                second = bcost;
                b[0] = o[0];
                b[1] = o[1];
                b[2] = o[2];
                bword = w;
                bcost = wcost;
            } else if (wcost < second) {
                second = wcost;
            }
        }
    }
    if (flags) {
        uint32_t f = 0;
        if (bword != PPO_NONE) {
            double tol = PPO_TIE_REL * (bcost > 1.0 ? bcost : 1.0);
            if (second - bcost <= tol) f |= PPO_FLAG_NEAR_TIE;
            /* a wrap in any word matters if un-wrapping it could change the winner or the cost:
             * the winner's own wraps, or another word whose cost minus 2pi (or 4pi) would undercut. */
            for (int w = 0; w < 6; ++w) {
                if (!wraps[w]) continue;
                if (w == bword || costs[w] - 2.0 * TWO_PI <= bcost + tol) f |= PPO_FLAG_NEAR_WRAP;
            }
            /* setup-angle wraps: theta/alpha/beta near 0/2pi are harmless by periodicity (Q14) */
        }
        /* feasibility margins */
        {
            double sa = sin(alpha), sb = sin(beta), c_ab = cos(alpha - beta);
            double m[6];
            m[0] = 2.0 + d * d - 2.0 * c_ab + 2.0 * d * (sa - sb);
            m[1] = 2.0 + d * d - 2.0 * c_ab + 2.0 * d * (sb - sa);
            m[2] = -2.0 + d * d + 2.0 * c_ab + 2.0 * d * (sa + sb);
            m[3] = -2.0 + d * d + 2.0 * c_ab - 2.0 * d * (sa + sb);
            m[4] = 1.0 - fabs((6.0 - d * d + 2.0 * c_ab + 2.0 * d * (sa - sb)) / 8.0);
            m[5] = 1.0 - fabs((6.0 - d * d + 2.0 * c_ab + 2.0 * d * (-sa + sb)) / 8.0);
            double scale = 1.0 + d * d;
            for (int w = 0; w < 6; ++w)
                if (fabs(m[w]) <= 1e-9 * scale) f |= PPO_FLAG_NEAR_FEAS;
        }
        *flags = f;
    }
    if (bword == PPO_NONE) {
        *cost = INFINITY;
        tpq[0] = tpq[1] = tpq[2] = NAN;
        return PPO_NONE;
    }
    *cost = bcost;
    tpq[0] = b[0];
    tpq[1] = b[1];
    tpq[2] = b[2];
    return bword;
}

/* src/dubins.rs:401-408 then :326-363 */
int ppo_dubins_eval(double sx, double sy, double syaw, double ex, double ey, double eyaw, double radius,
                    double *cost, double tpq[3], uint32_t *flags) {
    double ex_ = ex - sx;
    double ey_ = ey - sy;
    double c = 1.0 / radius;
    double lex = cos(syaw) * ex_ + sin(syaw) * ey_;
    double ley = -(sin(syaw)) * ex_ + cos(syaw) * ey_;
    double leyaw = eyaw - syaw;
    return eval_from_origin(lex, ley, leyaw, c, cost, tpq, flags);
}

void ppo_dubins_eval_batch(size_t n, const double *sx, const double *sy, const double *syaw, const double *ex,
                           const double *ey, const double *eyaw, const double *radius_arr, double radius,
                           double *cost, uint8_t *word, double *tpq, uint32_t *flags, int nthreads) {
    int nt = resolve_threads(nthreads);
    (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (long i = 0; i < (long)n; ++i) {
        double c, o[3];
        uint32_t f = 0;
        double r = radius_arr ? radius_arr[i] : radius;
        int w = ppo_dubins_eval(sx[i], sy[i], syaw[i], ex[i], ey[i], eyaw[i], r, &c, o, flags ? &f : NULL);
        if (cost) cost[i] = c;
        if (word) word[i] = (uint8_t)w;
        if (tpq) {
            tpq[3 * i] = o[0];
            tpq[3 * i + 1] = o[1];
            tpq[3 * i + 2] = o[2];
        }
        if (flags) flags[i] = f;
    }
}

/* src/dubins.rs:155-198.  mode: 0 L, 1 S, 2 R.  `directions` is computed and discarded
 * by the reference (Q13); not modelled. */
static void interpolate(long ind, double length, int mode, double max_curvature, double origin_x,
                        double origin_y, double origin_yaw, double *path_x, double *path_y, double *path_yaw) {
    if (mode == 1) {
        path_x[ind] = origin_x + length / max_curvature * cos(origin_yaw);
        path_y[ind] = origin_y + length / max_curvature * sin(origin_yaw);
        path_yaw[ind] = origin_yaw;
    } else {
        double ldx = sin(length) / max_curvature;
        double ldy = 0.0;
        if (mode == 0)
            ldy = (1.0 - cos(length)) / max_curvature;
        else if (mode == 2)
            ldy = (1.0 - cos(length)) / -max_curvature;
        double gdx = cos(-origin_yaw) * ldx + sin(-origin_yaw) * ldy;
        double gdy = -sin(-origin_yaw) * ldx + cos(-origin_yaw) * ldy;
        path_x[ind] = origin_x + gdx;
        path_y[ind] = origin_y + gdy;
    }
    if (mode == 0)
        path_yaw[ind] = origin_yaw + length;
    else if (mode == 2)
        path_yaw[ind] = origin_yaw - length;
}

/* src/dubins.rs:200-289.  Buffers are zero-initialised with n_point entries.
 * Returns the trimmed length, or -3 if an index reaches n_point (Rust would panic). */
static long generate_local_course(const double lengths[3], const int mode[3], double max_curvature,
                                  double step_size, double *px, double *py, double *pyaw, long n_point,
                                  uint32_t *flags) {
    long ind = 1;
    double ll = 0.0;
    /* harness only: how close the count-deciding comparisons below are to flipping.  An implementation whose
     * segment lengths agree to 1e-9 relative can take one iteration more or less of the loop at :239 exactly
     * when |pd| and |l| are that close at the last accepted or the first rejected iteration. */
    double count_tol = PPO_COUNT_REL * (1.0 + fabs(lengths[0]) + fabs(lengths[1]) + fabs(lengths[2]));
    for (int i = 0; i < 3; ++i) {
        int m = mode[i];
        double l = lengths[i];
        double d = (l > 0.0) ? step_size : -step_size; /* :228 */
        if (ind >= n_point) return -3;
        double ox = px[ind], oy = py[ind], oyaw = pyaw[ind]; /* :230 */
        ind -= 1;                                            /* :232 */
        double pd = (i >= 1 && (lengths[i - 1] * lengths[i]) > 0.0) ? (-d - ll) : (d - ll); /* :233-237 */
        while (fabs(pd) <= fabs(l)) { /* :239 */
            ind += 1;
            if (ind >= n_point) return -3;
            interpolate(ind, pd, m, max_curvature, ox, oy, oyaw, px, py, pyaw);
            if (flags && fabs(l) - fabs(pd) <= count_tol) *flags |= PPO_FLAG_NEAR_COUNT; /* accepted by a hair */
            pd += d;
        }
        if (flags && fabs(pd) - fabs(l) <= count_tol) *flags |= PPO_FLAG_NEAR_COUNT; /* rejected by a hair */
        ll = l - pd - d; /* :256 */
        ind += 1;
        if (ind >= n_point) return -3;
        interpolate(ind, l, m, max_curvature, ox, oy, oyaw, px, py, pyaw); /* :258-271 */
    }
    /* :274-279 never triggers (n_point >= 7) */
    long len = n_point;
    if (len <= 1) len = 0;
    if (len == 0) return -3; /* path_x[len-1] would panic */
    /* :281-288 trim: pops trailing zeros and then one more element */
    if (flags) { /* a written slot whose local x is nearly (not exactly) zero makes the trim length a knife edge */
        for (long k = ind; k >= 1 && k >= ind - 2; --k)
            if (px[k] != 0.0 && fabs(px[k]) <= count_tol / max_curvature) *flags |= PPO_FLAG_NEAR_COUNT;
    }
    double last = px[len - 1];
    while (len >= 1 && last == 0.0) {
        last = px[len - 1];
        len -= 1;
    }
    return len;
}

static long dubins_path_core(double sx, double sy, double syaw, double ex, double ey, double eyaw, double radius,
                             double step, int from_origin, double *out_x, double *out_y, double *out_yaw, size_t cap,
                             int *word, double *cost, long *n_point_out, uint32_t *flags);

long ppo_dubins_path(double sx, double sy, double syaw, double ex, double ey, double eyaw, double radius,
                     double step, int from_origin, double *out_x, double *out_y, double *out_yaw, size_t cap,
                     int *word, double *cost, long *n_point_out) {
    return dubins_path_core(sx, sy, syaw, ex, ey, eyaw, radius, step, from_origin, out_x, out_y, out_yaw, cap, word,
                            cost, n_point_out, NULL);
}

long ppo_dubins_path_flags(double sx, double sy, double syaw, double ex, double ey, double eyaw, double radius,
                           double step, int from_origin, double *out_x, double *out_y, double *out_yaw, size_t cap,
                           int *word, double *cost, uint32_t *flags) {
    return dubins_path_core(sx, sy, syaw, ex, ey, eyaw, radius, step, from_origin, out_x, out_y, out_yaw, cap, word,
                            cost, NULL, flags);
}

static long dubins_path_core(double sx, double sy, double syaw, double ex, double ey, double eyaw, double radius,
                             double step, int from_origin, double *out_x, double *out_y, double *out_yaw, size_t cap,
                             int *word, double *cost, long *n_point_out, uint32_t *flags) {
    double lex, ley, leyaw, c = 1.0 / radius;
    if (flags) {
        *flags = 0;
        /* one ulp of a yaw beyond ~2^20 rad is more than 1e-10 rad: the sample positions are then decided by
         * the last digits of the argument reduction, not by the path */
        if (fabs(syaw) > PPO_HUGE_ANGLE || fabs(eyaw) > PPO_HUGE_ANGLE) *flags |= PPO_FLAG_HUGE_ANGLE;
    }
    if (from_origin) {
        lex = ex;
        ley = ey;
        leyaw = eyaw;
    } else { /* src/dubins.rs:402-408 */
        double ex_ = ex - sx;
        double ey_ = ey - sy;
        lex = cos(syaw) * ex_ + sin(syaw) * ey_;
        ley = -(sin(syaw)) * ex_ + cos(syaw) * ey_;
        leyaw = eyaw - syaw;
    }
    double bcost, b[3];
    uint32_t eflags = 0;
    int w = eval_from_origin(lex, ley, leyaw, c, &bcost, b, flags ? &eflags : NULL);
    if (flags) *flags |= eflags;
    if (word) *word = w;
    if (cost) *cost = bcost;
    if (w == PPO_NONE) return -1; /* :397 */
    /* :367-375 */
    double total_length = 0.0 + b[0] + b[1] + b[2];
    double np_f = trunc(total_length / step);
    if (!(np_f >= 0.0) || np_f > 1e9) return -3;
    long n_point = (long)np_f + 3 + 4;
    if (n_point_out) *n_point_out = n_point;
    double *buf = (double *)calloc((size_t)n_point * 3, sizeof(double));
    if (!buf) return -3;
    double *px = buf, *py = buf + n_point, *pyaw = buf + 2 * n_point;
    long len = generate_local_course(b, WORD_MODES[w], c, step, px, py, pyaw, n_point, flags);
    /* bit-identical poses: d = alpha = beta = 0 exactly, every word length is an exact 0 and the path is empty
     * in any implementation -- nothing about it is a knife edge */
    if (flags && !from_origin && sx == ex && sy == ey && syaw == eyaw) *flags = 0;
    if (len < 0) {
        free(buf);
        return -3;
    }
    if ((size_t)len > cap) {
        free(buf);
        return -2;
    }
    if (from_origin) {
        for (long k = 0; k < len; ++k) {
            out_x[k] = px[k];
            out_y[k] = py[k];
            if (out_yaw) out_yaw[k] = pyaw[k];
        }
    } else { /* src/dubins.rs:412-422 */
        for (long k = 0; k < len; ++k) {
            double x = px[k], y = py[k];
            out_x[k] = cos(-syaw) * x + sin(-syaw) * y + sx;
            out_y[k] = -sin(-syaw) * x + cos(-syaw) * y + sy;
            if (out_yaw) out_yaw[k] = ppo_pi_2_pi(pyaw[k] + syaw);
        }
    }
    free(buf);
    return len;
}

void ppo_dubins_count_batch(size_t n, const double *sx, const double *sy, const double *syaw, const double *ex,
                            const double *ey, const double *eyaw, double radius, double step, int64_t *counts,
                            int nthreads) {
    int nt = resolve_threads(nthreads);
    (void)nt;
#pragma omp parallel num_threads(nt)
    {
        size_t cap = 1 << 16;
        double *bx = (double *)malloc(cap * sizeof(double)), *by = (double *)malloc(cap * sizeof(double));
#pragma omp for schedule(dynamic, 64)
        for (long i = 0; i < (long)n; ++i) {
            long r;
            for (;;) {
                r = ppo_dubins_path(sx[i], sy[i], syaw[i], ex[i], ey[i], eyaw[i], radius, step, 0, bx, by, NULL,
                                    cap, NULL, NULL, NULL);
                if (r != -2) break;
                cap *= 2;
                free(bx);
                free(by);
                bx = (double *)malloc(cap * sizeof(double));
                by = (double *)malloc(cap * sizeof(double));
            }
            counts[i] = r;
        }
        free(bx);
        free(by);
    }
}

/* ------------------------------------------------------------------ rrt.rs */

/* src/rrt.rs:43-60 + Polygon::new ring closing (geo-types 0.4: push first if last != first) */
long ppo_create_circle(double cx, double cy, double radius, double *rx, double *ry, size_t cap) {
    double circum = 2.0 * PI * radius;
    double n = ceil(circum / 1.0);
    if (!(n >= 0.0) || n > 1e8) return -1;
    size_t cnt = (size_t)(n + 1.0);
    if (cnt + 1 > cap) return -1;
    for (size_t k = 0; k < cnt; ++k) {
        double x = (double)k;
        rx[k] = (cos(2.0 * PI / n * x) * radius) + cx;
        ry[k] = (sin(2.0 * PI / n * x) * radius) + cy;
    }
    if (cnt > 0 && (rx[0] != rx[cnt - 1] || ry[0] != ry[cnt - 1])) {
        rx[cnt] = rx[0];
        ry[cnt] = ry[0];
        cnt += 1;
    }
    return (long)cnt;
}

/* src/rrt.rs:267-271 */
double ppo_compute_yaw(double fx, double fy, double tx, double ty) { return atan2(ty - fy, tx - fx); }

/* exact NN: intended contract of src/rrt.rs:378-391 */
void ppo_nn_brute(size_t n_nodes, const double *nx, const double *ny, size_t m, const double *qx,
                  const double *qy, uint32_t *idx, double *d2, size_t *hypot_disagree, int nthreads) {
    int nt = resolve_threads(nthreads);
    (void)nt;
    size_t disagree = 0;
#pragma omp parallel for schedule(static) num_threads(nt) reduction(+ : disagree)
    for (long j = 0; j < (long)m; ++j) {
        double best = INFINITY;
        uint32_t bi = 0xFFFFFFFFu;
        double x = qx[j], y = qy[j];
        for (size_t i = 0; i < n_nodes; ++i) {
            double dx = nx[i] - x, dy = ny[i] - y;
            double v = dx * dx + dy * dy;
            if (v < best) {
                best = v;
                bi = (uint32_t)i;
            }
        }
        if (hypot_disagree) { /* metric actually written in the reference: hypot, src/rrt.rs:239-246 */
            double hb = INFINITY;
            uint32_t hi = 0xFFFFFFFFu;
            for (size_t i = 0; i < n_nodes; ++i) {
                double h = hypot(x - nx[i], y - ny[i]);
                if (h < hb) {
                    hb = h;
                    hi = (uint32_t)i;
                }
            }
            if (hi != bi) disagree += 1;
        }
        idx[j] = bi;
        if (d2) d2[j] = best;
    }
    if (hypot_disagree) *hypot_disagree = disagree;
}

/* Exact NN through a uniform grid: ring-expanding search, same d2 arithmetic and tie-break.
 * Not in the reference; the CPU comparator for rstar's O(log N) query. */
void ppo_nn_grid(size_t n_nodes, const double *nx, const double *ny, size_t m, const double *qx, const double *qy,
                 uint32_t *idx, double *d2, int nthreads) {
    int nt = resolve_threads(nthreads);
    (void)nt;
    if (n_nodes == 0) {
        for (size_t j = 0; j < m; ++j) {
            idx[j] = 0xFFFFFFFFu;
            if (d2) d2[j] = INFINITY;
        }
        return;
    }
    double minx = nx[0], maxx = nx[0], miny = ny[0], maxy = ny[0];
    for (size_t i = 1; i < n_nodes; ++i) {
        if (nx[i] < minx) minx = nx[i];
        if (nx[i] > maxx) maxx = nx[i];
        if (ny[i] < miny) miny = ny[i];
        if (ny[i] > maxy) maxy = ny[i];
    }
    long g = (long)floor(sqrt((double)n_nodes / 2.0));
    if (g < 1) g = 1;
    if (g > 4096) g = 4096;
    double w = maxx - minx, h = maxy - miny;
    double cell = (w > h ? w : h) / (double)g;
    if (!(cell > 0.0)) cell = 1.0;
    double inv = 1.0 / cell;
    long gx = (long)floor(w * inv) + 1, gy = (long)floor(h * inv) + 1;
    uint32_t *start = (uint32_t *)calloc((size_t)(gx * gy + 1), sizeof(uint32_t));
    uint32_t *items = (uint32_t *)malloc(n_nodes * sizeof(uint32_t));
    uint32_t *cellof = (uint32_t *)malloc(n_nodes * sizeof(uint32_t));
    for (size_t i = 0; i < n_nodes; ++i) {
        long cx = (long)floor((nx[i] - minx) * inv), cy = (long)floor((ny[i] - miny) * inv);
        if (cx < 0) cx = 0;
        if (cx >= gx) cx = gx - 1;
        if (cy < 0) cy = 0;
        if (cy >= gy) cy = gy - 1;
        cellof[i] = (uint32_t)(cy * gx + cx);
        start[cellof[i] + 1]++;
    }
    for (long cidx = 0; cidx < gx * gy; ++cidx) start[cidx + 1] += start[cidx];
    uint32_t *fill = (uint32_t *)malloc((size_t)(gx * gy) * sizeof(uint32_t));
    memcpy(fill, start, (size_t)(gx * gy) * sizeof(uint32_t));
    for (size_t i = 0; i < n_nodes; ++i) items[fill[cellof[i]]++] = (uint32_t)i; /* ascending index per cell */
    free(fill);
    free(cellof);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (long j = 0; j < (long)m; ++j) {
        double x = qx[j], y = qy[j];
        long cx = (long)floor((x - minx) * inv), cy = (long)floor((y - miny) * inv);
        if (cx < 0) cx = 0;
        if (cx >= gx) cx = gx - 1;
        if (cy < 0) cy = 0;
        if (cy >= gy) cy = gy - 1;
        double best = INFINITY;
        uint32_t bi = 0xFFFFFFFFu;
        long maxr = (gx > gy ? gx : gy);
        for (long r = 0; r <= maxr; ++r) {
            /* every node outside the (2r-1)-cell square around the query cell is farther than
             * (r-1)*cell from the query: stop when the best so far is safely inside that. */
            if (r >= 2 && bi != 0xFFFFFFFFu) {
                double lim = (double)(r - 1) * cell * (1.0 - 1e-9);
                if (best < lim * lim) break;
            }
            long x0 = cx - r, x1 = cx + r, y0 = cy - r, y1 = cy + r;
            for (long yy = y0; yy <= y1; ++yy) {
                if (yy < 0 || yy >= gy) continue;
                int edge_row = (yy == y0 || yy == y1);
                for (long xx = x0; xx <= x1; xx += (edge_row ? 1 : (x1 - x0 > 0 ? x1 - x0 : 1))) {
                    if (xx < 0 || xx >= gx) continue;
                    uint32_t c0 = start[yy * gx + xx], c1 = start[yy * gx + xx + 1];
                    for (uint32_t k = c0; k < c1; ++k) {
                        uint32_t i = items[k];
                        double dx = nx[i] - x, dy = ny[i] - y;
                        double v = dx * dx + dy * dy;
                        if (v < best || (v == best && i < bi)) {
                            best = v;
                            bi = i;
                        }
                    }
                }
            }
        }
        idx[j] = bi;
        if (d2) d2[j] = best;
    }
    free(start);
    free(items);
}

/* ------------------------------------------------------------------ geo 0.12.2 predicates */

/* geo 0.12.2 `impl Contains<Point> for LineString` [RECALLED]: vertex equality, then per segment
 * tx/ty parameter test with f64::EPSILON */
int ppo_ring_has_point(const double *rx, const double *ry, size_t n, double px, double py) {
    if (n == 0) return 0;
    if (n == 1) return rx[0] == px && ry[0] == py;
    for (size_t i = 0; i < n; ++i)
        if (rx[i] == px && ry[i] == py) return 1;
    for (size_t i = 0; i + 1 < n; ++i) {
        double x0 = rx[i], y0 = ry[i];
        double dx = rx[i + 1] - x0, dy = ry[i + 1] - y0;
        int hit;
        if (dx == 0.0 && dy == 0.0) {
            hit = (px == x0 && py == y0);
        } else if (dy == 0.0) { /* (Some(t), None): horizontal */
            double t = (px - x0) / dx;
            hit = (py == y0 && 0.0 <= t && t <= 1.0);
        } else if (dx == 0.0) { /* (None, Some(t)): vertical */
            double t = (py - y0) / dy;
            hit = (px == x0 && 0.0 <= t && t <= 1.0);
        } else {
            double tx = (px - x0) / dx;
            double ty = (py - y0) / dy;
            hit = (fabs(tx - ty) <= 2.220446049250313e-16 && 0.0 <= tx && tx <= 1.0);
        }
        if (hit) return 1;
    }
    return 0;
}

/* geo 0.12.2 `get_position` [RECALLED]: 0 outside, 1 inside, 2 on boundary */
int ppo_point_position(const double *rx, const double *ry, size_t n, double px, double py) {
    if (n == 0) return 0;
    if (ppo_ring_has_point(rx, ry, n, px, py)) return 2;
    double xints = 0.0;
    long crossings = 0;
    for (size_t i = 0; i + 1 < n; ++i) {
        double x0 = rx[i], y0 = ry[i], x1 = rx[i + 1], y1 = ry[i + 1];
        double ymin = (y0 < y1) ? y0 : y1, ymax = (y0 > y1) ? y0 : y1; /* f64::min/max (no NaN here) */
        double xmax = (x0 > x1) ? x0 : x1;
        if (py > ymin && py <= ymax && px <= xmax) {
            if (y0 != y1) xints = (py - y0) * (x1 - x0) / (y1 - y0) + x0;
            if (x0 == x1 || px <= xints) crossings += 1;
        }
    }
    return (crossings % 2 == 1) ? 1 : 0;
}

/* geo 0.12.2 `impl Intersects<LineString> for LineString` [RECALLED]; a = self (ring), b = line */
int ppo_lines_intersect(const double *ax, const double *ay, size_t na, const double *bx, const double *by,
                        size_t nb) {
    if (na == 0 || nb == 0) return 0;
    for (size_t i = 0; i + 1 < na; ++i) {
        double a_dx = ax[i + 1] - ax[i], a_dy = ay[i + 1] - ay[i];
        for (size_t j = 0; j + 1 < nb; ++j) {
            double b_dx = bx[j + 1] - bx[j], b_dy = by[j + 1] - by[j];
            double u_b = b_dy * a_dx - b_dx * a_dy;
            if (u_b == 0.0) continue;
            double ua_t = b_dx * (ay[i] - by[j]) - b_dy * (ax[i] - bx[j]);
            double ub_t = a_dx * (ay[i] - by[j]) - a_dy * (ax[i] - bx[j]);
            double u_a = ua_t / u_b;
            double u_b2 = ub_t / u_b;
            if (0.0 <= u_a && u_a <= 1.0 && 0.0 <= u_b2 && u_b2 <= 1.0) return 1;
        }
    }
    return 0;
}

/* Polygon::contains(&Point) with no interiors */
static int poly_contains_point(const double *rx, const double *ry, size_t n, double px, double py) {
    return ppo_point_position(rx, ry, n, px, py) == 1;
}

/* src/rrt.rs:124-137 */
int ppo_verify(const ppo_world *w, const double *lx, const double *ly, size_t n) {
    /* bounds.contains(line): every point strictly inside (vacuously true for an empty line);
     * no interior rings to test */
    for (size_t k = 0; k < n; ++k)
        if (!poly_contains_point(w->bx, w->by, w->nb, lx[k], ly[k])) return 0;
    for (size_t r = 0; r < w->n_rings; ++r) {
        const double *rx = w->ox + w->ring_off[r], *ry = w->oy + w->ring_off[r];
        size_t rn = w->ring_off[r + 1] - w->ring_off[r];
        /* line.intersects(polygon) = polygon.intersects(line) */
        if (ppo_lines_intersect(rx, ry, rn, lx, ly, n)) return 0;
        for (size_t k = 0; k < n; ++k)
            if (poly_contains_point(rx, ry, rn, lx[k], ly[k])) return 0;
    }
    return 1;
}

/* ---- culled variant: must return exactly what ppo_verify returns (property-tested) ---- */
typedef struct {
    double minx, miny, maxx, maxy;
} aabb_t;

static aabb_t ring_aabb(const double *rx, const double *ry, size_t n) {
    aabb_t b = {INFINITY, INFINITY, -INFINITY, -INFINITY};
    for (size_t i = 0; i < n; ++i) {
        if (rx[i] < b.minx) b.minx = rx[i];
        if (rx[i] > b.maxx) b.maxx = rx[i];
        if (ry[i] < b.miny) b.miny = ry[i];
        if (ry[i] > b.maxy) b.maxy = ry[i];
    }
    return b;
}

/* conservative pad so that rounding in xints / parameter quotients cannot matter outside it */
static double aabb_pad(const aabb_t *b) {
    double m = fabs(b->minx);
    if (fabs(b->maxx) > m) m = fabs(b->maxx);
    if (fabs(b->miny) > m) m = fabs(b->miny);
    if (fabs(b->maxy) > m) m = fabs(b->maxy);
    return m * 0x1p-40 + 0x1p-1000;
}

/* circle filter, the second cull of the culled loop (same expressions as geo_predicates.cuh: pp_make_ring_circle /
 * pp_circle_class): per ring a centre, an inflated outer radius^2 no ring point exceeds and a deflated inner radius^2
 * whose disc lies inside the polygon.  class 0: the segment stays outside the outer circle (ring skipped); class 1:
 * both end points inside the inner circle (blocked); class 2: the exact geo predicates decide. */
typedef struct {
    double cx, cy, rout2, rin2;
} circle_t;

static circle_t ring_circle(const double *rx, const double *ry, size_t n, const aabb_t *b, double pad) {
    circle_t c;
    c.cx = c.cy = 0.0;
    c.rout2 = INFINITY;
    c.rin2 = 0.0;
    int finite = 1;
    for (size_t i = 0; i < n; ++i)
        if (!isfinite(rx[i]) || !isfinite(ry[i])) finite = 0;
    if (!finite || n < 3) return c;
    c.cx = 0.5 * (b->minx + b->maxx);
    c.cy = 0.5 * (b->miny + b->maxy);
    double rout = 0.0;
    for (size_t i = 0; i < n; ++i) {
        double d = hypot(rx[i] - c.cx, ry[i] - c.cy);
        if (d > rout) rout = d;
    }
    double ro = rout * (1.0 + 1.0e-6) + pad;
    c.rout2 = ro * ro;
    if (ppo_point_position(rx, ry, n, c.cx, c.cy) == 1) {
        double rin = INFINITY;
        for (size_t i = 0; i + 1 < n; ++i) {
            double x0 = rx[i], y0 = ry[i], dx = rx[i + 1] - x0, dy = ry[i + 1] - y0;
            double l2 = dx * dx + dy * dy;
            double t = (l2 > 0.0) ? ((c.cx - x0) * dx + (c.cy - y0) * dy) / l2 : 0.0;
            t = (t > 0.0) ? ((t < 1.0) ? t : 1.0) : 0.0;
            double d = hypot(c.cx - (x0 + t * dx), c.cy - (y0 + t * dy));
            if (d < rin) rin = d;
        }
        double ri = rin * (1.0 - 1.0e-6) - pad;
        if (ri > 0.0 && ri < INFINITY) c.rin2 = ri * ri;
    }
    return c;
}

static int circle_class(const circle_t *c, double ax, double ay, double bx, double by) {
    double acx = c->cx - ax, acy = c->cy - ay;
    double a2 = acx * acx + acy * acy;
    double bcx = c->cx - bx, bcy = c->cy - by;
    double b2 = bcx * bcx + bcy * bcy;
    if (a2 < c->rin2 && b2 < c->rin2) return 1;
    if (!(a2 > c->rout2) || !(b2 > c->rout2)) return 2;
    if (!(a2 < 1.0e6 * c->rout2) || !(b2 < 1.0e6 * c->rout2)) return 2;
    double abx = bx - ax, aby = by - ay;
    double e = acx * abx + acy * aby;
    double f = abx * abx + aby * aby;
    if (!(e > 0.0) || !(e < f)) return 0;
    return (a2 * f - e * e > c->rout2 * f) ? 0 : 2;
}

/* harness entry: the class the culled loop assigns to segment a-b against one ring (0 skip, 1 blocked, 2 exact) */
int ppo_circle_class(const double *rx, const double *ry, size_t n, double ax, double ay, double bx, double by) {
    aabb_t b = ring_aabb(rx, ry, n);
    circle_t c = ring_circle(rx, ry, n, &b, n ? aabb_pad(&b) : 0.0);
    return circle_class(&c, ax, ay, bx, by);
}

int ppo_verify_culled(const ppo_world *w, const double *lx, const double *ly, size_t n) {
    for (size_t k = 0; k < n; ++k)
        if (!poly_contains_point(w->bx, w->by, w->nb, lx[k], ly[k])) return 0;
    for (size_t r = 0; r < w->n_rings; ++r) {
        const double *rx = w->ox + w->ring_off[r], *ry = w->oy + w->ring_off[r];
        size_t rn = w->ring_off[r + 1] - w->ring_off[r];
        aabb_t b = ring_aabb(rx, ry, rn);
        double pad = aabb_pad(&b);
        /* quick reject of the whole line against the ring's padded box before the ring's circle is set up */
        int near = 0;
        for (size_t k = 0; k < n && !near; ++k) {
            size_t k1 = (k + 1 < n) ? k + 1 : k;
            double sminx = lx[k] < lx[k1] ? lx[k] : lx[k1], smaxx = lx[k] > lx[k1] ? lx[k] : lx[k1];
            double sminy = ly[k] < ly[k1] ? ly[k] : ly[k1], smaxy = ly[k] > ly[k1] ? ly[k] : ly[k1];
            if (!(smaxx < b.minx - pad || sminx > b.maxx + pad || smaxy < b.miny - pad || sminy > b.maxy + pad)) near = 1;
        }
        if (!near && n > 0) continue;
        circle_t c = ring_circle(rx, ry, rn, &b, rn ? pad : 0.0);
        for (size_t j = 0; j + 1 < n; ++j) {
            double sminx = lx[j] < lx[j + 1] ? lx[j] : lx[j + 1], smaxx = lx[j] > lx[j + 1] ? lx[j] : lx[j + 1];
            double sminy = ly[j] < ly[j + 1] ? ly[j] : ly[j + 1], smaxy = ly[j] > ly[j + 1] ? ly[j] : ly[j + 1];
            if (smaxx < b.minx - pad || sminx > b.maxx + pad || smaxy < b.miny - pad || sminy > b.maxy + pad)
                continue;
            int cls = circle_class(&c, lx[j], ly[j], lx[j + 1], ly[j + 1]);
            if (cls == 0) continue;
            if (cls == 1) return 0;
            if (ppo_lines_intersect(rx, ry, rn, lx + j, ly + j, 2)) return 0;
        }
        for (size_t k = 0; k < n; ++k) {
            double x = lx[k], y = ly[k];
            if (x < b.minx - pad || x > b.maxx + pad || y < b.miny - pad || y > b.maxy + pad) continue;
            int cls = circle_class(&c, x, y, x, y);
            if (cls == 0) continue;
            if (cls == 1) return 0;
            if (poly_contains_point(rx, ry, rn, x, y)) return 0;
        }
    }
    return 1;
}

void ppo_verify_segments(const ppo_world *w, size_t m, const double *ax, const double *ay, const double *bx,
                         const double *by, uint8_t *ok, int culled, int nthreads) {
    int nt = resolve_threads(nthreads);
    (void)nt;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
    for (long i = 0; i < (long)m; ++i) {
        double lx[2] = {ax[i], bx[i]}, ly[2] = {ay[i], by[i]};
        ok[i] = (uint8_t)(culled ? ppo_verify_culled(w, lx, ly, 2) : ppo_verify(w, lx, ly, 2));
    }
}

/* per-edge polyline of line_to_origin (src/rrt.rs:295-317) in node->root order:
 * samples(node->parent) followed by the parent's own point (= first sample of the next chunk, Q12). */
long ppo_dubins_edge_polyline(double sx, double sy, double syaw, double ex, double ey, double eyaw,
                              double radius, double step, double *lx, double *ly, size_t cap) {
    if (cap < 2) return -2;
    long n = ppo_dubins_path(sx, sy, syaw, ex, ey, eyaw, radius, step, 0, lx, ly, NULL, cap - 1, NULL, NULL, NULL);
    if (n == -1) { /* src/rrt.rs:313 fallback */
        lx[0] = sx;
        ly[0] = sy;
        n = 1;
    } else if (n < 0) {
        return n;
    }
    lx[n] = ex;
    ly[n] = ey;
    return n + 1;
}

void ppo_verify_dubins_edges(const ppo_world *w, size_t m, const double *sx, const double *sy, const double *syaw,
                             const double *ex, const double *ey, const double *eyaw, double radius, double step,
                             uint8_t *ok, int culled, int nthreads) {
    int nt = resolve_threads(nthreads);
    (void)nt;
#pragma omp parallel num_threads(nt)
    {
        size_t cap = 1 << 16;
        double *lx = (double *)malloc(cap * sizeof(double)), *ly = (double *)malloc(cap * sizeof(double));
#pragma omp for schedule(dynamic, 16)
        for (long i = 0; i < (long)m; ++i) {
            long n;
            for (;;) {
                n = ppo_dubins_edge_polyline(sx[i], sy[i], syaw[i], ex[i], ey[i], eyaw[i], radius, step, lx, ly, cap);
                if (n != -2) break;
                cap *= 2;
                free(lx);
                free(ly);
                lx = (double *)malloc(cap * sizeof(double));
                ly = (double *)malloc(cap * sizeof(double));
            }
            if (n < 0)
                ok[i] = 0xFF; /* reference would panic */
            else
                ok[i] = (uint8_t)(culled ? ppo_verify_culled(w, lx, ly, (size_t)n) : ppo_verify(w, lx, ly, (size_t)n));
        }
        free(lx);
        free(ly);
    }
}

/* src/rrt.rs:291-321 with deterministic node->root chunk order (Q12) */
long ppo_line_to_origin(const double *nx, const double *ny, const double *nyaw, const int32_t *parent,
                        uint32_t node, double radius, double step, double *lx, double *ly, size_t cap) {
    size_t n = 0;
    int64_t cur = node;
    for (;;) {
        int32_t par = parent[cur];
        if (par < 0) { /* :316 root contributes its own point */
            if (n + 1 > cap) return -2;
            lx[n] = nx[cur];
            ly[n] = ny[cur];
            n += 1;
            break;
        }
        long k = ppo_dubins_path(nx[cur], ny[cur], nyaw[cur], nx[par], ny[par], nyaw[par], radius, step, 0, lx + n,
                                 ly + n, NULL, cap - n, NULL, NULL, NULL);
        if (k == -1) { /* :313 */
            if (n + 1 > cap) return -2;
            lx[n] = nx[cur];
            ly[n] = ny[cur];
            k = 1;
        } else if (k < 0) {
            return k;
        }
        n += (size_t)k;
        cur = par;
    }
    return (long)n;
}

/* ------------------------------------------------------------------ harness: decision margins */

static double dist_point_seg(double px, double py, double x0, double y0, double x1, double y1) {
    double dx = x1 - x0, dy = y1 - y0;
    double l2 = dx * dx + dy * dy;
    double t = (l2 > 0.0) ? ((px - x0) * dx + (py - y0) * dy) / l2 : 0.0;
    if (!(t > 0.0)) t = 0.0; /* also catches NaN */
    if (t > 1.0) t = 1.0;
    return hypot(px - (x0 + t * dx), py - (y0 + t * dy));
}

/* true geometric crossing test (robust enough for a margin: orientation signs, touching counts) */
static int seg_cross(double ax, double ay, double bx, double by, double cx, double cy, double dx, double dy) {
    double d1 = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
    double d2 = (bx - ax) * (dy - ay) - (by - ay) * (dx - ax);
    double d3 = (dx - cx) * (ay - cy) - (dy - cy) * (ax - cx);
    double d4 = (dx - cx) * (by - cy) - (dy - cy) * (bx - cx);
    return ((d1 > 0.0) != (d2 > 0.0) || d1 == 0.0 || d2 == 0.0) && ((d3 > 0.0) != (d4 > 0.0) || d3 == 0.0 || d4 == 0.0);
}

static double dist_seg_seg(double ax, double ay, double bx, double by, double cx, double cy, double dx, double dy) {
    if (seg_cross(ax, ay, bx, by, cx, cy, dx, dy)) {
        /* the orientation test also fires for collinear disjoint segments; the end-point distances decide then */
        double d1 = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax), d2 = (bx - ax) * (dy - ay) - (by - ay) * (dx - ax);
        if (!(d1 == 0.0 && d2 == 0.0)) return 0.0;
    }
    double m = dist_point_seg(ax, ay, cx, cy, dx, dy);
    double v = dist_point_seg(bx, by, cx, cy, dx, dy);
    if (v < m) m = v;
    v = dist_point_seg(cx, cy, ax, ay, bx, by);
    if (v < m) m = v;
    v = dist_point_seg(dx, dy, ax, ay, bx, by);
    if (v < m) m = v;
    return m;
}

/* distance from a point to the boundary of a ring */
static double dist_point_ring(const double *rx, const double *ry, size_t n, double px, double py) {
    double m = INFINITY;
    if (n == 1) return hypot(px - rx[0], py - ry[0]);
    for (size_t i = 0; i + 1 < n; ++i) {
        double v = dist_point_seg(px, py, rx[i], ry[i], rx[i + 1], ry[i + 1]);
        if (v < m) m = v;
    }
    return m;
}

/* Space::verify (src/rrt.rs:124-137) together with a lower bound on how far every vertex of the line may move
 * without changing the verdict.
 *   verdict 1 (free): margin = the smallest distance between the line and any ring boundary (bounds included).
 *     Moving every vertex by less than that cannot create a contact, and without a contact no vertex changes side.
 *   verdict 0 (blocked): margin = the largest depth by which a point ON the line lies inside an obstacle or
 *     outside the bounds (vertices and PPO_MARGIN_SUB interior points per segment are probed).  A line whose
 *     vertices moved by less than that still passes within the obstacle / outside the bounds there.
 * Not part of the reference; only used to CLASSIFY verdict differences between two implementations whose
 * sample coordinates agree to a tolerance (margin < tolerance => "near graze"). */
#define PPO_MARGIN_SUB 4
static int verify_margin(const ppo_world *w, const double *lx, const double *ly, size_t n, double *margin, int culled);
int ppo_verify_margin(const ppo_world *w, const double *lx, const double *ly, size_t n, double *margin) {
    return verify_margin(w, lx, ly, n, margin, 0);
}
/* culled != 0: the verdict comes from ppo_verify_culled (same answer up to the near-parallel noise class, and
 * O(rings) instead of O(ring segments x line segments) -- for worlds of 1e5 rings) */
static int verify_margin(const ppo_world *w, const double *lx, const double *ly, size_t n, double *margin, int culled) {
    int verdict = culled ? ppo_verify_culled(w, lx, ly, n) : ppo_verify(w, lx, ly, n);
    double m;
    /* box of the whole line: only rings near it can matter (their boxes are computed once per call) */
    aabb_t lb = {INFINITY, INFINITY, -INFINITY, -INFINITY};
    for (size_t j = 0; j < n; ++j) {
        if (lx[j] < lb.minx) lb.minx = lx[j];
        if (lx[j] > lb.maxx) lb.maxx = lx[j];
        if (ly[j] < lb.miny) lb.miny = ly[j];
        if (ly[j] > lb.maxy) lb.maxy = ly[j];
    }
    if (verdict) {
        /* the bounds ring first: its distance bounds how far away an obstacle ring can still matter */
        m = INFINITY;
        for (size_t pass = 0; pass <= w->n_rings; ++pass) {
            const double *rx, *ry;
            size_t rn;
            if (pass == 0) {
                rx = w->bx, ry = w->by, rn = w->nb;
            } else {
                size_t r = pass - 1;
                rx = w->ox + w->ring_off[r], ry = w->oy + w->ring_off[r], rn = w->ring_off[r + 1] - w->ring_off[r];
            }
            if (rn == 0) continue;
            aabb_t b = ring_aabb(rx, ry, rn);
            if (lb.maxx < b.minx - m || lb.minx > b.maxx + m || lb.maxy < b.miny - m || lb.miny > b.maxy + m) continue;
            for (size_t j = 0; j < n; ++j) {
                size_t j1 = (j + 1 < n) ? j + 1 : j; /* last vertex: degenerate segment */
                double sminx = lx[j] < lx[j1] ? lx[j] : lx[j1], smaxx = lx[j] > lx[j1] ? lx[j] : lx[j1];
                double sminy = ly[j] < ly[j1] ? ly[j] : ly[j1], smaxy = ly[j] > ly[j1] ? ly[j] : ly[j1];
                if (smaxx < b.minx - m || sminx > b.maxx + m || smaxy < b.miny - m || sminy > b.maxy + m) continue;
                if (rn == 1) {
                    double v = dist_point_seg(rx[0], ry[0], lx[j], ly[j], lx[j1], ly[j1]);
                    if (v < m) m = v;
                }
                for (size_t i = 0; i + 1 < rn; ++i) {
                    double v = dist_seg_seg(lx[j], ly[j], lx[j1], ly[j1], rx[i], ry[i], rx[i + 1], ry[i + 1]);
                    if (v < m) m = v;
                }
            }
        }
    } else {
        m = 0.0;
        /* obstacle rings whose box meets the line's box: a probe point can only be inside one of those */
        size_t n_cand = 0;
        uint32_t *cand = (uint32_t *)malloc((w->n_rings ? w->n_rings : 1) * sizeof(uint32_t));
        aabb_t *cbox = (aabb_t *)malloc((w->n_rings ? w->n_rings : 1) * sizeof(aabb_t));
        for (size_t r = 0; r < w->n_rings && cand && cbox; ++r) {
            size_t rn = w->ring_off[r + 1] - w->ring_off[r];
            if (rn < 3) continue;
            aabb_t b = ring_aabb(w->ox + w->ring_off[r], w->oy + w->ring_off[r], rn);
            if (lb.maxx < b.minx || lb.minx > b.maxx || lb.maxy < b.miny || lb.miny > b.maxy) continue;
            cbox[n_cand] = b;
            cand[n_cand++] = (uint32_t)r;
        }
        for (size_t j = 0; j < n; ++j) {
            size_t j1 = (j + 1 < n) ? j + 1 : j;
            int subs = (j1 == j) ? 1 : PPO_MARGIN_SUB + 1;
            for (int k = 0; k < subs; ++k) {
                double t = (double)k / (double)(PPO_MARGIN_SUB + 1);
                double qx = lx[j] + t * (lx[j1] - lx[j]), qy = ly[j] + t * (ly[j1] - ly[j]);
                if (!(qx == qx) || !(qy == qy)) { /* a NaN coordinate is outside everything for every implementation */
                    m = INFINITY;
                    continue;
                }
                if (ppo_point_position(w->bx, w->by, w->nb, qx, qy) == 0) {
                    double v = w->nb ? dist_point_ring(w->bx, w->by, w->nb, qx, qy) : INFINITY;
                    if (v > m) m = v;
                }
                for (size_t c = 0; c < n_cand; ++c) {
                    const aabb_t b = cbox[c];
                    if (qx < b.minx || qx > b.maxx || qy < b.miny || qy > b.maxy) continue;
                    size_t r = cand[c];
                    const double *rx = w->ox + w->ring_off[r], *ry = w->oy + w->ring_off[r];
                    size_t rn = w->ring_off[r + 1] - w->ring_off[r];
                    if (ppo_point_position(rx, ry, rn, qx, qy) == 1) {
                        double v = dist_point_ring(rx, ry, rn, qx, qy);
                        if (v > m) m = v;
                    }
                }
            }
        }
        free(cand);
        free(cbox);
    }
    if (margin) *margin = m;
    return verdict;
}

/* Which flags make a verify DECISION fragile, given the verdict's margin: a possible other word / wrap / huge
 * angle changes the whole path; a knife-edge sample count adds or drops one vertex, which moves the polyline by
 * at most the sagitta of two sample spacings, (2 s)^2 / (8 R) = step^2 R / 2 with s = step R -- only a margin
 * below twice that can be affected; a margin below graze_tol * scale is fragile in any case. */
static uint32_t decision_flags(uint32_t path_flags, double margin, double scale, double graze_tol, double radius,
                               double step) {
    uint32_t f = path_flags & (PPO_FLAG_NEAR_WRAP | PPO_FLAG_NEAR_TIE | PPO_FLAG_NEAR_FEAS | PPO_FLAG_HUGE_ANGLE);
    double graze = graze_tol * scale;
    if (!(margin >= graze)) f |= PPO_FLAG_NEAR_GRAZE;
    if ((path_flags & PPO_FLAG_NEAR_COUNT) && !(margin >= step * step * radius + graze)) f |= PPO_FLAG_NEAR_COUNT;
    return f;
}

/* verify of Dubins edges with the classification flags of the parity harness: the Dubins flags of the edge's
 * path (near wrap / tie / feasibility, knife-edge sample count, huge angles) and PPO_FLAG_NEAR_GRAZE when the
 * verdict's margin is below graze_tol * max(1, largest |coordinate| of the polyline). */
void ppo_verify_dubins_edges_flags(const ppo_world *w, size_t m, const double *sx, const double *sy,
                                   const double *syaw, const double *ex, const double *ey, const double *eyaw,
                                   double radius, double step, double graze_tol, uint8_t *ok, uint32_t *flags,
                                   double *margins, int culled, int nthreads) {
    int nt = resolve_threads(nthreads);
    (void)nt;
#pragma omp parallel num_threads(nt)
    {
        size_t cap = 1 << 16;
        double *lx = (double *)malloc(cap * sizeof(double)), *ly = (double *)malloc(cap * sizeof(double));
#pragma omp for schedule(dynamic, 4)
        for (long i = 0; i < (long)m; ++i) {
            long n;
            uint32_t f = 0;
            for (;;) {
                n = ppo_dubins_path_flags(sx[i], sy[i], syaw[i], ex[i], ey[i], eyaw[i], radius, step, 0, lx, ly, NULL,
                                          cap - 1, NULL, NULL, &f);
                if (n != -2) break;
                cap *= 2;
                free(lx);
                free(ly);
                lx = (double *)malloc(cap * sizeof(double));
                ly = (double *)malloc(cap * sizeof(double));
            }
            if (n == -1) { /* src/rrt.rs:313 */
                lx[0] = sx[i];
                ly[0] = sy[i];
                n = 1;
            }
            if (n < 0) {
                ok[i] = 0xFF;
                flags[i] = f;
                if (margins) margins[i] = 0.0;
                continue;
            }
            lx[n] = ex[i];
            ly[n] = ey[i];
            n += 1;
            double mg = 0.0, scale = 1.0;
            for (long k = 0; k < n; ++k) {
                if (fabs(lx[k]) > scale) scale = fabs(lx[k]);
                if (fabs(ly[k]) > scale) scale = fabs(ly[k]);
            }
            ok[i] = (uint8_t)verify_margin(w, lx, ly, (size_t)n, &mg, culled);
            flags[i] = decision_flags(f, mg, scale, graze_tol, radius, step);
            if (margins) margins[i] = mg;
        }
        free(lx);
        free(ly);
    }
}

/* ------------------------------------------------------------------ rrt.rs: goal check and shortcutting */

/* growable node arena: slots [0, n_tree) are the tree, later slots are the Node values optimize() creates
 * (Arc<Node> in the reference; a slot index here) */
typedef struct {
    double *x, *y, *yaw;
    int32_t *parent;
    size_t n, cap;
    const ppo_world *w;
    double radius, step, graze_tol;
    uint32_t flags;      /* OR of the decision flags (decision_flags) of every verify decision taken */
    uint32_t line_flags; /* OR of the raw path flags of the edges of the final line (sample-level differences) */
    long verifies;       /* Space::verify calls made (the reference's cost driver) */
    double *lx, *ly;     /* polyline scratch */
    size_t lcap;
    int error;
} arena_t;

static int arena_push(arena_t *a, double x, double y, double yaw, int32_t parent) {
    if (a->n == a->cap) {
        size_t cap = a->cap * 2;
        double *nx = (double *)realloc(a->x, cap * sizeof(double));
        double *ny = (double *)realloc(a->y, cap * sizeof(double));
        double *nyaw = (double *)realloc(a->yaw, cap * sizeof(double));
        int32_t *np = (int32_t *)realloc(a->parent, cap * sizeof(int32_t));
        if (nx) a->x = nx;
        if (ny) a->y = ny;
        if (nyaw) a->yaw = nyaw;
        if (np) a->parent = np;
        if (!nx || !ny || !nyaw || !np) {
            a->error = 1;
            return -1;
        }
        a->cap = cap;
    }
    a->x[a->n] = x;
    a->y[a->n] = y;
    a->yaw[a->n] = yaw;
    a->parent[a->n] = parent;
    return (int)a->n++;
}

/* Node::new, src/rrt.rs:169-175: yaw = compute_yaw(point, parent.point) */
static int arena_node_new(arena_t *a, double x, double y, int32_t parent) {
    return arena_push(a, x, y, ppo_compute_yaw(x, y, a->x[parent], a->y[parent]), parent);
}

/* verify(line_to_origin(node)) with the harness flags; src/rrt.rs:414-426 as used at :476-477 */
static int arena_verify_chain(arena_t *a, int32_t node) {
    long n;
    uint32_t pf = 0;
    for (;;) {
        pf = 0;
        /* same chunks as ppo_line_to_origin, plus the Dubins flags of every edge */
        size_t cnt = 0;
        int32_t cur = node;
        n = 0;
        for (;;) {
            int32_t par = a->parent[cur];
            if (par < 0) {
                if (cnt + 1 > a->lcap) {
                    n = -2;
                    break;
                }
                a->lx[cnt] = a->x[cur];
                a->ly[cnt] = a->y[cur];
                cnt += 1;
                break;
            }
            uint32_t f = 0;
            long k = ppo_dubins_path_flags(a->x[cur], a->y[cur], a->yaw[cur], a->x[par], a->y[par], a->yaw[par], a->radius,
                                           a->step, 0, a->lx + cnt, a->ly + cnt, NULL, a->lcap - cnt, NULL, NULL, &f);
            if (k == -2) {
                n = -2;
                break;
            }
            pf |= f;
            if (k == -1) { /* :313 */
                if (cnt + 1 > a->lcap) {
                    n = -2;
                    break;
                }
                a->lx[cnt] = a->x[cur];
                a->ly[cnt] = a->y[cur];
                k = 1;
            } else if (k < 0) {
                a->error = 1;
                return 0;
            }
            cnt += (size_t)k;
            cur = par;
        }
        if (n != -2) {
            n = (long)cnt;
            break;
        }
        size_t cap = a->lcap * 2;
        free(a->lx);
        free(a->ly);
        a->lx = (double *)malloc(cap * sizeof(double));
        a->ly = (double *)malloc(cap * sizeof(double));
        a->lcap = cap;
    }
    double mg = 0.0, scale = 1.0;
    for (long k = 0; k < n; ++k) {
        if (fabs(a->lx[k]) > scale) scale = fabs(a->lx[k]);
        if (fabs(a->ly[k]) > scale) scale = fabs(a->ly[k]);
    }
    int v = ppo_verify_margin(a->w, a->lx, a->ly, (size_t)n, &mg);
    a->verifies += 1;
    a->flags |= decision_flags(pf, mg, scale, a->graze_tol, a->radius, a->step);
    return v;
}

#define PPO_RECURSION_LIMIT 16 /* src/rrt.rs:14 */

/* RRT::optimize, src/rrt.rs:463-487.  Returns the arena slot of the new node or -1 (None). */
static int32_t arena_optimize(arena_t *a, int32_t node, size_t i) {
    if (i >= PPO_RECURSION_LIMIT) return -1; /* :464-466 */
    size_t depth = 0;
    for (int32_t c = node; c >= 0; c = a->parent[c]) depth += 1;
    int32_t *nodes_vec = (int32_t *)malloc(depth * sizeof(int32_t)); /* :468-471 node, parent, ..., root */
    depth = 0;
    for (int32_t c = node; c >= 0; c = a->parent[c]) nodes_vec[depth++] = c;
    int32_t result = -1;
    for (size_t k = depth; k-- > 0;) { /* :473 .rev(): root first, the node itself last */
        int32_t to_node = nodes_vec[k];
        size_t mark = a->n;
        int32_t new_node = arena_node_new(a, a->x[node], a->y[node], to_node); /* :474 */
        if (new_node < 0) break;
        if (arena_verify_chain(a, new_node)) { /* :476-477 */
            int32_t deeper = arena_optimize(a, to_node, i + 1); /* :478 */
            result = (deeper >= 0) ? arena_node_new(a, a->x[node], a->y[node], deeper) /* :479-481 */
                                   : new_node;                                        /* :482 */
            break;
        }
        a->n = mark; /* nothing refers to a rejected candidate */
        if (a->error) break;
    }
    free(nodes_vec);
    return result;
}

static int arena_init(arena_t *a, const ppo_world *w, size_t n_nodes, const double *nx, const double *ny,
                      const double *nyaw, const int32_t *parent, double radius, double step, double graze_tol) {
    memset(a, 0, sizeof *a);
    a->cap = n_nodes + 64;
    a->x = (double *)malloc(a->cap * sizeof(double));
    a->y = (double *)malloc(a->cap * sizeof(double));
    a->yaw = (double *)malloc(a->cap * sizeof(double));
    a->parent = (int32_t *)malloc(a->cap * sizeof(int32_t));
    a->lcap = 1 << 14;
    a->lx = (double *)malloc(a->lcap * sizeof(double));
    a->ly = (double *)malloc(a->lcap * sizeof(double));
    if (!a->x || !a->y || !a->yaw || !a->parent || !a->lx || !a->ly) return -1;
    memcpy(a->x, nx, n_nodes * sizeof(double));
    memcpy(a->y, ny, n_nodes * sizeof(double));
    memcpy(a->yaw, nyaw, n_nodes * sizeof(double));
    memcpy(a->parent, parent, n_nodes * sizeof(int32_t));
    a->n = n_nodes;
    a->w = w;
    a->radius = radius;
    a->step = step;
    a->graze_tol = graze_tol;
    return 0;
}
static void arena_free(arena_t *a) {
    free(a->x);
    free(a->y);
    free(a->yaw);
    free(a->parent);
    free(a->lx);
    free(a->ly);
}

/* writes the chain node -> root as poses; returns its length or -2 when cap is too small */
static long arena_chain_out(const arena_t *a, int32_t node, double *cx, double *cy, double *cyaw, size_t cap) {
    size_t k = 0;
    for (int32_t c = node; c >= 0; c = a->parent[c]) {
        if (k >= cap) return -2;
        cx[k] = a->x[c];
        cy[k] = a->y[c];
        cyaw[k] = a->yaw[c];
        k += 1;
    }
    return (long)k;
}

long ppo_optimize(const ppo_world *w, size_t n_nodes, const double *nx, const double *ny, const double *nyaw,
                  const int32_t *parent, uint32_t node, double radius, double step, double graze_tol, double *cx,
                  double *cy, double *cyaw, size_t cap, uint32_t *flags, long *verifies) {
    arena_t a;
    if (node >= n_nodes || arena_init(&a, w, n_nodes, nx, ny, nyaw, parent, radius, step, graze_tol)) return -3;
    int32_t r = arena_optimize(&a, (int32_t)node, 0);
    long out = 0;
    if (a.error)
        out = -3;
    else if (r >= 0)
        out = arena_chain_out(&a, r, cx, cy, cyaw, cap);
    if (flags) *flags = a.flags;
    if (verifies) *verifies = a.verifies;
    arena_free(&a);
    return out;
}

/* RRT::check_finish (src/rrt.rs:428-438) = new_goal over `node`, finalize (:503-540, via optimize_from_goal
 * :489-501), verify.  Returns the number of points of the final line (start -> goal order), -1 for None
 * (the line does not verify; the line itself is still written), -2 when a capacity is too small, -3 when the
 * reference would panic (:529).  The optimised chain goal -> root is written to (cx, cy, cyaw). */
long ppo_check_finish(const ppo_world *w, size_t n_nodes, const double *nx, const double *ny, const double *nyaw,
                      const int32_t *parent, uint32_t node, double gx, double gy, double gyaw, double radius,
                      double step, double graze_tol, double *lx, double *ly, size_t cap, long *line_len, double *cx,
                      double *cy, double *cyaw, size_t ccap, long *chain_len, uint32_t *flags, uint32_t *line_flags,
                      long *verifies) {
    arena_t a;
    if (node >= n_nodes || arena_init(&a, w, n_nodes, nx, ny, nyaw, parent, radius, step, graze_tol)) return -3;
    long ret = -3;
    /* :429-430 Node::new_goal(point, node, goal_yaw) */
    int32_t goal_node = arena_push(&a, gx, gy, gyaw, (int32_t)node);
    /* :489-501 optimize_from_goal */
    int32_t top = goal_node;
    int32_t opt = arena_optimize(&a, (int32_t)node, 0);
    if (opt >= 0) top = arena_push(&a, gx, gy, gyaw, opt);
    if (a.error || goal_node < 0 || top < 0) goto done;
    if (chain_len) {
        *chain_len = arena_chain_out(&a, top, cx, cy, cyaw, ccap);
        if (*chain_len == -2) {
            ret = -2;
            goto done;
        }
    }
    /* :503-540 finalize: per node with a parent the Dubins samples, the root contributes nothing; chunks in
     * node -> root order (Q12), then the whole vector reversed */
    size_t n = 0;
    for (int32_t cur = top; a.parent[cur] >= 0; cur = a.parent[cur]) {
        int32_t par = a.parent[cur];
        uint32_t f = 0;
        long k = ppo_dubins_path_flags(a.x[cur], a.y[cur], a.yaw[cur], a.x[par], a.y[par], a.yaw[par], radius, step, 0,
                                       lx + n, ly + n, NULL, cap - n, NULL, NULL, &f);
        a.line_flags |= f;
        if (k == -2) {
            ret = -2;
            goto done;
        }
        if (k < 0) { /* :529 panic!("Should plan dubins curve") or an out-of-buffer index */
            ret = -3;
            goto done;
        }
        n += (size_t)k;
    }
    for (size_t i = 0, j = n; i + 1 < j; ++i) { /* :537 l.reverse() */
        --j;
        double t = lx[i];
        lx[i] = lx[j];
        lx[j] = t;
        t = ly[i];
        ly[i] = ly[j];
        ly[j] = t;
    }
    if (line_len) *line_len = (long)n;
    {
        double mg = 0.0, scale = 1.0;
        for (size_t k = 0; k < n; ++k) {
            if (fabs(lx[k]) > scale) scale = fabs(lx[k]);
            if (fabs(ly[k]) > scale) scale = fabs(ly[k]);
        }
        int v = ppo_verify_margin(w, lx, ly, n, &mg); /* :433 */
        a.verifies += 1;
        a.flags |= decision_flags(a.line_flags, mg, scale, graze_tol, radius, step);
        ret = v ? (long)n : -1;
    }
done:
    if (flags) *flags = a.flags;
    if (line_flags) *line_flags = a.line_flags;
    if (verifies) *verifies = a.verifies;
    arena_free(&a);
    return ret;
}

/* ------------------------------------------------------------------ synthetic inputs */

double ppo_uniform(uint64_t seed, uint64_t stream, uint64_t i) {
    uint64_t z = seed + (stream << 56) + (i + 1) * 0x9E3779B97F4A7C15ull;
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * 0x1p-53;
}

void ppo_fill_uniform(uint64_t seed, uint64_t stream, size_t n, double lo, double hi, double *out) {
    double span = hi - lo;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) {
        double u = ppo_uniform(seed, stream, (uint64_t)i);
        double s = span * u;
        out[i] = lo + s;
    }
}

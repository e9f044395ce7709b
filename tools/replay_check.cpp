// replay_check.cpp -- pp_replay.cuh (the exact fast-forward of generate_local_course's sample loop, src/dubins.rs:239-255)
// against the literal loop: iteration count AND final pd must be the same bits.
//   g++ -O2 -std=c++17 -ffp-contract=off tools/replay_check.cpp -o replay_check && ./replay_check 2000000
// Prints "<class> <mismatches>" lines; every count must be 0.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "../rs-pathplanning_b200/csrc/pp_replay.cuh"

static const uint32_t MAXI = 1u << 22;

static bool literal(double pd, double d, double al, uint32_t *cnt_out, double *pd_out) {
    uint32_t cnt = 0;
    while (fabs(pd) <= al) {
        pd += d;
        if (++cnt >= MAXI) {
            *cnt_out = MAXI;
            return false;
        }
    }
    *cnt_out = cnt;
    *pd_out = pd;
    return true;
}

static long check(double pd, double d, double al) {
    uint32_t c0 = 0, c1 = 0;
    double p0 = 0, p1 = 0;
    const bool ok0 = literal(pd, d, al, &c0, &p0), ok1 = pp_replay_segment(pd, d, al, MAXI, &c1, &p1);
    if (ok0 != ok1 || c0 != c1) return 1;
    if (ok0 && memcmp(&p0, &p1, 8) != 0 && !(p0 != p0 && p1 != p1)) return 1;  // same bits, or NaN on both sides
    return 0;
}

int main(int argc, char **argv) {
    const long n = argc > 1 ? atol(argv[1]) : 1000000;
    std::mt19937_64 rng(0x5EED5EEDull);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    long bad_call_pattern = 0, bad_steps = 0, bad_random = 0, bad_special = 0, bad_tiny = 0;
    const double steps[] = {0.05, 0.1, 0.01, 0.3, 0.25, 0.125, 1.0 / 3.0, 0.7, 1e-3, 0x1.fffffffffffffp-5, 0x1.0000000000001p-4};
    // (i) the reference's call pattern: d = +-step, pd0 in (0, 3 d], l up to ~100 turn radii
    for (long i = 0; i < n; ++i) {
        const double step = steps[rng() % (sizeof steps / sizeof steps[0])];
        const double l = U(rng) * ((rng() & 7) ? 12.0 : 120.0);
        const double pd0 = step * (U(rng) * 3.0);
        bad_call_pattern += check(pd0, step, l);
    }
    // (ii) arbitrary steps (full mantissas), either sign, pd0 of either sign
    for (long i = 0; i < n; ++i) {
        const double step = ldexp(0.5 + 0.5 * U(rng), -(int)(rng() % 12));
        const double sgn = (rng() & 1) ? 1.0 : -1.0;
        const double l = U(rng) * ldexp(1.0, (int)(rng() % 9));
        const double pd0 = (U(rng) * 4.0 - 1.0) * step * sgn;
        bad_steps += check(pd0, step * sgn, l);
    }
    // (iii) wild: limits just at iterates, limits equal to binade tops, steps with one-bit tails (tie binades)
    for (long i = 0; i < n / 4; ++i) {
        double step = ldexp((double)((rng() % 4095) + 1), -(int)(rng() % 20) - 4);  // short mantissas: exact adds, ties
        if (rng() & 1) step = nextafter(step, 2 * step);
        const int k = (int)(rng() % 3000);
        double l = step * k;                   // near an iterate of the exact progression
        if (rng() & 1) l = ldexp(1.0, (int)(rng() % 8) - 2);  // a power of two
        if (rng() & 1) l = nextafter(l, (rng() & 1) ? 0.0 : 1e9);
        const double pd0 = (rng() & 3) ? step : step * U(rng) * 3.0;
        bad_random += check(pd0, step, l);
    }
    // (iv) specials
    {
        const double nan = NAN, inf = INFINITY;
        const double v[] = {0.0, -0.0, 1.0, -1.0, 0.05, -0.05, 5e-324, 1e-310, 1e300, inf, -inf, nan, 0x1p-1022, 3.0, 1e-17};
        const int m = sizeof v / sizeof v[0];
        for (int a = 0; a < m; ++a)
            for (int b = 0; b < m; ++b)
                for (int c = 0; c < m; ++c) {
                    // keep the literal loop finite-time: it runs at most MAXI iterations anyway
                    bad_special += check(v[a], v[b], fabs(v[c]));
                }
    }
    // (v) steps far below the limit's ulp range boundaries: long runs, and steps that vanish against pd
    for (long i = 0; i < 2000; ++i) {
        const double step = ldexp(0.5 + 0.5 * U(rng), -10 - (int)(rng() % 8));
        const double l = ldexp(0.5 + 0.5 * U(rng), 1 + (int)(rng() % 4));
        bad_tiny += check(step * U(rng), step, l);
    }
    bad_tiny += check(1.0, 1e-17, 2.0);  // never moves: overflow on both sides
    bad_tiny += check(1.0, 0x1p-53, 1.5);  // exactly half an ulp: tie to even, never moves
    bad_tiny += check(1.0 + 0x1p-52, 0x1p-53, 1.5);  // odd start: first step moves, then stuck
    printf("bad_call_pattern %ld\nbad_steps %ld\nbad_random %ld\nbad_special %ld\nbad_tiny %ld\n", bad_call_pattern, bad_steps,
           bad_random, bad_special, bad_tiny);
    return 0;
}

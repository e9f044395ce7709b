// Host-side check of rs-pathplanning_b200/csrc/geo_predicates.cuh (the header compiles as plain C++):
//   1. pp_quot_in01(a, b) == (0.0 <= a / b && a / b <= 1.0) for the IEEE quotient, on random operands, on operands
//      one or two ulps apart (where the rounding of the quotient decides `<= 1`), on tiny / huge / zero / signed-zero /
//      non-finite operands (where the quotient underflows or overflows);
//   2. pp_ring_hits_segment / pp_point_position against a plain restatement that performs every division, on random
//      rings and segments including points placed exactly on ring vertices and edges.
// Prints the number of mismatches per group; tests/test_math_host.py asserts they are all zero.
//   g++ -O2 -std=c++17 -ffp-contract=off tools/predicate_check.cpp -o /tmp/predicate_check
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../rs-pathplanning_b200/csrc/geo_predicates.cuh"

static uint64_t st = 0x0123456789ABCDEFull;
static uint64_t next64() {  // splitmix64
    uint64_t z = (st += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double uni() { return (double)(next64() >> 11) * 0x1p-53; }
static double from_bits(uint64_t b) {
    double d;
    memcpy(&d, &b, 8);
    return d;
}
static uint64_t to_bits(double d) {
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
}
static volatile double sink_a, sink_b;  // keep the reference division a real division
static bool ref_in01(double a, double b) {
    sink_a = a;
    sink_b = b;
    const double q = sink_a / sink_b;
    return 0.0 <= q && q <= 1.0;
}

// every division performed, as in geo 0.12.2
static bool ref_has_point(const double *rx, const double *ry, uint32_t n, double px, double py) {
    if (n == 0) return false;
    if (n == 1) return rx[0] == px && ry[0] == py;
    for (uint32_t i = 0; i < n; ++i)
        if (rx[i] == px && ry[i] == py) return true;
    for (uint32_t i = 0; i + 1 < n; ++i) {
        const double x0 = rx[i], y0 = ry[i], dx = rx[i + 1] - x0, dy = ry[i + 1] - y0;
        bool hit;
        if (dx == 0.0 && dy == 0.0) {
            hit = (px == x0 && py == y0);
        } else if (dy == 0.0) {
            const double t = (px - x0) / dx;
            hit = (py == y0 && 0.0 <= t && t <= 1.0);
        } else if (dx == 0.0) {
            const double t = (py - y0) / dy;
            hit = (px == x0 && 0.0 <= t && t <= 1.0);
        } else {
            const double tx = (px - x0) / dx, ty = (py - y0) / dy;
            hit = (std::fabs(tx - ty) <= PP_F64_EPSILON && 0.0 <= tx && tx <= 1.0);
        }
        if (hit) return true;
    }
    return false;
}
static bool ref_hits(const double *rx, const double *ry, uint32_t n, double b0x, double b0y, double b1x, double b1y) {
    const double b_dx = b1x - b0x, b_dy = b1y - b0y;
    for (uint32_t i = 0; i + 1 < n; ++i) {
        const double a0x = rx[i], a0y = ry[i], a_dx = rx[i + 1] - a0x, a_dy = ry[i + 1] - a0y;
        const double u_b = b_dy * a_dx - b_dx * a_dy;
        if (u_b == 0.0) continue;
        const double ua_t = b_dx * (a0y - b0y) - b_dy * (a0x - b0x);
        const double ub_t = a_dx * (a0y - b0y) - a_dy * (a0x - b0x);
        const double u_a = ua_t / u_b, u_b2 = ub_t / u_b;
        if (0.0 <= u_a && u_a <= 1.0 && 0.0 <= u_b2 && u_b2 <= 1.0) return true;
    }
    return false;
}

int main(int argc, char **argv) {
    const long n = argc > 1 ? atol(argv[1]) : 2000000;
    long bad_random = 0, bad_close = 0, bad_special = 0, bad_ring = 0;
    // 1a. random magnitudes and signs over the whole exponent range
    for (long i = 0; i < n; ++i) {
        const double a = from_bits(next64()), b = from_bits(next64());
        if (b == 0.0 || std::isnan(b)) continue;
        bad_random += pp_quot_in01(a, b) != ref_in01(a, b);
    }
    // 1a'. independent operands inside the fast path's range (2^-900 .. 2^100), all sign combinations
    for (long i = 0; i < n; ++i) {
        double a = std::ldexp(0.5 + uni(), (int)(uni() * 990) - 895), b = std::ldexp(0.5 + uni(), (int)(uni() * 990) - 895);
        if (next64() & 1) a = -a;
        if (next64() & 1) b = -b;
        bad_random += pp_quot_in01(a, b) != ref_in01(a, b);
    }
    // 1b. world-scale operands a few ulps apart, all sign combinations
    for (long i = 0; i < n; ++i) {
        const double b = std::ldexp(0.5 + uni(), (int)(uni() * 60) - 30);
        const int k = (int)(next64() % 9) - 4;
        double a = from_bits(to_bits(b) + (uint64_t)(int64_t)k);
        if (next64() & 1) a = -a;
        const double bb = (next64() & 1) ? -b : b;
        bad_close += pp_quot_in01(a, bb) != ref_in01(a, bb);
    }
    // 1c. specials: zeros, denormals, the thresholds of the fast path, huge, infinities, NaN
    const double sp[] = {0.0, -0.0, 4.9e-324, -4.9e-324, 2.2e-308, -2.2e-308, 0x1p-900, -0x1p-900, 0x1p-901, 0x1p-899,
                         1e-300, -1e-300, 1.0, -1.0, 0.5, 2.0, 0x1p+100, -0x1p+100, 0x1p+101, 0x1p+99, 1e300, -1e300,
                         1.7e308, INFINITY, -INFINITY, NAN};
    for (double a : sp)
        for (double b : sp) {
            if (b == 0.0 || std::isnan(b)) continue;
            bad_special += pp_quot_in01(a, b) != ref_in01(a, b);
        }
    // 2. rings: regular polygons with jitter, points on vertices / edges / inside / outside, random segments
    std::vector<double> rx(40), ry(40);
    for (long it = 0; it < n / 20; ++it) {
        const uint32_t m = 3 + (uint32_t)(next64() % 20);
        const double cx = 1000.0 * uni(), cy = 1000.0 * uni(), r = 0.5 + 3.0 * uni();
        for (uint32_t k = 0; k < m; ++k) {
            const double th = 6.283185307179586 * k / m;
            rx[k] = cx + r * std::cos(th);
            ry[k] = cy + r * std::sin(th);
        }
        if (next64() % 4 == 0) {  // axis-aligned edges
            rx[1] = rx[0];
            ry[2] = ry[1];
        }
        rx[m] = rx[0];
        ry[m] = ry[0];
        const uint32_t np = m + 1;
        for (int q = 0; q < 8; ++q) {
            double px = cx + (uni() - 0.5) * 3.0 * r, py = cy + (uni() - 0.5) * 3.0 * r;
            const uint32_t e = (uint32_t)(next64() % m);
            if (q == 0) {
                px = rx[e];
                py = ry[e];
            } else if (q == 1) {  // on (or within rounding of) an edge
                const double t = uni();
                px = rx[e] + t * (rx[e + 1] - rx[e]);
                py = ry[e] + t * (ry[e + 1] - ry[e]);
            } else if (q == 2) {
                px = 0.5 * (rx[e] + rx[e + 1]);
                py = 0.5 * (ry[e] + ry[e + 1]);
            }
            const bool want_on = ref_has_point(rx.data(), ry.data(), np, px, py);
            bad_ring += pp_ring_has_point(rx.data(), ry.data(), np, px, py) != want_on;
            const double qx = cx + (uni() - 0.5) * 4.0 * r, qy = cy + (uni() - 0.5) * 4.0 * r;
            bad_ring += pp_ring_hits_segment(rx.data(), ry.data(), np, px, py, qx, qy) !=
                        ref_hits(rx.data(), ry.data(), np, px, py, qx, qy);
            // a segment that ends exactly on a vertex / starts on an edge: u_a or u_b2 equal to 0 or 1 up to rounding
            bad_ring += pp_ring_hits_segment(rx.data(), ry.data(), np, qx, qy, rx[e], ry[e]) !=
                        ref_hits(rx.data(), ry.data(), np, qx, qy, rx[e], ry[e]);
        }
    }
    printf("bad_random %ld\nbad_close %ld\nbad_special %ld\nbad_ring %ld\n", bad_random, bad_close, bad_special, bad_ring);
    return 0;
}

// Host-side accuracy harness for rs-pathplanning_b200/csrc/pp_math.cuh (the header compiles as plain C++).
// Reference: glibc's long double (80-bit) sinl / cosl / atan2l / acosl.  Prints the maximum error in ulps
// of the double result per function; tests/test_math_host.py asserts the thresholds.
//   g++ -O2 -std=c++17 -mfma -ffp-contract=off tools/math_accuracy.cpp -o /tmp/math_accuracy
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../rs-pathplanning_b200/csrc/pp_math.cuh"

static uint64_t st = 0x9E3779B97F4A7C15ull;
static double uni() {  // splitmix64
    uint64_t z = (st += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * 0x1p-53;
}
static double ulps(double got, long double want) {
    if (std::isnan(got) && std::isnan((double)want)) return 0;
    double w = (double)want;
    double u = std::fabs(std::nextafter(w, INFINITY) - w);
    if (w == 0.0) u = 4.9e-324;
    return (double)(fabsl((long double)got - want) / u);
}

int main(int argc, char **argv) {
    const long n = argc > 1 ? atol(argv[1]) : 1000000;
    double e_sin = 0, e_cos = 0, e_at = 0, e_ac = 0, e_sin_big = 0, e_at3 = 0;
    for (long i = 0; i < n; ++i) {
        double x = (uni() * 2 - 1) * (i % 3 == 0 ? 7.0 : (i % 3 == 1 ? 100.0 : 0.8));
        double s, c;
        pp_sincos1(x, &s, &c);
        e_sin = fmax(e_sin, ulps(s, sinl((long double)x)));
        e_cos = fmax(e_cos, ulps(c, cosl((long double)x)));
        double xb = (uni() * 2 - 1) * 9.0e4;
        pp_sincos1(xb, &s, &c);
        e_sin_big = fmax(e_sin_big, fmax(ulps(s, sinl((long double)xb)), ulps(c, cosl((long double)xb))));
        double yy = (uni() * 2 - 1) * (i % 2 ? 3.0 : 1e-3), xx = (uni() * 2 - 1) * (i % 5 ? 60.0 : 1e-2);
        e_at = fmax(e_at, ulps(pp_atan2(yy, xx), atan2l((long double)yy, (long double)xx)));
        double v = uni() * 2 - 1;
        if (i % 4 == 0) v = copysign(1.0 - uni() * 1e-6, v);
        // acos near +1 is ill-conditioned in ulps of the result; measure against max(ulp, 1e-16 absolute)
        long double wa = acosl((long double)v);
        double ga = pp_acos(v);
        double err = (double)fabsl((long double)ga - wa);
        e_ac = fmax(e_ac, fmin(ulps(ga, wa), err / 1.2e-16));
    }
    // 3-way batch must agree bit-for-bit with the single version
    for (long i = 0; i < 10000; ++i) {
        double y[3] = {uni() - 0.5, uni() * 5 - 2, -uni()}, x[3] = {uni() - 0.5, -uni() * 7, uni() * 1e-5}, o[3];
        pp_atan2_n<3>(y, x, o);
        for (int k = 0; k < 3; ++k) e_at3 = fmax(e_at3, o[k] == pp_atan2(y[k], x[k]) ? 0.0 : 1.0);
    }
    // special values
    int bad = 0;
    const double PI = 3.14159265358979323846;
    bad += !(pp_atan2(0.0, 1.0) == 0.0 && !std::signbit(pp_atan2(0.0, 1.0)));
    bad += !(std::signbit(pp_atan2(-0.0, 1.0)) && pp_atan2(-0.0, 1.0) == 0.0);
    bad += !(pp_atan2(0.0, -1.0) == PI && pp_atan2(-0.0, -1.0) == -PI);
    bad += !(pp_atan2(0.0, -0.0) == PI && pp_atan2(0.0, 0.0) == 0.0);
    bad += !(pp_atan2(1.0, 0.0) == PI / 2 && pp_atan2(-1.0, 0.0) == -PI / 2 && pp_atan2(-2.0, 0.0) == atan2(-2.0, 0.0));
    bad += !(pp_atan2(-2.0, 5.0) == -pp_atan2(2.0, 5.0));
    bad += !std::isnan(pp_atan2(NAN, 1.0)) + !std::isnan(pp_atan2(1.0, NAN));
    double s, c;
    pp_sincos1(0.0, &s, &c);
    bad += !(s == 0.0 && c == 1.0);
    pp_sincos1(-0.0, &s, &c);
    bad += !(std::signbit(s) && c == 1.0);
    pp_sincos1(NAN, &s, &c);
    bad += !(std::isnan(s) && std::isnan(c));
    pp_sincos1(1e9, &s, &c);
    bad += !(s == sin(1e9) && c == cos(1e9));
    bad += !(pp_acos(1.0) == 0.0 && pp_acos(-1.0) == PI && std::fabs(pp_acos(0.0) - PI / 2) < 3e-16);
    printf("sin_ulp %.3f\ncos_ulp %.3f\nsincos_big_ulp %.3f\natan2_ulp %.3f\nacos_err %.3f\nbatch_mismatch %.0f\nspecial_bad %d\n",
           e_sin, e_cos, e_sin_big, e_at, e_ac, e_at3, bad);
    return 0;
}

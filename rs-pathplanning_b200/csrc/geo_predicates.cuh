// geo_predicates.cuh -- the geo 0.12.2 predicates behind Space::verify (src/rrt.rs:124-137), written
// for host and device.  Semantics per SURVEY.md Appendix B.1:
//   Polygon::contains(&LineString)   -> every line point strictly inside the exterior ring
//   LineString::intersects(&Polygon) -> a ring segment meets a line segment (parameter test with two
//                                       IEEE divisions, parallel pairs skipped), or a line point is
//                                       inside the polygon
// All arithmetic is plain IEEE f64 in the reference's operation order; every translation unit that
// includes this header is compiled without multiply-add contraction (-fmad=false / -ffp-contract=off),
// which is what makes the flags bit-exact with the oracle.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define PP_HD __host__ __device__ __forceinline__
#else
#define PP_HD inline
#endif

#define PP_F64_EPSILON 2.220446049250313e-16

// 0.0 <= a / b && a / b <= 1.0 for the correctly rounded IEEE quotient (b != 0), without the division.
//   * a == 0: the quotient is +-0 and passes both tests.
//   * equal signs: the quotient is positive, and RN(|a| / |b|) <= 1 iff |a| <= |b| -- if |a| > |b| then
//     |a| >= |b| + ulp(|b|) > |b| (1 + 2^-53), which rounds to 1 + 2^-52 or more.
//   * opposite signs: the quotient is negative and fails, unless it underflows to -0 (which passes `0.0 <= q`).
// Operands outside [2^-900, 2^100] in magnitude (where underflow / overflow of the quotient is conceivable) and
// non-finite ones take the division itself; world-scale coordinates never get there.
PP_HD bool pp_quot_in01(double a, double b) {
    const double aa = fabs(a), ab = fabs(b);
    if (!(aa >= 0x1p-900 && aa <= 0x1p+100 && ab >= 0x1p-900 && ab <= 0x1p+100)) {
        const double q = a / b;
        return 0.0 <= q && q <= 1.0;
    }
    return ((a < 0.0) == (b < 0.0)) && aa <= ab;
}

// Necessary for -eps <= RN(a / b) <= 1 + eps (eps = 2^-52): |a| at most a hair above |b|, and a negative quotient
// only if it is tiny.  A cheap filter in front of the two divisions of the on-boundary test, never the decision.
PP_HD bool pp_quot_near01(double a, double b) {
    const double aa = fabs(a), ab = fabs(b);
    return aa <= 1.0000001 * ab && (((a < 0.0) == (b < 0.0)) || aa <= 1e-15 * ab);
}

// `impl Contains<Point> for LineString`: vertex equality, then the per-segment tx/ty test
PP_HD bool pp_ring_has_point(const double *rx, const double *ry, uint32_t n, double px, double py) {
    if (n == 0) return false;
    if (n == 1) return rx[0] == px && ry[0] == py;
    for (uint32_t i = 0; i < n; ++i)
        if (rx[i] == px && ry[i] == py) return true;
    for (uint32_t i = 0; i + 1 < n; ++i) {
        const double x0 = rx[i], y0 = ry[i];
        const double dx = rx[i + 1] - x0, dy = ry[i + 1] - y0;
        bool hit;
        if (dx == 0.0 && dy == 0.0) {
            hit = (px == x0 && py == y0);
        } else if (dy == 0.0) {
            hit = (py == y0 && pp_quot_in01(px - x0, dx));
        } else if (dx == 0.0) {
            hit = (px == x0 && pp_quot_in01(py - y0, dy));
        } else {
            // tx in [0, 1] is decided without dividing; |tx - ty| <= eps then needs ty in [-eps, 1 + eps]
            // (pp_quot_near01): only points inside the segment's box pay for the two quotients
            const double nx = px - x0, ny = py - y0;
            hit = false;
            if (pp_quot_in01(nx, dx) && pp_quot_near01(ny, dy)) {
                const double tx = nx / dx;
                const double ty = ny / dy;
                hit = (fabs(tx - ty) <= PP_F64_EPSILON);
            }
        }
        if (hit) return true;
    }
    return false;
}

// `get_position`: 0 outside, 1 inside, 2 on boundary
PP_HD int pp_point_position(const double *rx, const double *ry, uint32_t n, double px, double py) {
    if (n == 0) return 0;
    if (pp_ring_has_point(rx, ry, n, px, py)) return 2;
    double xints = 0.0;
    uint32_t crossings = 0;
    for (uint32_t i = 0; i + 1 < n; ++i) {
        const double x0 = rx[i], y0 = ry[i], x1 = rx[i + 1], y1 = ry[i + 1];
        const double ymin = (y0 < y1) ? y0 : y1, ymax = (y0 > y1) ? y0 : y1;
        const double xmax = (x0 > x1) ? x0 : x1;
        if (py > ymin && py <= ymax && px <= xmax) {
            if (y0 != y1) xints = (py - y0) * (x1 - x0) / (y1 - y0) + x0;
            if (x0 == x1 || px <= xints) crossings += 1;
        }
    }
    return (int)(crossings & 1u);
}

// `impl Intersects<LineString> for LineString` with the ring as `self` and one line segment b0->b1
PP_HD bool pp_ring_hits_segment(const double *rx, const double *ry, uint32_t n, double b0x, double b0y, double b1x,
                                double b1y) {
    const double b_dx = b1x - b0x, b_dy = b1y - b0y;
    for (uint32_t i = 0; i + 1 < n; ++i) {
        const double a0x = rx[i], a0y = ry[i];
        const double a_dx = rx[i + 1] - a0x, a_dy = ry[i + 1] - a0y;
        const double u_b = b_dy * a_dx - b_dx * a_dy;
        if (u_b == 0.0) continue;
        const double ua_t = b_dx * (a0y - b0y) - b_dy * (a0x - b0x);
        const double ub_t = a_dx * (a0y - b0y) - a_dy * (a0x - b0x);
        // u_a = ua_t / u_b and u_b2 = ub_t / u_b are only compared with 0 and 1 (pp_quot_in01: same answers)
        if (pp_quot_in01(ua_t, u_b) && pp_quot_in01(ub_t, u_b)) return true;
    }
    return false;
}

// Conservative pad of a ring's AABB.  Outside the padded box neither the crossing-number test nor
// (away from near-collinear rounding noise) the parameter test can fire: the
// rounding error of `xints` and of the quotients is below 2^-48 * max|coordinate|.
PP_HD double pp_aabb_pad(double minx, double miny, double maxx, double maxy) {
    double m = fabs(minx);
    if (fabs(maxx) > m) m = fabs(maxx);
    if (fabs(miny) > m) m = fabs(miny);
    if (fabs(maxy) > m) m = fabs(maxy);
    return m * 0x1p-40 + 0x1p-1000;
}


// ------------------------------------------------------------------------------------------------------------------
// Circle filter (second, exact-by-margin broad phase).  Per ring: a centre c, the square of an OUTER radius that
// every ring point stays inside (inflated by 1e-6 relative + the box pad) and the square of an INNER radius whose
// disc lies inside the polygon (deflated likewise; 0 when the centre is not inside the ring).  For a line segment
// a-b (a == b: a single vertex):
//   class 0  the whole segment stays outside the outer circle: no ring point is within reach, the polygon neither
//            meets the segment nor contains a point of it -> the ring is skipped without the exact predicates;
//   class 1  both end points lie inside the inner circle: the polygon contains them -> Space::verify fails;
//   class 2  undecided -> the exact geo predicates run (and are the answer).
// The margins (1e-6 relative) exceed every rounding error of the tests below and of geo's own formulas by orders of
// magnitude, so classes 0 / 1 agree with what the exact predicates would say -- except, as with the box cull, for
// line segments on the extension of a ring segment and parallel to it to ~2^-45 rad, where geo's parameter test is
// rounding noise (tests/test_oracle_geo.py).  The oracle's culled loop applies the same rule
// with the same arithmetic (the checker restates it as circle_class).
// ------------------------------------------------------------------------------------------------------------------
struct pp_ring_circle {
    double cx, cy, rout2, rin2;
};

PP_HD int pp_circle_class(const pp_ring_circle &c, double ax, double ay, double bx, double by) {
    const double acx = c.cx - ax, acy = c.cy - ay;
    const double a2 = acx * acx + acy * acy;
    const double bcx = c.cx - bx, bcy = c.cy - by;
    const double b2 = bcx * bcx + bcy * bcy;
    if (a2 < c.rin2 && b2 < c.rin2) return 1;
    if (!(a2 > c.rout2) || !(b2 > c.rout2)) return 2;  // an end point within reach of the ring (or NaN)
    // far end points make the cancellation below too coarse for the margin (and may overflow): leave it to the predicates
    if (!(a2 < 1.0e6 * c.rout2) || !(b2 < 1.0e6 * c.rout2)) return 2;
    const double abx = bx - ax, aby = by - ay;
    const double e = acx * abx + acy * aby;  // (c - a) . (b - a)
    const double f = abx * abx + aby * aby;
    if (!(e > 0.0) || !(e < f)) return 0;  // the closest point of the segment is an end point, and both are outside
    // closest point in the interior: |c - a|^2 - e^2 / f > rout2, multiplied through by f > 0
    return (a2 * f - e * e > c.rout2 * f) ? 0 : 2;
}

// host-side set-up of one ring's circle (n points, closed ring); the CPU checker restates the same expressions
inline pp_ring_circle pp_make_ring_circle(const double *rx, const double *ry, uint32_t n, double minx, double miny,
                                          double maxx, double maxy, double pad, bool finite) {
    pp_ring_circle c;
    c.cx = c.cy = 0.0;
    c.rout2 = INFINITY;  // never class 0 / 1: always undecided
    c.rin2 = 0.0;
    if (!finite || n < 3) return c;
    c.cx = 0.5 * (minx + maxx);
    c.cy = 0.5 * (miny + maxy);
    double rout = 0.0;
    for (uint32_t i = 0; i < n; ++i) {
        const double d = hypot(rx[i] - c.cx, ry[i] - c.cy);
        if (d > rout) rout = d;
    }
    const double ro = rout * (1.0 + 1.0e-6) + pad;
    c.rout2 = ro * ro;
    if (pp_point_position(rx, ry, n, c.cx, c.cy) == 1) {
        double rin = INFINITY;
        for (uint32_t i = 0; i + 1 < n; ++i) {
            const double x0 = rx[i], y0 = ry[i], dx = rx[i + 1] - x0, dy = ry[i + 1] - y0;
            const double l2 = dx * dx + dy * dy;
            double t = (l2 > 0.0) ? ((c.cx - x0) * dx + (c.cy - y0) * dy) / l2 : 0.0;
            t = (t > 0.0) ? ((t < 1.0) ? t : 1.0) : 0.0;
            const double d = hypot(c.cx - (x0 + t * dx), c.cy - (y0 + t * dy));
            if (d < rin) rin = d;
        }
        const double ri = rin * (1.0 - 1.0e-6) - pad;
        if (ri > 0.0 && ri < INFINITY) c.rin2 = ri * ri;
    }
    return c;
}

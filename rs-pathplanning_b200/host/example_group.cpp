// example_group.cpp -- a compiled host (what the Rust shim or a C++ planner would be) driving EVERY B200 of the box
// through ONE process: pp_group replicates tree + obstacles by ncclBroadcast, slices each batch contiguously over the
// devices and gathers the results; the answers must be byte-identical to a single context's.
//   ./example_group [queries = 2^18]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pathplanning.hpp"

using namespace pathplanning;

static double u01(uint64_t seed, uint64_t stream, uint64_t i) {  // the counter-based generator of SURVEY 8(d)
    uint64_t z = seed + (stream << 56) + (i + 1) * 0x9E3779B97F4A7C15ull;
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * 0x1p-53;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        int rc_ = (call);                                                                 \
        if (rc_ != PP_OK) {                                                               \
            std::fprintf(stderr, "%s -> %d (%s)\n", #call, rc_, pp_status_string(rc_));   \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

int main(int argc, char **argv) {
    const size_t m = argc > 1 ? (size_t)std::atoll(argv[1]) : (size_t)1 << 18;
    const size_t n_nodes = 1 << 16, n_rings = 600;
    const double world = 300.0;
    const int n_dev = pp_device_count();
    if (n_dev < 1) {
        std::fprintf(stderr, "no sm_100 device (there is no CPU fallback)\n");
        return 2;
    }
    std::vector<double> nx(n_nodes), ny(n_nodes), nyaw(n_nodes), qx(m), qy(m);
    for (size_t i = 0; i < n_nodes; ++i) {
        nx[i] = world * u01(5, 0, i);
        ny[i] = world * u01(5, 1, i);
        nyaw[i] = -M_PI + 2.0 * M_PI * u01(5, 2, i);
    }
    for (size_t i = 0; i < m; ++i) {
        qx[i] = world * u01(4, 0, i);
        qy[i] = world * u01(4, 1, i);
    }
    std::vector<double> ox, oy;
    std::vector<uint32_t> off{0};
    for (size_t r = 0; r < n_rings; ++r) {
        rrt::Polygon c = rrt::create_circle(rrt::Point{world * u01(6, 0, r), world * u01(6, 1, r)}, 1.0 + 2.0 * u01(6, 2, r));
        ox.insert(ox.end(), c.ring.x.begin(), c.ring.x.end());
        oy.insert(oy.end(), c.ring.y.begin(), c.ring.y.end());
        off.push_back((uint32_t)ox.size());
    }
    const double bx[5] = {0, 0, world, world, 0}, by[5] = {0, world, world, 0, 0};

    // reference answers: one context, one device
    pp_ctx *ctx = nullptr;
    CK(pp_ctx_create(0, &ctx));
    CK(pp_tree_upload(ctx, n_nodes, nx.data(), ny.data(), nyaw.data(), nullptr));
    CK(pp_obstacles_upload(ctx, bx, by, 5, ox.data(), oy.data(), off.data(), n_rings));
    std::vector<uint32_t> idx1(m), idx2(m);
    std::vector<double> yaw1(m), yaw2(m);
    std::vector<uint8_t> ok1(m), ok2(m), okd1(m / 8), okd2(m / 8);
    CK(pp_rrt_extend(ctx, m, qx.data(), qy.data(), idx1.data(), yaw1.data(), ok1.data(), 0, 0));
    CK(pp_rrt_extend_dubins(ctx, m / 8, qx.data(), qy.data(), 0.8, 0.1, idx2.data(), yaw2.data(), okd1.data(), 0, 0));
    pp_ctx_destroy(ctx);

    for (int g = 1; g <= n_dev; g *= 2) {
        std::vector<int> devs(g);
        for (int i = 0; i < g; ++i) devs[i] = i;
        pp_group *grp = nullptr;
        CK(pp_group_create(devs.data(), g, &grp));
        CK(pp_group_tree_upload(grp, n_nodes - 512, nx.data(), ny.data(), nyaw.data(), nullptr));
        // the insert site (src/rrt.rs:586-589): the last 512 nodes arrive as an append, i.e. a tail-only broadcast
        CK(pp_group_tree_append(grp, 512, nx.data() + n_nodes - 512, ny.data() + n_nodes - 512, nyaw.data() + n_nodes - 512,
                                nullptr));
        CK(pp_group_obstacles_upload(grp, bx, by, 5, ox.data(), oy.data(), off.data(), n_rings));
        std::fill(idx2.begin(), idx2.end(), 0u);
        CK(pp_group_rrt_extend(grp, m, qx.data(), qy.data(), idx2.data(), yaw2.data(), ok2.data(), 0, 0));
        const bool same = !std::memcmp(idx1.data(), idx2.data(), m * 4) && !std::memcmp(yaw1.data(), yaw2.data(), m * 8) &&
                          !std::memcmp(ok1.data(), ok2.data(), m);
        CK(pp_group_rrt_extend_dubins(grp, m / 8, qx.data(), qy.data(), 0.8, 0.1, idx2.data(), yaw2.data(), okd2.data(), 0, 0));
        const bool same_d = !std::memcmp(okd1.data(), okd2.data(), m / 8);
        size_t free_edges = 0;
        for (size_t i = 0; i < m; ++i) free_edges += ok2[i];
        std::printf("group of %d device(s): extend %s, extend_dubins %s, %zu of %zu straight edges free\n", g,
                    same ? "byte-identical" : "DIFFERS", same_d ? "byte-identical" : "DIFFERS", free_edges, m);
        pp_group_destroy(grp);
        if (!same || !same_d) return 1;
    }
    std::printf("group ok: %d device(s)\n", n_dev);
    return 0;
}

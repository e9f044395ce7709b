// bench_all.cpp -- the entry points of benches/all.rs on the GPU path: RRT::plan_one, RRT::plan_10 and
// Dubins::dubins_path_planning on the reference's bench world / pose (benches/all.rs:6-115), with
// `turn_radius: 1.0` where the stale bench file says `c: 1.0` (benches/all.rs:109 vs src/dubins.rs:322).
// criterion is replaced by a plain timing loop: warm-up W seconds, measure M seconds (defaults 1 / 3;
// the reference's are 5 / 15, benches/all.rs:117-121) -- usage: bench_all [W] [M] [--batch]
#include <chrono>
#include <cstdio>
#include <cstring>

#include "pathplanning.hpp"

using namespace pathplanning;
using Clock = std::chrono::steady_clock;

static std::shared_ptr<rrt::Space> bench_space() {  // benches/all.rs:8-32
    std::vector<rrt::Polygon> obstacle_list = {
        rrt::create_circle({5.0, 5.0}, 1.0), rrt::create_circle({3.0, 6.0}, 2.0), rrt::create_circle({3.0, 8.0}, 2.0),
        rrt::create_circle({3.0, 10.0}, 2.0), rrt::create_circle({7.0, 5.0}, 2.0), rrt::create_circle({9.0, 5.0}, 2.0)};
    rrt::LineString b;
    b.push(-6.0, -6.0);
    b.push(-6.0, 15.0);
    b.push(15.0, 15.0);
    b.push(15.0, -6.0);
    b.push(-6.0, -6.0);
    // geometry as given: no geo-offset inflation in this mirror (see Space::from_inflated)
    return rrt::Space::from_inflated(rrt::Polygon(b), rrt::Robot(1.0, 1.0, 0.8), obstacle_list, 42);
}
static rrt::RRT bench_planner() {  // benches/all.rs:34-42
    const double PI = 3.14159265358979323846;
    return rrt::RRT({-5.0, -5.0}, -45.0 * (PI / 180.0), {6.0, 10.0}, 45.0 * (PI / 180.0), 8000, 0.1, bench_space());
}

template <class F>
static void run(const char *name, double warm, double meas, F &&f) {
    auto t0 = Clock::now();
    size_t it = 0;
    while (std::chrono::duration<double>(Clock::now() - t0).count() < warm) f(), ++it;
    t0 = Clock::now();
    it = 0;
    double el = 0;
    do {
        f();
        ++it;
        el = std::chrono::duration<double>(Clock::now() - t0).count();
    } while (el < meas);
    std::printf("%-34s time: %12.3f us/iter  (%zu iterations)\n", name, el / (double)it * 1e6, it);
}

int main(int argc, char **argv) {
    double warm = argc > 1 ? std::atof(argv[1]) : 1.0, meas = argc > 2 ? std::atof(argv[2]) : 3.0;
    {
        rrt::RRT planner = bench_planner();
        run("RRT::plan_one", warm, meas, [&] { planner.plan_one(); });
        std::printf("  tree grew to %zu nodes (it persists across iterations, as under criterion)\n", planner.tree_size());
    }
    {
        rrt::RRT planner = bench_planner();
        run("RRT::plan_10", warm, meas, [&] {
            for (int k = 0; k < 10; ++k) planner.plan_one();
        });
    }
    {
        const double PI = 3.14159265358979323846;
        dubins::DubinsConfig conf{1.0, 1.0, 45.0 * (PI / 180.0), -3.0, -3.0, -45.0 * (PI / 180.0), 1.0, 0.1};
        run("Dubins::dubins_path_planning", warm, meas, [&] { (void)dubins::dubins_path_planning(conf); });
        auto r = dubins::dubins_path_planning(conf);
        std::printf("  cost %.17g, %zu samples\n", std::get<4>(*r), std::get<0>(*r).size());
    }
    if (argc > 3 && std::strcmp(argv[3], "--batch") == 0) {  // what the GPU path is for: the same pose 2^20 times
        const size_t n = 1 << 20;
        const double PI = 3.14159265358979323846;
        std::vector<double> sx(n, 1.0), sy(n, 1.0), syaw(n, 45.0 * (PI / 180.0)), ex(n, -3.0), ey(n, -3.0),
            eyaw(n, -45.0 * (PI / 180.0));
        run("dubins::batch::eval (2^20 pairs)", warm, meas, [&] { (void)dubins::batch::eval(sx, sy, syaw, ex, ey, eyaw, 1.0); });
    }
    return 0;
}

// example_dubins.cpp -- the entry point of examples/dubins/src/main.rs:7-61 on the GPU path:
// two Dubins paths (there and back) between (1,1,45deg) and (-3,-3,-45deg), turn radius 0.5, step 0.01.
// The gnuplot window of the reference is replaced by a summary on stdout (and the samples as CSV with
// --csv): plotting is out of scope.
#include <cstdio>
#include <cstring>

#include "pathplanning.hpp"

using pathplanning::dubins::dubins_path_planning;
using pathplanning::dubins::DubinsConfig;

int main(int argc, char **argv) {
    const bool csv = argc > 1 && std::strcmp(argv[1], "--csv") == 0;
    const double PI = 3.14159265358979323846;
    const double start_x = 1.0, start_y = 1.0, start_yaw = 45.0 * (PI / 180.0);
    const double end_x = -3.0, end_y = -3.0, end_yaw = -45.0 * (PI / 180.0);
    const double turn_radius = 0.5;
    DubinsConfig conf1{start_x, start_y, start_yaw, end_x, end_y, end_yaw, turn_radius, 0.01};
    DubinsConfig conf2{end_x, end_y, end_yaw, start_x, start_y, start_yaw, turn_radius, 0.01};
    std::printf("Start planner\n");
    int which = 1;
    for (const DubinsConfig &conf : {conf1, conf2}) {
        auto r = dubins_path_planning(conf);
        if (!r) {
            std::printf("Could not generate path\n");
            continue;
        }
        const auto &[px, py, pyaw, mode, cost] = *r;
        static const char *names = "LSR";
        std::printf("conf%d: word=%c%c%c cost=%.17g samples=%zu first=(%.17g,%.17g) last=(%.17g,%.17g)\n", which,
                    names[(int)(*mode)[0]], names[(int)(*mode)[1]], names[(int)(*mode)[2]], cost, px.size(), px.front(),
                    py.front(), px.back(), py.back());
        if (csv)
            for (size_t k = 0; k < px.size(); ++k) std::printf("%d,%.17g,%.17g,%.17g\n", which, px[k], py[k], pyaw[k]);
        ++which;
    }
    return 0;
}

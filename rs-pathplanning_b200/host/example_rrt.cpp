// example_rrt.cpp -- the entry point of examples/rrt/src/main.rs:22-96 on the GPU path: load a JSON world
// {"bounds": [[x,y],...], "obstacles": [[[x,y],...],...], "path": [], "start": [x,y,yaw], "goal": [x,y,yaw]}
// (the format of examples/rrt/transit.debug.json, main.rs:13-20), build Space / RRT with Robot(1.8, 3.0, 0.8),
// 8000 iterations, step 0.1 (main.rs:44, 55-63) and run plan().  serde_json is replaced by a 60-line
// number scanner; plotting is out of scope.  Usage: example_rrt world.json [max_iter] [seed]
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "pathplanning.hpp"

using namespace pathplanning::rrt;

struct Json {  // just enough for nested arrays of numbers keyed by name
    std::string s;
    size_t find_key(const std::string &k) const {
        size_t p = s.find("\"" + k + "\"");
        if (p == std::string::npos) throw std::runtime_error("missing key " + k);
        return s.find(':', p) + 1;
    }
    // flat list of numbers inside the bracket expression starting at p, with the nesting depth of each
    void numbers(size_t p, std::vector<double> &vals, std::vector<int> &depth) const {
        int d = 0;
        for (; p < s.size(); ++p) {
            char c = s[p];
            if (c == '[') ++d;
            else if (c == ']') {
                if (--d == 0) return;
            } else if (c == '-' || std::isdigit((unsigned char)c)) {
                char *end;
                vals.push_back(std::strtod(s.c_str() + p, &end));
                depth.push_back(d);
                p = (size_t)(end - s.c_str()) - 1;
            }
        }
    }
};

int main(int argc, char **argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s world.json [max_iter] [seed] [round_batch]\n", argv[0]);
        return 2;
    }
    std::ifstream f(argv[1]);
    if (!f) {
        std::fprintf(stderr, "file should open\n");
        return 2;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    Json j{ss.str()};
    const size_t max_iter = argc > 2 ? (size_t)std::atol(argv[2]) : 8000;
    const uint64_t seed = argc > 3 ? (uint64_t)std::atoll(argv[3]) : 1;
    const size_t round_batch = argc > 4 ? (size_t)std::atol(argv[4]) : 0;  // 0: the reference's plan() loop

    std::vector<double> v;
    std::vector<int> d;
    j.numbers(j.find_key("bounds"), v, d);
    LineString b;
    for (size_t i = 0; i + 1 < v.size(); i += 2) b.push(v[i], v[i + 1]);
    v.clear();
    d.clear();
    // obstacles: a new ring starts whenever the scanner re-enters depth 3 after leaving it
    std::vector<Polygon> obstacle_list;
    {
        const std::string &s = j.s;
        size_t p = j.find_key("obstacles");
        int depth = 0;
        LineString cur;
        std::vector<double> pt;
        for (; p < s.size(); ++p) {
            char c = s[p];
            if (c == '[') ++depth;
            else if (c == ']') {
                --depth;
                if (depth == 2 && pt.size() == 2) {
                    cur.push(pt[0], pt[1]);
                    pt.clear();
                }
                if (depth == 1 && cur.size()) {
                    obstacle_list.emplace_back(std::move(cur));
                    cur = LineString();
                }
                if (depth == 0) break;
            } else if (c == '-' || std::isdigit((unsigned char)c)) {
                char *end;
                pt.push_back(std::strtod(s.c_str() + p, &end));
                p = (size_t)(end - s.c_str()) - 1;
            }
        }
    }
    std::vector<double> st, gl;
    std::vector<int> dd;
    j.numbers(j.find_key("start"), st, dd);
    j.numbers(j.find_key("goal"), gl, dd);

    Polygon bounds(std::move(b));
    Robot robot(1.8, 3.0, 0.8);
    // the world file is taken as already carrying the robot's safety margin (Space::new's geo-offset inflation,
    // src/rrt.rs:81-111, is not available in this mirror: Space's plain constructor refuses a non-zero width)
    auto space = Space::from_inflated(bounds, robot, obstacle_list, seed);
    RRT planner(Coordinate{st[0], st[1]}, st[2], Coordinate{gl[0], gl[1]}, gl[2], max_iter, 0.1, space);
    std::printf("Start planner (bounds %zu pts, %zu obstacles, %zu iterations)\n", bounds.ring.size(),
                obstacle_list.size(), max_iter);
    auto path = round_batch ? planner.plan_rounds(round_batch) : planner.plan();
    if (round_batch) {
        // self-check of the round-level goal test: the same lines as check_finish node by node
        std::vector<pathplanning::rrt::NodePtr> some(planner.nodes().begin() + 1,
                                                    planner.nodes().begin() + std::min<size_t>(planner.nodes().size(), 40));
        auto many = planner.check_finish_many(some);
        size_t same = 0;
        for (size_t k = 0; k < some.size(); ++k) {
            auto one = planner.check_finish(some[k]);
            same += (bool(one) == bool(many[k])) && (!one || (one->x == many[k]->x && one->y == many[k]->y));
        }
        std::printf("check_finish_many agrees on %zu of %zu nodes\n", same, some.size());
    }
    if (path) {
        std::printf("Path generated!\nNum points: %zu\nlength: %.6f\ntree nodes: %zu\n", path->size(),
                    path->euclidean_length(), planner.tree_size());
        std::printf("verify(path) = %d\n", (int)space->verify(*path));
    } else {
        std::printf("Unable to generate path\ntree nodes: %zu\n", planner.tree_size());
    }
    return 0;
}

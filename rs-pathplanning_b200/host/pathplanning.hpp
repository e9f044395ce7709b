// pathplanning.hpp -- C++ host mirror of the Rust crate `pathplanning` v0.1.2 over the C-ABI of
// libpathplanning_b200.so.  No Rust toolchain exists in this image, so the host side above the C-ABI is
// written in C++ (the reference is compiled code); rs-pathplanning_b200/rust/ holds the same surface as a
// source-only Rust shim.  Names, argument order and return conventions follow the crate:
//
//   pathplanning::dubins   src/dubins.rs : Mode, mod2pi, pi_2_pi, lsl..lrl, DubinsConfig,
//                                          dubins_path_planning_from_origin, dubins_path_planning
//   pathplanning::rrt      src/rrt.rs    : Robot, create_circle, Space, Node, NodeIter, line_to_origin, RRT
//
// Option<T> -> std::optional<T>, Arc<Node> -> std::shared_ptr<Node>, geo's Point / LineString / Polygon ->
// the minimal structs below (Polygon's constructor closes the ring like geo-types 0.4's Polygon::new).
// Every numeric call runs on the GPU; there is no CPU fallback (a missing B200 throws pathplanning::Error).
// Out of scope: geo-offset inflation in Space::new (bounds / obstacles are
// taken as given) and thread_rng (a seedable std::mt19937_64 replaces it).
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <mutex>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/pathplanning_b200.h"

namespace pathplanning {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string &what) : std::runtime_error(what), status(s) {}
};

namespace detail {
inline pp_ctx *&ctx_slot() {
    static pp_ctx *c = nullptr;
    return c;
}
// crate-level lazy context (device from PP_DEVICE, default 0)
inline pp_ctx *ctx() {
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        if (const char *e = std::getenv("PP_DEVICE")) dev = std::atoi(e);
        int rc = pp_ctx_create(dev, &ctx_slot());
        if (rc != PP_OK) throw Error(rc, std::string("pp_ctx_create: ") + pp_status_string(rc));
    });
    return ctx_slot();
}
inline void check(int rc, const char *what) {
    if (rc != PP_OK)
        throw Error(rc, std::string(what) + ": " + pp_status_string(rc) + " (" + pp_last_error(ctx_slot()) + ")");
}
}  // namespace detail

// ================================================================================================ dubins
namespace dubins {

enum class Mode { L, S, R };  // src/dubins.rs:4-9
using ModeArray = std::array<Mode, 3>;
using ModeSlice = std::optional<ModeArray>;  // Option<&'static [Mode; 3]>
using PlannerResult = std::tuple<std::optional<double>, std::optional<double>, std::optional<double>, ModeSlice>;

// src/dubins.rs:26,50,73,94,115,135
inline constexpr ModeArray LSL_MODE{Mode::L, Mode::S, Mode::L}, RSR_MODE{Mode::R, Mode::S, Mode::R},
    LSR_MODE{Mode::L, Mode::S, Mode::R}, RSL_MODE{Mode::R, Mode::S, Mode::L}, RLR_MODE{Mode::R, Mode::L, Mode::R},
    LRL_MODE{Mode::L, Mode::R, Mode::L};
inline const ModeArray &word_modes(int w) {
    static const ModeArray all[6] = {LSL_MODE, RSR_MODE, LSR_MODE, RSL_MODE, RLR_MODE, LRL_MODE};
    return all[w];
}

inline double mod2pi(double theta) {  // src/dubins.rs:18
    double out;
    detail::check(pp_mod2pi(detail::ctx(), 1, &theta, &out, 0), "mod2pi");
    return out;
}
inline double pi_2_pi(double angle) {  // src/dubins.rs:22
    double out;
    detail::check(pp_mod2pi(detail::ctx(), 1, &angle, &out, 1), "pi_2_pi");
    return out;
}

namespace detail_words {
inline PlannerResult word(int w, double alpha, double beta, double d) {
    double tpq[18];
    uint8_t feas[6];
    detail::check(pp_dubins_words(detail::ctx(), 1, &alpha, &beta, &d, tpq, feas), "dubins_words");
    if (!feas[w]) return {std::nullopt, std::nullopt, std::nullopt, word_modes(w)};
    return {tpq[3 * w], tpq[3 * w + 1], tpq[3 * w + 2], word_modes(w)};
}
}  // namespace detail_words
inline PlannerResult lsl(double a, double b, double d) { return detail_words::word(PP_LSL, a, b, d); }  // :27
inline PlannerResult rsr(double a, double b, double d) { return detail_words::word(PP_RSR, a, b, d); }  // :51
inline PlannerResult lsr(double a, double b, double d) { return detail_words::word(PP_LSR, a, b, d); }  // :74
inline PlannerResult rsl(double a, double b, double d) { return detail_words::word(PP_RSL, a, b, d); }  // :95
inline PlannerResult rlr(double a, double b, double d) { return detail_words::word(PP_RLR, a, b, d); }  // :116
inline PlannerResult lrl(double a, double b, double d) { return detail_words::word(PP_LRL, a, b, d); }  // :136

struct DubinsConfig {  // src/dubins.rs:315-324
    double sx, sy, syaw, ex, ey, eyaw, turn_radius, step_size;
};

using DubinsPath = std::tuple<std::vector<double>, std::vector<double>, std::vector<double>, ModeSlice, double>;
using DubinsPathResult = std::optional<DubinsPath>;

namespace detail_path {
inline DubinsPathResult path(double sx, double sy, double syaw, double ex, double ey, double eyaw, double radius,
                             double step, int from_origin) {
    size_t cap = 4096, n = 0;
    for (;;) {
        std::vector<double> px(cap), py(cap), pyaw(cap);
        int word = PP_WORD_NONE;
        double cost = 0.0;
        int rc = pp_dubins_path(detail::ctx(), sx, sy, syaw, ex, ey, eyaw, radius, step, from_origin, px.data(),
                                py.data(), pyaw.data(), cap, &n, &word, &cost);
        if (rc == PP_ERR_OVERFLOW && n > cap) {
            cap = n;
            continue;
        }
        detail::check(rc, "dubins_path");
        if (word == PP_WORD_NONE) return std::nullopt;  // src/dubins.rs:397
        px.resize(n);
        py.resize(n);
        pyaw.resize(n);
        return DubinsPath{std::move(px), std::move(py), std::move(pyaw), word_modes(word), cost};
    }
}
}  // namespace detail_path

// src/dubins.rs:326 ; c = 1 / turn radius
inline DubinsPathResult dubins_path_planning_from_origin(double dx, double dy, double eyaw, double c,
                                                         double step_size) {
    return detail_path::path(0, 0, 0, dx, dy, eyaw, 1.0 / c, step_size, 1);
}
// src/dubins.rs:401
inline DubinsPathResult dubins_path_planning(const DubinsConfig &conf) {
    return detail_path::path(conf.sx, conf.sy, conf.syaw, conf.ex, conf.ey, conf.eyaw, conf.turn_radius,
                             conf.step_size, 0);
}

// ---- batched entry points: what the GPU path is for
namespace batch {
struct Eval {
    std::vector<double> cost;    // radius-normalised, +inf when no word is feasible
    std::vector<uint8_t> word;   // pp_word or PP_WORD_NONE
    std::vector<double> tpq;     // 3 per pair (empty unless requested)
};
inline Eval eval(const std::vector<double> &sx, const std::vector<double> &sy, const std::vector<double> &syaw,
                 const std::vector<double> &ex, const std::vector<double> &ey, const std::vector<double> &eyaw,
                 double turn_radius, bool want_tpq = false) {
    const size_t n = sx.size();
    Eval r;
    r.cost.resize(n);
    r.word.resize(n);
    if (want_tpq) r.tpq.resize(3 * n);
    detail::check(pp_dubins_eval(detail::ctx(), n, sx.data(), sy.data(), syaw.data(), ex.data(), ey.data(),
                                 eyaw.data(), nullptr, turn_radius, r.cost.data(), r.word.data(),
                                 want_tpq ? r.tpq.data() : nullptr),
                  "dubins_eval");
    return r;
}
}  // namespace batch
}  // namespace dubins

// ================================================================================================ rrt
namespace rrt {

struct Point {
    double x, y;
    std::pair<double, double> x_y() const { return {x, y}; }
};
using Coordinate = Point;
struct LineString {
    std::vector<double> x, y;
    size_t size() const { return x.size(); }
    void push(double px, double py) {
        x.push_back(px);
        y.push_back(py);
    }
    double euclidean_length() const {  // geo EuclideanLength: sum of hypot per segment
        double s = 0.0;
        for (size_t i = 0; i + 1 < x.size(); ++i) s += std::hypot(x[i + 1] - x[i], y[i + 1] - y[i]);
        return s;
    }
};
struct Polygon {
    LineString ring;  // exterior, closed (no interior rings: the reference never builds any)
    explicit Polygon(LineString exterior) : ring(std::move(exterior)) {
        if (!ring.x.empty() && (ring.x.front() != ring.x.back() || ring.y.front() != ring.y.back()))
            ring.push(ring.x.front(), ring.y.front());  // geo-types 0.4 Polygon::new closes rings
    }
    const LineString &exterior() const { return ring; }
};

class Robot {  // src/rrt.rs:17-40
   public:
    Robot(double width, double height, double max_steer) : width_(width), height_(height), max_steer_(max_steer) {}
    double get_width() const { return width_; }
    double get_steer() const { return max_steer_; }

   private:
    double width_, height_, max_steer_;
};

// src/rrt.rs:43-60 (std::cos / std::sin are glibc's, as Rust's f64::cos / sin)
inline Polygon create_circle(Point center, double radius) {
    const double PI = 3.14159265358979323846;
    const double circum = 2.0 * PI * radius;
    const double n = std::ceil(circum / 1.0);
    const size_t cnt = (size_t)(n + 1.0);
    LineString ls;
    for (size_t k = 0; k < cnt; ++k) {
        const double x = (double)k;
        ls.push((std::cos(2.0 * PI / n * x) * radius) + center.x, (std::sin(2.0 * PI / n * x) * radius) + center.y);
    }
    return Polygon(std::move(ls));
}

class Space {  // src/rrt.rs:70-159
   public:
    // Space::new (src/rrt.rs:81-122) shrinks the bounds and inflates every obstacle by width / 2 with the
    // geo-offset crate, whose source is not available here: this constructor therefore
    // REFUSES a robot of non-zero width instead of silently dropping the safety margin.  Geometry that already
    // carries the margin goes through Space::from_inflated.
    Space(Polygon bounds, Robot robot, std::vector<Polygon> obstacle_list, uint64_t seed = 0x5EED)
        : Space(std::move(bounds), robot, std::move(obstacle_list), seed, false) {}
    static std::shared_ptr<Space> from_inflated(Polygon bounds, Robot robot, std::vector<Polygon> obstacle_list,
                                                uint64_t seed = 0x5EED) {
        return std::shared_ptr<Space>(new Space(std::move(bounds), robot, std::move(obstacle_list), seed, true));
    }
    ~Space() { pp_ctx_destroy(ctx_); }
    Space(const Space &) = delete;
    Space &operator=(const Space &) = delete;
    // every Space owns its GPU context (one world + one tree, as an RRT owns its Space and its RTree in the
    // reference, src/rrt.rs:325-356): any number of planners can live side by side
    pp_ctx *ctx() const { return ctx_; }

   private:
    Space(Polygon bounds, Robot robot, std::vector<Polygon> obstacle_list, uint64_t seed, bool inflated)
        : bounds_(std::move(bounds)), robot_(robot), obstacles_(std::move(obstacle_list)), rng_(seed) {
        if (!inflated && robot_.get_width() != 0.0)
            throw Error(PP_ERR_INVALID,
                        "Space::new: geo-offset inflation by Robot.width / 2 is not available in this mirror; pass "
                        "pre-inflated bounds / obstacles through Space::from_inflated");
        int dev = 0;
        if (const char *e = std::getenv("PP_DEVICE")) dev = std::atoi(e);
        int rc = pp_ctx_create(dev, &ctx_);
        if (rc != PP_OK) throw Error(rc, std::string("pp_ctx_create: ") + pp_status_string(rc));
        const LineString &b = bounds_.ring;
        minx_ = maxx_ = b.x.empty() ? 0.0 : b.x[0];
        miny_ = maxy_ = b.y.empty() ? 0.0 : b.y[0];
        for (size_t i = 0; i < b.size(); ++i) {  // src/rrt.rs:84-106
            minx_ = std::min(minx_, b.x[i]);
            maxx_ = std::max(maxx_, b.x[i]);
            miny_ = std::min(miny_, b.y[i]);
            maxy_ = std::max(maxy_, b.y[i]);
        }
        std::vector<double> ox, oy;
        std::vector<uint32_t> off{0};
        for (const Polygon &p : obstacles_) {
            ox.insert(ox.end(), p.ring.x.begin(), p.ring.x.end());
            oy.insert(oy.end(), p.ring.y.begin(), p.ring.y.end());
            off.push_back((uint32_t)ox.size());
        }
        detail::check(pp_obstacles_upload(ctx_, b.x.data(), b.y.data(), b.size(), ox.data(), oy.data(),
                                          off.data(), obstacles_.size()),
                      "obstacles_upload");
    }

   public:
    bool verify(const LineString &line) const {  // src/rrt.rs:124-137
        uint32_t off[2] = {0, (uint32_t)line.size()};
        uint8_t ok = 0;
        detail::check(pp_verify_polylines(ctx_, 1, line.x.data(), line.y.data(), off, &ok, 0), "verify");
        return ok != 0;
    }
    std::vector<uint8_t> verify_many(const std::vector<LineString> &lines) const {  // one launch for all lines
        std::vector<uint32_t> off{0};
        std::vector<double> px, py;
        for (const LineString &l : lines) {
            px.insert(px.end(), l.x.begin(), l.x.end());
            py.insert(py.end(), l.y.begin(), l.y.end());
            off.push_back((uint32_t)px.size());
        }
        std::vector<uint8_t> ok(lines.size());
        if (!lines.empty())
            detail::check(pp_verify_polylines(ctx_, lines.size(), px.data(), py.data(), off.data(), ok.data(), 0),
                          "verify_many");
        return ok;
    }
    Point rand_point() {  // src/rrt.rs:139-146
        std::uniform_real_distribution<double> ux(minx_, maxx_), uy(miny_, maxy_);
        std::lock_guard<std::mutex> lk(mu_);
        return Point{ux(rng_), uy(rng_)};
    }
    double get_steer() const { return robot_.get_steer(); }
    std::vector<Polygon> get_obs() const { return obstacles_; }
    Polygon get_bounds() const { return bounds_; }

   private:
    Polygon bounds_;
    Robot robot_;
    std::vector<Polygon> obstacles_;
    double minx_, maxx_, miny_, maxy_;
    std::mt19937_64 rng_;
    std::mutex mu_;
    pp_ctx *ctx_ = nullptr;
};

inline double compute_yaw(const Point &from, const Point &to) {  // src/rrt.rs:267-271
    return std::atan2(to.y - from.y, to.x - from.x);
}

class Node;
using NodePtr = std::shared_ptr<Node>;
class NodeIter;

class Node {  // src/rrt.rs:161-214
   public:
    Node(Point point, NodePtr parent) : point_(point), parent_(std::move(parent)) {
        yaw_ = compute_yaw(point_, parent_->get_point());
    }
    static Node new_root(Point point, double yaw) { return Node(point, nullptr, yaw); }
    static Node new_goal(Point point, NodePtr parent, double yaw) { return Node(point, std::move(parent), yaw); }
    NodePtr get_parent() const { return parent_; }
    inline NodeIter get_above() const;
    const Point &get_point() const { return point_; }
    Coordinate get_coord() const { return point_; }
    double get_yaw() const { return yaw_; }
    int64_t slot = -1;  // index in the GPU tree mirror (set on insertion)

   private:
    Node(Point p, NodePtr parent, double yaw) : point_(p), parent_(std::move(parent)), yaw_(yaw) {}
    Point point_;
    NodePtr parent_;
    double yaw_;
};

class NodeIter {  // src/rrt.rs:248-265
   public:
    explicit NodeIter(NodePtr curr) : curr_(std::move(curr)) {}
    NodePtr next() {
        NodePtr c = curr_;
        if (c) curr_ = c->get_parent();
        return c;
    }

   private:
    NodePtr curr_;
};
inline NodeIter Node::get_above() const { return NodeIter(parent_); }

namespace detail_rrt {
struct Edges {
    std::vector<double> sx, sy, syaw, ex, ey, eyaw;
};
inline Edges chain_edges(const NodePtr &node) {
    Edges e;
    NodeIter it(node);
    while (NodePtr n = it.next()) {
        NodePtr p = n->get_parent();
        if (!p) break;
        e.sx.push_back(n->get_point().x);
        e.sy.push_back(n->get_point().y);
        e.syaw.push_back(n->get_yaw());
        e.ex.push_back(p->get_point().x);
        e.ey.push_back(p->get_point().y);
        e.eyaw.push_back(p->get_yaw());
    }
    return e;
}
}  // namespace detail_rrt

// src/rrt.rs:291-321 : per-edge Dubins samples in node->root order (one batched count + fill), then the
// root's own point.  An edge without a feasible word contributes its start point (:313); finalize's copy of the
// loop panics instead (:529), which `strict` reproduces as a std::runtime_error.
namespace detail_rrt {
constexpr const char *SHOULD_PLAN = "Should plan dubins curve";  // src/rrt.rs:529
inline LineString chain_line(const NodePtr &node, double turn_radius, double step_size, bool strict);
}  // namespace detail_rrt
inline LineString line_to_origin(const NodePtr &node, double turn_radius, double step_size) {
    return detail_rrt::chain_line(node, turn_radius, step_size, false);
}
inline LineString detail_rrt::chain_line(const NodePtr &node, double turn_radius, double step_size, bool strict) {
    detail_rrt::Edges e = detail_rrt::chain_edges(node);
    const size_t m = e.sx.size();
    LineString out;
    if (m) {
        std::vector<uint32_t> counts(m);
        std::vector<unsigned char> plan(m * PP_DUBINS_PLAN_BYTES);
        detail::check(pp_dubins_sample_count(detail::ctx(), m, e.sx.data(), e.sy.data(), e.syaw.data(), e.ex.data(),
                                             e.ey.data(), e.eyaw.data(), turn_radius, step_size, 0, counts.data(),
                                             plan.data()),
                      "sample_count");
        std::vector<uint64_t> offsets(m);
        uint64_t total = 0;
        for (size_t i = 0; i < m; ++i) {
            offsets[i] = total;
            total += (counts[i] == 0xFFFFFFFFu) ? 0 : counts[i];
        }
        std::vector<double> xyyaw(3 * total);
        detail::check(pp_dubins_sample_fill(detail::ctx(), m, plan.data(), offsets.data(), total, xyyaw.data()),
                      "sample_fill");
        for (size_t i = 0; i < m; ++i) {
            const uint8_t word = plan[i * PP_DUBINS_PLAN_BYTES + 104];
            if (word == PP_WORD_NONE) {  // src/rrt.rs:313
                if (strict) throw std::runtime_error(SHOULD_PLAN);
                out.push(e.sx[i], e.sy[i]);
                continue;
            }
            for (uint64_t k = offsets[i]; k < offsets[i] + counts[i]; ++k) out.push(xyyaw[3 * k], xyyaw[3 * k + 1]);
        }
    }
    NodePtr root = node;
    while (root->get_parent()) root = root->get_parent();
    out.push(root->get_point().x, root->get_point().y);  // src/rrt.rs:316
    return out;
}

constexpr size_t RECURSION_LIMIT = 16;  // src/rrt.rs:14

class RRT {  // src/rrt.rs:325-619
   public:
    RRT(Coordinate start, double start_yaw, Coordinate goal, double goal_yaw, size_t max_iter, double step_size,
        std::shared_ptr<Space> space)
        : goal_(goal), goal_yaw_(goal_yaw), max_iter_(max_iter), step_size_(step_size), space_(std::move(space)) {
        NodePtr root = std::make_shared<Node>(Node::new_root(start, start_yaw));
        root->slot = 0;
        nodes_.push_back(root);
        const int32_t par = -1;
        detail::check(pp_tree_upload(space_->ctx(), 1, &start.x, &start.y, &start_yaw, &par), "tree_upload");
    }

    std::optional<NodePtr> get_nearest_node(const Point &point) const {  // src/rrt.rs:378-391
        uint32_t idx = 0xFFFFFFFFu;
        detail::check(pp_nn(space_->ctx(), 1, &point.x, &point.y, &idx, nullptr, PP_NN_DEFAULT), "nn");
        if (idx == 0xFFFFFFFFu) return std::nullopt;
        std::lock_guard<std::mutex> lk(mu_);
        return nodes_[idx];
    }
    std::optional<NodePtr> get_random_node() const {  // src/rrt.rs:406-412
        Point p = space_->rand_point();
        auto nearest = get_nearest_node(p);
        if (!nearest) return std::nullopt;
        return std::make_shared<Node>(p, *nearest);
    }
    // src/rrt.rs:414-426 : Space::verify of line_to_origin(node) = AND over the chain's edges, each verified
    // as samples ++ [parent point] by the fused kernel (nothing is materialised)
    bool verify_node(const NodePtr &node) const {
        detail_rrt::Edges e = detail_rrt::chain_edges(node);
        const size_t m = e.sx.size();
        if (m == 0) {
            LineString one;
            one.push(node->get_point().x, node->get_point().y);
            return space_->verify(one);
        }
        std::vector<uint8_t> ok(m);
        detail::check(pp_collide_dubins(space_->ctx(), m, e.sx.data(), e.sy.data(), e.syaw.data(), e.ex.data(),
                                        e.ey.data(), e.eyaw.data(), space_->get_steer(), step_size_, ok.data(), 0),
                      "collide_dubins");
        for (uint8_t v : ok)
            if (!v) return false;
        return true;
    }
    std::optional<LineString> check_finish(const NodePtr &node) const {  // src/rrt.rs:428-438
        NodePtr goal_node = std::make_shared<Node>(Node::new_goal(goal_, node, goal_yaw_));
        LineString line = finalize(goal_node);
        if (space_->verify(line)) return line;
        return std::nullopt;
    }
    std::optional<NodePtr> optimize(const NodePtr &node, size_t i) const {  // src/rrt.rs:463-487
        // Same candidates, same (root-first) order and same verdicts as the reference's loop, but all shortcut
        // candidates of a level are verified in ONE fused launch: verify(line_to_origin(new)) =
        // verify(edge new -> to_node) AND verify(chain of to_node), the latter a suffix-AND over the chain's edges.
        if (i >= RECURSION_LIMIT) return std::nullopt;
        std::vector<NodePtr> chain;  // node, parent, ..., root
        NodeIter it(node);
        while (NodePtr n = it.next()) chain.push_back(n);
        const size_t n = chain.size();
        std::vector<NodePtr> cands(n);
        detail_rrt::Edges e;
        auto push = [&e](const Node &a, const Node &b) {
            e.sx.push_back(a.get_point().x);
            e.sy.push_back(a.get_point().y);
            e.syaw.push_back(a.get_yaw());
            e.ex.push_back(b.get_point().x);
            e.ey.push_back(b.get_point().y);
            e.eyaw.push_back(b.get_yaw());
        };
        for (size_t k = 0; k < n; ++k) {
            cands[k] = std::make_shared<Node>(node->get_coord(), chain[k]);
            push(*cands[k], *chain[k]);
        }
        for (size_t k = 0; k + 1 < n; ++k) push(*chain[k], *chain[k + 1]);
        std::vector<uint8_t> ok(e.sx.size());
        detail::check(pp_collide_dubins(space_->ctx(), ok.size(), e.sx.data(), e.sy.data(), e.syaw.data(), e.ex.data(),
                                        e.ey.data(), e.eyaw.data(), space_->get_steer(), step_size_, ok.data(), 0),
                      "collide_dubins");
        std::vector<uint8_t> chain_ok(n, 1);
        for (size_t k = n - 1; k-- > 0;) chain_ok[k] = (uint8_t)(chain_ok[k + 1] && ok[n + k]);
        for (size_t k = n; k-- > 0;) {  // .rev(): root first
            if (ok[k] && chain_ok[k]) {
                auto deeper = optimize(chain[k], i + 1);
                if (deeper) return std::make_shared<Node>(node->get_coord(), *deeper);
                return cands[k];
            }
        }
        return std::nullopt;
    }
    NodePtr optimize_from_goal(const NodePtr &goal_node) const {  // src/rrt.rs:489-501
        NodePtr parent = goal_node->get_parent();
        if (!parent) return goal_node;
        auto n = optimize(parent, 0);
        if (!n) return goal_node;
        return std::make_shared<Node>(Node::new_goal(goal_node->get_coord(), *n, goal_yaw_));
    }
    LineString finalize(const NodePtr &goal_node) const {  // src/rrt.rs:503-540
        NodePtr top = optimize_from_goal(goal_node);
        LineString l = detail_rrt::chain_line(top, space_->get_steer(), step_size_, true);  // :529 panics on None
        l.x.pop_back();  // the root contributes nothing here (None => vec![], :532)
        l.y.pop_back();
        LineString r;
        for (size_t k = l.size(); k-- > 0;) r.push(l.x[k], l.y[k]);  // :538 reverse
        return r;
    }
    std::optional<LineString> plan_one() {  // src/rrt.rs:583-597
        auto rnd = get_random_node();
        if (rnd && verify_node(*rnd)) {
            insert(*rnd);
            return check_finish(*rnd);
        }
        return std::nullopt;
    }
    // src/rrt.rs:599-619 : the reference's 4 racy rayon workers become sequential iterations
    std::optional<LineString> plan() {
        std::optional<LineString> best;
        double best_len = std::numeric_limits<double>::infinity();
        for (size_t it = 0; it < max_iter_; ++it) {
            auto r = plan_one();
            if (r) {
                double len = r->euclidean_length();
                if (len < best_len) {
                    best_len = len;
                    best = std::move(r);
                }
            }
        }
        return best;
    }
    // ---- SURVEY 8f-2/3 at round level.  optimize() for MANY start nodes: the candidates of all nodes of a recursion
    // level go into one fused launch (per node: the candidates, order and verdicts of optimize()).
    std::vector<std::optional<NodePtr>> optimize_many(const std::vector<NodePtr> &starts, size_t i = 0) const {
        std::vector<std::optional<NodePtr>> out(starts.size());
        if (i >= RECURSION_LIMIT || starts.empty()) return out;
        std::vector<std::vector<NodePtr>> chains(starts.size()), cands(starts.size());
        std::vector<size_t> first(starts.size());
        detail_rrt::Edges e;
        auto push = [&e](const Node &a, const Node &b) {
            e.sx.push_back(a.get_point().x);
            e.sy.push_back(a.get_point().y);
            e.syaw.push_back(a.get_yaw());
            e.ex.push_back(b.get_point().x);
            e.ey.push_back(b.get_point().y);
            e.eyaw.push_back(b.get_yaw());
        };
        for (size_t j = 0; j < starts.size(); ++j) {
            NodeIter it(starts[j]);
            while (NodePtr n = it.next()) chains[j].push_back(n);
            first[j] = e.sx.size();
            for (const NodePtr &to : chains[j]) {
                cands[j].push_back(std::make_shared<Node>(starts[j]->get_coord(), to));
                push(*cands[j].back(), *to);
            }
            for (size_t k = 0; k + 1 < chains[j].size(); ++k) push(*chains[j][k], *chains[j][k + 1]);
        }
        std::vector<uint8_t> ok(e.sx.size());
        detail::check(pp_collide_dubins(space_->ctx(), ok.size(), e.sx.data(), e.sy.data(), e.syaw.data(), e.ex.data(),
                                        e.ey.data(), e.eyaw.data(), space_->get_steer(), step_size_, ok.data(), 0),
                      "collide_dubins");
        std::vector<size_t> live;
        std::vector<size_t> pick(starts.size());
        std::vector<NodePtr> next;
        for (size_t j = 0; j < starts.size(); ++j) {
            const size_t n = chains[j].size(), a = first[j];
            std::vector<uint8_t> chain_ok(n, 1);
            for (size_t k = n - 1; k-- > 0;) chain_ok[k] = (uint8_t)(chain_ok[k + 1] && ok[a + n + k]);
            for (size_t k = n; k-- > 0;) {  // .rev(): root first
                if (ok[a + k] && chain_ok[k]) {
                    pick[j] = k;
                    live.push_back(j);
                    next.push_back(chains[j][k]);
                    break;
                }
            }
        }
        std::vector<std::optional<NodePtr>> deeper = optimize_many(next, i + 1);
        for (size_t q = 0; q < live.size(); ++q) {
            const size_t j = live[q];
            out[j] = deeper[q] ? std::make_shared<Node>(starts[j]->get_coord(), *deeper[q]) : cands[j][pick[j]];
        }
        return out;
    }
    // check_finish for many nodes: batched optimize, one count + fill over all final chains, one verify launch;
    // the same lines as check_finish(node) per node
    std::vector<std::optional<LineString>> check_finish_many(const std::vector<NodePtr> &from) const {
        std::vector<std::optional<LineString>> res(from.size());
        if (from.empty()) return res;
        std::vector<std::optional<NodePtr>> opt = optimize_many(from, 0);
        std::vector<detail_rrt::Edges> edges(from.size());
        detail_rrt::Edges all;
        for (size_t j = 0; j < from.size(); ++j) {
            NodePtr top = std::make_shared<Node>(Node::new_goal(goal_, opt[j] ? *opt[j] : from[j], goal_yaw_));
            edges[j] = detail_rrt::chain_edges(top);
            all.sx.insert(all.sx.end(), edges[j].sx.begin(), edges[j].sx.end());
            all.sy.insert(all.sy.end(), edges[j].sy.begin(), edges[j].sy.end());
            all.syaw.insert(all.syaw.end(), edges[j].syaw.begin(), edges[j].syaw.end());
            all.ex.insert(all.ex.end(), edges[j].ex.begin(), edges[j].ex.end());
            all.ey.insert(all.ey.end(), edges[j].ey.begin(), edges[j].ey.end());
            all.eyaw.insert(all.eyaw.end(), edges[j].eyaw.begin(), edges[j].eyaw.end());
        }
        const size_t m = all.sx.size();
        std::vector<uint32_t> counts(m);
        std::vector<unsigned char> plan(m * PP_DUBINS_PLAN_BYTES);
        detail::check(pp_dubins_sample_count(space_->ctx(), m, all.sx.data(), all.sy.data(), all.syaw.data(), all.ex.data(),
                                             all.ey.data(), all.eyaw.data(), space_->get_steer(), step_size_, 0,
                                             counts.data(), plan.data()),
                      "sample_count");
        std::vector<uint64_t> offsets(m);
        uint64_t total = 0;
        for (size_t i = 0; i < m; ++i) {
            offsets[i] = total;
            total += (counts[i] == 0xFFFFFFFFu) ? 0 : counts[i];
        }
        std::vector<double> xyyaw(3 * total);
        detail::check(pp_dubins_sample_fill(space_->ctx(), m, plan.data(), offsets.data(), total, xyyaw.data()),
                      "sample_fill");
        std::vector<LineString> lines(from.size());
        size_t pos = 0;
        for (size_t j = 0; j < from.size(); ++j) {
            LineString fwd;  // node -> root order, the root's own point left out (finalize drops it, :532)
            for (size_t k = 0; k < edges[j].sx.size(); ++k, ++pos) {
                if (plan[pos * PP_DUBINS_PLAN_BYTES + 104] == PP_WORD_NONE)  // finalize's loop, src/rrt.rs:529
                    throw std::runtime_error(detail_rrt::SHOULD_PLAN);
                for (uint64_t s = offsets[pos]; s < offsets[pos] + counts[pos]; ++s) fwd.push(xyyaw[3 * s], xyyaw[3 * s + 1]);
            }
            for (size_t k = fwd.size(); k-- > 0;) lines[j].push(fwd.x[k], fwd.y[k]);  // :538 reverse
        }
        std::vector<uint8_t> good = space_->verify_many(lines);
        for (size_t j = 0; j < from.size(); ++j)
            if (good[j]) res[j] = std::move(lines[j]);
        return res;
    }
    // Rounds of `batch` samples against one tree snapshot (the reference's four racy workers at width `batch`,
    // src/rrt.rs:600-609): one pp_rrt_extend_dubins call (NN -> Node::new yaw -> fused Dubins verify of the new
    // edges; the parents' chains are verified already, that is the tree invariant), one batched append,
    // check_finish_many for the round's fresh nodes; min_by length on the host.
    std::optional<LineString> plan_rounds(size_t batch = 256) {
        std::optional<LineString> best;
        double best_len = std::numeric_limits<double>::infinity();
        const double steer = space_->get_steer();
        for (size_t budget = max_iter_; budget > 0;) {
            const size_t b = std::min(batch, budget);
            budget -= b;
            std::vector<double> px(b), py(b), yaw(b);
            for (size_t k = 0; k < b; ++k) {
                const Point p = space_->rand_point();
                px[k] = p.x;
                py[k] = p.y;
            }
            std::vector<uint32_t> idx(b);
            std::vector<uint8_t> ok(b);
            detail::check(pp_rrt_extend_dubins(space_->ctx(), b, px.data(), py.data(), steer, step_size_, idx.data(),
                                               yaw.data(), ok.data(), 0, 0),
                          "rrt_extend_dubins");
            std::vector<NodePtr> fresh;
            std::vector<double> ax, ay, ayaw;
            std::vector<int32_t> apar;
            for (size_t k = 0; k < b; ++k) {
                if (!ok[k] || idx[k] == 0xFFFFFFFFu) continue;  // no nearest node: get_random_node's None
                NodePtr n = std::make_shared<Node>(Point{px[k], py[k]}, nodes_[idx[k]]);
                n->slot = (int64_t)(nodes_.size() + fresh.size());
                fresh.push_back(n);
                ax.push_back(px[k]);
                ay.push_back(py[k]);
                ayaw.push_back(n->get_yaw());
                apar.push_back((int32_t)nodes_[idx[k]]->slot);
            }
            if (fresh.empty()) continue;
            detail::check(pp_tree_append(space_->ctx(), fresh.size(), ax.data(), ay.data(), ayaw.data(), apar.data()),
                          "tree_append");
            nodes_.insert(nodes_.end(), fresh.begin(), fresh.end());
            // check_finish for every fresh node, as plan_one does (src/rrt.rs:591): the goal connects to the OPTIMIZED
            // node, whose yaw differs from the fresh node's, so a blocked goal -> fresh edge decides nothing
            const std::vector<NodePtr> &visible = fresh;
            for (std::optional<LineString> &line : check_finish_many(visible)) {
                if (!line) continue;
                const double len = line->euclidean_length();
                if (len < best_len) {
                    best_len = len;
                    best = std::move(line);
                }
            }
        }
        return best;
    }
    size_t tree_size() const { return nodes_.size(); }
    const std::vector<NodePtr> &nodes() const { return nodes_; }

   private:
    void insert(const NodePtr &n) {  // src/rrt.rs:586-589
        std::lock_guard<std::mutex> lk(mu_);
        n->slot = (int64_t)nodes_.size();
        nodes_.push_back(n);
        const double x = n->get_point().x, y = n->get_point().y, yaw = n->get_yaw();
        const int32_t par = n->get_parent() ? (int32_t)n->get_parent()->slot : -1;
        detail::check(pp_tree_append(space_->ctx(), 1, &x, &y, &yaw, &par), "tree_append");
    }
    Coordinate goal_;
    double goal_yaw_;
    size_t max_iter_;
    double step_size_;
    std::shared_ptr<Space> space_;
    std::vector<NodePtr> nodes_;  // slot i <-> GPU tree slot i (replaces the RTree, src/rrt.rs:345-346)
    mutable std::mutex mu_;
};

}  // namespace rrt
}  // namespace pathplanning

// fake_pp.cpp -- TEST DOUBLE of the C-ABI (include/pathplanning_b200.h) for the CPU test suite only.
// It answers the entry points the C++ host mirror (rs-pathplanning_b200/host/pathplanning.hpp) uses from the CPU
// oracle (oracle/pp_oracle.c), so that the mirror's host logic -- Arc-like node graph, tree slots, batched
// optimize / check_finish, plan_rounds, the JSON world reader of example_rrt -- can run without a device.
// NOT product code: it lives under tests/, is never installed next to libpathplanning_b200.so and the product has
// no CPU back end (pp_ctx_create of the real library fails with PP_ERR_NO_DEVICE without an sm_100 GPU).
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "pathplanning_b200.h"
#include "pp_oracle.h"

struct pp_ctx {
    std::vector<double> x, y, yaw;
    std::vector<int32_t> parent;
    std::vector<double> bx, by, ox, oy;
    std::vector<uint32_t> off;
    bool have_world = false;
    std::string err;
    ppo_world world() const { return ppo_world{bx.data(), by.data(), bx.size(), ox.data(), oy.data(), off.data(), off.size() - 1}; }
};

struct fake_plan {  // what the fill pass needs to replay the path; word sits where the real record keeps it
    double v[8];    // sx sy syaw ex ey eyaw radius step
    int32_t from_origin;
};
static_assert(sizeof(fake_plan) <= 104, "fake plan must leave byte 104 (the word) free");

static int fail(pp_ctx *c, int rc, const char *what) {
    if (c) c->err = what;
    return rc;
}

extern "C" {

const char *pp_status_string(int s) { return s == PP_OK ? "ok" : "error (fake ABI)"; }
const char *pp_last_error(pp_ctx *c) { return c ? c->err.c_str() : "null context"; }
int pp_ctx_create(int, pp_ctx **out) {
    if (!out) return PP_ERR_INVALID;
    *out = new pp_ctx();
    return PP_OK;
}
void pp_ctx_destroy(pp_ctx *c) { delete c; }

int pp_mod2pi(pp_ctx *, size_t n, const double *x, double *out, int pi_2_pi) {
    for (size_t i = 0; i < n; ++i) out[i] = pi_2_pi ? ppo_pi_2_pi(x[i]) : ppo_mod2pi(x[i]);
    return PP_OK;
}

int pp_dubins_words(pp_ctx *, size_t n, const double *alpha, const double *beta, const double *d, double *tpq, uint8_t *feasible) {
    for (size_t i = 0; i < n; ++i)
        for (int w = 0; w < 6; ++w) {
            double t[3];
            int f = ppo_dubins_word(w, alpha[i], beta[i], d[i], t);
            feasible[6 * i + w] = (uint8_t)f;
            for (int k = 0; k < 3; ++k) tpq[18 * i + 3 * w + k] = f ? t[k] : NAN;
        }
    return PP_OK;
}

int pp_dubins_eval(pp_ctx *c, size_t n, const double *sx, const double *sy, const double *syaw, const double *ex,
                   const double *ey, const double *eyaw, const double *radius_arr, double radius, double *cost,
                   uint8_t *word, double *tpq) {
    if (!radius_arr && !(radius > 0)) return fail(c, PP_ERR_INVALID, "radius");
    for (size_t i = 0; i < n; ++i) {
        double t[3], co = INFINITY;
        uint32_t fl;
        int w = ppo_dubins_eval(sx[i], sy[i], syaw[i], ex[i], ey[i], eyaw[i], radius_arr ? radius_arr[i] : radius, &co, t, &fl);
        word[i] = (uint8_t)w;
        cost[i] = w == PPO_NONE ? INFINITY : co;
        if (tpq)
            for (int k = 0; k < 3; ++k) tpq[3 * i + k] = w == PPO_NONE ? NAN : t[k];
    }
    return PP_OK;
}

static long run_path(const fake_plan &p, std::vector<double> &px, std::vector<double> &py, std::vector<double> &pyaw,
                     int *word, double *cost) {
    size_t cap = 4096;
    for (;;) {
        px.resize(cap), py.resize(cap), pyaw.resize(cap);
        long n = ppo_dubins_path(p.v[0], p.v[1], p.v[2], p.v[3], p.v[4], p.v[5], p.v[6], p.v[7], p.from_origin, px.data(),
                                 py.data(), pyaw.data(), cap, word, cost, nullptr);
        if (n != -2) return n;
        cap *= 4;
    }
}

int pp_dubins_sample_count(pp_ctx *c, size_t n, const double *sx, const double *sy, const double *syaw, const double *ex,
                           const double *ey, const double *eyaw, double radius, double step, int from_origin,
                           uint32_t *counts, void *plan) {
    if (!(radius > 0) || !(step > 0)) return fail(c, PP_ERR_INVALID, "radius/step");
    std::vector<double> px, py, pyaw;
    for (size_t i = 0; i < n; ++i) {
        fake_plan p{{sx[i], sy[i], syaw[i], ex[i], ey[i], eyaw[i], radius, step}, from_origin};
        int w = PPO_NONE;
        double co;
        long k = run_path(p, px, py, pyaw, &w, &co);
        if (k < -1) return fail(c, PP_ERR_OVERFLOW, "the reference would panic on this path");
        uint8_t *rec = (uint8_t *)plan + i * PP_DUBINS_PLAN_BYTES;
        std::memset(rec, 0, PP_DUBINS_PLAN_BYTES);
        std::memcpy(rec, &p, sizeof p);
        rec[104] = k < 0 ? (uint8_t)PP_WORD_NONE : (uint8_t)w;
        counts[i] = k < 0 ? 0u : (uint32_t)k;
    }
    return PP_OK;
}

int pp_dubins_sample_fill(pp_ctx *, size_t n, const void *plan, const uint64_t *offsets, uint64_t, double *out) {
    std::vector<double> px, py, pyaw;
    for (size_t i = 0; i < n; ++i) {
        const uint8_t *rec = (const uint8_t *)plan + i * PP_DUBINS_PLAN_BYTES;
        if (rec[104] == PP_WORD_NONE) continue;
        fake_plan p;
        std::memcpy(&p, rec, sizeof p);
        int w;
        double co;
        long k = run_path(p, px, py, pyaw, &w, &co);
        for (long j = 0; j < k; ++j) {
            double *o = out + 3 * (offsets[i] + (uint64_t)j);
            o[0] = px[(size_t)j], o[1] = py[(size_t)j], o[2] = pyaw[(size_t)j];
        }
    }
    return PP_OK;
}

int pp_dubins_path(pp_ctx *c, double sx, double sy, double syaw, double ex, double ey, double eyaw, double radius,
                   double step, int from_origin, double *px, double *py, double *pyaw, size_t cap, size_t *n_out,
                   int *word, double *cost) {
    if (!(radius > 0) || !(step > 0)) return fail(c, PP_ERR_INVALID, "radius/step");
    fake_plan p{{sx, sy, syaw, ex, ey, eyaw, radius, step}, from_origin};
    std::vector<double> x, y, yw;
    int w = PPO_NONE;
    double co = INFINITY;
    long k = run_path(p, x, y, yw, &w, &co);
    if (k < -1) return fail(c, PP_ERR_OVERFLOW, "the reference would panic on this path");
    if (word) *word = k < 0 ? (int)PP_WORD_NONE : w;
    if (cost) *cost = k < 0 ? INFINITY : co;
    if (n_out) *n_out = k < 0 ? 0 : (size_t)k;
    if (k > (long)cap) return fail(c, PP_ERR_OVERFLOW, "cap");
    for (long j = 0; j < k; ++j) px[j] = x[(size_t)j], py[j] = y[(size_t)j], pyaw[j] = yw[(size_t)j];
    return PP_OK;
}

int pp_tree_append(pp_ctx *c, size_t k, const double *x, const double *y, const double *yaw, const int32_t *parent) {
    for (size_t i = 0; i < k; ++i) {
        c->x.push_back(x[i]), c->y.push_back(y[i]);
        c->yaw.push_back(yaw ? yaw[i] : 0.0), c->parent.push_back(parent ? parent[i] : -1);
    }
    return PP_OK;
}
int pp_tree_upload(pp_ctx *c, size_t n, const double *x, const double *y, const double *yaw, const int32_t *parent) {
    c->x.clear(), c->y.clear(), c->yaw.clear(), c->parent.clear();
    return pp_tree_append(c, n, x, y, yaw, parent);
}
size_t pp_tree_size(pp_ctx *c) { return c ? c->x.size() : 0; }

int pp_obstacles_upload(pp_ctx *c, const double *bx, const double *by, size_t nb, const double *rx, const double *ry,
                        const uint32_t *off, size_t n_rings) {
    if (!bx || !by || nb < 3) return fail(c, PP_ERR_INVALID, "bounds");
    c->bx.assign(bx, bx + nb), c->by.assign(by, by + nb);
    c->off.assign(1, 0u);
    if (n_rings) c->off.assign(off, off + n_rings + 1);
    c->ox.assign(rx, rx + c->off.back()), c->oy.assign(ry, ry + c->off.back());
    c->have_world = true;
    return PP_OK;
}

int pp_nn(pp_ctx *c, size_t m, const double *qx, const double *qy, uint32_t *idx, double *d2, int) {
    std::vector<double> tmp(m);
    ppo_nn_brute(c->x.size(), c->x.data(), c->y.data(), m, qx, qy, idx, d2 ? d2 : tmp.data(), nullptr, 1);
    return PP_OK;
}

int pp_verify_polylines(pp_ctx *c, size_t n_lines, const double *px, const double *py, const uint32_t *line_off,
                        uint8_t *ok, int) {
    if (!c->have_world) return fail(c, PP_ERR_STATE, "obstacles not uploaded");
    ppo_world w = c->world();
    for (size_t i = 0; i < n_lines; ++i)
        ok[i] = (uint8_t)ppo_verify(&w, px + line_off[i], py + line_off[i], line_off[i + 1] - line_off[i]);
    return PP_OK;
}

int pp_collide_dubins(pp_ctx *c, size_t m, const double *sx, const double *sy, const double *syaw, const double *ex,
                      const double *ey, const double *eyaw, double radius, double step, uint8_t *ok, int) {
    if (!c->have_world) return fail(c, PP_ERR_STATE, "obstacles not uploaded");
    ppo_world w = c->world();
    ppo_verify_dubins_edges(&w, m, sx, sy, syaw, ex, ey, eyaw, radius, step, ok, 0, 1);
    return PP_OK;
}

int pp_rrt_extend_dubins(pp_ctx *c, size_t m, const double *qx, const double *qy, double radius, double step,
                         uint32_t *idx, double *yaw, uint8_t *ok, int, int) {
    if (!c->have_world || c->x.empty()) return fail(c, PP_ERR_STATE, "tree or obstacles not uploaded");
    std::vector<double> d2(m), nx(m), ny(m), nyaw(m);
    ppo_nn_brute(c->x.size(), c->x.data(), c->y.data(), m, qx, qy, idx, d2.data(), nullptr, 1);
    for (size_t i = 0; i < m; ++i) {
        nx[i] = c->x[idx[i]], ny[i] = c->y[idx[i]], nyaw[i] = c->yaw[idx[i]];
        yaw[i] = ppo_compute_yaw(qx[i], qy[i], nx[i], ny[i]);  // Node::new, src/rrt.rs:169-175
    }
    ppo_world w = c->world();
    ppo_verify_dubins_edges(&w, m, qx, qy, yaw, nx.data(), ny.data(), nyaw.data(), radius, step, ok, 0, 1);
    return PP_OK;
}

}  // extern "C"

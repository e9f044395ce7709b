//! extern "C" mirror of include/pathplanning_b200.h (ABI version 2): one context per Space (world + tree), one lazy
//! crate-level context for the stateless Dubins calls, and the multi-GPU group.
#![allow(non_camel_case_types, dead_code)]
use lazy_static::lazy_static;
use std::os::raw::{c_char, c_double, c_int, c_void};

#[repr(C)]
pub struct pp_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct pp_group {
    _private: [u8; 0],
}

pub const PP_OK: c_int = 0;
pub const PP_ERR_OVERFLOW: c_int = -6;
pub const PP_WORD_NONE: c_int = 0xFF;
pub const PP_DUBINS_PLAN_BYTES: usize = 112;
pub const PP_PLAN_WORD_OFFSET: usize = 104;

extern "C" {
    pub fn pp_abi_version() -> c_int;
    pub fn pp_status_string(status: c_int) -> *const c_char;
    pub fn pp_device_count() -> c_int;
    pub fn pp_ctx_create(device: c_int, out: *mut *mut pp_ctx) -> c_int;
    pub fn pp_ctx_destroy(ctx: *mut pp_ctx);
    pub fn pp_last_error(ctx: *mut pp_ctx) -> *const c_char;
    pub fn pp_sync(ctx: *mut pp_ctx) -> c_int;
    pub fn pp_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn pp_host_free(p: *mut c_void) -> c_int;

    pub fn pp_mod2pi(ctx: *mut pp_ctx, n: usize, x: *const c_double, out: *mut c_double, pi_2_pi: c_int) -> c_int;
    pub fn pp_dubins_words(ctx: *mut pp_ctx, n: usize, alpha: *const c_double, beta: *const c_double,
                           d: *const c_double, tpq: *mut c_double, feasible: *mut u8) -> c_int;
    pub fn pp_dubins_eval(ctx: *mut pp_ctx, n: usize, sx: *const c_double, sy: *const c_double,
                          syaw: *const c_double, ex: *const c_double, ey: *const c_double, eyaw: *const c_double,
                          radius_arr: *const c_double, radius: c_double, cost: *mut c_double, word: *mut u8,
                          tpq: *mut c_double) -> c_int;
    pub fn pp_dubins_sample_count(ctx: *mut pp_ctx, n: usize, sx: *const c_double, sy: *const c_double,
                                  syaw: *const c_double, ex: *const c_double, ey: *const c_double,
                                  eyaw: *const c_double, radius: c_double, step: c_double, from_origin: c_int,
                                  counts: *mut u32, plan: *mut c_void) -> c_int;
    pub fn pp_dubins_sample_fill(ctx: *mut pp_ctx, n: usize, plan: *const c_void, offsets: *const u64, total: u64,
                                 out_xyyaw: *mut c_double) -> c_int;
    pub fn pp_dubins_path(ctx: *mut pp_ctx, sx: c_double, sy: c_double, syaw: c_double, ex: c_double, ey: c_double,
                          eyaw: c_double, radius: c_double, step: c_double, from_origin: c_int, px: *mut c_double,
                          py: *mut c_double, pyaw: *mut c_double, cap: usize, n_out: *mut usize, word: *mut c_int,
                          cost: *mut c_double) -> c_int;

    pub fn pp_tree_upload(ctx: *mut pp_ctx, n: usize, x: *const c_double, y: *const c_double, yaw: *const c_double,
                          parent: *const i32) -> c_int;
    pub fn pp_tree_append(ctx: *mut pp_ctx, k: usize, x: *const c_double, y: *const c_double, yaw: *const c_double,
                          parent: *const i32) -> c_int;
    pub fn pp_tree_size(ctx: *mut pp_ctx) -> usize;
    pub fn pp_obstacles_upload(ctx: *mut pp_ctx, bounds_x: *const c_double, bounds_y: *const c_double,
                               n_bounds: usize, ring_x: *const c_double, ring_y: *const c_double,
                               ring_off: *const u32, n_rings: usize) -> c_int;
    pub fn pp_nn(ctx: *mut pp_ctx, m: usize, qx: *const c_double, qy: *const c_double, idx: *mut u32,
                 d2: *mut c_double, flags: c_int) -> c_int;
    pub fn pp_collide_segments(ctx: *mut pp_ctx, m: usize, ax: *const c_double, ay: *const c_double,
                               bx: *const c_double, by: *const c_double, ok: *mut u8, flags: c_int) -> c_int;
    pub fn pp_verify_polylines(ctx: *mut pp_ctx, n_lines: usize, px: *const c_double, py: *const c_double,
                               line_off: *const u32, ok: *mut u8, flags: c_int) -> c_int;
    pub fn pp_collide_dubins(ctx: *mut pp_ctx, m: usize, sx: *const c_double, sy: *const c_double,
                             syaw: *const c_double, ex: *const c_double, ey: *const c_double, eyaw: *const c_double,
                             radius: c_double, step: c_double, ok: *mut u8, flags: c_int) -> c_int;
    pub fn pp_rrt_extend(ctx: *mut pp_ctx, m: usize, qx: *const c_double, qy: *const c_double, idx: *mut u32,
                         yaw: *mut c_double, ok: *mut u8, nn_flags: c_int, collide_flags: c_int) -> c_int;
    pub fn pp_rrt_extend_dubins(ctx: *mut pp_ctx, m: usize, qx: *const c_double, qy: *const c_double,
                                radius: c_double, step: c_double, idx: *mut u32, yaw: *mut c_double, ok: *mut u8,
                                nn_flags: c_int, collide_flags: c_int) -> c_int;

    // one host process, several B200s: replicated tree / obstacles (ncclBroadcast), sliced batches
    pub fn pp_group_create(devices: *const c_int, n_dev: c_int, out: *mut *mut pp_group) -> c_int;
    pub fn pp_group_destroy(g: *mut pp_group);
    pub fn pp_group_size(g: *mut pp_group) -> c_int;
    pub fn pp_group_ctx(g: *mut pp_group, i: c_int) -> *mut pp_ctx;
    pub fn pp_group_last_error(g: *mut pp_group) -> *const c_char;
    pub fn pp_group_tree_upload(g: *mut pp_group, n: usize, x: *const c_double, y: *const c_double,
                                yaw: *const c_double, parent: *const i32) -> c_int;
    pub fn pp_group_tree_append(g: *mut pp_group, k: usize, x: *const c_double, y: *const c_double,
                                yaw: *const c_double, parent: *const i32) -> c_int;
    pub fn pp_group_obstacles_upload(g: *mut pp_group, bounds_x: *const c_double, bounds_y: *const c_double,
                                     n_bounds: usize, ring_x: *const c_double, ring_y: *const c_double,
                                     ring_off: *const u32, n_rings: usize) -> c_int;
    pub fn pp_group_dubins_eval(g: *mut pp_group, n: usize, sx: *const c_double, sy: *const c_double,
                                syaw: *const c_double, ex: *const c_double, ey: *const c_double,
                                eyaw: *const c_double, radius_arr: *const c_double, radius: c_double,
                                cost: *mut c_double, word: *mut u8, tpq: *mut c_double) -> c_int;
    pub fn pp_group_collide_dubins(g: *mut pp_group, m: usize, sx: *const c_double, sy: *const c_double,
                                   syaw: *const c_double, ex: *const c_double, ey: *const c_double,
                                   eyaw: *const c_double, radius: c_double, step: c_double, ok: *mut u8,
                                   flags: c_int) -> c_int;
    pub fn pp_group_rrt_extend(g: *mut pp_group, m: usize, qx: *const c_double, qy: *const c_double, idx: *mut u32,
                               yaw: *mut c_double, ok: *mut u8, nn_flags: c_int, collide_flags: c_int) -> c_int;
    pub fn pp_group_rrt_extend_dubins(g: *mut pp_group, m: usize, qx: *const c_double, qy: *const c_double,
                                      radius: c_double, step: c_double, idx: *mut u32, yaw: *mut c_double,
                                      ok: *mut u8, nn_flags: c_int, collide_flags: c_int) -> c_int;
}

/// an owned pp_ctx; internally synchronised, so `&Ctx` may be shared by rayon workers.  A context holds ONE
/// obstacle set and ONE tree: every `rrt::Space` creates its own (`Ctx::new`), so planners never clobber each other.
pub struct Ctx(pub *mut pp_ctx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}

impl Ctx {
    pub fn new() -> std::sync::Arc<Ctx> {
        let dev = std::env::var("PP_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut p: *mut pp_ctx = std::ptr::null_mut();
        let rc = unsafe { pp_ctx_create(dev, &mut p) };
        // no CPU fallback: same failure style as the reference's expect()s (src/rrt.rs:64,67)
        assert!(rc == PP_OK, "pathplanning_b200: no sm_100 GPU context (status {})", rc);
        std::sync::Arc::new(Ctx(p))
    }
}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { pp_ctx_destroy(self.0) }
    }
}

lazy_static! {
    /// crate-level context of the STATELESS calls only (the `dubins` module, Dubins sampling of chains): they never
    /// touch a tree or an obstacle set, so sharing it is harmless
    pub static ref CTX: std::sync::Arc<Ctx> = Ctx::new();
}

pub fn check(rc: c_int, what: &str) {
    if rc != PP_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(pp_last_error(CTX.0)) }.to_string_lossy().into_owned();
        panic!("pathplanning_b200: {} failed with status {}: {}", what, rc, msg);
    }
}

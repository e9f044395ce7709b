//! Public API of the reference's src/dubins.rs, unchanged; arithmetic on the GPU through the C-ABI.
use crate::ffi::{self, CTX};

#[derive(Clone)]
pub enum Mode {
    L,
    S,
    R,
}

type ModeSlice = Option<&'static [Mode; 3]>;
type PlannerResult = (Option<f64>, Option<f64>, Option<f64>, ModeSlice);

const LSL_MODE: ModeSlice = Some(&[Mode::L, Mode::S, Mode::L]);
const RSR_MODE: ModeSlice = Some(&[Mode::R, Mode::S, Mode::R]);
const LSR_MODE: ModeSlice = Some(&[Mode::L, Mode::S, Mode::R]);
const RSL_MODE: ModeSlice = Some(&[Mode::R, Mode::S, Mode::L]);
const RLR_MODE: ModeSlice = Some(&[Mode::R, Mode::L, Mode::R]);
const LRL_MODE: ModeSlice = Some(&[Mode::L, Mode::R, Mode::L]);
const WORD_MODES: [ModeSlice; 6] = [LSL_MODE, RSR_MODE, LSR_MODE, RSL_MODE, RLR_MODE, LRL_MODE];

pub fn mod2pi(theta: f64) -> f64 {
    let mut out = 0.0;
    ffi::check(unsafe { ffi::pp_mod2pi(CTX.0, 1, &theta, &mut out, 0) }, "mod2pi");
    out
}

pub fn pi_2_pi(angle: f64) -> f64 {
    let mut out = 0.0;
    ffi::check(unsafe { ffi::pp_mod2pi(CTX.0, 1, &angle, &mut out, 1) }, "pi_2_pi");
    out
}

fn word(w: usize, alpha: f64, beta: f64, d: f64) -> PlannerResult {
    let mut tpq = [0.0f64; 18];
    let mut feas = [0u8; 6];
    ffi::check(
        unsafe { ffi::pp_dubins_words(CTX.0, 1, &alpha, &beta, &d, tpq.as_mut_ptr(), feas.as_mut_ptr()) },
        "dubins_words",
    );
    if feas[w] == 0 {
        (None, None, None, WORD_MODES[w])
    } else {
        (Some(tpq[3 * w]), Some(tpq[3 * w + 1]), Some(tpq[3 * w + 2]), WORD_MODES[w])
    }
}

pub fn lsl(alpha: f64, beta: f64, d: f64) -> PlannerResult { word(0, alpha, beta, d) }
pub fn rsr(alpha: f64, beta: f64, d: f64) -> PlannerResult { word(1, alpha, beta, d) }
pub fn lsr(alpha: f64, beta: f64, d: f64) -> PlannerResult { word(2, alpha, beta, d) }
pub fn rsl(alpha: f64, beta: f64, d: f64) -> PlannerResult { word(3, alpha, beta, d) }
pub fn rlr(alpha: f64, beta: f64, d: f64) -> PlannerResult { word(4, alpha, beta, d) }
pub fn lrl(alpha: f64, beta: f64, d: f64) -> PlannerResult { word(5, alpha, beta, d) }

type DubinsPath = (Vec<f64>, Vec<f64>, Vec<f64>, ModeSlice, f64);
type DubinsPathResult = Option<DubinsPath>;

pub struct DubinsConfig {
    pub sx: f64,
    pub sy: f64,
    pub syaw: f64,
    pub ex: f64,
    pub ey: f64,
    pub eyaw: f64,
    pub turn_radius: f64,
    pub step_size: f64,
}

fn path(sx: f64, sy: f64, syaw: f64, ex: f64, ey: f64, eyaw: f64, radius: f64, step: f64, from_origin: i32)
    -> DubinsPathResult {
    let mut cap = 4096usize;
    loop {
        let (mut px, mut py, mut pyaw) = (vec![0.0; cap], vec![0.0; cap], vec![0.0; cap]);
        let (mut n, mut w, mut cost) = (0usize, 0i32, 0.0f64);
        let rc = unsafe {
            ffi::pp_dubins_path(CTX.0, sx, sy, syaw, ex, ey, eyaw, radius, step, from_origin, px.as_mut_ptr(),
                                py.as_mut_ptr(), pyaw.as_mut_ptr(), cap, &mut n, &mut w, &mut cost)
        };
        if rc == ffi::PP_ERR_OVERFLOW && n > cap {
            cap = n;
            continue;
        }
        ffi::check(rc, "dubins_path");
        if w == ffi::PP_WORD_NONE {
            return None;
        }
        px.truncate(n);
        py.truncate(n);
        pyaw.truncate(n);
        return Some((px, py, pyaw, WORD_MODES[w as usize], cost));
    }
}

pub fn dubins_path_planning_from_origin(dx: f64, dy: f64, eyaw: f64, c: f64, step_size: f64) -> DubinsPathResult {
    path(0.0, 0.0, 0.0, dx, dy, eyaw, 1.0 / c, step_size, 1)
}

pub fn dubins_path_planning(conf: &DubinsConfig) -> DubinsPathResult {
    path(conf.sx, conf.sy, conf.syaw, conf.ex, conf.ey, conf.eyaw, conf.turn_radius, conf.step_size, 0)
}

/// batched entry points (new): what the GPU path is for
pub mod batch {
    use crate::ffi::{self, CTX};

    /// cost (radius-normalised; +inf = no feasible word) and word id (0..5 in ALL_PLANNERS order, 0xFF = None)
    pub fn eval(sx: &[f64], sy: &[f64], syaw: &[f64], ex: &[f64], ey: &[f64], eyaw: &[f64], turn_radius: f64)
        -> (Vec<f64>, Vec<u8>) {
        let n = sx.len();
        assert!(sy.len() == n && syaw.len() == n && ex.len() == n && ey.len() == n && eyaw.len() == n);
        let (mut cost, mut word) = (vec![0.0f64; n], vec![0u8; n]);
        ffi::check(
            unsafe {
                ffi::pp_dubins_eval(CTX.0, n, sx.as_ptr(), sy.as_ptr(), syaw.as_ptr(), ex.as_ptr(), ey.as_ptr(),
                                    eyaw.as_ptr(), std::ptr::null(), turn_radius, cost.as_mut_ptr(),
                                    word.as_mut_ptr(), std::ptr::null_mut())
            },
            "dubins_eval",
        );
        (cost, word)
    }
}

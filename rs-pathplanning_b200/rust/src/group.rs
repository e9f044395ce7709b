//! Every B200 of the box from ONE host process: a safe wrapper over `pp_group` (include/pathplanning_b200.h,
//! "multi-GPU").  Tree and obstacles are replicated by `ncclBroadcast` inside the library (the whole tree at
//! `RRT::new`, src/rrt.rs:345-346; only the appended tail at the insert site, src/rrt.rs:586-589); every batch call is
//! cut into contiguous slices, one per device, and the results land in the caller's vectors -- byte-identical for
//! every device count.  Source only (no rustc in this image).
use crate::ffi;
use std::os::raw::c_int;

pub struct Group(*mut ffi::pp_group);
unsafe impl Send for Group {}
unsafe impl Sync for Group {}

fn check(g: &Group, rc: c_int, what: &str) {
    if rc != ffi::PP_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(ffi::pp_group_last_error(g.0)) }.to_string_lossy().into_owned();
        panic!("pathplanning_b200: {} failed with status {}: {}", what, rc, msg);
    }
}

impl Group {
    /// `devices`: CUDA ordinals, e.g. `&[0, 1, 2, 3, 4, 5, 6, 7]`
    pub fn new(devices: &[i32]) -> Group {
        let mut p: *mut ffi::pp_group = std::ptr::null_mut();
        let rc = unsafe { ffi::pp_group_create(devices.as_ptr(), devices.len() as c_int, &mut p) };
        assert!(rc == ffi::PP_OK, "pathplanning_b200: pp_group_create failed with status {} (B200s + libnccl.so.2)", rc);
        Group(p)
    }
    pub fn len(&self) -> usize { unsafe { ffi::pp_group_size(self.0) as usize } }

    pub fn tree_upload(&self, x: &[f64], y: &[f64], yaw: &[f64], parent: &[i32]) {
        let rc = unsafe { ffi::pp_group_tree_upload(self.0, x.len(), x.as_ptr(), y.as_ptr(), yaw.as_ptr(), parent.as_ptr()) };
        check(self, rc, "group_tree_upload");
    }
    /// src/rrt.rs:586-589: the new nodes reach every replica as a tail-only broadcast
    pub fn tree_append(&self, x: &[f64], y: &[f64], yaw: &[f64], parent: &[i32]) {
        let rc = unsafe { ffi::pp_group_tree_append(self.0, x.len(), x.as_ptr(), y.as_ptr(), yaw.as_ptr(), parent.as_ptr()) };
        check(self, rc, "group_tree_append");
    }
    /// bounds ring + obstacle rings in CSR form (ring r = points off[r] .. off[r + 1])
    pub fn obstacles_upload(&self, bx: &[f64], by: &[f64], ox: &[f64], oy: &[f64], off: &[u32]) {
        let rc = unsafe {
            ffi::pp_group_obstacles_upload(self.0, bx.as_ptr(), by.as_ptr(), bx.len(), ox.as_ptr(), oy.as_ptr(),
                                           off.as_ptr(), off.len().saturating_sub(1))
        };
        check(self, rc, "group_obstacles_upload");
    }
    /// batched `dubins_path_planning` evaluation half: (cost, word) per pose pair
    pub fn dubins_eval(&self, sx: &[f64], sy: &[f64], syaw: &[f64], ex: &[f64], ey: &[f64], eyaw: &[f64],
                       turn_radius: f64) -> (Vec<f64>, Vec<u8>) {
        let n = sx.len();
        let (mut cost, mut word) = (vec![0.0f64; n], vec![0u8; n]);
        let rc = unsafe {
            ffi::pp_group_dubins_eval(self.0, n, sx.as_ptr(), sy.as_ptr(), syaw.as_ptr(), ex.as_ptr(), ey.as_ptr(),
                                      eyaw.as_ptr(), std::ptr::null(), turn_radius, cost.as_mut_ptr(),
                                      word.as_mut_ptr(), std::ptr::null_mut())
        };
        check(self, rc, "group_dubins_eval");
        (cost, word)
    }
    /// one batched extend step with the reference's real edge (src/rrt.rs:406-426): (nearest index, yaw, ok)
    pub fn rrt_extend_dubins(&self, qx: &[f64], qy: &[f64], turn_radius: f64, step_size: f64)
        -> (Vec<u32>, Vec<f64>, Vec<u8>) {
        let m = qx.len();
        let (mut idx, mut yaw, mut ok) = (vec![0u32; m], vec![0.0f64; m], vec![0u8; m]);
        let rc = unsafe {
            ffi::pp_group_rrt_extend_dubins(self.0, m, qx.as_ptr(), qy.as_ptr(), turn_radius, step_size,
                                            idx.as_mut_ptr(), yaw.as_mut_ptr(), ok.as_mut_ptr(), 0, 0)
        };
        check(self, rc, "group_rrt_extend_dubins");
        (idx, yaw, ok)
    }
    /// `RRT::verify_node` per edge on all devices
    pub fn collide_dubins(&self, e: &[Vec<f64>; 6], turn_radius: f64, step_size: f64) -> Vec<u8> {
        let m = e[0].len();
        let mut ok = vec![0u8; m];
        let rc = unsafe {
            ffi::pp_group_collide_dubins(self.0, m, e[0].as_ptr(), e[1].as_ptr(), e[2].as_ptr(), e[3].as_ptr(),
                                         e[4].as_ptr(), e[5].as_ptr(), turn_radius, step_size, ok.as_mut_ptr(), 0)
        };
        check(self, rc, "group_collide_dubins");
        ok
    }
}

impl Drop for Group {
    fn drop(&mut self) {
        unsafe { ffi::pp_group_destroy(self.0) }
    }
}

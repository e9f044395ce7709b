//! Crate `pathplanning`, B200 back end.  Module layout as in the reference (`dubins`, `rrt`), plus `ffi`:
//! the extern "C" block for libpathplanning_b200.so and the process-wide GPU context.
//! Source only: this image has no rustc, see ../../INTEGRATION.md for how a maintainer builds it.
pub mod ffi;

pub mod dubins;
pub mod rrt;

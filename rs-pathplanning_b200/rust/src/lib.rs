//! Crate `pathplanning`, B200 back end.  Module layout as in the reference (`dubins`, `rrt`), plus `ffi`:
//! the extern "C" block for libpathplanning_b200.so (one context per `rrt::Space`), and `group`: every GPU of the box
//! from one process.
//! Source only: this image has no rustc, see ../../INTEGRATION.md for how a maintainer builds it.
pub mod ffi;

pub mod dubins;
pub mod group;
pub mod rrt;

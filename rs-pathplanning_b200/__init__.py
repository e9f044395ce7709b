"""pathplanning-b200: B200-native (sm_100a CUDA) hot path of the Rust crate `pathplanning`
(tsturzl/rs-pathplanning): batched Dubins evaluation / sampling and the RRT extend step.

Layout
  csrc/      hand-written CUDA kernels + the C-ABI (libpathplanning_b200.so)
  _ffi.py    ctypes binding of include/pathplanning_b200.h
  dubins.py  host mirror of the crate's `dubins` module (same names and argument meaning)
  rrt.py     host mirror of the crate's `rrt` module
  synth.py   counter-based synthetic inputs of SURVEY.md section 8(d)
  host/      C++ mirror of the crate API over the C-ABI (+ examples / benches entry points)
  rust/      source-only Rust shim (no rustc in this image)

The directory name carries a hyphen (it is the reference's name + `_b200`), so the package is
imported under the module name `rs_pathplanning_b200` through `__graft_entry__.import_package()`.
There is no CPU fallback anywhere in this package.
"""
from . import _ffi  # noqa: F401  (raises ImportError when the CUDA library has not been built)
from ._ffi import Context, Group, PathPlanningError, PinnedArray, comm_unique_id, device_count, slice_bounds  # noqa: F401
from . import synth  # noqa: F401

_default_ctx = None


def default_context(device: int = 0) -> "Context":
    """crate-level lazy context used by the scalar drop-in functions"""
    global _default_ctx
    if _default_ctx is None or _default_ctx._h is None:
        _default_ctx = Context(device)
    return _default_ctx


from . import dubins, rrt  # noqa: E402,F401

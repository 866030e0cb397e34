"""rosdyn_b200 - B200-native batched engine for the hot path of rosdyn_core's `rosdyn::Chain`.

`Chain` (host mirror of the reference interface) sits on the C-ABI library `librosdyn_b200.so`
(include/rosdyn_b200.h), whose kernels are hand-written sm_100a fp64 CUDA.  No CPU fallback.
"""
from .descriptor import FIXED, PRISMATIC, REVOLUTE, ChainDesc, JointDesc, LinkDesc, rpy_to_rot  # noqa: F401
from . import fixtures  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not need the built library (descriptor-only users)
    if name in ("Chain", "createChain", "fill_uniform", "fp64_peak", "kernel_launch_count"):
        from . import chain as _chain
        return getattr(_chain, name)
    raise AttributeError(name)

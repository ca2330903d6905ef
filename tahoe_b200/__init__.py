"""tahoe_b200 -- B200-native (sm_100a, FP64) implementation of Tahoe's Hex8 continuum-solid hot path.

Layout: csrc/ hand-written CUDA kernels + the extern "C" layer (include/tahoe_b200.h), host/ the C++ plugin classes that
forward Tahoe's ElementBaseT / GlobalMatrixT / integrator interfaces to it, capi.py the ctypes binding used by the harness
(tests, bench), mesh.py synthetic structured hex meshes and the element partitioner.
"""
from . import capi, mesh  # noqa: F401

__all__ = ["capi", "mesh"]

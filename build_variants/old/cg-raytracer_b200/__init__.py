"""cg-raytracer_b200: B200-native hot path of CG-RayTracer (BVH traversal + ray/AABB + ray/triangle under getFinalColor).

Layout:
  csrc/   CUDA kernels (sm_100a) + the extern "C" boundary (include/cgrt_b200.h) + the host BVH builder
  host/   C++ mirror of the reference interface (BoundingVolumeHierarchy, intersectRayWith*, renderRayTracing, loader)
  capi.py ctypes binding of the C ABI, used by tests/ and bench.py (plumbing only)

The directory name carries a hyphen, so import it through `__graft_entry__.load_package()` (registers the module as
`cg_raytracer_b200`).
"""
from . import capi  # noqa: F401
from .capi import CgrtError, Scene, load_library  # noqa: F401

"""raytracinginaweekend_b200 — B200-native render hot path of martingleich/RayTracingInAWeekend.

Layout:
  csrc/   hand-written sm_100a CUDA kernels + the C ABI (include/b200rt.h)  -> libb200rt.so
  host/   C++ host mirror of the reference's src/lib API + scene authoring  -> librtiaw_host.so (include/rtiaw_host.h)
  api.py  Python face of the host mirror (World, render, RenderMode, Size2i)
  dist.py one-process-per-GPU sample-lane sharding + NCCL film combine
"""
from .api import FlatScene, RenderMode, Size2i, World, earth_rgb8, load_obj_mesh, render, to_rgb8_gamma2  # noqa: F401

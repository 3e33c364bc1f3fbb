"""lens_flare_b200 -- B200-native lens-flare ghost engine (sm_100a CUDA behind a C ABI).

  capi        ctypes binding of include/lfb200.h (liblfb200.so, built in-tree)
  pathtracer  host-side mirror of the reference's call surface for this path
              (CameraApertureTexture, PathTracer.find_sun_pos / generate_ghost_buffer / ghost_buffer)
  sharding    (light x ghost pair x wavelength) sharding across GPUs + the NCCL reduce
  csrc/       the kernels and the C ABI; host/ the C++ facade for the reference application
"""
from . import capi  # noqa: F401

__all__ = ["capi"]

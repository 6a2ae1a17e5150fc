"""meshfem_b200 -- B200-native linear-elasticity assemble-and-solve path behind
MeshFEM's operator surface.

Layout:
  csrc/      hand-written sm_100a CUDA + the C ABI (include/mfem_b200.h) -> lib/libmfem_b200.so
  capi.py    ctypes binding of the C ABI (what tests and bench.py call)
  build.py   nvcc build recipe
The host C++ mirror of the reference classes lives in include/MeshFEM/ and src/.
"""
from .capi import Handle, MfemB200Error, SolveInfo, load_library, LIB_PATH  # noqa: F401

__all__ = ["Handle", "MfemB200Error", "SolveInfo", "load_library", "LIB_PATH"]

"""mesh_deform_b200 -- B200 (sm_100a) engine for the ARAP solve path of cheind/mesh-deform.

The product is the C++ header API in inc/deform/ on top of the C ABI in include/arap_b200.h,
implemented by hand-written CUDA kernels in mesh_deform_b200/csrc/ (libarap_b200.so).
This Python package is only a ctypes binding of that C ABI (used by tests and bench.py) plus the
synthetic workload generators. There is no CPU fallback anywhere in the package.
"""
from . import capi, meshgen  # noqa: F401
from .capi import AsRigidAsPossibleDeformation, ArapError, BatchDeformation, EngineMissingError  # noqa: F401

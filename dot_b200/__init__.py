"""dot_b200 - B200-native DOT hot path (libdotgpu) and its thin Python host mirror.

The product is the C-ABI shared library `libdotgpu.so` (include/dotgpu.h) built from csrc/ by
`python -m dot_b200.build`.  This package only binds it with ctypes for tests and bench.py, naming
things after the reference's classes (Energy, LinSysSolver -> Solver, AnimScripter -> Anim,
DOTTimeStepper -> Stepper).  There is no CPU fallback: compute entry points raise DotGpuError when no
CUDA device is visible, and importing the binding fails loudly if the library has not been built.
"""
from .api import (ANIM_KINDS, ENERGY_FCR, ENERGY_SNH, Anim, DD, DotGpuError, Energy, FrameStats, Solver, Stepper,  # noqa: F401
                  device_count, lib, lib_path, mesh_features, nccl_unique_id, owned_subdomains, partition, partition_nodes, balanced_owner)

"""hypre_b200 — B200-native BoomerAMG solve phase (PCG/GMRES) behind hypre's own interface.

The package is a thin host-side mirror of the reference API over libhb200.so (the C-ABI in
include/hb200.h, hand-written sm_100a CUDA + NCCL).  Importing it requires the built library:
there is no CPU fallback.
"""
from .solver import *  # noqa: F401,F403
from .solver import __all__  # noqa: F401

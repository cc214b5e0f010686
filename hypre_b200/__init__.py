"""hypre_b200 — B200-native BoomerAMG solve phase (PCG/GMRES) behind hypre's own interface.

The package is a thin host-side mirror of the reference API over libhb200.so (the C-ABI in
include/hb200.h, hand-written sm_100a CUDA + NCCL).  Using it requires the built library:
there is no CPU fallback.  (`hypre_b200.build` can be imported without the library — it is what
builds it; everything else loads libhb200.so on first access and raises if it is missing or
does not match the header.)
"""
import importlib

_EXPORTS = None


def __getattr__(name):
    if name in ("build", "_lib", "solver"):
        return importlib.import_module("." + name, __name__)
    solver = importlib.import_module(".solver", __name__)
    if name == "__all__":
        return solver.__all__
    try:
        return getattr(solver, name)
    except AttributeError:
        raise AttributeError(f"module 'hypre_b200' has no attribute {name!r}") from None

"""Test infrastructure: run the `-m gpu` parity tests against the host emulation of the kernels.

Activated by tests/conftest.py only when HB200_EMU_TEST=1 (set by tests/test_emu_kernels.py in a
child pytest process).  It changes nothing in the product package:
  * `hypre_b200._lib` is pre-loaded from the product's own _lib.py source with LIB_PATH pointing at
    oracle/_ref/libhb200_emu.so (the product .cu sources compiled by g++ against oracle/emu);
  * torch "cuda" tensors become CPU tensors (the emulated device memory is host memory).
"""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_LIB = os.environ.get("HB200_EMU_LIB") or os.path.join(ROOT, "oracle", "_ref", "libhb200_emu.so")


def activate():
    if not os.path.exists(EMU_LIB):
        raise RuntimeError(f"{EMU_LIB} missing: make -C oracle emu")
    # ---- the ctypes binding, bound to the emulated library
    path = os.path.join(ROOT, "hypre_b200", "_lib.py")
    src = open(path).read()
    needle = 'LIB_PATH = os.path.join(HERE, "libhb200.so")'
    assert needle in src
    src = src.replace(needle, f"LIB_PATH = {EMU_LIB!r}")
    import hypre_b200                                    # the package itself (lazy, loads nothing yet)
    mod = types.ModuleType("hypre_b200._lib")
    mod.__file__ = path
    mod.__package__ = "hypre_b200"
    sys.modules["hypre_b200._lib"] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    hypre_b200._lib = mod
    # ---- torch: "cuda" means host memory here
    import torch
    torch.cuda.is_available = lambda: True
    torch.cuda.device_count = lambda: 1
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.set_device = lambda *a, **k: None
    # host <-> "device" transfers copy, as the real ones do (no aliasing with the numpy source)
    torch.Tensor.cuda = lambda self, *a, **k: self.clone()
    torch.Tensor.cpu = lambda self, *a, **k: self.clone()
    for name in ("empty", "zeros", "ones", "full", "randn", "rand", "tensor", "empty_like", "zeros_like"):
        orig = getattr(torch, name)

        def wrapped(*a, __orig=orig, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k["device"] = "cpu"
            return __orig(*a, **k)
        setattr(torch, name, wrapped)

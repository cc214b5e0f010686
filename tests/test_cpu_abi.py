"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/hb200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb200_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from hypre_b200._lib import LIB_PATH, SYMBOLS
    lib = C.CDLL(LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 50
    missing = [s for s in decl if not hasattr(lib, s)]
    assert not missing, f"declared in hb200.h but not exported: {missing}"
    # the Python binding covers the whole header as well
    assert sorted(SYMBOLS) == decl


def test_header_cites_reference_lines():
    src = open(os.path.join(ROOT, "include", "hb200.h")).read()
    for ref in ("par_csr_matvec.c:241", "par_csr_matvec.c:523", "par_relax.c:23", "par_cycle.c:23",
                "par_amg_solve.c:22", "pcg.c:313", "gmres.c:294", "par_gauss_elim.c:457",
                "par_csr_communication.h:51-75"):
        assert ref in src, ref


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hypre_b200._lib import lib
    assert lib.hb200_init(0) != 0
    assert b"no CPU fallback" in lib.hb200_last_error()
    # every compute entry point refuses to run
    out = C.c_void_p()
    assert lib.hb200_malloc(C.byref(out), 64) != 0
    assert lib.hb200_vec_set(None, 0.0, 16) != 0
    assert lib.hb200_parcsr_matvec(None, 1.0, None, 0.0, None, None) != 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hypre_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c")) and f != "hypre_shim.c":
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "refbridge" not in txt and "oracle/" not in txt.replace("oracle/_ref", "").replace(
                    "reference bridge", ""), os.path.join(dp, f)


def test_dropin_ij_fails_loudly_without_gpu():
    """The unmodified ij driver linked in front of the shim: without a GPU the overridden solve must
    refuse (hypre error + message), never fall back to the reference's CPU solve behind it."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ij = os.path.join(ROOT, "oracle", "_ref", "ij_b200")
    if not os.path.exists(ij):
        pytest.skip("oracle/_ref/ij_b200 not built (needs /root/reference)")
    r = subprocess.run([ij, "-n", "12", "12", "12", "-solver", "1", "-rlx", "18"], capture_output=True, text=True,
                       timeout=300, cwd=os.path.dirname(ij))
    out = r.stdout + r.stderr
    assert "[hypre_b200] FATAL" in out and "no CPU fallback" in out, out[-2000:]
    # the reference's own CPU PCG would report 8 iterations here; the refused solve reports none
    assert "Iterations = 0" in out, out[-2000:]

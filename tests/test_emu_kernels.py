"""The `-m gpu` parity tests, run on the CPU against a host emulation of the kernels.

oracle/_ref/libhb200_emu.so is the product's own hypre_b200/csrc/*.cu compiled by g++ against
oracle/emu/cuda_runtime.h (one fiber per CUDA thread, switched at __syncthreads and the warp
shuffles).  It checks the LOGIC of every kernel and of the host code around it — formats,
epilogues, reductions, wavefront GS, cycle, PCG, GMRES — against the compiled reference on a
machine without a GPU; it says nothing about performance and the product never loads it.  The
real parity gate stays `pytest -m gpu` on a B200."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "oracle", "_ref", "libhb200_emu.so")
BRIDGE = os.path.join(ROOT, "oracle", "_ref", "libref_bridge.so")


def build_emu():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "emu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert os.path.exists(EMU)


def run_child(extra_env, *pytest_args, timeout=1500):
    env = dict(os.environ, HB200_EMU_TEST="1", **extra_env)
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", *pytest_args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


def test_gpu_parity_suite_on_the_host_emulation():
    if not os.path.exists(BRIDGE):
        pytest.skip("oracle/_ref/libref_bridge.so not built (needs /root/reference)")
    build_emu()
    # everything that runs on one device except the tests that exec the real shim binary
    r = run_child({}, os.path.join("tests", "test_gpu_parity.py"), "-m", "gpu", "-n", "6",
                  "-k", "not ij_dropin and not multi_gpu")
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail


def test_wide_pattern_format_on_the_host_emulation():
    """the experimental 16-bit-code / global-table variant of the row-pattern kernel (HB200_PAT_WIDE=1):
    a block the 1-byte format rejects runs through spmv_pat<..., WIDE> and matches the row sums"""
    build_emu()
    r = run_child({"HB200_PAT_WIDE": "1"}, os.path.join("tests", "emu_wide_case.py"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_fused_dots_on_the_host_emulation(tmp_path):
    """<s,p> out of the Krylov matvec epilogue and <r,z> out of the cycle's last Jacobi sweep: same
    iteration count and residual as with separate dot kernels, two launches fewer per iteration"""
    import json
    if not os.path.exists(BRIDGE):
        pytest.skip("oracle/_ref/libref_bridge.so not built (needs /root/reference)")
    build_emu()
    reports = {}
    for name, env in (("fused", {"HB200_FUSED_DOTS": "1"}), ("separate", {})):
        path = str(tmp_path / f"{name}.json")
        r = run_child(dict(env, HB200_EMU_REPORT=path), os.path.join("tests", "emu_fused_dots_case.py"))
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
        reports[name] = json.load(open(path))
    f, s = reports["fused"], reports["separate"]
    assert f["iterations"] == s["iterations"] == f["ref_iterations"]
    assert abs(f["rel_res"] - s["rel_res"]) <= 1e-9 * s["rel_res"]
    assert abs(f["x_norm"] - s["x_norm"]) <= 1e-12 * s["x_norm"]
    assert s["launches"] - f["launches"] == 2 * f["iterations"], (f, s)


def test_kernels_under_address_sanitizer():
    """the same emulation compiled with -fsanitize=address: no kernel (or host path around it) reads or
    writes outside its "device" buffers on the SpMV formats, the relaxation family and a PCG solve"""
    if not os.path.exists(BRIDGE):
        pytest.skip("oracle/_ref/libref_bridge.so not built (needs /root/reference)")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "emu_asan"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan.so not found")
    env = {"HB200_EMU_LIB": os.path.join(ROOT, "oracle", "_ref", "libhb200_emu_asan.so"), "LD_PRELOAD": asan,
           "ASAN_OPTIONS": "detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1"}
    r = run_child(env, os.path.join("tests", "test_gpu_parity.py"), "-m", "gpu", "-n", "6",
                  "-k", "matvec or format or pattern or relax or cheby or coarse or pcg_amg or gmres")
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0 and "AddressSanitizer" not in tail, tail
    assert " passed" in r.stdout, tail

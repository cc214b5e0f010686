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
    r = run_child({}, os.path.join("tests", "test_gpu_parity.py"), os.path.join("tests", "test_ij_formats.py"), "-m", "gpu",
                  "-n", "8", "-k", "not ij_dropin and not multi_gpu")
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
    for name, env in (("fused", {"HB200_FUSED_DOTS": "1"}), ("separate", {"HB200_FUSED_DOTS": "0"})):
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
    writes outside its "device" buffers on the SpMV formats, the Jacobi / Chebyshev sweeps, the BLAS-1 kernels and the IJ
    assembly (the chunked Gauss-Seidel sweep, the batched dots and the new Krylov drivers pass under the sanitizer as well:
    add their test names to -k)
    (the whole suite passes under the sanitizer as well: `make -C oracle emu_asan`, then the command below without -k)"""
    if not os.path.exists(BRIDGE):
        pytest.skip("oracle/_ref/libref_bridge.so not built (needs /root/reference)")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "emu_asan"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan.so not found")
    env = {"HB200_EMU_LIB": os.path.join(ROOT, "oracle", "_ref", "libhb200_emu_asan.so"), "LD_PRELOAD": asan,
           "ASAN_OPTIONS": "detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1"}
    r = run_child(env, os.path.join("tests", "test_gpu_parity.py"), os.path.join("tests", "test_ij_formats.py"), "-m", "gpu",
                  "-n", "6", "-k", "matvec or format_detection or pattern or blas1 or binary_ij")
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0 and "AddressSanitizer" not in tail, tail
    assert " passed" in r.stdout, tail


def build_emu_mpi():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "emu_mpi"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def run_ranks(nproc, script, *args, timeout=900, extra_env=None):
    """one host process per rank under torch.distributed.run: emulated kernels, NCCL calls over
    oracle/minimpi, the IPC arena of the peer-put halo in POSIX shared memory"""
    env = dict(os.environ, HB200_EMU_TEST="1", OMP_NUM_THREADS="1",
               HB200_EMU_LIB=os.path.join(ROOT, "oracle", "_ref", "libhb200_emu_mpi.so"))
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + nproc + 10 * len(args)),
           os.path.join(ROOT, "tests", script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


@pytest.mark.parametrize("nproc,halo,fuse_wait", [(8, "nccl", False)])   # (4, "peer", True) passes too: on demand
def test_multi_rank_parity_on_the_host_emulation(nproc, halo, fuse_wait):
    """tests/mp_parity_worker.py (the worker of the multi-GPU parity tests) on N host processes: maps,
    SpMV / SpMV-T, relaxation, cycle, PCG and GMRES against the reference running on the same ranks —
    the NCCL halo at 8 ranks (2 x 2 x 2 bricks, 7 neighbours).  The peer-put protocol runs in the bench and
    ij tests below (2 ranks); `run_ranks(4, "mp_parity_worker.py", "27pt", "peer")` and 8 ranks pass as well."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_bridge_mpi.so")):
        pytest.skip("oracle/_ref/libref_bridge_mpi.so not built (needs /root/reference)")
    build_emu_mpi()
    # fuse_wait: the opt-in kernel that waits for the peer puts and runs the offd pass in one launch
    r = run_ranks(nproc, "mp_parity_worker.py", "27pt", halo, extra_env={"HB200_FUSE_WAIT": "1"} if fuse_wait else None)
    assert r.returncode == 0 and "MULTI-RANK PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_peer_halo_watchdog_on_the_host_emulation():
    """a polling halo kernel whose peer never shows up gives up after HB200_HALO_TIMEOUT_S and the
    library reports it (hb_peer.cuh, halo_check_error) — no unbounded spin on the device"""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_bridge_mpi.so")):
        pytest.skip("oracle/_ref/libref_bridge_mpi.so not built (needs /root/reference)")
    build_emu_mpi()
    r = run_ranks(2, "emu_halo_timeout_worker.py", timeout=300, extra_env={"HB200_HALO_TIMEOUT_S": "1"})
    assert r.returncode == 0 and "WATCHDOG OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_bench_code_path_on_the_host_emulation():
    """bench.py itself, unchanged, on 2 emulated ranks: stages and watchdog, halo choice (auto -> peer
    puts at N = 2), per-level kernel timing, the JSON contract.  Its numbers mean nothing here."""
    import json
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_bridge_mpi.so")):
        pytest.skip("oracle/_ref/libref_bridge_mpi.so not built (needs /root/reference)")
    build_emu_mpi()
    r = run_ranks(2, "emu_bench_runner.py", "--gpus", "2", "--size", "12", "--steps", "1", "--warmup", "1",
                  "--no-cpu-baseline")
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1, r.stdout[-3000:] + r.stderr[-3000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == 2 and d["config"]["halo"] == "peer" and d["gpu_launches"] > 0
    assert d["roofline"]["bound"] == "hbm" and d["e2e"]["h2d_bytes_per_step"] > 0
    assert "workload" in d["config"] and d["config"]["iterations"] > 0


def test_random_blocks_through_every_kernel_on_the_host_emulation():
    """awkward CSR blocks (1 x 1, rectangular, empty rows, one dense row, no entries at all) through
    every SpMV kernel kind, every (alpha, beta) branch and the stored transpose, against scipy"""
    build_emu()
    r = run_child({}, os.path.join("tests", "emu_fuzz_case.py"), "-n", "6")
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_smoke_entry_point_on_the_host_emulation():
    """__graft_entry__.smoke() — the call the driver makes on the GPU box — runs its code path here"""
    if not os.path.exists(BRIDGE):
        pytest.skip("oracle/_ref/libref_bridge.so not built (needs /root/reference)")
    build_emu()
    r = run_child({}, os.path.join("tests", "emu_smoke_case.py"), "-s")
    assert r.returncode == 0 and "smoke ok" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def _ij_run(binary, args, nprocs=1, env_extra=None):
    import re
    ref = os.path.join(ROOT, "oracle", "_ref")
    exe = os.path.join(ref, binary)
    cmd = [exe, *args.split()]
    if nprocs > 1:
        cmd = [os.path.join(ref, "mpirun"), "-np", str(nprocs)] + cmd
    env = dict(os.environ, OMP_NUM_THREADS="1", HYPRE_B200_VERBOSE="1")
    env.update(env_extra or {})
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ref, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    its = re.findall(r"Iterations = (\d+)", r.stdout)
    res = re.findall(r"Final (?:\w+ )?Relative Residual Norm = ([0-9.eE+-]+)", r.stdout)
    return int(its[-1]), float(res[-1]), r.stderr


@pytest.mark.parametrize("args,nprocs", [("-27pt -n 18 18 18 -solver 1 -rlx 18", 1),
                                         ("-27pt -n 24 14 14 -P 2 1 1 -solver 1 -rlx 18", 2),
                                         ("-vardifconv -n 10 10 10 -solver 9 -rlx 18", 1),     # AMG-BiCGSTAB
                                         ("-vardifconv -n 16 9 9 -P 2 1 1 -solver 16 -rlx 18", 2),     # AMG-COGMRES
                                         ("-27pt -n 20 12 12 -P 2 1 1 -solver 10", 2)])        # DS-BiCGSTAB
def test_ij_dropin_through_the_shim_on_the_host_emulation(args, nprocs):
    """the UNMODIFIED reference driver linked in front of hypre_shim.c, the shim bound to the emulated
    library: the interposition, the hierarchy hand-over and (2 ranks) the halo-transport choice of
    the shim run here; iteration counts and residuals against the reference's own solve"""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("needs /root/reference to build the ij driver")
    for target in ("ij", "ij_mpi", "emu_shim"):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", target], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    ref_bin, dev_bin = ("ij_ref", "ij_b200_emu") if nprocs == 1 else ("ij_refmpi", "ij_b200_emu_mpi")
    its_ref, res_ref, _ = _ij_run(ref_bin, args, nprocs)
    its_dev, res_dev, err = _ij_run(dev_bin, args, nprocs)
    assert "on device" in err, err[-1500:]
    assert its_dev == its_ref, (args, its_dev, its_ref)
    solver_id = int(args.split("-solver")[1].split()[0])
    rtol = 2e-6 if solver_id in (1, 2) else 5e-2
    assert abs(res_dev - res_ref) <= rtol * res_ref and res_dev < 1e-8, (args, res_dev, res_ref)


def test_default_ij_invocation_prints_the_reference_tables_on_the_host_emulation():
    """`ij` with its defaults is the stand-alone BoomerAMG solver at print level 3: parameter table, residual and
    convergence factor of every cycle, grid / operator / cycle complexities.  The drop-in runs it on the device and
    prints the reference's output, line for line."""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("needs /root/reference to build the ij driver")
    for target in ("ij", "emu_shim"):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", target], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    ref = os.path.join(ROOT, "oracle", "_ref")
    env = dict(os.environ, OMP_NUM_THREADS="1", HYPRE_B200_VERBOSE="1")
    outs = {}
    for binary in ("ij_ref", "ij_b200_emu"):
        r = subprocess.run([os.path.join(ref, binary), "-27pt", "-n", "8", "8", "8", "-rlx", "18", "-mu", "2"],
                           capture_output=True, text=True, timeout=600, cwd=ref, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[binary] = ([l for l in r.stdout.splitlines() if l.strip() and "wall clock" not in l and "cpu clock" not in l
                         and "seconds" not in l], r.stderr)
    assert "BoomerAMG on device" in outs["ij_b200_emu"][1], outs["ij_b200_emu"][1][-1500:]
    assert any(l.startswith("    Cycle  1") for l in outs["ij_ref"][0]) and any("cycle =" in l for l in outs["ij_ref"][0])
    assert outs["ij_b200_emu"][0] == outs["ij_ref"][0]


def test_hybrid_gs_chunks_through_the_shim_on_the_host_emulation():
    """the reference's default smoother (hybrid l1-GS 13 / 14) depends on its thread count: with
    HYPRE_B200_GS_CHUNKS=host the drop-in reproduces the 3-thread reference digit for digit (one launch per sweep);
    with a chunk count of its own (5) it reproduces the 5-thread reference, l1 norms of that partition included,
    whatever the host's thread count is"""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("needs /root/reference to build the ij driver")
    for target in ("ij", "emu_shim"):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", target], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    args = "-27pt -n 8 7 7 -solver 1"
    its3, res3, _ = _ij_run("ij_ref", args, env_extra={"OMP_NUM_THREADS": "3"})
    its5, res5, _ = _ij_run("ij_ref", args, env_extra={"OMP_NUM_THREADS": "5"})
    assert abs(res3 - res5) > 1e-6 * res3          # the thread count does change the reference's numbers
    its, res, err = _ij_run("ij_b200_emu", args, env_extra={"OMP_NUM_THREADS": "3", "HYPRE_B200_GS_CHUNKS": "host"})
    assert "on device" in err and its == its3 and abs(res - res3) <= 2e-6 * res3, (its, its3, res, res3)
    its, res, err = _ij_run("ij_b200_emu", args, env_extra={"OMP_NUM_THREADS": "2", "HYPRE_B200_GS_CHUNKS": "5"})
    assert "on device" in err and its == its5 and abs(res - res5) <= 2e-6 * res5, (its, its5, res, res5)


def test_random_krylov_options_on_the_host_emulation():
    """12 random PCG / GMRES option combinations (norms, flexible, relative change, recomputed residuals,
    tolerances, restart, preconditioner, x0, b, early max_iter exit) against the 1-thread reference.
    (`tests/emu_option_sweep_case.py` does the same for BoomerAMG options; run on demand, minutes.)"""
    if not os.path.exists(BRIDGE):
        pytest.skip("oracle/_ref/libref_bridge.so not built (needs /root/reference)")
    build_emu()
    r = run_child({"HB200_SWEEP_CASES": "12"}, os.path.join("tests", "emu_krylov_sweep_case.py"), "-n", "4")
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]

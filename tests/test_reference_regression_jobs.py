"""The reference's own regression jobs for this path (src/test/TEST_ij/smoother.jobs, solvers.jobs) through
the drop-in: the unmodified ij driver in front of hypre_shim.c, iteration counts and final residuals
against smoother.saved / solvers.saved (tests/ref_golden_jobs.py holds the table).

CPU suite: four short jobs on the host emulation of the kernels (2, 3 and 4 ranks: PCG, stand-alone
BoomerAMG, Chebyshev smoothing, and a smoother outside the path that the shim must hand back to the
reference); HB200_ALL_GOLDENS=1 runs all 23 (about 8 minutes; last full run: profiles/r1_reference_regression_jobs.txt).
`-m gpu`: every job whose rank count fits the GPUs of the box, on the real library."""
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_golden_jobs as jobs  # noqa: E402

ROOT = jobs.ROOT
QUICK = ("solvers.0", "smoother.8", "smoother.12", "smoother.14")


def _have(*names):
    return all(os.path.exists(os.path.join(jobs.REF, n)) for n in names)


@pytest.mark.parametrize("case", jobs.CASES, ids=[c[0] for c in jobs.CASES])
def test_reference_job_through_the_dropin_on_the_host_emulation(case):
    if case[0] not in QUICK and not os.environ.get("HB200_ALL_GOLDENS"):
        pytest.skip("long emulation case: HB200_ALL_GOLDENS=1 runs it")
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("needs /root/reference to build the ij driver")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "emu_shim"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    o = jobs.run_case(case, "emu")
    assert o["ok"], (case[0], o)


@pytest.mark.gpu
@pytest.mark.parametrize("case", jobs.CASES, ids=[c[0] for c in jobs.CASES])
def test_reference_job_through_the_dropin(case):
    import torch
    if not _have("ij_b200_mpi", "mpirun"):
        pytest.skip("oracle/_ref/ij_b200_mpi not built (needs /root/reference at build time)")
    if torch.cuda.device_count() < case[1]:
        pytest.skip(f"{case[1]} ranks need {case[1]} GPUs")
    o = jobs.run_case(case, "gpu")
    assert o["ok"], (case[0], o)


def test_sub_communicator_matrices_stay_in_the_reference():
    """seq_threshold (TEST_ij solvers.jobs out.105-108): the coarse problem of the sequential coarse AMG lives on
    a sub-communicator (or COMM_SELF); its BoomerAMG solve must not reach the device path, whose NCCL
    communicator, halo plans and all-reduces belong to the communicator the library was bound on (the
    8-rank job hung before hypre_shim.c checked the matrix communicator).  Against the reference's own run."""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("needs /root/reference to build the ij driver")
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import ref_jobs_sweep as sweep
    for target in ("ij_mpi", "emu_shim"):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", target], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for args in ("-n 12 12 12 -P 2 2 1 -seq_th 50 -solver 1 -rlx 18", "-n 12 12 12 -P 2 2 1 -seq_th 50 -solver 1 -rlx 18 -red 1"):
        name, status, note = sweep.one(("seq_th", 4, args.split()), "emu", 300)
        assert status == "ok", (status, note)

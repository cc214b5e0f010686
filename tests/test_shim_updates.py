"""The reference-side binding under the update patterns of real hypre applications: values changed in
place on the same IJ object (HYPRE_IJMatrixInitialize + SetValues + Assemble, the time-stepping /
Newton pattern), a new Setup on it, a HYPRE_BoomerAMGSet* call between two solves without a setup, and
a change of preconditioner.  tests/csrc/shim_update_case.c does all of it through the public HYPRE
API only; it is linked once against the reference alone and once in front of libHYPRE_b200.so, and the
two programs must print the same iteration counts, residuals and solution norms.

CPU suite: the shim in front of the host emulation of the kernels.  `-m gpu`: the real library."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
LINE = re.compile(r"^(step \d.*): iterations (\d+), final rel\. residual ([0-9.eE+-]+), \|x\| ([0-9.eE+-]+)$", re.M)


def _run(binary, n=14, env_extra=None):
    exe = os.path.join(REF, binary)
    if not os.path.exists(exe):
        pytest.skip(f"{binary} not built (make -C oracle shim_tests; needs /root/reference at build time)")
    env = dict(os.environ, OMP_NUM_THREADS="1", HYPRE_B200_VERBOSE="1")
    env.update(env_extra or {})
    r = subprocess.run([exe, str(n)], capture_output=True, text=True, timeout=900, cwd=REF, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = LINE.findall(r.stdout)
    assert len(rows) == 4, r.stdout
    return rows, r.stderr


def _compare(dev_binary):
    ref, _ = _run("shim_update_ref")
    dev, err = _run(dev_binary)
    assert err.count("on device") == 4, err[-1500:]                 # every solve ran through libhb200
    for (what, its_r, res_r, x_r), (_, its_d, res_d, x_d) in zip(ref, dev):
        assert int(its_d) == int(its_r), (what, its_d, its_r)
        assert abs(float(res_d) - float(res_r)) <= 2e-6 * float(res_r), (what, res_d, res_r)
        assert abs(float(x_d) - float(x_r)) <= 1e-9 * float(x_r), (what, x_d, x_r)
    # the three operators differ: a stale device copy would have reproduced step 1's numbers
    assert len({r[3] for r in ref}) >= 3


def test_in_place_updates_through_the_shim_on_the_host_emulation():
    _compare("shim_update_b200_emu")


@pytest.mark.gpu
def test_in_place_updates_through_the_shim():
    if os.environ.get("HB200_EMU_TEST"):
        pytest.skip("covered by the CPU test above")
    _compare("shim_update_b200")

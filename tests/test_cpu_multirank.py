"""world_size-2 (and 3) gloo tests of the N>1 host logic on CPU: the reference on minimpi
reproduces its golden outputs, and the halo plan handed to the GPU path is consistent."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BRIDGE = os.path.join(ROOT, "oracle", "_ref", "libref_bridge_mpi.so")


@pytest.mark.parametrize("world", [2, 3])
def test_multirank_reference_and_halo_plan(world):
    if not os.path.exists(BRIDGE):
        pytest.skip("oracle/_ref/libref_bridge_mpi.so not built (needs /root/reference)")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29610 + world),
           os.path.join(ROOT, "tests", "mp_cpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and "CPU MULTI-RANK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_minimpi_reproduces_reference_golden_outputs():
    """src/test/TEST_ij/solvers.saved, out.19 / out.20 (np = 3 and 4, -rlx 18) through the
    unmodified ij driver on minimpi."""
    ij = os.path.join(ROOT, "oracle", "_ref", "ij_refmpi")
    mpirun = os.path.join(ROOT, "oracle", "_ref", "mpirun")
    if not (os.path.exists(ij) and os.path.exists(mpirun)):
        pytest.skip("oracle/_ref/ij_refmpi not built (needs /root/reference)")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cases = [
        (3, "-n 23 29 31 -solver 1 -rhsrand -precon_cycles 2 -rlx 18", "Iterations = 8", "2.463625e-09"),
        (4, "-n 23 29 31 -solver 3 -rhsrand -precon_cycles 3 -rlx 18", "GMRES Iterations = 7", "5.912905e-10"),
    ]
    for np_, args, its, res in cases:
        r = subprocess.run([mpirun, "-np", str(np_), ij, *args.split()], capture_output=True, text=True,
                           timeout=600, env=env, cwd=os.path.join(ROOT, "oracle", "_ref"))
        assert r.returncode == 0, r.stderr[-2000:]
        assert its in r.stdout and res in r.stdout, r.stdout[-1500:]

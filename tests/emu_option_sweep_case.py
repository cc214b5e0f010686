"""Child of tests/test_emu_kernels.py (needs HB200_EMU_TEST=1, also runs on a GPU): randomised
BoomerAMG option combinations — relaxation types per leg, CF ordering, sweep counts, cycle type,
relaxation weights, coarsening / interpolation — each solved by the reference and by hb200 with
the same hierarchy: same PCG iteration count, same final residual."""
import os

import numpy as np
import pytest


def _runnable():
    if os.environ.get("HB200_EMU_TEST"):
        return True
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


pytestmark = pytest.mark.skipif(not _runnable(), reason="needs the host emulation or a GPU")

JACOBI = [0, 7, 18]
GS = [3, 4, 6, 8, 13, 14, 88, 89]


def draw(rng):
    kind = ["27pt", "laplacian", "vardifconv"][int(rng.integers(3))]
    n = tuple(int(v) for v in rng.integers(9, 17, size=3))
    o = {}
    fam = rng.integers(3)
    if fam == 0:
        o["relax_type"] = int(rng.choice(JACOBI))
    elif fam == 1:
        o["relax_type"] = int(rng.choice(GS))
    else:
        o["relax_type"] = 16
        o["cheby_order"] = int(rng.integers(1, 5))
        o["cheby_scale"] = int(rng.integers(0, 2))
    if rng.random() < 0.3:
        o["relax_down"] = int(rng.choice(JACOBI + GS))
        o["relax_up"] = int(rng.choice(JACOBI + GS))
    if rng.random() < 0.5:
        o["relax_order"] = int(rng.integers(0, 2))
    if rng.random() < 0.4:
        o["num_sweeps"] = int(rng.integers(1, 4))
    if rng.random() < 0.4:
        o["cycle_type"] = int(rng.integers(1, 3))
    if rng.random() < 0.3:
        o["relax_wt"] = float(rng.choice([0.8, 0.9, 1.0]))
    if rng.random() < 0.3:
        o["outer_wt"] = float(rng.choice([0.9, 1.0, 1.1]))
    if rng.random() < 0.3:
        o["coarsen_type"] = int(rng.choice([6, 8, 10]))
    if rng.random() < 0.3:
        o["interp_type"] = int(rng.choice([0, 6, 14]))
    if rng.random() < 0.2:
        o["max_levels"] = int(rng.integers(2, 5))
    return kind, n, o


@pytest.mark.parametrize("seed", range(int(os.environ.get("HB200_SWEEP_CASES", "12"))))
def test_random_amg_options(seed):
    import torch
    import hypre_b200 as hb
    from oracle import refbridge as rb
    hb.init(0)
    rb.load()
    rb.set_num_threads(1)      # the reference's hybrid GS is Gauss-Seidel per OpenMP thread: 1 thread = the sweep hb200 runs
    rng = np.random.default_rng(7000 + seed)
    kind, n, o = draw(rng)
    pb = rb.Problem(kind, n)
    try:
        pb.setup_amg(**o)
    except RuntimeError:
        pytest.skip(f"the reference rejects {o}")
    try:
        mats, amg = hb.amg_from_hierarchy(pb.hierarchy())
    except hb.HB200Error as e:
        # combinations outside the path (the shim forwards those to the reference) must say so loudly
        assert "not on the B200 path" in str(e) or "unsupported" in str(e), (kind, n, o, str(e))
        return
    A = mats[0][0]
    ref = pb.pcg(precond="amg", tol=1e-8, max_iter=60, two_norm=1)
    pcg = hb.ParCSRPCG(tol=1e-8, max_iter=60, two_norm=1)
    pcg.set_precond(amg)
    x = torch.zeros(A.num_rows, dtype=torch.float64).cuda()
    try:
        res = pcg.solve(A, torch.from_numpy(np.array(pb.b)).cuda(), x)
    except hb.HB200Error as e:
        assert "not on the B200 path" in str(e) or "unsupported" in str(e), (kind, n, o, str(e))
        return
    assert abs(res.num_iterations - ref["iterations"]) <= 1, (kind, n, o, res.num_iterations, ref["iterations"])
    if res.num_iterations == ref["iterations"] and ref["final_rel_res"] > 0:
        assert abs(res.rel_residual_norm - ref["final_rel_res"]) <= 1e-5 * ref["final_rel_res"], \
            (kind, n, o, res.rel_residual_norm, ref["final_rel_res"])

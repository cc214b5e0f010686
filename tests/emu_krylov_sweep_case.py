"""Child of tests/test_emu_kernels.py (needs HB200_EMU_TEST=1, also runs on a GPU): randomised PCG /
GMRES option combinations (norm, flexible, relative change, residual recomputation, absolute
tolerance, restart length, preconditioner, non-zero initial guess, random right-hand side, early
max_iter exit) against the reference: iteration count, solution, final residual."""
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
_cache = {}


def problem(kind, n):
    import hypre_b200 as hb
    from oracle import refbridge as rb
    key = (kind, n)
    if key not in _cache:
        pb = rb.Problem(kind, n)
        pb.setup_amg(relax_type=18)
        mats, amg = hb.amg_from_hierarchy(pb.hierarchy())
        _cache[key] = (pb, mats, amg)
    return _cache[key]


@pytest.mark.parametrize("seed", range(int(os.environ.get("HB200_SWEEP_CASES", "24"))))
def test_random_krylov_options(seed):
    import torch
    import hypre_b200 as hb
    from oracle import refbridge as rb
    hb.init(0)
    rb.load()
    rb.set_num_threads(1)
    rng = np.random.default_rng(9000 + seed)
    kind = ["laplacian", "27pt", "vardifconv"][int(rng.integers(3))]
    pb, mats, amg = problem(kind, (12, 11, 10))
    A = mats[0][0]
    n = A.num_rows
    precond = ["amg", "amg", "diagscale", "none"][int(rng.integers(4))]
    b = rng.standard_normal(n) if rng.random() < 0.5 else None
    x0 = rng.standard_normal(n) if rng.random() < 0.5 else None
    tol = float(rng.choice([1e-4, 1e-8, 1e-10]))
    atol = float(rng.choice([0.0, 0.0, 1e-6]))
    max_iter = int(rng.choice([3, 40, 100]))
    symmetric = kind != "vardifconv"
    use_pcg = symmetric and rng.random() < 0.6
    bb = np.array(pb.b) if b is None else b
    xx = np.array(pb.x0) if x0 is None else x0
    if use_pcg:
        opts = dict(two_norm=int(rng.integers(2)), rel_change=int(rng.integers(2)), flex=int(rng.integers(2)),
                    recompute_res=int(rng.integers(2)))
        ref = pb.pcg(precond=precond, tol=tol, atol=atol, max_iter=max_iter, b=b, x0=x0, **opts)
        s = hb.ParCSRPCG(tol=tol, a_tol=atol, max_iter=max_iter, two_norm=opts["two_norm"], rel_change=opts["rel_change"],
                         flex=opts["flex"], recompute_residual=opts["recompute_res"])
    else:
        opts = dict(k_dim=int(rng.choice([2, 5, 10])), rel_change=int(rng.integers(2)))
        ref = pb.gmres(precond=precond, tol=tol, atol=atol, max_iter=max_iter, b=b, x0=x0, **opts)
        s = hb.ParCSRGMRES(tol=tol, a_tol=atol, max_iter=max_iter, k_dim=opts["k_dim"], rel_change=opts["rel_change"])
    s.set_precond(amg if precond == "amg" else ("diagscale" if precond == "diagscale" else None))
    x = torch.from_numpy(xx.copy()).cuda()
    try:
        s.solve(A, torch.from_numpy(bb.copy()).cuda(), x)
    except hb.HB200Error as e:
        # the only error a solve may report here is "not converged within max_iter" (hypre's flag 256)
        assert e.flag & 256, (kind, precond, opts, str(e))
    what = (kind, "pcg" if use_pcg else "gmres", precond, opts, tol, atol, max_iter, b is not None, x0 is not None)
    assert abs(s.num_iterations - ref["iterations"]) <= 1, what + (s.num_iterations, ref["iterations"])
    if s.num_iterations == ref["iterations"]:
        scale = max(np.max(np.abs(ref["x"])), 1e-300)
        assert np.max(np.abs(x.cpu().numpy() - ref["x"])) <= 1e-6 * scale, what

"""Child test of tests/test_emu_kernels.py (needs HB200_EMU_TEST=1): one BoomerAMG-PCG solve on the
host emulation; prints iterations, launches and the residual history for the parent to compare
between HB200_FUSED_DOTS=1 and the default."""
import json
import os

import numpy as np
import pytest

def _runnable():
    if os.environ.get("HB200_EMU_TEST"):
        return True
    try:
        import torch
        return torch.cuda.is_available()         # also runs on a real GPU (scripts/gpu_opt_in_checks.sh)
    except Exception:
        return False


pytestmark = pytest.mark.skipif(not _runnable(), reason="needs the host emulation or a GPU")


def test_pcg_solve_report():
    import torch
    import hypre_b200 as hb
    from oracle import refbridge as rb
    hb.init(0)
    rb.load()
    pb = rb.Problem("27pt", (12, 11, 10))
    pb.setup_amg(relax_type=18)
    mats, amg = hb.amg_from_hierarchy(pb.hierarchy(), use_graph=True)    # the bench's configuration: captured V-cycle
    A = mats[0][0]
    assert A.format_info()["kernel"] in (7, 9)
    ref = pb.pcg(precond="amg", tol=1e-8, max_iter=100, two_norm=1)
    pcg = hb.ParCSRPCG(tol=1e-8, max_iter=100, two_norm=1, logging=1)
    pcg.set_precond(amg)
    x = torch.zeros(A.num_rows, dtype=torch.float64).cuda()
    res = pcg.solve(A, torch.from_numpy(np.array(pb.b)).cuda(), x)
    out = {"iterations": int(res.num_iterations), "launches": int(res.kernel_launches),
           "rel_res": float(res.rel_residual_norm), "ref_iterations": int(ref["iterations"]),
           "ref_rel_res": float(ref["final_rel_res"]), "x_norm": float(torch.linalg.norm(x.cpu()))}
    print("REPORT " + json.dumps(out))
    if os.environ.get("HB200_EMU_REPORT"):
        with open(os.environ["HB200_EMU_REPORT"], "w") as f:
            json.dump(out, f)
    assert out["iterations"] == out["ref_iterations"]
    assert abs(out["rel_res"] - out["ref_rel_res"]) <= 1e-6 * out["ref_rel_res"]

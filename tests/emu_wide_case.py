"""Child test of tests/test_emu_kernels.py (needs HB200_EMU_TEST=1 and HB200_PAT_WIDE=1): not
collected by the normal runs (no test_ prefix in the file name)."""
import os

import numpy as np
import pytest

from test_cpu_formats import stencil7

def _runnable():
    if os.environ.get("HB200_EMU_TEST"):
        return True
    try:
        import torch
        return torch.cuda.is_available()         # also runs on a real GPU (scripts/gpu_opt_in_checks.sh)
    except Exception:
        return False


pytestmark = pytest.mark.skipif(not _runnable(), reason="needs the host emulation or a GPU")


def build_case():
    n, ai, aj, aa = stencil7(16, 16, 16)
    rng = np.random.default_rng(17)
    aa = aa.copy()
    interior = [r for r in range(n) if ai[r + 1] - ai[r] == 7]
    pick = rng.choice(interior, size=2000, replace=False)
    for g in range(0, 2000, 5):
        scale = 1.0 + rng.random(7)
        for r in pick[g:g + 5]:
            aa[ai[r]:ai[r + 1]] *= scale
    return n, ai, aj, aa, rng


def test_wide_pattern_spmv_and_fused_jacobi():
    import torch
    import hypre_b200 as hb
    hb.init(0)
    n, ai, aj, aa, rng = build_case()
    M = hb.ParCSRMatrix(n, n, ai, aj, aa)
    fi = M.format_info()
    assert fi["pattern"] and fi["kernel"] in (7, 9) and fi["patterns"] > 255, fi     # only the wide variant holds that many
    x = rng.standard_normal(n)
    b = rng.standard_normal(n)
    yref = np.zeros(n)
    for r in range(n):
        s = 0.0
        for p in range(ai[r], ai[r + 1]):
            s += aa[p] * x[aj[p]]
        yref[r] = s
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()      # a copy in host memory under the emulation
    y = dev(np.zeros(n))
    M.matvec(1.0, dev(x), 0.0, y)
    got = y.cpu().numpy()
    # pattern rows: CSR order, separate multiply/add -> bit-identical; rows outside the table: CSR lanes
    assert int((got != yref).sum()) <= fi["pattern_irregular_rows"]
    assert np.max(np.abs(got - yref)) <= 1e-13 * np.max(np.abs(yref))
    M.matvec(-1.0, dev(x), 1.0, y, b=dev(b))
    assert np.max(np.abs(y.cpu().numpy() - (b - yref))) <= 1e-13 * np.max(np.abs(yref))
    # the fused l1-Jacobi sweep (EPI_JACOBI7) through the same kernel
    l1 = np.abs(aa[ai[:-1]]) * 1.5
    u = rng.standard_normal(n)
    unew = hb.relax(M, dev(b), dev(u), relax_type=18, l1_norms=dev(l1)).cpu().numpy()
    uref = u + (b - np.array([sum(aa[p] * u[aj[p]] for p in range(ai[r], ai[r + 1])) for r in range(n)])) / l1
    assert np.max(np.abs(np.asarray(unew) - uref)) <= 1e-12 * np.max(np.abs(uref))
    M.destroy()

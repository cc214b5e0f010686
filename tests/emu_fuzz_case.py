"""Child test of tests/test_emu_kernels.py (needs HB200_EMU_TEST=1): random CSR blocks of awkward
shapes through every SpMV kernel kind and the fused epilogues, against scipy."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(not os.environ.get("HB200_EMU_TEST"), reason="emulation child only")


def random_csr(rng, n, m, kind):
    import scipy.sparse as sp
    if kind == "empty":
        return sp.csr_matrix((n, m))
    if kind == "diag":
        A = sp.eye(n, m, format="csr") * 2.0
    elif kind == "dense_row":
        A = sp.random(n, m, density=min(1.0, 3.0 / max(m, 1)), random_state=rng.integers(1 << 30), format="lil")
        A[n // 2, :] = rng.standard_normal(m)            # one full row (longer than any lane count)
        A = A.tocsr()
    elif kind == "empty_rows":
        A = sp.random(n, m, density=min(1.0, 6.0 / max(m, 1)), random_state=rng.integers(1 << 30), format="lil")
        for r in range(0, n, 3):
            A[r, :] = 0.0
        A = A.tocsr()
        A.eliminate_zeros()
    else:
        A = sp.random(n, m, density=min(1.0, rng.integers(1, 40) / max(m, 1)), random_state=rng.integers(1 << 30),
                      format="csr")
    A.sort_indices()
    return A.tocsr()


@pytest.mark.parametrize("seed", range(2))
def test_random_blocks_all_kernels(seed):
    import torch
    import hypre_b200 as hb
    hb.init(0)
    rng = np.random.default_rng(1000 + seed)
    shapes = [(1, 1), (1, 7), (7, 1), (31, 33), (257, 255), (1500, 1500), (2100, 300), (300, 2100)]
    kinds = ["random", "diag", "dense_row", "empty_rows", "empty"]
    for (n, m) in shapes:
        kind = kinds[int(rng.integers(len(kinds)))]
        A = random_csr(rng, n, m, kind)
        ai, aj, aa = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
        M = hb.ParCSRMatrix(n, m, ai, aj, aa)
        x = rng.standard_normal(m)
        b = rng.standard_normal(n)
        yref = A @ x
        scale = max(1.0, float(np.max(np.abs(yref))) if n else 1.0)
        for k, lanes in ((0, 0), (1, 1), (1, 4), (1, 32), (2, 0), (6, 0), (7, 0), (8, 0), (9, 0)):
            M.set_spmv_kernel(k, lanes)
            for alpha, beta in ((1.0, 0.0), (-1.0, 1.0), (0.5, -2.0), (0.0, 3.0)):
                y = torch.from_numpy(b.copy())
                M.matvec(alpha, torch.from_numpy(x.copy()), beta, y)
                err = np.max(np.abs(y.numpy() - (alpha * yref + beta * b))) / scale if n else 0.0
                assert err <= 1e-12, (n, m, kind, k, lanes, alpha, beta, err)
        M.set_spmv_kernel(0, 0)
        # transposed product (stored transpose, built lazily)
        z = torch.from_numpy(x.copy())
        M.matvecT(2.0, torch.from_numpy(b.copy()), -1.0, z)
        zref = 2.0 * (A.T @ b) - x
        if m:
            assert np.max(np.abs(z.numpy() - zref)) <= 1e-12 * max(1.0, float(np.max(np.abs(zref)))), (n, m, kind, "T")
        M.destroy()

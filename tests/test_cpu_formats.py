"""CPU checks of the host half of the upload: the row-pattern analysis (kernels_pat.cu) is lossless
— decoding (row code, base, pattern table, irregular-row list) reproduces the CSR block entry for
entry — qualifies the blocks DESIGN.md says it does, and leaves irregular blocks alone."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BRIDGE = os.path.join(ROOT, "oracle", "_ref", "libref_bridge.so")


def analyze(n, ncols, ai, aj, aa):
    from hypre_b200._lib import lib, check
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    aj = np.ascontiguousarray(aj, dtype=np.int32)
    aa = np.ascontiguousarray(aa, dtype=np.float64)
    code = np.zeros(max(n, 1), np.uint8)
    base = np.zeros(max(n, 1), np.int32)
    ptr = np.zeros(257, np.int32)
    off = np.zeros(8192, np.int32)
    val = np.zeros(8192, np.float64)
    irr = np.zeros(max(n, 1), np.int32)
    npat, nirr = C.c_int(0), C.c_int(0)
    check(lib.hb200_host_pattern_analyze(n, ncols, ai.ctypes.data, aj.ctypes.data, aa.ctypes.data,
                                         code.ctypes.data, base.ctypes.data, C.byref(npat), ptr.ctypes.data,
                                         off.ctypes.data, val.ctypes.data, C.byref(nirr), irr.ctypes.data))
    return {"npat": npat.value, "nirr": nirr.value, "code": code[:n], "base": base[:n],
            "ptr": ptr[:npat.value + 1], "off": off, "val": val, "irr": irr[:nirr.value]}


def analyze_wide(n, ncols, ai, aj, aa, cap=1 << 20):
    from hypre_b200._lib import lib, check
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    aj = np.ascontiguousarray(aj, dtype=np.int32)
    aa = np.ascontiguousarray(aa, dtype=np.float64)
    code = np.zeros(max(n, 1), np.uint16)
    base = np.zeros(max(n, 1), np.int32)
    ptr = np.zeros(65536, np.int32)
    off = np.zeros(cap, np.int32)
    val = np.zeros(cap, np.float64)
    irr = np.zeros(max(n, 1), np.int32)
    npat, nent, nirr = C.c_int(0), C.c_int(0), C.c_int(0)
    check(lib.hb200_host_pattern_analyze_wide(n, ncols, ai.ctypes.data, aj.ctypes.data, aa.ctypes.data,
                                              code.ctypes.data, base.ctypes.data, C.byref(npat), C.byref(nent), cap,
                                              ptr.ctypes.data, off.ctypes.data, val.ctypes.data, C.byref(nirr),
                                              irr.ctypes.data))
    return {"npat": npat.value, "nent": nent.value, "nirr": nirr.value, "code": code[:n], "base": base[:n],
            "ptr": ptr[:npat.value + 1], "off": off, "val": val, "irr": irr[:nirr.value], "none": 65535}


def assert_lossless(n, ai, aj, aa, r):
    """every row with a code decodes to its CSR entries, in CSR order, bit for bit"""
    ai = np.asarray(ai)
    aj = np.asarray(aj)
    aa = np.asarray(aa, dtype=np.float64)
    irr = set(int(x) for x in r["irr"])
    assert len(irr) == r["nirr"]
    for row in range(n):
        c = int(r["code"][row])
        if c == r.get("none", 255):
            assert row in irr
            continue
        assert row not in irr and c < r["npat"]
        b, e = int(r["ptr"][c]), int(r["ptr"][c + 1])
        s, t = int(ai[row]), int(ai[row + 1])
        assert e - b == t - s, row
        assert np.array_equal(r["base"][row] + r["off"][b:e], aj[s:t]), row
        assert np.array_equal(r["val"][b:e].view(np.uint64), aa[s:t].view(np.uint64)), row


def stencil7(nx, ny, nz):
    """7-point Laplacian in the reference's storage: diagonal first, then ascending columns"""
    n = nx * ny * nz
    ai, aj, aa = [0], [], []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                r = i + nx * (j + ny * k)
                aj.append(r)
                aa.append(6.0)
                for d, ok in ((-nx * ny, k > 0), (-nx, j > 0), (-1, i > 0), (1, i < nx - 1), (nx, j < ny - 1),
                              (nx * ny, k < nz - 1)):
                    if ok:
                        aj.append(r + d)
                        aa.append(-1.0)
                ai.append(len(aj))
    return n, np.array(ai, np.int32), np.array(aj, np.int32), np.array(aa)


def test_constant_stencil_has_27_patterns_and_is_lossless():
    n, ai, aj, aa = stencil7(12, 11, 10)
    r = analyze(n, n, ai, aj, aa)
    assert r["npat"] == 27 and r["nirr"] == 0           # 3 x 3 x 3 kinds of boundary
    assert_lossless(n, ai, aj, aa, r)
    assert np.array_equal(r["base"], np.arange(n))      # square block: base = the row


def test_perturbed_rows_go_to_the_irregular_list():
    # 409 one-off rows + 27 regular patterns: more candidates than table slots, so a pattern has to
    # earn its slot with >= 8 rows (with fewer candidates every pattern is kept, see the next test)
    n, ai, aj, aa = stencil7(16, 16, 16)
    rng = np.random.default_rng(5)
    odd = np.sort(rng.choice(n, size=n // 10, replace=False))
    aa = aa.copy()
    for row in odd:
        aa[ai[row]:ai[row + 1]] *= 1.0 + rng.random(ai[row + 1] - ai[row])
    r = analyze(n, n, ai, aj, aa)
    assert r["npat"] > 0
    assert set(odd.tolist()) <= set(r["irr"].tolist())
    # beyond the perturbed rows only patterns too rare to earn a table slot (the 8 corners) are left out
    assert r["nirr"] <= len(odd) + 8
    assert_lossless(n, ai, aj, aa, r)


def test_few_one_off_rows_are_kept_in_the_table():
    n, ai, aj, aa = stencil7(16, 12, 10)
    rng = np.random.default_rng(6)
    aa = aa.copy()
    for row in rng.choice(n, size=40, replace=False):
        aa[ai[row]:ai[row + 1]] *= 1.0 + rng.random(ai[row + 1] - ai[row])
    r = analyze(n, n, ai, aj, aa)
    assert r["nirr"] == 0 and 27 < r["npat"] <= 27 + 40
    assert_lossless(n, ai, aj, aa, r)


def test_irregular_and_tiny_blocks_do_not_qualify():
    rng = np.random.default_rng(9)
    n = 4096
    ai = np.arange(0, 8 * n + 1, 8, dtype=np.int32)
    aj = np.sort(rng.integers(0, n, size=(n, 8)), axis=1).astype(np.int32).ravel()
    aa = rng.standard_normal(8 * n)
    assert analyze(n, n, ai, aj, aa)["npat"] == 0        # unstructured: every row its own pattern
    n2, bi, bj, ba = stencil7(8, 8, 8)
    assert analyze(n2, n2, bi, bj, ba)["npat"] == 0      # 512 rows: below the size worth a table


def test_mostly_irregular_block_is_rejected_below_70_percent_coverage():
    n, ai, aj, aa = stencil7(16, 16, 8)
    rng = np.random.default_rng(11)
    aa = aa.copy()
    for row in rng.choice(n, size=n // 2, replace=False):
        aa[ai[row]:ai[row + 1]] *= 1.0 + rng.random(ai[row + 1] - ai[row])
    assert analyze(n, n, ai, aj, aa)["npat"] == 0


@pytest.mark.parametrize("kind,dims", [("27pt", (24, 24, 24)), ("laplacian", (32, 32, 32))])
def test_reference_hierarchy_blocks(kind, dims):
    """A_0, P_0, P_0^T and A_1 of hierarchies built by the reference's own BoomerAMGSetup"""
    if not os.path.exists(BRIDGE):
        pytest.skip("oracle/_ref/libref_bridge.so not built (needs /root/reference)")
    import scipy.sparse as sp
    from oracle import refbridge as rb
    rb.load()
    pb = rb.Problem(kind, dims)
    pb.setup_amg(relax_type=18)
    h = pb.hierarchy()
    A0 = h["levels"][0]["A"]
    a = A0.arrays()
    r = analyze(A0.num_rows, A0.num_cols, a["diag_i"], a["diag_j"], a["diag_data"])
    assert 0 < r["npat"] <= 27 and r["nirr"] == 0
    assert_lossless(A0.num_rows, a["diag_i"], a["diag_j"], a["diag_data"], r)
    P0 = h["levels"][0]["P"]
    p = P0.arrays()
    r = analyze(P0.num_rows, P0.num_cols, p["diag_i"], p["diag_j"], p["diag_data"])
    if r["npat"]:                                        # rectangular: base = first column of the row
        assert_lossless(P0.num_rows, p["diag_i"], p["diag_j"], p["diag_data"], r)
        first = np.asarray(p["diag_j"])[np.minimum(np.asarray(p["diag_i"])[:-1], len(p["diag_j"]) - 1)]
        nonempty = np.diff(np.asarray(p["diag_i"])) > 0
        assert np.array_equal(r["base"][nonempty], first[nonempty])
        Pm = sp.csr_matrix((np.asarray(p["diag_data"]), np.asarray(p["diag_j"]), np.asarray(p["diag_i"])),
                           shape=(P0.num_rows, P0.num_cols))
        R = Pm.T.tocsr()
        R.sort_indices()
        if R.shape[0] >= 1024:
            rr = analyze(R.shape[0], R.shape[1], R.indptr, R.indices, R.data)
            if rr["npat"]:
                assert_lossless(R.shape[0], R.indptr, R.indices, R.data, rr)
    if len(h["levels"]) > 1 and h["levels"][1]["A"].num_rows >= 1024:
        A1 = h["levels"][1]["A"]
        a1 = A1.arrays()
        r1 = analyze(A1.num_rows, A1.num_cols, a1["diag_i"], a1["diag_j"], a1["diag_data"])
        if r1["npat"]:
            assert_lossless(A1.num_rows, a1["diag_i"], a1["diag_j"], a1["diag_data"], r1)


# ----------------------------------------------------------------------------------------
# stored transpose (restriction) and hybrid-GS wavefront schedule: host logic against the oracle
# ----------------------------------------------------------------------------------------
def host_transpose(n, ncols, ai, aj, aa):
    from hypre_b200._lib import lib, check
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    aj = np.ascontiguousarray(aj, dtype=np.int32)
    aa = np.ascontiguousarray(aa, dtype=np.float64)
    ti = np.zeros(ncols + 1, np.int32)
    tj = np.zeros(max(len(aj), 1), np.int32)
    ta = np.zeros(max(len(aa), 1), np.float64)
    check(lib.hb200_host_csr_transpose(n, ncols, ai.ctypes.data, aj.ctypes.data, aa.ctypes.data,
                                       ti.ctypes.data, tj.ctypes.data, ta.ctypes.data))
    return ti, tj[:len(aj)], ta[:len(aa)]


def gs_schedule(n, ai, aj, forward):
    from hypre_b200._lib import lib, check
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    aj = np.ascontiguousarray(aj, dtype=np.int32)
    perm = np.zeros(max(n, 1), np.int32)
    lptr = np.zeros(n + 2, np.int32)
    nlev = C.c_int(0)
    check(lib.hb200_host_gs_schedule(n, ai.ctypes.data, aj.ctypes.data, 1 if forward else 0,
                                     perm.ctypes.data, lptr.ctypes.data, C.byref(nlev)))
    return perm[:n], lptr[:nlev.value + 1]


def random_rect(rng, n, ncols, per_row):
    ai = [0]
    aj, aa = [], []
    for _ in range(n):
        k = int(rng.integers(0, per_row + 1))
        cols = np.sort(rng.choice(ncols, size=min(k, ncols), replace=False))
        aj.extend(cols.tolist())
        aa.extend(rng.standard_normal(len(cols)).tolist())
        ai.append(len(aj))
    return np.array(ai, np.int32), np.array(aj, np.int32), np.array(aa)


def test_stored_transpose_sums_in_the_reference_order():
    """row sums of the stored transpose == the oracle's restatement of hypre_CSRMatrixMatvecT
    (csr_matvec.c:1095-1110, one thread), bit for bit: same products, same order"""
    from oracle import restatement as orc
    rng = np.random.default_rng(42)
    for n, ncols in ((300, 90), (64, 64), (5, 200)):
        ai, aj, aa = random_rect(rng, n, ncols, 7)
        ti, tj, ta = host_transpose(n, ncols, ai, aj, aa)
        assert ti[0] == 0 and ti[-1] == len(aj) and np.all(np.diff(ti) >= 0)
        for c in range(ncols):                       # ascending source rows within each output row
            assert np.all(np.diff(tj[ti[c]:ti[c + 1]]) > 0)
        x = rng.standard_normal(n)
        yref = orc.csr_matvecT(ai, aj, aa, ncols, 1.0, x, 0.0, np.zeros(ncols))
        y = np.zeros(ncols)
        for c in range(ncols):
            s = 0.0
            for p in range(ti[c], ti[c + 1]):
                s += ta[p] * x[tj[p]]
            y[c] = s
        assert np.array_equal(y, yref)


@pytest.mark.parametrize("forward", [True, False])
def test_gs_wavefront_schedule_reproduces_the_sequential_sweep(forward):
    """levels are independent sets that respect the sweep's dependencies, and sweeping them in
    order gives exactly the oracle's sequential Gauss-Seidel sweep (relax type 3 forward / 4 backward)"""
    from oracle import restatement as orc
    n, ai, aj, aa = stencil7(9, 8, 7)
    rng = np.random.default_rng(3)
    aa = aa * (1.0 + 0.1 * rng.random(len(aa)))      # generic values, diagonal stays first and dominant
    perm, lptr = gs_schedule(n, ai, aj, forward)
    assert sorted(perm.tolist()) == list(range(n))
    lev = np.empty(n, np.int64)
    for l in range(len(lptr) - 1):
        lev[perm[lptr[l]:lptr[l + 1]]] = l
    for i in range(n):
        for p in range(ai[i], ai[i + 1]):
            j = int(aj[p])
            if j == i:
                continue
            earlier = (j < i) if forward else (j > i)
            # a coupled row that the sequential sweep visits earlier must sit in an earlier level,
            # one it visits later in a later level: never the same level
            assert (lev[j] < lev[i]) if earlier else (lev[j] > lev[i]), (i, j)
    f = rng.standard_normal(n)
    u0 = rng.standard_normal(n)
    uref = orc.relax(ai, aj, aa, f, u0.copy(), 3 if forward else 4)
    u = u0.copy()
    for l in range(len(lptr) - 1):
        rows = perm[lptr[l]:lptr[l + 1]]
        new = {}
        for i in rows:                                # all rows of a level read the same state
            s = f[i]
            for p in range(ai[i] + 1, ai[i + 1]):
                s -= aa[p] * u[aj[p]]
            new[i] = s / aa[ai[i]]
        for i, v in new.items():
            u[i] = v
    assert np.array_equal(u, uref)


def test_wide_variant_takes_blocks_the_one_byte_format_rejects():
    """half of the rows perturbed in groups of 5 equal rows: ~400 patterns of >= 4 rows, too many for
    1-byte codes at < 70 % coverage, fine for 16-bit codes; one-off rows go to the irregular list"""
    n, ai, aj, aa = stencil7(16, 16, 16)
    rng = np.random.default_rng(17)
    aa = aa.copy()
    interior = [r for r in range(n) if ai[r + 1] - ai[r] == 7]
    pick = rng.choice(interior, size=2000, replace=False)
    for g in range(0, 2000, 5):                       # groups of 5 rows share one new pattern
        scale = 1.0 + rng.random(7)
        for r in pick[g:g + 5]:
            aa[ai[r]:ai[r + 1]] *= scale
    singles = [r for r in interior if r not in set(pick.tolist())][:50]
    for r in singles:                                 # and 50 one-off rows
        aa[ai[r]:ai[r + 1]] *= 1.0 + rng.random(7)
    assert analyze(n, n, ai, aj, aa)["npat"] == 0
    w = analyze_wide(n, n, ai, aj, aa)
    assert w["npat"] >= 400 and set(singles) <= set(w["irr"].tolist())
    assert_lossless(n, ai, aj, aa, w)


def _same_analysis(a, b):
    assert a["npat"] == b["npat"] and a["nirr"] == b["nirr"]
    for k in ("code", "base", "ptr", "irr"):
        assert np.array_equal(a[k], b[k]), k
    ne = int(a["ptr"][-1]) if a["npat"] else 0
    assert np.array_equal(a["off"][:ne], b["off"][:ne])
    assert np.array_equal(a["val"][:ne].view(np.uint64), b["val"][:ne].view(np.uint64))


def test_threaded_classification_is_the_sequential_one():
    """blocks of >= 2^18 rows are classified by several host threads (chunks of rows, candidate lists merged in chunk
    order): the row codes, the table and the irregular-row list are those of the one-thread pass — on a regular
    stencil, on one with scattered one-off rows, on a rectangular block and on an irregular block (left to the
    sequential pass)"""
    import scipy.sparse as sp
    n, ai, aj, aa = stencil7(64, 64, 64)
    assert n >= 1 << 18
    rng = np.random.default_rng(9)
    cases = [("regular", n, n, ai, aj, aa.copy())]
    pert = aa.copy()
    rows = rng.choice(n, 3000, replace=False)
    pert[ai[rows]] += rng.standard_normal(3000)            # 3000 rows with a one-off diagonal
    cases.append(("one-off rows", n, n, ai, aj, pert))
    # rectangular: every second column dropped from the column space (an interpolation-like block)
    keep = (aj % 2 == 0)
    rowsz = np.add.reduceat(keep.astype(np.int64), ai[:-1])
    ai2 = np.concatenate([[0], np.cumsum(rowsz)]).astype(np.int32)
    cases.append(("rectangular", n, n // 2, ai2, (aj[keep] // 2).astype(np.int32), aa[keep].copy()))
    noisy = aa * (1.0 + 0.1 * rng.standard_normal(aa.size))   # every row its own values: not this format's business
    cases.append(("irregular", n, n, ai, aj, noisy))
    old = os.environ.get("HB200_ANALYSIS_THREADS")
    try:
        for name, nr, nc, i_, j_, a_ in cases:
            os.environ["HB200_ANALYSIS_THREADS"] = "1"
            seq = analyze(nr, nc, i_, j_, a_)
            os.environ["HB200_ANALYSIS_THREADS"] = "7"
            par = analyze(nr, nc, i_, j_, a_)
            _same_analysis(seq, par)
            if name == "regular":
                assert seq["npat"] == 27 and seq["nirr"] == 0
            if name == "one-off rows":
                assert seq["npat"] > 0 and seq["nirr"] >= 3000
            if name == "irregular":
                assert seq["npat"] == 0
    finally:
        if old is None:
            os.environ.pop("HB200_ANALYSIS_THREADS", None)
        else:
            os.environ["HB200_ANALYSIS_THREADS"] = old

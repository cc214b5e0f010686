"""IJ interface and on-disk formats.

CPU part (no GPU): the host assembly of coordinate triplets — insertion order inside a row, repeated entries,
the diagonal moved to the front, col_map_offd — against the ParCSR arrays the reference's own HYPRE_IJMatrixRead
assembles from the same file, bit for bit.  The N-rank part (offd block, CommPkg) runs in tests/mp_cpu_worker.py.
GPU part (`-m gpu`; the CPU suite runs it on the host emulation): read / print round trips against the reference's
files and the SpMV of a matrix that entered through hb200_parcsr_read_ij."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BRIDGE = os.path.join(ROOT, "oracle", "_ref", "libref_bridge.so")


def read_ij_text(path):
    with open(path) as f:
        head = [int(t) for t in f.readline().split()]
        rows, cols, vals = [], [], []
        for line in f:
            t = line.split()
            rows.append(int(t[0])); cols.append(int(t[1])); vals.append(float(t[2]))
    return head, np.array(rows, np.int64), np.array(cols, np.int64), np.array(vals, np.float64)


def write_ij_text(path, head, rows, cols, vals):
    with open(path, "w") as f:
        f.write("%d %d %d %d\n" % tuple(head))
        for i, j, v in zip(rows, cols, vals):
            f.write("%d %d %.14e\n" % (i, j, v))


def host_assemble(head, rows, cols, vals, add=False):
    from hypre_b200._lib import lib, check
    r = np.ascontiguousarray(rows, np.int64); c = np.ascontiguousarray(cols, np.int64); v = np.ascontiguousarray(vals, np.float64)
    dn, on, nco = C.c_int(0), C.c_int(0), C.c_int(0)
    args = (head[0], head[1], head[2], head[3], len(r), r.ctypes.data, c.ctypes.data, v.ctypes.data, 1 if add else 0)
    check(lib.hb200_host_ij_assemble(*args, C.byref(dn), C.byref(on), C.byref(nco), *([None] * 7)))
    n = head[1] - head[0] + 1
    out = {"diag_i": np.zeros(n + 1, np.int32), "diag_j": np.zeros(dn.value, np.int32), "diag_data": np.zeros(dn.value),
           "offd_i": np.zeros(n + 1, np.int32), "offd_j": np.zeros(on.value, np.int32), "offd_data": np.zeros(on.value),
           "col_map_offd": np.zeros(nco.value, np.int64)}
    check(lib.hb200_host_ij_assemble(*args, C.byref(dn), C.byref(on), C.byref(nco),
                                     *[out[k].ctypes.data for k in ("diag_i", "diag_j", "diag_data", "offd_i", "offd_j", "offd_data", "col_map_offd")]))
    return out


@pytest.fixture(scope="module")
def rb():
    from oracle import refbridge
    if not os.path.exists(BRIDGE):
        pytest.skip("oracle/_ref/libref_bridge.so not built (needs /root/reference)")
    refbridge.load()
    refbridge.set_num_threads(1)
    return refbridge


@pytest.mark.parametrize("kind,n", [("27pt", (6, 5, 4)), ("laplacian", (7, 3, 5)), ("vardifconv", (5, 5, 5))])
def test_host_assembly_matches_the_reference_reader(rb, tmp_path, kind, n):
    pb = rb.Problem(kind, n)
    name = str(tmp_path / "A")
    pb.print_ij(name)
    head, rows, cols, vals = read_ij_text(name + ".00000")
    # the file in another order, with repeated entries (HYPRE_IJMatrixRead sets: the last value of an (i, j) wins)
    rng = np.random.default_rng(3)
    perm = rng.permutation(len(rows))
    extra = rng.integers(0, len(rows), 25)
    r2 = np.concatenate([rows[perm], rows[extra]]); c2 = np.concatenate([cols[perm], cols[extra]])
    v2 = np.concatenate([vals[perm], 10.0 + np.arange(25.0)])
    name2 = str(tmp_path / "B")
    write_ij_text(name2 + ".00000", head, r2, c2, v2)
    pb2 = rb.Problem.from_ij_file(name2)               # (keeps the arrays of the view alive)
    ref = pb2.level_view(0, 0).arrays()
    mine = host_assemble(head, r2, c2, v2)
    for k in ("diag_i", "diag_j", "diag_data"):
        assert np.array_equal(mine[k], ref[k]), k
    assert mine["col_map_offd"].size == 0 and mine["offd_j"].size == 0
    # every row starts with its diagonal entry (par_relax.c:274)
    di, dj = mine["diag_i"], mine["diag_j"]
    assert all(dj[di[i]] == i for i in range(len(di) - 1))
    # AddToValues semantics: repeated entries are summed
    added = host_assemble(head, r2, c2, v2, add=True)
    assert np.array_equal(added["diag_j"], mine["diag_j"])
    assert abs(added["diag_data"].sum() - v2.sum()) <= 1e-9 * np.abs(v2).sum()


def test_host_assembly_splits_off_range_columns():
    # rows 4..7 of an 12 x 12 matrix whose columns 4..7 are "mine": the others go to the offd block through an
    # ascending col_map_offd, in insertion order
    head = (4, 7, 4, 7)
    rows = np.array([4, 4, 4, 5, 5, 6, 7, 7, 7], np.int64)
    cols = np.array([9, 4, 1, 5, 11, 6, 0, 7, 9], np.int64)
    vals = np.arange(1.0, 10.0)
    m = host_assemble(head, rows, cols, vals)
    assert m["col_map_offd"].tolist() == [0, 1, 9, 11]
    assert m["diag_i"].tolist() == [0, 1, 2, 3, 4] and m["diag_j"].tolist() == [0, 1, 2, 3]
    assert m["offd_i"].tolist() == [0, 2, 3, 3, 5] and m["offd_j"].tolist() == [2, 1, 3, 0, 2]
    assert m["offd_data"].tolist() == [1.0, 3.0, 5.0, 7.0, 9.0]


def test_host_commpkg_of_a_three_rank_chain():
    from hypre_b200._lib import lib, check
    # three ranks own columns 0..3, 4..7, 8..11; rank 1 needs {2, 3, 8}, rank 0 needs {4}, rank 2 needs {3, 7}
    own = np.array([0, 3, 0, 3, 1, 4, 7, 4, 7, 3, 8, 11, 8, 11, 2], np.int64)
    need = np.array([4, -1, -1, 2, 3, 8, 3, 7, -1], np.int64)
    got = {}
    for me in range(3):
        ns, nr = C.c_int(0), C.c_int(0)
        sp, sms, sme = np.zeros(3, np.int32), np.zeros(4, np.int32), np.zeros(16, np.int32)
        rp, rvs = np.zeros(3, np.int32), np.zeros(4, np.int32)
        check(lib.hb200_host_ij_commpkg(3, me, own.ctypes.data, need.ctypes.data, 3, C.byref(ns), sp.ctypes.data, sms.ctypes.data,
                                        sme.ctypes.data, 16, C.byref(nr), rp.ctypes.data, rvs.ctypes.data))
        got[me] = (sp[:ns.value].tolist(), sms[:ns.value + 1].tolist(), sme[:sms[ns.value]].tolist(),
                   rp[:nr.value].tolist(), rvs[:nr.value + 1].tolist())
    assert got[0] == ([1, 2], [0, 2, 3], [2, 3, 3], [1], [0, 1])
    assert got[1] == ([0, 2], [0, 1, 2], [0, 3], [0, 2], [0, 2, 3])
    assert got[2] == ([1], [0, 1], [0], [0, 1], [0, 1, 2])


# ----------------------------------------------------------------------------------------------------------
# device part
# ----------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def hb():
    import hypre_b200 as h
    h.init(0)
    return h


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n", [("27pt", (9, 8, 7)), ("vardifconv", (8, 8, 8))])
def test_read_ij_matvec_and_print_round_trip(rb, hb, tmp_path, kind, n):
    import torch
    pb = rb.Problem(kind, n)
    name = str(tmp_path / "A")
    pb.print_ij(name)
    A = hb.ParCSRMatrix.read_ij(name)
    ref = pb.level_view(0, 0).arrays()
    maps = A.download_maps()
    assert np.array_equal(maps["diag_i"], ref["diag_i"]) and np.array_equal(maps["diag_j"], ref["diag_j"])
    rng = np.random.default_rng(11)
    x = rng.standard_normal(A.num_cols)
    yref = pb.matvec(1.0, x, 0.0)
    y = torch.zeros(A.num_rows, dtype=torch.float64, device="cuda")
    A.matvec(1.0, torch.from_numpy(x).cuda(), 0.0, y)
    # (the file holds 15 significant digits: exact for the stencil operators, 1e-14 for the variable coefficients;
    # a block this small runs the CSR kernel, whose lanes add in another order)
    err = float(np.max(np.abs(y.cpu().numpy() - yref)) / np.max(np.abs(yref)))
    assert err <= 1e-12, err
    # the writer produces the reference's text
    out = str(tmp_path / "B")
    A.print_ij(out)
    if kind == "27pt":
        assert open(out + ".00000").read() == open(name + ".00000").read()
    # ... and the reference reads it back to the same matrix
    pb_back = rb.Problem.from_ij_file(out)
    back = pb_back.level_view(0, 0).arrays()
    assert np.array_equal(back["diag_i"], ref["diag_i"]) and np.array_equal(back["diag_j"], ref["diag_j"])
    assert np.max(np.abs(back["diag_data"] - ref["diag_data"])) <= 1e-13 * np.max(np.abs(ref["diag_data"]))


@pytest.mark.gpu
def test_binary_ij_files_are_lossless_both_ways(rb, hb, tmp_path):
    pb = rb.Problem("vardifconv", (7, 6, 5))
    ref = pb.level_view(0, 0).arrays()
    name = str(tmp_path / "A")
    pb.print_ij(name, binary=True)                     # written by the reference (HYPRE_IJMatrixPrintBinary)
    A = hb.ParCSRMatrix.read_ij(name, binary=True)
    maps = A.download_maps()
    assert np.array_equal(maps["diag_i"], ref["diag_i"]) and np.array_equal(maps["diag_j"], ref["diag_j"])
    out = str(tmp_path / "B")
    A.print_ij(out, binary=True)
    assert open(out + ".00000.bin", "rb").read() == open(name + ".00000.bin", "rb").read()
    pb_back = rb.Problem.from_ij_file(out, binary=True)    # ... and read back by the reference: the same bits
    back = pb_back.level_view(0, 0).arrays()
    for k in ("diag_i", "diag_j", "diag_data"):
        assert np.array_equal(back[k], ref[k]), k


@pytest.mark.gpu
def test_from_ij_solves_like_the_reference(rb, hb, tmp_path):
    """a matrix that entered as triplets (shuffled, with duplicates to add up) drives the same diag-scaled PCG"""
    import torch
    pb = rb.Problem("laplacian", (8, 7, 6))
    name = str(tmp_path / "A")
    pb.print_ij(name)
    head, rows, cols, vals = read_ij_text(name + ".00000")
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(rows))
    # every entry split into two halves that AddToValues sums again (exact for these values)
    r2 = np.concatenate([rows[perm], rows]); c2 = np.concatenate([cols[perm], cols]); v2 = np.concatenate([0.5 * vals[perm], 0.5 * vals])
    A = hb.ParCSRMatrix.from_ij(*head, r2, c2, v2, add_duplicates=True)
    assert A.num_rows == pb.local_rows and A.diag_nnz == len(rows)
    ref = pb.pcg(precond="diagscale", tol=1e-8, max_iter=200, two_norm=1)
    pcg = hb.ParCSRPCG(tol=1e-8, max_iter=200, two_norm=1)
    pcg.set_precond("diagscale")
    x = torch.zeros(A.num_rows, dtype=torch.float64, device="cuda")
    pcg.solve(A, torch.from_numpy(np.array(pb.b)).cuda(), x)
    assert abs(pcg.num_iterations - ref["iterations"]) <= 1
    assert np.max(np.abs(x.cpu().numpy() - ref["x"])) <= 1e-7 * np.max(np.abs(ref["x"]))


@pytest.mark.gpu
def test_vector_files_and_matrix_market(rb, hb, tmp_path):
    import torch
    pb = rb.Problem("laplacian", (5, 4, 3))
    rng = np.random.default_rng(2)
    v = rng.standard_normal(pb.local_rows)
    name = str(tmp_path / "v")
    pb.print_vector_ij(v, name)                       # written by the reference
    lo, x = hb.vector_read_ij(name)
    assert lo == 0 and np.max(np.abs(x.cpu().numpy() - v)) <= 1e-14 * np.max(np.abs(v))
    out = str(tmp_path / "w")
    hb.vector_print_ij(x, 0, out)
    assert open(out + ".00000").read() == open(name + ".00000").read()
    # Matrix Market, symmetric storage: the reference's reader (HYPRE_IJMatrixReadMM) and ours build the same matrix
    mm = str(tmp_path / "S.mtx")
    with open(mm, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real symmetric\n% a comment\n4 4 6\n1 1 4.0\n2 1 -1.0\n2 2 4.0\n3 2 -1.5\n3 3 5.0\n4 4 2.0\n")
    A = hb.ParCSRMatrix.read_ij(mm, matrix_market=True)
    ref = rb.Problem.from_ij_file(mm, matrix_market=True)
    ra = ref.level_view(0, 0).arrays()
    maps = A.download_maps()
    assert np.array_equal(maps["diag_i"], ra["diag_i"]) and np.array_equal(maps["diag_j"], ra["diag_j"])
    xs = np.array([1.0, 2.0, 3.0, 4.0])
    y = torch.zeros(4, dtype=torch.float64, device="cuda")
    A.matvec(1.0, torch.from_numpy(xs).cuda(), 0.0, y)
    assert np.max(np.abs(y.cpu().numpy() - ref.matvec(1.0, xs, 0.0))) <= 1e-14


_HIER_CASES = [("27pt", dict(relax_type=18)), ("laplacian", dict(relax_type=16)), ("vardifconv", dict(relax_type=8, relax_order=1))]
if os.environ.get("HB200_EMU_TEST") == "1":
    _HIER_CASES = _HIER_CASES[2:]         # (the CPU suite's emulation run takes one of them)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,smoother", _HIER_CASES)
def test_hierarchy_saved_and_loaded_solves_identically(rb, hb, tmp_path, kind, smoother):
    """hb200_amg_save / hb200_amg_load ("ship hierarchies between boxes"): the loaded hierarchy — level matrices from
    hypre's binary IJ files, everything else from the record — runs the same PCG / GMRES solve, bit for bit, as the
    hierarchy uploaded straight from the reference's setup; the level files are readable by hypre itself"""
    import torch
    pb = rb.Problem(kind, (11, 10, 9))
    pb.setup_amg(**smoother)
    mats, amg = hb.amg_from_hierarchy(pb.hierarchy())
    d = str(tmp_path / "hier")
    amg.save(d)
    loaded = hb.BoomerAMG.load(d)
    assert len(loaded.mats) == len(mats)
    for (A, P), (A2, P2) in zip(mats, loaded.mats):
        m1, m2 = A.download_maps(), A2.download_maps()
        assert np.array_equal(m1["diag_i"], m2["diag_i"]) and np.array_equal(m1["diag_j"], m2["diag_j"])
        assert (P is None) == (P2 is None)
        if P is not None:
            p1, p2 = P.download_maps(), P2.download_maps()
            assert np.array_equal(p1["diag_i"], p2["diag_i"]) and np.array_equal(p1["diag_j"], p2["diag_j"])
    b = torch.from_numpy(np.array(pb.b)).cuda()
    xs = []
    for M, pc in ((mats[0][0], amg), (loaded.mats[0][0], loaded)):
        ks = hb.ParCSRGMRES(tol=1e-8, max_iter=60, k_dim=5) if kind == "vardifconv" else hb.ParCSRPCG(tol=1e-8, max_iter=60, two_norm=1)
        ks.set_precond(pc)
        x = torch.zeros(M.num_rows, dtype=torch.float64, device="cuda")
        ks.solve(M, b, x)
        xs.append((ks.num_iterations, x.cpu().numpy()))
    assert xs[0][0] == xs[1][0] and np.array_equal(xs[0][1], xs[1][1])
    # hypre reads the fine-level file of the saved hierarchy back to the reference's own operator
    back = rb.Problem.from_ij_file(os.path.join(d, "A0"), binary=True)
    ra, rr = back.level_view(0, 0).arrays(), pb.level_view(0, 0).arrays()
    for k in ("diag_i", "diag_j", "diag_data"):
        assert np.array_equal(ra[k], rr[k]), k

"""CPU multi-rank worker (gloo): validates the host-side logic of the N>1 path without a GPU.

 * the reference runs on `world` ranks over oracle/minimpi and reproduces the golden output its
   own regression suite pins (src/test/TEST_ij/solvers.saved, out.0: 7 its, 3.095059e-09);
 * the hierarchy view that gets uploaded to the GPUs is self-consistent across ranks: what rank r
   packs for rank q (send_map_elmts -> global ids) is exactly, element for element, the segment
   of q's col_map_offd that q expects from r — i.e. the halo plan the NCCL path executes is right;
 * executing that plan with plain gloo send/recv yields x_ext == x_global[col_map_offd]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    from oracle import refbridge as rb
    rb.load(mpi=True)
    rb.set_num_threads(1)

    # ---- golden: mpirun -np 2 ./ij -solver 1 -rhsrand  (default 10^3 7-pt Laplacian, P = 1 x np x 1)
    if world == 2:
        pb = rb.Problem("laplacian", (10, 10, 10), P=(1, 2, 1), rhs="rand", mpi=True)
        pb.setup_amg()                      # ij defaults: hybrid l1-GS 13/14
        r = pb.pcg(precond="amg", tol=1e-8, max_iter=100, two_norm=1)
        assert r["iterations"] == 7, r["iterations"]
        assert abs(r["final_rel_res"] - 3.095059e-09) < 5e-16 * 1e7, r["final_rel_res"]
        g = pb.gmres(precond="amg", tol=1e-8, max_iter=100, k_dim=5)
        assert g["iterations"] == 7 and abs(g["final_rel_res"] - 4.842561e-09) < 5e-16 * 1e7, g
        pb.destroy()

    # ---- halo plan consistency on a 27-pt problem, every level, A and P
    P = {2: (2, 1, 1), 3: (3, 1, 1), 4: (2, 2, 1)}[world]
    pb = rb.Problem("27pt", (7 * P[0], 6 * P[1], 5 * P[2]), P=P, mpi=True)
    pb.setup_amg(relax_type=18)
    h = pb.hierarchy()
    nl = pb.num_levels
    rng = np.random.default_rng(7 + rank)
    for l in range(nl):
        for which in (0, 1):
            v = h["levels"][l]["A" if which == 0 else "P"]
            if v is None:
                continue
            a = v.arrays()
            ns, nr = v.num_sends, v.num_recvs
            first_col = v.first_col
            # what I send, as global column ids, per destination
            my_sends = {}
            for s in range(ns):
                lo, hi = a["send_map_starts"][s], a["send_map_starts"][s + 1]
                my_sends[int(a["send_procs"][s])] = (a["send_map_elmts"][lo:hi].astype(np.int64) + first_col)
            my_recvs = {}
            for q in range(nr):
                lo, hi = a["recv_vec_starts"][q], a["recv_vec_starts"][q + 1]
                my_recvs[int(a["recv_procs"][q])] = np.array(a["col_map_offd"][lo:hi], dtype=np.int64)
            all_sends = [None] * world
            dist.all_gather_object(all_sends, my_sends)
            for src, want in my_recvs.items():
                got = all_sends[src].get(rank)
                assert got is not None and np.array_equal(got, want), (l, which, src, rank)
            # and nobody sends me something I do not expect
            for src in range(world):
                if rank in all_sends[src]:
                    assert src in my_recvs, (l, which, src, rank)
            # execute the plan with gloo: x_ext must equal x_global[col_map_offd]
            ncols = v.num_cols
            x = rng.standard_normal(ncols)
            gathered = [None] * world
            dist.all_gather_object(gathered, (int(first_col), x))
            xg = {}
            for fc, xx in gathered:
                for k, val in enumerate(xx):
                    xg[fc + k] = val
            reqs, bufs = [], {}
            for q in range(nr):
                src = int(a["recv_procs"][q])
                cnt = int(a["recv_vec_starts"][q + 1] - a["recv_vec_starts"][q])
                bufs[q] = torch.zeros(cnt, dtype=torch.float64)
                reqs.append(dist.irecv(bufs[q], src=src, tag=l * 2 + which))
            for s in range(ns):
                lo, hi = a["send_map_starts"][s], a["send_map_starts"][s + 1]
                packed = torch.from_numpy(x[a["send_map_elmts"][lo:hi]].copy())
                reqs.append(dist.isend(packed, dst=int(a["send_procs"][s]), tag=l * 2 + which))
            for rq in reqs:
                rq.wait()
            if v.num_cols_offd:
                x_ext = np.concatenate([bufs[q].numpy() for q in range(nr)]) if nr else np.zeros(0)
                want = np.array([xg[int(c)] for c in a["col_map_offd"]])
                assert np.array_equal(x_ext, want), (l, which)
    # ---- IJ interface, host halves: this rank's rows of A_0 as coordinate triplets through
    # hb200_host_ij_assemble give the reference's diag / offd blocks and col_map_offd bit for bit (A_0 comes from
    # the generator: diagonal first, then ascending columns — the order the triplets are handed over in); the
    # CommPkg built from the all-gathered ownership and need lists is the reference's (hypre_MatvecCommPkgCreate)
    import ctypes as C
    from hypre_b200._lib import lib, check
    v = h["levels"][0]["A"]
    a = v.arrays()
    n = v.num_rows
    rows, cols, vals = [], [], []
    for i in range(n):
        for p in range(a["diag_i"][i], a["diag_i"][i + 1]):
            rows.append(v.first_row + i); cols.append(v.first_col + int(a["diag_j"][p])); vals.append(a["diag_data"][p])
        if v.num_cols_offd:
            for p in range(a["offd_i"][i], a["offd_i"][i + 1]):
                rows.append(v.first_row + i); cols.append(int(a["col_map_offd"][a["offd_j"][p]])); vals.append(a["offd_data"][p])
    r = np.array(rows, np.int64); c = np.array(cols, np.int64); w = np.array(vals, np.float64)
    head = (v.first_row, v.first_row + n - 1, v.first_col, v.first_col + v.num_cols - 1)
    dn, on, nco = C.c_int(0), C.c_int(0), C.c_int(0)
    args = (*head, len(r), r.ctypes.data, c.ctypes.data, w.ctypes.data, 0)
    check(lib.hb200_host_ij_assemble(*args, C.byref(dn), C.byref(on), C.byref(nco), *([None] * 7)))
    assert dn.value == v.diag_nnz and on.value == v.offd_nnz and nco.value == v.num_cols_offd
    out = {"diag_i": np.zeros(n + 1, np.int32), "diag_j": np.zeros(dn.value, np.int32), "diag_data": np.zeros(dn.value),
           "offd_i": np.zeros(n + 1, np.int32), "offd_j": np.zeros(on.value, np.int32), "offd_data": np.zeros(on.value),
           "col_map_offd": np.zeros(nco.value, np.int64)}
    keys = ("diag_i", "diag_j", "diag_data", "offd_i", "offd_j", "offd_data", "col_map_offd")
    check(lib.hb200_host_ij_assemble(*args, C.byref(dn), C.byref(on), C.byref(nco), *[out[k].ctypes.data for k in keys]))
    for k in keys:
        if a[k] is not None:
            assert np.array_equal(out[k], a[k]), ("ij assemble", k, rank)
    own_all = [None] * world
    dist.all_gather_object(own_all, (list(head) + [nco.value], out["col_map_offd"].tolist()))
    max_offd = max(o[0][4] for o in own_all)
    own5 = np.array([x for o in own_all for x in o[0]], np.int64)
    need = np.full(world * max(max_offd, 1), -1, np.int64)
    for q, o in enumerate(own_all):
        need[q * max_offd: q * max_offd + len(o[1])] = o[1]
    ns, nr = C.c_int(0), C.c_int(0)
    cap = max(int(a["send_map_starts"][v.num_sends]) if v.num_sends else 0, 1) + 8
    sp, sms, sme = np.zeros(world, np.int32), np.zeros(world + 1, np.int32), np.zeros(cap, np.int32)
    rp, rvs = np.zeros(world, np.int32), np.zeros(world + 1, np.int32)
    check(lib.hb200_host_ij_commpkg(world, rank, own5.ctypes.data, need.ctypes.data, max_offd, C.byref(ns), sp.ctypes.data,
                                    sms.ctypes.data, sme.ctypes.data, cap, C.byref(nr), rp.ctypes.data, rvs.ctypes.data))
    assert ns.value == v.num_sends and nr.value == v.num_recvs, ("ij commpkg", ns.value, v.num_sends, nr.value, v.num_recvs)
    if v.num_sends:
        assert np.array_equal(sp[:ns.value], a["send_procs"]) and np.array_equal(sms[:ns.value + 1], a["send_map_starts"])
        assert np.array_equal(sme[:sms[ns.value]], a["send_map_elmts"]), ("ij commpkg send_map_elmts", rank)
    if v.num_recvs:
        assert np.array_equal(rp[:nr.value], a["recv_procs"]) and np.array_equal(rvs[:nr.value + 1], a["recv_vec_starts"])
    dist.barrier()
    if rank == 0:
        print("CPU MULTI-RANK OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

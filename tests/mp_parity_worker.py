"""Multi-rank parity worker (one process per GPU, launched by torch.distributed.run).

Every rank: builds its part of an ij problem with the reference running on minimpi, runs the
reference's BoomerAMGSetup, uploads its part of the hierarchy through the hb200 C-ABI, then checks
the NCCL path against the reference's MPI path on the same inputs: bit-exact maps, SpMV / SpMV-T
/ relax / cycle to 1e-12-ish, PCG / GMRES iteration counts and residual histories."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def relerr(a, b, den=None):
    d = np.max(np.abs(b)) if den is None else den
    return float(np.max(np.abs(a - b)) / (d if d > 0 else 1.0)) if a.size else 0.0


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    kind = sys.argv[1] if len(sys.argv) > 1 else "27pt"
    halo = sys.argv[2] if len(sys.argv) > 2 else "nccl"
    if os.environ.get("HB200_EMU_TEST"):
        # CPU run against the host emulation of the kernels (tests/test_emu_kernels.py): one host
        # process per rank, NCCL calls over oracle/minimpi
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import emu_env
        emu_env.activate()
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import hypre_b200 as hb
    from hypre_b200._lib import lib, check
    from oracle import refbridge as rb
    hb.init(local)
    uid = [hb.comm_get_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    hb.comm_init(rank, world, uid[0])
    if halo == "peer":
        check(lib.hb200_set_halo_mode(1))
    rb.load(mpi=True)
    rb.set_num_threads(1)
    P = {1: (1, 1, 1), 2: (2, 1, 1), 3: (3, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    scale = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    n = (12 * scale * P[0], 11 * scale * P[1], 10 * scale * P[2])
    pb = rb.Problem(kind, n, P=P, mpi=True)
    setup_relax = int(sys.argv[4]) if len(sys.argv) > 4 else 18      # -1 = ij's default (hybrid l1-GS 13/14)
    pb.setup_amg(relax_type=setup_relax)
    h = pb.hierarchy()
    mats, amg = hb.amg_from_hierarchy(h)
    nl = pb.num_levels

    def gmax(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()

    fails = []

    def expect(cond, what):
        ok = gmax(0.0 if cond else 1.0) == 0.0
        if not ok:
            fails.append(what)
        if rank == 0:
            print(("ok   " if ok else "FAIL ") + what, flush=True)

    rng = np.random.default_rng(100 + rank)
    # ---- maps: bit-exact round trip
    same = True
    for l, (A, Pm) in enumerate(mats):
        for M, view in ((A, h["levels"][l]["A"]), (Pm, h["levels"][l]["P"])):
            if M is None:
                continue
            ref = view.arrays()
            got = M.download_maps()
            for key in ("diag_i", "diag_j", "offd_i", "offd_j", "col_map_offd", "send_map_starts",
                        "send_map_elmts", "recv_vec_starts", "send_procs", "recv_procs"):
                r_, g_ = ref.get(key), got.get(key)
                if r_ is None or g_ is None:
                    continue
                same = same and np.array_equal(np.asarray(r_), np.asarray(g_)[: len(r_)])
    expect(same, "CommPkg / col_map_offd / CSR index arrays bit-exact after device round trip")

    # ---- IJ interface: the reference prints its fine-level operator as IJ files, every rank reads its
    # part back through hb200_parcsr_read_ij (host assembly, CommPkg from two NCCL all-gathers): same maps, same SpMV
    import tempfile
    tmp = [tempfile.mkdtemp(prefix="hb200_ij_") if rank == 0 else None]
    dist.broadcast_object_list(tmp, src=0)
    name = os.path.join(tmp[0], "A")
    pb.print_ij(name)
    dist.barrier()
    Aij = hb.ParCSRMatrix.read_ij(name)
    ref0 = h["levels"][0]["A"].arrays()
    got0 = Aij.download_maps()
    same = Aij.num_rows == mats[0][0].num_rows and Aij.num_cols_offd == mats[0][0].num_cols_offd
    for key in ("diag_i", "diag_j", "offd_i", "offd_j", "col_map_offd", "send_map_starts", "send_map_elmts",
                "recv_vec_starts", "send_procs", "recv_procs"):
        r_, g_ = ref0.get(key), got0.get(key)
        if r_ is None and g_ is None:
            continue
        same = same and r_ is not None and g_ is not None and np.array_equal(np.asarray(r_), np.asarray(g_)[: len(r_)])
    expect(same, "IJ files -> ParCSR: CSR blocks, col_map_offd and CommPkg are the reference's, bit for bit")
    x = rng.standard_normal(Aij.num_cols)
    yref = pb.matvec(1.0, x, 0.0, level=0, which=0)
    y = torch.empty(Aij.num_rows, dtype=torch.float64, device="cuda")
    Aij.matvec(1.0, dev(x), 0.0, y)
    den = gmax(float(np.max(np.abs(yref))) if yref.size else 0.0)
    e_ij = gmax(relerr(y.cpu().numpy(), yref, den))
    expect(e_ij <= 1e-12, f"ParCSR matvec of the matrix read from IJ files ({e_ij:.2e})")
    # ---- off-processor contributions: the first rows of every rank keep half of their values in the rank's own part
    # file, the other half sits in the NEXT rank's file (rows outside that part's range: HYPRE_IJMatrixRead adds such
    # entries at the owner, hypre_IJMatrixAssembleOffProcValsParCSR).  Reference and hb200 read the same files; halves
    # add up exactly, so both must give the original operator again
    if world > 1:
        with open(name + ".%05d" % rank) as fh:
            head = fh.readline()
            lines = fh.readlines()
        lo = int(head.split()[0])
        mine, give = [], []
        for ln in lines:
            i_, j_, v_ = ln.split()
            if int(i_) < lo + 3:
                half = "%d %d %.14e\n" % (int(i_), int(j_), 0.5 * float(v_))
                mine.append(half); give.append(half)
            else:
                mine.append(ln)
        given = [None] * world
        dist.all_gather_object(given, give)
        name2 = os.path.join(tmp[0], "B")
        with open(name2 + ".%05d" % rank, "w") as fh:
            fh.write(head)
            fh.writelines(mine)
            fh.writelines(given[(rank - 1) % world])
        dist.barrier()
        pb2 = rb.Problem.from_ij_file(name2, mpi=True)
        ref2 = pb2.level_view(0, 0).arrays()
        Aoff = hb.ParCSRMatrix.read_ij(name2)
        got2 = Aoff.download_maps()
        same = True
        for key in ("diag_i", "diag_j", "offd_i", "offd_j", "col_map_offd", "send_map_starts", "send_map_elmts",
                    "recv_vec_starts", "send_procs", "recv_procs"):
            for r_ in (ref0.get(key), ref2.get(key)):
                g_ = got2.get(key)
                if r_ is None and g_ is None:
                    continue
                same = same and r_ is not None and g_ is not None and np.array_equal(np.asarray(r_), np.asarray(g_)[: len(r_)])
        expect(same, "IJ files with off-processor entries: the reference's ParCSR (and the original operator's), bit for bit")
        y2 = torch.empty(Aoff.num_rows, dtype=torch.float64, device="cuda")
        Aoff.matvec(1.0, dev(x), 0.0, y2)
        expect(bool(np.array_equal(y2.cpu().numpy(), y.cpu().numpy())), "ParCSR matvec after off-processor assembly is bit-identical")
        pb2.destroy()

    # ---- the hierarchy through files (hb200_amg_save / hb200_amg_load, one set per rank): the loaded hierarchy holds the
    # same level matrices and CommPkgs and its V-cycle gives the same vector, bit for bit
    hdir = os.path.join(tmp[0], "hier")
    amg.save(hdir)
    dist.barrier()
    loaded = hb.BoomerAMG.load(hdir)
    same = len(loaded.mats) == len(mats)
    for (A1, P1), (A2, P2) in zip(mats, loaded.mats):
        for M1, M2 in ((A1, A2), (P1, P2)):
            if M1 is None or M2 is None:
                same = same and M1 is None and M2 is None
                continue
            g1, g2 = M1.download_maps(), M2.download_maps()
            for key in g1:
                if g1[key] is None or g2[key] is None:
                    same = same and g1[key] is None and g2[key] is None
                else:
                    same = same and np.array_equal(g1[key], g2[key])
    expect(same, "hierarchy saved and loaded: level matrices, col_map_offd and CommPkgs identical on every level")
    fz = rng.standard_normal(mats[0][0].num_rows)
    u1 = torch.zeros(mats[0][0].num_rows, dtype=torch.float64, device="cuda")
    u2 = torch.zeros(mats[0][0].num_rows, dtype=torch.float64, device="cuda")
    amg.cycle(dev(fz), u1, u_all_zeros=True)
    loaded.cycle(dev(fz), u2, u_all_zeros=True)
    expect(bool(np.array_equal(u1.cpu().numpy(), u2.cpu().numpy())), "V-cycle of the loaded hierarchy is bit-identical")
    loaded.destroy()
    dist.barrier()
    if rank == 0:
        import shutil
        shutil.rmtree(tmp[0], ignore_errors=True)

    # ---- SpMV, SpMV-T on every level
    worst = 0.0
    for l, (A, Pm) in enumerate(mats):
        for which, M in ((0, A), (1, Pm)):
            if M is None:
                continue
            x = rng.standard_normal(M.num_cols)
            b = rng.standard_normal(M.num_rows)
            for al, be in ((1.0, 0.0), (-1.0, 1.0), (0.7, -0.3)):
                yref = pb.matvec(al, x, be, b, level=l, which=which)
                y = torch.empty(M.num_rows, dtype=torch.float64, device="cuda")
                M.matvec(al, dev(x), be, y, b=dev(b))
                den = gmax(float(np.max(np.abs(yref))) if yref.size else 0.0)
                worst = max(worst, relerr(y.cpu().numpy(), yref, den))
    expect(gmax(worst) <= 1e-12, f"ParCSR matvec, all levels, A and P (worst {gmax(worst):.2e})")
    worst = 0.0
    for l, (A, Pm) in enumerate(mats):
        if Pm is None:
            continue
        x = rng.standard_normal(Pm.num_rows)
        y0 = rng.standard_normal(Pm.num_cols)
        for al, be in ((1.0, 0.0), (-2.0, 0.5)):
            yref = pb.matvecT(al, x, be, y0, level=l, which=1)
            y = dev(y0)
            Pm.matvecT(al, dev(x), be, y)
            den = gmax(float(np.max(np.abs(yref))) if yref.size else 0.0)
            worst = max(worst, relerr(y.cpu().numpy(), yref, den))
    expect(gmax(worst) <= 1e-12, f"ParCSR matvecT (restriction), all levels (worst {gmax(worst):.2e})")

    # ---- relaxation
    worst = 0.0
    for l in range(min(nl - 1, 3)):
        A = mats[l][0]
        L = h["levels"][l]
        nloc = A.num_rows
        f = rng.standard_normal(nloc)
        u = rng.standard_normal(nloc)
        cf = torch.from_numpy(np.ascontiguousarray(L["cf_marker"])).cuda() if L["cf_marker"] is not None else None
        l1 = dev(L["l1_norms"]) if L["l1_norms"] is not None else None
        for rt, pts, w, om in ((18, 0, 1.0, 1.0), (18, 1, 0.9, 1.0), (0, 0, 0.8, 1.0), (7, -1, 1.0, 1.0),
                               (13, 0, 1.0, 1.0), (14, 0, 1.0, 1.0), (8, 0, 0.9, 1.1), (6, 1, 1.0, 1.0), (3, 0, 1.0, 1.0)):
            use_l1 = rt in (7, 18, 8, 13, 14, 88, 89)
            if (use_l1 and l1 is None) or (pts != 0 and cf is None):
                continue
            uref = pb.relax(l, rt, f, u, relax_points=pts, relax_weight=w, omega=om, use_l1=use_l1)
            du = dev(u)
            hb.relax(A, dev(f), du, rt, relax_points=pts, relax_weight=w, omega=om,
                     l1_norms=l1 if use_l1 else None, cf_marker=cf)
            den = gmax(float(np.max(np.abs(uref))) if uref.size else 0.0)
            e = gmax(relerr(du.cpu().numpy(), uref, den))
            if e > 1e-12 and rank == 0:
                print(f"     relax type {rt} points {pts} w {w} omega {om} level {l}: err {e:.2e}", flush=True)
            worst = max(worst, e)
    expect(worst <= 1e-12, f"relaxation sweeps (Jacobi + hybrid GS families) (worst {worst:.2e})")

    # ---- one V-cycle
    A0 = mats[0][0]
    n0 = A0.num_rows
    f = rng.standard_normal(n0)
    uref = pb.amg_solve(f, np.zeros(n0), u_all_zeros=True)
    du = torch.zeros(n0, dtype=torch.float64, device="cuda")
    amg.cycle(dev(f), du, u_all_zeros=True)
    den = gmax(float(np.max(np.abs(uref))))
    e = gmax(relerr(du.cpu().numpy(), uref, den))
    expect(e <= 1e-12, f"V(1,1)-cycle, zero initial guess (err {e:.2e})")

    # ---- PCG
    ref = pb.pcg(precond="amg", tol=1e-8, max_iter=100, two_norm=1)
    pcg = hb.ParCSRPCG(tol=1e-8, max_iter=100, two_norm=1)
    pcg.set_precond(amg)
    x = torch.zeros(n0, dtype=torch.float64, device="cuda")
    pcg.solve(A0, dev(pb.b), x)
    k = ref["iterations"]
    ok = (pcg.num_iterations == k) and relerr(pcg.norms[: k + 1], ref["norms"]) <= 1e-9
    expect(ok, f"AMG-PCG: {pcg.num_iterations} its (reference {k}), rel.res {pcg.final_relative_residual_norm:.6e} "
               f"(reference {ref['final_rel_res']:.6e})")
    den = gmax(float(np.max(np.abs(ref["x"]))))
    e = gmax(relerr(x.cpu().numpy(), ref["x"], den))
    expect(e <= 1e-8, f"AMG-PCG solution (err {e:.2e})")

    # ---- GMRES
    ref = pb.gmres(precond="amg", tol=1e-8, max_iter=100, k_dim=5)
    gm = hb.ParCSRGMRES(tol=1e-8, max_iter=100, k_dim=5)
    gm.set_precond(amg)
    x = torch.zeros(n0, dtype=torch.float64, device="cuda")
    gm.solve(A0, dev(pb.b), x)
    expect(abs(gm.num_iterations - ref["iterations"]) <= 1,
           f"AMG-GMRES(5): {gm.num_iterations} its (reference {ref['iterations']})")
    den = gmax(float(np.max(np.abs(ref["x"]))))
    e = gmax(relerr(x.cpu().numpy(), ref["x"], den))
    expect(e <= 1e-7, f"AMG-GMRES solution (err {e:.2e})")

    dist.barrier()
    dist.destroy_process_group()
    if fails:
        raise SystemExit(f"rank {rank}: {len(fails)} parity failures: {fails}")
    if rank == 0:
        print("MULTI-RANK PARITY OK", flush=True)


if __name__ == "__main__":
    main()

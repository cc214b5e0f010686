"""Generates tests/golden/*.npz from the compiled reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run where /root/reference exists:

    OMP_NUM_THREADS=1 python tests/golden/make_golden.py

Each fixture holds one small ij problem: the hierarchy the reference's BoomerAMGSetup built
(plain CSR arrays per level, l1 norms, CF markers, Chebyshev data, coarse dense matrix, cycle
parameters) plus the reference's own outputs on seeded inputs: SpMV / SpMV-T per level, one sweep
of every relaxation type on the accelerated path, one V-cycle (fine result + every level's F and
U), the PCG and GMRES residual histories and solutions.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbridge as rb  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
rb.load()
rb.set_num_threads(1)

CASES = {
    "lap7_rlx18": dict(kind="laplacian", n=(9, 8, 7), amg=dict(relax_type=18)),
    "lap27_rlx18": dict(kind="27pt", n=(8, 7, 6), amg=dict(relax_type=18)),
    "lap7_default_gs": dict(kind="laplacian", n=(8, 8, 6), amg=dict()),
    "lap7_cheby": dict(kind="laplacian", n=(8, 7, 7), amg=dict(relax_type=16, cheby_order=3)),
    "lap7_rlx18_cf": dict(kind="laplacian", n=(9, 7, 6), amg=dict(relax_type=18, relax_order=1)),
    "vdc_rlx18": dict(kind="vardifconv", n=(8, 8, 8), amg=dict(relax_type=18)),
}
RELAX_TYPES = [0, 7, 18, 3, 4, 6, 8, 13, 14, 88, 89]


def main():
    for name, c in CASES.items():
        pb = rb.Problem(c["kind"], c["n"])
        pb.setup_amg(**c["amg"])
        h = pb.hierarchy()
        d = {}
        nl = pb.num_levels
        p = h["params"]
        d["num_levels"] = nl
        for k in ("num_grid_sweeps", "grid_relax_type"):
            d[k] = np.array(p[k], dtype=np.int32)
        for k in ("relax_order", "cycle_type", "fcycle", "cheby_order", "cheby_scale", "cheby_variant",
                  "user_relax_type", "max_iter", "min_iter", "converge_type"):
            d[k] = np.int32(p[k])
        d["tol"] = np.float64(p["tol"])
        rng = np.random.default_rng(2026)
        for l, L in enumerate(h["levels"]):
            A = L["A"].arrays()
            d[f"A{l}_i"], d[f"A{l}_j"], d[f"A{l}_a"] = A["diag_i"].copy(), A["diag_j"].copy(), A["diag_data"].copy()
            n = L["A"].num_rows
            if L["P"] is not None:
                P = L["P"].arrays()
                d[f"P{l}_i"], d[f"P{l}_j"], d[f"P{l}_a"] = P["diag_i"].copy(), P["diag_j"].copy(), P["diag_data"].copy()
                d[f"P{l}_ncols"] = np.int32(L["P"].num_cols)
            for key in ("l1_norms", "cf_marker", "cheby_ds", "cheby_coefs"):
                if L[key] is not None:
                    d[f"{key}{l}"] = np.array(L[key]).copy()
            d[f"relax_weight{l}"] = np.float64(L["relax_weight"])
            d[f"omega{l}"] = np.float64(L["omega"])
            # SpMV / SpMV-T
            x = rng.standard_normal(n)
            b = rng.standard_normal(n)
            d[f"mv_x{l}"], d[f"mv_b{l}"] = x, b
            for tag, (al, be) in {"10": (1.0, 0.0), "m11": (-1.0, 1.0), "gen": (-0.7, 0.3)}.items():
                d[f"mv_y{l}_{tag}"] = pb.matvec(al, x, be, b, level=l)
            if L["P"] is not None:
                nc = L["P"].num_cols
                xc = rng.standard_normal(nc)
                d[f"p_x{l}"] = xc
                d[f"p_y{l}"] = pb.matvec(1.0, xc, 1.0, b, level=l, which=1)
                d[f"pt_y{l}"] = pb.matvecT(1.0, x, 0.0, np.zeros(nc), level=l, which=1)
            # relaxation sweeps (levels that have CF markers and l1 norms)
            if L["cf_marker"] is not None and l < 3:
                f = rng.standard_normal(n)
                u = rng.standard_normal(n)
                d[f"rx_f{l}"], d[f"rx_u{l}"] = f, u
                for rt in RELAX_TYPES:
                    use_l1 = rt in (7, 18, 8, 13, 14, 88, 89)
                    if use_l1 and L["l1_norms"] is None:
                        continue
                    for pts in (0, 1, -1):
                        for (w, om) in ((1.0, 1.0), (0.8, 1.1)):
                            key = f"rx{l}_t{rt}_p{pts}_w{w}_o{om}"
                            d[key] = pb.relax(l, rt, f, u, relax_points=pts, relax_weight=w, omega=om, use_l1=use_l1)
                    if rt in (7, 18):
                        d[f"rx{l}_t{rt}_zero"] = pb.relax(l, rt, f, np.zeros(n), u_all_zeros=True)
                if L["cheby_coefs"] is not None:
                    d[f"cheby{l}"] = pb.cheby(l, f, u)
        ge = h["coarse_ge"]
        if ge is not None:
            d["ge_A_mat"], d["ge_n"] = ge["A_mat"], np.int32(ge["n"])
        # one cycle
        n0 = h["levels"][0]["A"].num_rows
        f = rng.standard_normal(n0)
        d["cyc_f"] = f
        d["cyc_u_zero"] = pb.amg_solve(f, np.zeros(n0), u_all_zeros=True)
        for l in range(1, nl):
            d[f"cyc_F{l}"] = pb.level_vector(l, 0)
            d[f"cyc_U{l}"] = pb.level_vector(l, 1)
        u0 = rng.standard_normal(n0)
        d["cyc_u0"] = u0
        d["cyc_u_nonzero"] = pb.amg_solve(f, u0, u_all_zeros=False)
        # Krylov
        d["b"] = np.array(pb.b)
        r = pb.pcg(precond="amg", tol=1e-8, max_iter=100, two_norm=1)
        d["pcg_its"], d["pcg_relres"], d["pcg_norms"], d["pcg_x"] = np.int32(r["iterations"]), np.float64(r["final_rel_res"]), r["norms"], r["x"]
        r = pb.pcg(precond="diagscale", tol=1e-8, max_iter=500, two_norm=1)
        d["dspcg_its"], d["dspcg_relres"], d["dspcg_norms"] = np.int32(r["iterations"]), np.float64(r["final_rel_res"]), r["norms"]
        r = pb.gmres(precond="amg", tol=1e-8, max_iter=100, k_dim=5)
        d["gmres_its"], d["gmres_relres"], d["gmres_x"] = np.int32(r["iterations"]), np.float64(r["final_rel_res"]), r["x"]
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, "levels", nl, "rows", n0, "pcg its", int(d["pcg_its"]), "gmres its", int(d["gmres_its"]),
              "size", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

"""ctypes wrapper over oracle/hypre_oracle.c — the CPU restatement of the path (1 rank).

TEST INFRASTRUCTURE (see the header of hypre_oracle.c).  Builds the shared object on demand
with gcc; needs neither the reference tree nor a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libhypre_oracle.so")
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(HERE, "hypre_oracle.c")
    if (not os.path.exists(SO)) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "restatement"], check=True, capture_output=True)
    lib = C.CDLL(SO)
    vp, d, i = C.c_void_p, C.c_double, C.c_int
    sig = {
        "ho_csr_matvec": ([i, vp, vp, vp, d, vp, d, vp, vp], None),
        "ho_csr_matvecT": ([i, i, vp, vp, vp, d, vp, d, vp], None),
        "ho_inner_prod": ([i, vp, vp], d),
        "ho_relax_raw": ([i, vp, vp, vp, vp, vp, i, i, d, d, vp, vp, i, vp], i),
        "ho_cheby_raw": ([i, vp, vp, vp, vp, vp, vp, i, i, vp], i),
        "ho_gselim": ([vp, vp, i], None),
        "ho_amg_create": ([i], vp),
        "ho_amg_destroy": ([vp], None),
        "ho_amg_set_level": ([vp, i, i, vp, vp, vp, i, vp, vp, vp, vp, vp, d, d, vp, vp], None),
        "ho_amg_set_params": ([vp, vp, vp, i, i, i, i, i, i, d, i, i, i], None),
        "ho_amg_set_coarse_ge": ([vp, vp, i], None),
        "ho_amg_cycle": ([vp, vp, vp, i], i),
        "ho_amg_solve": ([vp, vp, vp, i, vp, vp], i),
        "ho_amg_level_vector": ([vp, i, i, vp], None),
        "ho_pcg": ([i, vp, vp, vp, i, vp, d, d, i, i, i, i, i, vp, vp, vp, vp, vp], i),
        "ho_gmres": ([i, vp, vp, vp, i, vp, d, d, i, i, vp, vp, vp, vp], i),
    }
    for name, (a, r) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = a, r
    _lib = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def csr_matvec(ai, aj, aa, alpha, x, beta, b=None):
    lib = load()
    ai, aj, aa, x = _i(ai), _i(aj), _f(aa), _f(x)
    n = ai.shape[0] - 1
    b = np.zeros(n) if b is None else _f(b)
    y = np.zeros(n)
    lib.ho_csr_matvec(n, _p(ai), _p(aj), _p(aa), alpha, _p(x), beta, _p(b), _p(y))
    return y


def csr_matvecT(ai, aj, aa, ncols, alpha, x, beta, y):
    lib = load()
    ai, aj, aa, x = _i(ai), _i(aj), _f(aa), _f(x)
    y = np.array(y, dtype=np.float64, copy=True)
    lib.ho_csr_matvecT(ai.shape[0] - 1, ncols, _p(ai), _p(aj), _p(aa), alpha, _p(x), beta, _p(y))
    return y


def relax(ai, aj, aa, f, u, relax_type, relax_points=0, w=1.0, omega=1.0, l1=None, cf=None, u_all_zeros=False):
    lib = load()
    ai, aj, aa, f, l1, cf = _i(ai), _i(aj), _f(aa), _f(f), _f(l1), _i(cf)
    u = np.array(u, dtype=np.float64, copy=True)
    vt = np.zeros_like(u)
    rc = lib.ho_relax_raw(ai.shape[0] - 1, _p(ai), _p(aj), _p(aa), _p(f), _p(cf), relax_type, relax_points,
                          w, omega, _p(l1), _p(u), 1 if u_all_zeros else 0, _p(vt))
    if rc:
        raise ValueError(f"relax type {relax_type} is not restated")
    return u


def cheby(ai, aj, aa, f, u, coefs, order, scale, ds=None):
    lib = load()
    ai, aj, aa, f, ds, coefs = _i(ai), _i(aj), _f(aa), _f(f), _f(ds), _f(coefs)
    u = np.array(u, dtype=np.float64, copy=True)
    lib.ho_cheby_raw(ai.shape[0] - 1, _p(ai), _p(aj), _p(aa), _p(f), _p(ds), _p(coefs), order, scale, _p(u))
    return u


class AMG:
    """Restated hypre_ParAMGData (solve part) over plain numpy arrays: levels = list of dicts with
    A=(i,j,a), P=(i,j,a,ncols)|None, l1_norms, cf_marker, relax_weight, omega, cheby_ds, cheby_coefs."""

    def __init__(self, levels, params, coarse_ge=None):
        self.lib = load()
        self.keep = []
        self.h = self.lib.ho_amg_create(len(levels))
        self.nrows = []
        for l, L in enumerate(levels):
            ai, aj, aa = _i(L["A"][0]), _i(L["A"][1]), _f(L["A"][2])
            P = L.get("P")
            if P is not None:
                pi, pj, pa, pnc = _i(P[0]), _i(P[1]), _f(P[2]), int(P[3])
            else:
                pi = pj = pa = None
                pnc = 0
            l1, cf = _f(L.get("l1_norms")), _i(L.get("cf_marker"))
            ds, cc = _f(L.get("cheby_ds")), _f(L.get("cheby_coefs"))
            self.keep.append((ai, aj, aa, pi, pj, pa, l1, cf, ds, cc))
            n = ai.shape[0] - 1
            self.nrows.append(n)
            self.lib.ho_amg_set_level(self.h, l, n, _p(ai), _p(aj), _p(aa), pnc, _p(pi), _p(pj), _p(pa),
                                      _p(l1), _p(cf), float(L.get("relax_weight", 1.0)),
                                      float(L.get("omega", 1.0)), _p(ds), _p(cc))
        ngs, grt = _i(params["num_grid_sweeps"]), _i(params["grid_relax_type"])
        self.lib.ho_amg_set_params(self.h, _p(ngs), _p(grt), params["relax_order"], params["cycle_type"],
                                   params["fcycle"], params["cheby_order"], params["cheby_scale"],
                                   params["user_relax_type"], params["tol"], params["min_iter"],
                                   params["max_iter"], params["converge_type"])
        if coarse_ge is not None:
            am = _f(coarse_ge["A_mat"])
            self.lib.ho_amg_set_coarse_ge(self.h, _p(am), int(coarse_ge["n"]))

    def __del__(self):
        try:
            self.lib.ho_amg_destroy(self.h)
        except Exception:
            pass

    def solve(self, f, u, u_all_zeros=False):
        f = _f(f)
        u = np.array(u, dtype=np.float64, copy=True)
        self.lib.ho_amg_solve(self.h, _p(f), _p(u), 1 if u_all_zeros else 0, None, None)
        return u

    def level_vector(self, level, which):
        out = np.zeros(self.nrows[level])
        self.lib.ho_amg_level_vector(self.h, level, which, _p(out))
        return out


def pcg(A, b, x0, precond="amg", amg=None, tol=1e-8, a_tol=0.0, max_iter=100, two_norm=1, rel_change=0,
        flex=0, recompute_residual=0):
    lib = load()
    ai, aj, aa = _i(A[0]), _i(A[1]), _f(A[2])
    b = _f(b)
    x = np.array(x0, dtype=np.float64, copy=True)
    its, rr = C.c_int(0), C.c_double(0.0)
    norms = np.zeros(max_iter + 2)
    pk = {"none": 0, "amg": 1, "diagscale": 2}[precond]
    flag = lib.ho_pcg(ai.shape[0] - 1, _p(ai), _p(aj), _p(aa), pk, amg.h if amg is not None else None, tol, a_tol,
                      max_iter, two_norm, rel_change, flex, recompute_residual, _p(b), _p(x), C.addressof(its),
                      C.addressof(rr), _p(norms))
    return {"iterations": its.value, "final_rel_res": rr.value, "norms": norms[: its.value + 1], "x": x,
            "error_flag": flag}


def gmres(A, b, x0, precond="amg", amg=None, tol=1e-8, a_tol=0.0, max_iter=100, k_dim=5):
    lib = load()
    ai, aj, aa = _i(A[0]), _i(A[1]), _f(A[2])
    b = _f(b)
    x = np.array(x0, dtype=np.float64, copy=True)
    its, rr = C.c_int(0), C.c_double(0.0)
    pk = {"none": 0, "amg": 1, "diagscale": 2}[precond]
    flag = lib.ho_gmres(ai.shape[0] - 1, _p(ai), _p(aj), _p(aa), pk, amg.h if amg is not None else None, tol, a_tol,
                        max_iter, k_dim, _p(b), _p(x), C.addressof(its), C.addressof(rr))
    return {"iterations": its.value, "final_rel_res": rr.value, "x": x, "error_flag": flag}

"""ctypes wrapper over oracle/_ref/libref_bridge.so — the UNMODIFIED compiled reference
(hypre 3.1.0, CPU, OpenMP) driven through oracle/ref_bridge.c.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py (setup provider +
cpu_baseline / --impl reference legs) import this.  The product (hypre_b200/) never does.

Roles:
  * builds ij's test problems with the reference's generators and runs the reference's own
    BoomerAMGSetup (north_star: "consumes the hierarchy built by the reference's own
    BoomerAMGSetup, uploaded once and timed separately");
  * exposes the hierarchy as plain arrays (`hierarchy()`), the input of
    hypre_b200.amg_from_hierarchy;
  * runs the reference's CPU matvec / relax / cycle / PCG / GMRES on the same inputs = the
    parity oracle and the CPU baseline.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

c_int_p = C.POINTER(C.c_int)
c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)


class ParCSRView(C.Structure):
    _fields_ = [
        ("num_rows", C.c_int), ("num_cols", C.c_int), ("num_cols_offd", C.c_int),
        ("diag_nnz", C.c_int), ("offd_nnz", C.c_int),
        ("diag_i", c_int_p), ("diag_j", c_int_p), ("diag_data", c_double_p),
        ("offd_i", c_int_p), ("offd_j", c_int_p), ("offd_data", c_double_p),
        ("col_map_offd", c_int64_p),
        ("first_row", C.c_int64), ("first_col", C.c_int64),
        ("global_rows", C.c_int64), ("global_cols", C.c_int64),
        ("num_sends", C.c_int), ("num_recvs", C.c_int),
        ("send_procs", c_int_p), ("send_map_starts", c_int_p), ("send_map_elmts", c_int_p),
        ("recv_procs", c_int_p), ("recv_vec_starts", c_int_p),
    ]

    def arrays(self) -> dict:
        """numpy views (no copies) of every array in the view"""
        def a(p, n):
            return np.ctypeslib.as_array(p, shape=(n,)) if (p and n > 0) else None
        ns, nr = self.num_sends, self.num_recvs
        sms = a(self.send_map_starts, ns + 1)
        return {
            "diag_i": a(self.diag_i, self.num_rows + 1), "diag_j": a(self.diag_j, self.diag_nnz),
            "diag_data": a(self.diag_data, self.diag_nnz),
            "offd_i": a(self.offd_i, self.num_rows + 1) if self.num_cols_offd else None,
            "offd_j": a(self.offd_j, self.offd_nnz), "offd_data": a(self.offd_data, self.offd_nnz),
            "col_map_offd": a(self.col_map_offd, self.num_cols_offd),
            "send_procs": a(self.send_procs, ns), "send_map_starts": sms,
            "send_map_elmts": a(self.send_map_elmts, int(sms[ns]) if ns else 0),
            "recv_procs": a(self.recv_procs, nr), "recv_vec_starts": a(self.recv_vec_starts, nr + 1),
        }


_lib = None
_mpi = False


def available(mpi: bool = False) -> bool:
    name = "libref_bridge_mpi.so" if mpi else "libref_bridge.so"
    return os.path.exists(os.path.join(REF_DIR, name))


def load(mpi=None) -> C.CDLL:
    """Loads the bridge (and through it libHYPRE_ref*.so).  One flavour per process;
    mpi=None means "whatever is loaded" (serial if nothing is)."""
    global _lib, _mpi
    if _lib is not None:
        if mpi is not None and bool(mpi) != _mpi:
            raise RuntimeError("the serial and the mini-MPI reference builds cannot share a process")
        return _lib
    mpi = bool(mpi)
    name = "libref_bridge_mpi.so" if mpi else "libref_bridge.so"
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: run `make -C oracle ref bridge` where /root/reference exists")
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    vp = C.c_void_p
    dp = c_double_p
    sig = {
        "rb_init": ([], C.c_int), "rb_finalize": ([], C.c_int),
        "rb_num_threads": ([], C.c_int), "rb_set_num_threads": ([C.c_int], None),
        "rb_comm_rank": ([], C.c_int), "rb_comm_size": ([], C.c_int),
        "rb_error_flag": ([], C.c_int), "rb_clear_errors": ([], None),
        "rb_sizeof_bigint": ([], C.c_int),
        "rb_problem_create": ([C.c_int] * 7 + [C.c_double, C.c_int, C.c_int], vp),
        "rb_problem_destroy": ([vp], None),
        "rb_problem_from_ij_file": ([C.c_char_p, C.c_int], vp),
        "rb_print_ij": ([vp, C.c_char_p], C.c_int),
        "rb_print_ij_binary": ([vp, C.c_char_p], C.c_int),
        "rb_print_vector_ij": ([vp, vp, C.c_char_p], C.c_int),
        "rb_problem_b": ([vp], dp), "rb_problem_x": ([vp], dp),
        "rb_problem_local_rows": ([vp], C.c_int), "rb_problem_global_rows": ([vp], C.c_longlong),
        "rb_amg_setup": ([vp] + [C.c_int] * 15 + [C.c_double] * 4 + [C.c_int] * 3 + [dp], C.c_int),
        "rb_num_levels": ([vp], C.c_int),
        "rb_level_matrix": ([vp, C.c_int, C.c_int, C.POINTER(ParCSRView)], C.c_int),
        "rb_level_l1_norms": ([vp, C.c_int], dp), "rb_level_cf_marker": ([vp, C.c_int], c_int_p),
        "rb_level_relax_weight": ([vp, C.c_int], C.c_double), "rb_level_omega": ([vp, C.c_int], C.c_double),
        "rb_level_cheby_ds": ([vp, C.c_int], dp), "rb_level_cheby_coefs": ([vp, C.c_int], dp),
        "rb_amg_params": ([vp, c_int_p, dp], C.c_int),
        "rb_coarse_ge": ([vp, C.POINTER(dp), c_int_p, c_int_p], C.c_int),
        "rb_matvec": ([vp, C.c_int, C.c_int, C.c_double, vp, C.c_double, vp, vp], C.c_int),
        "rb_matvecT": ([vp, C.c_int, C.c_int, C.c_double, vp, C.c_double, vp], C.c_int),
        "rb_matvec_time": ([vp, C.c_int], C.c_double),
        "rb_relax": ([vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp, C.c_int, C.c_int, C.c_int], C.c_int),
        "rb_cheby": ([vp, C.c_int, vp, vp], C.c_int),
        "rb_amg_solve": ([vp, vp, vp, C.c_int], C.c_int),
        "rb_amg_set_solve": ([vp, C.c_double, C.c_int], C.c_int),
        "rb_level_vector": ([vp, C.c_int, C.c_int, vp], C.c_int),
        "rb_pcg_solve": ([vp, C.c_int, C.c_double, C.c_double] + [C.c_int] * 5 + [vp, vp, c_int_p, dp, vp, dp], C.c_int),
        "rb_gmres_solve": ([vp, C.c_int, C.c_double, C.c_double] + [C.c_int] * 3 + [vp, vp, c_int_p, dp, vp, dp], C.c_int),
        "rb_krylov_ext_solve": ([vp, C.c_int, C.c_int, C.c_double, C.c_double] + [C.c_int] * 4 + [vp, vp, c_int_p, dp, vp, dp], C.c_int),
        "rb_inner_prod": ([vp, vp, vp], C.c_double),
        "rb_axpy": ([vp, C.c_double, vp, vp], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, res
    lib.rb_init()
    _lib, _mpi = lib, mpi
    return lib


def _p(a):
    return None if a is None else a.ctypes.data


PROBLEM_TYPES = {"laplacian": 0, "27pt": 1, "vardifconv": 2}


class Problem:
    """One ij test problem: matrix from the reference's generator, b, x0, and (after
    setup_amg) the reference's BoomerAMG hierarchy."""

    def __init__(self, kind: str, n, P=(1, 1, 1), eps: float = 1.0, rhs: str = "ones",
                 x0rand: bool = False, mpi=None):
        self.lib = load(mpi)
        nx, ny, nz = n
        rhs_type = {"ones": 0, "rand": 1}[rhs]
        self.h = self.lib.rb_problem_create(PROBLEM_TYPES[kind], nx, ny, nz, P[0], P[1], P[2],
                                            eps, rhs_type, 1 if x0rand else 0)
        if not self.h:
            raise RuntimeError("rb_problem_create failed")
        self.kind, self.n = kind, tuple(n)
        self.local_rows = self.lib.rb_problem_local_rows(self.h)
        self.global_rows = self.lib.rb_problem_global_rows(self.h)
        self.setup_seconds = None

    @classmethod
    def from_ij_file(cls, filename: str, matrix_market: bool = False, binary: bool = False, mpi=None) -> "Problem":
        """ij -fromfile / -frombinfile: the matrix the reference's own HYPRE_IJMatrixRead (/ ReadMM / ReadBinary)
        assembles from the file(s)"""
        self = cls.__new__(cls)
        self.lib = load(mpi)
        self.h = self.lib.rb_problem_from_ij_file(filename.encode(), 2 if binary else (1 if matrix_market else 0))
        if not self.h:
            raise RuntimeError(f"the reference could not read {filename}")
        self.kind, self.n = "file", None
        self.local_rows = self.lib.rb_problem_local_rows(self.h)
        self.global_rows = self.lib.rb_problem_global_rows(self.h)
        self.setup_seconds = None
        return self

    def print_ij(self, filename: str, binary: bool = False) -> None:
        """HYPRE_IJMatrixPrint[Binary] of the fine-level operator: `<filename>.<5-digit rank>[.bin]`"""
        fn = self.lib.rb_print_ij_binary if binary else self.lib.rb_print_ij
        if fn(self.h, filename.encode()):
            raise RuntimeError("reference IJ print failed")

    def print_vector_ij(self, values, filename: str) -> None:
        v = np.ascontiguousarray(values, np.float64)
        if self.lib.rb_print_vector_ij(self.h, _p(v), filename.encode()):
            raise RuntimeError("reference IJ vector print failed")

    def destroy(self):
        if getattr(self, "h", None):
            self.lib.rb_problem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    @property
    def b(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.lib.rb_problem_b(self.h), shape=(self.local_rows,))

    @property
    def x0(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.lib.rb_problem_x(self.h), shape=(self.local_rows,))

    def setup_amg(self, relax_type=-1, relax_down=-1, relax_up=-1, relax_coarse=-1, relax_order=-1,
                  num_sweeps=-1, coarsen_type=-1, interp_type=-1, Pmx=-1, agg_nl=-1, cycle_type=-1,
                  cheby_order=-1, cheby_eig_est=-1, cheby_scale=-1, cheby_variant=-1,
                  cheby_fraction=-1.0, strong_threshold=-1.0, relax_wt=-1e300, outer_wt=-1e300,
                  keep_transpose=-1, max_levels=-1, print_level=0) -> float:
        """The reference's HYPRE_BoomerAMGSetup with ij's defaults; -1 = ij default.
        NB: like `ij`, pass relax_type=18 explicitly for l1-Jacobi (CPU default is 13/14)."""
        t = C.c_double(0.0)
        flag = self.lib.rb_amg_setup(self.h, relax_type, relax_down, relax_up, relax_coarse, relax_order,
                                     num_sweeps, coarsen_type, interp_type, Pmx, agg_nl, cycle_type,
                                     cheby_order, cheby_eig_est, cheby_scale, cheby_variant,
                                     cheby_fraction, strong_threshold, relax_wt, outer_wt,
                                     keep_transpose, max_levels, print_level, C.byref(t))
        if flag:
            raise RuntimeError(f"reference BoomerAMGSetup raised hypre error flag {flag}")
        self.setup_seconds = t.value
        return t.value

    # ---- hierarchy -----------------------------------------------------------------------
    @property
    def num_levels(self) -> int:
        return self.lib.rb_num_levels(self.h)

    def level_view(self, level: int, which: int = 0) -> ParCSRView:
        v = ParCSRView()
        if self.lib.rb_level_matrix(self.h, level, which, C.byref(v)):
            raise RuntimeError(f"no matrix at level {level} kind {which}")
        return v

    def amg_params(self) -> dict:
        out = (C.c_int * 24)()
        tol = C.c_double(0.0)
        self.lib.rb_amg_params(self.h, out, C.byref(tol))
        o = list(out)
        return {
            "num_grid_sweeps": o[0:4], "grid_relax_type": o[4:8], "relax_order": o[8],
            "cycle_type": o[9], "fcycle": o[10], "cheby_order": o[11], "cheby_scale": o[12],
            "cheby_variant": o[13], "user_relax_type": o[14], "max_iter": o[15], "min_iter": o[16],
            "converge_type": o[17], "restriction": o[18], "block_mode": o[19],
            "smooth_num_levels": o[20], "additive": o[21], "mult_additive": o[22], "simple": o[23],
            "tol": tol.value,
        }

    def hierarchy(self) -> dict:
        """Plain-array description of the reference's hierarchy (SURVEY Appendix B manifest)."""
        nl = self.num_levels
        params = self.amg_params()
        levels = []
        for l in range(nl):
            A = self.level_view(l, 0)
            P = self.level_view(l, 1) if l < nl - 1 else None
            n = A.num_rows
            l1p = self.lib.rb_level_l1_norms(self.h, l)
            cfp = self.lib.rb_level_cf_marker(self.h, l)
            dsp = self.lib.rb_level_cheby_ds(self.h, l)
            ccp = self.lib.rb_level_cheby_coefs(self.h, l)
            uses_cheby = 16 in params["grid_relax_type"][1:4]
            levels.append({
                "A": A, "P": P,
                "l1_norms": np.ctypeslib.as_array(l1p, shape=(n,)) if (l1p and n) else None,
                "cf_marker": np.ctypeslib.as_array(cfp, shape=(n,)) if (cfp and n) else None,
                "relax_weight": self.lib.rb_level_relax_weight(self.h, l),
                "omega": self.lib.rb_level_omega(self.h, l),
                "cheby_ds": np.ctypeslib.as_array(dsp, shape=(n,)) if (uses_cheby and dsp and n) else None,
                "cheby_coefs": (np.ctypeslib.as_array(ccp, shape=(params["cheby_order"] + 1,)).copy()
                                if (uses_cheby and ccp) else None),
            })
        amat = c_double_p()
        fr, nloc = C.c_int(0), C.c_int(0)
        n = self.lib.rb_coarse_ge(self.h, C.byref(amat), C.byref(fr), C.byref(nloc))
        ge = None
        if n > 0:
            # ranks without coarse rows have no A_mat (par_gauss_elim.c:277-286)
            mat = (np.ctypeslib.as_array(amat, shape=(n * n,)).copy() if amat
                   else np.zeros(n * n))
            ge = {"A_mat": mat, "n": n, "first_row": fr.value, "num_local": nloc.value}
        return {"levels": levels, "params": params, "coarse_ge": ge}

    # ---- reference compute (oracle) -----------------------------------------------------------
    def matvec(self, alpha, x, beta, b=None, level=0, which=0) -> np.ndarray:
        v = self.level_view(level, which)
        x = np.ascontiguousarray(x, np.float64)
        bb = np.zeros(v.num_rows) if b is None else np.ascontiguousarray(b, np.float64)
        y = np.zeros(v.num_rows)
        self.lib.rb_matvec(self.h, level, which, alpha, _p(x), beta, _p(bb), _p(y))
        return y

    def matvecT(self, alpha, x, beta, y, level=0, which=1) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float64)
        y = np.array(y, dtype=np.float64, copy=True)
        self.lib.rb_matvecT(self.h, level, which, alpha, _p(x), beta, _p(y))
        return y

    def matvec_time(self, nmv: int) -> float:
        return self.lib.rb_matvec_time(self.h, nmv)

    def relax(self, level, relax_type, f, u, relax_points=0, relax_weight=1.0, omega=1.0,
              u_all_zeros=False, use_l1=True, use_cf=True) -> np.ndarray:
        f = np.ascontiguousarray(f, np.float64)
        u = np.array(u, dtype=np.float64, copy=True)
        flag = self.lib.rb_relax(self.h, level, relax_type, relax_points, relax_weight, omega,
                                 _p(f), _p(u), 1 if u_all_zeros else 0, 1 if use_l1 else 0,
                                 1 if use_cf else 0)
        if flag:
            self.lib.rb_clear_errors()
            raise RuntimeError(f"reference relax raised hypre error flag {flag}")
        return u

    def cheby(self, level, f, u) -> np.ndarray:
        f = np.ascontiguousarray(f, np.float64)
        u = np.array(u, dtype=np.float64, copy=True)
        self.lib.rb_cheby(self.h, level, _p(f), _p(u))
        return u

    def amg_solve(self, f, u, u_all_zeros=False) -> np.ndarray:
        f = np.ascontiguousarray(f, np.float64)
        u = np.array(u, dtype=np.float64, copy=True)
        self.lib.rb_amg_solve(self.h, _p(f), _p(u), 1 if u_all_zeros else 0)
        return u

    def level_vector(self, level, which) -> np.ndarray:
        n = self.level_view(level, 0).num_rows
        out = np.zeros(n)
        self.lib.rb_level_vector(self.h, level, which, _p(out))
        return out

    def pcg(self, precond="amg", tol=1e-8, atol=0.0, max_iter=100, two_norm=1, rel_change=0, flex=0,
            recompute_res=0, b=None, x0=None) -> dict:
        """ij -solver 1 (precond="amg"), -solver 2 ("diagscale"); max_iter = ij's mg_max_iter"""
        pk = {"none": 0, "amg": 1, "diagscale": 2}[precond]
        its, fr, t = C.c_int(0), C.c_double(0.0), C.c_double(0.0)
        norms = np.zeros(max_iter + 2)
        x = np.array(self.x0 if x0 is None else x0, dtype=np.float64, copy=True)
        bb = None if b is None else np.ascontiguousarray(b, np.float64)
        flag = self.lib.rb_pcg_solve(self.h, pk, tol, atol, max_iter, two_norm, rel_change, flex,
                                     recompute_res, _p(bb), _p(x), C.byref(its), C.byref(fr),
                                     _p(norms), C.byref(t))
        return {"iterations": its.value, "final_rel_res": fr.value, "norms": norms[: its.value + 1],
                "x": x, "seconds": t.value, "error_flag": flag}

    def gmres(self, precond="amg", tol=1e-8, atol=0.0, max_iter=100, k_dim=5, rel_change=0,
              b=None, x0=None) -> dict:
        """ij -solver 3 (precond="amg"), -solver 4 ("diagscale")"""
        pk = {"none": 0, "amg": 1, "diagscale": 2}[precond]
        its, fr, t = C.c_int(0), C.c_double(0.0), C.c_double(0.0)
        norms = np.zeros(max_iter + 2)
        x = np.array(self.x0 if x0 is None else x0, dtype=np.float64, copy=True)
        bb = None if b is None else np.ascontiguousarray(b, np.float64)
        flag = self.lib.rb_gmres_solve(self.h, pk, tol, atol, max_iter, k_dim, rel_change, _p(bb),
                                       _p(x), C.byref(its), C.byref(fr), _p(norms), C.byref(t))
        return {"iterations": its.value, "final_rel_res": fr.value, "norms": norms[: its.value + 1],
                "x": x, "seconds": t.value, "error_flag": flag}

    def krylov_ext(self, which, precond="amg", tol=1e-8, atol=0.0, max_iter=100, k_dim=5, cgs=1, rel_change=0,
                   aug_dim=2, b=None, x0=None) -> dict:
        """ij -solver 9/10 (which="bicgstab"), 61/60 ("flexgmres"), 16/17 ("cogmres"), 50/51 ("lgmres") with
        BoomerAMG / diagonal scaling"""
        wk = {"bicgstab": 0, "flexgmres": 1, "cogmres": 2, "lgmres": 3}[which]
        if which == "lgmres":
            cgs = aug_dim
        pk = {"none": 0, "amg": 1, "diagscale": 2}[precond]
        its, fr, t = C.c_int(0), C.c_double(0.0), C.c_double(0.0)
        norms = np.zeros(max_iter + 2)
        x = np.array(self.x0 if x0 is None else x0, dtype=np.float64, copy=True)
        bb = None if b is None else np.ascontiguousarray(b, np.float64)
        flag = self.lib.rb_krylov_ext_solve(self.h, wk, pk, tol, atol, max_iter, k_dim, cgs, rel_change, _p(bb),
                                            _p(x), C.byref(its), C.byref(fr), _p(norms), C.byref(t))
        return {"iterations": its.value, "final_rel_res": fr.value, "norms": norms[: its.value + 1],
                "x": x, "seconds": t.value, "error_flag": flag}

    def inner_prod(self, x, y) -> float:
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        return self.lib.rb_inner_prod(self.h, _p(x), _p(y))


def num_threads() -> int:
    return load().rb_num_threads()


def set_num_threads(n: int) -> None:
    load().rb_set_num_threads(n)

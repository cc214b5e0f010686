"""Host-side mirror of the reference interface for the BoomerAMG solve path.

Class and method names follow the reference's C API for this path so that parity tests read
like the reference's own drivers (src/test/ij.c):

    ParCSRMatrix   ~ HYPRE_ParCSRMatrix      (matvec = HYPRE_ParCSRMatrixMatvec[T])
    BoomerAMG      ~ HYPRE_BoomerAMG*        (solve phase only; hierarchy comes from the
                                              reference's own BoomerAMGSetup, uploaded once)
    ParCSRPCG      ~ HYPRE_ParCSRPCG* / HYPRE_PCG*
    ParCSRGMRES    ~ HYPRE_ParCSRGMRES* / HYPRE_GMRES*

Everything here is plumbing over the C-ABI in include/hb200.h; the arithmetic lives in
libhb200.so (hand-written sm_100a CUDA).  Device vectors are torch CUDA float64 tensors
(torch is used for device memory only); host vectors are numpy float64 arrays.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from ._lib import (BiCGSTABParams, GMRESParams, HB200Error, KrylovResult, PCGParams, check, lib)

_initialized = False


def init(device: int = 0) -> None:
    """hb200_init: bind this process to one GPU.  Raises if there is no usable B200."""
    global _initialized
    if _initialized:
        return
    check(lib.hb200_init(int(device)))
    _initialized = True


def finalize() -> None:
    global _initialized
    if _initialized:
        lib.hb200_finalize()
        _initialized = False


def comm_init(rank: int, nranks: int, unique_id: Optional[bytes]) -> None:
    buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
    check(lib.hb200_comm_init(rank, nranks, buf))


def comm_get_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib.hb200_comm_get_unique_id(buf))
    return buf.raw


def sync() -> None:
    check(lib.hb200_sync())


def launch_count(reset: bool = False) -> int:
    return int(lib.hb200_launch_count(1 if reset else 0))


def _np(a, dtype):
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a) -> Optional[int]:
    """host numpy array / torch cuda tensor / raw int -> address"""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(f"cannot take the address of {type(a)}")


def _torch_sync(*tensors) -> None:
    # torch runs on its own stream; libhb200 on its compute stream
    for t in tensors:
        if t is not None and hasattr(t, "is_cuda") and t.is_cuda:
            import torch
            torch.cuda.synchronize()
            return


class ParCSRMatrix:
    """Device-resident hypre_ParCSRMatrix (diag + offd CSR, col_map_offd, CommPkg)."""

    def __init__(self, num_rows: int, num_cols: int, diag_i, diag_j, diag_data,
                 offd_i=None, offd_j=None, offd_data=None, col_map_offd=None,
                 first_row: int = 0, first_col: int = 0,
                 global_rows: Optional[int] = None, global_cols: Optional[int] = None,
                 send_procs=None, send_map_starts=None, send_map_elmts=None,
                 recv_procs=None, recv_vec_starts=None):
        init_required()
        self.num_rows, self.num_cols = int(num_rows), int(num_cols)
        di, dj, dd = _np(diag_i, np.int32), _np(diag_j, np.int32), _np(diag_data, np.float64)
        cm = _np(col_map_offd, np.int64)
        self.num_cols_offd = 0 if cm is None else int(cm.shape[0])
        oi = _np(offd_i, np.int32) if self.num_cols_offd else None
        oj = _np(offd_j, np.int32) if self.num_cols_offd else None
        od = _np(offd_data, np.float64) if self.num_cols_offd else None
        sp, sms, sme = _np(send_procs, np.int32), _np(send_map_starts, np.int32), _np(send_map_elmts, np.int32)
        rp, rvs = _np(recv_procs, np.int32), _np(recv_vec_starts, np.int32)
        ns = 0 if sp is None else int(sp.shape[0])
        nr = 0 if rp is None else int(rp.shape[0])
        self.num_sends, self.num_recvs = ns, nr
        self.n_send_elmts = int(sms[ns]) if ns else 0
        self.diag_nnz = int(di[self.num_rows]) if self.num_rows else 0
        self.offd_nnz = int(oi[self.num_rows]) if (oi is not None and self.num_rows) else 0
        h = C.c_void_p()
        check(lib.hb200_parcsr_create(
            C.byref(h), self.num_rows, self.num_cols, self.num_cols_offd,
            _ptr(di), _ptr(dj), _ptr(dd), _ptr(oi), _ptr(oj), _ptr(od), _ptr(cm),
            int(first_row), int(first_col),
            int(global_rows if global_rows is not None else num_rows),
            int(global_cols if global_cols is not None else num_cols),
            ns, _ptr(sp), _ptr(sms), _ptr(sme), nr, _ptr(rp), _ptr(rvs)))
        self.handle = h

    @classmethod
    def _adopt(cls, handle) -> "ParCSRMatrix":
        """wrap a matrix the library assembled itself (from_ij / read_ij): sizes from hb200_parcsr_info"""
        info = (C.c_int64 * 12)()
        check(lib.hb200_parcsr_info(handle, info))
        m = cls.__new__(cls)
        m.handle = handle
        m.num_rows, m.num_cols, m.num_cols_offd = int(info[0]), int(info[1]), int(info[2])
        m.diag_nnz, m.offd_nnz = int(info[3]), int(info[4])
        m.num_sends, m.num_recvs, m.n_send_elmts = int(info[5]), int(info[6]), int(info[7])
        m.first_row, m.first_col, m.global_rows, m.global_cols = (int(info[k]) for k in range(8, 12))
        return m

    @classmethod
    def from_ij(cls, ilower: int, iupper: int, jlower: int, jupper: int, rows, cols, values,
                add_duplicates: bool = False) -> "ParCSRMatrix":
        """HYPRE_IJMatrixCreate + SetValues / AddToValues + Assemble: this rank's coordinate triplets (global
        indices) become its rows of a device-resident ParCSR matrix.  Collective."""
        init_required()
        r, c_, v = _np(rows, np.int64), _np(cols, np.int64), _np(values, np.float64)
        n = 0 if r is None else int(r.shape[0])
        h = C.c_void_p()
        check(lib.hb200_parcsr_from_ij(C.byref(h), int(ilower), int(iupper), int(jlower), int(jupper), n,
                                       _ptr(r), _ptr(c_), _ptr(v), 1 if add_duplicates else 0))
        return cls._adopt(h)

    @classmethod
    def read_ij(cls, filename: str, matrix_market: bool = False, binary: bool = False) -> "ParCSRMatrix":
        """HYPRE_IJMatrixRead / ReadMM / ReadBinary: `<filename>.<5-digit rank>[.bin]` per rank, or one Matrix Market file"""
        init_required()
        h = C.c_void_p()
        check(lib.hb200_parcsr_read_ij(C.byref(h), filename.encode(), 2 if binary else (1 if matrix_market else 0)))
        return cls._adopt(h)

    def print_ij(self, filename: str, binary: bool = False) -> None:
        """HYPRE_IJMatrixPrint / PrintBinary: the reference's formats, one file per rank"""
        fn = lib.hb200_parcsr_print_ij_binary if binary else lib.hb200_parcsr_print_ij
        check(fn(self.handle, filename.encode()))

    @classmethod
    def from_view(cls, v) -> "ParCSRMatrix":
        """Build from any object exposing the hypre_ParCSRMatrix fields as arrays/pointers
        (e.g. the reference bridge's view of A_array[l] / P_array[l])."""
        def arr(ptr, n, dtype):
            if ptr is None or n <= 0:
                return None
            if isinstance(ptr, np.ndarray):
                return np.ascontiguousarray(ptr[:n], dtype=dtype)
            if not ptr:          # NULL ctypes pointer
                return None
            a = np.ctypeslib.as_array(ptr, shape=(n,))
            assert a.dtype == np.dtype(dtype), (a.dtype, dtype)
            return a
        nr = int(v.num_rows)
        nco = int(v.num_cols_offd)
        ns, nrecv = int(v.num_sends), int(v.num_recvs)
        sms = arr(v.send_map_starts, ns + 1, np.int32) if ns else None
        nse = int(sms[ns]) if ns else 0
        return cls(
            nr, int(v.num_cols),
            arr(v.diag_i, nr + 1, np.int32) if nr else np.zeros(1, np.int32),
            arr(v.diag_j, int(v.diag_nnz), np.int32), arr(v.diag_data, int(v.diag_nnz), np.float64),
            arr(v.offd_i, nr + 1, np.int32) if nco else None,
            arr(v.offd_j, int(v.offd_nnz), np.int32) if nco else None,
            arr(v.offd_data, int(v.offd_nnz), np.float64) if nco else None,
            arr(v.col_map_offd, nco, np.int64) if nco else None,
            int(v.first_row), int(v.first_col), int(v.global_rows), int(v.global_cols),
            arr(v.send_procs, ns, np.int32) if ns else None, sms,
            arr(v.send_map_elmts, nse, np.int32) if nse else (np.zeros(0, np.int32) if ns else None),
            arr(v.recv_procs, nrecv, np.int32) if nrecv else None,
            arr(v.recv_vec_starts, nrecv + 1, np.int32) if nrecv else None)

    def destroy(self) -> None:
        if getattr(self, "handle", None):
            if not getattr(self, "_borrowed", False):      # (a level matrix of a loaded hierarchy belongs to it)
                lib.hb200_parcsr_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    @property
    def num_nonzeros(self) -> int:
        return int(lib.hb200_parcsr_num_nonzeros(self.handle))

    def set_spmv_kernel(self, kind: int = 0, lanes_per_row: int = 0) -> None:
        check(lib.hb200_parcsr_set_spmv_kernel(self.handle, kind, lanes_per_row))

    def set_gs_chunks(self, num_chunks: int) -> None:
        """hybrid Gauss-Seidel on this matrix: the reference's thread count (hypre_NumThreads(), par_relax.c:727);
        0 / 1 = the sequential sweep, T > 1 = the reference's result at OMP_NUM_THREADS = T in one launch"""
        check(lib.hb200_parcsr_set_gs_chunks(self.handle, num_chunks))

    def format_info(self) -> dict:
        """storage formats of the diag block (hb200_parcsr_format_info)"""
        info = (C.c_longlong * 10)()
        check(lib.hb200_parcsr_format_info(self.handle, info))
        return {"sell": bool(info[0]), "sell_entries": int(info[1]), "sell_bytes_per_entry": int(info[2]),
                "sell_values": int(info[3]), "pattern": bool(info[4] & 1), "box": bool(info[4] & 2),
                "box_uniform": bool(info[4] & 4), "box_geo": bool(info[4] & 8), "patterns": int(info[5]),
                "pattern_entries": int(info[6]), "kernel": int(info[7]),
                "pattern_irregular_rows": int(info[8]), "pattern_irregular_nnz": int(info[9])}

    def matvec(self, alpha: float, x, beta: float, y, b=None):
        """HYPRE_ParCSRMatrixMatvec / hypre_ParCSRMatrixMatvecOutOfPlace: y = alpha*A*x + beta*b
        (b defaults to y).  torch CUDA tensors -> device path; numpy arrays -> host path."""
        if isinstance(x, np.ndarray):
            assert b is None, "the host entry point is the in-place form"
            check(lib.hb200_parcsr_matvec_host(self.handle, alpha, _ptr(x), beta, _ptr(y)))
            return y
        _torch_sync(x, y, b)
        bb = y if b is None else b
        check(lib.hb200_parcsr_matvec(self.handle, alpha, _ptr(x), beta, _ptr(bb), _ptr(y)))
        sync()
        return y

    def matvecT(self, alpha: float, x, beta: float, y):
        """HYPRE_ParCSRMatrixMatvecT: y = alpha*A^T*x + beta*y"""
        _torch_sync(x, y)
        check(lib.hb200_parcsr_matvecT(self.handle, alpha, _ptr(x), beta, _ptr(y)))
        sync()
        return y

    def download_maps(self) -> dict:
        """Device -> host round trip of every integer map (bit-exactness check)."""
        out = {
            "diag_i": np.zeros(self.num_rows + 1, np.int32),
            "diag_j": np.zeros(self.diag_nnz, np.int32),
            "offd_i": np.zeros(self.num_rows + 1, np.int32) if self.num_cols_offd else None,
            "offd_j": np.zeros(self.offd_nnz, np.int32) if self.num_cols_offd else None,
            "col_map_offd": np.zeros(self.num_cols_offd, np.int64) if self.num_cols_offd else None,
            "send_map_starts": np.zeros(self.num_sends + 1, np.int32),
            "send_map_elmts": np.zeros(self.n_send_elmts, np.int32) if self.n_send_elmts else None,
            "recv_vec_starts": np.zeros(self.num_recvs + 1, np.int32),
            "send_procs": np.zeros(self.num_sends, np.int32) if self.num_sends else None,
            "recv_procs": np.zeros(self.num_recvs, np.int32) if self.num_recvs else None,
        }
        order = ["diag_i", "diag_j", "offd_i", "offd_j", "col_map_offd", "send_map_starts",
                 "send_map_elmts", "recv_vec_starts", "send_procs", "recv_procs"]
        check(lib.hb200_parcsr_download_maps(self.handle, *[_ptr(out[k]) for k in order]))
        return out


def init_required() -> None:
    if not _initialized:
        init(0)


class BoomerAMG:
    """Solve-phase mirror of HYPRE_BoomerAMG: a device-resident hierarchy + cycle parameters.

    levels: sequence of dicts with keys A (ParCSRMatrix), P (ParCSRMatrix or None),
    l1_norms, cf_marker (numpy or None), relax_weight, omega, cheby_ds, cheby_coefs.
    """

    def __init__(self, levels: Sequence[dict], num_grid_sweeps=(1, 1, 1, 1),
                 grid_relax_type=(18, 18, 18, 9), relax_order: int = 0, cycle_type: int = 1,
                 fcycle: int = 0, cheby_order: int = 2, cheby_scale: int = 1,
                 cheby_variant: int = 0, user_relax_type: int = -1, tol: float = 0.0,
                 min_iter: int = 0, max_iter: int = 1, converge_type: int = 0,
                 coarse_ge: Optional[dict] = None, use_graph: bool = False):
        init_required()
        self.levels = list(levels)
        h = C.c_void_p()
        check(lib.hb200_amg_create(C.byref(h), len(self.levels)))
        self.handle = h
        for l, L in enumerate(self.levels):
            P = L.get("P")
            l1 = _np(L.get("l1_norms"), np.float64)
            cf = _np(L.get("cf_marker"), np.int32)
            check(lib.hb200_amg_set_level(h, l, L["A"].handle, P.handle if P is not None else None,
                                          _ptr(l1), _ptr(cf), float(L.get("relax_weight", 1.0)),
                                          float(L.get("omega", 1.0))))
            coefs = _np(L.get("cheby_coefs"), np.float64)
            if coefs is not None:
                ds = _np(L.get("cheby_ds"), np.float64)
                check(lib.hb200_amg_set_level_cheby(h, l, _ptr(ds), _ptr(coefs), int(coefs.shape[0]) - 1))
        ngs = np.ascontiguousarray(num_grid_sweeps, np.int32)
        grt = np.ascontiguousarray(grid_relax_type, np.int32)
        check(lib.hb200_amg_set_cycle(h, _ptr(ngs), _ptr(grt), relax_order, cycle_type, fcycle,
                                      cheby_order, cheby_scale, cheby_variant, user_relax_type))
        check(lib.hb200_amg_set_solve(h, tol, min_iter, max_iter, converge_type))
        if coarse_ge is not None:
            amat = _np(coarse_ge["A_mat"], np.float64)
            check(lib.hb200_amg_set_coarse_ge(h, _ptr(amat), int(coarse_ge["n"]),
                                              int(coarse_ge["first_row"]), int(coarse_ge["num_local"])))
        if use_graph:
            check(lib.hb200_amg_set_use_graph(h, 1))
        self.grid_relax_type = tuple(int(v) for v in grt)
        self.num_grid_sweeps = tuple(int(v) for v in ngs)

    def set_use_graph(self, enable: bool) -> None:
        check(lib.hb200_amg_set_use_graph(self.handle, 1 if enable else 0))

    def set_solve(self, tol: float, min_iter: int, max_iter: int, converge_type: int = 0) -> None:
        check(lib.hb200_amg_set_solve(self.handle, tol, min_iter, max_iter, converge_type))

    def cycle(self, f, u, u_all_zeros: bool = False):
        """hypre_BoomerAMGCycle on (F_array[0], U_array[0]) = (f, u)."""
        _torch_sync(f, u)
        check(lib.hb200_amg_cycle(self.handle, _ptr(f), _ptr(u), 1 if u_all_zeros else 0))
        sync()
        return u

    def solve(self, f, u, u_all_zeros: bool = False):
        """HYPRE_BoomerAMGSolve: returns (num_iterations, final relative residual)."""
        _torch_sync(f, u)
        its, rr = C.c_int(0), C.c_double(0.0)
        flag = lib.hb200_amg_solve(self.handle, _ptr(f), _ptr(u), 1 if u_all_zeros else 0,
                                   C.byref(its), C.byref(rr))
        check(flag, allow_conv=True)
        sync()
        return its.value, rr.value

    def level_vector(self, level: int, which: int):
        """Device pointer + length of F_array[level] (which=0) or U_array[level] (which=1)."""
        p, n = C.c_void_p(), C.c_int(0)
        check(lib.hb200_amg_level_vector(self.handle, level, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def save(self, dirname: str) -> None:
        """hb200_amg_save: the hierarchy as files (level matrices in hypre's binary IJ format + one record per rank)"""
        check(lib.hb200_amg_save(self.handle, dirname.encode()))

    @classmethod
    def load(cls, dirname: str) -> "BoomerAMG":
        """hb200_amg_load: the hierarchy hb200_amg_save wrote, in a process without hypre; `mats[l] = (A_l, P_l)` are
        borrowed handles (mats[0][0] is the operator the Krylov solvers take)"""
        init_required()
        h = C.c_void_p()
        check(lib.hb200_amg_load(C.byref(h), dirname.encode()))
        self = cls.__new__(cls)
        self.handle = h
        self.levels = []
        self.mats = []
        for l in range(int(lib.hb200_amg_num_levels(h))):
            pair = []
            for which in (0, 1):
                m = C.c_void_p()
                check(lib.hb200_amg_level_matrix(h, l, which, C.byref(m)))
                if m.value:
                    M = ParCSRMatrix._adopt(m)
                    M._borrowed = True
                    pair.append(M)
                else:
                    pair.append(None)
            self.mats.append(tuple(pair))
        return self

    def destroy(self) -> None:
        if getattr(self, "handle", None):
            lib.hb200_amg_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


PRECOND_NONE, PRECOND_AMG, PRECOND_DIAGSCALE = 0, 1, 2


class _Krylov:
    def __init__(self):
        init_required()
        self.precond_kind = PRECOND_NONE
        self.precond = None
        self.result: Optional[KrylovResult] = None
        self.norms: Optional[np.ndarray] = None
        self.rel_norms: Optional[np.ndarray] = None

    def set_precond(self, precond) -> None:
        """HYPRE_PCGSetPrecond / HYPRE_GMRESSetPrecond: a BoomerAMG object, "diagscale" or None."""
        if precond is None:
            self.precond_kind, self.precond = PRECOND_NONE, None
        elif isinstance(precond, str) and precond == "diagscale":
            self.precond_kind, self.precond = PRECOND_DIAGSCALE, None
        else:
            self.precond_kind, self.precond = PRECOND_AMG, precond

    # HYPRE_*GetNumIterations / GetFinalRelativeResidualNorm / GetConverged
    @property
    def num_iterations(self) -> int:
        return int(self.result.num_iterations)

    @property
    def final_relative_residual_norm(self) -> float:
        return float(self.result.rel_residual_norm)

    @property
    def converged(self) -> int:
        return int(self.result.converged)


class ParCSRPCG(_Krylov):
    """HYPRE_ParCSRPCG: hypre_PCGSolve over the ParCSR function table."""

    def __init__(self, tol: float = 1e-6, max_iter: int = 1000, two_norm: int = 0,
                 rel_change: int = 0, flex: int = 0, recompute_residual: int = 0,
                 recompute_residual_p: int = 0, a_tol: float = 0.0, rtol: float = 0.0,
                 cf_tol: float = 0.0, atolf: float = 0.0, stop_crit: int = 0,
                 skip_break: int = 0, logging: int = 1, print_level: int = 0):
        super().__init__()
        p = PCGParams()
        lib.hb200_pcg_default_params(C.byref(p))
        p.tol, p.max_iter, p.two_norm, p.rel_change, p.flex = tol, max_iter, two_norm, rel_change, flex
        p.recompute_residual, p.recompute_residual_p = recompute_residual, recompute_residual_p
        p.a_tol, p.rtol, p.cf_tol, p.atolf, p.stop_crit, p.skip_break = a_tol, rtol, cf_tol, atolf, stop_crit, skip_break
        p.logging, p.print_level = logging, print_level
        self.params = p

    def solve(self, A: ParCSRMatrix, b, x) -> KrylovResult:
        """HYPRE_PCGSolve(solver, A, b, x).  torch CUDA tensors: device-resident solve; numpy
        arrays: the host-buffer entry point (H2D of b, x0 and D2H of x inside)."""
        n = self.params.max_iter + 2
        self.norms = np.zeros(n)
        self.rel_norms = np.zeros(n)
        res = KrylovResult()
        amg = self.precond.handle if self.precond_kind == PRECOND_AMG else None
        host = isinstance(b, np.ndarray)
        fn = lib.hb200_pcg_solve_host if host else lib.hb200_pcg_solve
        if not host:
            _torch_sync(b, x)
        flag = fn(A.handle, self.precond_kind, amg, C.byref(self.params), _ptr(b), _ptr(x),
                  _ptr(self.norms), _ptr(self.rel_norms), C.byref(res))
        check(flag, allow_conv=True)
        self.result = res
        return res


class ParCSRGMRES(_Krylov):
    """HYPRE_ParCSRGMRES: hypre_GMRESSolve over the ParCSR function table."""

    def __init__(self, tol: float = 1e-6, max_iter: int = 1000, k_dim: int = 5, a_tol: float = 0.0,
                 min_iter: int = 0, rel_change: int = 0, skip_real_r_check: int = 0,
                 cf_tol: float = 0.0, logging: int = 1, print_level: int = 0):
        super().__init__()
        p = GMRESParams()
        lib.hb200_gmres_default_params(C.byref(p))
        p.tol, p.max_iter, p.k_dim, p.a_tol, p.min_iter = tol, max_iter, k_dim, a_tol, min_iter
        p.rel_change, p.skip_real_r_check, p.cf_tol = rel_change, skip_real_r_check, cf_tol
        p.logging, p.print_level = logging, print_level
        self.params = p

    def solve(self, A: ParCSRMatrix, b, x) -> KrylovResult:
        self.norms = np.zeros(self.params.max_iter + 2)
        res = KrylovResult()
        amg = self.precond.handle if self.precond_kind == PRECOND_AMG else None
        host = isinstance(b, np.ndarray)
        fn = lib.hb200_gmres_solve_host if host else lib.hb200_gmres_solve
        if not host:
            _torch_sync(b, x)
        flag = fn(A.handle, self.precond_kind, amg, C.byref(self.params), _ptr(b), _ptr(x),
                  _ptr(self.norms), C.byref(res))
        check(flag, allow_conv=True)
        self.result = res
        return res


class ParCSRFlexGMRES(ParCSRGMRES):
    """HYPRE_ParCSRFlexGMRES: hypre_FlexGMRESSolve (src/krylov/flexgmres.c:288) over the ParCSR function
    table, default modify_pc."""
    _dev, _host = "hb200_flexgmres_solve", "hb200_flexgmres_solve_host"

    def __init__(self, tol: float = 1e-6, max_iter: int = 1000, k_dim: int = 20, a_tol: float = 0.0,
                 min_iter: int = 0, cf_tol: float = 0.0, logging: int = 1, print_level: int = 0):
        # hypre_FlexGMRESCreate: k_dim 20 (flexgmres.c:88)
        super().__init__(tol=tol, max_iter=max_iter, k_dim=k_dim, a_tol=a_tol, min_iter=min_iter,
                         cf_tol=cf_tol, logging=logging, print_level=print_level)

    def solve(self, A: ParCSRMatrix, b, x) -> KrylovResult:
        self.norms = np.zeros(self.params.max_iter + 2)
        res = KrylovResult()
        amg = self.precond.handle if self.precond_kind == PRECOND_AMG else None
        host = isinstance(b, np.ndarray)
        fn = getattr(lib, self._host if host else self._dev)
        if not host:
            _torch_sync(b, x)
        flag = fn(A.handle, self.precond_kind, amg, C.byref(self.params), _ptr(b), _ptr(x),
                  _ptr(self.norms), C.byref(res))
        check(flag, allow_conv=True)
        self.result = res
        return res


class ParCSRCOGMRES(ParCSRFlexGMRES):
    """HYPRE_ParCSRCOGMRES: hypre_COGMRESSolve (src/krylov/cogmres.c:270): classical Gram-Schmidt over the
    batched inner products / updates; cgs = 2 re-orthogonalises."""
    _dev, _host = "hb200_cogmres_solve", "hb200_cogmres_solve_host"

    def __init__(self, tol: float = 1e-6, max_iter: int = 1000, k_dim: int = 5, a_tol: float = 0.0,
                 min_iter: int = 0, rel_change: int = 0, skip_real_r_check: int = 0, cgs: int = 1,
                 cf_tol: float = 0.0, logging: int = 1, print_level: int = 0):
        ParCSRGMRES.__init__(self, tol=tol, max_iter=max_iter, k_dim=k_dim, a_tol=a_tol, min_iter=min_iter,
                             rel_change=rel_change, skip_real_r_check=skip_real_r_check, cf_tol=cf_tol,
                             logging=logging, print_level=print_level)
        self.params.cgs = cgs


class ParCSRLGMRES(ParCSRFlexGMRES):
    """HYPRE_ParCSRLGMRES: hypre_LGMRESSolve (src/krylov/lgmres.c:320): GMRES(k_dim) augmented with up to aug_dim
    error approximations of the previous restart cycles."""
    _dev, _host = "hb200_lgmres_solve", "hb200_lgmres_solve_host"

    def __init__(self, tol: float = 1e-6, max_iter: int = 1000, k_dim: int = 20, aug_dim: int = 2, a_tol: float = 0.0,
                 min_iter: int = 0, cf_tol: float = 0.0, logging: int = 1, print_level: int = 0):
        ParCSRGMRES.__init__(self, tol=tol, max_iter=max_iter, k_dim=k_dim, a_tol=a_tol, min_iter=min_iter,
                             cf_tol=cf_tol, logging=logging, print_level=print_level)
        self.params.aug_dim = aug_dim


class ParCSRBiCGSTAB(_Krylov):
    """HYPRE_ParCSRBiCGSTAB: hypre_BiCGSTABSolve (src/krylov/bicgstab.c:246) over the ParCSR function table."""

    def __init__(self, tol: float = 1e-6, max_iter: int = 1000, a_tol: float = 0.0, min_iter: int = 0,
                 cf_tol: float = 0.0, stop_crit: int = 0, logging: int = 1, print_level: int = 0):
        super().__init__()
        p = BiCGSTABParams()
        lib.hb200_bicgstab_default_params(C.byref(p))
        p.tol, p.max_iter, p.a_tol, p.min_iter, p.cf_tol, p.stop_crit = tol, max_iter, a_tol, min_iter, cf_tol, stop_crit
        p.logging, p.print_level = logging, print_level
        self.params = p

    def solve(self, A: ParCSRMatrix, b, x) -> KrylovResult:
        self.norms = np.zeros(self.params.max_iter + 2)
        res = KrylovResult()
        amg = self.precond.handle if self.precond_kind == PRECOND_AMG else None
        host = isinstance(b, np.ndarray)
        fn = lib.hb200_bicgstab_solve_host if host else lib.hb200_bicgstab_solve
        if not host:
            _torch_sync(b, x)
        flag = fn(A.handle, self.precond_kind, amg, C.byref(self.params), _ptr(b), _ptr(x),
                  _ptr(self.norms), C.byref(res))
        check(flag, allow_conv=True)
        self.result = res
        return res


def vector_print_ij(x, jlower: int, filename: str) -> None:
    """HYPRE_IJVectorPrint of a device vector: `<filename>.<5-digit rank>`"""
    _torch_sync(x)
    check(lib.hb200_vector_print_ij(_ptr(x), int(jlower), int(x.shape[0]), filename.encode()))


def vector_read_ij(filename: str):
    """HYPRE_IJVectorRead into a new device vector; returns (jlower, tensor)"""
    import torch
    lo, hi = C.c_int64(0), C.c_int64(-1)
    check(lib.hb200_vector_read_ij(filename.encode(), C.byref(lo), C.byref(hi), None, 0))
    n = int(hi.value - lo.value + 1)
    x = torch.zeros(max(n, 0), dtype=torch.float64, device="cuda")
    check(lib.hb200_vector_read_ij(filename.encode(), C.byref(lo), C.byref(hi), _ptr(x), n))
    sync()
    return int(lo.value), x


def relax(A: ParCSRMatrix, f, u, relax_type: int, relax_points: int = 0, relax_weight: float = 1.0,
          omega: float = 1.0, l1_norms=None, cf_marker=None, u_all_zeros: bool = False, vtemp=None):
    """hypre_BoomerAMGRelax on device vectors (l1_norms / cf_marker: torch CUDA tensors)."""
    import torch
    if vtemp is None:
        vtemp = torch.empty_like(u)
    _torch_sync(f, u, vtemp)
    check(lib.hb200_relax(A.handle, _ptr(f), _ptr(cf_marker), relax_type, relax_points,
                          relax_weight, omega, _ptr(l1_norms), _ptr(u), 1 if u_all_zeros else 0,
                          _ptr(vtemp)))
    sync()
    return u


def cheby_solve(A: ParCSRMatrix, f, u, coefs, order: int, scale: int, ds=None, variant: int = 0):
    """hypre_ParCSRRelax_Cheby_Solve on device vectors."""
    c = np.ascontiguousarray(coefs, np.float64)
    _torch_sync(f, u, ds)
    check(lib.hb200_cheby_solve(A.handle, _ptr(f), _ptr(ds), _ptr(c), order, scale, variant, _ptr(u)))
    sync()
    return u


def inner_prod(x, y) -> float:
    """hypre_ParVectorInnerProd (local dot + all-reduce)."""
    _torch_sync(x, y)
    out = C.c_double(0.0)
    check(lib.hb200_vec_inner_prod(_ptr(x), _ptr(y), x.numel(), C.byref(out)))
    return out.value


def axpy(alpha: float, x, y):
    """hypre_ParVectorAxpy: y += alpha*x"""
    _torch_sync(x, y)
    check(lib.hb200_vec_axpy(alpha, _ptr(x), _ptr(y), x.numel()))
    sync()
    return y


def amg_from_hierarchy(h, use_graph: bool = False, gs_chunks: int = 0):
    """Upload a hierarchy description (any object with the fields below, e.g. the reference
    bridge's view of hypre_ParAMGData after the reference's own BoomerAMGSetup) and return
    (list of level matrices, BoomerAMG).  Fields: levels = [dict(A=view, P=view|None,
    l1_norms, cf_marker, relax_weight, omega, cheby_ds, cheby_coefs)], params dict,
    coarse_ge dict|None.  gs_chunks: the thread count the reference's setup ran with when the smoothers are
    hybrid Gauss-Seidel (its sweeps and its l1 norms depend on it, ParCSRMatrix.set_gs_chunks)."""
    levels = []
    mats = []
    for L in h["levels"]:
        A = ParCSRMatrix.from_view(L["A"])
        P = ParCSRMatrix.from_view(L["P"]) if L.get("P") is not None else None
        if gs_chunks > 1:
            A.set_gs_chunks(gs_chunks)
        mats.append((A, P))
        d = dict(L)
        d["A"], d["P"] = A, P
        levels.append(d)
    p = h["params"]
    amg = BoomerAMG(levels, num_grid_sweeps=p["num_grid_sweeps"], grid_relax_type=p["grid_relax_type"],
                    relax_order=p["relax_order"], cycle_type=p["cycle_type"], fcycle=p["fcycle"],
                    cheby_order=p["cheby_order"], cheby_scale=p["cheby_scale"],
                    cheby_variant=p["cheby_variant"], user_relax_type=p["user_relax_type"],
                    tol=p["tol"], min_iter=p["min_iter"], max_iter=p["max_iter"],
                    converge_type=p["converge_type"], coarse_ge=h.get("coarse_ge"),
                    use_graph=use_graph)
    return mats, amg


__all__ = [
    "init", "finalize", "comm_init", "comm_get_unique_id", "sync", "launch_count",
    "ParCSRMatrix", "BoomerAMG", "ParCSRPCG", "ParCSRGMRES", "ParCSRFlexGMRES", "ParCSRCOGMRES",
    "ParCSRLGMRES", "ParCSRBiCGSTAB", "relax", "cheby_solve", "vector_print_ij", "vector_read_ij",
    "inner_prod", "axpy", "amg_from_hierarchy", "HB200Error", "KrylovResult",
    "PRECOND_NONE", "PRECOND_AMG", "PRECOND_DIAGSCALE",
]

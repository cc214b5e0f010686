"""ctypes binding of libhb200.so (include/hb200.h).

The library is the product; there is no Python or CPU fallback.  Importing this module
without the built CUDA library raises, and every entry point returns hypre-style error
flags that `check()` turns into exceptions.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhb200.so")

c_int_p = C.POINTER(C.c_int)
c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)


class HB200Error(RuntimeError):
    def __init__(self, flag: int, msg: str):
        super().__init__(f"hb200 error flag {flag}: {msg}")
        self.flag = flag


class PCGParams(C.Structure):
    _fields_ = [
        ("tol", C.c_double), ("a_tol", C.c_double), ("atolf", C.c_double),
        ("cf_tol", C.c_double), ("rtol", C.c_double),
        ("max_iter", C.c_int), ("two_norm", C.c_int), ("rel_change", C.c_int),
        ("recompute_residual", C.c_int), ("recompute_residual_p", C.c_int),
        ("stop_crit", C.c_int), ("skip_break", C.c_int), ("flex", C.c_int), ("hybrid", C.c_int),
        ("logging", C.c_int), ("print_level", C.c_int),
    ]


class GMRESParams(C.Structure):
    _fields_ = [
        ("tol", C.c_double), ("a_tol", C.c_double), ("cf_tol", C.c_double),
        ("k_dim", C.c_int), ("min_iter", C.c_int), ("max_iter", C.c_int),
        ("rel_change", C.c_int), ("skip_real_r_check", C.c_int), ("stop_crit", C.c_int),
        ("hybrid", C.c_int), ("logging", C.c_int), ("print_level", C.c_int),
        ("cgs", C.c_int), ("unroll", C.c_int), ("aug_dim", C.c_int), ("approx_constant", C.c_int),
    ]


class BiCGSTABParams(C.Structure):
    _fields_ = [
        ("tol", C.c_double), ("a_tol", C.c_double), ("cf_tol", C.c_double),
        ("min_iter", C.c_int), ("max_iter", C.c_int), ("stop_crit", C.c_int), ("hybrid", C.c_int),
        ("logging", C.c_int), ("print_level", C.c_int),
    ]


class KrylovResult(C.Structure):
    _fields_ = [
        ("num_iterations", C.c_int), ("converged", C.c_int),
        ("rel_residual_norm", C.c_double), ("error_flag", C.c_int),
        ("solve_ms", C.c_double), ("kernel_launches", C.c_longlong),
    ]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m hypre_b200.build` "
            "(hypre_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp = C.c_void_p
    sigs = {
        "hb200_init": ([C.c_int], C.c_int),
        "hb200_finalize": ([], C.c_int),
        "hb200_last_error": ([], C.c_char_p),
        "hb200_version": ([], C.c_char_p),
        "hb200_comm_get_unique_id": ([vp], C.c_int),
        "hb200_comm_init": ([C.c_int, C.c_int, vp], C.c_int),
        "hb200_comm_rank": ([], C.c_int),
        "hb200_comm_size": ([], C.c_int),
        "hb200_comm_barrier": ([], C.c_int),
        "hb200_set_halo_mode": ([C.c_int], C.c_int),
        "hb200_halo_mode": ([], C.c_int),
        "hb200_malloc": ([C.POINTER(vp), C.c_size_t], C.c_int),
        "hb200_free": ([vp], C.c_int),
        "hb200_memcpy_h2d": ([vp, vp, C.c_size_t], C.c_int),
        "hb200_memcpy_d2h": ([vp, vp, C.c_size_t], C.c_int),
        "hb200_memcpy_d2d": ([vp, vp, C.c_size_t], C.c_int),
        "hb200_sync": ([], C.c_int),
        "hb200_compute_stream": ([], vp),
        "hb200_launch_count": ([C.c_int], C.c_longlong),
        "hb200_parcsr_create": ([C.POINTER(vp), C.c_int, C.c_int, C.c_int,
                                 vp, vp, vp, vp, vp, vp, vp,
                                 C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                 C.c_int, vp, vp, vp, C.c_int, vp, vp], C.c_int),
        "hb200_parcsr_destroy": ([vp], C.c_int),
        "hb200_parcsr_num_rows": ([vp], C.c_int),
        "hb200_parcsr_num_cols": ([vp], C.c_int),
        "hb200_parcsr_num_nonzeros": ([vp], C.c_longlong),
        "hb200_parcsr_download_maps": ([vp] + [vp] * 10, C.c_int),
        "hb200_parcsr_set_spmv_kernel": ([vp, C.c_int, C.c_int], C.c_int),
        "hb200_parcsr_format_info": ([vp, C.POINTER(C.c_longlong)], C.c_int),
        "hb200_host_csr_transpose": ([C.c_int, C.c_int, vp, vp, vp, vp, vp, vp], C.c_int),
        "hb200_host_gs_schedule": ([C.c_int, vp, vp, C.c_int, vp, vp, c_int_p], C.c_int),
        "hb200_host_pattern_analyze_wide": ([C.c_int, C.c_int, vp, vp, vp, vp, vp, c_int_p, c_int_p, C.c_int, vp, vp, vp, c_int_p, vp], C.c_int),
        "hb200_host_pattern_analyze": ([C.c_int, C.c_int, vp, vp, vp, vp, vp, c_int_p, vp, vp, vp, c_int_p, vp], C.c_int),
        "hb200_parcsr_matvec": ([vp, C.c_double, vp, C.c_double, vp, vp], C.c_int),
        "hb200_parcsr_matvecT": ([vp, C.c_double, vp, C.c_double, vp], C.c_int),
        "hb200_parcsr_matvec_host": ([vp, C.c_double, vp, C.c_double, vp], C.c_int),
        "hb200_vec_set": ([vp, C.c_double, C.c_size_t], C.c_int),
        "hb200_vec_copy": ([vp, vp, C.c_size_t], C.c_int),
        "hb200_vec_scale": ([C.c_double, vp, C.c_size_t], C.c_int),
        "hb200_vec_axpy": ([C.c_double, vp, vp, C.c_size_t], C.c_int),
        "hb200_vec_inner_prod": ([vp, vp, C.c_size_t, c_double_p], C.c_int),
        "hb200_vec_pointwise_divpy": ([vp, vp, vp, vp, C.c_int, C.c_size_t], C.c_int),
        "hb200_relax": ([vp, vp, vp, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp, C.c_int, vp], C.c_int),
        "hb200_relax_if": ([vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp, C.c_int, vp], C.c_int),
        "hb200_cheby_solve": ([vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp], C.c_int),
        "hb200_amg_create": ([C.POINTER(vp), C.c_int], C.c_int),
        "hb200_amg_destroy": ([vp], C.c_int),
        "hb200_parcsr_set_gs_chunks": ([vp, C.c_int], C.c_int),
        "hb200_gs_auto_chunks": ([C.c_int], C.c_int),
        "hb200_amg_set_level": ([vp, C.c_int, vp, vp, vp, vp, C.c_double, C.c_double], C.c_int),
        "hb200_amg_set_level_cheby": ([vp, C.c_int, vp, vp, C.c_int], C.c_int),
        "hb200_amg_set_cycle": ([vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int], C.c_int),
        "hb200_amg_set_level_weights": ([vp, C.c_int, C.c_double, C.c_double], C.c_int),
        "hb200_amg_set_solve": ([vp, C.c_double, C.c_int, C.c_int, C.c_int], C.c_int),
        "hb200_amg_set_coarse_ge": ([vp, vp, C.c_int, C.c_int, C.c_int], C.c_int),
        "hb200_amg_set_use_graph": ([vp, C.c_int], C.c_int),
        "hb200_amg_cycle": ([vp, vp, vp, C.c_int], C.c_int),
        "hb200_amg_solve": ([vp, vp, vp, C.c_int, c_int_p, c_double_p], C.c_int),
        "hb200_amg_save": ([vp, C.c_char_p], C.c_int),
        "hb200_amg_load": ([C.POINTER(vp), C.c_char_p], C.c_int),
        "hb200_amg_level_matrix": ([vp, C.c_int, C.c_int, C.POINTER(vp)], C.c_int),
        "hb200_amg_num_levels": ([vp], C.c_int),
        "hb200_amg_solve_logged": ([vp, vp, vp, C.c_int, c_int_p, c_double_p, vp, c_double_p], C.c_int),
        "hb200_amg_cycle_sweeps": ([vp, vp], C.c_int),
        "hb200_amg_level_vector": ([vp, C.c_int, C.c_int, C.POINTER(vp), c_int_p], C.c_int),
        "hb200_pcg_default_params": ([C.POINTER(PCGParams)], None),
        "hb200_pcg_solve": ([vp, C.c_int, vp, C.POINTER(PCGParams), vp, vp, vp, vp, C.POINTER(KrylovResult)], C.c_int),
        "hb200_pcg_solve_host": ([vp, C.c_int, vp, C.POINTER(PCGParams), vp, vp, vp, vp, C.POINTER(KrylovResult)], C.c_int),
        "hb200_krylov_warmup": ([vp, C.c_int, vp, C.c_int, C.c_int], C.c_int),
        "hb200_gmres_default_params": ([C.POINTER(GMRESParams)], None),
        "hb200_gmres_solve": ([vp, C.c_int, vp, C.POINTER(GMRESParams), vp, vp, vp, C.POINTER(KrylovResult)], C.c_int),
        "hb200_gmres_solve_host": ([vp, C.c_int, vp, C.POINTER(GMRESParams), vp, vp, vp, C.POINTER(KrylovResult)], C.c_int),
        "hb200_flexgmres_solve": ([vp, C.c_int, vp, C.POINTER(GMRESParams), vp, vp, vp, C.POINTER(KrylovResult)], C.c_int),
        "hb200_flexgmres_solve_host": ([vp, C.c_int, vp, C.POINTER(GMRESParams), vp, vp, vp, C.POINTER(KrylovResult)], C.c_int),
        "hb200_cogmres_solve": ([vp, C.c_int, vp, C.POINTER(GMRESParams), vp, vp, vp, C.POINTER(KrylovResult)], C.c_int),
        "hb200_cogmres_solve_host": ([vp, C.c_int, vp, C.POINTER(GMRESParams), vp, vp, vp, C.POINTER(KrylovResult)], C.c_int),
        "hb200_lgmres_solve": ([vp, C.c_int, vp, C.POINTER(GMRESParams), vp, vp, vp, C.POINTER(KrylovResult)], C.c_int),
        "hb200_lgmres_solve_host": ([vp, C.c_int, vp, C.POINTER(GMRESParams), vp, vp, vp, C.POINTER(KrylovResult)], C.c_int),
        "hb200_bicgstab_default_params": ([C.POINTER(BiCGSTABParams)], None),
        "hb200_bicgstab_solve": ([vp, C.c_int, vp, C.POINTER(BiCGSTABParams), vp, vp, vp, C.POINTER(KrylovResult)], C.c_int),
        "hb200_bicgstab_solve_host": ([vp, C.c_int, vp, C.POINTER(BiCGSTABParams), vp, vp, vp, C.POINTER(KrylovResult)], C.c_int),
        "hb200_parcsr_from_ij": ([C.POINTER(vp)] + [C.c_int64] * 5 + [vp, vp, vp, C.c_int], C.c_int),
        "hb200_parcsr_read_ij": ([C.POINTER(vp), C.c_char_p, C.c_int], C.c_int),
        "hb200_parcsr_info": ([vp, vp], C.c_int),
        "hb200_parcsr_print_ij": ([vp, C.c_char_p], C.c_int),
        "hb200_parcsr_print_ij_binary": ([vp, C.c_char_p], C.c_int),
        "hb200_vector_print_ij": ([vp, C.c_int64, C.c_int, C.c_char_p], C.c_int),
        "hb200_vector_read_ij": ([C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), vp, C.c_int], C.c_int),
        "hb200_host_ij_assemble": ([C.c_int64] * 5 + [vp, vp, vp, C.c_int, c_int_p, c_int_p, c_int_p] + [vp] * 7, C.c_int),
        "hb200_host_ij_commpkg": ([C.c_int, C.c_int, vp, vp, C.c_int64, c_int_p, vp, vp, vp, C.c_int, c_int_p, vp, vp], C.c_int),
    }
    for name, (argtypes, restype) in sigs.items():
        fn = getattr(lib, name)   # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = restype
    lib._hb200_symbols = sorted(sigs)
    return lib


lib = _load()
SYMBOLS = lib._hb200_symbols


def check(flag: int, allow_conv: bool = False) -> int:
    if flag == 0:
        return 0
    if allow_conv and (flag & ~256) == 0:
        return flag
    raise HB200Error(flag, (lib.hb200_last_error() or b"").decode(errors="replace"))

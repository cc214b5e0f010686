/*
 * hb200.h — C-ABI of libhb200.so: the B200-native (sm_100a CUDA + NCCL) BoomerAMG
 * *solve phase* inside PCG/GMRES.
 *
 * This is the drop-in boundary for ONE hot path of hypre 3.1.0 (paths below are relative
 * to the reference tree, /root/reference): plain pointers and sizes, no hypre types, no
 * torch types.  Each entry point names the reference interface it replaces.  The
 * reference-side binding (the C shim that overrides hypre's own symbols and forwards to
 * these calls) is hypre_b200/csrc/hypre_shim.c; INTEGRATION.md shows how a maintainer
 * links it.
 *
 * Conventions (mirroring the reference):
 *   - HYPRE_Int  = int32 (src/utilities/HYPRE_utilities.h:83-84); HYPRE_BigInt is passed
 *     as int64 here (the shim widens it when the reference is built without MIXEDINT).
 *   - HYPRE_Real = HYPRE_Complex = double (HYPRE_utilities.h:130,156).  All arithmetic fp64.
 *   - Every function returns an int error flag with hypre's bit meaning
 *     (src/utilities/HYPRE_utilities.h:273-277): 0 ok, 1 generic (incl. CUDA/NCCL
 *     failure), 2 memory, 4 argument, 256 convergence.  hb200_last_error() has the text.
 *   - "dev" pointers are device (HBM) pointers owned by the caller (hb200_malloc or any
 *     CUDA allocation, e.g. a torch tensor's data_ptr()); "host" pointers are host memory.
 *   - One process drives one GPU (src/utilities/device_utils.c:2764 hypre_bind_device_id);
 *     all ranks of a job call collectively, exactly like the MPI reference.
 *   - There is NO CPU fallback: every entry point fails with flag 1 when no sm_100 device
 *     or no CUDA runtime is usable.  The only calls that work without a GPU are the
 *     hb200_host_* helpers: the host halves of the upload (format analysis, stored transpose,
 *     GS schedule), exposed so that they can be checked on a CPU-only machine; they do none of
 *     the solve-path arithmetic.
 *   - What "no CPU fallback" covers.  THIS library (libhb200.so) never computes on the CPU.  The
 *     reference-side binding (libHYPRE_b200.so, hypre_shim.c) sits in front of a complete hypre and is
 *     a drop-in for ONE path: a call that is not this path at all — another Krylov function table
 *     (struct / sstruct), block-mode or additive BoomerAMG, Schwarz/ILU/Euclid smoothers, AIR, a
 *     matrix on a sub-communicator, multi-vectors, per-cycle printing — is handed back to the
 *     application's own hypre with a one-line notice on stderr, because refusing it would break
 *     applications that use those features next to the accelerated solve.  HYPRE_B200_STRICT=1
 *     turns every hand-back into a hypre error instead.  The decision is taken collectively (the
 *     same on every rank) and never depends on whether a GPU is present: a missing device or a CUDA /
 *     NCCL failure is always an error.
 *   - A polling halo kernel that loses its peer gives up after HB200_HALO_TIMEOUT_S seconds
 *     (default 30); the call in flight and every later one return flag 1 ("halo exchange timed
 *     out") until hb200_finalize.
 */
#ifndef HB200_H
#define HB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB200_ERROR_GENERIC 1
#define HB200_ERROR_MEMORY  2
#define HB200_ERROR_ARG     4
#define HB200_ERROR_CONV    256

typedef struct hb200_parcsr hb200_parcsr;   /* device-resident hypre_ParCSRMatrix + CommPkg */
typedef struct hb200_amg    hb200_amg;      /* device-resident hypre_ParAMGData (solve part) */

/* ------------------------------------------------------------------------------------ */
/* runtime                                                                                */
/* ------------------------------------------------------------------------------------ */

/* Binds this process to `device`, creates the compute + comm streams.  Replaces
 * hypre_bind_device_id / HYPRE_Init device part (src/utilities/general.c HYPRE_Init). */
int hb200_init(int device);
int hb200_finalize(void);
const char *hb200_last_error(void);
const char *hb200_version(void);

/* Multi-GPU: one NCCL communicator over all ranks replaces the MPI communicator of
 * hypre_ParCSRMatrixComm for the solve phase (src/utilities/mpistubs.c:940+ hot-path
 * wrappers Isend/Irecv/Waitall/Allreduce/Allgatherv).  Rank 0 calls get_unique_id, the
 * host program broadcasts the 128 bytes by any means (torch.distributed, MPI, a file),
 * then every rank calls comm_init. */
int hb200_comm_get_unique_id(void *id128);
int hb200_comm_init(int rank, int nranks, const void *id128);
int hb200_comm_rank(void);
int hb200_comm_size(void);
int hb200_comm_barrier(void);
/* halo transport (the Isend/Irecv/Waitall of hypre_ParCSRCommHandleCreate,
 * src/parcsr_mv/par_csr_communication.c:520-640): 0 = NCCL send/recv (default), 1 = direct
 * NVLink peer stores into IPC-mapped receive buffers (pack+put fused in one kernel; lets the
 * whole V-cycle stay a CUDA graph on N>1), 2 = mode 1 when every rank can map every peer,
 * else mode 0.  Collective over the ranks of hb200_comm_init for modes 1 and 2.
 * hb200_halo_mode returns the transport in force. */
int hb200_set_halo_mode(int mode);
int hb200_halo_mode(void);

/* device memory + transfers (hypre_TAlloc/hypre_TMemcpy, src/utilities/memory.c:956-990) */
int hb200_malloc(void **dev, size_t bytes);
int hb200_free(void *dev);
int hb200_memcpy_h2d(void *dev, const void *host, size_t bytes);
int hb200_memcpy_d2h(void *host, const void *dev, size_t bytes);
int hb200_memcpy_d2d(void *dst, const void *src, size_t bytes);
int hb200_sync(void);                 /* hypre_SyncComputeStream */
void *hb200_compute_stream(void);     /* cudaStream_t, for callers that time with events */
/* launch counter: number of kernels launched by this library since the last reset */
long long hb200_launch_count(int reset);

/* ------------------------------------------------------------------------------------ */
/* (a1,a2) ParCSR matrix + CommPkg                                                        */
/* ------------------------------------------------------------------------------------ */

/* Uploads one hypre_ParCSRMatrix (src/parcsr_mv/par_csr_matrix.h:27-92): diag and offd
 * hypre_CSRMatrix blocks (src/seq_mv/csr_matrix.h:33-62; local / compressed column
 * indices, first entry of every diag row = diagonal element as the reference's relaxation
 * assumes, par_relax.c:274), col_map_offd, the row/col ownership, and its
 * hypre_ParCSRCommPkg (src/parcsr_mv/par_csr_communication.h:51-75) by value.
 * All array arguments are HOST pointers and are copied; they may be freed on return.
 * For 1 rank pass num_cols_offd = 0, num_sends = num_recvs = 0 and NULL arrays. */
int hb200_parcsr_create(hb200_parcsr **A,
                        int num_rows, int num_cols, int num_cols_offd,
                        const int *diag_i, const int *diag_j, const double *diag_data,
                        const int *offd_i, const int *offd_j, const double *offd_data,
                        const int64_t *col_map_offd,
                        int64_t first_row_index, int64_t first_col_diag,
                        int64_t global_num_rows, int64_t global_num_cols,
                        int num_sends, const int *send_procs, const int *send_map_starts,
                        const int *send_map_elmts,
                        int num_recvs, const int *recv_procs, const int *recv_vec_starts);
int hb200_parcsr_destroy(hb200_parcsr *A);
int hb200_parcsr_num_rows(const hb200_parcsr *A);
int hb200_parcsr_num_cols(const hb200_parcsr *A);
long long hb200_parcsr_num_nonzeros(const hb200_parcsr *A);   /* local diag+offd */
/* Device -> host round trip of the integer maps (bit-exact parity check).
 * Any output pointer may be NULL.  Sizes are those given at creation. */
int hb200_parcsr_download_maps(const hb200_parcsr *A, int *diag_i, int *diag_j,
                               int *offd_i, int *offd_j, int64_t *col_map_offd,
                               int *send_map_starts, int *send_map_elmts,
                               int *recv_vec_starts, int *send_procs, int *recv_procs);
/* Storage format of the diag block (10 values): info[0] = 1 when a dictionary-packed SELL-32
 * copy exists (structured operators: <= 256 distinct column offsets), info[1] = its stored
 * entries incl. padding, info[2] = bytes per stored entry (2 = offset + value codes, 9 = offset
 * code + fp64 value), info[3] = number of distinct values (0 when values are stored raw);
 * info[4] bit 0 = a row-pattern copy exists (>= 70% of the rows fall into <= 255 distinct rows
 * as lists of (column - base, value): constant-coefficient stencils and the regular part of
 * their first coarse level; 1 byte per row), bit 1 = its patterns are compact 3 x 3 x 3 stencils (the
 * stencil-sweep kernel applies), bit 2 = all patterns are one reference pattern with slots missing,
 * bit 3 = the missing slots are exactly the neighbours outside the grid box (no row codes are read);
 * info[5] = patterns, info[6] = table entries;
 * info[7] = the kernel kind in force (numbering of hb200_parcsr_set_spmv_kernel);
 * info[8] = rows outside the pattern table (swept in CSR), info[9] = their nonzeros. */
int hb200_parcsr_format_info(const hb200_parcsr *A, long long *info10);
/* Selects the SpMV kernel for this matrix: 0 = auto (compact-stencil, else row-pattern, else packed
 * SELL, else vector with lanes from nnz/row; default), 1 = vector-per-row (sub-warp of K lanes: the
 * general CSR kernel), 2 = nnz-balanced stream, one thread per row adding in CSR order (the
 * order-preserving cross-check of the parity tests; lanes_per_row is ignored), 6 = packed SELL,
 * 7 = row-pattern, 8 = vector-per-row over 16-bit column offsets from the row (10 B per nonzero;
 * square blocks whose entries stay within +-32767 of the diagonal), 9 = the compact-stencil
 * (3 x 3 x 3 box) kernel over the row-pattern table.  6 - 9 fall back to the next format when the
 * block does not qualify; auto prefers 9, 7, 6, 8, 1 in that order.  (3, 4, 5 — the stream / unrolled
 * variants of the first release, slower on every level measured — are gone: argument error.) */
int hb200_parcsr_set_spmv_kernel(hb200_parcsr *A, int kind, int lanes_per_row);
/* Host-side half of the upload, no GPU needed: the row-pattern analysis of one CSR
 * block (a hypre_CSRMatrix: i / j / data, src/seq_mv/csr_matrix.h:33-62) exactly as
 * hb200_parcsr_create runs it.  Outputs (any may be NULL): row_code[num_rows] (pattern id, 255 =
 * row outside the table), row_base[num_rows] (column of entry k = row_base + pattern_offset[k];
 * the row itself for square blocks), pattern_ptr[*num_patterns + 1], pattern_offset / pattern_value
 * (<= 8192 entries), irregular_rows[*num_irregular].  *num_patterns = 0 when the block does not
 * qualify.  Decoding the outputs reproduces the block entry for entry (lossless). */
int hb200_host_pattern_analyze(int num_rows, int num_cols, const int *row_ptr, const int *col_ind,
                               const double *values, unsigned char *row_code, int *row_base,
                               int *num_patterns, int *pattern_ptr, int *pattern_offset,
                               double *pattern_value, int *num_irregular, int *irregular_rows);
/* Experimental wide variant of the same analysis (enabled in the upload by HB200_PAT_WIDE=1 for
 * blocks the 1-byte format rejects): 16-bit row codes (65535 = row outside the table), up to 65534
 * patterns and 2^20 table entries read from global memory; patterns need >= 4 rows.  The table is
 * copied out only when it fits pattern_capacity entries (*num_entries says how many it has). */
int hb200_host_pattern_analyze_wide(int num_rows, int num_cols, const int *row_ptr, const int *col_ind,
                                    const double *values, unsigned short *row_code, int *row_base,
                                    int *num_patterns, int *num_entries, int pattern_capacity,
                                    int *pattern_ptr, int *pattern_offset, double *pattern_value,
                                    int *num_irregular, int *irregular_rows);
/* The stored transpose that restriction runs on (hypre_ParCSRMatrixMatvecT with keepTranspose,
 * src/parcsr_mv/par_csr_matvec.c:298-299, 430-468): entries of each output row in ascending
 * source-row order = the order hypre_CSRMatrixMatvecT accumulates in (src/seq_mv/csr_matvec.c:1095-1110).
 * t_row_ptr[num_cols + 1], t_col_ind / t_values[nnz].  Host only, no GPU needed. */
int hb200_host_csr_transpose(int num_rows, int num_cols, const int *row_ptr, const int *col_ind,
                             const double *values, int *t_row_ptr, int *t_col_ind, double *t_values);
/* Hybrid Gauss-Seidel (relax types 3, 4, 6, 8, 13, 14, 88, 89): the reference's result depends on its thread
 * count (hypre_NumThreads(), src/parcsr_ls/par_relax.c:727, 868-896): rows are cut into that many chunks
 * (hypre_partition1D), Gauss-Seidel inside a chunk, Jacobi between chunks.  num_chunks <= 1 (default): the
 * 1-thread sweep, exactly (wavefront schedule, one launch per wavefront); num_chunks = T > 1: the sweep of the
 * reference at OMP_NUM_THREADS = T, one launch per call.  The l1 norms of the l1 variants must be the ones of
 * the same T (hypre_ParCSRComputeL1NormsThreads, src/parcsr_ls/ams.c:4523). */
int hb200_parcsr_set_gs_chunks(hb200_parcsr *A, int num_chunks);
/* the chunk count the device prefers for a block of num_rows rows (no GPU needed) */
int hb200_gs_auto_chunks(int num_rows);
/* Wavefront schedule of the hybrid Gauss-Seidel sweeps (src/parcsr_ls/par_relax.h:12-330 run with
 * one thread): rows of one level are mutually uncoupled and every row comes after the rows it
 * reads updated values from, so sweeping the levels in order reproduces the sequential sweep.
 * perm[num_rows] = rows grouped by level, level_ptr[*num_levels + 1].  Host only. */
int hb200_host_gs_schedule(int num_rows, const int *row_ptr, const int *col_ind, int forward,
                           int *perm, int *level_ptr, int *num_levels);

/* (a3) hypre_ParCSRMatrixMatvecOutOfPlace (src/parcsr_mv/par_csr_matvec.c:241-262):
 *   y = alpha*A*x + beta*b, halo exchange (job 1) overlapped with the diag block.
 * x: num_cols doubles, b,y: num_rows doubles (dev).  y may alias b, must not alias x. */
int hb200_parcsr_matvec(hb200_parcsr *A, double alpha, const double *x_dev,
                        double beta, const double *b_dev, double *y_dev);
/* (a4) hypre_ParCSRMatrixMatvecT (par_csr_matvec.c:523-546): y = alpha*A^T*x + beta*y,
 * reverse halo (job 2) + deterministic scatter-add into y[send_map_elmts]. */
int hb200_parcsr_matvecT(hb200_parcsr *A, double alpha, const double *x_dev,
                         double beta, double *y_dev);
/* HYPRE_ParCSRMatrixMatvec with HOST vectors (src/parcsr_mv/HYPRE_parcsr_matrix.c:385):
 * copies x (and y when beta != 0) to the device, runs the kernel, copies y back. */
int hb200_parcsr_matvec_host(hb200_parcsr *A, double alpha, const double *x_host,
                             double beta, double *y_host);

/* ------------------------------------------------------------------------------------ */
/* (a14) ParVector BLAS-1 (src/parcsr_mv/par_vector.c:416-627, src/seq_mv/vector.c)        */
/* ------------------------------------------------------------------------------------ */
int hb200_vec_set(double *y_dev, double value, size_t n);        /* SetConstantValues/SetZeros */
int hb200_vec_copy(const double *x_dev, double *y_dev, size_t n);  /* hypre_ParVectorCopy       */
int hb200_vec_scale(double alpha, double *y_dev, size_t n);      /* hypre_ParVectorScale        */
int hb200_vec_axpy(double alpha, const double *x_dev, double *y_dev, size_t n); /* ...Axpy      */
/* hypre_ParVectorInnerProd: local dot + allreduce(SUM) over all ranks (par_vector.c:600-627) */
int hb200_vec_inner_prod(const double *x_dev, const double *y_dev, size_t n, double *result);
/* hypre_ParVectorPointwiseDivpy[Marked]: y += x ./ b (vector.c:1083-1296); marker may be NULL */
int hb200_vec_pointwise_divpy(const double *x_dev, const double *b_dev, double *y_dev,
                              const int *marker_dev, int marker_val, size_t n);

/* ------------------------------------------------------------------------------------ */
/* (a9-a13) relaxation                                                                    */
/* ------------------------------------------------------------------------------------ */

/* hypre_BoomerAMGRelax (src/parcsr_ls/par_relax.c:23-173), same argument meaning.
 * relax_type: 0 weighted Jacobi, 7 Jacobi (matvec form), 18 l1-Jacobi,
 *             3/4/6 hybrid GS fwd/bwd/symmetric, 8/13/14/88/89 l1 hybrid GS variants.
 * cf_marker_dev / relax_points: CF ordering (par_relax.c:262); NULL / 0 = all points.
 * l1_norms_dev: required for 8/13/14/18/88/89.  vtemp_dev: num_rows doubles of scratch.
 * u_all_zeros: hypre_ParVectorAllZeros(u) (par_vector.h:40-41): legalises the SpMV-free
 * first Jacobi sweep (par_relax.c:1221-1228). */
int hb200_relax(hb200_parcsr *A, const double *f_dev, const int *cf_marker_dev,
                int relax_type, int relax_points, double relax_weight, double omega,
                const double *l1_norms_dev, double *u_dev, int u_all_zeros,
                double *vtemp_dev);
/* hypre_BoomerAMGRelaxIF (src/parcsr_ls/par_relax_interface.c:19-65): C/F ordered sweeps. */
int hb200_relax_if(hb200_parcsr *A, const double *f_dev, const int *cf_marker_dev,
                   int relax_type, int relax_order, int cycle_param, double relax_weight,
                   double omega, const double *l1_norms_dev, double *u_dev, int u_all_zeros,
                   double *vtemp_dev);
/* hypre_ParCSRRelax_Cheby_Solve (src/parcsr_ls/par_cheby_solve.c:352-390):
 * u += p(A) r, coefs[order+1] on the host, ds = D^{-1/2} (dev) when scale != 0. */
int hb200_cheby_solve(hb200_parcsr *A, const double *f_dev, const double *ds_dev,
                      const double *coefs_host, int order, int scale, int variant,
                      double *u_dev);

/* ------------------------------------------------------------------------------------ */
/* (a7,a8) BoomerAMG hierarchy + cycle                                                    */
/* ------------------------------------------------------------------------------------ */

/* Mirror of the solve-time part of hypre_ParAMGData (src/parcsr_ls/par_amg.h:18-310). */
int hb200_amg_create(hb200_amg **amg, int num_levels);
int hb200_amg_destroy(hb200_amg *amg);   /* does NOT destroy the level matrices */
/* level l: A_array[l]; P_array[l] (NULL on the coarsest level; R = P^T is applied through
 * a transpose built at upload, par_amg_setup.c:831); l1_norms[l] / cf_marker[l] are HOST
 * arrays of A's num_rows entries or NULL.  relax_weight[l], omega[l] (par_amg.h:66-67). */
int hb200_amg_set_level(hb200_amg *amg, int level, hb200_parcsr *A, hb200_parcsr *P,
                        const double *l1_norms_host, const int *cf_marker_host,
                        double relax_weight, double omega);
/* Chebyshev data of level l: ds (num_rows, host, may be NULL when scale == 0), coefs[order+1] */
int hb200_amg_set_level_cheby(hb200_amg *amg, int level, const double *ds_host,
                              const double *coefs_host, int order);
/* cycle parameters: num_grid_sweeps[4], grid_relax_type[4] (index = cycle_param: 1 down,
 * 2 up, 3 coarse; par_cycle.c:317-342), relax_order, cycle_type (1 V, 2 W), fcycle,
 * cheby_order/scale/variant, user_relax_type for 1-level hierarchies (par_cycle.c:349-356) */
int hb200_amg_set_cycle(hb200_amg *amg, const int *num_grid_sweeps4,
                        const int *grid_relax_type4, int relax_order, int cycle_type,
                        int fcycle, int cheby_order, int cheby_scale, int cheby_variant,
                        int user_relax_type);
/* relax_weight[l] / omega[l] of a level changed after the upload (HYPRE_BoomerAMGSetRelaxWt /
 * SetOuterWt / SetLevelRelaxWt between two solves: the reference's cycle reads them from
 * hypre_ParAMGData on every call, par_cycle.c:60-110).  A change drops the captured cycle graphs. */
int hb200_amg_set_level_weights(hb200_amg *amg, int level, double relax_weight, double omega);
/* solver parameters of hypre_BoomerAMGSolve (par_amg_solve.c:22): tol, min/max_iter,
 * converge_type.  As a preconditioner ij sets tol = 0, max_iter = 1 (ij.c:320,324). */
int hb200_amg_set_solve(hb200_amg *amg, double tol, int min_iter, int max_iter,
                        int converge_type);
/* coarsest-level direct solve, relax types 9/99 (src/parcsr_ls/par_gauss_elim.c:457):
 * A_mat is the dense n x n matrix exactly as hypre_GaussElimSetup stores it (row-major,
 * global n = coarsest global rows), first_row = this rank's first coarse row. */
int hb200_amg_set_coarse_ge(hb200_amg *amg, const double *A_mat_host, int n, int first_row,
                            int num_local_rows);
/* Capture the whole V-cycle in a CUDA graph (launch-latency bound coarse levels). */
int hb200_amg_set_use_graph(hb200_amg *amg, int enable);
/* hypre_BoomerAMGCycle (par_cycle.c:23): one cycle on F_array[0]=f, U_array[0]=u (dev).
 * u_all_zeros as in hb200_relax. */
int hb200_amg_cycle(hb200_amg *amg, const double *f_dev, double *u_dev, int u_all_zeros);
/* hypre_BoomerAMGSolve (par_amg_solve.c:22-424): cycles until tol / max_iter; outputs
 * may be NULL. */
int hb200_amg_solve(hb200_amg *amg, const double *f_dev, double *u_dev, int u_all_zeros,
                    int *num_iterations, double *rel_resid_norm);
/* The same solve with what the reference prints at print_level > 1 / keeps at logging > 1 (par_amg_solve.c:130-280):
 * resid_norms[0] = the residual norm before the first cycle, resid_norms[k] = after cycle k (max_iter + 1 values, host),
 * *rhs_norm = the norm of f (converge_type 0).  The norms are computed after every cycle even when tol == 0, as the
 * reference does in that mode. */
int hb200_amg_solve_logged(hb200_amg *amg, const double *f_dev, double *u_dev, int u_all_zeros,
                           int *num_iterations, double *rel_resid_norm, double *resid_norms, double *rhs_norm);
/* relaxation sweeps ONE cycle makes on every level (host only): the reference's "cycle complexity" adds the nonzeros of
 * a level once per sweep (par_cycle.c:455-474, hypre_ParAMGDataCycleOpCount) */
int hb200_amg_cycle_sweeps(const hb200_amg *amg, int *sweeps_per_level);
/* A hierarchy on disk ("ship hierarchies between boxes"): hb200_amg_save writes, per rank, the level matrices
 * A_l, P_l as binary IJ files (`<dir>/A<l>.<rank>.bin`, `<dir>/P<l>.<rank>.bin`: HYPRE_IJMatrixPrintBinary's lossless
 * format, readable by hypre itself) and one record `<dir>/amg.<rank>.bin` with l1 norms, CF markers, weights, Chebyshev
 * data, the dense coarse matrix and the cycle parameters.  hb200_amg_load builds the same hierarchy from it in a
 * process that holds no hypre at all, on the same number of ranks; the loaded hierarchy owns its matrices
 * (hb200_amg_destroy frees them; hb200_amg_level_matrix hands out borrowed handles: level 0 of `which` = 0 is the
 * operator the Krylov solvers take).  Both are collective. */
int hb200_amg_save(const hb200_amg *amg, const char *dirname);
int hb200_amg_load(hb200_amg **amg, const char *dirname);
int hb200_amg_level_matrix(hb200_amg *amg, int level, int which, hb200_parcsr **M);   /* which: 0 = A_l, 1 = P_l */
int hb200_amg_num_levels(const hb200_amg *amg);
/* per-level device vectors after a cycle, for parity tests: which = 0 F_array, 1 U_array */
int hb200_amg_level_vector(hb200_amg *amg, int level, int which, double **dev, int *n);

/* ------------------------------------------------------------------------------------ */
/* (a15,a16) Krylov drivers                                                               */
/* ------------------------------------------------------------------------------------ */

#define HB200_PRECOND_NONE      0
#define HB200_PRECOND_AMG       1   /* HYPRE_BoomerAMGSolve */
#define HB200_PRECOND_DIAGSCALE 2   /* HYPRE_ParCSRDiagScale (HYPRE_parcsr_pcg.c) */

/* Fields = the user-settable part of hypre_PCGData (src/krylov/pcg.h:104-149) */
typedef struct {
   double tol, a_tol, atolf, cf_tol, rtol;
   int    max_iter, two_norm, rel_change, recompute_residual, recompute_residual_p;
   int    stop_crit, skip_break, flex, hybrid;
   int    logging;        /* > 0: fill norms[] / rel_norms[] (pcg.c:774-778) */
   int    print_level;    /* > 1: rank 0 prints the reference's iteration table */
} hb200_pcg_params;

typedef struct {
   int    num_iterations;
   int    converged;
   double rel_residual_norm;
   int    error_flag;           /* hypre error bits raised inside the solve */
   double solve_ms;             /* device time of the solve (CUDA events) */
   long long kernel_launches;   /* hb200 kernels launched during the solve */
} hb200_krylov_result;

void hb200_pcg_default_params(hb200_pcg_params *p);   /* hypre_PCGCreate defaults, pcg.c:52-110 */
/* hypre_PCGSolve (src/krylov/pcg.c:313-1016) with the ParCSR function table of
 * HYPRE_ParCSRPCGCreate (src/parcsr_ls/HYPRE_parcsr_pcg.c:15-38).  b, x dev pointers;
 * norms/rel_norms host arrays of max_iter+1 entries or NULL. */
int hb200_pcg_solve(hb200_parcsr *A, int precond_kind, hb200_amg *amg,
                    const hb200_pcg_params *params, const double *b_dev, double *x_dev,
                    double *norms, double *rel_norms, hb200_krylov_result *result);
/* Setup-time warm-up (hypre_PCGSetup / hypre_GMRESSetup, src/krylov/pcg.c:198-283, gmres.c:185-283,
 * allocate the work vectors there): allocates the persistent Krylov workspace and runs a few
 * iterations on b = 1, x = 0 inside it so that halo plans, smoother scratch and the captured V-cycle
 * graphs of the solve exist before the application times its Solve call.  is_gmres: 0 = PCG, 1 =
 * GMRES(k_dim), 2 = FlexGMRES(k_dim), 3 = COGMRES(k_dim), 4 = BiCGSTAB, 5 = LGMRES(k_dim).  Collective. */
int hb200_krylov_warmup(hb200_parcsr *A, int precond_kind, hb200_amg *amg, int is_gmres, int k_dim);

/* Same call with HOST b and x (what HYPRE_PCGSolve sees in a CPU-memory application):
 * H2D of b and x0, solve, D2H of x, all inside. */
int hb200_pcg_solve_host(hb200_parcsr *A, int precond_kind, hb200_amg *amg,
                         const hb200_pcg_params *params, const double *b_host,
                         double *x_host, double *norms, double *rel_norms,
                         hb200_krylov_result *result);

/* user-settable part of hypre_GMRESData (src/krylov/gmres.h:74-114) */
typedef struct {
   double tol, a_tol, cf_tol;
   int    k_dim, min_iter, max_iter, rel_change, skip_real_r_check, stop_crit, hybrid;
   int    logging, print_level;
   int    cgs, unroll;    /* COGMRES only (cogmres.h: cgs = 2 re-orthogonalises; the device kernels have one unrolling) */
   int    aug_dim, approx_constant;   /* LGMRES only (lgmres.h: augmentation vectors, constant size of the space) */
} hb200_gmres_params;

void hb200_gmres_default_params(hb200_gmres_params *p);   /* hypre_GMRESCreate, gmres.c:52-110 */
/* hypre_GMRESSolve (src/krylov/gmres.c:294-1100), ParCSR table of HYPRE_ParCSRGMRESCreate
 * (src/parcsr_ls/HYPRE_parcsr_gmres.c:15-46). */
int hb200_gmres_solve(hb200_parcsr *A, int precond_kind, hb200_amg *amg,
                      const hb200_gmres_params *params, const double *b_dev, double *x_dev,
                      double *norms, hb200_krylov_result *result);
int hb200_gmres_solve_host(hb200_parcsr *A, int precond_kind, hb200_amg *amg,
                           const hb200_gmres_params *params, const double *b_host,
                           double *x_host, double *norms, hb200_krylov_result *result);

/* ------------------------------------------------------------------------------------ */
/* (f4) the other Krylov drivers of the ParCSR function table                              */
/* ------------------------------------------------------------------------------------ */

/* hypre_FlexGMRESSolve (src/krylov/flexgmres.c:288-812), table of HYPRE_ParCSRFlexGMRESCreate
 * (src/parcsr_ls/HYPRE_parcsr_flexgmres.c:15-45) with the default modify_pc (a no-op).  Reads k_dim, tol,
 * a_tol, cf_tol, min_iter, max_iter, logging, print_level (<= 2). */
int hb200_flexgmres_solve(hb200_parcsr *A, int precond_kind, hb200_amg *amg,
                          const hb200_gmres_params *params, const double *b_dev, double *x_dev,
                          double *norms, hb200_krylov_result *result);
int hb200_flexgmres_solve_host(hb200_parcsr *A, int precond_kind, hb200_amg *amg,
                               const hb200_gmres_params *params, const double *b_host,
                               double *x_host, double *norms, hb200_krylov_result *result);

/* hypre_COGMRESSolve (src/krylov/cogmres.c:270-896), table of HYPRE_ParCSRCOGMRESCreate
 * (src/parcsr_ls/HYPRE_parcsr_cogmres.c:15-50): classical Gram-Schmidt through the batched vector
 * operations hypre_ParVectorMassInnerProd / MassDotpTwo / MassAxpy (src/parcsr_mv/par_vector_batched.c:
 * 17-135).  params->cgs = 2: the re-orthogonalising variant (k_dim <= 50). */
int hb200_cogmres_solve(hb200_parcsr *A, int precond_kind, hb200_amg *amg,
                        const hb200_gmres_params *params, const double *b_dev, double *x_dev,
                        double *norms, hb200_krylov_result *result);
int hb200_cogmres_solve_host(hb200_parcsr *A, int precond_kind, hb200_amg *amg,
                             const hb200_gmres_params *params, const double *b_host,
                             double *x_host, double *norms, hb200_krylov_result *result);

/* hypre_LGMRESSolve (src/krylov/lgmres.c:320-940), table of HYPRE_ParCSRLGMRESCreate
 * (src/parcsr_ls/HYPRE_parcsr_lgmres.c): GMRES(k_dim) whose space holds up to params->aug_dim error
 * approximations of the previous restart cycles. */
int hb200_lgmres_solve(hb200_parcsr *A, int precond_kind, hb200_amg *amg,
                       const hb200_gmres_params *params, const double *b_dev, double *x_dev,
                       double *norms, hb200_krylov_result *result);
int hb200_lgmres_solve_host(hb200_parcsr *A, int precond_kind, hb200_amg *amg,
                            const hb200_gmres_params *params, const double *b_host,
                            double *x_host, double *norms, hb200_krylov_result *result);

/* user-settable part of hypre_BiCGSTABData (src/krylov/bicgstab.h:70-108) */
typedef struct {
   double tol, a_tol, cf_tol;
   int    min_iter, max_iter, stop_crit, hybrid;
   int    logging, print_level;
} hb200_bicgstab_params;

void hb200_bicgstab_default_params(hb200_bicgstab_params *p);   /* hypre_BiCGSTABCreate, bicgstab.c:64-106 */
/* hypre_BiCGSTABSolve (src/krylov/bicgstab.c:246-606), table of HYPRE_ParCSRBiCGSTABCreate
 * (src/parcsr_ls/HYPRE_parcsr_bicgstab.c:15-40). */
int hb200_bicgstab_solve(hb200_parcsr *A, int precond_kind, hb200_amg *amg,
                         const hb200_bicgstab_params *params, const double *b_dev, double *x_dev,
                         double *norms, hb200_krylov_result *result);
int hb200_bicgstab_solve_host(hb200_parcsr *A, int precond_kind, hb200_amg *amg,
                              const hb200_bicgstab_params *params, const double *b_host,
                              double *x_host, double *norms, hb200_krylov_result *result);

/* ------------------------------------------------------------------------------------ */
/* (f3) IJ interface and on-disk formats: matrices and vectors enter without hypre        */
/* ------------------------------------------------------------------------------------ */

/* HYPRE_IJMatrixCreate(ilower, iupper, jlower, jupper) + SetValues / AddToValues + Assemble
 * (src/IJ_mv/HYPRE_IJMatrix.c:45, 539, 731, 869; hypre_IJMatrixAssembleParCSR, src/IJ_mv/IJMatrix_parcsr.c:2641)
 * in one collective call: this rank's coordinate triplets (global indices, rows inside [ilower, iupper]; HOST
 * arrays) become its rows of a ParCSR matrix on the device.  A row keeps the order its entries were given in,
 * a repeated (i, j) lands on its first occurrence (add_duplicates: summed, else the last value wins), the
 * diagonal entry of a square block comes first (the relaxation sweeps rely on it, par_relax.c:274),
 * col_map_offd ascends.  On N ranks the CommPkg (hypre_MatvecCommPkgCreate, src/parcsr_mv/
 * par_csr_communication.c) is built from two NCCL all-gathers.  Entries for rows of other ranks
 * (HYPRE_IJMatrixAddToValues off-processor; what hypre_IJMatrixRead does with the entries of a part file outside its
 * row range) travel to the owner, who adds them after its own entries (one more all-gather, only when there are any). */
int hb200_parcsr_from_ij(hb200_parcsr **A, int64_t ilower, int64_t iupper, int64_t jlower, int64_t jupper,
                         int64_t num_entries, const int64_t *rows, const int64_t *cols,
                         const double *values, int add_duplicates);
/* HYPRE_IJMatrixRead (src/IJ_mv/IJMatrix.c:110-249): format 0 — every rank reads `<filename>.<5-digit rank>`
 * (header `ilower iupper jlower jupper`, then `i j value` lines); format 1: HYPRE_IJMatrixReadMM, the whole
 * Matrix Market file on one rank (coordinate, real / integer, general / symmetric); format 2:
 * HYPRE_IJMatrixReadBinary (IJMatrix.c:252-470), `<filename>.<5-digit rank>.bin` (88-byte header, then row
 * indices, column indices, values; 4- or 8-byte indices and values).  Collective. */
int hb200_parcsr_read_ij(hb200_parcsr **A, const char *filename, int format);
/* sizes of a matrix (what the caller of hb200_parcsr_create passed; needed after from_ij / read_ij): info[0..11]
 * = num_rows, num_cols, num_cols_offd, diag nonzeros, offd nonzeros, num_sends, num_recvs, send_map_starts[num_sends],
 * first_row_index, first_col_diag, global_num_rows, global_num_cols (hypre_ParCSRMatrix, par_csr_matrix.h:27-92) */
int hb200_parcsr_info(const hb200_parcsr *A, int64_t *info12);
/* HYPRE_IJMatrixPrint -> hypre_ParCSRMatrixPrintIJ (src/parcsr_mv/par_csr_matrix.c): the same text the
 * reference writes for the same matrix (diag entries, then offd entries of a row; `%.14e`). */
int hb200_parcsr_print_ij(const hb200_parcsr *A, const char *filename);
/* HYPRE_IJMatrixPrintBinary -> hypre_ParCSRMatrixPrintBinaryIJ (par_csr_matrix.c:1120-1400): the lossless format
 * (fp64 values, 64-bit indices).  Collective (the header carries the global nonzero count). */
int hb200_parcsr_print_ij_binary(const hb200_parcsr *A, const char *filename);
/* HYPRE_IJVectorPrint / HYPRE_IJVectorRead (hypre_ParVectorPrintIJ, src/parcsr_mv/par_vector.c): `jlower
 * jupper`, then `j value` lines, per rank file.  read with x_dev == NULL only returns the range. */
int hb200_vector_print_ij(const double *x_dev, int64_t jlower, int num_values, const char *filename);
int hb200_vector_read_ij(const char *filename, int64_t *jlower, int64_t *jupper, double *x_dev, int capacity);
/* The host halves (no GPU): the assembly of one rank's rows (call with NULL arrays for the sizes first) and
 * the CommPkg of rank `me` from the all-gathered ownership (own5[5 q ..] = ilower, iupper, jlower, jupper,
 * number of off-range columns of rank q) and need lists (need[q * max_offd ..] = rank q's col_map_offd). */
int hb200_host_ij_assemble(int64_t ilower, int64_t iupper, int64_t jlower, int64_t jupper, int64_t num_entries,
                           const int64_t *rows, const int64_t *cols, const double *values, int add_duplicates,
                           int *diag_nnz, int *offd_nnz, int *num_cols_offd, int *diag_i, int *diag_j,
                           double *diag_data, int *offd_i, int *offd_j, double *offd_data,
                           int64_t *col_map_offd);
int hb200_host_ij_commpkg(int num_ranks, int me, const int64_t *own5, const int64_t *need, int64_t max_offd,
                          int *num_sends, int *send_procs, int *send_map_starts, int *send_map_elmts,
                          int send_capacity, int *num_recvs, int *recv_procs, int *recv_vec_starts);

#ifdef __cplusplus
}
#endif
#endif /* HB200_H */

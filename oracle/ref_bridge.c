/*
 * ref_bridge.c — thin C bridge over the UNMODIFIED compiled reference (oracle/_ref/libHYPRE_ref*.so).
 *
 * TEST / BASELINE INFRASTRUCTURE and the north_star-mandated *setup provider*: it (1) builds
 * the ij test problems with the reference's own generators, (2) runs the reference's own
 * HYPRE_BoomerAMGSetup with ij's parameter defaults, (3) exposes the resulting hierarchy
 * (hypre_ParAMGData, hypre_ParCSRMatrix, hypre_ParCSRCommPkg) as plain pointers + sizes so that
 * Python can hand it to the hb200 C-ABI, and (4) runs the reference's own CPU matvec / relax /
 * cycle / PCG / GMRES on the same data as the parity oracle and the CPU baseline.
 * It contains no solver arithmetic of its own.  Compiled against the reference headers where
 * they lie (never copied into this repository).
 *
 * ij parameter defaults mirrored here: src/test/ij.c:150-420 (variables), :5563-5749 (PCG+AMG),
 * :7322-7560 (GMRES+AMG); problem builders :10310-10480 (laplacian), :11807-11920 (27pt),
 * :12173-12300 (vardifconv).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <sys/time.h>

#include "_hypre_utilities.h"
#include "HYPRE.h"
#include "HYPRE_parcsr_mv.h"
#include "_hypre_parcsr_mv.h"
#include "HYPRE_parcsr_ls.h"
#include "_hypre_parcsr_ls.h"
#include "HYPRE_krylov.h"
#include "_hypre_krylov.h"
#ifdef HYPRE_USING_OPENMP
#include <omp.h>
#endif

typedef struct
{
   MPI_Comm            comm;
   int                 myid, nprocs;
   HYPRE_ParCSRMatrix  A;
   HYPRE_ParVector     b, x;
   int                 b_from_generator;
   HYPRE_Solver        amg;
   HYPRE_IJMatrix      ij;                /* the owner of A when the problem was read from an IJ file */
   int64_t            *colmap64[64][2];   /* widened col_map_offd per level / matrix kind */
} rb_problem;

/* plain view of one hypre_ParCSRMatrix (src/parcsr_mv/par_csr_matrix.h:27-92) */
typedef struct
{
   int      num_rows, num_cols, num_cols_offd;
   int      diag_nnz, offd_nnz;
   int     *diag_i, *diag_j;
   double  *diag_data;
   int     *offd_i, *offd_j;
   double  *offd_data;
   int64_t *col_map_offd;
   int64_t  first_row, first_col, global_rows, global_cols;
   int      num_sends, num_recvs;
   int     *send_procs, *send_map_starts, *send_map_elmts, *recv_procs, *recv_vec_starts;
} rb_parcsr_view;

static int g_initialized = 0;

static double wall(void)
{
   struct timeval tv;
   gettimeofday(&tv, NULL);
   return (double) tv.tv_sec + 1e-6 * (double) tv.tv_usec;
}

int rb_init(void)
{
   if (!g_initialized)
   {
#ifndef HYPRE_SEQUENTIAL
      int flag = 0;
      MPI_Initialized(&flag);
      if (!flag) { MPI_Init(NULL, NULL); }
#endif
      HYPRE_Initialize();
      g_initialized = 1;
   }
   return 0;
}

int rb_finalize(void)
{
   if (g_initialized)
   {
      HYPRE_Finalize();
#ifndef HYPRE_SEQUENTIAL
      MPI_Finalize();
#endif
      g_initialized = 0;
   }
   return 0;
}

int rb_num_threads(void) { return hypre_NumThreads(); }
void rb_set_num_threads(int n)
{
#ifdef HYPRE_USING_OPENMP
   omp_set_num_threads(n);
#else
   (void) n;
#endif
}
int rb_comm_rank(void) { int r; hypre_MPI_Comm_rank(hypre_MPI_COMM_WORLD, &r); return r; }
int rb_comm_size(void) { int s; hypre_MPI_Comm_size(hypre_MPI_COMM_WORLD, &s); return s; }
int rb_error_flag(void) { return (int) HYPRE_GetError(); }
void rb_clear_errors(void) { HYPRE_ClearAllErrors(); }
int rb_sizeof_bigint(void) { return (int) sizeof(HYPRE_BigInt); }

/* problem type: 0 = -laplacian (7-pt), 1 = -27pt, 2 = -vardifconv (eps) ;
 * rhs_type: 0 = ones (ij default "RHS vector has unit coefficients"), 1 = -rhsrand,
 *           6 = from the generator (vardifconv, ij.c:3221) ; x0 = 0 (or x0rand = 1) */
rb_problem *rb_problem_create(int type, int nx, int ny, int nz, int P, int Q, int R,
                              double eps, int rhs_type, int x0rand)
{
   rb_problem *pb = (rb_problem *) calloc(1, sizeof(rb_problem));
   HYPRE_Int p, q, r;
   HYPRE_Real values[4];
   rb_init();
   pb->comm = hypre_MPI_COMM_WORLD;
   hypre_MPI_Comm_rank(pb->comm, &pb->myid);
   hypre_MPI_Comm_size(pb->comm, &pb->nprocs);
   if (P * Q * R != pb->nprocs)
   {
      fprintf(stderr, "rb_problem_create: P*Q*R = %d != nprocs = %d\n", P * Q * R, pb->nprocs);
      free(pb);
      return NULL;
   }
   p = pb->myid % P;
   q = ((pb->myid - p) / P) % Q;
   r = (pb->myid - p - P * q) / (P * Q);
   if (type == 0)
   {
      values[1] = -1.0; values[2] = -1.0; values[3] = -1.0;
      values[0] = 0.0;
      if (nx > 1) { values[0] += 2.0; }
      if (ny > 1) { values[0] += 2.0; }
      if (nz > 1) { values[0] += 2.0; }
      pb->A = GenerateLaplacian(pb->comm, nx, ny, nz, P, Q, R, p, q, r, values);
   }
   else if (type == 1)
   {
      values[0] = 26.0;
      if (nx == 1 || ny == 1 || nz == 1) { values[0] = 8.0; }
      if (nx * ny == 1 || nx * nz == 1 || ny * nz == 1) { values[0] = 2.0; }
      values[1] = -1.0;
      pb->A = GenerateLaplacian27pt(pb->comm, nx, ny, nz, P, Q, R, p, q, r, values);
   }
   else
   {
      pb->A = GenerateVarDifConv(pb->comm, nx, ny, nz, P, Q, R, p, q, r, eps, &pb->b);
      pb->b_from_generator = 1;
      rhs_type = 6;
   }
   {
      hypre_ParCSRMatrix *A = (hypre_ParCSRMatrix *) pb->A;
      if (!hypre_ParCSRMatrixCommPkg(A)) { hypre_MatvecCommPkgCreate(A); }
      if (rhs_type != 6)
      {
         pb->b = (HYPRE_ParVector) hypre_ParVectorCreate(pb->comm, hypre_ParCSRMatrixGlobalNumRows(A),
                                                        hypre_ParCSRMatrixRowStarts(A));
         hypre_ParVectorInitialize((hypre_ParVector *) pb->b);
         if (rhs_type == 1)
         {
            /* ij -rhsrand (ij.c:3777-3800): seed 22775 */
            HYPRE_ParVectorSetRandomValues(pb->b, 22775);
         }
         else
         {
            hypre_ParVectorSetConstantValues((hypre_ParVector *) pb->b, 1.0);
         }
      }
      pb->x = (HYPRE_ParVector) hypre_ParVectorCreate(pb->comm, hypre_ParCSRMatrixGlobalNumCols(A),
                                                     hypre_ParCSRMatrixColStarts(A));
      hypre_ParVectorInitialize((hypre_ParVector *) pb->x);
      if (x0rand) { HYPRE_ParVectorSetRandomValues(pb->x, 775); }   /* ij.c:4272-4290 */
      else { hypre_ParVectorSetConstantValues((hypre_ParVector *) pb->x, 0.0); }
   }
   return pb;
}

/* ij -fromfile <name> (ij.c:3174: HYPRE_IJMatrixRead) / a Matrix Market file (HYPRE_IJMatrixReadMM): the matrix the
 * reference assembles from `<name>.<5-digit rank>`; b = ones, x = 0 */
rb_problem *rb_problem_from_ij_file(const char *filename, int is_mm)
{
   rb_problem *pb = (rb_problem *) calloc(1, sizeof(rb_problem));
   hypre_ParCSRMatrix *A;
   void *obj = NULL;
   rb_init();
   pb->comm = hypre_MPI_COMM_WORLD;
   hypre_MPI_Comm_rank(pb->comm, &pb->myid);
   hypre_MPI_Comm_size(pb->comm, &pb->nprocs);
   if (is_mm == 2) { HYPRE_IJMatrixReadBinary(filename, pb->comm, HYPRE_PARCSR, &pb->ij); }
   else if (is_mm) { HYPRE_IJMatrixReadMM(filename, pb->comm, HYPRE_PARCSR, &pb->ij); }
   else { HYPRE_IJMatrixRead(filename, pb->comm, HYPRE_PARCSR, &pb->ij); }
   if (HYPRE_GetError() || !pb->ij) { HYPRE_ClearAllErrors(); free(pb); return NULL; }
   HYPRE_IJMatrixGetObject(pb->ij, &obj);
   pb->A = (HYPRE_ParCSRMatrix) obj;
   A = (hypre_ParCSRMatrix *) pb->A;
   if (!hypre_ParCSRMatrixCommPkg(A)) { hypre_MatvecCommPkgCreate(A); }
   pb->b = (HYPRE_ParVector) hypre_ParVectorCreate(pb->comm, hypre_ParCSRMatrixGlobalNumRows(A), hypre_ParCSRMatrixRowStarts(A));
   hypre_ParVectorInitialize((hypre_ParVector *) pb->b);
   hypre_ParVectorSetConstantValues((hypre_ParVector *) pb->b, 1.0);
   pb->x = (HYPRE_ParVector) hypre_ParVectorCreate(pb->comm, hypre_ParCSRMatrixGlobalNumCols(A), hypre_ParCSRMatrixColStarts(A));
   hypre_ParVectorInitialize((hypre_ParVector *) pb->x);
   hypre_ParVectorSetConstantValues((hypre_ParVector *) pb->x, 0.0);
   return pb;
}

/* HYPRE_IJMatrixPrint of the fine-level operator / HYPRE_IJVectorPrint-format file of the right-hand side */
int rb_print_ij(rb_problem *pb, const char *filename)
{
   hypre_ParCSRMatrixPrintIJ((hypre_ParCSRMatrix *) pb->A, 0, 0, filename);
   return (int) HYPRE_GetError();
}
int rb_print_ij_binary(rb_problem *pb, const char *filename)
{
   hypre_ParCSRMatrixPrintBinaryIJ((hypre_ParCSRMatrix *) pb->A, 0, 0, filename);
   return (int) HYPRE_GetError();
}
int rb_print_vector_ij(rb_problem *pb, const double *values, const char *filename)
{
   hypre_ParCSRMatrix *M = (hypre_ParCSRMatrix *) pb->A;
   hypre_ParVector *v = hypre_ParVectorCreate(pb->comm, hypre_ParCSRMatrixGlobalNumRows(M), hypre_ParCSRMatrixRowStarts(M));
   int n = hypre_ParCSRMatrixNumRows(M), k;
   hypre_ParVectorInitialize(v);
   for (k = 0; k < n; k++) { hypre_VectorData(hypre_ParVectorLocalVector(v))[k] = values[k]; }
   hypre_ParVectorPrintIJ(v, 0, filename);
   hypre_ParVectorDestroy(v);
   return (int) HYPRE_GetError();
}

void rb_problem_destroy(rb_problem *pb)
{
   int l, k;
   if (!pb) { return; }
   if (pb->amg) { HYPRE_BoomerAMGDestroy(pb->amg); }
   if (pb->ij) { HYPRE_IJMatrixDestroy(pb->ij); }
   else if (pb->A) { HYPRE_ParCSRMatrixDestroy(pb->A); }
   if (pb->b) { HYPRE_ParVectorDestroy(pb->b); }
   if (pb->x) { HYPRE_ParVectorDestroy(pb->x); }
   for (l = 0; l < 64; l++) for (k = 0; k < 2; k++) { free(pb->colmap64[l][k]); }
   free(pb);
}

double *rb_problem_b(rb_problem *pb) { return hypre_VectorData(hypre_ParVectorLocalVector((hypre_ParVector *) pb->b)); }
double *rb_problem_x(rb_problem *pb) { return hypre_VectorData(hypre_ParVectorLocalVector((hypre_ParVector *) pb->x)); }
int rb_problem_local_rows(rb_problem *pb) { return hypre_ParCSRMatrixNumRows((hypre_ParCSRMatrix *) pb->A); }
long long rb_problem_global_rows(rb_problem *pb) { return (long long) hypre_ParCSRMatrixGlobalNumRows((hypre_ParCSRMatrix *) pb->A); }

/* BoomerAMG with ij's defaults (ij.c:5578-5749): only the setters whose ij default differs
 * from the library default (par_amg.c:160-316) or that the caller overrides are needed:
 * tol = pc_tol = 0, max_iter = precon_cycles = 1, max_row_sum = 1.0.  Arguments < 0 (or NaN-free
 * negative doubles) mean "ij default". */
int rb_amg_setup(rb_problem *pb, int relax_type, int relax_down, int relax_up, int relax_coarse,
                 int relax_order, int num_sweeps, int coarsen_type, int interp_type, int Pmx,
                 int agg_nl, int cycle_type, int cheby_order, int cheby_eig_est, int cheby_scale,
                 int cheby_variant, double cheby_fraction, double strong_threshold,
                 double relax_wt, double outer_wt, int keep_transpose, int max_levels,
                 int print_level, double *setup_seconds)
{
   HYPRE_Solver amg;
   double t0;
   if (pb->amg) { HYPRE_BoomerAMGDestroy(pb->amg); pb->amg = NULL; }
   HYPRE_BoomerAMGCreate(&amg);
   HYPRE_BoomerAMGSetTol(amg, 0.0);
   HYPRE_BoomerAMGSetMaxIter(amg, 1);
   HYPRE_BoomerAMGSetMaxRowSum(amg, 1.0);
   HYPRE_BoomerAMGSetPrintLevel(amg, print_level);
   if (coarsen_type >= 0) { HYPRE_BoomerAMGSetCoarsenType(amg, coarsen_type); }
   if (interp_type >= 0) { HYPRE_BoomerAMGSetInterpType(amg, interp_type); }
   if (Pmx >= 0) { HYPRE_BoomerAMGSetPMaxElmts(amg, Pmx); }
   if (agg_nl >= 0) { HYPRE_BoomerAMGSetAggNumLevels(amg, agg_nl); }
   if (cycle_type >= 1) { HYPRE_BoomerAMGSetCycleType(amg, cycle_type); }
   if (num_sweeps >= 0) { HYPRE_BoomerAMGSetNumSweeps(amg, num_sweeps); }
   if (relax_type >= 0) { HYPRE_BoomerAMGSetRelaxType(amg, relax_type); }
   if (relax_down >= 0) { HYPRE_BoomerAMGSetCycleRelaxType(amg, relax_down, 1); }
   if (relax_up >= 0) { HYPRE_BoomerAMGSetCycleRelaxType(amg, relax_up, 2); }
   if (relax_coarse >= 0) { HYPRE_BoomerAMGSetCycleRelaxType(amg, relax_coarse, 3); }
   HYPRE_BoomerAMGSetRelaxOrder(amg, relax_order >= 0 ? relax_order : 0);   /* ij default 0 */
   if (cheby_order >= 0) { HYPRE_BoomerAMGSetChebyOrder(amg, cheby_order); }
   if (cheby_eig_est >= 0) { HYPRE_BoomerAMGSetChebyEigEst(amg, cheby_eig_est); }
   if (cheby_scale >= 0) { HYPRE_BoomerAMGSetChebyScale(amg, cheby_scale); }
   if (cheby_variant >= 0) { HYPRE_BoomerAMGSetChebyVariant(amg, cheby_variant); }
   if (cheby_fraction > 0) { HYPRE_BoomerAMGSetChebyFraction(amg, cheby_fraction); }
   if (strong_threshold >= 0) { HYPRE_BoomerAMGSetStrongThreshold(amg, strong_threshold); }
   if (relax_wt > -1e30) { HYPRE_BoomerAMGSetRelaxWt(amg, relax_wt); }
   if (outer_wt > -1e30) { HYPRE_BoomerAMGSetOuterWt(amg, outer_wt); }
   if (keep_transpose >= 0) { HYPRE_BoomerAMGSetKeepTranspose(amg, keep_transpose); }
   if (max_levels >= 1) { HYPRE_BoomerAMGSetMaxLevels(amg, max_levels); }
   t0 = wall();
   HYPRE_BoomerAMGSetup(amg, pb->A, pb->b, pb->x);
   if (setup_seconds) { *setup_seconds = wall() - t0; }
   pb->amg = amg;
   return (int) HYPRE_GetError();
}

/* ---- hierarchy access -------------------------------------------------------------- */

int rb_num_levels(rb_problem *pb)
{
   return hypre_ParAMGDataNumLevels((hypre_ParAMGData *) pb->amg);
}

static hypre_ParCSRMatrix *level_matrix(rb_problem *pb, int level, int which)
{
   hypre_ParAMGData *amg = (hypre_ParAMGData *) pb->amg;
   if (!amg) { return level == 0 && which == 0 ? (hypre_ParCSRMatrix *) pb->A : NULL; }
   if (which == 0) { return hypre_ParAMGDataAArray(amg)[level]; }
   return hypre_ParAMGDataPArray(amg)[level];
}

int rb_level_matrix(rb_problem *pb, int level, int which, rb_parcsr_view *v)
{
   hypre_ParCSRMatrix  *M = level_matrix(pb, level, which);
   hypre_CSRMatrix     *diag, *offd;
   hypre_ParCSRCommPkg *pkg;
   HYPRE_BigInt        *cmap;
   int                  k;
   if (!M) { return 1; }
   if (!hypre_ParCSRMatrixCommPkg(M)) { hypre_MatvecCommPkgCreate(M); }   /* par_csr_matvec.c:102-106 */
   diag = hypre_ParCSRMatrixDiag(M);
   offd = hypre_ParCSRMatrixOffd(M);
   pkg  = hypre_ParCSRMatrixCommPkg(M);
   memset(v, 0, sizeof(*v));
   v->num_rows      = hypre_CSRMatrixNumRows(diag);
   v->num_cols      = hypre_CSRMatrixNumCols(diag);
   v->num_cols_offd = hypre_CSRMatrixNumCols(offd);
   v->diag_i = hypre_CSRMatrixI(diag); v->diag_j = hypre_CSRMatrixJ(diag); v->diag_data = hypre_CSRMatrixData(diag);
   v->offd_i = hypre_CSRMatrixI(offd); v->offd_j = hypre_CSRMatrixJ(offd); v->offd_data = hypre_CSRMatrixData(offd);
   v->diag_nnz = v->num_rows ? v->diag_i[v->num_rows] : 0;
   v->offd_nnz = (v->num_rows && v->offd_i) ? v->offd_i[v->num_rows] : 0;
   cmap = hypre_ParCSRMatrixColMapOffd(M);
   if (v->num_cols_offd > 0 && level < 64)
   {
      free(pb->colmap64[level][which]);
      pb->colmap64[level][which] = (int64_t *) malloc(sizeof(int64_t) * (size_t) v->num_cols_offd);
      for (k = 0; k < v->num_cols_offd; k++) { pb->colmap64[level][which][k] = (int64_t) cmap[k]; }
      v->col_map_offd = pb->colmap64[level][which];
   }
   v->first_row   = (int64_t) hypre_ParCSRMatrixFirstRowIndex(M);
   v->first_col   = (int64_t) hypre_ParCSRMatrixFirstColDiag(M);
   v->global_rows = (int64_t) hypre_ParCSRMatrixGlobalNumRows(M);
   v->global_cols = (int64_t) hypre_ParCSRMatrixGlobalNumCols(M);
   if (pkg)
   {
      v->num_sends       = hypre_ParCSRCommPkgNumSends(pkg);
      v->num_recvs       = hypre_ParCSRCommPkgNumRecvs(pkg);
      v->send_procs      = hypre_ParCSRCommPkgSendProcs(pkg);
      v->send_map_starts = hypre_ParCSRCommPkgSendMapStarts(pkg);
      v->send_map_elmts  = hypre_ParCSRCommPkgSendMapElmts(pkg);
      v->recv_procs      = hypre_ParCSRCommPkgRecvProcs(pkg);
      v->recv_vec_starts = hypre_ParCSRCommPkgRecvVecStarts(pkg);
   }
   return 0;
}

double *rb_level_l1_norms(rb_problem *pb, int level)
{
   hypre_ParAMGData *amg = (hypre_ParAMGData *) pb->amg;
   hypre_Vector **l1 = hypre_ParAMGDataL1Norms(amg);
   if (!l1 || !l1[level]) { return NULL; }
   return hypre_VectorData(l1[level]);
}

int *rb_level_cf_marker(rb_problem *pb, int level)
{
   hypre_ParAMGData *amg = (hypre_ParAMGData *) pb->amg;
   hypre_IntArray **cf = hypre_ParAMGDataCFMarkerArray(amg);
   if (!cf || !cf[level]) { return NULL; }
   return hypre_IntArrayData(cf[level]);
}

double rb_level_relax_weight(rb_problem *pb, int level) { return hypre_ParAMGDataRelaxWeight((hypre_ParAMGData *) pb->amg)[level]; }
double rb_level_omega(rb_problem *pb, int level) { return hypre_ParAMGDataOmega((hypre_ParAMGData *) pb->amg)[level]; }

double *rb_level_cheby_ds(rb_problem *pb, int level)
{
   hypre_Vector **ds = hypre_ParAMGDataChebyDS((hypre_ParAMGData *) pb->amg);
   if (!ds || !ds[level]) { return NULL; }
   return hypre_VectorData(ds[level]);
}
double *rb_level_cheby_coefs(rb_problem *pb, int level)
{
   HYPRE_Real **c = hypre_ParAMGDataChebyCoefs((hypre_ParAMGData *) pb->amg);
   if (!c) { return NULL; }
   return c[level];
}

/* out[0..3] num_grid_sweeps, out[4..7] grid_relax_type, out[8] relax_order, [9] cycle_type,
 * [10] fcycle, [11] cheby_order, [12] cheby_scale, [13] cheby_variant, [14] user_relax_type,
 * [15] max_iter, [16] min_iter, [17] converge_type, [18] restriction type, [19] block_mode,
 * [20] smooth_num_levels, [21] additive, [22] mult_additive, [23] simple */
int rb_amg_params(rb_problem *pb, int *out, double *tol)
{
   hypre_ParAMGData *amg = (hypre_ParAMGData *) pb->amg;
   int k;
   for (k = 0; k < 4; k++)
   {
      out[k]     = hypre_ParAMGDataNumGridSweeps(amg)[k];
      out[4 + k] = hypre_ParAMGDataGridRelaxType(amg)[k];
   }
   out[8]  = hypre_ParAMGDataRelaxOrder(amg);
   out[9]  = hypre_ParAMGDataCycleType(amg);
   out[10] = hypre_ParAMGDataFCycle(amg);
   out[11] = hypre_ParAMGDataChebyOrder(amg);
   out[12] = hypre_ParAMGDataChebyScale(amg);
   out[13] = hypre_ParAMGDataChebyVariant(amg);
   out[14] = hypre_ParAMGDataUserRelaxType(amg);
   out[15] = hypre_ParAMGDataMaxIter(amg);
   out[16] = hypre_ParAMGDataMinIter(amg);
   out[17] = hypre_ParAMGDataConvergeType(amg);
   out[18] = hypre_ParAMGDataRestriction(amg);
   out[19] = hypre_ParAMGDataBlockMode(amg);
   out[20] = hypre_ParAMGDataSmoothNumLevels(amg);
   out[21] = hypre_ParAMGDataAdditive(amg);
   out[22] = hypre_ParAMGDataMultAdditive(amg);
   out[23] = hypre_ParAMGDataSimple(amg);
   if (tol) { *tol = hypre_ParAMGDataTol(amg); }
   return 0;
}

/* coarsest-level dense matrix of relax type 9 (par_gauss_elim.c:33-440); returns n or 0 */
int rb_coarse_ge(rb_problem *pb, double **A_mat, int *first_row, int *num_local)
{
   hypre_ParAMGData *amg = (hypre_ParAMGData *) pb->amg;
   int nl = hypre_ParAMGDataNumLevels(amg);
   hypre_ParCSRMatrix *Ac = hypre_ParAMGDataAArray(amg)[nl - 1];
   int rt = hypre_ParAMGDataGridRelaxType(amg)[3];
   if (nl == 1) { rt = hypre_ParAMGDataUserRelaxType(amg); }
   if (!(rt == 9 || rt == 19)) { return 0; }
   if (hypre_ParAMGDataGSSetup(amg) == 0) { hypre_GaussElimSetup(amg, nl - 1, rt); }
   *A_mat     = hypre_ParAMGDataAMat(amg);
   *first_row = (int) hypre_ParCSRMatrixFirstRowIndex(Ac);
   *num_local = hypre_ParCSRMatrixNumRows(Ac);
   return (int) hypre_ParCSRMatrixGlobalNumRows(Ac);
}

/* ---- reference compute (the oracle) --------------------------------------------------- */

static hypre_ParVector *make_vec(rb_problem *pb, HYPRE_BigInt global, HYPRE_BigInt *starts, const double *src)
{
   hypre_ParVector *v = hypre_ParVectorCreate(pb->comm, global, starts);
   hypre_ParVectorInitialize(v);
   if (src)
   {
      memcpy(hypre_VectorData(hypre_ParVectorLocalVector(v)), src,
             sizeof(double) * (size_t) hypre_VectorSize(hypre_ParVectorLocalVector(v)));
   }
   return v;
}
static void take_vec(hypre_ParVector *v, double *dst)
{
   memcpy(dst, hypre_VectorData(hypre_ParVectorLocalVector(v)),
          sizeof(double) * (size_t) hypre_VectorSize(hypre_ParVectorLocalVector(v)));
}

/* y = alpha*M*x + beta*b  (hypre_ParCSRMatrixMatvecOutOfPlace) */
int rb_matvec(rb_problem *pb, int level, int which, double alpha, const double *x, double beta,
              const double *b, double *y)
{
   hypre_ParCSRMatrix *M = level_matrix(pb, level, which);
   hypre_ParVector *vx = make_vec(pb, hypre_ParCSRMatrixGlobalNumCols(M), hypre_ParCSRMatrixColStarts(M), x);
   hypre_ParVector *vb = make_vec(pb, hypre_ParCSRMatrixGlobalNumRows(M), hypre_ParCSRMatrixRowStarts(M), b);
   hypre_ParVector *vy = make_vec(pb, hypre_ParCSRMatrixGlobalNumRows(M), hypre_ParCSRMatrixRowStarts(M), NULL);
   hypre_ParCSRMatrixMatvecOutOfPlace(alpha, M, vx, beta, vb, vy);
   take_vec(vy, y);
   hypre_ParVectorDestroy(vx); hypre_ParVectorDestroy(vb); hypre_ParVectorDestroy(vy);
   return 0;
}

/* y = alpha*M^T*x + beta*y  (hypre_ParCSRMatrixMatvecT) */
int rb_matvecT(rb_problem *pb, int level, int which, double alpha, const double *x, double beta, double *y)
{
   hypre_ParCSRMatrix *M = level_matrix(pb, level, which);
   hypre_ParVector *vx = make_vec(pb, hypre_ParCSRMatrixGlobalNumRows(M), hypre_ParCSRMatrixRowStarts(M), x);
   hypre_ParVector *vy = make_vec(pb, hypre_ParCSRMatrixGlobalNumCols(M), hypre_ParCSRMatrixColStarts(M), y);
   hypre_ParCSRMatrixMatvecT(alpha, M, vx, beta, vy);
   take_vec(vy, y);
   hypre_ParVectorDestroy(vx); hypre_ParVectorDestroy(vy);
   return 0;
}

/* timed loop of HYPRE_ParCSRMatrixMatvec(1, A, x, 0, b), as `ij -solver -1 -nmv N` (ij.c:4758-4792) */
double rb_matvec_time(rb_problem *pb, int nmv)
{
   double t0;
   int k;
   HYPRE_ParCSRMatrixMatvec(1.0, pb->A, pb->x, 0.0, pb->b);
   t0 = wall();
   for (k = 0; k < nmv; k++) { HYPRE_ParCSRMatrixMatvec(1.0, pb->A, pb->x, 0.0, pb->b); }
   return (wall() - t0) / (double) nmv;
}

/* one hypre_BoomerAMGRelax sweep on level `level` (par_relax.c:23) */
int rb_relax(rb_problem *pb, int level, int relax_type, int relax_points, double relax_weight,
             double omega, const double *f, double *u, int u_all_zeros, int use_l1, int use_cf)
{
   hypre_ParAMGData *amg = (hypre_ParAMGData *) pb->amg;
   hypre_ParCSRMatrix *M = hypre_ParAMGDataAArray(amg)[level];
   HYPRE_BigInt g = hypre_ParCSRMatrixGlobalNumRows(M), *st = hypre_ParCSRMatrixRowStarts(M);
   hypre_ParVector *vf = make_vec(pb, g, st, f), *vu = make_vec(pb, g, st, u);
   hypre_ParVector *vt = make_vec(pb, g, st, NULL), *zt = make_vec(pb, g, st, NULL);
   double *l1 = use_l1 ? rb_level_l1_norms(pb, level) : NULL;
   int *cf = use_cf ? rb_level_cf_marker(pb, level) : NULL;
   hypre_ParVectorAllZeros(vu) = u_all_zeros;
   hypre_BoomerAMGRelax(M, vf, cf, relax_type, relax_points, relax_weight, omega, l1, vu, vt, zt);
   take_vec(vu, u);
   hypre_ParVectorDestroy(vf); hypre_ParVectorDestroy(vu); hypre_ParVectorDestroy(vt); hypre_ParVectorDestroy(zt);
   return (int) HYPRE_GetError();
}

/* hypre_ParCSRRelax_Cheby_Solve on level `level` with the setup's ds/coefs */
int rb_cheby(rb_problem *pb, int level, const double *f, double *u)
{
   hypre_ParAMGData *amg = (hypre_ParAMGData *) pb->amg;
   hypre_ParCSRMatrix *M = hypre_ParAMGDataAArray(amg)[level];
   HYPRE_BigInt g = hypre_ParCSRMatrixGlobalNumRows(M), *st = hypre_ParCSRMatrixRowStarts(M);
   hypre_ParVector *vf = make_vec(pb, g, st, f), *vu = make_vec(pb, g, st, u);
   hypre_ParVector *v1 = make_vec(pb, g, st, NULL), *v2 = make_vec(pb, g, st, NULL);
   hypre_ParVector *v3 = make_vec(pb, g, st, NULL), *v4 = make_vec(pb, g, st, NULL);
   hypre_ParCSRRelax_Cheby_Solve(M, vf, rb_level_cheby_ds(pb, level), rb_level_cheby_coefs(pb, level),
                                 hypre_ParAMGDataChebyOrder(amg), hypre_ParAMGDataChebyScale(amg),
                                 hypre_ParAMGDataChebyVariant(amg), vu, v1, v2, v3, v4);
   take_vec(vu, u);
   hypre_ParVectorDestroy(vf); hypre_ParVectorDestroy(vu); hypre_ParVectorDestroy(v1);
   hypre_ParVectorDestroy(v2); hypre_ParVectorDestroy(v3); hypre_ParVectorDestroy(v4);
   return (int) HYPRE_GetError();
}

/* HYPRE_BoomerAMGSolve(amg, A, f, u) — with ij's precond settings that is exactly one
 * hypre_BoomerAMGCycle (par_amg_solve.c:265).  Optionally returns level vectors. */
int rb_amg_solve(rb_problem *pb, const double *f, double *u, int u_all_zeros)
{
   hypre_ParCSRMatrix *M = (hypre_ParCSRMatrix *) pb->A;
   HYPRE_BigInt g = hypre_ParCSRMatrixGlobalNumRows(M), *st = hypre_ParCSRMatrixRowStarts(M);
   hypre_ParVector *vf = make_vec(pb, g, st, f), *vu = make_vec(pb, g, st, u);
   hypre_ParVectorAllZeros(vu) = u_all_zeros;
   HYPRE_BoomerAMGSolve(pb->amg, pb->A, (HYPRE_ParVector) vf, (HYPRE_ParVector) vu);
   take_vec(vu, u);
   hypre_ParVectorDestroy(vf); hypre_ParVectorDestroy(vu);
   return (int) HYPRE_GetError();
}

int rb_amg_set_solve(rb_problem *pb, double tol, int max_iter)
{
   HYPRE_BoomerAMGSetTol(pb->amg, tol);
   HYPRE_BoomerAMGSetMaxIter(pb->amg, max_iter);
   return 0;
}

/* after rb_amg_solve: copy F_array[level] / U_array[level] (level >= 1) */
int rb_level_vector(rb_problem *pb, int level, int which, double *out)
{
   hypre_ParAMGData *amg = (hypre_ParAMGData *) pb->amg;
   hypre_ParVector *v = which == 0 ? hypre_ParAMGDataFArray(amg)[level] : hypre_ParAMGDataUArray(amg)[level];
   take_vec(v, out);
   return 0;
}

/* ij -solver 1 / 2: PCG (ij.c:5563-5576): two_norm = 1, tol, max_iter = mg_max_iter (100) with AMG.
 * precond: 1 = BoomerAMG (pb->amg, already set up), 2 = diagonal scaling, 0 = none.
 * b and x0 are taken from the problem unless b_in / x_io given.  norms: max_iter+1 doubles. */
int rb_pcg_solve(rb_problem *pb, int precond, double tol, double atol, int max_iter, int two_norm,
                 int rel_change, int flex, int recompute_res, const double *b_in, double *x_io,
                 int *num_iterations, double *final_res_norm, double *norms, double *solve_seconds)
{
   HYPRE_Solver pcg;
   hypre_ParCSRMatrix *M = (hypre_ParCSRMatrix *) pb->A;
   HYPRE_BigInt g = hypre_ParCSRMatrixGlobalNumRows(M), *st = hypre_ParCSRMatrixRowStarts(M);
   hypre_ParVector *vb = make_vec(pb, g, st, b_in ? b_in : rb_problem_b(pb));
   hypre_ParVector *vx = make_vec(pb, g, st, x_io ? x_io : rb_problem_x(pb));
   HYPRE_Int its = 0;
   HYPRE_Real fr = 0.0;
   double t0;
   int k;
   HYPRE_ParCSRPCGCreate(pb->comm, &pcg);
   HYPRE_PCGSetMaxIter(pcg, max_iter);
   HYPRE_PCGSetTol(pcg, tol);
   HYPRE_PCGSetTwoNorm(pcg, two_norm);
   HYPRE_PCGSetFlex(pcg, flex);
   HYPRE_PCGSetRelChange(pcg, rel_change);
   HYPRE_PCGSetPrintLevel(pcg, 0);
   HYPRE_PCGSetLogging(pcg, 1);
   HYPRE_PCGSetAbsoluteTol(pcg, atol);
   HYPRE_PCGSetRecomputeResidual(pcg, recompute_res);
   if (precond == 1)
   {
      /* the hierarchy is already built (rb_amg_setup, timed separately): install a no-op
         setup so that HYPRE_PCGSetup does not run BoomerAMGSetup a second time */
      HYPRE_PCGSetPrecond(pcg, (HYPRE_PtrToSolverFcn) HYPRE_BoomerAMGSolve,
                          (HYPRE_PtrToSolverFcn) HYPRE_ParCSRDiagScaleSetup, pb->amg);
   }
   else if (precond == 2)
   {
      HYPRE_PCGSetPrecond(pcg, (HYPRE_PtrToSolverFcn) HYPRE_ParCSRDiagScale,
                          (HYPRE_PtrToSolverFcn) HYPRE_ParCSRDiagScaleSetup, NULL);
   }
   HYPRE_PCGSetup(pcg, (HYPRE_Matrix) pb->A, (HYPRE_Vector) vb, (HYPRE_Vector) vx);
   t0 = wall();
   HYPRE_PCGSolve(pcg, (HYPRE_Matrix) pb->A, (HYPRE_Vector) vb, (HYPRE_Vector) vx);
   if (solve_seconds) { *solve_seconds = wall() - t0; }
   HYPRE_PCGGetNumIterations(pcg, &its);
   HYPRE_PCGGetFinalRelativeResidualNorm(pcg, &fr);
   if (norms)
   {
      hypre_PCGData *pd = (hypre_PCGData *) pcg;
      for (k = 0; k <= its && k <= max_iter; k++) { norms[k] = pd->norms[k]; }
   }
   if (num_iterations) { *num_iterations = (int) its; }
   if (final_res_norm) { *final_res_norm = fr; }
   if (x_io) { take_vec(vx, x_io); }
   HYPRE_ParCSRPCGDestroy(pcg);
   hypre_ParVectorDestroy(vb); hypre_ParVectorDestroy(vx);
   k = (int) HYPRE_GetError();
   HYPRE_ClearAllErrors();
   return k;
}

/* ij -solver 3 / 4: GMRES (ij.c:7322-7330): k_dim = 5, logging 1 */
int rb_gmres_solve(rb_problem *pb, int precond, double tol, double atol, int max_iter, int k_dim,
                   int rel_change, const double *b_in, double *x_io, int *num_iterations,
                   double *final_res_norm, double *norms, double *solve_seconds)
{
   HYPRE_Solver gm;
   hypre_ParCSRMatrix *M = (hypre_ParCSRMatrix *) pb->A;
   HYPRE_BigInt g = hypre_ParCSRMatrixGlobalNumRows(M), *st = hypre_ParCSRMatrixRowStarts(M);
   hypre_ParVector *vb = make_vec(pb, g, st, b_in ? b_in : rb_problem_b(pb));
   hypre_ParVector *vx = make_vec(pb, g, st, x_io ? x_io : rb_problem_x(pb));
   HYPRE_Int its = 0;
   HYPRE_Real fr = 0.0;
   double t0;
   int k;
   HYPRE_ParCSRGMRESCreate(pb->comm, &gm);
   HYPRE_GMRESSetKDim(gm, k_dim);
   HYPRE_GMRESSetMaxIter(gm, max_iter);
   HYPRE_GMRESSetTol(gm, tol);
   HYPRE_GMRESSetAbsoluteTol(gm, atol);
   HYPRE_GMRESSetLogging(gm, 1);
   HYPRE_GMRESSetPrintLevel(gm, 0);
   HYPRE_GMRESSetRelChange(gm, rel_change);
   if (precond == 1)
   {
      HYPRE_GMRESSetPrecond(gm, (HYPRE_PtrToSolverFcn) HYPRE_BoomerAMGSolve,
                            (HYPRE_PtrToSolverFcn) HYPRE_ParCSRDiagScaleSetup, pb->amg);
   }
   else if (precond == 2)
   {
      HYPRE_GMRESSetPrecond(gm, (HYPRE_PtrToSolverFcn) HYPRE_ParCSRDiagScale,
                            (HYPRE_PtrToSolverFcn) HYPRE_ParCSRDiagScaleSetup, NULL);
   }
   HYPRE_GMRESSetup(gm, (HYPRE_Matrix) pb->A, (HYPRE_Vector) vb, (HYPRE_Vector) vx);
   t0 = wall();
   HYPRE_GMRESSolve(gm, (HYPRE_Matrix) pb->A, (HYPRE_Vector) vb, (HYPRE_Vector) vx);
   if (solve_seconds) { *solve_seconds = wall() - t0; }
   HYPRE_GMRESGetNumIterations(gm, &its);
   HYPRE_GMRESGetFinalRelativeResidualNorm(gm, &fr);
   if (norms)
   {
      hypre_GMRESData *gd = (hypre_GMRESData *) gm;
      for (k = 0; k <= its && k <= max_iter; k++) { norms[k] = gd->norms[k]; }
   }
   if (num_iterations) { *num_iterations = (int) its; }
   if (final_res_norm) { *final_res_norm = fr; }
   if (x_io) { take_vec(vx, x_io); }
   HYPRE_ParCSRGMRESDestroy(gm);
   hypre_ParVectorDestroy(vb); hypre_ParVectorDestroy(vx);
   k = (int) HYPRE_GetError();
   HYPRE_ClearAllErrors();
   return k;
}

/* ij -solver 9 / 10 (BiCGSTAB, ij.c:8660-8690), 16 / 17 (COGMRES, ij.c:9088-9110), 61 / 60 (FlexGMRES,
 * ij.c:8226-8245), 50 / 51 (LGMRES) over the same ParCSR table: which = 0 BiCGSTAB, 1 FlexGMRES, 2 COGMRES, 3 LGMRES;
 * cgs: COGMRES' variant, LGMRES' aug_dim */
int rb_krylov_ext_solve(rb_problem *pb, int which, int precond, double tol, double atol, int max_iter,
                        int k_dim, int cgs, int rel_change, const double *b_in, double *x_io,
                        int *num_iterations, double *final_res_norm, double *norms, double *solve_seconds)
{
   HYPRE_Solver ks;
   hypre_ParCSRMatrix *M = (hypre_ParCSRMatrix *) pb->A;
   HYPRE_BigInt g = hypre_ParCSRMatrixGlobalNumRows(M), *st = hypre_ParCSRMatrixRowStarts(M);
   hypre_ParVector *vb = make_vec(pb, g, st, b_in ? b_in : rb_problem_b(pb));
   hypre_ParVector *vx = make_vec(pb, g, st, x_io ? x_io : rb_problem_x(pb));
   HYPRE_PtrToSolverFcn pc = NULL, pcs = (HYPRE_PtrToSolverFcn) HYPRE_ParCSRDiagScaleSetup;
   HYPRE_Solver pcd = NULL;
   HYPRE_Int its = 0;
   HYPRE_Real fr = 0.0, *nr = NULL;
   double t0;
   int k;
   if (precond == 1) { pc = (HYPRE_PtrToSolverFcn) HYPRE_BoomerAMGSolve; pcd = pb->amg; }
   else if (precond == 2) { pc = (HYPRE_PtrToSolverFcn) HYPRE_ParCSRDiagScale; }
   if (which == 0)
   {
      HYPRE_ParCSRBiCGSTABCreate(pb->comm, &ks);
      HYPRE_BiCGSTABSetMaxIter(ks, max_iter);
      HYPRE_BiCGSTABSetTol(ks, tol);
      HYPRE_BiCGSTABSetAbsoluteTol(ks, atol);
      HYPRE_BiCGSTABSetLogging(ks, 1);
      HYPRE_BiCGSTABSetPrintLevel(ks, 0);
      if (pc) { HYPRE_BiCGSTABSetPrecond(ks, pc, pcs, pcd); }
      HYPRE_BiCGSTABSetup(ks, (HYPRE_Matrix) pb->A, (HYPRE_Vector) vb, (HYPRE_Vector) vx);
      t0 = wall();
      HYPRE_BiCGSTABSolve(ks, (HYPRE_Matrix) pb->A, (HYPRE_Vector) vb, (HYPRE_Vector) vx);
      if (solve_seconds) { *solve_seconds = wall() - t0; }
      HYPRE_BiCGSTABGetNumIterations(ks, &its);
      HYPRE_BiCGSTABGetFinalRelativeResidualNorm(ks, &fr);
      nr = ((hypre_BiCGSTABData *) ks)->norms;
   }
   else if (which == 1)
   {
      HYPRE_ParCSRFlexGMRESCreate(pb->comm, &ks);
      HYPRE_FlexGMRESSetKDim(ks, k_dim);
      HYPRE_FlexGMRESSetMaxIter(ks, max_iter);
      HYPRE_FlexGMRESSetTol(ks, tol);
      HYPRE_FlexGMRESSetAbsoluteTol(ks, atol);
      HYPRE_FlexGMRESSetLogging(ks, 1);
      HYPRE_FlexGMRESSetPrintLevel(ks, 0);
      if (pc) { HYPRE_FlexGMRESSetPrecond(ks, pc, pcs, pcd); }
      HYPRE_FlexGMRESSetup(ks, (HYPRE_Matrix) pb->A, (HYPRE_Vector) vb, (HYPRE_Vector) vx);
      t0 = wall();
      HYPRE_FlexGMRESSolve(ks, (HYPRE_Matrix) pb->A, (HYPRE_Vector) vb, (HYPRE_Vector) vx);
      if (solve_seconds) { *solve_seconds = wall() - t0; }
      HYPRE_FlexGMRESGetNumIterations(ks, &its);
      HYPRE_FlexGMRESGetFinalRelativeResidualNorm(ks, &fr);
      nr = ((hypre_FlexGMRESData *) ks)->norms;
   }
   else if (which == 3)
   {
      /* ij -solver 50 / 51 (ij.c:7990-8015); `cgs` carries aug_dim here */
      HYPRE_ParCSRLGMRESCreate(pb->comm, &ks);
      HYPRE_LGMRESSetKDim(ks, k_dim);
      HYPRE_LGMRESSetAugDim(ks, cgs);
      HYPRE_LGMRESSetMaxIter(ks, max_iter);
      HYPRE_LGMRESSetTol(ks, tol);
      HYPRE_LGMRESSetAbsoluteTol(ks, atol);
      HYPRE_LGMRESSetLogging(ks, 1);
      HYPRE_LGMRESSetPrintLevel(ks, 0);
      if (pc) { HYPRE_LGMRESSetPrecond(ks, pc, pcs, pcd); }
      HYPRE_LGMRESSetup(ks, (HYPRE_Matrix) pb->A, (HYPRE_Vector) vb, (HYPRE_Vector) vx);
      t0 = wall();
      HYPRE_LGMRESSolve(ks, (HYPRE_Matrix) pb->A, (HYPRE_Vector) vb, (HYPRE_Vector) vx);
      if (solve_seconds) { *solve_seconds = wall() - t0; }
      HYPRE_LGMRESGetNumIterations(ks, &its);
      HYPRE_LGMRESGetFinalRelativeResidualNorm(ks, &fr);
      nr = ((hypre_LGMRESData *) ks)->norms;
   }
   else
   {
      HYPRE_ParCSRCOGMRESCreate(pb->comm, &ks);
      HYPRE_COGMRESSetKDim(ks, k_dim);
      HYPRE_COGMRESSetCGS(ks, cgs);
      HYPRE_COGMRESSetMaxIter(ks, max_iter);
      HYPRE_COGMRESSetTol(ks, tol);
      HYPRE_COGMRESSetAbsoluteTol(ks, atol);
      HYPRE_COGMRESSetLogging(ks, 1);
      HYPRE_COGMRESSetPrintLevel(ks, 0);
      ((hypre_COGMRESData *) ks)->rel_change = rel_change;   /* no public setter in 3.1.0 */
      if (pc) { HYPRE_COGMRESSetPrecond(ks, pc, pcs, pcd); }
      HYPRE_COGMRESSetup(ks, (HYPRE_Matrix) pb->A, (HYPRE_Vector) vb, (HYPRE_Vector) vx);
      t0 = wall();
      HYPRE_COGMRESSolve(ks, (HYPRE_Matrix) pb->A, (HYPRE_Vector) vb, (HYPRE_Vector) vx);
      if (solve_seconds) { *solve_seconds = wall() - t0; }
      HYPRE_COGMRESGetNumIterations(ks, &its);
      HYPRE_COGMRESGetFinalRelativeResidualNorm(ks, &fr);
      nr = ((hypre_COGMRESData *) ks)->norms;
   }
   if (norms && nr)
   {
      /* BiCGSTAB logs every iteration; the GMRES family keeps norms[iter] only when printing (norms[0] always) */
      for (k = 0; k <= its && k <= max_iter; k++) { norms[k] = nr[k]; }
   }
   if (num_iterations) { *num_iterations = (int) its; }
   if (final_res_norm) { *final_res_norm = fr; }
   if (x_io) { take_vec(vx, x_io); }
   if (which == 0) { HYPRE_ParCSRBiCGSTABDestroy(ks); }
   else if (which == 1) { HYPRE_ParCSRFlexGMRESDestroy(ks); }
   else if (which == 3) { HYPRE_ParCSRLGMRESDestroy(ks); }
   else { HYPRE_ParCSRCOGMRESDestroy(ks); }
   hypre_ParVectorDestroy(vb); hypre_ParVectorDestroy(vx);
   k = (int) HYPRE_GetError();
   HYPRE_ClearAllErrors();
   return k;
}

/* reference BLAS-1 (hypre_ParVectorInnerProd / Axpy) on raw arrays of the fine-level size */
double rb_inner_prod(rb_problem *pb, const double *x, const double *y)
{
   hypre_ParCSRMatrix *M = (hypre_ParCSRMatrix *) pb->A;
   HYPRE_BigInt g = hypre_ParCSRMatrixGlobalNumRows(M), *st = hypre_ParCSRMatrixRowStarts(M);
   hypre_ParVector *vx = make_vec(pb, g, st, x), *vy = make_vec(pb, g, st, y);
   double r = hypre_ParVectorInnerProd(vx, vy);
   hypre_ParVectorDestroy(vx); hypre_ParVectorDestroy(vy);
   return r;
}
int rb_axpy(rb_problem *pb, double alpha, const double *x, double *y)
{
   hypre_ParCSRMatrix *M = (hypre_ParCSRMatrix *) pb->A;
   HYPRE_BigInt g = hypre_ParCSRMatrixGlobalNumRows(M), *st = hypre_ParCSRMatrixRowStarts(M);
   hypre_ParVector *vx = make_vec(pb, g, st, x), *vy = make_vec(pb, g, st, y);
   hypre_ParVectorAxpy(alpha, vx, vy);
   take_vec(vy, y);
   hypre_ParVectorDestroy(vx); hypre_ParVectorDestroy(vy);
   return 0;
}

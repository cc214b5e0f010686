/*
 * hypre_oracle.c — CPU restatement (plain C, one rank, one thread) of the reference's
 * BoomerAMG solve path.  TEST INFRASTRUCTURE ONLY: nothing in hypre_b200/ links, imports or
 * calls this file; tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg are the only
 * users.  Every function cites the reference lines (hypre 3.1.0, /root/reference/src) it
 * restates; the restatement keeps the reference's operation order so that, compiled without FMA
 * contraction, it reproduces the reference bit for bit on one thread.
 *
 * Pinning: tests/test_oracle_golden.py checks this file against tests/golden/ (npz files), fixtures
 * produced by the compiled reference itself (tests/golden/make_golden.py, run where
 * /root/reference exists) — SpMV / relax / cycle outputs, PCG / GMRES residual histories —
 * and against the iteration counts the reference's own regression suite pins
 * (src/test/TEST_ij/solvers.saved).
 *
 * Scope: 1 MPI rank (no offd block).  The multi-rank reference behaviour is pinned through the
 * compiled reference running on oracle/minimpi instead.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct
{
   int           nrows, ncols;
   const int    *i, *j;
   const double *a;
} ho_csr;

typedef struct
{
   ho_csr        A, P;
   int           has_P;
   const double *l1;
   const int    *cf;
   double        relax_weight, omega;
   const double *cheby_ds, *cheby_coefs;
   double       *F, *U, *Vtemp, *Ztemp, *Ptemp, *Rtemp;   /* level work vectors (owned) */
   int           u_zero;
} ho_level;

typedef struct
{
   int       num_levels;
   ho_level *lev;
   int       num_grid_sweeps[4], grid_relax_type[4];
   int       relax_order, cycle_type, fcycle, cheby_order, cheby_scale, user_relax_type;
   double    tol;
   int       min_iter, max_iter, converge_type;
   double   *A_mat;   /* coarsest dense matrix, row-major n x n (par_gauss_elim.c:203-221) */
   int       ge_n;
} ho_amg;

/* ---------------------------------------------------------------------------------------------
 * hypre_CSRMatrixMatvecOutOfPlaceHost, single vector, no rownnz (src/seq_mv/csr_matvec.c:683-845)
 * y = alpha*A*x + beta*b with the reference's branch structure on temp = beta/alpha.
 * ------------------------------------------------------------------------------------------- */
void ho_csr_matvec(int nrows, const int *ai, const int *aj, const double *aa, double alpha,
                   const double *x, double beta, const double *b, double *y)
{
   int r, p;
   double temp, tempx;
   if (alpha == 0.0)   /* csr_matvec.c:92-105 */
   {
      for (r = 0; r < nrows; r++) { y[r] = beta * b[r]; }
      return;
   }
   temp = beta / alpha;   /* csr_matvec.c:115 */
   for (r = 0; r < nrows; r++)
   {
      if (temp == 0.0)
      {
         tempx = 0.0;
         if (alpha == -1.0) { for (p = ai[r]; p < ai[r + 1]; p++) { tempx -= aa[p] * x[aj[p]]; } y[r] = tempx; }
         else
         {
            for (p = ai[r]; p < ai[r + 1]; p++) { tempx += aa[p] * x[aj[p]]; }
            y[r] = (alpha == 1.0) ? tempx : alpha * tempx;
         }
      }
      else if (temp == -1.0)
      {
         if (alpha == 1.0) { y[r] = -b[r]; tempx = 0.0; for (p = ai[r]; p < ai[r + 1]; p++) { tempx += aa[p] * x[aj[p]]; } y[r] += tempx; }
         else if (alpha == -1.0) { y[r] = b[r]; tempx = 0.0; for (p = ai[r]; p < ai[r + 1]; p++) { tempx -= aa[p] * x[aj[p]]; } y[r] += tempx; }
         else { y[r] = -alpha * b[r]; tempx = 0.0; for (p = ai[r]; p < ai[r + 1]; p++) { tempx += aa[p] * x[aj[p]]; } y[r] += alpha * tempx; }
      }
      else if (temp == 1.0)
      {
         if (alpha == 1.0) { y[r] = b[r]; tempx = 0.0; for (p = ai[r]; p < ai[r + 1]; p++) { tempx += aa[p] * x[aj[p]]; } y[r] += tempx; }
         else if (alpha == -1.0) { y[r] = -b[r]; tempx = 0.0; for (p = ai[r]; p < ai[r + 1]; p++) { tempx -= aa[p] * x[aj[p]]; } y[r] += tempx; }
         else { y[r] = alpha * b[r]; tempx = 0.0; for (p = ai[r]; p < ai[r + 1]; p++) { tempx += aa[p] * x[aj[p]]; } y[r] += alpha * tempx; }
      }
      else
      {
         if (alpha == 1.0) { y[r] = b[r] * temp; tempx = 0.0; for (p = ai[r]; p < ai[r + 1]; p++) { tempx += aa[p] * x[aj[p]]; } y[r] += tempx; }
         else if (alpha == -1.0) { y[r] = -b[r] * temp; tempx = 0.0; for (p = ai[r]; p < ai[r + 1]; p++) { tempx -= aa[p] * x[aj[p]]; } y[r] += tempx; }
         else { y[r] = b[r] * beta; tempx = 0.0; for (p = ai[r]; p < ai[r + 1]; p++) { tempx += aa[p] * x[aj[p]]; } y[r] += alpha * tempx; }
      }
   }
}

/* hypre_CSRMatrixMatvecTHost, one thread (src/seq_mv/csr_matvec.c:914-1139): y = alpha*A^T*x + beta*y;
 * y *= beta/alpha, scatter-add in row order, y *= alpha. */
void ho_csr_matvecT(int nrows, int ncols, const int *ai, const int *aj, const double *aa, double alpha,
                    const double *x, double beta, double *y)
{
   int r, p, c;
   double temp;
   if (alpha == 0.0) { for (c = 0; c < ncols; c++) { y[c] *= beta; } return; }
   temp = beta / alpha;
   if (temp != 1.0)
   {
      if (temp == 0.0) { for (c = 0; c < ncols; c++) { y[c] = 0.0; } }
      else { for (c = 0; c < ncols; c++) { y[c] *= temp; } }
   }
   for (r = 0; r < nrows; r++)
   {
      for (p = ai[r]; p < ai[r + 1]; p++) { y[aj[p]] += aa[p] * x[r]; }
   }
   if (alpha != 1.0) { for (c = 0; c < ncols; c++) { y[c] *= alpha; } }
}

/* hypre_SeqVectorInnerProdHost (src/seq_mv/vector.c:1303-1358), one thread */
double ho_inner_prod(int n, const double *x, const double *y)
{
   int k;
   double r = 0.0;
   for (k = 0; k < n; k++) { r += y[k] * x[k]; }
   return r;
}
/* hypre_SeqVectorAxpyHost (vector.c:980-1010) */
void ho_axpy(int n, double alpha, const double *x, double *y)
{
   int k;
   for (k = 0; k < n; k++) { y[k] += alpha * x[k]; }
}
/* hypre_SeqVectorScaleHost (vector.c:905-945) */
void ho_scale(int n, double alpha, double *y)
{
   int k;
   if (alpha == 1.0) { return; }
   if (alpha == 0.0) { for (k = 0; k < n; k++) { y[k] = 0.0; } return; }
   for (k = 0; k < n; k++) { y[k] *= alpha; }
}

/* ---------------------------------------------------------------------------------------------
 * relaxation
 * ------------------------------------------------------------------------------------------- */

/* hypre_BoomerAMGRelaxWeightedJacobi_core (src/parcsr_ls/par_relax.c:180-314), one rank */
static void jacobi_core(const ho_csr *A, const double *f, const int *cf, int relax_points, double w,
                        const double *l1, double *u, double *vtemp, int skip_diag)
{
   int r, p, n = A->nrows;
   const double one_minus_weight = 1.0 - w;
   memcpy(vtemp, u, sizeof(double) * (size_t) n);
   for (r = 0; r < n; r++)
   {
      const double di = l1 ? l1[r] : A->a[A->i[r]];
      if ((relax_points == 0 || cf[r] == relax_points) && di != 0.0)
      {
         double res = f[r];
         for (p = A->i[r] + skip_diag; p < A->i[r + 1]; p++) { res -= A->a[p] * vtemp[A->j[p]]; }
         if (skip_diag) { u[r] *= one_minus_weight; u[r] += w * res / di; }
         else { u[r] += w * res / di; }
      }
   }
}

/* hypre_BoomerAMGRelax7Jacobi (par_relax.c:1178-1254): Vtemp = f; Vtemp = w*f - w*A*u (or w*f when
 * u is flagged all-zero); u += Vtemp ./ l1 [marked] */
static void jacobi7(const ho_csr *A, const double *f, const int *cf, int relax_points, double w,
                    const double *l1, double *u, int u_all_zeros, double *vtemp)
{
   int r, n = A->nrows;
   memcpy(vtemp, f, sizeof(double) * (size_t) n);
   if (u_all_zeros) { ho_scale(n, w, vtemp); }
   else { ho_csr_matvec(n, A->i, A->j, A->a, -w, u, w, vtemp, vtemp); }
   for (r = 0; r < n; r++)
   {
      if (relax_points == 0 || cf[r] == relax_points) { u[r] += vtemp[r] / l1[r]; }
   }
}

/* hypre_HybridGaussSeidelNS / hypre_HybridGaussSeidel (src/parcsr_ls/par_relax.h:12-110, 238-330),
 * driven as in hypre_BoomerAMGRelaxHybridGaussSeidel_core, num_threads == 1 (par_relax.c:905-936) */
static void hybrid_gs(const ho_csr *A, const double *f, const int *cf, int relax_points, double w,
                      double omega, const double *l1, double *u, double *vtemp, int gs_order, int symm,
                      int skip_diag)
{
   const int n = A->nrows;
   const int non_scale = (w == 1.0 && omega == 1.0);
   const double one_minus_omega = 1.0 - omega, prod = 1.0 - w * omega;
   const int num_sweeps = symm ? 2 : 1;
   int sweep, r, p;
   if (!non_scale) { memcpy(vtemp, u, sizeof(double) * (size_t) n); }
   for (sweep = 0; sweep < num_sweeps; sweep++)
   {
      const int iorder = num_sweeps == 1 ? (gs_order > 0 ? 1 : -1) : (sweep == 0 ? 1 : -1);
      const int ibegin = iorder > 0 ? 0 : n - 1, iend = iorder > 0 ? n : -1;
      for (r = ibegin; r != iend; r += iorder)
      {
         const double d = l1 ? l1[r] : A->a[A->i[r]];
         if (!((relax_points == 0 || cf[r] == relax_points) && d != 0.0)) { continue; }
         if (non_scale)
         {
            double res = f[r];
            for (p = A->i[r] + skip_diag; p < A->i[r + 1]; p++) { res -= A->a[p] * u[A->j[p]]; }
            if (skip_diag) { u[r] = res / d; } else { u[r] += res / d; }
         }
         else
         {
            double res = f[r], res0 = 0.0, res2 = 0.0;
            for (p = A->i[r] + skip_diag; p < A->i[r + 1]; p++)
            {
               res0 -= A->a[p] * u[A->j[p]];
               res2 += A->a[p] * vtemp[A->j[p]];
            }
            if (skip_diag) { u[r] *= prod; }
            u[r] += w * (omega * res + res0 + one_minus_omega * res2) / d;
         }
      }
   }
}

/* hypre_BoomerAMGRelax dispatch (par_relax.c:23-173) for the types on the accelerated path */
int ho_relax(const ho_csr *A, const double *f, const int *cf, int relax_type, int relax_points, double w,
             double omega, const double *l1, double *u, int u_all_zeros, double *vtemp)
{
   const int skip_l1 = (w == 1.0 && omega == 1.0) ? 0 : 1;
   switch (relax_type)
   {
      case 0:  jacobi_core(A, f, cf, relax_points, w, NULL, u, vtemp, 1); break;
      case 7:  jacobi7(A, f, cf, relax_points, w, l1, u, u_all_zeros, vtemp); break;
      case 18:   /* par_relax.c:355-368 */
         if (relax_points == 0) { jacobi7(A, f, cf, relax_points, w, l1, u, u_all_zeros, vtemp); }
         else { jacobi_core(A, f, cf, relax_points, w, l1, u, vtemp, 0); }
         break;
      case 3:  hybrid_gs(A, f, cf, relax_points, w, omega, NULL, u, vtemp, 1, 0, 1); break;
      case 4:  hybrid_gs(A, f, cf, relax_points, w, omega, NULL, u, vtemp, -1, 0, 1); break;
      case 6:  hybrid_gs(A, f, cf, relax_points, w, omega, NULL, u, vtemp, 1, 1, 1); break;
      case 8:
      case 88: hybrid_gs(A, f, cf, relax_points, w, omega, l1, u, vtemp, 1, 1, skip_l1); break;
      case 13: hybrid_gs(A, f, cf, relax_points, w, omega, l1, u, vtemp, 1, 0, skip_l1); break;
      case 14: hybrid_gs(A, f, cf, relax_points, w, omega, l1, u, vtemp, -1, 0, skip_l1); break;
      case 89:
         hybrid_gs(A, f, cf, relax_points, w, omega, l1, u, vtemp, 1, 0, skip_l1);
         hybrid_gs(A, f, cf, relax_points, w, omega, l1, u, vtemp, -1, 0, skip_l1);
         break;
      default: return 1;
   }
   return 0;
}

int ho_relax_raw(int nrows, const int *ai, const int *aj, const double *aa, const double *f, const int *cf,
                 int relax_type, int relax_points, double w, double omega, const double *l1, double *u,
                 int u_all_zeros, double *vtemp)
{
   ho_csr A;
   A.nrows = nrows; A.ncols = nrows; A.i = ai; A.j = aj; A.a = aa;
   return ho_relax(&A, f, cf, relax_type, relax_points, w, omega, l1, u, u_all_zeros, vtemp);
}

/* hypre_ParCSRRelax_Cheby_SolveHost (src/parcsr_ls/par_cheby_solve.c:194-345) */
void ho_cheby(const ho_csr *A, const double *f, const double *ds, const double *coefs, int order, int scale,
              double *u, double *v, double *r, double *orig_u, double *tmp)
{
   const int n = A->nrows;
   int i, k, cheby_order;
   double mult;
   if (order > 4) { order = 4; }
   if (order < 1) { order = 1; }
   cheby_order = order - 1;
   if (!scale)
   {
      ho_csr_matvec(n, A->i, A->j, A->a, -1.0, u, 1.0, f, r);
      for (k = 0; k < n; k++) { orig_u[k] = u[k]; u[k] = r[k] * coefs[cheby_order]; }
      for (i = cheby_order - 1; i >= 0; i--)
      {
         ho_csr_matvec(n, A->i, A->j, A->a, 1.0, u, 0.0, v, v);
         mult = coefs[i];
         for (k = 0; k < n; k++) { u[k] = mult * r[k] + v[k]; }
      }
      for (k = 0; k < n; k++) { u[k] = orig_u[k] + u[k]; }
   }
   else
   {
      ho_csr_matvec(n, A->i, A->j, A->a, -1.0, u, 0.0, tmp, tmp);
      for (k = 0; k < n; k++) { r[k] = ds[k] * (f[k] + tmp[k]); }
      for (k = 0; k < n; k++) { orig_u[k] = u[k]; u[k] = r[k] * coefs[cheby_order]; }
      for (i = cheby_order - 1; i >= 0; i--)
      {
         for (k = 0; k < n; k++) { tmp[k] = ds[k] * u[k]; }
         ho_csr_matvec(n, A->i, A->j, A->a, 1.0, tmp, 0.0, v, v);
         mult = coefs[i];
         for (k = 0; k < n; k++) { u[k] = mult * r[k] + ds[k] * v[k]; }
      }
      for (k = 0; k < n; k++) { u[k] = orig_u[k] + ds[k] * u[k]; }
   }
}

int ho_cheby_raw(int nrows, const int *ai, const int *aj, const double *aa, const double *f, const double *ds,
                 const double *coefs, int order, int scale, double *u)
{
   ho_csr A;
   double *w = (double *) malloc(sizeof(double) * 4 * (size_t) (nrows ? nrows : 1));
   A.nrows = nrows; A.ncols = nrows; A.i = ai; A.j = aj; A.a = aa;
   ho_cheby(&A, f, ds, coefs, order, scale, u, w, w + nrows, w + 2 * nrows, w + 3 * nrows);
   free(w);
   return 0;
}

/* hypre_gselim (src/utilities/gselim.h:11-66): unpivoted elimination on a row-major copy */
void ho_gselim(double *A, double *x, int n)
{
   int j, k, m;
   double factor, divA;
   if (n == 1) { if (A[0] != 0.0) { x[0] = x[0] / A[0]; } return; }
   for (k = 0; k < n - 1; k++)
   {
      if (A[k * n + k] != 0.0)
      {
         divA = 1.0 / A[k * n + k];
         for (j = k + 1; j < n; j++)
         {
            if (A[j * n + k] != 0.0)
            {
               factor = A[j * n + k] * divA;
               for (m = k + 1; m < n; m++) { A[j * n + m] -= factor * A[k * n + m]; }
               x[j] -= factor * x[k];
            }
         }
      }
   }
   for (k = n - 1; k > 0; --k)
   {
      if (A[k * n + k] != 0.0)
      {
         x[k] /= A[k * n + k];
         for (j = 0; j < k; j++) { if (A[j * n + k] != 0.0) { x[j] -= x[k] * A[j * n + k]; } }
      }
   }
   if (A[0] != 0.0) { x[0] /= A[0]; }
}

/* ---------------------------------------------------------------------------------------------
 * hierarchy + cycle
 * ------------------------------------------------------------------------------------------- */

ho_amg *ho_amg_create(int num_levels)
{
   ho_amg *amg = (ho_amg *) calloc(1, sizeof(ho_amg));
   int k;
   amg->num_levels = num_levels;
   amg->lev = (ho_level *) calloc((size_t) num_levels, sizeof(ho_level));
   for (k = 0; k < 4; k++) { amg->num_grid_sweeps[k] = 1; amg->grid_relax_type[k] = 18; }
   amg->grid_relax_type[3] = 9;
   amg->cycle_type = 1; amg->cheby_order = 2; amg->cheby_scale = 1; amg->user_relax_type = -1;
   amg->tol = 0.0; amg->max_iter = 1;
   return amg;
}

void ho_amg_destroy(ho_amg *amg)
{
   int l;
   if (!amg) { return; }
   for (l = 0; l < amg->num_levels; l++)
   {
      ho_level *L = &amg->lev[l];
      if (l > 0) { free(L->F); free(L->U); }
      free(L->Vtemp); free(L->Ztemp); free(L->Ptemp); free(L->Rtemp);
   }
   free(amg->A_mat);
   free(amg->lev);
   free(amg);
}

/* arrays are borrowed (must outlive the hierarchy) */
void ho_amg_set_level(ho_amg *amg, int l, int nrows, const int *ai, const int *aj, const double *aa,
                      int p_ncols, const int *pi, const int *pj, const double *pa, const double *l1,
                      const int *cf, double relax_weight, double omega, const double *cheby_ds,
                      const double *cheby_coefs)
{
   ho_level *L = &amg->lev[l];
   size_t n = (size_t) (nrows ? nrows : 1);
   L->A.nrows = nrows; L->A.ncols = nrows; L->A.i = ai; L->A.j = aj; L->A.a = aa;
   L->has_P = pi != NULL;
   if (L->has_P) { L->P.nrows = nrows; L->P.ncols = p_ncols; L->P.i = pi; L->P.j = pj; L->P.a = pa; }
   L->l1 = l1; L->cf = cf; L->relax_weight = relax_weight; L->omega = omega;
   L->cheby_ds = cheby_ds; L->cheby_coefs = cheby_coefs;
   if (l > 0) { L->F = (double *) calloc(n, sizeof(double)); L->U = (double *) calloc(n, sizeof(double)); }
   L->Vtemp = (double *) calloc(n, sizeof(double)); L->Ztemp = (double *) calloc(n, sizeof(double));
   L->Ptemp = (double *) calloc(n, sizeof(double)); L->Rtemp = (double *) calloc(n, sizeof(double));
}

void ho_amg_set_params(ho_amg *amg, const int *ngs4, const int *grt4, int relax_order, int cycle_type,
                       int fcycle, int cheby_order, int cheby_scale, int user_relax_type, double tol,
                       int min_iter, int max_iter, int converge_type)
{
   int k;
   for (k = 0; k < 4; k++) { amg->num_grid_sweeps[k] = ngs4[k]; amg->grid_relax_type[k] = grt4[k]; }
   amg->relax_order = relax_order; amg->cycle_type = cycle_type; amg->fcycle = fcycle;
   amg->cheby_order = cheby_order; amg->cheby_scale = cheby_scale; amg->user_relax_type = user_relax_type;
   amg->tol = tol; amg->min_iter = min_iter; amg->max_iter = max_iter; amg->converge_type = converge_type;
}

void ho_amg_set_coarse_ge(ho_amg *amg, const double *A_mat, int n)
{
   free(amg->A_mat);
   amg->A_mat = (double *) malloc(sizeof(double) * (size_t) n * (size_t) n);
   memcpy(amg->A_mat, A_mat, sizeof(double) * (size_t) n * (size_t) n);
   amg->ge_n = n;
}

/* one relaxation "sweep" as the cycle issues it: GE / Chebyshev / hypre_BoomerAMGRelaxIF
 * (par_cycle.c:535-645, par_relax_interface.c:19-65) */
static int cycle_relax(ho_amg *amg, int level, int relax_type, int cycle_param, int one_level)
{
   ho_level *L = &amg->lev[level];
   int q;
   if (relax_type == 9 || relax_type == 19)
   {
      /* hypre_GaussElimSolve, types 9/19, one rank (par_gauss_elim.c:640-652) */
      int n = amg->ge_n;
      double *work = (double *) malloc(sizeof(double) * (size_t) n * (size_t) n);
      double *b = (double *) malloc(sizeof(double) * (size_t) n);
      memcpy(work, amg->A_mat, sizeof(double) * (size_t) n * (size_t) n);
      memcpy(b, L->F, sizeof(double) * (size_t) n);
      ho_gselim(work, b, n);
      memcpy(L->U, b, sizeof(double) * (size_t) n);
      free(work); free(b);
      L->u_zero = 0;
      return 0;
   }
   if (relax_type == 16)
   {
      ho_cheby(&L->A, L->F, L->cheby_ds, L->cheby_coefs, amg->cheby_order, amg->cheby_scale, L->U,
               L->Vtemp, L->Ztemp, L->Ptemp, L->Rtemp);
      L->u_zero = 0;
      return 0;
   }
   if (!one_level && amg->relax_order == 1 && cycle_param < 3)
   {
      int pts[2];
      pts[0] = cycle_param < 2 ? 1 : -1; pts[1] = -pts[0];
      for (q = 0; q < 2; q++)
      {
         if (ho_relax(&L->A, L->F, L->cf, relax_type, pts[q], L->relax_weight, L->omega, L->l1, L->U,
                      L->u_zero, L->Vtemp)) { return 1; }
         L->u_zero = 0;   /* par_relax.c:170 */
      }
      return 0;
   }
   if (ho_relax(&L->A, L->F, L->cf, relax_type, 0, L->relax_weight, L->omega, L->l1, L->U, L->u_zero, L->Vtemp)) { return 1; }
   L->u_zero = 0;
   return 0;
}

/* hypre_BoomerAMGCycle (src/parcsr_ls/par_cycle.c:23-889): level-counter state machine
 * (:225-243), relaxation (:439-659), residual + restriction (:742-790), interpolation (:815-843) */
int ho_amg_cycle(ho_amg *amg, const double *f, double *u, int u_all_zeros)
{
   const int nl = amg->num_levels;
   int *lev_counter = (int *) calloc((size_t) nl, sizeof(int));
   int k, j, level = 0, cycle_param = 1, not_finished = 1, fcycle_lev = nl - 2, err = 0;
   amg->lev[0].F = (double *) f;
   amg->lev[0].U = u;
   amg->lev[0].u_zero = u_all_zeros;
   lev_counter[0] = 1;
   for (k = 1; k < nl; k++) { lev_counter[k] = amg->fcycle ? 1 : amg->cycle_type; }
   while (not_finished && !err)
   {
      int num_sweep, relax_type, one_level = 0;
      if (nl > 1) { num_sweep = amg->num_grid_sweeps[cycle_param]; relax_type = amg->grid_relax_type[cycle_param]; }
      else
      {
         num_sweep = amg->num_grid_sweeps[0];
         relax_type = amg->user_relax_type == -1 ? 6 : amg->user_relax_type;   /* par_cycle.c:349-356 */
         one_level = 1;
      }
      for (j = 0; j < num_sweep && !err; j++) { err = cycle_relax(amg, level, relax_type, cycle_param, one_level); }
      --lev_counter[level];
      if (lev_counter[level] >= 0 && level != nl - 1)
      {
         ho_level *Lf = &amg->lev[level], *Lc = &amg->lev[level + 1];
         memset(Lc->U, 0, sizeof(double) * (size_t) Lc->A.nrows);   /* hypre_ParVectorSetZeros */
         Lc->u_zero = 1;
         ho_csr_matvec(Lf->A.nrows, Lf->A.i, Lf->A.j, Lf->A.a, -1.0, Lf->U, 1.0, Lf->F, Lf->Vtemp);
         ho_csr_matvecT(Lf->P.nrows, Lf->P.ncols, Lf->P.i, Lf->P.j, Lf->P.a, 1.0, Lf->Vtemp, 0.0, Lc->F);
         ++level;
         if (lev_counter[level] < amg->cycle_type) { lev_counter[level] = amg->cycle_type; }
         cycle_param = (level == nl - 1) ? 3 : 1;
      }
      else if (level != 0)
      {
         ho_level *Lf = &amg->lev[level - 1], *Lc = &amg->lev[level];
         ho_csr_matvec(Lf->P.nrows, Lf->P.i, Lf->P.j, Lf->P.a, 1.0, Lc->U, 1.0, Lf->U, Lf->U);
         Lf->u_zero = 0;   /* par_cycle.c:843 */
         --level;
         cycle_param = 2;
         if (amg->fcycle && fcycle_lev == level)
         {
            if (lev_counter[level] < 1) { lev_counter[level] = 1; }
            fcycle_lev--;
         }
      }
      else { not_finished = 0; }
   }
   free(lev_counter);
   return err;
}

/* hypre_BoomerAMGSolve (src/parcsr_ls/par_amg_solve.c:22-424) without printing */
int ho_amg_solve(ho_amg *amg, const double *f, double *u, int u_all_zeros, int *num_iterations, double *rel_resid)
{
   ho_level *L0 = &amg->lev[0];
   const int n = L0->A.nrows;
   double resid_nrm = 1.0, resid_init = 0.0, rhs_norm = 0.0, relative_resid = 1.0;
   int cycle_count = 0, flag = 0;
   if (amg->tol > 0.0)
   {
      ho_csr_matvec(n, L0->A.i, L0->A.j, L0->A.a, 1.0, u, -1.0, f, L0->Vtemp);
      resid_nrm = sqrt(ho_inner_prod(n, L0->Vtemp, L0->Vtemp));
      resid_init = resid_nrm;
      if (amg->converge_type == 0)
      {
         rhs_norm = sqrt(ho_inner_prod(n, f, f));
         relative_resid = rhs_norm != 0.0 ? resid_init / rhs_norm : resid_init;
      }
      else { relative_resid = 1.0; }
   }
   while ((relative_resid >= amg->tol || cycle_count < amg->min_iter) && cycle_count < amg->max_iter)
   {
      ho_amg_cycle(amg, f, u, u_all_zeros);
      u_all_zeros = 0;
      if (amg->tol > 0.0)
      {
         ho_csr_matvec(n, L0->A.i, L0->A.j, L0->A.a, 1.0, u, -1.0, f, L0->Vtemp);
         resid_nrm = sqrt(ho_inner_prod(n, L0->Vtemp, L0->Vtemp));
         if (amg->converge_type == 0) { relative_resid = rhs_norm != 0.0 ? resid_nrm / rhs_norm : resid_nrm; }
         else { relative_resid = resid_nrm / resid_init; }
      }
      ++cycle_count;
   }
   if (cycle_count == amg->max_iter && amg->tol > 0.0) { flag = 256; }
   if (num_iterations) { *num_iterations = cycle_count; }
   if (rel_resid) { *rel_resid = relative_resid; }
   return flag;
}

void ho_amg_level_vector(ho_amg *amg, int level, int which, double *out)
{
   ho_level *L = &amg->lev[level];
   memcpy(out, which == 0 ? L->F : L->U, sizeof(double) * (size_t) L->A.nrows);
}

/* ---------------------------------------------------------------------------------------------
 * Krylov: precond kinds 0 = identity copy (hypre_ParKrylovIdentity), 1 = BoomerAMG, 2 = diagonal
 * scaling (hypre_ParCSRDiagScaleVector)
 * ------------------------------------------------------------------------------------------- */
static void precond(int kind, ho_amg *amg, const ho_csr *A, const double *r, double *z)
{
   int k, n = A->nrows;
   if (kind == 1) { memset(z, 0, sizeof(double) * (size_t) n); ho_amg_solve(amg, r, z, 1, NULL, NULL); }
   else if (kind == 2) { for (k = 0; k < n; k++) { z[k] = r[k] / A->a[A->i[k]]; } }
   else { memcpy(z, r, sizeof(double) * (size_t) n); }
}

/* hypre_PCGSolve (src/krylov/pcg.c:313-1016), default options + two_norm / flex / rel_change /
 * recompute_residual; returns the hypre error flag (256 = not converged) */
int ho_pcg(int nrows, const int *ai, const int *aj, const double *aa, int precond_kind, ho_amg *amg,
           double tol, double a_tol, int max_iter, int two_norm, int rel_change, int flex, int recompute_residual,
           const double *b, double *x, int *num_iterations, double *rel_residual_norm, double *norms)
{
   ho_csr A;
   const int n = nrows;
   double *p = (double *) calloc((size_t) (n ? n : 1), sizeof(double)), *s = (double *) calloc((size_t) (n ? n : 1), sizeof(double));
   double *r = (double *) calloc((size_t) (n ? n : 1), sizeof(double)), *r_old = (double *) calloc((size_t) (n ? n : 1), sizeof(double));
   double alpha = 0.0, beta, delta = 0.0, gamma, gamma_old, bi_prod, eps, i_prod = 0.0, i_prod_0 = 0.0, sdotp;
   int i = 0, tentatively_converged = 0, flag = 0, converged = 0;
   A.nrows = n; A.ncols = n; A.i = ai; A.j = aj; A.a = aa;
   if (two_norm) { bi_prod = ho_inner_prod(n, b, b); }
   else { precond(precond_kind, amg, &A, b, p); bi_prod = ho_inner_prod(n, p, b); }
   eps = tol * tol;
   if (bi_prod > 0.0) { eps = fmax(tol * tol, a_tol * a_tol / bi_prod); }   /* pcg.c:452-475, default branch */
   else
   {
      memcpy(x, b, sizeof(double) * (size_t) n);
      if (norms) { norms[0] = 0.0; }
      *num_iterations = 0; *rel_residual_norm = 0.0;
      free(p); free(s); free(r); free(r_old);
      return 0;
   }
   memcpy(r, b, sizeof(double) * (size_t) n);
   ho_csr_matvec(n, ai, aj, aa, -1.0, x, 1.0, r, r);
   precond(precond_kind, amg, &A, r, p);
   gamma = ho_inner_prod(n, r, p);
   i_prod_0 = two_norm ? ho_inner_prod(n, r, r) : gamma;
   if (norms) { norms[0] = sqrt(i_prod_0); }
   while ((i + 1) <= max_iter)
   {
      i++;
      ho_csr_matvec(n, ai, aj, aa, 1.0, p, 0.0, s, s);
      sdotp = ho_inner_prod(n, s, p);
      if (sdotp == 0.0) { flag |= 256; if (i == 1) { i_prod = i_prod_0; } break; }
      alpha = gamma / sdotp;
      if (!(alpha >= DBL_MIN)) { flag |= 256; if (i == 1) { i_prod = i_prod_0; } break; }   /* skip_break = 0 */
      gamma_old = gamma;
      ho_axpy(n, alpha, p, x);
      if (flex) { memcpy(r_old, r, sizeof(double) * (size_t) n); }
      ho_axpy(n, -alpha, s, r);
      precond(precond_kind, amg, &A, r, s);
      gamma = ho_inner_prod(n, r, s);
      if (flex) { delta = gamma - ho_inner_prod(n, r_old, s); }
      i_prod = two_norm ? ho_inner_prod(n, r, r) : gamma;
      if (norms) { norms[i] = sqrt(i_prod); }
      if (i_prod / bi_prod < eps) { tentatively_converged = 1; }
      if (tentatively_converged && recompute_residual)
      {
         memcpy(r, b, sizeof(double) * (size_t) n);
         ho_csr_matvec(n, ai, aj, aa, -1.0, x, 1.0, r, r);
         if (two_norm) { i_prod = ho_inner_prod(n, r, r); }
         else { precond(precond_kind, amg, &A, r, s); i_prod = ho_inner_prod(n, r, s); gamma = i_prod; }
         if (i_prod / bi_prod >= eps) { tentatively_converged = 0; }
      }
      if (tentatively_converged && rel_change && i_prod > 0.0)
      {
         double pi_prod = ho_inner_prod(n, p, p), xi_prod = ho_inner_prod(n, x, x);
         if (alpha * alpha * pi_prod / xi_prod >= eps) { tentatively_converged = 0; }
      }
      if (tentatively_converged) { converged = 1; break; }
      if (!(gamma >= DBL_MIN)) { flag |= 256; if (i == 1) { i_prod = i_prod_0; } break; }
      beta = flex ? delta / gamma_old : gamma / gamma_old;
      ho_scale(n, beta, p);
      ho_axpy(n, 1.0, s, p);
   }
   if (i >= max_iter && (i_prod / bi_prod) >= eps && eps > 0) { flag |= 256; }
   (void) converged;
   *num_iterations = i;
   *rel_residual_norm = sqrt(i_prod / bi_prod);
   free(p); free(s); free(r); free(r_old);
   return flag;
}

/* hypre_GMRESSolve (src/krylov/gmres.c:294-1100), default options (no rel_change, real-residual
 * check on convergence) */
int ho_gmres(int nrows, const int *ai, const int *aj, const double *aa, int precond_kind, ho_amg *amg,
             double tol, double a_tol, int max_iter, int k_dim, const double *b, double *x,
             int *num_iterations, double *rel_residual_norm)
{
   ho_csr A;
   const int n = nrows;
   const size_t na = (size_t) (n ? n : 1);
   double **p = (double **) malloc(sizeof(double *) * (size_t) (k_dim + 1));
   double *r = (double *) calloc(na, sizeof(double)), *w = (double *) calloc(na, sizeof(double));
   double *rs = (double *) calloc((size_t) (k_dim + 1), sizeof(double));
   double *c = (double *) calloc((size_t) k_dim, sizeof(double)), *s = (double *) calloc((size_t) k_dim, sizeof(double));
   double **hh = (double **) malloc(sizeof(double *) * (size_t) (k_dim + 1));
   double epsilon, gamma, t, r_norm, b_norm, den_norm, real_r_norm_old, real_r_norm_new;
   const double epsmac = 1.e-16;
   int i = 0, j, k, iter = 0, flag = 0;
   A.nrows = n; A.ncols = n; A.i = ai; A.j = aj; A.a = aa;
   for (i = 0; i <= k_dim; i++) { p[i] = (double *) calloc(na, sizeof(double)); hh[i] = (double *) calloc((size_t) k_dim, sizeof(double)); }
   memcpy(p[0], b, sizeof(double) * (size_t) n);
   ho_csr_matvec(n, ai, aj, aa, -1.0, x, 1.0, p[0], p[0]);
   b_norm = sqrt(ho_inner_prod(n, b, b));
   real_r_norm_old = b_norm;
   r_norm = sqrt(ho_inner_prod(n, p[0], p[0]));
   den_norm = b_norm > 0.0 ? b_norm : r_norm;
   epsilon = fmax(a_tol, tol * den_norm);
   i = 0;
   while (iter < max_iter)
   {
      rs[0] = r_norm;
      if (r_norm == 0.0) { break; }
      if (r_norm <= epsilon)
      {
         memcpy(r, b, sizeof(double) * (size_t) n);
         ho_csr_matvec(n, ai, aj, aa, -1.0, x, 1.0, r, r);
         r_norm = sqrt(ho_inner_prod(n, r, r));
         if (r_norm <= epsilon) { break; }
      }
      t = 1.0 / r_norm;
      ho_scale(n, t, p[0]);
      i = 0;
      while (i < k_dim && iter < max_iter)
      {
         i++; iter++;
         precond(precond_kind, amg, &A, p[i - 1], r);
         ho_csr_matvec(n, ai, aj, aa, 1.0, r, 0.0, p[i], p[i]);
         for (j = 0; j < i; j++)
         {
            hh[j][i - 1] = ho_inner_prod(n, p[j], p[i]);
            ho_axpy(n, -hh[j][i - 1], p[j], p[i]);
         }
         t = sqrt(ho_inner_prod(n, p[i], p[i]));
         hh[i][i - 1] = t;
         if (t != 0.0) { t = 1.0 / t; ho_scale(n, t, p[i]); }
         for (j = 1; j < i; j++)
         {
            t = hh[j - 1][i - 1];
            hh[j - 1][i - 1] = s[j - 1] * hh[j][i - 1] + c[j - 1] * t;
            hh[j][i - 1] = -s[j - 1] * t + c[j - 1] * hh[j][i - 1];
         }
         t = hh[i][i - 1] * hh[i][i - 1];
         t += hh[i - 1][i - 1] * hh[i - 1][i - 1];
         gamma = sqrt(t);
         if (gamma == 0.0) { gamma = epsmac; }
         c[i - 1] = hh[i - 1][i - 1] / gamma;
         s[i - 1] = hh[i][i - 1] / gamma;
         rs[i] = -hh[i][i - 1] * rs[i - 1];
         rs[i] /= gamma;
         rs[i - 1] = c[i - 1] * rs[i - 1];
         hh[i - 1][i - 1] = s[i - 1] * hh[i][i - 1] + c[i - 1] * hh[i - 1][i - 1];
         r_norm = fabs(rs[i]);
         if (r_norm <= epsilon) { break; }
      }
      rs[i - 1] = rs[i - 1] / hh[i - 1][i - 1];
      for (k = i - 2; k >= 0; k--)
      {
         t = 0.0;
         for (j = k + 1; j < i; j++) { t -= hh[k][j] * rs[j]; }
         t += rs[k];
         rs[k] = t / hh[k][k];
      }
      memcpy(w, p[i - 1], sizeof(double) * (size_t) n);
      ho_scale(n, rs[i - 1], w);
      for (j = i - 2; j >= 0; j--) { ho_axpy(n, rs[j], p[j], w); }
      precond(precond_kind, amg, &A, w, r);
      ho_axpy(n, 1.0, r, x);
      if (r_norm <= epsilon)
      {
         memcpy(r, b, sizeof(double) * (size_t) n);
         ho_csr_matvec(n, ai, aj, aa, -1.0, x, 1.0, r, r);
         real_r_norm_new = r_norm = sqrt(ho_inner_prod(n, r, r));
         if (r_norm <= epsilon) { break; }
         if (real_r_norm_new >= real_r_norm_old) { break; }
         memcpy(p[0], r, sizeof(double) * (size_t) n);
         i = 0;
         real_r_norm_old = real_r_norm_new;
      }
      for (j = i; j > 0; j--) { rs[j - 1] = -s[j - 1] * rs[j]; rs[j] = c[j - 1] * rs[j]; }
      if (i) { ho_axpy(n, rs[i] - 1.0, p[i], p[i]); }
      for (j = i - 1; j > 0; j--) { ho_axpy(n, rs[j], p[j], p[i]); }
      if (i) { ho_axpy(n, rs[0] - 1.0, p[0], p[0]); ho_axpy(n, 1.0, p[i], p[0]); }
   }
   *num_iterations = iter;
   *rel_residual_norm = b_norm > 0.0 ? r_norm / b_norm : r_norm;
   if (iter >= max_iter && r_norm > epsilon && epsilon > 0) { flag |= 256; }
   for (i = 0; i <= k_dim; i++) { free(p[i]); free(hh[i]); }
   free(p); free(hh); free(r); free(w); free(rs); free(c); free(s);
   return flag;
}

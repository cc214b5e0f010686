/* Test program of tests/test_shim_updates.py: the in-place update pattern of hypre applications
 * (time stepping / Newton loops) through the public HYPRE API only.  Linked once against the
 * reference alone and once in front of libHYPRE_b200.so; both print the same lines.
 *
 *   step 1  assemble a 7-point operator, BoomerAMG-PCG setup + solve
 *   step 2  HYPRE_IJMatrixInitialize + SetValues (new diagonal) + Assemble on the SAME object,
 *           setup + solve again                                  (stale-mirror check, with re-setup)
 *   step 3  HYPRE_BoomerAMGSetRelaxWt between two solves, no setup  (run-time parameter refresh)
 *   step 4  values changed in place again, diagonal-scaled PCG, setup + solve (no AMG)
 */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "HYPRE.h"
#include "_hypre_utilities.h"   /* hypre_MPI_* : the serial stubs or a real MPI, as the library was built */
#include "HYPRE_IJ_mv.h"
#include "HYPRE_parcsr_ls.h"
#include "HYPRE_krylov.h"

static int N;
static int idx(int i, int j, int k) { return (k * N + j) * N + i; }

static void fill(HYPRE_IJMatrix ij, double diag)
{
   int i, j, k;
   for (k = 0; k < N; k++) for (j = 0; j < N; j++) for (i = 0; i < N; i++)
   {
      HYPRE_BigInt row = idx(i, j, k), cols[7];
      double vals[7];
      int nnz = 0;
      cols[nnz] = row; vals[nnz++] = diag + 0.01 * (double) ((i + 2 * j + 3 * k) % 5);
      if (i > 0)     { cols[nnz] = idx(i - 1, j, k); vals[nnz++] = -1.0; }
      if (i < N - 1) { cols[nnz] = idx(i + 1, j, k); vals[nnz++] = -1.0; }
      if (j > 0)     { cols[nnz] = idx(i, j - 1, k); vals[nnz++] = -1.0; }
      if (j < N - 1) { cols[nnz] = idx(i, j + 1, k); vals[nnz++] = -1.0; }
      if (k > 0)     { cols[nnz] = idx(i, j, k - 1); vals[nnz++] = -1.0; }
      if (k < N - 1) { cols[nnz] = idx(i, j, k + 1); vals[nnz++] = -1.0; }
      HYPRE_IJMatrixSetValues(ij, 1, &nnz, &row, cols, vals);
   }
}

static void report(const char *what, HYPRE_Solver pcg, HYPRE_ParVector x)
{
   HYPRE_Int its;
   double res, xx;
   HYPRE_PCGGetNumIterations(pcg, &its);
   HYPRE_PCGGetFinalRelativeResidualNorm(pcg, &res);
   HYPRE_ParVectorInnerProd(x, x, &xx);
   printf("%s: iterations %d, final rel. residual %.6e, |x| %.12e\n", what, (int) its, res, sqrt(xx));
}

int main(int argc, char **argv)
{
   HYPRE_IJMatrix ij;
   HYPRE_ParCSRMatrix A;
   HYPRE_IJVector ib, ix;
   HYPRE_ParVector b, x;
   HYPRE_Solver pcg, amg;
   int n, r;
   hypre_MPI_Init(&argc, &argv);
   HYPRE_Initialize();
   N = argc > 1 ? atoi(argv[1]) : 14;
   n = N * N * N;
   HYPRE_IJMatrixCreate(hypre_MPI_COMM_WORLD, 0, n - 1, 0, n - 1, &ij);
   HYPRE_IJMatrixSetObjectType(ij, HYPRE_PARCSR);
   HYPRE_IJMatrixInitialize(ij);
   fill(ij, 6.0);
   HYPRE_IJMatrixAssemble(ij);
   HYPRE_IJMatrixGetObject(ij, (void **) &A);
   HYPRE_IJVectorCreate(hypre_MPI_COMM_WORLD, 0, n - 1, &ib); HYPRE_IJVectorSetObjectType(ib, HYPRE_PARCSR); HYPRE_IJVectorInitialize(ib);
   HYPRE_IJVectorCreate(hypre_MPI_COMM_WORLD, 0, n - 1, &ix); HYPRE_IJVectorSetObjectType(ix, HYPRE_PARCSR); HYPRE_IJVectorInitialize(ix);
   for (r = 0; r < n; r++) { HYPRE_BigInt row = r; double one = 1.0 + 0.001 * (r % 7), zero = 0.0; HYPRE_IJVectorSetValues(ib, 1, &row, &one); HYPRE_IJVectorSetValues(ix, 1, &row, &zero); }
   HYPRE_IJVectorAssemble(ib); HYPRE_IJVectorGetObject(ib, (void **) &b);
   HYPRE_IJVectorAssemble(ix); HYPRE_IJVectorGetObject(ix, (void **) &x);

   HYPRE_ParCSRPCGCreate(hypre_MPI_COMM_WORLD, &pcg);
   HYPRE_PCGSetTol(pcg, 1e-8); HYPRE_PCGSetMaxIter(pcg, 200); HYPRE_PCGSetTwoNorm(pcg, 1);
   HYPRE_BoomerAMGCreate(&amg);
   HYPRE_BoomerAMGSetTol(amg, 0.0); HYPRE_BoomerAMGSetMaxIter(amg, 1); HYPRE_BoomerAMGSetRelaxType(amg, 18);
   HYPRE_BoomerAMGSetPrintLevel(amg, 0);
   HYPRE_PCGSetPrecond(pcg, (HYPRE_PtrToSolverFcn) HYPRE_BoomerAMGSolve, (HYPRE_PtrToSolverFcn) HYPRE_BoomerAMGSetup, amg);

   /* step 1 */
   HYPRE_PCGSetup(pcg, (HYPRE_Matrix) A, (HYPRE_Vector) b, (HYPRE_Vector) x);
   HYPRE_PCGSolve(pcg, (HYPRE_Matrix) A, (HYPRE_Vector) b, (HYPRE_Vector) x);
   report("step 1 (first operator)", pcg, x);

   /* step 2: same IJ object, new values in place */
   HYPRE_IJMatrixInitialize(ij);
   fill(ij, 9.5);
   HYPRE_IJMatrixAssemble(ij);
   HYPRE_IJMatrixGetObject(ij, (void **) &A);
   HYPRE_ParVectorSetConstantValues(x, 0.0);
   HYPRE_PCGSetup(pcg, (HYPRE_Matrix) A, (HYPRE_Vector) b, (HYPRE_Vector) x);
   HYPRE_PCGSolve(pcg, (HYPRE_Matrix) A, (HYPRE_Vector) b, (HYPRE_Vector) x);
   report("step 2 (values updated in place, new setup)", pcg, x);

   /* step 3: a run-time parameter changed between two solves, no setup */
   HYPRE_BoomerAMGSetRelaxWt(amg, 0.6);
   HYPRE_ParVectorSetConstantValues(x, 0.0);
   HYPRE_PCGSolve(pcg, (HYPRE_Matrix) A, (HYPRE_Vector) b, (HYPRE_Vector) x);
   report("step 3 (relax weight 0.6, no setup)", pcg, x);

   /* step 4: another in-place update, diagonal scaling instead of AMG */
   {
      HYPRE_Solver pcg2;
      HYPRE_IJMatrixInitialize(ij);
      fill(ij, 7.25);
      HYPRE_IJMatrixAssemble(ij);
      HYPRE_IJMatrixGetObject(ij, (void **) &A);
      HYPRE_ParCSRPCGCreate(hypre_MPI_COMM_WORLD, &pcg2);
      HYPRE_PCGSetTol(pcg2, 1e-8); HYPRE_PCGSetMaxIter(pcg2, 500); HYPRE_PCGSetTwoNorm(pcg2, 1);
      HYPRE_PCGSetPrecond(pcg2, (HYPRE_PtrToSolverFcn) HYPRE_ParCSRDiagScale, (HYPRE_PtrToSolverFcn) HYPRE_ParCSRDiagScaleSetup, NULL);
      HYPRE_ParVectorSetConstantValues(x, 0.0);
      HYPRE_PCGSetup(pcg2, (HYPRE_Matrix) A, (HYPRE_Vector) b, (HYPRE_Vector) x);
      HYPRE_PCGSolve(pcg2, (HYPRE_Matrix) A, (HYPRE_Vector) b, (HYPRE_Vector) x);
      report("step 4 (third operator, diagonal scaling)", pcg2, x);
      HYPRE_ParCSRPCGDestroy(pcg2);
   }
   HYPRE_BoomerAMGDestroy(amg);
   HYPRE_ParCSRPCGDestroy(pcg);
   HYPRE_IJMatrixDestroy(ij); HYPRE_IJVectorDestroy(ib); HYPRE_IJVectorDestroy(ix);
   HYPRE_Finalize();
   hypre_MPI_Finalize();
   return 0;
}

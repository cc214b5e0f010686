// krylov_ext.cu — the other Krylov drivers of the ParCSR function table on the same device kernels
// (SURVEY §8 row f4): BiCGSTAB, FlexGMRES, COGMRES and LGMRES.
//
// Reference: hypre_BiCGSTABSolve (src/krylov/bicgstab.c:246-606), hypre_FlexGMRESSolve
// (src/krylov/flexgmres.c:288-812), hypre_LGMRESSolve (src/krylov/lgmres.c:320-940),
// hypre_COGMRESSolve (src/krylov/cogmres.c:270-896) with the
// batched vector operations COGMRES is built on: hypre_ParVectorMassInnerProd / MassDotpTwo / MassAxpy
// (src/parcsr_mv/par_vector_batched.c:17-135 over src/seq_mv/vector_batched.c:212-1208).
//
// As in krylov.cu the stopping tests, restarts and messages are the reference's; the scalars live in
// device slots.  What the device adds:
//  * COGMRES' one-reduce Gram-Schmidt is what it was designed for: i dots in ceil(i/4) passes over the
//    new basis vector and ONE all-reduce of i doubles, then one pass that subtracts all i projections
//    (the coefficients never leave the device between the two);
//  * BiCGSTAB: the two vector updates that follow alpha, the two that follow gamma and the three-step
//    update of p are one pass each (same rounding as the separate Axpy / ScaleVector calls), the two dots
//    that close an iteration (<r,r>, <r0,r>) are one kernel and one all-reduce: 3 host reads per iteration
//    where the reference blocks on 5 all-reduces.
#include "krylov.cuh"
#include <vector>

namespace hb {

enum { S_DUMP = 14, S_U0 = S_H0 + 51 };   // S_U0: second coefficient block of the re-orthogonalising COGMRES

// ---------------------------------------------------------------------------------------
// batched vector kernels (vector_batched.c)
// ---------------------------------------------------------------------------------------
// y += sum_j (sign * S[slot0 + j]) * X_j, added in the order j = 0 .. k-1 with one rounding per product and
// per sum: the unroll = 0 form of hypre_SeqVectorMassAxpy (vector_batched.c:236-247)
struct FMassAxpyDev {
   const double *X; size_t xstride; double *y; const double *S; int slot0, k; double sign;
   __device__ void operator()(size_t i) const
   {
      double t = y[i];
      for (int j = 0; j < k; j++) t = __dadd_rn(t, __dmul_rn(sign * S[slot0 + j], X[(size_t) j * xstride + i]));
      y[i] = t;
   }
};

// BiCGSTAB passes (bicgstab.c:497-498, 515-516, 579-591): two Axpy calls / the p update in one sweep
struct FAxpyPair {   // x += a * v ; r -= a * q
   const double *v, *q; double *x, *r; double a;
   __device__ void operator()(size_t i) const
   {
      x[i] = __dadd_rn(x[i], __dmul_rn(a, v[i]));
      r[i] = __dadd_rn(r[i], __dmul_rn(-a, q[i]));
   }
};
struct FBicgUpdateP {   // Axpy(-gamma, q, p); ScaleVector(c, p); Axpy(1, r, p)
   const double *q, *r; double *p; double gamma, c;
   __device__ void operator()(size_t i) const
   {
      const double t = __dadd_rn(p[i], __dmul_rn(-gamma, q[i]));
      p[i] = __dadd_rn((c == 1.0) ? t : __dmul_rn(t, c), r[i]);
   }
};

static int mass_dot(const double *x, const double *Z, size_t zstride, int k, size_t n, int slot0, cudaStream_t st)
{
   Ctx &c = ctx();
   timer_tick(T_BLAS1);
   for (int j0 = 0; j0 < k; j0 += kMassNV) {
      const int nv = (k - j0 < kMassNV) ? k - j0 : kMassNV;
      HB_CHECK(vec_mass_dot_dev(x, Z + (size_t) j0 * zstride, zstride, nv, n, slot0 + j0, S_DUMP, st));
   }
   timer_tick(T_OTHER);
   return scalars_allreduce(slot0, k, c.s_comp);
}

// host values into device scalar slots (through the pinned mirror; the stream is idle after the fetch
// that produced them, and is drained again before the mirror is reused)
static int scalars_store(int slot, int count, const double *vals, cudaStream_t st)
{
   Ctx &c = ctx();
   for (int k = 0; k < count; k++) c.h_scalars[slot + k] = vals[k];
   HB_CUDA(cudaMemcpyAsync(c.d_scalars + slot, c.h_scalars + slot, sizeof(double) * count, cudaMemcpyHostToDevice, st));
   HB_CUDA(cudaStreamSynchronize(st));
   return 0;
}

static void print_residual_header(bool new_style, double b_norm)
{
   // hypre_KrylovResPrintHeader (krylov_res_print.h:47-66) / the older table of cogmres.c:462-474
   printf("=============================================\n\n");
   if (new_style) {
      if (b_norm > 0.0) {
         printf("Iters      resid.norm     conv.rate   rel.res.norm\n");
         printf("-----    ------------    ----------   ------------\n");
      } else {
         printf("Iters      resid.norm     conv.rate\n");
         printf("-----    ------------    ----------\n");
      }
   } else {
      if (b_norm > 0.0) {
         printf("Iters     resid.norm     conv.rate  rel.res.norm\n");
         printf("-----    ------------    ---------- ------------\n");
      } else {
         printf("Iters     resid.norm     conv.rate\n");
         printf("-----    ------------    ----------\n");
      }
   }
}

static void print_residual_row(bool new_style, int iter, double norm, double prev, double b_norm)
{
   if (new_style) {   // hypre_KrylovResPrintScalarRow
      if (b_norm > 0.0) printf("%5d    %e      %f   %e\n", iter, norm, norm / prev, norm / b_norm);
      else printf("%5d    %e      %f\n", iter, norm, norm / prev);
   } else {
      if (b_norm > 0.0) printf("% 5d    %e    %f   %e\n", iter, norm, norm / prev, norm / b_norm);
      else printf("% 5d    %e    %f\n", iter, norm, norm / prev);
   }
}

// =======================================================================================
// BiCGSTAB
// =======================================================================================
static int bicgstab_solve_dev(hb200_parcsr *A, int pk, hb200_amg *amg, const hb200_bicgstab_params *P,
                              const double *b, double *x, double *norms, hb200_krylov_result *res)
{
   Ctx &c = ctx();
   const size_t n = (size_t) A->num_rows;
   const size_t na = n ? n : 1;
   cudaStream_t st = c.s_comp;
   const int my_id = c.rank;
   const int min_iter = P->min_iter, max_iter = P->max_iter;
   const double r_tol = P->tol, cf_tol = P->cf_tol, a_tol = P->a_tol;
   const bool log = (P->logging > 0 || P->print_level > 0) && norms;
   const bool prt = P->print_level > 0 && my_id == 0 && norms;
   ProfRange pr_solve("BiCGSTAB-Solve");

   // r, r0, s, v, p, q (hypre_BiCGSTABSetup, bicgstab.c:176-199) as one slab of the persistent workspace
   double *slab = nullptr;
   HB_CHECK(ws_get(8, sizeof(double) * na * 6, &slab));
   double *r = slab, *r0 = slab + na, *s = slab + 2 * na, *v = slab + 3 * na, *p = slab + 4 * na, *q = slab + 5 * na;
   auto cleanup = [&]() { cudaStreamSynchronize(st); };
#define BI_CHECK(expr) do { int f_ = (expr); if (f_) { cleanup(); return f_; } } while (0)

   int iter = 0, converged = 0, eflag = 0;
   double alpha, beta, gamma, epsilon, temp, res_, r_norm, b_norm, den_norm;
   const double epsmac = DBL_MIN;   // HYPRE_REAL_MIN
   double ieee_check = 0.0, cf_ave_0 = 0.0, cf_ave_1 = 0.0, weight, r_norm_0;

   // r0 = b - A x ; r = p = r0 (bicgstab.c:314-321)
   BI_CHECK(parcsr_matvec(A, -1.0, x, 1.0, b, r0));
   BI_CHECK(vec_copy(r0, r, n, st));
   BI_CHECK(vec_copy(r0, p, n, st));
   {
      double v2[2];
      BI_CHECK(vec_dot2_dev(b, b, r0, r0, n, S_T0, S_T1, st));
      BI_CHECK(scalars_allreduce(S_T0, 2, st));
      BI_CHECK(scalars_fetch(S_T0, 2, v2, st));
      b_norm = sqrt(v2[0]);
      res_ = v2[1];
   }
   if (b_norm != 0.0) ieee_check = b_norm / b_norm;
   if (ieee_check != ieee_check) {
      cleanup();
      res->error_flag = HB200_ERROR_GENERIC;
      return set_error(HB200_ERROR_GENERIC, "hb200_bicgstab_solve: INFs and/or NaNs detected in input b");
   }
   r_norm = sqrt(res_);
   r_norm_0 = r_norm;
   if (r_norm != 0.0) ieee_check = r_norm / r_norm;
   if (ieee_check != ieee_check) {
      cleanup();
      res->error_flag = HB200_ERROR_GENERIC;
      return set_error(HB200_ERROR_GENERIC, "hb200_bicgstab_solve: INFs and/or NaNs detected in A or x_0");
   }
   if (log) norms[0] = r_norm;
   if (prt) {
      printf("L2 norm of b: %e\n", b_norm);
      if (b_norm == 0.0) printf("Rel_resid_norm actually contains the residual norm\n");
      printf("Initial L2 norm of residual: %e\n", r_norm);
   }
   den_norm = b_norm > 0.0 ? b_norm : r_norm;
   if (P->stop_crit) epsilon = (a_tol == 0.0) ? r_tol : a_tol;   // bicgstab.c:407-419
   else epsilon = fmax(a_tol, r_tol * den_norm);
   if (prt) print_residual_header(false, b_norm);

   res->num_iterations = 0;
   if (b_norm > 0.0) res->rel_residual_norm = r_norm / b_norm;
   if (r_norm == 0.0) { cleanup(); return 0; }
   if (r_norm <= epsilon && iter >= min_iter) {
      if (prt) {
         printf("\n\n");
         printf("Tolerance and min_iter requirements satisfied by initial data.\n");
         printf("Final L2 norm of residual: %e\n\n", r_norm);
      }
      res->converged = 1;
      cleanup();
      return 0;
   }

   while (iter < max_iter) {
      iter++;
      // v = C p ; q = A v ; alpha = <r0,r> / <r0,q>
      BI_CHECK(precond_apply(pk, amg, A, p, v));
      BI_CHECK(parcsr_matvec(A, 1.0, v, 0.0, q, q));
      BI_CHECK(dot_global_host(r0, q, n, &temp));
      if (fabs(temp) >= epsmac) alpha = res_ / temp;
      else {
         cleanup();
         res->num_iterations = iter; res->error_flag = HB200_ERROR_GENERIC;
         return set_error(HB200_ERROR_GENERIC, "BiCGSTAB broke down!! divide by near zero");
      }
      { FAxpyPair f{v, q, x, r, alpha}; HB_EW(f, n, st); }
      // v = C r ; s = A v ; gamma = <r,s> / <s,s>
      BI_CHECK(precond_apply(pk, amg, A, r, v));
      BI_CHECK(parcsr_matvec(A, 1.0, v, 0.0, s, s));
      {
         double g2[2];
         BI_CHECK(vec_dot2_dev(r, s, s, s, n, S_T0, S_T1, st));
         BI_CHECK(scalars_allreduce(S_T0, 2, st));
         BI_CHECK(scalars_fetch(S_T0, 2, g2, st));
         gamma = (g2[0] == 0.0 && g2[1] == 0.0) ? 0.0 : g2[0] / g2[1];
      }
      { FAxpyPair f{v, s, x, r, gamma}; HB_EW(f, n, st); }
      // <r,r> for the convergence test and <r0,r> for beta, in one pass (the reference takes the second
      // one after the test, bicgstab.c:577; the value is the same)
      double res_new;
      {
         double d2[2];
         BI_CHECK(vec_dot2_dev(r, r, r0, r, n, S_T0, S_T1, st));
         BI_CHECK(scalars_allreduce(S_T0, 2, st));
         BI_CHECK(scalars_fetch(S_T0, 2, d2, st));
         r_norm = sqrt(d2[0]);
         res_new = d2[1];
      }
      if (log) norms[iter] = r_norm;
      if (prt) print_residual_row(false, iter, norms[iter], norms[iter - 1], b_norm);
      if (r_norm <= epsilon && iter >= min_iter) {
         // evaluate the actual residual (bicgstab.c:534-548)
         double rr;
         BI_CHECK(parcsr_matvec(A, -1.0, x, 1.0, b, r));
         BI_CHECK(dot_global_host(r, r, n, &rr));
         r_norm = sqrt(rr);
         if (r_norm <= epsilon) {
            if (prt) { printf("\n\n"); printf("Final L2 norm of residual: %e\n\n", r_norm); }
            converged = 1;
            break;
         }
         // r was replaced by the true residual: <r0,r> with it
         BI_CHECK(dot_global_host(r0, r, n, &res_new));
      }
      if (cf_tol > 0.0) {
         cf_ave_0 = cf_ave_1;
         cf_ave_1 = pow(r_norm / r_norm_0, 1.0 / (2.0 * (double) iter));
         weight = fabs(cf_ave_1 - cf_ave_0);
         weight = weight / fmax(cf_ave_1, cf_ave_0);
         weight = 1.0 - weight;
         if (weight * cf_ave_1 > cf_tol) break;
      }
      if (fabs(res_) >= epsmac) beta = 1.0 / res_;
      else {
         cleanup();
         res->num_iterations = iter; res->error_flag = HB200_ERROR_GENERIC;
         return set_error(HB200_ERROR_GENERIC, "BiCGSTAB broke down!! res=0");
      }
      res_ = res_new;
      beta *= res_;
      if (fabs(gamma) >= epsmac) {
         FBicgUpdateP f{q, r, p, gamma, beta * alpha / gamma};
         HB_EW(f, n, st);
      } else {
         cleanup();
         res->num_iterations = iter; res->error_flag = HB200_ERROR_GENERIC;
         return set_error(HB200_ERROR_GENERIC, "BiCGSTAB broke down!! gamma=0");
      }
   }

   res->num_iterations = iter;
   res->converged = converged;
   res->rel_residual_norm = b_norm > 0.0 ? r_norm / b_norm : r_norm;
   if (iter >= max_iter && r_norm > epsilon && epsilon > 0 && P->hybrid != -1) eflag |= HB200_ERROR_CONV;
   res->error_flag = eflag;
   cleanup();
#undef BI_CHECK
   return eflag;
}

// =======================================================================================
// FlexGMRES and COGMRES: restarted Arnoldi with the reference's two ways of keeping the basis
// =======================================================================================
enum { AV_FLEX = 0, AV_CO = 1 };

static int arnoldi_solve_dev(int variant, hb200_parcsr *A, int pk, hb200_amg *amg, const hb200_gmres_params *P,
                             const double *b, double *x, double *norms, hb200_krylov_result *res)
{
   Ctx &c = ctx();
   const bool flex = (variant == AV_FLEX);
   const size_t n = (size_t) A->num_rows;
   const size_t na = n ? n : 1;
   cudaStream_t st = c.s_comp;
   const int my_id = c.rank;
   const int k_dim = P->k_dim, min_iter = P->min_iter, max_iter = P->max_iter;
   const int rel_change = flex ? 0 : P->rel_change;          // FlexGMRES stores the flag and never reads it
   const int skip_real_r_check = flex ? 0 : P->skip_real_r_check;
   const int cgs = flex ? 1 : P->cgs;
   const double r_tol = P->tol, cf_tol = P->cf_tol, a_tol = P->a_tol;
   const bool log = (P->logging > 0 || P->print_level > 0) && norms;
   HB_REQUIRE(k_dim >= 1 && k_dim <= 100, HB200_ERROR_ARG, "k_dim out of range (1..100)");
   HB_REQUIRE(cgs <= 1 || k_dim <= 50, HB200_ERROR_ARG, "COGMRES with re-orthogonalisation: k_dim out of range (1..50)");
   int eflag = 0;
   ProfRange pr_solve(flex ? "FlexGMRES-Solve" : "COGMRES-Solve");

   // basis p[0..k_dim] as one slab; FlexGMRES keeps the preconditioned vectors pre_vecs[0..k_dim] as a
   // second one (flexgmres.c:225-245); r, w [, w_2]
   double *slab = nullptr;
   const int nvec = (k_dim + 1) * (flex ? 2 : 1) + 2 + (rel_change ? 1 : 0);
   HB_CHECK(ws_get(8, sizeof(double) * na * (size_t) nvec, &slab));
   double *basis = slab;
   double *pre = flex ? slab + (size_t) (k_dim + 1) * na : nullptr;
   double *r = slab + (size_t) (k_dim + 1) * (flex ? 2 : 1) * na;
   double *w = r + na;
   double *w_2 = rel_change ? w + na : nullptr;
   auto pv = [&](int q) { return basis + (size_t) q * na; };
   auto zv = [&](int q) { return pre + (size_t) q * na; };
   auto cleanup = [&]() { cudaStreamSynchronize(st); };
#define AR_CHECK(expr) do { int f_ = (expr); if (f_) { cleanup(); return f_; } } while (0)

   // Hessenberg columns: H(j, col) = entry j of column col (cogmres.c keeps them flat, column after column)
   const int ld = k_dim + 1;
   std::vector<double> hh((size_t) ld * k_dim, 0.0), uu((size_t) ld * k_dim, 0.0);
   std::vector<double> rs(k_dim + 1, 0.0), cc(k_dim, 0.0), ss(k_dim, 0.0), rs_2(k_dim + 1, 0.0), rv(k_dim + 1, 0.0);
   auto H = [&](int j, int col) -> double & { return hh[(size_t) col * ld + j]; };
   int i = 0, j, k, iter = 0, break_value = 0, converged = 0;
   double epsilon, gamma, t, r_norm, b_norm, den_norm, x_norm, w_norm;
   const double epsmac = 1.e-16, guard_zero_residual = 0.0;
   double ieee_check = 0.0, cf_ave_0 = 0.0, cf_ave_1 = 0.0, weight, r_norm_0, relative_error = 1.0;
   int rel_change_passed = 0, num_rel_change_check = 0;
   double real_r_norm_old, real_r_norm_new;
   const char *who = flex ? "hb200_flexgmres_solve" : "hb200_cogmres_solve";

   // p[0] = b - A x
   AR_CHECK(parcsr_matvec(A, -1.0, x, 1.0, b, pv(0)));
   double bb, rr;
   AR_CHECK(vec_dot2_dev(b, b, pv(0), pv(0), n, S_T0, S_T1, st));
   AR_CHECK(scalars_allreduce(S_T0, 2, st));
   { double v2[2]; AR_CHECK(scalars_fetch(S_T0, 2, v2, st)); bb = v2[0]; rr = v2[1]; }
   b_norm = sqrt(bb);
   real_r_norm_old = b_norm;
   if (b_norm != 0.0) ieee_check = b_norm / b_norm;
   if (ieee_check != ieee_check) {
      cleanup();
      res->error_flag = HB200_ERROR_GENERIC;
      return set_error(HB200_ERROR_GENERIC, "%s: INFs and/or NaNs detected in input b", who);
   }
   r_norm = sqrt(rr);
   r_norm_0 = r_norm;
   if (r_norm != 0.0) ieee_check = r_norm / r_norm;
   if (ieee_check != ieee_check) {
      cleanup();
      res->error_flag = HB200_ERROR_GENERIC;
      return set_error(HB200_ERROR_GENERIC, "%s: INFs and/or NaNs detected in A or x_0", who);
   }
   if (log) norms[0] = r_norm;
   if (!my_id && P->print_level > 1 && (P->logging > 0 || P->print_level > 0)) {
      printf("L2 norm of b: %e\n", b_norm);
      if (b_norm == 0.0) printf("Rel_resid_norm actually contains the residual norm\n");
      printf("Initial L2 norm of residual: %e\n", r_norm);
   }
   den_norm = b_norm > 0.0 ? b_norm : r_norm;
   epsilon = fmax(a_tol, r_tol * den_norm);
   if (!my_id && P->print_level > 1) print_residual_header(flex, b_norm);

   while (iter < max_iter) {
      rs[0] = r_norm;
      if (r_norm == 0.0) {
         cleanup();
         res->num_iterations = iter; res->converged = 0; res->error_flag = 0;
         res->rel_residual_norm = 0.0;
         return 0;
      }
      if (r_norm <= epsilon && iter >= min_iter && !rel_change) {
         AR_CHECK(parcsr_matvec(A, -1.0, x, 1.0, b, r));
         AR_CHECK(dot_global_host(r, r, n, &rr));
         r_norm = sqrt(rr);
         if (r_norm <= epsilon) {
            if (!flex && !my_id && P->print_level > 1) { printf("\n\n"); printf("Final L2 norm of residual: %e\n\n", r_norm); }
            break;
         } else if (!my_id && P->print_level > 0) printf("false convergence 1\n");
      }
      t = 1.0 / r_norm;
      AR_CHECK(vec_scale(t, pv(0), n, st));
      i = 0;
      while (i < k_dim && iter < max_iter) {
         i++;
         iter++;
         const int col = i - 1;
         // z = C p[i-1] ; p[i] = A z   (FlexGMRES keeps z, flexgmres.c:562-570)
         double *z = flex ? zv(i - 1) : r;
         AR_CHECK(precond_apply(pk, amg, A, pv(i - 1), z));
         AR_CHECK(parcsr_matvec(A, 1.0, z, 0.0, pv(i), pv(i)));
         if (flex) {
            // modified Gram-Schmidt, coefficients stay on the device (flexgmres.c:572-587)
            for (j = 0; j < i; j++) {
               AR_CHECK(dot_global(pv(j), pv(i), n, S_H0 + j));
               FAxpyDev fa{pv(j), pv(i), c.d_scalars, S_H0 + j, -1.0};
               HB_EW(fa, n, st);
            }
         } else if (cgs > 1) {
            // classical Gram-Schmidt with the low-synchronisation correction (cogmres.c:547-563)
            std::vector<double> hx(i), hy(i);
            AR_CHECK(mass_dot(pv(i), basis, na, i, n, S_H0, st));
            AR_CHECK(mass_dot(pv(i - 1), basis, na, i, n, S_U0, st));
            AR_CHECK(scalars_fetch(S_H0, i, hx.data(), st));
            AR_CHECK(scalars_fetch(S_U0, i, hy.data(), st));
            for (j = 0; j < i; j++) { H(j, col) = hx[j]; uu[(size_t) col * ld + j] = hy[j]; }
            for (j = 0; j < i - 1; j++) uu[(size_t) j * ld + i - 1] = uu[(size_t) col * ld + j];
            for (j = 0; j < i; j++) rv[j] = H(j, col);
            for (k = 0; k < i; k++)
               for (j = 0; j < i; j++) H(j, col) -= uu[(size_t) k * ld + j] * rv[j];
            for (j = 0; j < i; j++) H(j, col) = -rv[j] - H(j, col);
            for (j = 0; j < i; j++) hx[j] = H(j, col);
            AR_CHECK(scalars_store(S_H0, i, hx.data(), st));
            { FMassAxpyDev fm{basis, na, pv(i), c.d_scalars, S_H0, i, 1.0}; HB_EW(fm, n, st); }
            for (j = 0; j < i; j++) H(j, col) = -H(j, col);
         } else {
            // classical Gram-Schmidt: i dots, one reduction, one update (cogmres.c:564-577)
            AR_CHECK(mass_dot(pv(i), basis, na, i, n, S_H0, st));
            FMassAxpyDev fm{basis, na, pv(i), c.d_scalars, S_H0, i, -1.0};
            HB_EW(fm, n, st);
         }
         AR_CHECK(dot_global(pv(i), pv(i), n, S_H0 + i));
         { FScaleInvSqrtDev fs{pv(i), c.d_scalars, S_H0 + i}; HB_EW(fs, n, st); }
         {
            double hcol[kScalarSlots];
            AR_CHECK(scalars_fetch(S_H0, i + 1, hcol, st));
            if (cgs <= 1) for (j = 0; j < i; j++) H(j, col) = hcol[j];
            H(i, col) = sqrt(hcol[i]);
         }
         // Givens rotations on the new column
         for (j = 1; j < i; j++) {
            t = H(j - 1, col);
            H(j - 1, col) = ss[j - 1] * H(j, col) + cc[j - 1] * t;
            H(j, col) = -ss[j - 1] * t + cc[j - 1] * H(j, col);
         }
         t = H(i, col) * H(i, col);
         t += H(i - 1, col) * H(i - 1, col);
         gamma = sqrt(t);
         if (gamma == 0.0) gamma = epsmac;
         cc[i - 1] = H(i - 1, col) / gamma;
         ss[i - 1] = H(i, col) / gamma;
         rs[i] = -H(i, col) * rs[i - 1];
         rs[i] /= gamma;
         rs[i - 1] = cc[i - 1] * rs[i - 1];
         H(i - 1, col) = ss[i - 1] * H(i, col) + cc[i - 1] * H(i - 1, col);
         r_norm = fabs(rs[i]);
         if (P->print_level > 0) {
            if (norms) norms[iter] = r_norm;
            if (!my_id && P->print_level > 1 && norms) print_residual_row(flex, iter, norms[iter], norms[iter - 1], b_norm);
         }
         if (cf_tol > 0.0) {
            cf_ave_0 = cf_ave_1;
            cf_ave_1 = pow(r_norm / r_norm_0, 1.0 / (2.0 * (double) iter));
            weight = fabs(cf_ave_1 - cf_ave_0);
            weight = weight / fmax(cf_ave_1, cf_ave_0);
            weight = 1.0 - weight;
            if (weight * cf_ave_1 > cf_tol) { break_value = 1; break; }
         }
         if (r_norm <= epsilon && iter >= min_iter) {
            if (rel_change && !rel_change_passed) {
               // relative change of the iterate inside the restart cycle (cogmres.c:636-731)
               for (k = 0; k < i; k++) rs_2[k] = rs[k];
               rs_2[i - 1] = rs_2[i - 1] / H(i - 1, col);
               for (k = i - 2; k >= 0; k--) {
                  t = 0.0;
                  for (j = k + 1; j < i; j++) t -= H(k, j) * rs_2[j];
                  t += rs_2[k];
                  rs_2[k] = t / H(k, k);
               }
               AR_CHECK(vec_copy(pv(i - 1), w, n, st));
               AR_CHECK(vec_scale(rs_2[i - 1], w, n, st));
               for (j = i - 2; j >= 0; j--) AR_CHECK(vec_axpy(rs_2[j], pv(j), w, n, st));
               AR_CHECK(precond_apply(pk, amg, A, w, r));
               AR_CHECK(vec_copy(x, w, n, st));
               AR_CHECK(vec_axpy(1.0, r, w, n, st));
               double ww;
               AR_CHECK(dot_global_host(w, w, n, &ww));
               x_norm = sqrt(ww);
               if (!(x_norm <= guard_zero_residual)) {
                  if (num_rel_change_check) {
                     AR_CHECK(vec_copy(w, r, n, st));
                     AR_CHECK(vec_axpy(-1.0, w_2, r, n, st));
                     AR_CHECK(vec_copy(w, w_2, n, st));
                  } else {
                     AR_CHECK(vec_copy(w, w_2, n, st));
                     AR_CHECK(vec_set(w, 0.0, n, st));
                     AR_CHECK(vec_axpy(rs_2[i - 1], pv(i - 1), w, n, st));
                     AR_CHECK(precond_apply(pk, amg, A, w, r));
                  }
                  AR_CHECK(dot_global_host(r, r, n, &ww));
                  w_norm = sqrt(ww);
                  relative_error = w_norm / x_norm;
                  if (relative_error <= r_tol) { rel_change_passed = 1; break; }
               } else {
                  rel_change_passed = 1;
                  break;
               }
               num_rel_change_check++;
            } else {
               break;
            }
         }
      }   // restart cycle

      if (break_value) break;

      // back substitution, then the correction: FlexGMRES combines the stored preconditioned vectors,
      // COGMRES combines the basis and applies the preconditioner once
      rs[i - 1] = rs[i - 1] / H(i - 1, i - 1);
      for (k = i - 2; k >= 0; k--) {
         t = 0.0;
         for (j = k + 1; j < i; j++) t -= H(k, j) * rs[j];
         t += rs[k];
         rs[k] = t / H(k, k);
      }
      if (flex) {
         AR_CHECK(vec_copy(zv(i - 1), w, n, st));
         AR_CHECK(vec_scale(rs[i - 1], w, n, st));
         for (j = i - 2; j >= 0; j--) AR_CHECK(vec_axpy(rs[j], zv(j), w, n, st));
         AR_CHECK(vec_axpy(1.0, w, x, n, st));
      } else {
         AR_CHECK(vec_copy(pv(i - 1), w, n, st));
         AR_CHECK(vec_scale(rs[i - 1], w, n, st));
         for (j = i - 2; j >= 0; j--) AR_CHECK(vec_axpy(rs[j], pv(j), w, n, st));
         AR_CHECK(precond_apply(pk, amg, A, w, r));
         AR_CHECK(vec_axpy(1.0, r, x, n, st));
      }

      if (r_norm <= epsilon && iter >= min_iter) {
         if (skip_real_r_check) { converged = 1; break; }
         AR_CHECK(parcsr_matvec(A, -1.0, x, 1.0, b, r));
         AR_CHECK(dot_global_host(r, r, n, &rr));
         real_r_norm_new = r_norm = sqrt(rr);
         if (r_norm <= epsilon) {
            if (rel_change && !rel_change_passed) {
               double xx;
               AR_CHECK(dot_global_host(x, x, n, &xx));
               x_norm = sqrt(xx);
               if (!(x_norm <= guard_zero_residual)) {
                  AR_CHECK(vec_set(w, 0.0, n, st));
                  AR_CHECK(vec_axpy(rs[i - 1], pv(i - 1), w, n, st));
                  AR_CHECK(precond_apply(pk, amg, A, w, r));
                  AR_CHECK(dot_global_host(r, r, n, &xx));
                  w_norm = sqrt(xx);
                  relative_error = w_norm / x_norm;
                  if (relative_error < r_tol) { converged = 1; }
               } else { converged = 1; }
            } else { converged = 1; }
            if (converged) {
               if (!flex && !my_id && P->print_level > 1) { printf("\n\n"); printf("Final L2 norm of residual: %e\n\n", r_norm); }
               break;
            }
         } else {
            if (flex) {
               if (!my_id && P->print_level > 0) printf("false convergence 2\n");
            } else {
               // the real residual norm did not decrease: give up (cogmres.c:837-848)
               if (real_r_norm_new >= real_r_norm_old) {
                  if (!my_id && P->print_level > 1) { printf("\n\n"); printf("Final L2 norm of residual: %e\n\n", r_norm); }
                  converged = 1;
                  break;
               }
               if (!my_id && P->print_level > 0) printf("false convergence 2, L2 norm of residual: %e\n", r_norm);
            }
            AR_CHECK(vec_copy(r, pv(0), n, st));
            i = 0;
            real_r_norm_old = real_r_norm_new;
         }
      }

      // residual vector for the restart
      for (j = i; j > 0; j--) {
         rs[j - 1] = -ss[j - 1] * rs[j];
         rs[j] = cc[j - 1] * rs[j];
      }
      if (i) AR_CHECK(vec_axpy(rs[i] - 1.0, pv(i), pv(i), n, st));
      for (j = i - 1; j > 0; j--) AR_CHECK(vec_axpy(rs[j], pv(j), pv(i), n, st));
      if (i) {
         AR_CHECK(vec_axpy(rs[0] - 1.0, pv(0), pv(0), n, st));
         AR_CHECK(vec_axpy(1.0, pv(i), pv(0), n, st));
      }
   }

   if (flex && P->print_level > 0 && !my_id) { printf("Final L2 norm of residual: %e\n", r_norm); printf("\n"); }
   if (flex && P->print_level > 1 && !my_id) printf("\n\n");
   res->num_iterations = iter;
   res->converged = converged;
   res->rel_residual_norm = b_norm > 0.0 ? r_norm / b_norm : r_norm;
   if (iter >= max_iter && r_norm > epsilon && epsilon > 0) eflag |= HB200_ERROR_CONV;
   res->error_flag = eflag;
   cleanup();
#undef AR_CHECK
   return eflag;
}

// =======================================================================================
// LGMRES: restarted GMRES augmented with the error approximations of the previous cycles
// =======================================================================================
static int lgmres_solve_dev(hb200_parcsr *A, int pk, hb200_amg *amg, const hb200_gmres_params *P,
                            const double *b, double *x, double *norms, hb200_krylov_result *res)
{
   Ctx &c = ctx();
   const size_t n = (size_t) A->num_rows;
   const size_t na = n ? n : 1;
   cudaStream_t st = c.s_comp;
   const int my_id = c.rank;
   const int k_dim = P->k_dim, min_iter = P->min_iter, max_iter = P->max_iter;
   int aug_dim = P->aug_dim < 0 ? 0 : P->aug_dim;
   if (aug_dim > k_dim - 1) aug_dim = k_dim - 1;             // hypre_LGMRESSetAugDim (lgmres.c:975-990)
   if (aug_dim < 0) aug_dim = 0;
   const int approx_constant = P->approx_constant;
   const double r_tol = P->tol, cf_tol = P->cf_tol, a_tol = P->a_tol;
   const bool log = (P->logging > 0 || P->print_level > 0) && norms;
   HB_REQUIRE(k_dim >= 1 && k_dim <= 100, HB200_ERROR_ARG, "k_dim out of range (1..100)");
   int eflag = 0;
   ProfRange pr_solve("LGMRES-Solve");

   // p[0..k_dim], aug_vecs[0..aug_dim], a_aug_vecs[0..aug_dim-1] (hypre_LGMRESSetup, lgmres.c:240-275), r, w
   double *slab = nullptr;
   const int nvec = (k_dim + 1) + (aug_dim + 1) + aug_dim + 2;
   HB_CHECK(ws_get(8, sizeof(double) * na * (size_t) nvec, &slab));
   auto pv = [&](int q) { return slab + (size_t) q * na; };
   auto augv = [&](int q) { return slab + (size_t) (k_dim + 1 + q) * na; };
   auto aaugv = [&](int q) { return slab + (size_t) (k_dim + 1 + aug_dim + 1 + q) * na; };
   double *r = slab + (size_t) (k_dim + 1 + aug_dim + 1 + aug_dim) * na;
   double *w = r + na;
   auto cleanup = [&]() { cudaStreamSynchronize(st); };
#define LG_CHECK(expr) do { int f_ = (expr); if (f_) { cleanup(); return f_; } } while (0)

   const int kd = k_dim + aug_dim;
   std::vector<double> rs(kd + 1, 0.0), cc(kd, 0.0), ss(kd, 0.0);
   std::vector<std::vector<double>> hh(kd + 1, std::vector<double>(kd, 0.0));
   std::vector<int> aug_order(aug_dim > 0 ? aug_dim : 1, 0);
   int i = 0, j, k, ii, iter = 0, break_value = 0, converged = 0;
   int aug_ct = 0, it_arnoldi, it_total, it_aug, order, spot = 0;
   double epsilon, gamma, t, r_norm, b_norm, den_norm, r_norm_last = 0.0, tmp_norm;
   const double epsmac = 1.e-16;
   double ieee_check = 0.0, cf_ave_0 = 0.0, cf_ave_1 = 0.0, weight, r_norm_0;

   LG_CHECK(parcsr_matvec(A, -1.0, x, 1.0, b, pv(0)));
   double bb, rr;
   LG_CHECK(vec_dot2_dev(b, b, pv(0), pv(0), n, S_T0, S_T1, st));
   LG_CHECK(scalars_allreduce(S_T0, 2, st));
   { double v2[2]; LG_CHECK(scalars_fetch(S_T0, 2, v2, st)); bb = v2[0]; rr = v2[1]; }
   b_norm = sqrt(bb);
   if (b_norm != 0.0) ieee_check = b_norm / b_norm;
   if (ieee_check != ieee_check) {
      cleanup();
      res->error_flag = HB200_ERROR_GENERIC;
      return set_error(HB200_ERROR_GENERIC, "hb200_lgmres_solve: INFs and/or NaNs detected in input b");
   }
   r_norm = sqrt(rr);
   r_norm_0 = r_norm;
   if (r_norm != 0.0) ieee_check = r_norm / r_norm;
   if (ieee_check != ieee_check) {
      cleanup();
      res->error_flag = HB200_ERROR_GENERIC;
      return set_error(HB200_ERROR_GENERIC, "hb200_lgmres_solve: INFs and/or NaNs detected in A or x_0");
   }
   if (log) norms[0] = r_norm;
   if (!my_id && P->print_level > 1 && (P->logging > 0 || P->print_level > 0)) {
      printf("L2 norm of b: %e\n", b_norm);
      if (b_norm == 0.0) printf("Rel_resid_norm actually contains the residual norm\n");
      printf("Initial L2 norm of residual: %e\n", r_norm);
   }
   den_norm = b_norm > 0.0 ? b_norm : r_norm;
   epsilon = fmax(a_tol, r_tol * den_norm);
   if (!my_id && P->print_level > 1) print_residual_header(false, b_norm);

   while (iter < max_iter) {
      rs[0] = r_norm;
      if (r_norm == 0.0) {
         cleanup();
         res->num_iterations = iter; res->converged = 0; res->error_flag = 0;
         res->rel_residual_norm = 0.0;
         return 0;
      }
      if (r_norm <= epsilon && iter >= min_iter) {
         LG_CHECK(parcsr_matvec(A, -1.0, x, 1.0, b, r));
         LG_CHECK(dot_global_host(r, r, n, &rr));
         r_norm = sqrt(rr);
         if (r_norm <= epsilon) {
            if (!my_id && P->print_level > 1) { printf("\n\n"); printf("Final L2 norm of residual: %e\n\n", r_norm); }
            break;
         } else if (!my_id && P->print_level > 0) printf("false convergence 1\n");
      }
      t = 1.0 / r_norm;
      r_norm_last = r_norm;
      LG_CHECK(vec_scale(t, pv(0), n, st));
      i = 0;
      // the approximation space keeps its size: Krylov vectors make up for the augmentation vectors not there yet
      it_arnoldi = approx_constant ? k_dim - aug_ct : k_dim - aug_dim;
      it_total = it_arnoldi + aug_ct;
      it_aug = 0;
      while (i < it_total && iter < max_iter) {
         i++;
         iter++;
         if (i <= it_arnoldi) {
            LG_CHECK(precond_apply(pk, amg, A, pv(i - 1), r));
            LG_CHECK(parcsr_matvec(A, 1.0, r, 0.0, pv(i), pv(i)));
         } else {
            it_aug++;
            order = i - it_arnoldi - 1;
            for (ii = 0; ii < aug_dim; ii++) { if (aug_order[ii] == order) { spot = ii; break; } }
            LG_CHECK(vec_copy(aaugv(spot), pv(i), n, st));
         }
         for (j = 0; j < i; j++) {
            LG_CHECK(dot_global(pv(j), pv(i), n, S_H0 + j));
            FAxpyDev fa{pv(j), pv(i), c.d_scalars, S_H0 + j, -1.0};
            HB_EW(fa, n, st);
         }
         LG_CHECK(dot_global(pv(i), pv(i), n, S_H0 + i));
         { FScaleInvSqrtDev fs{pv(i), c.d_scalars, S_H0 + i}; HB_EW(fs, n, st); }
         {
            double hcol[kScalarSlots];
            LG_CHECK(scalars_fetch(S_H0, i + 1, hcol, st));
            for (j = 0; j < i; j++) hh[j][i - 1] = hcol[j];
            hh[i][i - 1] = sqrt(hcol[i]);
         }
         for (j = 1; j < i; j++) {
            t = hh[j - 1][i - 1];
            hh[j - 1][i - 1] = ss[j - 1] * hh[j][i - 1] + cc[j - 1] * t;
            hh[j][i - 1] = -ss[j - 1] * t + cc[j - 1] * hh[j][i - 1];
         }
         t = hh[i][i - 1] * hh[i][i - 1];
         t += hh[i - 1][i - 1] * hh[i - 1][i - 1];
         gamma = sqrt(t);
         if (gamma == 0.0) gamma = epsmac;
         cc[i - 1] = hh[i - 1][i - 1] / gamma;
         ss[i - 1] = hh[i][i - 1] / gamma;
         rs[i] = -hh[i][i - 1] * rs[i - 1];
         rs[i] /= gamma;
         rs[i - 1] = cc[i - 1] * rs[i - 1];
         hh[i - 1][i - 1] = ss[i - 1] * hh[i][i - 1] + cc[i - 1] * hh[i - 1][i - 1];
         r_norm = fabs(rs[i]);
         if (P->print_level > 0) {
            if (norms) norms[iter] = r_norm;
            if (!my_id && P->print_level > 1 && norms) print_residual_row(false, iter, norms[iter], norms[iter - 1], b_norm);
         }
         if (cf_tol > 0.0) {
            cf_ave_0 = cf_ave_1;
            cf_ave_1 = pow(r_norm / r_norm_0, 1.0 / (2.0 * (double) iter));
            weight = fabs(cf_ave_1 - cf_ave_0);
            weight = weight / fmax(cf_ave_1, cf_ave_0);
            weight = 1.0 - weight;
            if (weight * cf_ave_1 > cf_tol) { break_value = 1; break; }
         }
         if (r_norm <= epsilon && iter >= min_iter) break;
      }   // restart cycle

      if (break_value) break;

      rs[i - 1] = rs[i - 1] / hh[i - 1][i - 1];
      for (k = i - 2; k >= 0; k--) {
         t = 0.0;
         for (j = k + 1; j < i; j++) t -= hh[k][j] * rs[j];
         t += rs[k];
         rs[k] = t / hh[k][k];
      }
      if (it_arnoldi > i) it_arnoldi = i;
      if (!it_aug) {
         LG_CHECK(vec_copy(pv(i - 1), w, n, st));
         LG_CHECK(vec_scale(rs[i - 1], w, n, st));
         for (j = i - 2; j >= 0; j--) LG_CHECK(vec_axpy(rs[j], pv(j), w, n, st));
      } else {
         // the correction holds Krylov vectors and augmentation vectors (lgmres.c:782-806)
         LG_CHECK(vec_copy(pv(0), w, n, st));
         LG_CHECK(vec_scale(rs[0], w, n, st));
         for (j = 1; j < it_arnoldi; j++) LG_CHECK(vec_axpy(rs[j], pv(j), w, n, st));
         for (ii = 0; ii < it_aug; ii++) {
            for (j = 0; j < aug_dim; j++) { if (aug_order[j] == ii) { spot = j; break; } }
            LG_CHECK(vec_axpy(rs[it_arnoldi + ii], augv(spot), w, n, st));
         }
      }
      // w is the error approximation of this cycle: kept as the next augmentation vector
      LG_CHECK(vec_copy(w, augv(aug_dim), n, st));
      LG_CHECK(precond_apply(pk, amg, A, w, r));
      LG_CHECK(vec_axpy(1.0, r, x, n, st));

      if (r_norm <= epsilon && iter >= min_iter) {
         LG_CHECK(parcsr_matvec(A, -1.0, x, 1.0, b, r));
         LG_CHECK(dot_global_host(r, r, n, &rr));
         r_norm = sqrt(rr);
         if (r_norm <= epsilon) {
            if (!my_id && P->print_level > 1) { printf("\n\n"); printf("Final L2 norm of residual: %e\n\n", r_norm); }
            converged = 1;
            break;
         }
         if (!my_id && P->print_level > 0) printf("false convergence 2\n");
         LG_CHECK(vec_copy(r, pv(0), n, st));
         i = 0;
      }

      // w = r_0 of this cycle (lgmres.c:844-845), then the residual vector for the restart
      LG_CHECK(vec_copy(pv(0), w, n, st));
      LG_CHECK(vec_scale(r_norm_last, w, n, st));
      for (j = i; j > 0; j--) {
         rs[j - 1] = -ss[j - 1] * rs[j];
         rs[j] = cc[j - 1] * rs[j];
      }
      if (i) LG_CHECK(vec_axpy(rs[i] - 1.0, pv(i), pv(i), n, st));
      for (j = i - 1; j > 0; j--) LG_CHECK(vec_axpy(rs[j], pv(j), pv(i), n, st));
      if (i) {
         LG_CHECK(vec_axpy(rs[0] - 1.0, pv(0), pv(0), n, st));
         LG_CHECK(vec_axpy(1.0, pv(i), pv(0), n, st));
      }

      // the new augmentation vector and A times it (= (r_0 - r_m) / |z|: independent of the preconditioner)
      if (aug_dim > 0) {
         if (!aug_ct) { spot = 0; aug_ct++; }
         else if (aug_ct < aug_dim) { spot = aug_ct; aug_ct++; }
         else { for (ii = 0; ii < aug_dim; ii++) if (aug_order[ii] == aug_dim - 1) spot = ii; }
         LG_CHECK(vec_copy(augv(aug_dim), augv(spot), n, st));
         double zz;
         LG_CHECK(dot_global_host(augv(spot), augv(spot), n, &zz));
         tmp_norm = 1.0 / sqrt(zz);
         LG_CHECK(vec_scale(tmp_norm, augv(spot), n, st));
         for (ii = 0; ii < aug_dim; ii++) aug_order[ii]++;
         aug_order[spot] = 0;
         LG_CHECK(vec_copy(w, aaugv(spot), n, st));
         LG_CHECK(vec_scale(-1.0, aaugv(spot), n, st));
         LG_CHECK(vec_axpy(1.0, pv(0), aaugv(spot), n, st));
         LG_CHECK(vec_scale(-tmp_norm, aaugv(spot), n, st));
      }
   }

   if (!my_id && P->print_level > 1) printf("\n\n");
   res->num_iterations = iter;
   res->converged = converged;
   res->rel_residual_norm = b_norm > 0.0 ? r_norm / b_norm : r_norm;
   if (iter >= max_iter && r_norm > epsilon && epsilon > 0) eflag |= HB200_ERROR_CONV;
   res->error_flag = eflag;
   cleanup();
#undef LG_CHECK
   return eflag;
}

}  // namespace hb

using namespace hb;

static int check_precond_ext(int kind, hb200_amg *amg)
{
   HB_REQUIRE(kind >= 0 && kind <= 2, HB200_ERROR_ARG, "unknown preconditioner kind");
   HB_REQUIRE(kind != HB200_PRECOND_AMG || amg != nullptr, HB200_ERROR_ARG, "AMG preconditioner requested but amg is NULL");
   return 0;
}

template <class Fn>
static int timed_solve(hb200_krylov_result *result, Fn &&solve)
{
   Ctx &c = ctx();
   memset(result, 0, sizeof(*result));
   const long long l0 = c.launches;
   HB_CUDA(cudaEventRecord(c.ev_c, c.s_comp));
   const int f = solve();
   HB_CUDA(cudaEventRecord(c.ev_d, c.s_comp));
   HB_CUDA(cudaEventSynchronize(c.ev_d));
   float ms = 0.f;
   cudaEventElapsedTime(&ms, c.ev_c, c.ev_d);
   result->solve_ms = ms;
   result->kernel_launches = c.launches - l0;
   return f;
}

// host-buffer entry points (what HYPRE_BiCGSTABSolve / HYPRE_FlexGMRESSolve / HYPRE_COGMRESSolve see from a
// CPU application): H2D of b and x0, solve, D2H of x
template <class Fn>
static int host_solve(hb200_parcsr *A, const double *b_host, double *x_host, Fn &&solve)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A && ((b_host && x_host) || A->num_rows == 0), HB200_ERROR_ARG, "null argument");
   Ctx &c = ctx();
   const size_t n = (size_t) A->num_rows, na = n ? n : 1;
   double *db = nullptr, *dx = nullptr;
   HB_CHECK(ws_get(6, sizeof(double) * na, &db));
   HB_CHECK(ws_get(7, sizeof(double) * na, &dx));
   HB_CUDA(cudaMemcpyAsync(db, b_host, sizeof(double) * n, cudaMemcpyHostToDevice, c.s_comp));
   HB_CUDA(cudaMemcpyAsync(dx, x_host, sizeof(double) * n, cudaMemcpyHostToDevice, c.s_comp));
   const int f = solve(db, dx);
   cudaMemcpyAsync(x_host, dx, sizeof(double) * n, cudaMemcpyDeviceToHost, c.s_comp);
   cudaStreamSynchronize(c.s_comp);
   return f;
}

extern "C" {

void hb200_bicgstab_default_params(hb200_bicgstab_params *p)
{
   memset(p, 0, sizeof(*p));
   p->tol = 1.0e-06; p->max_iter = 1000;   // bicgstab.c:76-86
}

int hb200_bicgstab_solve(hb200_parcsr *A, int precond_kind, hb200_amg *amg, const hb200_bicgstab_params *params,
                         const double *b, double *x, double *norms, hb200_krylov_result *result)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A && params && result && ((b && x) || A->num_rows == 0), HB200_ERROR_ARG, "null argument");
   HB_CHECK(check_precond_ext(precond_kind, amg));
   return timed_solve(result, [&]() { return bicgstab_solve_dev(A, precond_kind, amg, params, b, x, norms, result); });
}

int hb200_flexgmres_solve(hb200_parcsr *A, int precond_kind, hb200_amg *amg, const hb200_gmres_params *params,
                          const double *b, double *x, double *norms, hb200_krylov_result *result)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A && params && result && ((b && x) || A->num_rows == 0), HB200_ERROR_ARG, "null argument");
   HB_CHECK(check_precond_ext(precond_kind, amg));
   return timed_solve(result, [&]() { return arnoldi_solve_dev(AV_FLEX, A, precond_kind, amg, params, b, x, norms, result); });
}

int hb200_cogmres_solve(hb200_parcsr *A, int precond_kind, hb200_amg *amg, const hb200_gmres_params *params,
                        const double *b, double *x, double *norms, hb200_krylov_result *result)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A && params && result && ((b && x) || A->num_rows == 0), HB200_ERROR_ARG, "null argument");
   HB_CHECK(check_precond_ext(precond_kind, amg));
   return timed_solve(result, [&]() { return arnoldi_solve_dev(AV_CO, A, precond_kind, amg, params, b, x, norms, result); });
}

int hb200_lgmres_solve(hb200_parcsr *A, int precond_kind, hb200_amg *amg, const hb200_gmres_params *params,
                       const double *b, double *x, double *norms, hb200_krylov_result *result)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A && params && result && ((b && x) || A->num_rows == 0), HB200_ERROR_ARG, "null argument");
   HB_CHECK(check_precond_ext(precond_kind, amg));
   return timed_solve(result, [&]() { return lgmres_solve_dev(A, precond_kind, amg, params, b, x, norms, result); });
}

int hb200_bicgstab_solve_host(hb200_parcsr *A, int precond_kind, hb200_amg *amg, const hb200_bicgstab_params *params,
                              const double *b_host, double *x_host, double *norms, hb200_krylov_result *result)
{
   return host_solve(A, b_host, x_host, [&](double *db, double *dx) {
      return hb200_bicgstab_solve(A, precond_kind, amg, params, db, dx, norms, result); });
}

int hb200_flexgmres_solve_host(hb200_parcsr *A, int precond_kind, hb200_amg *amg, const hb200_gmres_params *params,
                               const double *b_host, double *x_host, double *norms, hb200_krylov_result *result)
{
   return host_solve(A, b_host, x_host, [&](double *db, double *dx) {
      return hb200_flexgmres_solve(A, precond_kind, amg, params, db, dx, norms, result); });
}

int hb200_cogmres_solve_host(hb200_parcsr *A, int precond_kind, hb200_amg *amg, const hb200_gmres_params *params,
                             const double *b_host, double *x_host, double *norms, hb200_krylov_result *result)
{
   return host_solve(A, b_host, x_host, [&](double *db, double *dx) {
      return hb200_cogmres_solve(A, precond_kind, amg, params, db, dx, norms, result); });
}

int hb200_lgmres_solve_host(hb200_parcsr *A, int precond_kind, hb200_amg *amg, const hb200_gmres_params *params,
                            const double *b_host, double *x_host, double *norms, hb200_krylov_result *result)
{
   return host_solve(A, b_host, x_host, [&](double *db, double *dx) {
      return hb200_lgmres_solve(A, precond_kind, amg, params, db, dx, norms, result); });
}

}  // extern "C"

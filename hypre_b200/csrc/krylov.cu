// krylov.cu — PCG and restarted GMRES drivers over device-resident ParCSR data.
//
// Reference: hypre_PCGSolve (src/krylov/pcg.c:313-1016) and hypre_GMRESSolve
// (src/krylov/gmres.c:294-1100) driven through the ParCSR function tables of
// HYPRE_ParCSRPCGCreate / HYPRE_ParCSRGMRESCreate (src/parcsr_ls/HYPRE_parcsr_pcg.c:15-38,
// HYPRE_parcsr_gmres.c:15-46; adaptors src/parcsr_ls/par_krylov_func.c:40-317).
//
// The control flow, stopping tests, breakdown tests and logging are the reference's.  What
// changes is where the scalars live: dot products land in device slots, alpha/beta are
// formed on the device by the kernels that consume them, and the host reads ONE block of
// scalars per iteration for the convergence / breakdown decisions (the reference blocks on
// MPI_Allreduce three times per iteration).
#include "krylov.cuh"

namespace hb {

// =======================================================================================
// PCG
// =======================================================================================
static int pcg_solve_dev(hb200_parcsr *A, int pk, hb200_amg *amg, const hb200_pcg_params *P,
                         const double *b, double *x, double *norms, double *rel_norms,
                         hb200_krylov_result *res)
{
   Ctx &c = ctx();
   const size_t n = (size_t) A->num_rows;
   const size_t na = n ? n : 1;
   cudaStream_t st = c.s_comp;
   const int my_id = c.rank;
   const bool log = (P->logging > 0 || P->print_level > 0) && norms;
   int eflag = 0;
   ProfRange pr_solve("PCG-Solve");
   HB_TRACE("pcg_solve: %zu local rows, halo mode %d", n, c.halo_mode);

   // work vectors p, s, r [, r_old, v] (hypre_PCGSetup, pcg.c:233-250) from the persistent workspace
   double *p = nullptr, *s = nullptr, *r = nullptr, *r_old = nullptr, *v = nullptr;
   HB_CHECK(ws_get(0, sizeof(double) * na, &p));
   HB_CHECK(ws_get(1, sizeof(double) * na, &s));
   HB_CHECK(ws_get(2, sizeof(double) * na, &r));
   if (P->flex) HB_CHECK(ws_get(3, sizeof(double) * na, &r_old));
   if (P->rtol > 0.0 && !P->two_norm && P->recompute_residual_p) HB_CHECK(ws_get(4, sizeof(double) * na, &v));
   auto cleanup = [&]() { cudaStreamSynchronize(st); };
#define PCG_CHECK(expr) do { int f_ = (expr); if (f_) { cleanup(); return f_; } } while (0)

   const double r_tol = P->tol, a_tol = P->a_tol, atolf = P->atolf, cf_tol = P->cf_tol, rtol = P->rtol;
   const int max_iter = P->max_iter, two_norm = P->two_norm, rel_change = P->rel_change;
   const int skip_break = P->skip_break, flex = P->flex, stop_crit = P->stop_crit;

   double alpha = 0.0, beta, delta = 0.0, gamma, gamma_old, bi_prod, eps;
   double i_prod = 0.0, i_prod_0 = 0.0, cf_ave_0 = 0.0, cf_ave_1 = 0.0, weight, ieee_check = 0.0;
   const double guard_zero_residual = 0.0;
   int tentatively_converged = 0, converged = 0;
   int i = 0;
   int blk_cur = S_BLK0, blk_old = S_BLK1;   // blk_cur + B_GAMMA holds the current gamma

   // ---- bi_prod (pcg.c:403-421)
   if (two_norm) {
      PCG_CHECK(dot_global_host(b, b, n, &bi_prod));
      if (P->print_level > 1 && my_id == 0) printf("<b,b>: %e\n", bi_prod);
   } else {
      PCG_CHECK(precond_apply(pk, amg, A, b, p));
      PCG_CHECK(dot_global_host(p, b, n, &bi_prod));
      if (P->print_level > 1 && my_id == 0) printf("<C*b,b>: %e\n", bi_prod);
   }
   if (bi_prod != 0.0) ieee_check = bi_prod / bi_prod;
   if (ieee_check != ieee_check) {
      // pcg.c:426-450
      if (P->print_level > 0 || P->logging > 0) {
         printf("\n\nERROR detected by Hypre ...  BEGIN\n");
         printf("ERROR -- hypre_PCGSolve: INFs and/or NaNs detected in input.\n");
         printf("User probably placed non-numerics in supplied b.\n");
         printf("Returning error flag += 101.  Program not terminated.\n");
         printf("ERROR detected by Hypre ...  END\n\n\n");
      }
      cleanup();
      res->error_flag = HB200_ERROR_GENERIC;
      return set_error(HB200_ERROR_GENERIC, "hb200_pcg_solve: INFs and/or NaNs detected in input b");
   }
   eps = r_tol * r_tol;
   if (bi_prod > 0.0) {
      if (stop_crit && !rel_change && atolf <= 0) { eps = eps / bi_prod; }
      else if (atolf > 0) { bi_prod += atolf; }
      else { eps = fmax(r_tol * r_tol, a_tol * a_tol / bi_prod); }
   } else {
      // zero right-hand side: x = b (= 0) and return (pcg.c:482-497)
      PCG_CHECK(vec_copy(b, x, n, st));
      if (log) { norms[0] = 0.0; if (rel_norms) rel_norms[0] = 0.0; }
      cleanup();
      res->num_iterations = 0; res->rel_residual_norm = 0.0; res->converged = 0;
      return 0;
   }

   // ---- r = b - A x ; p = C r ; gamma = <r,p> (pcg.c:499-510)
   PCG_CHECK(parcsr_matvec(A, -1.0, x, 1.0, b, r));
   PCG_CHECK(precond_apply(pk, amg, A, r, p));
   PCG_CHECK(vec_dot2_dev(r, p, r, r, n, blk_cur + B_GAMMA, blk_cur + B_RR, st));
   {
      PCG_CHECK(scalars_allreduce(blk_cur, 2, st));
      double tmp[S_NFETCH];
      PCG_CHECK(scalars_fetch(0, S_NFETCH, tmp, st));
      gamma = tmp[blk_cur + B_GAMMA];
      if (two_norm) i_prod_0 = tmp[blk_cur + B_RR]; else i_prod_0 = gamma;
   }
   if (gamma != 0.0) ieee_check = gamma / gamma;
   if (ieee_check != ieee_check) {
      if (P->print_level > 0 || P->logging > 0) {
         printf("\n\nERROR detected by Hypre ...  BEGIN\n");
         printf("ERROR -- hypre_PCGSolve: INFs and/or NaNs detected in input.\n");
         printf("User probably placed non-numerics in supplied A or x_0.\n");
         printf("Returning error flag += 101.  Program not terminated.\n");
         printf("ERROR detected by Hypre ...  END\n\n\n");
      }
      cleanup();
      res->error_flag = HB200_ERROR_GENERIC;
      return set_error(HB200_ERROR_GENERIC, "hb200_pcg_solve: INFs and/or NaNs detected in A or x_0");
   }
   if (log) norms[0] = sqrt(i_prod_0);
   if (P->print_level > 1 && my_id == 0) {
      printf("\n\n");
      if (two_norm) {
         if (stop_crit && !rel_change && atolf == 0) {
            printf("Iters       ||r||_2     conv.rate\n");
            printf("-----    ------------   ---------\n");
         } else {
            printf("Iters       ||r||_2     conv.rate  ||r||_2/||b||_2\n");
            printf("-----    ------------   ---------  ------------ \n");
         }
      } else {
         printf("Iters       ||r||_C     conv.rate  ||r||_C/||b||_C\n");
         printf("-----    ------------    ---------  ------------ \n");
      }
   }

   double prev_norm = sqrt(i_prod_0);
   while ((i + 1) <= max_iter) {
      i++;
      const bool recompute_true_residual = P->recompute_residual_p && !(i % P->recompute_residual_p);

      // s = A p ; sdotp = <s,p>  (the dot comes out of the matvec epilogue when the format fuses it)
      if (fused_dots_enabled() && c.nranks == 1) { c.dot_req_armed = true; c.dot_req_w = p; c.dot_req_slot = S_SDOTP; }
      PCG_CHECK(parcsr_matvec(A, 1.0, p, 0.0, s, s));
      if (!c.last_dot_fused) PCG_CHECK(dot_global(s, p, n, S_SDOTP));

      double sdotp = 0.0;
      int dflag = 0;
      const int blk_new = blk_old;          // this iteration's {gamma, rr, delta}
      const int g_cur = blk_cur + B_GAMMA, g_new = blk_new + B_GAMMA;
      const int s_rr = blk_new + B_RR, s_delta = blk_new + B_DELTA;
      if (!recompute_true_residual) {
         if (flex) PCG_CHECK(vec_copy(r, r_old, n, st));
         // x += alpha p ; r -= alpha s ; <r,r>   with alpha, and its breakdown tests, on the device
         timer_tick(T_BLAS1);
         PCG_CHECK(pcg_update_xr(p, s, x, r, n, g_cur, S_SDOTP, s_rr, S_FLAG, skip_break, st));
         timer_tick(T_OTHER);
      } else {
         // rare path (pcg.c:653-702): host-driven
         double tmp[S_NFETCH];
         PCG_CHECK(scalars_fetch(0, S_NFETCH, tmp, st));
         sdotp = tmp[S_SDOTP];
         if (sdotp == 0.0) { eflag |= HB200_ERROR_CONV; if (i == 1) i_prod = i_prod_0; break; }
         alpha = tmp[g_cur] / sdotp;
         if (alpha <= 0.0) { eflag |= HB200_ERROR_CONV; if (skip_break < 3) { if (i == 1) i_prod = i_prod_0; break; } }
         else if (!(alpha >= 4.9406564584124654e-324)) { eflag |= HB200_ERROR_CONV; if (skip_break < 2) { if (i == 1) i_prod = i_prod_0; break; } }
         else if (!(alpha >= DBL_MIN)) { eflag |= HB200_ERROR_CONV; if (skip_break < 1) { if (i == 1) i_prod = i_prod_0; break; } }
         PCG_CHECK(vec_axpy(alpha, p, x, n, st));
         if (P->print_level > 1 && my_id == 0) printf("Recomputing the residual...\n");
         PCG_CHECK(vec_copy(r, s, n, st));
         if (flex) PCG_CHECK(vec_copy(r, r_old, n, st));
         PCG_CHECK(parcsr_matvec(A, -1.0, x, 1.0, b, r));
         if (rtol > 0.0) {
            PCG_CHECK(vec_axpy(-1.0, r, s, n, st));   // s = r_old - r_new
            if (two_norm) {
               double ss;
               PCG_CHECK(dot_global_host(s, s, n, &ss));
               const double drob2 = ss / bi_prod;
               if (drob2 < rtol * rtol) {
                  if (P->print_level > 1 && my_id == 0) printf("\n\n||r_old-r_new||/||b||: %e\n", sqrt(drob2));
                  break;
               }
            } else {
               double sv;
               PCG_CHECK(precond_apply(pk, amg, A, s, v));
               PCG_CHECK(dot_global_host(s, v, n, &sv));
               const double r2ob2 = sv / bi_prod;
               if (r2ob2 < rtol * rtol) {
                  if (P->print_level > 1 && my_id == 0) printf("\n\n||r_old-r_new||_C/||b||_C: %e\n", sqrt(r2ob2));
                  break;
               }
            }
         }
         PCG_CHECK(vec_dot_dev(r, r, n, s_rr, st));
      }

      if (rtol > 0.0 && two_norm && !recompute_true_residual) {
         // pcg.c:705-719 (needs alpha on the host)
         double ss, tmp[S_NFETCH];
         PCG_CHECK(dot_global_host(s, s, n, &ss));
         PCG_CHECK(scalars_fetch(0, S_NFETCH, tmp, st));
         const double al = tmp[S_ALPHA];
         const double drob2 = al * al * ss / bi_prod;
         if (tmp[S_FLAG] <= 0.0 && drob2 < rtol * rtol) {
            if (P->print_level > 1 && my_id == 0) printf("\n\n||r_old-r_new||/||b||: %e\n", sqrt(drob2));
            break;
         }
      }

      // s = C r ; gamma = <r,s>
      bool gamma_done = false;
      PCG_CHECK(precond_apply(pk, amg, A, r, s, g_new, &gamma_done));
      timer_tick(T_BLAS1);
      if (!gamma_done) PCG_CHECK(vec_dot_dev(r, s, n, g_new, st));
      if (flex) PCG_CHECK(vec_dot_dev(r_old, s, n, s_delta, st));
      timer_tick(T_OTHER);
      // one all-reduce for <r,s>, <r,r> [, <r_old,s>] (the reference: one MPI_Allreduce per dot)
      PCG_CHECK(scalars_allreduce(blk_new, flex ? 3 : 2, st));

      // ---- the one host read of this iteration
      double S[S_NFETCH];
      PCG_CHECK(scalars_fetch(0, S_NFETCH, S, st));
      if (!recompute_true_residual) {
         sdotp = S[S_SDOTP];
         dflag = (int) S[S_FLAG];
         alpha = S[S_ALPHA];
         // breakdown tests in the reference's order (pcg.c:588-636); the device kernel has
         // already refused to touch x and r when the test demands a break
         if (dflag == 1) {
            eflag |= HB200_ERROR_CONV;
            set_error(HB200_ERROR_CONV, "Zero sdotp value in PCG");
            if (i == 1) i_prod = i_prod_0;
            break;
         }
         if (dflag != 0) {
            eflag |= HB200_ERROR_CONV;
            const int kind = dflag > 0 ? dflag : -dflag;
            set_error(HB200_ERROR_CONV, kind == 2 ? "Negative or zero alpha value in PCG"
                                      : kind == 3 ? "alpha value less than TRUE_MIN in PCG"
                                                  : "Subnormal alpha value in PCG");
            if (P->print_level > 1 && my_id == 0) printf("alpha %e", alpha);
            if (dflag > 0) { if (i == 1) i_prod = i_prod_0; break; }
         }
      }
      gamma_old = S[g_cur];
      gamma = S[g_new];
      if (flex) delta = gamma - S[s_delta];

      if (rtol > 0.0 && !two_norm && !recompute_true_residual) {
         const double r2ob2 = (gamma + gamma_old) / bi_prod;
         if (r2ob2 < rtol * rtol) {
            if (P->print_level > 1 && my_id == 0) printf("\n\n||r_old-r_new||_C/||b||_C: %e\n", sqrt(r2ob2));
            break;
         }
      }

      i_prod = two_norm ? S[s_rr] : gamma;

      if (log) {
         norms[i] = sqrt(i_prod);
         if (rel_norms) rel_norms[i] = bi_prod > 0.0 ? sqrt(i_prod / bi_prod) : 0.0;
      }
      if (P->print_level > 1 && my_id == 0) {
         const double ni = sqrt(i_prod);
         if (two_norm && stop_crit && !rel_change && atolf == 0) printf("% 5d    %e    %f\n", i, ni, ni / prev_norm);
         else printf("% 5d    %e    %f    %e\n", i, ni, ni / prev_norm, bi_prod > 0.0 ? sqrt(i_prod / bi_prod) : 0.0);
      }
      prev_norm = sqrt(i_prod);

      // ---- convergence (pcg.c:805-860)
      if (i_prod / bi_prod < eps) tentatively_converged = 1;
      if (tentatively_converged && P->recompute_residual) {
         PCG_CHECK(parcsr_matvec(A, -1.0, x, 1.0, b, r));
         if (two_norm) {
            PCG_CHECK(dot_global_host(r, r, n, &i_prod));
         } else {
            PCG_CHECK(precond_apply(pk, amg, A, r, s));
            PCG_CHECK(dot_global(r, s, n, g_new));
            PCG_CHECK(scalars_fetch(g_new, 1, &i_prod, st));
            gamma = i_prod;
         }
         if (i_prod / bi_prod >= eps) tentatively_converged = 0;
      }
      if (tentatively_converged && rel_change && (i_prod > guard_zero_residual)) {
         double pi_prod, xi_prod;
         PCG_CHECK(dot_global_host(p, p, n, &pi_prod));
         PCG_CHECK(dot_global_host(x, x, n, &xi_prod));
         const double ratio = alpha * alpha * pi_prod / xi_prod;
         if (ratio >= eps) tentatively_converged = 0;
      }
      if (tentatively_converged) { converged = 1; break; }

      // ---- gamma breakdown tests (pcg.c:862-902)
      if (gamma <= 0.0) {
         eflag |= HB200_ERROR_CONV;
         set_error(HB200_ERROR_CONV, "Negative or zero gamma value in PCG");
         if (P->print_level > 1 && my_id == 0) printf("gamma %e", gamma);
         if (skip_break < 3) { if (i == 1) i_prod = i_prod_0; break; }
      } else if (!(gamma >= 4.9406564584124654e-324)) {
         eflag |= HB200_ERROR_CONV;
         set_error(HB200_ERROR_CONV, "gamma value less than TRUE_MIN in PCG");
         if (skip_break < 2) { if (i == 1) i_prod = i_prod_0; break; }
      } else if (!(gamma >= DBL_MIN)) {
         eflag |= HB200_ERROR_CONV;
         set_error(HB200_ERROR_CONV, "Subnormal gamma value in PCG");
         if (skip_break < 1) { if (i == 1) i_prod = i_prod_0; break; }
      }

      // ---- convergence-factor test (pcg.c:912-962)
      if (cf_tol > 0.0) {
         cf_ave_0 = cf_ave_1;
         if (i_prod_0 <= 0.0) { eflag |= HB200_ERROR_CONV; if (skip_break < 3) break; }
         else if (!(i_prod_0 >= 4.9406564584124654e-324)) { eflag |= HB200_ERROR_CONV; if (skip_break < 2) break; }
         else if (!(i_prod_0 >= DBL_MIN)) { eflag |= HB200_ERROR_CONV; if (skip_break < 1) break; }
         cf_ave_1 = pow(i_prod / i_prod_0, 1.0 / (2.0 * (double) i));
         weight = fabs(cf_ave_1 - cf_ave_0);
         weight = weight / fmax(cf_ave_1, cf_ave_0);
         weight = 1.0 - weight;
         if (weight * cf_ave_1 > cf_tol) break;
      }

      // ---- p = s + beta p (pcg.c:968-984)
      if (!recompute_true_residual) {
         if (!flex) {
            timer_tick(T_BLAS1);
            PCG_CHECK(pcg_update_p(s, p, n, g_new, g_cur, st));
            timer_tick(T_OTHER);
         } else {
            beta = delta / gamma_old;
            PCG_CHECK(vec_scale(beta, p, n, st));
            PCG_CHECK(vec_axpy(1.0, s, p, n, st));
         }
      } else {
         PCG_CHECK(vec_copy(s, p, n, st));
      }
      // rotate the scalar blocks
      blk_old = blk_cur;
      blk_cur = blk_new;
   }

   if (P->print_level > 1 && my_id == 0) printf("\n\n");
   if (i >= max_iter && (i_prod / bi_prod) >= eps && eps > 0 && P->hybrid != -1) {
      eflag |= HB200_ERROR_CONV;
      set_error(HB200_ERROR_CONV, "Reached max iterations %d in PCG before convergence", max_iter);
   }
   res->num_iterations = i;
   res->converged = converged;
   res->rel_residual_norm = bi_prod > 0.0 ? sqrt(i_prod / bi_prod) : 0.0;
   res->error_flag = eflag;
   cleanup();
   HB_TRACE("pcg_solve: %d iterations, rel.res %.6e", i, res->rel_residual_norm);
#undef PCG_CHECK
   return eflag;
}

// =======================================================================================
// GMRES
// =======================================================================================
static int gmres_solve_dev(hb200_parcsr *A, int pk, hb200_amg *amg, const hb200_gmres_params *P,
                           const double *b, double *x, double *norms, hb200_krylov_result *res)
{
   Ctx &c = ctx();
   const size_t n = (size_t) A->num_rows;
   const size_t na = n ? n : 1;
   cudaStream_t st = c.s_comp;
   const int my_id = c.rank;
   const int k_dim = P->k_dim, min_iter = P->min_iter, max_iter = P->max_iter;
   const int rel_change = P->rel_change, skip_real_r_check = P->skip_real_r_check;
   const double r_tol = P->tol, cf_tol = P->cf_tol, a_tol = P->a_tol;
   const bool log = (P->logging > 0 || P->print_level > 0) && norms;
   HB_REQUIRE(k_dim >= 1 && k_dim <= 100, HB200_ERROR_ARG, "k_dim out of range (1..100)");
   int eflag = 0;
   ProfRange pr_solve("GMRES-Solve");

   // basis p[0..k_dim] as one slab (par_krylov_func.c:66-104), plus r, w [, w_2]
   double *slab = nullptr;
   const int nvec = k_dim + 1 + 2 + (rel_change ? 1 : 0);
   HB_CHECK(ws_get(5, sizeof(double) * na * (size_t) nvec, &slab));
   std::vector<double *> p(k_dim + 1);
   for (int q = 0; q <= k_dim; q++) p[q] = slab + (size_t) q * na;
   double *r = slab + (size_t) (k_dim + 1) * na;
   double *w = r + na;
   double *w_2 = rel_change ? w + na : nullptr;
   auto cleanup = [&]() { cudaStreamSynchronize(st); };
#define GM_CHECK(expr) do { int f_ = (expr); if (f_) { cleanup(); return f_; } } while (0)

   std::vector<double> rs(k_dim + 1, 0.0), cc(k_dim, 0.0), ss(k_dim, 0.0), rs_2(k_dim + 1, 0.0);
   std::vector<std::vector<double>> hh(k_dim + 1, std::vector<double>(k_dim, 0.0));
   int i = 0, j, k, iter = 0, break_value = 0, converged = 0;
   double epsilon, gamma, t, r_norm, b_norm, den_norm, x_norm, w_norm;
   const double epsmac = 1.e-16, guard_zero_residual = 0.0;
   double ieee_check = 0.0, cf_ave_0 = 0.0, cf_ave_1 = 0.0, weight, r_norm_0, relative_error = 1.0;
   int rel_change_passed = 0, num_rel_change_check = 0;
   double real_r_norm_old, real_r_norm_new;

   // p[0] = b - A x (gmres.c:385-388)
   GM_CHECK(parcsr_matvec(A, -1.0, x, 1.0, b, p[0]));
   double bb, rr;
   GM_CHECK(vec_dot2_dev(b, b, p[0], p[0], n, S_T0, S_T1, st));
   GM_CHECK(scalars_allreduce(S_T0, 2, st));
   { double v2[2]; GM_CHECK(scalars_fetch(S_T0, 2, v2, st)); bb = v2[0]; rr = v2[1]; }
   b_norm = sqrt(bb);
   real_r_norm_old = b_norm;
   if (b_norm != 0.0) ieee_check = b_norm / b_norm;
   if (ieee_check != ieee_check) {
      cleanup();
      res->error_flag = HB200_ERROR_GENERIC;
      return set_error(HB200_ERROR_GENERIC, "hb200_gmres_solve: INFs and/or NaNs detected in input b");
   }
   r_norm = sqrt(rr);
   r_norm_0 = r_norm;
   if (r_norm != 0.0) ieee_check = r_norm / r_norm;
   if (ieee_check != ieee_check) {
      cleanup();
      res->error_flag = HB200_ERROR_GENERIC;
      return set_error(HB200_ERROR_GENERIC, "hb200_gmres_solve: INFs and/or NaNs detected in A or x_0");
   }
   if (log) norms[0] = r_norm;
   if (!my_id && P->print_level > 0) {
      printf("L2 norm of b: %e\n", b_norm);
      if (b_norm == 0.0) printf("Rel_resid_norm actually contains the residual norm\n");
      printf("Initial L2 norm of residual: %e\n", r_norm);
   }
   den_norm = b_norm > 0.0 ? b_norm : r_norm;
   epsilon = fmax(a_tol, r_tol * den_norm);
   if (P->print_level > 1 && my_id == 0) {
      // hypre_KrylovResPrintHeader, scalar-residual mode
      if (b_norm > 0.0) {
         printf("=============================================\n\n");
         printf("Iters     resid.norm     conv.rate  rel.res.norm\n");
         printf("-----    ------------    ---------- ------------\n");
      } else {
         printf("=============================================\n\n");
         printf("Iters     resid.norm     conv.rate\n");
         printf("-----    ------------    ----------\n");
      }
   }

   while (iter < max_iter) {
      rs[0] = r_norm;
      if (r_norm == 0.0) {
         cleanup();
         res->num_iterations = iter; res->converged = 0; res->error_flag = 0;
         res->rel_residual_norm = 0.0;
         return 0;
      }
      if (r_norm <= epsilon && iter >= min_iter) {
         if (!rel_change) {
            GM_CHECK(parcsr_matvec(A, -1.0, x, 1.0, b, r));
            GM_CHECK(dot_global_host(r, r, n, &rr));
            r_norm = sqrt(rr);
            if (r_norm <= epsilon) break;
            else if (!my_id && P->print_level > 0) printf("false convergence 1\n");
         }
      }
      t = 1.0 / r_norm;
      GM_CHECK(vec_scale(t, p[0], n, st));
      i = 0;
      while (i < k_dim && iter < max_iter) {
         i++;
         iter++;
         // r = C p[i-1] ; p[i] = A r
         GM_CHECK(precond_apply(pk, amg, A, p[i - 1], r));
         GM_CHECK(parcsr_matvec(A, 1.0, r, 0.0, p[i], p[i]));
         // modified Gram-Schmidt: coefficients stay on the device, read back once
         for (j = 0; j < i; j++) {
            GM_CHECK(dot_global(p[j], p[i], n, S_H0 + j));
            FAxpyDev fa{p[j], p[i], c.d_scalars, S_H0 + j, -1.0};
            HB_EW(fa, n, st);
         }
         GM_CHECK(dot_global(p[i], p[i], n, S_H0 + i));
         { FScaleInvSqrtDev fs{p[i], c.d_scalars, S_H0 + i}; HB_EW(fs, n, st); }
         {
            double hcol[kScalarSlots];
            GM_CHECK(scalars_fetch(S_H0, i + 1, hcol, st));
            for (j = 0; j < i; j++) hh[j][i - 1] = hcol[j];
            hh[i][i - 1] = sqrt(hcol[i]);
         }
         // Givens rotations (gmres.c:638-659)
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
         if (P->print_level > 0) {   // gmres.c:662-664: norms[iter] is only kept when printing
            if (norms) norms[iter] = r_norm;
            if (!my_id && P->print_level > 1 && norms) {
               if (b_norm > 0.0) printf("% 5d    %e    %f   %e\n", iter, norms[iter], norms[iter] / norms[iter - 1], norms[iter] / b_norm);
               else printf("% 5d    %e    %f\n", iter, norms[iter], norms[iter] / norms[iter - 1]);
            }
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
               for (k = 0; k < i; k++) rs_2[k] = rs[k];
               rs_2[i - 1] = rs_2[i - 1] / hh[i - 1][i - 1];
               for (k = i - 2; k >= 0; k--) {
                  t = 0.0;
                  for (j = k + 1; j < i; j++) t -= hh[k][j] * rs_2[j];
                  t += rs_2[k];
                  rs_2[k] = t / hh[k][k];
               }
               GM_CHECK(vec_copy(p[i - 1], w, n, st));
               GM_CHECK(vec_scale(rs_2[i - 1], w, n, st));
               for (j = i - 2; j >= 0; j--) GM_CHECK(vec_axpy(rs_2[j], p[j], w, n, st));
               GM_CHECK(precond_apply(pk, amg, A, w, r));
               GM_CHECK(vec_copy(x, w, n, st));
               GM_CHECK(vec_axpy(1.0, r, w, n, st));
               double ww;
               GM_CHECK(dot_global_host(w, w, n, &ww));
               x_norm = sqrt(ww);
               if (!(x_norm <= guard_zero_residual)) {
                  if (num_rel_change_check) {
                     GM_CHECK(vec_copy(w, r, n, st));
                     GM_CHECK(vec_axpy(-1.0, w_2, r, n, st));
                     GM_CHECK(vec_copy(w, w_2, n, st));
                  } else {
                     GM_CHECK(vec_copy(w, w_2, n, st));
                     GM_CHECK(vec_set(w, 0.0, n, st));
                     GM_CHECK(vec_axpy(rs_2[i - 1], p[i - 1], w, n, st));
                     GM_CHECK(precond_apply(pk, amg, A, w, r));
                  }
                  GM_CHECK(dot_global_host(r, r, n, &ww));
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

      // solve the upper triangular system, update x (gmres.c:890-917)
      rs[i - 1] = rs[i - 1] / hh[i - 1][i - 1];
      for (k = i - 2; k >= 0; k--) {
         t = 0.0;
         for (j = k + 1; j < i; j++) t -= hh[k][j] * rs[j];
         t += rs[k];
         rs[k] = t / hh[k][k];
      }
      GM_CHECK(vec_copy(p[i - 1], w, n, st));
      GM_CHECK(vec_scale(rs[i - 1], w, n, st));
      for (j = i - 2; j >= 0; j--) GM_CHECK(vec_axpy(rs[j], p[j], w, n, st));
      GM_CHECK(precond_apply(pk, amg, A, w, r));
      GM_CHECK(vec_axpy(1.0, r, x, n, st));

      if (r_norm <= epsilon && iter >= min_iter) {
         if (skip_real_r_check) { converged = 1; break; }
         GM_CHECK(parcsr_matvec(A, -1.0, x, 1.0, b, r));
         GM_CHECK(dot_global_host(r, r, n, &rr));
         real_r_norm_new = r_norm = sqrt(rr);
         if (r_norm <= epsilon) {
            if (rel_change && !rel_change_passed) {
               double xx;
               GM_CHECK(dot_global_host(x, x, n, &xx));
               x_norm = sqrt(xx);
               if (!(x_norm <= guard_zero_residual)) {
                  GM_CHECK(vec_set(w, 0.0, n, st));
                  GM_CHECK(vec_axpy(rs[i - 1], p[i - 1], w, n, st));
                  GM_CHECK(precond_apply(pk, amg, A, w, r));
                  GM_CHECK(dot_global_host(r, r, n, &xx));
                  w_norm = sqrt(xx);
                  relative_error = w_norm / x_norm;
                  if (relative_error < r_tol) { converged = 1; break; }
               } else { converged = 1; break; }
            } else { converged = 1; break; }
         } else {
            if (real_r_norm_new >= real_r_norm_old) { converged = 1; break; }
            if (!my_id && P->print_level > 0) printf("false convergence 2, L2 norm of residual: %e\n", r_norm);
            GM_CHECK(vec_copy(r, p[0], n, st));
            i = 0;
            real_r_norm_old = real_r_norm_new;
         }
      }

      // residual vector for the restart (gmres.c:1002-1020)
      for (j = i; j > 0; j--) {
         rs[j - 1] = -ss[j - 1] * rs[j];
         rs[j] = cc[j - 1] * rs[j];
      }
      if (i) GM_CHECK(vec_axpy(rs[i] - 1.0, p[i], p[i], n, st));
      for (j = i - 1; j > 0; j--) GM_CHECK(vec_axpy(rs[j], p[j], p[i], n, st));
      if (i) {
         GM_CHECK(vec_axpy(rs[0] - 1.0, p[0], p[0], n, st));
         GM_CHECK(vec_axpy(1.0, p[i], p[0], n, st));
      }
   }

   if (!my_id && P->print_level > 1) printf("\n\n");
   if (!my_id && P->print_level > 0) printf("Final L2 norm of residual: %e\n\n", r_norm);
   res->num_iterations = iter;
   res->converged = converged;
   res->rel_residual_norm = b_norm > 0.0 ? r_norm / b_norm : r_norm;
   if (iter >= max_iter && r_norm > epsilon && epsilon > 0 && P->hybrid != -1) eflag |= HB200_ERROR_CONV;
   res->error_flag = eflag;
   cleanup();
#undef GM_CHECK
   return eflag;
}

}  // namespace hb

using namespace hb;

extern "C" {

void hb200_pcg_default_params(hb200_pcg_params *p)
{
   memset(p, 0, sizeof(*p));
   p->tol = 1.0e-06; p->max_iter = 1000;   // pcg.c:97-113
}

void hb200_gmres_default_params(hb200_gmres_params *p)
{
   memset(p, 0, sizeof(*p));
   p->k_dim = 5; p->tol = 1.0e-06; p->max_iter = 1000;   // gmres.c:60-75
   p->cgs = 1;                                            // cogmres.c:97
   p->aug_dim = 2; p->approx_constant = 1;                // lgmres.c:111-112
}

static int check_precond(int kind, hb200_amg *amg, hb200_parcsr *A)
{
   HB_REQUIRE(kind >= 0 && kind <= 2, HB200_ERROR_ARG, "unknown preconditioner kind");
   if (kind == HB200_PRECOND_AMG) {
      HB_REQUIRE(amg != nullptr, HB200_ERROR_ARG, "AMG preconditioner requested but amg is NULL");
      (void) A;
   }
   return 0;
}

int hb200_pcg_solve(hb200_parcsr *A, int precond_kind, hb200_amg *amg, const hb200_pcg_params *params,
                    const double *b, double *x, double *norms, double *rel_norms,
                    hb200_krylov_result *result)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A && params && result && ((b && x) || A->num_rows == 0), HB200_ERROR_ARG, "null argument");
   HB_CHECK(check_precond(precond_kind, amg, A));
   Ctx &c = ctx();
   memset(result, 0, sizeof(*result));
   const long long l0 = c.launches;
   timers_begin();
   HB_CUDA(cudaEventRecord(c.ev_c, c.s_comp));
   int f = pcg_solve_dev(A, precond_kind, amg, params, b, x, norms, rel_norms, result);
   timers_report("hb200_pcg_solve");
   HB_CUDA(cudaEventRecord(c.ev_d, c.s_comp));
   HB_CUDA(cudaEventSynchronize(c.ev_d));
   float ms = 0.f;
   cudaEventElapsedTime(&ms, c.ev_c, c.ev_d);
   result->solve_ms = ms;
   result->kernel_launches = c.launches - l0;
   return f;
}

int hb200_gmres_solve(hb200_parcsr *A, int precond_kind, hb200_amg *amg, const hb200_gmres_params *params,
                      const double *b, double *x, double *norms, hb200_krylov_result *result)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A && params && result && ((b && x) || A->num_rows == 0), HB200_ERROR_ARG, "null argument");
   HB_CHECK(check_precond(precond_kind, amg, A));
   Ctx &c = ctx();
   memset(result, 0, sizeof(*result));
   const long long l0 = c.launches;
   HB_CUDA(cudaEventRecord(c.ev_c, c.s_comp));
   int f = gmres_solve_dev(A, precond_kind, amg, params, b, x, norms, result);
   HB_CUDA(cudaEventRecord(c.ev_d, c.s_comp));
   HB_CUDA(cudaEventSynchronize(c.ev_d));
   float ms = 0.f;
   cudaEventElapsedTime(&ms, c.ev_c, c.ev_d);
   result->solve_ms = ms;
   result->kernel_launches = c.launches - l0;
   return f;
}

// Setup-time warm-up of a Krylov solve (hypre_PCGSetup / hypre_GMRESSetup in the reference allocate
// the work vectors, pcg.c:233-250): allocates the persistent workspace the host entry points use, and
// runs a few iterations on b = 1, x = 0 in that workspace, twice, so that every lazily built piece of
// the solve — halo plans, Chebyshev / GS scratch, the captured V-cycle graphs of exactly the (f, u)
// pairs the real solve will present — exists before the application starts its timer around Solve.
int hb200_krylov_warmup(hb200_parcsr *A, int precond_kind, hb200_amg *amg, int is_gmres, int k_dim)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A != nullptr, HB200_ERROR_ARG, "null matrix");
   HB_CHECK(check_precond(precond_kind, amg, A));
   Ctx &c = ctx();
   const size_t n = (size_t) A->num_rows, na = n ? n : 1;
   double *db = nullptr, *dx = nullptr;
   HB_CHECK(ws_get(6, sizeof(double) * na, &db));
   HB_CHECK(ws_get(7, sizeof(double) * na, &dx));
   hb200_krylov_result R;
   for (int pass = 0; pass < 2; pass++) {
      HB_CHECK(vec_set(db, 1.0, n, c.s_comp));
      HB_CHECK(vec_set(dx, 0.0, n, c.s_comp));
      int f;
      if (is_gmres == 4) {
         hb200_bicgstab_params P;
         hb200_bicgstab_default_params(&P);
         P.tol = 0.0; P.max_iter = 2;
         f = hb200_bicgstab_solve(A, precond_kind, amg, &P, db, dx, nullptr, &R);
      } else if (is_gmres) {
         hb200_gmres_params P;
         hb200_gmres_default_params(&P);
         P.k_dim = k_dim > 0 ? k_dim : 5; P.tol = 0.0; P.max_iter = P.k_dim + 1; P.skip_real_r_check = 1;
         if (is_gmres == 2) f = hb200_flexgmres_solve(A, precond_kind, amg, &P, db, dx, nullptr, &R);
         else if (is_gmres == 3) f = hb200_cogmres_solve(A, precond_kind, amg, &P, db, dx, nullptr, &R);
         else if (is_gmres == 5) f = hb200_lgmres_solve(A, precond_kind, amg, &P, db, dx, nullptr, &R);
         else f = hb200_gmres_solve(A, precond_kind, amg, &P, db, dx, nullptr, &R);
      } else {
         hb200_pcg_params P;
         hb200_pcg_default_params(&P);
         P.tol = 0.0; P.max_iter = 3; P.two_norm = 1;
         f = hb200_pcg_solve(A, precond_kind, amg, &P, db, dx, nullptr, nullptr, &R);
      }
      if (f & ~HB200_ERROR_CONV) return f;   // "did not converge in 3 iterations" is expected here
   }
   HB_CUDA(cudaStreamSynchronize(c.s_comp));
   return 0;
}

// host-buffer entry points: what HYPRE_PCGSolve / HYPRE_GMRESSolve see from a CPU application
static int host_wrap(hb200_parcsr *A, const double *b_host, double *x_host, double **db, double **dx)
{
   Ctx &c = ctx();
   const size_t n = (size_t) A->num_rows, na = n ? n : 1;
   HB_CHECK(ws_get(6, sizeof(double) * na, db));
   HB_CHECK(ws_get(7, sizeof(double) * na, dx));
   HB_CUDA(cudaMemcpyAsync(*db, b_host, sizeof(double) * n, cudaMemcpyHostToDevice, c.s_comp));
   HB_CUDA(cudaMemcpyAsync(*dx, x_host, sizeof(double) * n, cudaMemcpyHostToDevice, c.s_comp));
   return 0;
}

int hb200_pcg_solve_host(hb200_parcsr *A, int precond_kind, hb200_amg *amg,
                         const hb200_pcg_params *params, const double *b_host, double *x_host,
                         double *norms, double *rel_norms, hb200_krylov_result *result)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A && ((b_host && x_host) || A->num_rows == 0), HB200_ERROR_ARG, "null argument");
   double *db = nullptr, *dx = nullptr;
   HB_CHECK(host_wrap(A, b_host, x_host, &db, &dx));
   int f = hb200_pcg_solve(A, precond_kind, amg, params, db, dx, norms, rel_norms, result);
   cudaMemcpyAsync(x_host, dx, sizeof(double) * (size_t) A->num_rows, cudaMemcpyDeviceToHost, ctx().s_comp);
   cudaStreamSynchronize(ctx().s_comp);
   return f;
}

int hb200_gmres_solve_host(hb200_parcsr *A, int precond_kind, hb200_amg *amg,
                           const hb200_gmres_params *params, const double *b_host, double *x_host,
                           double *norms, hb200_krylov_result *result)
{
   HB_CHECK(require_ready());
   HB_REQUIRE(A && ((b_host && x_host) || A->num_rows == 0), HB200_ERROR_ARG, "null argument");
   double *db = nullptr, *dx = nullptr;
   HB_CHECK(host_wrap(A, b_host, x_host, &db, &dx));
   int f = hb200_gmres_solve(A, precond_kind, amg, params, db, dx, norms, result);
   cudaMemcpyAsync(x_host, dx, sizeof(double) * (size_t) A->num_rows, cudaMemcpyDeviceToHost, ctx().s_comp);
   cudaStreamSynchronize(ctx().s_comp);
   return f;
}

}  // extern "C"
